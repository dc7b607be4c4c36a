"""How fast do two trajectories of the SAME walking fly separate?  (B200; writes profiles-ready JSON)

  (a) float32 kernel vs float64 kernel (same source, same inputs)            -> what the product path loses against the oracle-grade build
  (b) float64 kernel vs float64 kernel started 1 ulp(float32) away in qpos    -> what ANY float32-resolution difference grows into

If (b) grows at the rate of (a), the float32 error is amplification of rounding by the contact dynamics (chaos), not an
algorithmic difference: no float32 implementation -- MuJoCo-Warp included -- can shadow a float64 trajectory beyond that horizon.
Error = max over qpos entries of |difference| / max |qpos|, per fly; percentiles over 4096 CPG flies."""
import argparse, json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--mesh", action="store_true")
ap.add_argument("--out", default="gpurun_out/divergence.json")
args = ap.parse_args()
CHECK = (10, 30, 100, 200, 300, 500, 700, 1000)
model = NMFModel.bench(simplify_geom=not args.mesh)
n = args.n
table = torch.from_numpy(cpg_table(model, n, max(CHECK))).cuda()


def run(precision, perturb=False):
    sim = B200Simulation(model, n_worlds=n, outputs=False)
    sim.set_precision(precision)
    sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
    sim.warmup()                                     # 500 steps at the neutral pose: every fly stands on its tarsi
    if perturb:                                      # one float32 ulp on every hinge angle (what storing the state in float32 costs)
        q = sim.qpos[:, 7:]
        sim.qpos[:, 7:] = torch.nextafter(q, torch.full_like(q, float("inf")))
    out, done = {}, 0
    for cp in CHECK:
        sim.step(cp - done, table, done); done = cp
        out[cp] = sim.qpos.cpu().numpy().astype(np.float64)
    return out


a64 = run(64)
a32 = run(32)
b64 = run(64, perturb=True)
res = {"n_flies": n, "geoms": "mesh" if args.mesh else "capsule", "percentiles": [50, 90, 99, 100], "checkpoints": list(CHECK), "f32_vs_f64": {}, "f64_vs_f64_perturbed_1ulp_f32": {}}
for cp in CHECK:
    ref = a64[cp]; scale = np.abs(ref).max(axis=1)
    for key, other in (("f32_vs_f64", a32[cp]), ("f64_vs_f64_perturbed_1ulp_f32", b64[cp])):
        e = np.abs(other - ref).max(axis=1) / scale
        res[key][cp] = [float(np.percentile(e, p)) for p in (50, 90, 99, 100)] + [float((e > 1e-4).mean())]
    print(cp, "f32-f64", ["%.1e" % x for x in res["f32_vs_f64"][cp]], "| f64-f64'", ["%.1e" % x for x in res["f64_vs_f64_perturbed_1ulp_f32"][cp]], flush=True)
res["columns"] = "p50, p90, p99, max of the per-fly qpos rel Linf; last = fraction of flies above 1e-4"
Path(args.out).parent.mkdir(exist_ok=True)
Path(args.out).write_text(json.dumps(res, indent=1))
