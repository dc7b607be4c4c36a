"""B200 A/B: flies per block (1 / 2 / 4) of the f32 step kernels on the bench workload; checks that the records are bit-identical
and prints throughput per setting and launch length.  python tools/fpb_sweep.py [--terrain blocks] [--mesh]"""
import argparse, json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--chunks", type=int, nargs="+", default=[100, 20])
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--fpb", type=int, nargs="+", default=[1, 2, 4])
ap.add_argument("--mesh", action="store_true")
ap.add_argument("--terrain", default=None)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
model = NMFModel.bench(simplify_geom=not args.mesh, terrain=args.terrain)
n, T = args.n, 2500
table = torch.from_numpy(cpg_table(model, n, T)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
final = {}
for fpb in args.fpb:
    sim = B200Simulation(model, n_worlds=n, outputs=False)
    sim.set_flies_per_block(fpb)
    sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
    sim.warmup()
    for chunk in args.chunks:
        best = 0.0
        for rep in range(args.reps):
            t0, done, ms = 0, 0, 0.0
            while done < args.steps:
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); sim.step(chunk, table, t0); b.record()
                torch.cuda.synchronize()
                ms += a.elapsed_time(b); t0 = (t0 + chunk) % T; done += chunk
            best = max(best, n * done / ms * 1e3)
        print(json.dumps({"fpb": fpb, "chunk": chunk, "env_steps_per_s": best}), flush=True)
    final[fpb] = sim.state.clone()
    del sim
ks = sorted(final)
for k in ks[1:]:
    same = torch.equal(final[ks[0]].view(torch.int32), final[k].view(torch.int32))
    print(json.dumps({"bit_identical": [ks[0], k], "ok": bool(same)}), flush=True)
