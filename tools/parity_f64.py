"""fp64 vs fp32 instantiation of the step kernel against the fp64 oracle after 1000 CPG walking steps (one launch), per world."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table
from oracle.oracle import Oracle

T = 1000
out = {}
worlds = {"flat_capsule": (NMFModel.bench(True), 8), "flat_mesh": (NMFModel.bench(False), 4), "terrain_blocks": (NMFModel.bench(True, terrain="blocks"), 4),
          "terrain_gapped": (NMFModel.bench(True, terrain="gapped"), 4), "tethered": (NMFModel.tethered(), 4),
          "legs_active_only": (NMFModel.bench(True, joint_preset="legs_active_only"), 4)}
for name, (m, n) in worlds.items():
    a = dict(m.arrays); opt = a["opt"].copy(); opt[5] = 1e-16; a["opt"] = opt     # converge the oracle's Newton loop fully
    mt = NMFModel(a, m.names, m.meta)
    tab = cpg_table(m, n, T)
    q0 = np.tile(m.arrays["key_qpos"], (n, 1))
    if name != "tethered":
        q0[:, 2] = -0.17; q0[:, 0] += np.linspace(0, 1.0, n); q0[:, 1] += np.linspace(0, 0.6, n)
    q0 = q0.astype(np.float32)
    res = {}
    for bits in (64, 32):
        sim = B200Simulation(m, n_worlds=n, outputs=False); sim.set_precision(bits)
        sim.qpos.copy_(torch.from_numpy(q0)); sim.ctrl[:, 42:] = 1.0
        sim.step(T, torch.from_numpy(tab).cuda(), 0)
        res[bits] = (sim.qpos.cpu().numpy().astype(np.float64), sim.qvel.cpu().numpy().astype(np.float64))
    errs = {64: [], 32: []}
    for k in range(n):
        o = Oracle(mt); o.reset(); o.qpos[:] = q0[k].astype(np.float64); o.ctrl[42:] = 1.0
        o.step_table(tab[k].astype(np.float64))
        for bits in (64, 32):
            errs[bits].append({"qpos_rel_linf": float(np.abs(res[bits][0][k] - o.qpos).max() / np.abs(o.qpos).max()),
                               "qvel_abs_linf": float(np.abs(res[bits][1][k] - o.qvel).max()), "qvel_max": float(np.abs(o.qvel).max())})
    out[name] = {"steps": T, "fp64": errs[64], "fp32": errs[32]}
    print(f"{name:18s} fp64 max {max(e['qpos_rel_linf'] for e in errs[64]):.1e}   fp32 " + " ".join("%.0e" % e["qpos_rel_linf"] for e in errs[32]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/parity_f64.json").write_text(json.dumps(out, indent=1))
