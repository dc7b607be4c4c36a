"""B200: end-to-end step_host (H2D actions, 1 step, D2H qpos, sync) vs pipeline form (call by call / CUDA graph), number of slices
and flies per block.  python tools/e2e_sweep.py [--quick]"""
import json, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table

model = NMFModel.bench(True)
n = 4096
tab = cpg_table(model, n, 200)
act = torch.from_numpy(np.ascontiguousarray(tab.transpose(1, 0, 2))).pin_memory()
res = torch.empty((n, 73), dtype=torch.float32).pin_memory()
cases = [(0, p, f) for p in (1, 4) for f in (0,)] + [(1, p, f) for p in (2, 4, 6, 8, 12, 16) for f in (0, 4, 1)]
for graph, parts, fpb in cases:
    os.environ["NMF_HOST_GRAPH"] = str(graph); os.environ["NMF_HOST_PARTS"] = str(parts)
    sim = B200Simulation(model, n_worlds=n, outputs=False)
    sim.set_flies_per_block(fpb)
    sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
    sim.warmup()
    for s in range(5):
        sim.step_host(act[s].numpy(), 1, res.numpy())
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for s in range(200):
            sim.step_host(act[s].numpy(), 1, res.numpy())
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"graph": graph, "parts": parts, "fpb": fpb, "e2e_env_steps_per_s": n * 200 / best, "us_per_step": best / 200 * 1e6}), flush=True)
    del sim
