"""B200: end-to-end step_host (H2D actions, 1 step, D2H qpos, sync) vs number of pipeline slices / flies per block."""
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
for parts in (1, 2, 4):
    os.environ["NMF_HOST_PARTS"] = str(parts)
    for fpb in (0, 8, 4):
        sim = B200Simulation(model, n_worlds=n, outputs=False)
        sim.set_flies_per_block(fpb)
        sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
        sim.warmup()
        for s in range(5):
            sim.step_host(act[s].numpy(), 1, res.numpy())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(200):
            sim.step_host(act[s].numpy(), 1, res.numpy())
        dt = time.perf_counter() - t0
        print(json.dumps({"parts": parts, "fpb": fpb, "e2e_env_steps_per_s": n * 200 / dt, "us_per_step": dt / 200 * 1e6}), flush=True)
        del sim
