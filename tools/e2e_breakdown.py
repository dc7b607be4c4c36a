"""B200: where the time of one end-to-end step (nmf_step_host: H2D actions, 1 step, D2H qpos, sync) goes.
Prints the wall time per call next to (a) the device time of a bare 1-step launch, (b) the two copies alone, (c) an empty
synchronised call, so that the pipeline's overhead can be read off."""
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
dev_tab = torch.from_numpy(tab).cuda()
sim = B200Simulation(model, n_worlds=n, outputs=False)
sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
sim.warmup()
out = {}
for s in range(5):
    sim.step_host(act[s].numpy(), 1, res.numpy())
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(200):
    sim.step_host(act[s].numpy(), 1, res.numpy())
out["step_host_us"] = (time.perf_counter() - t0) / 200 * 1e6
for fpb in (0, 8, 4):
    sim.set_flies_per_block(fpb)
    ev = []
    for s in range(100):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sim.step(1, dev_tab, s); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    out[f"one_step_launch_device_us_fpb{fpb}"] = float(np.median([a.elapsed_time(b) for a, b in ev]) * 1e3)
    t0 = time.perf_counter()
    for s in range(100):
        sim.step(1, dev_tab, s); torch.cuda.synchronize()
    out[f"one_step_launch_sync_wall_us_fpb{fpb}"] = (time.perf_counter() - t0) / 100 * 1e6
sim.set_flies_per_block(0)
d_act = torch.empty((n, 42), dtype=torch.float32, device="cuda"); d_q = torch.empty((n, 73), dtype=torch.float32, device="cuda")
t0 = time.perf_counter()
for s in range(200):
    d_act.copy_(act[s], non_blocking=True); res.copy_(d_q, non_blocking=True); torch.cuda.synchronize()
out["copies_sync_wall_us"] = (time.perf_counter() - t0) / 200 * 1e6
t0 = time.perf_counter()
for s in range(200):
    torch.cuda.synchronize()
out["empty_sync_us"] = (time.perf_counter() - t0) / 200 * 1e6
for parts in (1, 2, 4):
    os.environ["NMF_HOST_PARTS"] = str(parts)
    s2 = B200Simulation(model, n_worlds=n, outputs=False)
    s2.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32)); s2.warmup()
    for s in range(5):
        s2.step_host(act[s].numpy(), 1, res.numpy())
    t0 = time.perf_counter()
    for s in range(200):
        s2.step_host(act[s].numpy(), 1, res.numpy())
    out[f"step_host_us_parts{parts}"] = (time.perf_counter() - t0) / 200 * 1e6
    del s2
print(json.dumps(out, indent=1))
