"""B200 sweep of the work-queue item length (steps per item) for a given launch length; python tools/queue_sweep.py --chunk 20"""
import argparse, json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--chunk", type=int, default=20)
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--fpb", type=int, nargs="+", default=[8, 4])
ap.add_argument("--subs", type=int, nargs="+", default=[0, 3, 4, 5, 7, 10, -1])
args = ap.parse_args()
model = NMFModel.bench(True)
n, T = args.n, 2500
table = torch.from_numpy(cpg_table(model, n, T)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sim = B200Simulation(model, n_worlds=n, outputs=False)
sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
sim.warmup()
for fpb in args.fpb:
    sim.set_flies_per_block(fpb)
    for sub in args.subs:
        sim.set_schedule(sub)
        best = 0.0
        for rep in range(2):
            t0, done, ms = 0, 0, 0.0
            while done < args.steps:
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); sim.step(args.chunk, table, t0); b.record()
                torch.cuda.synchronize()
                ms += a.elapsed_time(b); t0 = (t0 + args.chunk) % T; done += args.chunk
            best = max(best, n * done / ms * 1e3)
        print(json.dumps({"fpb": fpb, "chunk": args.chunk, "sub_steps": sub, "env_steps_per_s": best}), flush=True)
