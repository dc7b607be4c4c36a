"""GPU-side exploration: solver statistics and timing sweeps (not part of the test-suite)."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table, replay_table

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[4096])
ap.add_argument("--chunk", type=int, default=100)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--mesh", action="store_true")
ap.add_argument("--stats", action="store_true")
ap.add_argument("--actions", default="cpg")
args = ap.parse_args()
model = NMFModel.bench(simplify_geom=not args.mesh)
T = 2500
for n in args.n:
    sim = B200Simulation(model, n_worlds=n, outputs=False, debug=args.stats)
    table = torch.from_numpy(cpg_table(model, n, T) if args.actions == 'cpg' else replay_table(model, n, 1000)).cuda()
    T = table.shape[1]
    sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
    sim.warmup()
    t0 = 0
    for _ in range(3):
        sim.step(args.chunk, table, t0); t0 = (t0 + args.chunk) % T
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    done = 0
    while done < args.steps:
        sim.step(args.chunk, table, t0); t0 = (t0 + args.chunk) % T; done += args.chunk
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    out = {"n": n, "chunk": args.chunk, "ms_per_step": ms / done, "env_steps_per_s": n * done / ms * 1e3}
    if args.stats:
        d = sim.debug.cpu().numpy()
        out.update(niter_mean=float(d[:, 0].mean()), niter_max=float(d[:, 0].max()), ncon_mean=float(d[:, 1].mean()),
                   nls_mean=float(d[:, 2].mean()), nls_max=float(d[:, 2].max()),
                   z_mean=float(sim.qpos[:, 2].mean()), finite=bool(torch.isfinite(sim.state).all()))
    print(json.dumps(out), flush=True)
    del sim, table
