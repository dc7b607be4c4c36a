"""Throughput of the general-topology kernels (csrc/nmf_tree.cuh): env-steps/s for the ALL_BIOLOGICAL / ALL_POSSIBLE skeletons and,
for comparison, the benchmark skeleton through both kernel families.  Usage: python tools/tree_bench.py [n_flies] [steps]"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table


def run(label, model, n, steps, force_tree=False, precision=32, reps=3):
    os.environ["NMF_FORCE_TREE"] = "1" if force_tree else "0"
    try:
        sim = B200Simulation(model, n_worlds=n, outputs=False)
    finally:
        os.environ.pop("NMF_FORCE_TREE", None)
    sim.set_precision(precision)
    nu_pos = model.dim("nu_pos")
    tab = torch.as_tensor(cpg_table(model, n, steps)).cuda().contiguous()
    q = sim.qpos.clone(); q[:, 2] = -0.17; sim.qpos.copy_(q)
    sim.step(500)                                  # settle
    sim.step(steps, tab, 0)
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); sim.step(steps, tab, 0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    rec = dict(model=label, n_flies=n, steps=steps, precision=precision, nv=model.nv, ms=best, env_steps_per_s=n * steps / best * 1e3,
               finite=bool(torch.isfinite(sim.state).all()), status_or=int(sim.status.max()))
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    run("legs_only (star kernels)", NMFModel.bench(True), n, steps)
    run("legs_only (tree kernels)", NMFModel.bench(True), n, steps, force_tree=True)
    run("all_biological", NMFModel.bench(True, joint_preset="all_biological"), n, steps)
    run("all_biological mesh", NMFModel.bench(False, joint_preset="all_biological"), n, steps)
    run("all_possible + all contacts", NMFModel.bench(True, joint_preset="all_possible", contact_preset="all"), n, steps)
    run("all_biological f64", NMFModel.bench(True, joint_preset="all_biological"), n, steps, precision=64)
