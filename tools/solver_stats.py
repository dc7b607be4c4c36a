"""Distribution of Newton / line-search trip counts over flies during CPG walking (input to scheduling decisions)."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table
m = NMFModel.bench(simplify_geom=True)
n, T = 4096, 2500
sim = B200Simulation(m, n_worlds=n, outputs=False, debug=True)
table = torch.from_numpy(cpg_table(m, n, T)).cuda()
sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
sim.warmup()
t0 = 0
hist_it, hist_ls, hist_con = np.zeros(16), np.zeros(64), np.zeros(64)
for rep in range(40):
    sim.step(37, table, t0); t0 = (t0 + 37) % T
    sim.step(1, table, t0); t0 = (t0 + 1) % T
    d = sim.debug.cpu().numpy()
    it, ncon, nls = d[:, 0].astype(int), d[:, 1].astype(int), d[:, 2].astype(int)
    hist_it += np.bincount(np.clip(it, 0, 15), minlength=16); hist_ls += np.bincount(np.clip(nls, 0, 63), minlength=64)
    hist_con += np.bincount(np.clip(ncon, 0, 63), minlength=64)
tot = hist_it.sum()
print(json.dumps({"newton_iters_hist": (hist_it / tot).round(4).tolist(), "ls_evals_hist": (hist_ls / tot).round(4).tolist()[:24],
                  "ncon_hist": (hist_con / tot).round(4).tolist()[:32],
                  "mean_iters": float((hist_it * np.arange(16)).sum() / tot), "mean_ls": float((hist_ls * np.arange(64)).sum() / tot)}))
