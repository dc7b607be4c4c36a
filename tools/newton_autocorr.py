"""B200: how predictable is a fly's Newton pass count from its previous step?  (Would regrouping the flies of a lockstep block by
their last pass count shorten the wait at the pass barrier?)  4096 CPG walkers, 300 single-step launches with the debug dump."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table

model = NMFModel.bench(True)
n, T = 4096, 2500
sim = B200Simulation(model, n_worlds=n, outputs=False, debug=True)
table = torch.from_numpy(cpg_table(model, n, T)).cuda()
sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
sim.warmup()
sim.step(500, table, 0)
its = []
for s in range(300):
    sim.step(1, table, 500 + s)
    its.append(sim.debug[:, 0].clone())
it = torch.stack(its).cpu().numpy()            # (steps, flies)
a, b = it[:-1].ravel(), it[1:].ravel()
out = {"mean": float(it.mean()), "hist": (np.bincount(it.astype(int).ravel()) / it.size).round(4).tolist(),
       "lag1_corr": float(np.corrcoef(a, b)[0, 1]), "p_same_as_previous": float((a == b).mean())}
blocks = it.reshape(it.shape[0], n // 8, 8)
out["mean_max_of_8_consecutive"] = float(blocks.max(2).mean())
# oracle regrouping: sort the flies of every step by their PREVIOUS step's count, then blocks of 8
order = np.argsort(it[:-1], axis=1, kind="stable")
nxt = np.take_along_axis(it[1:], order, axis=1).reshape(it.shape[0] - 1, n // 8, 8)
out["mean_max_of_8_sorted_by_previous"] = float(nxt.max(2).mean())
srt = np.sort(it, axis=1).reshape(it.shape[0], n // 8, 8)
out["mean_max_of_8_perfect_sort"] = float(srt.max(2).mean())
# over an 8-step work item the block pays the per-step maxima
it8 = it[: it.shape[0] // 8 * 8].reshape(-1, 8, n)
out["item8_sum_of_step_maxima_consecutive"] = float(it8.reshape(-1, 8, n // 8, 8).max(3).sum(1).mean())
o8 = np.argsort(it8[:-1].sum(1), axis=1, kind="stable")          # regroup by the previous ITEM's total
nx8 = np.take_along_axis(it8[1:], o8[:, None, :], axis=2).reshape(it8.shape[0] - 1, 8, n // 8, 8)
out["item8_sum_of_step_maxima_sorted_by_previous_item"] = float(nx8.max(3).sum(1).mean())
out["item8_sum_of_means"] = float(it8.sum(1).mean())
print(json.dumps(out, indent=1))
