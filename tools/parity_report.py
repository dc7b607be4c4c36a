"""GPU-vs-oracle trajectory parity report for the BASELINE.json configurations (writes JSON)."""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table
from oracle.oracle import Oracle

CHECK = (1, 10, 100, 300, 1000)
out = {}
for simplify in (True, False):
    model = NMFModel.bench(simplify_geom=simplify)
    nu_pos = model.dim("nu_pos")
    scen = {}
    # scenarios: (name, qpos0, ctrl-table or None, adhesion ctrl)
    key = model.arrays["key_qpos"].copy()
    stand = key.copy(); stand[2] = -0.17
    tab = cpg_table(model, 4, 1000)
    cases = [("1a_hold_neutral_drop", key, None, 0.0), ("1b_zero_actions_drop", key, np.zeros((1000, nu_pos)), 0.0),
             ("stand_hold_neutral", stand, None, 1.0), ("2_cpg_fly0", stand, tab[0], 1.0), ("2_cpg_fly1", stand, tab[1], 1.0),
             ("2_cpg_fly2", stand, tab[2], 1.0), ("2_cpg_fly3", stand, tab[3], 1.0)]
    n = len(cases)
    sim = B200Simulation(model, n_worlds=n, outputs=False)
    T = np.zeros((n, 1000, nu_pos), np.float32)
    for i, (name, q0, table, adh) in enumerate(cases):
        sim.qpos[i].copy_(torch.as_tensor(q0, dtype=torch.float32))
        sim.ctrl[i, nu_pos:] = adh
        T[i] = np.tile(model.arrays["key_ctrl"][:nu_pos], (1000, 1)) if table is None else table
    Td = torch.from_numpy(T).cuda()
    got, done = {}, 0
    for cp in CHECK:
        sim.step(cp - done, Td, done); done = cp
        got[cp] = (sim.qpos.cpu().numpy().astype(np.float64), sim.qvel.cpu().numpy().astype(np.float64))
    for i, (name, q0, table, adh) in enumerate(cases):
        o = Oracle(model); o.reset(); o.qpos[:] = q0; o.ctrl[nu_pos:] = adh
        res, done = {}, 0
        for cp in CHECK:
            o.step_table(T[i, done:cp].astype(np.float64)); done = cp
            rq = o.qpos.copy(); rv = o.qvel.copy()
            res[cp] = {"qpos_rel_linf": float(np.abs(got[cp][0][i] - rq).max() / np.abs(rq).max()),
                       "qpos_abs_linf": float(np.abs(got[cp][0][i] - rq).max()),
                       "qvel_abs_linf": float(np.abs(got[cp][1][i] - rv).max()), "qvel_max": float(np.abs(rv).max()),
                       "ncon_oracle": o.dim("ncon")}
        scen[name] = res
        print(("capsule " if simplify else "mesh    ") + name, {cp: "%.1e" % res[cp]["qpos_rel_linf"] for cp in CHECK}, flush=True)
    out["capsule" if simplify else "mesh"] = scen
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/parity_report.json").write_text(json.dumps(out, indent=1))
