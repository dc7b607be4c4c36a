"""GPU-vs-oracle trajectory parity report for the BASELINE.json configurations and every world / model variant:
qpos (rel + abs), qvel and per-leg contact force (the contact sensor's net force) L-inf at 1 ... 1000 steps (writes JSON)."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table, cpg_parameters, TRIPOD_PHASE
from oracle.oracle import Oracle

CHECK = (1, 10, 100, 300, 1000)
T = max(CHECK)


def stance_adhesion(model, n):
    t = np.arange(T) * model.timestep
    legph = np.array([TRIPOD_PHASE[l] for l in model.names["legs"]])
    psi = 2 * np.pi * np.arange(n) / n
    return np.where(np.sin(2 * np.pi * 12.0 * t[None, :, None] + psi[:, None, None] + legph[None, None, :]) < 0, 100.0, 1.0)


def worlds():
    flat = NMFModel.bench(True)
    yield "flat_capsule", flat, -0.17, False
    yield "flat_mesh", NMFModel.bench(False), -0.17, False
    yield "terrain_blocks", NMFModel.bench(True, terrain="blocks"), -0.15, True
    yield "terrain_gapped", NMFModel.bench(True, terrain="gapped"), -0.17, True
    yield "tethered", NMFModel.tethered(), None, False
    yield "legs_active_only", NMFModel.bench(True, joint_preset="legs_active_only"), -0.17, False
    yield "contacts_tibia_tarsus_only", flat.with_contact_bodies("tibia_tarsus_only"), -0.17, False


out = {}
for wname, model, stand_z, stance in worlds():
    nu_pos, nu = model.dim("nu_pos"), model.nu
    key = model.arrays["key_qpos"].copy()
    stand = key.copy()
    if stand_z is not None:
        stand[2] = stand_z
    cpg = cpg_table(model, 4, T).astype(np.float64)
    hold = np.tile(model.arrays["key_ctrl"][:nu_pos], (T, 1))
    adh_on = np.ones((T, 6))
    adh_st = stance_adhesion(model, 4)
    cases = [("1a_hold_neutral_from_keyframe", key, hold, np.zeros((T, 6))), ("1b_zero_actions_from_keyframe", key, np.zeros((T, nu_pos)), np.zeros((T, 6))),
             ("stand_hold_neutral", stand, hold, adh_on)]
    for k in range(4):
        cases.append((f"cpg_fly{k}", stand + np.r_[0.35 * k, 0.22 * k, np.zeros(model.nq - 2)] * (stand_z is not None), cpg[k], adh_st[k] if stance else adh_on))
    n = len(cases)
    sim = B200Simulation(model, n_worlds=n, outputs=True)
    tab = np.zeros((n, T, nu), np.float32)
    for i, (name, q0, pos, adh) in enumerate(cases):
        sim.qpos[i].copy_(torch.as_tensor(q0, dtype=torch.float32))
        tab[i, :, :nu_pos] = pos; tab[i, :, nu_pos:] = adh
    tabd = torch.from_numpy(tab).cuda()
    got, done = {}, 0
    for cp in CHECK:
        sim.step(cp - done, tabd, done); done = cp
        got[cp] = (sim.qpos.cpu().numpy().astype(np.float64), sim.qvel.cpu().numpy().astype(np.float64),
                   sim.sensordata.cpu().numpy().astype(np.float64).reshape(n, 6, 16))
    scen = {}
    for i, (name, q0, pos, adh) in enumerate(cases):
        o = Oracle(model); o.reset(); o.qpos[:] = q0
        res, done = {}, 0
        for cp in CHECK:
            o.step_table_full(tab[i, done:cp].astype(np.float64)); done = cp
            rq, rv, rs = o.qpos.copy(), o.qvel.copy(), o.get("sensordata").copy().reshape(6, 16)
            fo, fg = rs[:, 1:4], got[cp][2][i][:, 1:4]
            res[cp] = {"qpos_rel_linf": float(np.abs(got[cp][0][i] - rq).max() / np.abs(rq).max()),
                       "qpos_abs_linf": float(np.abs(got[cp][0][i] - rq).max()),
                       "qvel_abs_linf": float(np.abs(got[cp][1][i] - rv).max()), "qvel_max": float(np.abs(rv).max()),
                       "contact_force_abs_linf_uN": float(np.abs(fg - fo).max()), "contact_force_max_uN": float(np.abs(fo).max()),
                       "legs_in_contact_oracle": int(rs[:, 0].astype(bool).sum()), "legs_in_contact_gpu": int(got[cp][2][i][:, 0].astype(bool).sum()),
                       "ncon_oracle": o.dim("ncon")}
        scen[name] = res
        print(f"{wname:28s} {name:30s}", {cp: "%.1e" % res[cp]["qpos_rel_linf"] for cp in CHECK}, flush=True)
    out[wname] = scen
    del sim
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/parity_report.json").write_text(json.dumps(out, indent=1))
