"""GPU parity: the sm_100a step kernel (through the C ABI) vs the CPU fp64 oracle.

Tolerances (fp32 kernel vs fp64 oracle, PARITY UNPINNED against real MuJoCo — see
oracle/nmf_oracle.c): after N steps from identical state and action sequence
``|qpos_gpu - qpos_oracle|_inf / |qpos_oracle|_inf <= 1e-4`` (BASELINE.json's
north_star tolerance), checked at 1, 100 and 1000 steps.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_traj(model, qpos0, qvel0, ctrl_fn, nsteps, checkpoints):
    from oracle.oracle import Oracle
    o = Oracle(model)
    o.reset()
    o.qpos[:] = qpos0
    o.qvel[:] = qvel0
    out = {}
    for s in range(nsteps):
        c = ctrl_fn(s)
        if c is not None:
            o.ctrl[:] = c
        o.step()
        if s + 1 in checkpoints:
            out[s + 1] = (o.qpos.copy(), o.qvel.copy())
    return out


def _perturbed_states(model, n, seed=0):
    rng = np.random.default_rng(seed)
    q = np.tile(model.arrays["key_qpos"], (n, 1))
    v = np.zeros((n, model.nv))
    for i in range(1, n):
        q[i, 7:] += 0.15 * rng.standard_normal(model.nq - 7)
        q[i, 2] = rng.uniform(-0.3, 0.9)
        quat = np.array([1.0, 0, 0, 0]) + 0.1 * rng.standard_normal(4)
        q[i, 3:7] = quat / np.linalg.norm(quat)
        v[i] = rng.standard_normal(model.nv) * np.r_[np.full(3, 5.0), np.full(3, 1.0), np.full(model.nv - 6, 3.0)]
    return q, v


@pytest.mark.parametrize("simplify", [True, False])
def test_trajectory_parity(simplify):
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    model = NMFModel.bench(simplify_geom=simplify)
    n = 6
    q0, v0 = _perturbed_states(model, n, seed=3)
    sim = B200Simulation(model, n_worlds=n)
    sim.qpos.copy_(torch.as_tensor(q0, dtype=torch.float32))
    sim.qvel.copy_(torch.as_tensor(v0, dtype=torch.float32))
    checkpoints = (1, 100, 1000)
    done = 0
    got = {}
    for cp in checkpoints:
        sim.step(cp - done)
        done = cp
        got[cp] = (sim.qpos.cpu().numpy().astype(np.float64), sim.qvel.cpu().numpy().astype(np.float64))
    worst = {}
    for i in range(n):
        ref = _oracle_traj(model, q0[i], v0[i], lambda s: None, 1000, checkpoints)
        for cp in checkpoints:
            rq = ref[cp][0]
            err = np.abs(got[cp][0][i] - rq).max() / np.abs(rq).max()
            worst[cp] = max(worst.get(cp, 0.0), err)
    print("qpos rel Linf vs oracle:", worst)
    assert worst[1] < 1e-5
    assert worst[100] < 1e-4
    assert worst[1000] < 1e-4


def test_api_shapes_and_time():
    import torch
    from flygym_b200 import B200Simulation, ActuatorType
    n = 4
    sim = B200Simulation(None, n_worlds=n)
    dt = sim.timestep
    assert sim.time == 0.0
    sim.step(10)
    assert abs(sim.time - 10 * dt) < 1e-7
    fly = "nmf"
    assert sim.get_joint_angles(fly).shape == (n, 66)
    assert sim.get_joint_velocities(fly).shape == (n, 66)
    assert sim.get_body_positions(fly).shape == (n, 69, 3)
    q = sim.get_body_rotations(fly)
    assert q.shape == (n, 69, 4)
    assert torch.allclose(q.norm(dim=-1), torch.ones_like(q[..., 0]), atol=1e-5)
    assert sim.get_site_positions(fly).shape == (n, 68, 3)
    assert sim.get_actuator_forces(fly, ActuatorType.POSITION).shape == (n, 42)
    info = sim.get_ground_contact_info(fly)
    assert [tuple(t.shape) for t in info] == [(n, 6)] + [(n, 6, 3)] * 5
    with pytest.raises(ValueError):
        sim.set_actuator_inputs(fly, ActuatorType.POSITION, np.zeros((n, 41), np.float32))
    with pytest.raises(ValueError):
        sim.set_leg_adhesion_states(fly, np.zeros((n, 5), np.float32))
    sim.set_actuator_inputs(fly, ActuatorType.POSITION, np.zeros((n, 42), np.float32))
    sim.set_leg_adhesion_states(fly, torch.ones((n, 6), device="cuda"))
    assert float(sim.ctrl[:, :42].abs().max()) == 0.0
    assert float(sim.ctrl[:, 42:].min()) == 1.0
    sim.reset()
    assert sim.time == 0.0
    assert float(sim.qvel.abs().max()) == 0.0
