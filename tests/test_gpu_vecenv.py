"""Vector-env wrapper and nmf_forward (mj_forward semantics) on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_forward_refreshes_outputs_without_advancing():
    import torch
    from flygym_b200 import B200Simulation
    sim = B200Simulation(None, n_worlds=3)
    sim.step(5)
    before = sim.state.clone()
    sim.qpos[1, 0] += 2.5                      # move one fly by hand: poses are stale until forward()
    moved = sim.state.clone()
    stale = sim.get_body_positions("nmf").clone()
    sim.forward()
    torch.cuda.synchronize()
    assert torch.equal(sim.state, moved)        # nothing advanced (time, qpos, qvel, warm start untouched)
    fresh = sim.get_body_positions("nmf")
    assert torch.allclose(fresh[1, :, 0], stale[1, :, 0] + 2.5, atol=1e-3)
    assert not torch.equal(before, moved)
    # forward after reset gives the keyframe poses; a subsequent step starts from the same state as without forward
    a, b = B200Simulation(None, n_worlds=2), B200Simulation(None, n_worlds=2)
    a.forward(); a.step(3); b.step(3)
    assert torch.equal(a.state, b.state)


def test_vector_env_loop_and_masked_autoreset():
    import torch
    from flygym_b200 import NMFVectorEnv, NMFModel
    from flygym_b200.actions import cpg_table
    n = 64
    env = NMFVectorEnv(NMFModel.bench(True), n_envs=n, physics_steps_per_action=10, episode_steps=7,
                       odor=([[12.0, 4.0, 1.5]], [[1.0]]), vision=True)
    obs, info = env.reset()
    assert obs["joint_angles"].shape == (n, 66) and obs["vision"].shape == (n, 2, 721, 2) and obs["odor"].shape == (n, 1, 4)
    assert obs["thorax_position"].shape == (n, 3) and float(obs["thorax_position"][:, 2].min()) > 1.0   # spawned above ground
    acts = torch.from_numpy(cpg_table(env.sim.model, n, 200)).cuda()
    env.sim.qpos[:, 2] = -0.15
    total = torch.zeros(n, device="cuda")
    resets = 0
    for t in range(20):
        a = acts[:, 10 * t]
        if t % 2:                                    # alternate the 42- and the 48-column action forms
            a = torch.cat([a, torch.ones(n, 6, device="cuda")], dim=1)
        obs, rew, term, trunc, info = env.step(a)
        assert rew.shape == (n,) and term.dtype == torch.bool and trunc.dtype == torch.bool
        total += rew
        resets += int(info["reset_mask"].sum())
        if t == 6:
            assert bool(trunc.all())                 # episode_steps = 7
            assert abs(env.sim.time) < 1e-9          # ... and every fly was reset on the device
    assert resets >= 2 * n and torch.isfinite(total).all()
    assert torch.isfinite(obs["vision"]).all() and float(obs["vision"].max()) <= 1.0 + 1e-6
    with pytest.raises(ValueError):
        env.step(np.zeros((n, 40), np.float32))
    # masked reset leaves the other flies alone
    env.step(acts[:, 0])
    mask = torch.zeros(n, dtype=torch.bool, device="cuda"); mask[::2] = True
    t_before = env.sim.state[:, env.sim.info.off_time].clone()
    env.reset(mask)
    t_after = env.sim.state[:, env.sim.info.off_time]
    assert bool((t_after[::2] == 0).all()) and torch.equal(t_after[1::2], t_before[1::2])


def test_several_flies_per_world():
    """BaseWorld.add_fly called twice (reference compose/world.py:95-150): flies never collide with each other, so each fly of a
    two-fly world must follow exactly the single-fly trajectory translated by its spawn offset (the ground is translation
    invariant), and per-fly setters must only touch their own fly."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.anatomy import ActuatorType
    from flygym_b200.multifly import B200MultiFlySimulation, B200World
    from flygym_b200.actions import cpg_table
    m = NMFModel.bench(True)
    w = B200World(m)
    w.add_fly("alice", (0.0, 0.0, 0.8))
    w.add_fly("bob", (4.0, -3.0, 0.8))
    with pytest.raises(ValueError):
        w.add_fly("bob", (1.0, 1.0, 0.8))
    n = 3
    multi = B200MultiFlySimulation(w, n_worlds=n)
    single = B200Simulation(m, n_worlds=n)
    acts = torch.from_numpy(cpg_table(m, n, 400)).cuda()
    multi.set_leg_adhesion_states("alice", np.ones(6, dtype=bool)); multi.set_leg_adhesion_states("bob", np.ones(6, dtype=bool))
    single.set_leg_adhesion_states("nmf", np.ones(6, dtype=bool))
    for t in range(0, 400, 20):
        multi.set_actuator_inputs("alice", ActuatorType.POSITION, acts[:, t])
        multi.set_actuator_inputs("bob", ActuatorType.POSITION, acts[:, t])
        single.set_actuator_inputs("nmf", ActuatorType.POSITION, acts[:, t])
        multi.step(20); single.step(20)
    ref = single.get_body_positions("nmf")
    a, b = multi.get_body_positions("alice"), multi.get_body_positions("bob")
    assert a.shape == (n, 69, 3) and torch.equal(a, ref)
    off = torch.tensor([4.0, -3.0, 0.0], device="cuda")
    assert torch.allclose(b - off, ref, atol=2e-4)                  # fp32 round-off of the translated coordinates only
    assert torch.allclose(multi.get_joint_angles("bob"), single.get_joint_angles("nmf"), atol=2e-4)
    assert float(ref[:, 0, 2].max()) < 2.0 and bool((multi.get_ground_contact_info("bob")[0].sum(dim=1) >= 0).all())
    # a setter addressed to one fly leaves the other alone
    multi.set_actuator_inputs("alice", ActuatorType.POSITION, np.zeros(42))
    multi.step(50); single.step(50)
    assert not torch.allclose(multi.get_joint_angles("alice"), single.get_joint_angles("nmf"), atol=1e-3)
    assert torch.allclose(multi.get_joint_angles("bob"), single.get_joint_angles("nmf"), atol=5e-4)
    with pytest.raises(ValueError):
        multi.set_actuator_inputs("alice", ActuatorType.POSITION, np.zeros(40))
    with pytest.raises(KeyError):
        multi.get_joint_angles("carol")


def test_trajectory_recorder_and_state_export(tmp_path):
    import torch
    from flygym_b200 import B200Simulation
    from flygym_b200.trajectory import TrajectoryRecorder
    sim = B200Simulation(None, n_worlds=5)
    rec = TrajectoryRecorder(sim, capacity=4, every=2, with_qvel=True)
    kept = []
    for t in range(10):
        sim.step(3)
        if rec.record():
            kept.append(sim.qpos.clone())
    assert rec.count == 4 and len(kept) == 4                       # calls 0, 2, 4, 6; the buffer is full afterwards
    data = rec.gather()
    assert data["qpos"].shape == (4, 5, 73) and data["qvel"].shape == (4, 5, 72)
    assert np.array_equal(data["qpos"][2], kept[2].cpu().numpy())
    assert np.allclose(data["time"], [3e-4, 9e-4, 15e-4, 21e-4], rtol=1e-5)
    assert rec.save(tmp_path / "traj.npz") and np.load(tmp_path / "traj.npz")["qpos"].shape == (4, 5, 73)
    st = sim.export_state(3)
    assert st["qpos"].shape == (73,) and st["qvel"].shape == (72,) and abs(st["time"] - 30e-4) < 1e-7
    assert np.array_equal(st["qpos"].astype(np.float32), sim.qpos[3].cpu().numpy())


def test_two_handles_on_two_gpus_in_one_process():
    """Every C-ABI entry point runs under a device guard: simulations (and their eye cameras) on different GPUs coexist in one process,
    whatever the caller's current device is, and leave it as they found it.  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from flygym_b200 import B200Simulation
    from flygym_b200.retina import EyeCameras, Retina
    torch.cuda.set_device(0)
    a = B200Simulation(None, n_worlds=8, device="cuda:0")
    b = B200Simulation(None, n_worlds=8, device="cuda:1")
    assert torch.cuda.current_device() == 0
    for s in (a, b):
        s.qpos[:, 2] = -0.15
    a.step(20); b.step(20); a.step(5); b.step(5)
    ea, eb = EyeCameras(a, Retina(device="cuda:0")), EyeCameras(b, Retina(device="cuda:1"))
    ra, rb = ea.retina(), eb.retina()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0
    assert a.qpos.device.index == 0 and b.qpos.device.index == 1 and rb.device.index == 1
    assert torch.equal(a.state.cpu(), b.state.cpu()) and torch.equal(ra.cpu(), rb.cpu())       # same model, same steps: same bits on both GPUs
    qh = np.empty((8, a.info.nq), np.float32)
    b.step_host(np.tile(a.model.arrays["key_ctrl"][:42].astype(np.float32), (8, 1)), 1, qh)
    assert np.isfinite(qh).all() and torch.cuda.current_device() == 0
