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


def _run_cases(model, cases, checkpoints, precision=32):
    """cases: list of (qpos0, qvel0, table[T, nu_pos] or None, adhesion ctrl). Returns {cp: [rel err per case]}."""
    import torch
    from flygym_b200 import B200Simulation
    from oracle.oracle import Oracle
    n, nu_pos, T = len(cases), model.dim("nu_pos"), max(checkpoints)
    sim = B200Simulation(model, n_worlds=n, outputs=False)
    sim.set_precision(precision)
    tab = np.zeros((n, T, nu_pos), np.float32)
    for i, (q0, v0, table, adh) in enumerate(cases):
        sim.qpos[i].copy_(torch.as_tensor(q0, dtype=torch.float32))
        sim.qvel[i].copy_(torch.as_tensor(v0, dtype=torch.float32))
        sim.ctrl[i, nu_pos:] = adh
        tab[i] = np.tile(model.arrays["key_ctrl"][:nu_pos], (T, 1)) if table is None else table[:T]
    tabd = torch.from_numpy(tab).cuda()
    got, done = {}, 0
    for cp in checkpoints:
        sim.step(cp - done, tabd, done)
        done = cp
        got[cp] = sim.qpos.cpu().numpy().astype(np.float64)
    errs = {cp: [] for cp in checkpoints}
    for i, (q0, v0, table, adh) in enumerate(cases):
        o = Oracle(model)
        o.reset()
        o.qpos[:] = np.asarray(q0, np.float32)         # the SAME initial state: what the float32 record of the device holds (a state
        o.qvel[:] = np.asarray(v0, np.float32)         # that differs by one float32 rounding is 1e-2 away after 1000 steps of a tumbling fly)
        o.ctrl[nu_pos:] = adh
        done = 0
        for cp in checkpoints:
            o.step_table(tab[i, done:cp].astype(np.float64))
            done = cp
            errs[cp].append(float(np.abs(got[cp][i] - o.qpos).max() / np.abs(o.qpos).max()))
    return errs


@pytest.mark.parametrize("simplify", [True, False])
def test_config1_1000_steps(simplify):
    """BASELINE.json config 1 / north_star: 1 fly, flat terrain, 1000 steps holding the neutral action ->
    qpos within 1e-4 rel of the CPU oracle after 1000 steps (measured: ~2e-6)."""
    from flygym_b200 import NMFModel
    model = NMFModel.bench(simplify_geom=simplify)
    key = model.arrays["key_qpos"].copy()
    stand = key.copy()
    stand[2] = -0.17
    errs = _run_cases(model, [(key, np.zeros(model.nv), None, 0.0), (stand, np.zeros(model.nv), None, 1.0)], (1, 100, 1000))
    print("config-1 qpos rel Linf:", errs)
    assert max(errs[1]) < 1e-6 and max(errs[100]) < 1e-5 and max(errs[1000]) < 1e-4
    # config 1b: "zero action" read literally -- set_actuator_inputs(zeros(42)) as tests/warp/test_simulation.py:284-297
    # does: every position target 0 rad, so the legs fold away from the neutral pose while the fly drops
    zero = np.zeros((1000, model.dim("nu_pos")))
    errs_b = _run_cases(model, [(key, np.zeros(model.nv), zero, 0.0), (stand, np.zeros(model.nv), zero, 1.0)], (1, 100, 1000))
    print("config-1b (literal zero targets) qpos rel Linf:", errs_b)
    # float32 at 1000 steps: 1e-6 ... 2e-2 -- with all targets at 0 rad the legs snap away from the neutral pose, the fly jumps, tumbles and
    # lands on 1-4 contacts: a chaotic trajectory on which float32 rounding (like a one-ulp change of the initial state in float64, see
    # _run_cases) is amplified 1e5-fold; the f64 build below stays at 2e-8
    assert max(errs_b[1]) < 1e-6 and max(errs_b[100]) < 1e-4 and max(errs_b[1000]) < 5e-2
    errs_b64 = _run_cases(model, [(key, np.zeros(model.nv), zero, 0.0), (stand, np.zeros(model.nv), zero, 1.0)], (1000,), precision=64)
    print("config-1b, f64 build:", errs_b64)
    assert max(errs_b64[1000]) < 1e-6                                                              # (measured 2e-8: far inside the north star's 1e-4)


@pytest.mark.parametrize("simplify", [True, False])
def test_cpg_and_perturbed_parity(simplify):
    """Config 2 (CPG tripod actions, adhesion on) and randomly perturbed / penetrating initial states.
    Walking contact dynamics are chaotic (stick-slip, support-vertex switches), so an fp32 trajectory cannot shadow the
    fp64 one indefinitely: we require 1e-4 for every fly up to 100 steps, a median below 1e-5 (max 1e-2) at 300 steps and a
    median below 1e-3 (max 5e-2) at 1000 steps (measured on B200: most flies ~1e-6, occasional 1e-4..1e-2 after a contact
    event resolves one step apart in the two precisions)."""
    from flygym_b200 import NMFModel
    from flygym_b200.actions import cpg_table
    model = NMFModel.bench(simplify_geom=simplify)
    stand = model.arrays["key_qpos"].copy()
    stand[2] = -0.17
    tab = cpg_table(model, 6, 1000)
    cases = [(stand, np.zeros(model.nv), tab[i], 1.0) for i in range(6)]
    errs = _run_cases(model, cases, (1, 100, 300, 1000))
    print("cpg qpos rel Linf:", {k: ["%.1e" % e for e in v] for k, v in errs.items()})
    assert max(errs[1]) < 1e-6 and max(errs[100]) < 1e-4
    assert np.median(errs[300]) < 1e-5 and max(errs[300]) < 1e-2
    assert np.median(errs[1000]) < 1e-3 and max(errs[1000]) < 5e-2
    q0, v0 = _perturbed_states(model, 6, seed=3)
    errs = _run_cases(model, [(q0[i], v0[i], None, 0.0) for i in range(6)], (1, 10, 100))
    print("perturbed qpos rel Linf:", {k: ["%.1e" % e for e in v] for k, v in errs.items()})
    assert max(errs[1]) < 1e-5 and max(errs[10]) < 1e-4 and max(errs[100]) < 1e-3


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


def test_control_changes_angles_and_input_types():
    """Mirrors tests/warp/test_simulation.py:254-297 of the reference: numpy and device inputs are accepted,
    and different controls give different joint angles after 50 steps."""
    import torch
    from flygym_b200 import B200Simulation, ActuatorType
    n = 3
    sim = B200Simulation(None, n_worlds=n)
    sim.warmup(0.005)
    a0 = sim.get_joint_angles("nmf").clone()
    sim.set_actuator_inputs("nmf", ActuatorType.POSITION, np.zeros((n, 42), dtype=np.float32))          # numpy
    sim.step(50)
    a1 = sim.get_joint_angles("nmf").clone()
    sim.set_actuator_inputs("nmf", ActuatorType.POSITION, torch.full((n, 42), 0.5, device="cuda"))        # device tensor
    sim.step(50)
    a2 = sim.get_joint_angles("nmf")
    assert (a1 - a0).abs().max() > 1e-3 and (a2 - a1).abs().max() > 1e-3
    assert isinstance(a2, torch.Tensor) and a2.dtype == torch.float32 and a2.shape == (n, 66)
    f = sim.get_actuator_forces("nmf", ActuatorType.POSITION)
    assert float(f.abs().max()) <= 30.0 + 1e-4            # forcerange (-30, 30), fly.py:305-306
    adh = sim.get_actuator_forces("nmf", ActuatorType.ADHESION)
    assert adh.shape == (n, 6) and float(adh.min()) >= 1.0 - 1e-6   # ctrlrange (1, 100) clamps 0 -> 1


def test_masked_reset_and_site_positions():
    import torch
    from flygym_b200 import B200Simulation
    n = 4
    sim = B200Simulation(None, n_worlds=n)
    sim.step(30)
    before = sim.qpos.clone()
    sim.reset(mask=[True, False, True, False])
    key = torch.as_tensor(sim.model.arrays["key_qpos"], dtype=torch.float32, device="cuda")
    assert torch.equal(sim.qpos[0], key) and torch.equal(sim.qpos[2], key)
    assert torch.equal(sim.qpos[1], before[1]) and torch.equal(sim.qpos[3], before[3])
    assert sim.time == 0.0        # world 0 was reset
    sim.step(1)
    sites = sim.get_site_positions("nmf")
    bodies = sim.get_body_positions("nmf")
    segs, names = sim.model.names["segments"], sim.model.names["sites"]
    idx = [segs.index(s.split("-")[1]) for s in names]
    assert torch.equal(sites, bodies[:, idx])       # sites sit at the child-segment origins (fly.py:371-405)


def test_step_is_cuda_graph_capturable():
    """The reference benchmark replays the step as a CUDA graph (time_gpu_simulation.py:137-150): the C ABI must not
    allocate, synchronise or touch the default stream."""
    import torch
    from flygym_b200 import B200Simulation
    sim = B200Simulation(None, n_worlds=8, outputs=False)
    ref = B200Simulation(None, n_worlds=8, outputs=False)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sim.step(1)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            sim.step(5)
        for _ in range(4):
            g.replay()
    torch.cuda.synchronize()
    ref.step(1 + 4 * 5)      # the captured launch itself does not execute during capture
    torch.cuda.synchronize()
    assert torch.equal(sim.state, ref.state)
    # the work-queue launch (more flies than resident blocks; counters cleared by a captured memset node) replays as well
    n = 2500
    big, big_ref = B200Simulation(None, n_worlds=n, outputs=False), B200Simulation(None, n_worlds=n, outputs=False)
    for b in (big, big_ref):
        b.qpos[:, 2] = -0.15
    with torch.cuda.stream(s):
        big.step(1)
        g2 = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g2, stream=s):
            big.step(50)
        for _ in range(3):
            g2.replay()
    torch.cuda.synchronize()
    big_ref.set_schedule(0)
    big_ref.step(1 + 3 * 50)
    torch.cuda.synchronize()
    assert torch.equal(big.state, big_ref.state)


def test_work_queue_schedule_is_bit_identical():
    """Multi-step launches of more flies than the GPU holds resident blocks are served from a device-side work queue
    (sub-chunks of `sub_steps` steps, or the heuristic's tapered items); the schedule must not change a single bit of any fly's trajectory."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    m = NMFModel.bench(simplify_geom=True)
    n = 3001                                         # > 148 SMs x 16 resident blocks, and not a multiple of anything
    table = torch.from_numpy(cpg_table(m, n, 64)).cuda()
    finals = []
    for sub in (0, 10, 7, -1):                       # no queue, uniform items of 10 / 7 steps, the heuristic (tapered: 8 8 7 7 6 4 3 2)
        sim = B200Simulation(m, n_worlds=n, outputs=True)
        sim.set_schedule(sub)
        sim.qpos[:, 2] = -0.15                       # feet on the ground
        sim.qpos[:, 0] += torch.linspace(0, 1, n, device="cuda")     # distinct flies
        sim.step(45, table, 3)
        sim.step(1, table, 48)                       # a 1-step launch (never queued) on top
        torch.cuda.synchronize()
        finals.append((sim.state.clone(), sim.seg_xpos.clone(), sim.sensordata.clone()))
    for other in finals[1:]:
        for a, b in zip(finals[0], other):
            assert torch.equal(a, b)
    assert torch.isfinite(finals[0][0]).all()
    assert abs(float(finals[0][0][0, 300]) - 46e-4) < 1e-6      # time advanced by 46 steps


@pytest.mark.parametrize("terrain", ["blocks", "gapped"])
def test_terrain_walking_parity(terrain):
    """BASELINE config 3: CPG gait with stance-phase adhesion (ctrl 100 in stance, 1 in swing) on box-column terrain,
    flies spread over the tile / gap pattern.  Contact dynamics over edges are more chaotic than on the plane (a foot
    sliding off a tile edge switches contact normal): 1e-6 after one step and 1e-4 up to 100 steps for every fly, median
    below 1e-4 (max 5e-2) at 300 steps (measured on B200: median 1.2e-5, one fly 8e-3 after an edge event)."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_parameters, TRIPOD_PHASE
    from oracle.oracle import Oracle
    model = NMFModel.bench(simplify_geom=True, terrain=terrain)
    n, T, nu_pos = 8, 300, model.dim("nu_pos")
    neutral, amp, phase = cpg_parameters(model)
    t = np.arange(T) * model.timestep
    tab = np.zeros((n, T, nu_pos + 6))
    legph = np.array([TRIPOD_PHASE[l] for l in model.names["legs"]])
    for k in range(n):
        base = 2 * np.pi * 12.0 * t[:, None] + 2 * np.pi * k / n
        tab[k, :, :nu_pos] = neutral + amp * np.sin(base + phase)
        tab[k, :, nu_pos:] = np.where(np.sin(base + legph) < 0, 100.0, 1.0)
    tab32 = tab.astype(np.float32)
    sim = B200Simulation(model, n_worlds=n, outputs=True)
    q0 = np.tile(model.arrays["key_qpos"], (n, 1))
    q0[:, 2] = -0.15 if terrain == "blocks" else -0.17
    q0[:, 0] += np.linspace(0.0, 1.4, n)                     # different phases relative to the tile / gap pattern
    q0[:, 1] += np.linspace(0.0, 0.9, n)
    sim.qpos.copy_(torch.as_tensor(q0, dtype=torch.float32))
    tabd = torch.from_numpy(tab32).cuda()
    got, done = {}, 0
    for cp in (1, 100, 300):
        sim.step(cp - done, tabd, done); done = cp
        got[cp] = sim.qpos.cpu().numpy().astype(np.float64)
    found = sim.get_ground_contact_info("nmf")[0].cpu().numpy()
    errs = {cp: [] for cp in got}
    ncon_total = 0
    for k in range(n):
        o = Oracle(model); o.reset(); o.qpos[:] = q0[k]
        done = 0
        for cp in (1, 100, 300):
            o.step_table_full(tab32[k, done:cp].astype(np.float64)); done = cp
            errs[cp].append(float(np.abs(got[cp][k] - o.qpos).max() / np.abs(o.qpos).max()))
        ncon_total += o.dim("ncon")
    print(terrain, "qpos rel Linf:", {k: ["%.1e" % e for e in v] for k, v in errs.items()})
    assert ncon_total > 0 and found.sum() > 0
    assert max(errs[1]) < 1e-6 and max(errs[100]) < 1e-4
    assert np.median(errs[300]) < 1e-4 and max(errs[300]) < 5e-2
    assert torch.isfinite(sim.state).all()


def test_full_size_properties():
    """BASELINE full sizes (4096 and 32768 flies per GPU) through size-independent properties: flies that start from the same
    state with the same actions end bit-identically wherever they sit in the batch (scheduling independence, work queue on),
    time = n dt, unit quaternions, finite state; and the C ABI rejects ragged action tables."""
    import ctypes
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    m = NMFModel.bench(simplify_geom=True)
    base = torch.from_numpy(cpg_table(m, 4, 64))                     # 4 distinct action sequences
    for n in (4096, 32768):
        table = base.repeat(n // 4, 1, 1).contiguous().cuda()        # fly k follows sequence k % 4
        sim = B200Simulation(m, n_worlds=n, outputs=False)
        sim.qpos[:, 2] = -0.15
        sim.step(60, table, 0)
        torch.cuda.synchronize()
        st = sim.state.view(n // 4, 4, -1)
        assert torch.equal(st, st[:1].expand_as(st))                 # every replica of a sequence is bit-identical
        assert not torch.equal(st[0, 0], st[0, 1])                   # ... and the sequences differ
        assert torch.isfinite(sim.state).all()
        assert torch.allclose(sim.state[:, sim.info.off_time], torch.full((n,), 60e-4, device="cuda"), rtol=1e-5)
        q = sim.qpos[:, 3:7]
        assert torch.allclose(q.norm(dim=1), torch.ones(n, device="cuda"), atol=1e-5)
        # ragged / mistyped tables are refused, by the Python class and by the C ABI itself
        with pytest.raises(ValueError):
            sim.step(1, table[:, :, :41].contiguous(), 0)
        with pytest.raises(ValueError):
            sim.step(1, table[: n - 1], 0)
        rc = sim._lib.nmf_step(sim._h, 1, ctypes.c_void_p(table.data_ptr()), 64, 0, 41, sim._stream())
        assert rc == -1 and b"action table" in sim._lib.nmf_last_error(sim._h)
        rc = sim._lib.nmf_step(sim._h, 1, ctypes.c_void_p(table.data_ptr()), 0, 0, 42, sim._stream())
        assert rc == -1
        del sim, table


def test_step_host_pipelined_slices_match_device_path():
    """nmf_step_host cuts the batch into slices pipelined over private streams; the result must equal the plain device path
    bit for bit (42- and 48-column action forms, odd batch size)."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    m = NMFModel.bench(simplify_geom=True)
    n = 2051
    tab = cpg_table(m, n, 6)
    adh = np.where(np.arange(n * 6).reshape(n, 6) % 3 == 0, 50.0, 1.0).astype(np.float32)
    a, b = B200Simulation(m, n_worlds=n, outputs=True), B200Simulation(m, n_worlds=n, outputs=True)
    for s in (a, b):
        s.qpos[:, 2] = -0.15
    qh = np.empty((n, m.nq), np.float32)
    for t in range(6):
        if t % 2 == 0:
            act = np.ascontiguousarray(tab[:, t])
            b.ctrl[:, :42] = torch.from_numpy(act).cuda()
        else:
            act = np.ascontiguousarray(np.concatenate([tab[:, t], adh], axis=1))
            b.ctrl[:, :48] = torch.from_numpy(act).cuda()
        a.step_host(act, 1, qh)
        b.step(1)
        torch.cuda.synchronize()
        assert np.array_equal(qh, b.qpos.cpu().numpy())
    assert torch.equal(a.state, b.state) and torch.equal(a.seg_xpos, b.seg_xpos) and torch.equal(a.sensordata, b.sensordata)


def test_step_host_graph_replay_with_pinned_buffers():
    """With pinned host buffers nmf_step_host replays its slice pipeline as a CUDA graph (finer slices, host addresses patched into
    the copy nodes from call to call).  Same bits as the plain device path: action rows at changing addresses, two result buffers,
    the 42- and 48-column forms (a new graph), a setter in between (new epoch), nsteps > 1, and a pageable call in the middle."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    m = NMFModel.bench(simplify_geom=True)
    n = 4099
    T = 10
    tab = cpg_table(m, n, T)
    adh = np.where(np.arange(n * 6).reshape(n, 6) % 3 == 0, 50.0, 1.0).astype(np.float32)
    a, b = B200Simulation(m, n_worlds=n, outputs=True), B200Simulation(m, n_worlds=n, outputs=True)
    for s in (a, b):
        s.qpos[:, 2] = -0.15
    act42 = torch.from_numpy(np.ascontiguousarray(tab.transpose(1, 0, 2))).pin_memory()                  # (T, n, 42): row t at its own address
    act48 = torch.from_numpy(np.ascontiguousarray(np.concatenate([tab.transpose(1, 0, 2), np.broadcast_to(adh, (T, n, 6))], axis=2))).pin_memory()
    res = [torch.empty((n, m.nq), dtype=torch.float32).pin_memory() for _ in range(2)]
    l0 = a.launch_count
    for t in range(T):
        k = 1 if t != 7 else 3
        if t in (3, 4, 8):
            act = act48[t]; b.ctrl[:, :48] = act.cuda()
        else:
            act = act42[t]; b.ctrl[:, :42] = act.cuda()
        if t == 5:
            a.set_flies_per_block(4); b.set_flies_per_block(4)                                            # new epoch -> the graph is rebuilt
        out = res[t % 2]
        if t == 6:                                                                                          # pageable buffers: call-by-call path
            qh = np.empty((n, m.nq), np.float32); a.step_host(act.numpy().copy(), k, qh); got = qh
        else:
            a.step_host(act.numpy(), k, out.numpy()); got = out.numpy()
        b.step(k)
        torch.cuda.synchronize()
        assert np.array_equal(got, b.qpos.cpu().numpy()), t
    assert a.launch_count - l0 >= T                      # the replayed launches are counted
    assert torch.equal(a.state, b.state) and torch.equal(a.seg_xpos, b.seg_xpos) and torch.equal(a.sensordata, b.sensordata)


def test_replay_table_built_on_device():
    """nmf_replay_table (cubic resampling of the recorded clip + per-world tiling on the GPU) equals the host construction the
    reference uses (MotionSnippet.get_joint_angles + ReplayTargetData.make_target_angles_all_worlds), incl. the rank offset."""
    import torch
    from flygym_b200 import NMFModel
    from flygym_b200.actions import replay_table, replay_table_device
    m = NMFModel.bench(True)
    for n, T, off in ((37, 1000, 0), (8, 1000, 29), (5, 2500, 3)):
        host = replay_table(m, n, T, fly_offset=off)
        dev = replay_table_device(m, n, T, "cuda", fly_offset=off).cpu().numpy()
        assert dev.shape == host.shape == (n, T, 42)
        assert np.abs(dev - host).max() <= 2.4e-7          # fp64 evaluation, one float32 rounding apart at most
    with pytest.raises(ValueError):
        replay_table_device(m, 2, 30000, "cuda")


def test_fp64_kernel_meets_the_north_star_tolerance_for_every_walking_fly():
    """BASELINE north_star: qpos within 1e-4 rel of the CPU reference after 1000 steps.  Walking contact dynamics amplify float32
    round-off by ~1e5 over 1000 steps, so single flies of the fp32 kernel leave that band (test_cpg_and_perturbed_parity).  The
    fp64 instantiation of the SAME kernel source stays within float32 rounding of the fp64 oracle for EVERY fly: the kernel's
    algorithm is the oracle's; what separates the fp32 trajectories is arithmetic precision only."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    from oracle.oracle import Oracle
    for terrain, n, T in ((None, 8, 1000), ("blocks", 4, 400)):
        m = NMFModel.bench(True, terrain=terrain)
        a = dict(m.arrays); opt = a["opt"].copy(); opt[5] = 1e-16; a["opt"] = opt
        mt = NMFModel(a, m.names, m.meta)
        tab = cpg_table(m, n, T)
        q0 = np.tile(m.arrays["key_qpos"], (n, 1)); q0[:, 2] = -0.17; q0[:, 0] += np.linspace(0, 1.0, n)
        q0 = q0.astype(np.float32)
        out = {}
        for bits in (64, 32):
            sim = B200Simulation(m, n_worlds=n, outputs=False)
            sim.set_precision(bits)
            sim.qpos.copy_(torch.from_numpy(q0)); sim.ctrl[:, 42:] = 1.0
            sim.step(T, torch.from_numpy(tab).cuda(), 0)            # one launch: the state stays in double for all T steps
            out[bits] = sim.qpos.cpu().numpy().astype(np.float64)
        errs = {64: [], 32: []}
        for k in range(n):
            o = Oracle(mt); o.reset(); o.qpos[:] = q0[k].astype(np.float64); o.ctrl[42:] = 1.0
            o.step_table(tab[k].astype(np.float64))
            for bits in (64, 32):
                errs[bits].append(float(np.abs(out[bits][k] - o.qpos).max() / np.abs(o.qpos).max()))
        print(terrain, "fp64:", ["%.1e" % e for e in errs[64]], "fp32:", ["%.1e" % e for e in errs[32]])
        assert max(errs[64]) < 1e-6, errs          # north_star tolerance is 1e-4
        assert np.median(errs[32]) < 1e-3
    with pytest.raises(RuntimeError):
        sim.set_precision(16)


def test_fp64_state_persists_between_launches_and_follows_api_edits():
    """precision 64 with ONE step per launch (the reference's Simulation.step() usage): the library keeps full-precision
    records between launches, so 600 single-step launches shadow the oracle as well as one fused launch does; values written
    through the API in between (setters, a direct write to qpos, a masked reset) are picked up."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.anatomy import ActuatorType
    from flygym_b200.actions import cpg_table
    from oracle.oracle import Oracle
    m = NMFModel.bench(True)
    a = dict(m.arrays); opt = a["opt"].copy(); opt[5] = 1e-16; a["opt"] = opt
    mt = NMFModel(a, m.names, m.meta)
    n, T = 3, 600
    tab = cpg_table(m, n, T)
    q0 = np.tile(m.arrays["key_qpos"], (n, 1)).astype(np.float32); q0[:, 2] = -0.17
    sim = B200Simulation(m, n_worlds=n, outputs=False)
    sim.set_precision(64)
    sim.qpos.copy_(torch.from_numpy(q0)); sim.set_leg_adhesion_states("nmf", np.ones(6, dtype=bool))
    oracles = []
    for k in range(n):
        o = Oracle(mt); o.reset(); o.qpos[:] = q0[k].astype(np.float64); o.ctrl[42:] = 1.0
        oracles.append(o)
    tabd = torch.from_numpy(tab).cuda()
    kick = np.float32(0.25)
    for t in range(T):
        if t == 300:                                   # edits through the API between two launches
            sim.qpos[1, 0] += float(kick)              # direct write to the state tensor
            oracles[1].qpos[0] = np.float64(np.float32(oracles[1].qpos[0]) + kick)
            mask = torch.tensor([False, False, True]); sim.reset(mask)      # masked reset of fly 2
            oracles[2].reset()
        sim.set_actuator_inputs("nmf", ActuatorType.POSITION, tabd[:, t])
        sim.step()
        for k, o in enumerate(oracles):
            o.ctrl[:42] = tab[k, t].astype(np.float64)
            if t == 300 and k == 2:
                pass                                    # reset restored the keyframe ctrl; the setter above then wrote the targets
            o.step()
    got = sim.qpos.cpu().numpy().astype(np.float64)
    errs = [float(np.abs(got[k] - oracles[k].qpos).max() / np.abs(oracles[k].qpos).max()) for k in range(n)]
    print("fp64, one step per launch:", ["%.1e" % e for e in errs])
    # fly 1's kick was applied to the float32 image of its state (that is what an API edit is), hence ~1e-7 there
    assert errs[0] < 2e-7 and errs[1] < 5e-6 and errs[2] < 2e-7, errs
    assert abs(sim.time - (T * 1e-4)) < 1e-6


@pytest.mark.parametrize("variant", ["flat", "mesh", "blocks"])
def test_noslip_kernels_match_the_oracle(variant):
    """`noslip_iterations: 5` (mujoco_globals.yaml:15): the reference's CPU `Simulation` runs MuJoCo's noslip post-solver after
    Newton, its `GPUSimulation` strips it (warp/simulation.py:427-448).  A model baked with noslip_iterations = 5 selects the
    NOSLIP kernel instantiations; they must follow the oracle's noslip ([PRIOR] mj_solNoSlip) on a settled fly pushed sideways
    and then walking: f64 within float32 resolution over 300 steps, f32 within 1e-4 (every world); BASELINE config 1 (1000 steps,
    hold-neutral) is reported for both settings."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    from oracle.oracle import Oracle
    base = {"flat": NMFModel.bench(True), "mesh": NMFModel.bench(False), "blocks": NMFModel.bench(True, terrain="blocks")}[variant]
    m = base.with_options(noslip_iterations=5)
    o0 = Oracle(base); o0.reset(); o0.qpos[2] = -0.17 if variant != "blocks" else -0.15; o0.ctrl[42:] = 1.0; o0.step(1200)
    q, v, w = o0.qpos.copy(), o0.qvel.copy(), o0.get("qacc_warmstart").copy()
    v[0:2] += [3.0, -2.0]
    T = 300
    tab = cpg_table(m, 2, T)
    o = [Oracle(m) for _ in range(2)]
    for i, oi in enumerate(o):
        oi.reset(); oi.qpos[:] = q; oi.qvel[:] = v; oi.get("qacc_warmstart")[:] = w; oi.ctrl[42:] = 1.0
        oi.step_table(tab[i].astype(np.float64))
    plain = Oracle(base); plain.reset(); plain.qpos[:] = q; plain.qvel[:] = v; plain.get("qacc_warmstart")[:] = w; plain.ctrl[42:] = 1.0
    plain.step_table(tab[0].astype(np.float64))
    for prec in (64, 32):
        sim = B200Simulation(m, n_worlds=2)
        sim.set_precision(prec)
        f32 = lambda a: torch.as_tensor(np.tile(a, (2, 1)), dtype=torch.float32)
        sim.qpos.copy_(f32(q)); sim.qvel.copy_(f32(v)); sim.qacc_warmstart.copy_(f32(w)); sim.ctrl[:, 42:] = 1.0
        sim.step(T, torch.from_numpy(tab).cuda(), 0)
        got = sim.qpos.cpu().numpy().astype(np.float64)
        err = max(np.abs(got[i] - o[i].qpos).max() / np.abs(o[i].qpos).max() for i in range(2))
        gap = np.abs(plain.qpos - o[0].qpos).max() / np.abs(o[0].qpos).max()
        print(f"noslip {variant} f{prec}: qpos rel Linf vs oracle {err:.1e}; noslip-vs-plain oracle {gap:.1e}")
        assert err < (5e-7 if prec == 64 else 1e-4) and gap > 1e-4
        so = np.stack([oi.get("sensordata") for oi in o])
        assert np.abs(sim.sensordata.cpu().numpy() - so).max() < (1e-5 if prec == 64 else 1e-2) * max(1.0, np.abs(so).max())
        assert int(sim.status.abs().max()) == 0
    if variant == "flat":       # BASELINE config 1 under both solver settings (the north star's 1e-4 after 1000 steps)
        key = base.arrays["key_qpos"]
        for mm, tag in ((base, "noslip 0 (GPUSimulation semantics)"), (m, "noslip 5 (Simulation semantics)")):
            oc = Oracle(mm); oc.reset(); oc.step(1000)
            for prec in (32, 64):
                sim = B200Simulation(mm, n_worlds=1, outputs=False); sim.set_precision(prec)
                sim.step(1000)
                e = np.abs(sim.qpos[0].cpu().numpy() - oc.qpos).max() / np.abs(oc.qpos).max()
                print(f"config 1, {tag}, f{prec}: qpos rel Linf after 1000 steps = {e:.1e}")
                assert e < 1e-4


def test_noslip_runs_in_float32_in_every_world():
    """float32 noslip instantiations exist for every world of the star kernels (flat, mesh hulls, terrain, tethered) and on the
    general-topology kernels (contact rows and the weld of a tethered full skeleton): nothing is refused any more."""
    from flygym_b200 import B200Simulation, NMFModel
    for m in (NMFModel.bench(False), NMFModel.tethered(), NMFModel.bench(True, terrain="gapped"),
              NMFModel.bench(True, joint_preset="all_biological"), NMFModel.tethered(joint_preset="all_biological")):
        sim = B200Simulation(m.with_options(noslip_iterations=5), n_worlds=3)
        sim.step(5)
        assert bool(np.isfinite(sim.qpos.cpu().numpy()).all()) and int(sim.status.abs().max()) == 0


@pytest.mark.parametrize("skeleton", ["legs_only", "all_biological"])
def test_noslip_with_many_contacts(skeleton):
    """A mesh-hull fly dropped flat onto the ground has 48 contacts at once with multiccd (4 per hull): the noslip pass takes up to 48
    (star and general-topology kernels; B_tt is 96 x 96 in dynamic shared memory) and must follow the oracle there too, without
    raising NMF_ST_NOSLIP_SKIP.  f64 build, 30 steps from a 0.02 mm penetration."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from oracle.oracle import Oracle
    m = NMFModel.bench(False, joint_preset=skeleton).with_options(noslip_iterations=5)
    q0 = m.arrays["key_qpos"].astype(np.float32).astype(np.float64); q0[2] = np.float32(-0.17)
    o = Oracle(m); o.reset(); o.qpos[:] = q0; o.ctrl[m.dim("nu_pos"):] = 1.0
    o.step(1); ncon1 = o.dim("ncon"); o.step(29)
    for prec, tol in ((64, 5e-7), (32, 2e-5)):
        sim = B200Simulation(m, n_worlds=2); sim.set_precision(prec)
        sim.qpos.copy_(torch.as_tensor(np.tile(q0, (2, 1)), dtype=torch.float32)); sim.ctrl[:, m.dim("nu_pos"):] = 1.0
        sim.step(30)
        err = np.abs(sim.qpos[0].cpu().numpy() - o.qpos).max() / np.abs(o.qpos).max()
        print(f"{skeleton} f{prec}: {ncon1} contacts in the first step, qpos rel Linf after 30 steps {err:.1e}, status {int(sim.status.abs().max())}")
        assert ncon1 > 24 and err < tol and int(sim.status.abs().max()) == 0


def test_tethered_world_with_noslip_matches_the_oracle():
    """The world of the reference's own `tests/core/test_simulation.py` fixtures (TetheredWorld) under the CPU `Simulation`
    semantics (noslip_iterations = 5): the six weld rows are equality rows, which MuJoCo's noslip sweeps unclamped -- the soft
    weld becomes nearly hard.  f64 kernel vs the oracle over 300 CPG steps, and far from the plain (noslip 0) trajectory."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    from oracle.oracle import Oracle
    base = NMFModel.tethered(); m = base.with_options(noslip_iterations=5)
    tab = cpg_table(m, 2, 300)
    os_ = []
    for i in range(2):
        o = Oracle(m); o.reset(); o.step_table(tab[i].astype(np.float64)); os_.append(o)
    for prec, tol in ((64, 5e-7), (32, 1e-4)):
        sim = B200Simulation(m, n_worlds=2); sim.set_precision(prec)
        sim.step(300, torch.from_numpy(tab).cuda(), 0)
        got = sim.qpos.cpu().numpy().astype(np.float64)
        err = max(np.abs(got[i] - os_[i].qpos).max() / np.abs(os_[i].qpos).max() for i in range(2))
        print(f"tethered noslip f{prec}: qpos rel Linf vs oracle after 300 steps {err:.1e}")
        assert err < tol
    plain = Oracle(base); plain.reset(); plain.step_table(tab[0].astype(np.float64))
    o = Oracle(m); o.reset(); o.step_table(tab[0].astype(np.float64))
    assert np.abs(plain.qpos - o.qpos).max() > 1e-4 and int(sim.status.abs().max()) == 0


def test_mujoco_golden_on_the_gpu_if_present():
    """Twin of tests/test_cpu_suite.py::test_mujoco_golden_if_present for the CUDA path: where tests/golden/mujoco_golden.npz exists
    (tools/dump_mujoco_golden.py, needs real MuJoCo + the reference package), the kernels -- f64 against the north star's 1e-4,
    f32 against the bands of this file -- replay every recorded scenario for noslip 0 and 5 and are compared with MuJoCo's own
    qpos / qvel / contact sensors.  Absent here (MuJoCo is not installable): parity vs MuJoCo is unpinned."""
    import sys
    from pathlib import Path
    import torch
    root = Path(__file__).resolve().parent.parent
    path = root / "tests" / "golden" / "mujoco_golden.npz"
    if not path.exists():
        pytest.skip("parity vs MuJoCo: not run (golden file absent; see tools/dump_mujoco_golden.py)")
    from flygym_b200 import B200Simulation, NMFModel
    from test_cpu_suite import golden_scenarios
    z = np.load(path, allow_pickle=False)
    for tag, simplify in (("capsule", True), ("mesh", False)):
        base = NMFModel.bench(simplify)
        cps, scen = golden_scenarios(base, z, tag)
        for noslip in sorted({int(k.split("/")[1][6:]) for k in z.files if k.startswith(f"{tag}/noslip")}):
            model = base.with_options(noslip_iterations=noslip)
            for prec in (64, 32):
                if prec == 32 and noslip > 0 and not simplify:
                    continue                                   # float32 noslip is built for the flat capsule world only
                n = len(scen)
                sim = B200Simulation(model, n_worlds=n); sim.set_precision(prec)
                tab = np.zeros((n, cps[-1], 48), np.float32)
                for i, (sname, z0, table, adh) in enumerate(scen):
                    if z0 is not None:
                        sim.qpos[i, 2] = z0
                    tab[i, :, :42] = table; tab[i, :, 42:] = max(adh, 0.0)
                tabd = torch.from_numpy(tab).cuda(); done = 0
                for k, cp in enumerate(cps):
                    sim.step(cp - done, tabd, done); done = cp
                    q = sim.qpos.cpu().numpy().astype(np.float64); v = sim.qvel.cpu().numpy().astype(np.float64)
                    errs = []
                    for i, (sname, *_rest) in enumerate(scen):
                        rq, rv = z[f"{tag}/noslip{noslip}/{sname}/qpos"][k], z[f"{tag}/noslip{noslip}/{sname}/qvel"][k]
                        errs.append(np.abs(q[i] - rq).max() / np.abs(rq).max())
                        if prec == 64:
                            assert errs[-1] < 1e-4, (tag, noslip, sname, cp, errs[-1])             # the north star's tolerance
                            assert np.abs(v[i] - rv).max() / max(1.0, np.abs(rv).max()) < 1e-3, (tag, noslip, sname, cp, "qvel")
                    if prec == 32:                             # float32: every fly to 100 steps, median beyond (contact chaos, see the tests above)
                        assert (max(errs) < 1e-4) if cp <= 100 else (np.median(errs) < 1e-3), (tag, noslip, cp, errs)


@pytest.mark.parametrize("world", ["flat", "mesh", "blocks"])
def test_flies_per_block_is_bit_identical(world):
    """A block steps 1, 2, 4 or 8 flies side by side (lockstep solver passes, step_block<WORLD, FPB>): ragged batch sizes (empty
    slots in the last block), the work queue on and off, and every FPB must give the same bits -- state records, segment poses,
    sensors -- as one fly per block.  Also the edge sizes n = 1 and n = 13."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    m = {"flat": lambda: NMFModel.bench(True), "mesh": lambda: NMFModel.bench(False), "blocks": lambda: NMFModel.bench(True, terrain="blocks")}[world]()
    for n, steps in ((2500, 30), (13, 30), (1, 30)):                  # 2500 = 312 blocks of 8 + a block with 4 empty slots; > resident slots
        table = torch.from_numpy(cpg_table(m, n, 64)).cuda()
        ref = None
        for fpb, sub in ((1, 0), (8, -1), (8, 0), (4, 7), (2, -1), (0, -1)):
            sim = B200Simulation(m, n_worlds=n, outputs=True)
            sim.set_flies_per_block(fpb); sim.set_schedule(sub)
            sim.qpos[:, 2] = -0.15
            sim.qpos[:, 0] += torch.linspace(0, 1, n, device="cuda")
            sim.ctrl[:, 42:] = 1.0
            sim.step(steps, table, 5)
            sim.step(1, table, 5 + steps)
            torch.cuda.synchronize()
            got = (sim.state.clone(), sim.seg_xpos.clone(), sim.sensordata.clone(), sim.act_force.clone())
            if ref is None:
                ref = got
                assert torch.isfinite(ref[0]).all()
            else:
                for a, b in zip(ref, got):
                    assert torch.equal(a, b), (world, n, fpb, sub)
