"""The reference's own Simulation test-suite (reference tests/core/test_simulation.py, whose fixture world is a
TetheredWorld with the LEGS_ONLY / YAW_PITCH_ROLL / kp = 50 / adhesion fly at spawn (0, 0, 1.5): tests/conftest.py:74-152),
restated against B200Simulation.  Batched getters return (n_worlds, ...) like the reference's GPUSimulation, so the
single-world assertions are applied to every world."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 3


@pytest.fixture(scope="module")
def model():
    from flygym_b200 import NMFModel
    return NMFModel.tethered(spawn_position=(0.0, 0.0, 1.5))


@pytest.fixture(scope="module")
def simulation(model):
    from flygym_b200 import B200Simulation
    sim = B200Simulation(model, n_worlds=N, fly_name="sim_fly")
    sim.reset()
    return sim


class TestSimulationConstruction:
    def test_construction_succeeds(self, simulation):
        assert simulation is not None

    def test_world_empty_raises(self, model):
        from flygym_b200 import B200Simulation
        from flygym_b200.simulation import WorldView
        w = WorldView(model); w.fly_lookup.clear()
        with pytest.raises(ValueError):
            B200Simulation(w)

    def test_time_starts_at_zero_after_reset(self, simulation):
        simulation.reset()
        assert simulation.time == pytest.approx(0.0)


class TestSimulationStep:
    def test_step_advances_time(self, simulation):
        simulation.reset()
        simulation.step()
        assert simulation.time == pytest.approx(simulation.timestep, rel=1e-6)

    def test_multiple_steps(self, simulation):
        simulation.reset()
        for _ in range(10):
            simulation.step()
        assert simulation.time == pytest.approx(10 * simulation.timestep, rel=1e-6)

    def test_reset_resets_time(self, simulation):
        simulation.reset()
        for _ in range(5):
            simulation.step()
        simulation.reset()
        assert simulation.time == pytest.approx(0.0)


class TestGetters:
    def test_joint_angles_length_and_neutral_pose(self, simulation, model):
        simulation.reset()
        angles = simulation.get_joint_angles("sim_fly").cpu().numpy()
        order = simulation.world.fly_lookup["sim_fly"].get_jointdofs_order()
        assert angles.shape == (N, len(order)) and len(order) == 66
        neutral = model.arrays["dof_springref"][6:]                # springref = neutral angle (fly.py:285-295)
        assert np.abs(angles - neutral[None]).max() < 0.2

    def test_velocities_near_zero_at_reset(self, simulation):
        simulation.reset()
        vels = simulation.get_joint_velocities("sim_fly").cpu().numpy()
        assert vels.shape == (N, 66)
        np.testing.assert_allclose(vels, 0.0, atol=1e-8)

    def test_body_positions_and_unit_quaternions(self, simulation):
        simulation.reset()
        simulation.step()
        pos = simulation.get_body_positions("sim_fly").cpu().numpy()
        quat = simulation.get_body_rotations("sim_fly").cpu().numpy()
        nseg = len(simulation.world.fly_lookup["sim_fly"].get_bodysegs_order())
        assert pos.shape == (N, nseg, 3) and quat.shape == (N, nseg, 4) and nseg == 69
        np.testing.assert_allclose(np.linalg.norm(quat, axis=-1), 1.0, atol=1e-5)


class TestActuatorIO:
    def test_set_and_get_actuator_forces(self, simulation):
        from flygym_b200.anatomy import ActuatorType
        simulation.reset()
        n_act = len(simulation.world.fly_lookup["sim_fly"].get_actuated_jointdofs_order(ActuatorType.POSITION))
        simulation.set_actuator_inputs("sim_fly", ActuatorType.POSITION, np.zeros(n_act))
        simulation.step()
        forces = simulation.get_actuator_forces("sim_fly", ActuatorType.POSITION).cpu().numpy()
        assert forces.shape == (N, n_act) and n_act == 42 and np.isfinite(forces).all() and np.abs(forces).max() > 0

    def test_set_actuator_inputs_wrong_length_raises(self, simulation):
        from flygym_b200.anatomy import ActuatorType
        simulation.reset()
        with pytest.raises(ValueError):
            simulation.set_actuator_inputs("sim_fly", ActuatorType.POSITION, np.zeros(42 + 5))


class TestLegAdhesion:
    def test_set_all_adhesion_on_and_off(self, simulation):
        simulation.reset()
        simulation.set_leg_adhesion_states("sim_fly", np.ones(6, dtype=bool))
        simulation.step()
        simulation.set_leg_adhesion_states("sim_fly", np.zeros(6, dtype=bool))
        simulation.step()
        assert np.isfinite(simulation.state.cpu().numpy()).all()

    def test_set_adhesion_wrong_length_raises(self, simulation):
        simulation.reset()
        with pytest.raises(ValueError):
            simulation.set_leg_adhesion_states("sim_fly", np.ones(5, dtype=bool))


class TestWarmupAndProfiling:
    def test_warmup_advances_time(self, simulation):
        simulation.reset()
        simulation.warmup(duration_s=0.001)
        assert simulation.time > 0.0

    def test_warmup_zero_duration_does_not_change_time(self, simulation):
        simulation.reset()
        simulation.warmup(duration_s=0.0)
        assert simulation.time == pytest.approx(0.0)

    def test_step_with_profile(self, simulation, capsys):
        simulation.reset()
        simulation.step_with_profile()
        assert simulation.time == pytest.approx(simulation.timestep, rel=1e-6)
        assert simulation._curr_step == 1 and simulation._total_physics_time_ns > 0
        simulation.print_performance_report()
        assert "physics" in capsys.readouterr().out
        simulation.reset()
        assert simulation._curr_step == 0 and simulation._total_physics_time_ns == 0


def test_tethered_parity_with_oracle(model):
    """Legs driven by the CPG targets while the thorax hangs on the weld: no contacts, so the fp32 trajectory must stay
    within the north_star tolerance (1e-4 rel) of the fp64 oracle over 1000 steps."""
    import torch
    from flygym_b200 import B200Simulation
    from flygym_b200.actions import cpg_table
    from oracle.oracle import Oracle
    n, T = 4, 1000
    tab = cpg_table(model, n, T)
    sim = B200Simulation(model, n_worlds=n, outputs=True)
    tabd = torch.from_numpy(tab).cuda()
    got, done = {}, 0
    for cp in (1, 10, 100, 1000):
        sim.step(cp - done, tabd, done); done = cp
        got[cp] = sim.qpos.cpu().numpy().astype(np.float64)
    errs = {cp: [] for cp in got}
    for k in range(n):
        o = Oracle(model); o.reset(); done = 0
        for cp in (1, 10, 100, 1000):
            o.step_table(tab[k, done:cp].astype(np.float64)); done = cp
            errs[cp].append(float(np.abs(got[cp][k] - o.qpos).max() / np.abs(o.qpos).max()))
    print("tethered qpos rel Linf:", {k: ["%.1e" % e for e in v] for k, v in errs.items()})
    assert max(errs[1]) < 1e-4 and max(errs[10]) < 1e-4 and max(errs[100]) < 1e-4 and max(errs[1000]) < 1e-4
    thorax = sim.model.names["segments"].index("c_thorax")
    assert (sim.get_body_positions("nmf")[:, thorax].cpu() - torch.tensor([0.0, 0.0, -1.5])).abs().max() < 5e-3
    found = sim.get_ground_contact_info("nmf")[0]
    assert float(found.abs().sum()) == 0.0        # no ground in the tethered world


def test_joint_and_contact_presets_on_gpu():
    """JointPreset.LEGS_ACTIVE_ONLY (42 hinge DoFs; reference tests/core/test_compose.py:74-76,178-183 count joints = len(iter_jointdofs))
    and the ContactBodiesPreset variants through the public API, each against the oracle on the same blob."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    from oracle.oracle import Oracle
    base = NMFModel.bench(True)
    variants = {
        "legs_active_only": NMFModel.bench(True, joint_preset="legs_active_only"),
        "contacts_legs_only": base.with_contact_bodies("legs_only"),
        "contacts_tibia_tarsus": base.with_contact_bodies("tibia_tarsus_only"),
        "kp150": base.with_actuator_gains(kp=150.0),
    }
    for name, model in variants.items():
        n, T = 3, 200
        tab = cpg_table(model, n, T)
        sim = B200Simulation(model, n_worlds=n)
        sim.qpos[:, 2] = -0.17
        sim.set_leg_adhesion_states("nmf", np.ones(6, dtype=bool))
        ndof = len(sim.world.fly_lookup["nmf"].get_jointdofs_order())
        assert sim.get_joint_angles("nmf").shape == (n, ndof) and ndof == (42 if name == "legs_active_only" else 66)
        sim.step(T, torch.from_numpy(tab).cuda(), 0)
        got = sim.qpos.cpu().numpy().astype(np.float64)
        errs = []
        for k in range(n):
            o = Oracle(model); o.reset(); o.qpos[2] = -0.17; o.ctrl[42:] = 1.0
            o.step_table(tab[k].astype(np.float64))
            errs.append(float(np.abs(got[k] - o.qpos).max() / np.abs(o.qpos).max()))
        print(name, ["%.1e" % e for e in errs])
        assert np.median(errs) < 1e-4 and max(errs) < 1e-2, (name, errs)
        ang = sim.get_joint_angles("nmf").cpu().numpy()
        ex = model.exposed_hinge_dofs()
        assert np.allclose(ang, got[:, 7 + ex], atol=1e-6)
        if name == "legs_active_only":
            locked = np.delete(got[:, 7:], ex, axis=1)
            assert np.abs(locked).max() < 1e-6


@pytest.fixture(scope="module")
def sim():
    import torch
    from flygym_b200 import B200Simulation
    s = B200Simulation(None, n_worlds=7)
    g = torch.Generator(device="cuda").manual_seed(1)
    s.state.copy_(torch.randn(s.state.shape, device="cuda", generator=g))
    return s


class TestGatherScatterKernels:
    """The reference checks its indexed gather / scatter kernels against numpy fancy indexing (tests/warp/test_utils.py:26-150:
    correct columns, a single column, all columns, scatter preserves the other values); same checks through the C ABI
    (nmf_gather_state / nmf_scatter_ctrl, which stand behind the getters / setters)."""

    def _gather(self, sim, off, cols):
        import ctypes, torch
        c = torch.tensor(cols, dtype=torch.int32, device="cuda")
        dst = torch.full((sim.n_worlds, len(cols)), float("nan"), device="cuda")
        rc = sim._lib.nmf_gather_state(sim._h, off, ctypes.c_void_p(c.data_ptr()), len(cols), ctypes.c_void_p(dst.data_ptr()), sim._stream())
        assert rc == 0
        return dst.cpu().numpy()

    @pytest.mark.parametrize("cols", [[3, 0, 17, 65, 9], [41], list(range(72))])
    def test_gather_matches_numpy_indexing(self, sim, cols):
        host = sim.state.cpu().numpy()
        off = sim.info.off_qvel
        assert np.array_equal(self._gather(sim, off, cols), host[:, off + np.array(cols)])

    def test_scatter_writes_only_the_selected_columns(self, sim):
        import ctypes, torch
        before = sim.state.cpu().numpy().copy()
        cols = [5, 0, 47, 20]
        src = torch.arange(sim.n_worlds * len(cols), dtype=torch.float32, device="cuda").reshape(sim.n_worlds, len(cols)) + 100
        c = torch.tensor(cols, dtype=torch.int32, device="cuda")
        rc = sim._lib.nmf_scatter_ctrl(sim._h, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(c.data_ptr()), len(cols), sim._stream())
        assert rc == 0
        after = sim.state.cpu().numpy()
        off = sim.info.off_ctrl
        assert np.array_equal(after[:, off + np.array(cols)], src.cpu().numpy())
        mask = np.ones(after.shape[1], bool); mask[off + np.array(cols)] = False
        assert np.array_equal(after[:, mask], before[:, mask])
        # bad arguments are refused with a status, not a fault
        assert sim._lib.nmf_gather_state(sim._h, 10_000, ctypes.c_void_p(c.data_ptr()), 4, ctypes.c_void_p(src.data_ptr()), sim._stream()) == -1
        assert sim._lib.nmf_scatter_ctrl(sim._h, None, ctypes.c_void_p(c.data_ptr()), 4, sim._stream()) == -1
