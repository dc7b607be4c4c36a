"""General-topology kernels (csrc/nmf_tree.cuh) on the B200, through the C ABI: JointPreset.ALL_BIOLOGICAL / ALL_POSSIBLE skeletons
(reference src/flygym/anatomy.py:388-460), ContactBodiesPreset.ALL (anatomy.py:519-526), mesh hulls with multiccd, box-column terrain.
Same comparison as tests/test_gpu_outputs.py: qpos, qvel, actuator forces, segment poses and the per-leg contact sensor against the fp64
oracle after 1 and 100 steps (a standing fly + 4 CPG walkers per world).  PARITY UNPINNED against real MuJoCo."""
import os

import numpy as np
import pytest

from test_gpu_outputs import CHECK, run_world

pytestmark = pytest.mark.gpu

TREE_WORLDS = ["allbio_capsule", "allbio_mesh", "allbio_blocks", "allpossible_allcontacts", "legsonly_allcontacts"]


@pytest.mark.parametrize("wname", TREE_WORLDS)
def test_every_output_of_the_tree_kernels_matches_the_oracle(wname):
    errs = run_world(wname)
    e1, e100 = errs[1], errs[100]
    print(wname, {cp: {k: "%.1e" % max(v) for k, v in e.items()} for cp, e in errs.items()})
    assert max(e1["ncon"]) >= 6
    assert max(e1["qpos_rel"]) < 1e-6 and max(e1["qvel_rel"]) < 5e-4 and max(e1["actf_abs"]) < 1e-4
    assert max(e1["xpos_abs"]) < 2e-6 and max(e1["xquat_abs"]) < 2e-6
    assert max(e1["found_mismatch"]) == 0 and max(e1["frame_abs"]) == 0
    assert max(e1["force_rel"]) < 5e-5 and max(e1["torque_rel"]) < 5e-5 and max(e1["pos_abs"]) < 1e-5
    worst = max if wname in ("allbio_capsule", "legsonly_allcontacts") else np.median
    assert worst(e100["qpos_rel"]) < 1e-4 and worst(e100["qvel_rel"]) < 1e-3 and worst(e100["actf_abs"]) < 2e-3
    assert worst(e100["xpos_abs"]) < 1e-4 and worst(e100["xquat_abs"]) < 1e-4 and worst(e100["found_mismatch"]) == 0
    assert worst(e100["force_rel"]) < 1e-3 and worst(e100["torque_rel"]) < 1e-3 and worst(e100["pos_abs"]) < 1e-3
    for k in ("qpos_rel", "qvel_rel", "force_rel"):
        assert e100[k][0] < 1e-3, k


def test_tethered_world_with_the_full_skeleton():
    """TetheredWorld (world.py:334-366) + ALL_BIOLOGICAL: the weld rows on the root body of the general-topology kernels."""
    errs = run_world("allbio_tethered")
    e1, e100 = errs[1], errs[100]
    print({cp: {k: "%.1e" % max(v) for k, v in e.items()} for cp, e in errs.items()})
    assert max(e1["qpos_rel"]) < 1e-4 and max(e1["qvel_rel"]) < 5e-4 and max(e1["actf_abs"]) < 1e-4      # the weld snaps the thorax by ~1 mm in the first step
    assert max(e1["xpos_abs"]) < 2e-6 and max(e1["xquat_abs"]) < 2e-6 and max(e1["found_mismatch"]) == 0
    assert max(e100["qpos_rel"]) < 1e-4 and max(e100["qvel_rel"]) < 1e-3 and max(e100["xpos_abs"]) < 5e-4     # measured 3.8e-5 / 5e-5 / 1.5e-4 (float32 against the stiff weld)


@pytest.mark.parametrize("wname", ["allbio_capsule", "allpossible_allcontacts"])
def test_tree_kernels_in_double_precision(wname):
    errs = run_world(wname, precision=64)
    print({cp: {k: "%.1e" % max(v) for k, v in e.items()} for cp, e in errs.items()})
    for cp in CHECK:
        e = errs[cp]
        # float32 rounding of the buffers + the oracle's own Newton tolerance (1e-8 scaled: it stops a hair before the exact minimiser,
        # measured qvel 8.7e-6 / force 3.9e-5 on the 210-DoF model with two contacts left)
        assert max(e["qpos_rel"]) < 5e-7 and max(e["qvel_rel"]) < 3e-5
        assert max(e["actf_abs"]) < 2e-5 and max(e["xpos_abs"]) < 5e-6 and max(e["xquat_abs"]) < 1e-6
        assert max(e["found_mismatch"]) == 0 and max(e["force_rel"]) < 2e-4 and max(e["torque_rel"]) < 2e-4 and max(e["pos_abs"]) < 2e-6


@pytest.mark.parametrize("wname", ["allbio_capsule_noslip", "allbio_mesh_noslip"])
def test_noslip_on_the_tree_kernels(wname):
    """`noslip_iterations: 5` on the general-topology kernels (ALL_BIOLOGICAL skeleton; capsule geoms and mesh hulls with multiccd):
    every output against the oracle's noslip after 1 and 100 steps -- double precision to the float32 resolution of the buffers,
    float32 within the bands of the plain solve."""
    errs = run_world(wname, precision=64, settle=800)      # (a fly dropped from the keyframe height lands on more than 24 hull contacts at once)
    print(wname, "f64", {cp: {k: "%.1e" % max(v) for k, v in e.items()} for cp, e in errs.items()})
    for cp in CHECK:
        e = errs[cp]
        assert max(e["qpos_rel"]) < 5e-7 and max(e["qvel_rel"]) < 3e-5 and max(e["actf_abs"]) < 2e-5 and max(e["xpos_abs"]) < 5e-6
        assert max(e["found_mismatch"]) == 0 and max(e["force_rel"]) < 2e-4 and max(e["torque_rel"]) < 2e-4 and max(e["pos_abs"]) < 2e-6
        assert max(e["status"]) == 0                      # in particular: the post-solver was never skipped
    errs = run_world(wname, precision=32, settle=800)
    print(wname, "f32", {cp: {k: "%.1e" % max(v) for k, v in e.items()} for cp, e in errs.items()})
    e1, e100 = errs[1], errs[100]
    assert max(e1["qpos_rel"]) < 1e-6 and max(e1["qvel_rel"]) < 5e-4 and max(e1["actf_abs"]) < 1e-4 and max(e1["found_mismatch"]) == 0
    assert max(e1["force_rel"]) < 1e-4 and max(e1["torque_rel"]) < 1e-4
    assert np.median(e100["qpos_rel"]) < 1e-4 and np.median(e100["qvel_rel"]) < 1e-3 and np.median(e100["force_rel"]) < 1e-3
    assert e100["qpos_rel"][0] < 1e-4 and e100["force_rel"][0] < 1e-3          # the standing fly


def test_noslip_on_the_tethered_full_skeleton():
    """The weld rows of the tethered ALL_BIOLOGICAL world are equality rows, which noslip sweeps unclamped (as the star kernels do for the
    LEGS_ONLY skeleton of the reference's CPU fixtures): double precision vs the oracle's noslip, float32 within the tethered band."""
    errs = run_world("allbio_tethered_noslip", precision=64)
    print({cp: {k: "%.1e" % max(v) for k, v in e.items()} for cp, e in errs.items()})
    for cp in CHECK:
        e = errs[cp]
        assert max(e["qpos_rel"]) < 5e-7 and max(e["qvel_rel"]) < 3e-5 and max(e["xpos_abs"]) < 5e-6 and max(e["status"]) == 0
    errs = run_world("allbio_tethered_noslip", precision=32)
    print({cp: {k: "%.1e" % max(v) for k, v in e.items()} for cp, e in errs.items()})
    assert max(errs[1]["qpos_rel"]) < 1e-4 and max(errs[100]["qpos_rel"]) < 1e-4 and max(errs[100]["xpos_abs"]) < 5e-4


def test_tree_and_star_kernels_agree_on_the_benchmark_model():
    """NMF_FORCE_TREE routes the benchmark skeleton through the general kernels: both families must walk the same 64 flies (300
    CPG steps, float32) to float32 rounding, and every reference-facing call works on both layouts."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table
    m = NMFModel.bench(True)
    n, T = 64, 300
    tab = torch.zeros((n, T, 48), dtype=torch.float32); tab[:, :, :42] = torch.as_tensor(cpg_table(m, n, T)); tab[:, :, 42:] = 1.0
    tab = tab.cuda()
    sims = []
    for force in ("0", "1"):
        os.environ["NMF_FORCE_TREE"] = force
        try:
            sim = B200Simulation(m, n_worlds=n)
        finally:
            os.environ.pop("NMF_FORCE_TREE", None)
        q = sim.qpos.clone(); q[:, 2] = -0.17; sim.qpos.copy_(q)
        sim.step(100, tab, 0)
        sims.append(sim)
    star, tree = sims
    assert star.info.state_stride == 304 and tree.info.state_stride != 304
    d100 = (star.qpos - tree.qpos).abs().max(dim=1).values.cpu().numpy()
    print("tree vs star after 100 steps: median %.1e max %.1e" % (np.median(d100), d100.max()))
    assert np.median(d100) < 5e-6 and np.percentile(d100, 90) < 1e-4 and d100.max() < 5e-2   # a walker may resolve a contact switch one step apart (both are float32)
    assert (star.get_joint_angles("nmf") - tree.get_joint_angles("nmf")).abs().median() < 5e-6
    fs, ft = star.get_ground_contact_info("nmf")[1], tree.get_ground_contact_info("nmf")[1]
    assert (fs - ft).abs().median() < 1e-3


def test_reference_api_on_the_all_biological_skeleton():
    """Counts and orders of the reference's own compose tests for the full skeleton (tests/core/test_compose.py:74-76,178-188:
    joints = len(skeleton.iter_jointdofs()), 69 body segments), reset / masked reset, setters, getters, step_host, status."""
    import torch
    from flygym_b200 import ActuatorType, B200Simulation, NMFModel
    from flygym_b200 import anatomy as A
    m = NMFModel.bench(joint_preset="all_biological")
    sim = B200Simulation(m, n_worlds=5)
    dofs = sim.world.fly_lookup["nmf"].get_jointdofs_order()
    assert len(dofs) == 126 == len(A.jointdofs_order("all_biological")) and sim.info.nv == 132 and sim.info.nq == 133
    assert sim.get_joint_angles("nmf").shape == (5, 126) and sim.get_joint_velocities("nmf").shape == (5, 126)
    assert float(sim.get_joint_velocities("nmf").abs().max()) == 0.0
    neutral = torch.as_tensor(m.arrays["key_qpos"][7:], dtype=torch.float32).cuda()
    assert torch.equal(sim.get_joint_angles("nmf")[2], neutral)
    sim.set_actuator_inputs("nmf", ActuatorType.POSITION, np.zeros(42, np.float32))
    sim.set_leg_adhesion_states("nmf", np.ones((5, 6), np.float32))
    assert float(sim.ctrl[:, :42].abs().max()) == 0.0 and float(sim.ctrl[:, 42:].min()) == 1.0
    with pytest.raises(ValueError):
        sim.set_actuator_inputs("nmf", ActuatorType.POSITION, np.zeros(41, np.float32))
    sim.step(25)
    assert abs(sim.time - 25 * m.timestep) < 1e-9 and int(sim.status.abs().max()) == 0
    assert sim.get_body_positions("nmf").shape == (5, 69, 3) and sim.get_body_rotations("nmf").shape == (5, 69, 4)
    assert torch.allclose(sim.get_body_rotations("nmf").norm(dim=-1), torch.ones(5, 69, device="cuda"), atol=1e-5)
    assert sim.get_actuator_forces("nmf", ActuatorType.POSITION).shape == (5, 42)
    moved = sim.qpos.clone()
    sim.reset(mask=[True, False, False, False, True])
    assert torch.equal(sim.qpos[1:4], moved[1:4]) and torch.equal(sim.qpos[0, 7:], neutral) and float(sim.state[4, sim.info.off_time]) == 0.0
    # host-buffer call: one more step for every fly, qpos comes back packed
    act = np.tile(m.arrays["key_ctrl"][:42].astype(np.float32), (5, 1)); out = np.zeros((5, 133), np.float32)
    before = sim.qpos.clone()
    sim.step_host(act, 1, out)
    assert np.array_equal(out, sim.qpos.cpu().numpy()) and not torch.equal(before, sim.qpos)
    sim.qvel[3, 40] = float("nan"); sim.step(2)
    st = sim.status.cpu().numpy()
    assert st[3] & sim.ST_NONFINITE and not (np.delete(st, 3) & sim.ST_NONFINITE).any()


def test_mjmodel_shaped_all_biological_world_is_ingested():
    """Drop-in path: an MjModel-shaped world with the full skeleton (71 bodies, joint-less ones static, actuators adhesion-first)
    goes through from_mjmodel into the tree kernels and steps like the baked model."""
    import torch
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.convert import mjmodel_like
    m = NMFModel.bench(joint_preset="all_biological")
    a = B200Simulation(m, n_worlds=2); b = B200Simulation(mjmodel_like(m), n_worlds=2)
    for s in (a, b):
        q = s.qpos.clone(); q[:, 2] = -0.17; s.qpos.copy_(q); s.step(50)
    assert float((a.qpos - b.qpos).abs().max()) < 1e-6
