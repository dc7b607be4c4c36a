"""General-topology step kernels (flygym_b200/csrc/nmf_tree.cuh) on the CPU: the REAL kernel source runs through the SIMT
emulator (tests/simt_emu, test infrastructure) against the fp64 oracle, for the skeletons the star kernels reject:
JointPreset.ALL_BIOLOGICAL / ALL_POSSIBLE (reference src/flygym/anatomy.py:388-460) and ContactBodiesPreset.ALL
(anatomy.py:519-526).  PARITY UNPINNED against real MuJoCo (see oracle/nmf_oracle.c)."""
import ctypes

import numpy as np
import pytest

import emu
from flygym_b200 import NMFModel
from oracle.oracle import Oracle


def _settled(model, settle=1200):
    """(oracle, float32 tree-layout record): the fly after `settle` oracle steps from the keyframe (standing on its tarsi), the
    oracle restarted from the float32-rounded state so that both sides start from identical numbers."""
    info = emu.tree_info(model)
    o = Oracle(model); o.reset(); o.step(settle)
    st = np.zeros((1, info["s_stride"]), np.float32)
    st[0, :info["nq"]] = o.qpos
    st[0, info["s_qvel"]:info["s_qvel"] + info["nv"]] = o.qvel
    st[0, info["s_warm"]:info["s_warm"] + info["nv"]] = o.get("qacc_warmstart")
    st[0, info["s_ctrl"]:info["s_ctrl"] + info["nu"]] = o.ctrl
    o.qpos[:] = st[0, :info["nq"]]; o.qvel[:] = st[0, info["s_qvel"]:info["s_qvel"] + info["nv"]]
    o.get("qacc_warmstart")[:] = st[0, info["s_warm"]:info["s_warm"] + info["nv"]]
    return o, st, info


def _compare(model, nsteps, precision, settle=1200):
    o, st, info = _settled(model, settle)
    r = emu.tree_step(model, st, nsteps, precision=precision, outputs=True, dbg=True)
    o.step(nsteps)
    qpos = st[0, :info["nq"]].astype(np.float64); qvel = st[0, info["s_qvel"]:info["s_qvel"] + info["nv"]].astype(np.float64)
    so = o.get("sensordata").reshape(-1, 16); sg = r["sensor"][0].reshape(-1, 16)
    oq = o.get("seg_xquat").reshape(-1, 4); gq = r["xquat"][0]
    return dict(ncon=o.dim("ncon"), qpos=np.abs(qpos - o.qpos).max(), qvel=np.abs(qvel - o.qvel).max() / max(1.0, np.abs(o.qvel).max()),
                actf=np.abs(r["actf"][0] - o.get("actuator_force")).max(), xpos=np.abs(r["xpos"][0].ravel() - o.get("seg_xpos")).max(),
                xquat=np.abs(gq * np.sign((gq * oq).sum(1, keepdims=True)) - oq).max(),
                found=np.abs(sg[:, 0] - so[:, 0]).max(), force=np.abs(sg[:, 1:10] - so[:, 1:10]).max() / max(1.0, np.abs(so[:, 1:4]).max()),
                energy=np.abs(r["energy"][0] - o.get("energy")).max() / np.abs(o.get("energy")).max(), ncon_kernel=int(r["dbg"][0, 1]))


def test_all_biological_f64_kernel_source_matches_the_oracle():
    """126 hinge DoFs (head, proboscis, antennae, eyes, abdomen, wings, halteres articulated): every output after 12 steps of a
    standing fly, double-precision instantiation -> what is left is the float32 rounding of the buffers."""
    m = NMFModel.bench(joint_preset="all_biological")
    assert m.nv == 132 and m.dim("ngeom") == 55
    e = _compare(m, 12, 64)
    print(e)
    assert e["ncon"] >= 6 and e["ncon_kernel"] == e["ncon"]
    assert e["qpos"] < 3e-7 and e["qvel"] < 2e-5 and e["actf"] < 2e-5 and e["xpos"] < 5e-7 and e["xquat"] < 2e-7
    assert e["found"] == 0 and e["force"] < 1e-5 and e["energy"] < 1e-6


def test_all_biological_f32_kernel_source_matches_the_oracle():
    m = NMFModel.bench(joint_preset="all_biological")
    e = _compare(m, 12, 32)
    print(e)
    assert e["qpos"] < 2e-6 and e["qvel"] < 5e-4 and e["actf"] < 2e-4 and e["xpos"] < 3e-6 and e["found"] == 0 and e["force"] < 5e-4


@pytest.mark.parametrize("kw", [dict(joint_preset="all_possible", contact_preset="all"), dict(joint_preset="all_biological", simplify_geom=False),
                                dict(joint_preset="all_biological", terrain="blocks"), dict(contact_preset="all")],
                         ids=["all_possible+all_contacts", "all_biological_mesh_multiccd", "all_biological_blocks", "legs_only+all_contacts"])
def test_other_general_models_f64(kw):
    """ALL_POSSIBLE (204 hinge DoFs) with ContactBodiesPreset.ALL, mesh hulls with multiccd, box-column terrain, and the LEGS_ONLY
    skeleton with 21 contact geoms on the hub (more than the star kernels' 16 hub lanes)."""
    m = NMFModel.bench(**kw)
    e = _compare(m, 5, 64, settle=1000)
    print(kw, e)
    assert e["ncon"] >= 6 and e["ncon_kernel"] == e["ncon"]
    assert e["qpos"] < 3e-7 and e["qvel"] < 2e-5 and e["xpos"] < 5e-7 and e["found"] == 0 and e["force"] < 1e-5


@pytest.mark.parametrize("kw,precision", [(dict(joint_preset="all_biological"), 64), (dict(joint_preset="all_biological"), 32),
                                          (dict(joint_preset="all_biological", simplify_geom=False), 64)],
                         ids=["capsule_f64", "capsule_f32", "mesh_multiccd_f64"])
def test_noslip_on_the_tree_kernels_matches_the_oracle(kw, precision):
    """`noslip_iterations: 5` (mujoco_globals.yaml:15, the CPU `Simulation` semantics) on the general-topology kernels: a settled
    ALL_BIOLOGICAL fly pushed sideways, 10 steps, against the oracle's noslip ([PRIOR] mj_solNoSlip) -- and the post-solver has to
    matter (the plain solve of the same state ends somewhere else)."""
    base = NMFModel.bench(**kw)
    m = base.with_options(noslip_iterations=5)
    o, st, info = _settled(m, 800)
    v0 = info["s_qvel"]
    st[0, v0:v0 + 2] += np.float32([3.0, -2.0]); o.qvel[0:2] = st[0, v0:v0 + 2]
    plain = Oracle(base); plain.reset()
    plain.qpos[:] = o.qpos; plain.qvel[:] = o.qvel; plain.get("qacc_warmstart")[:] = o.get("qacc_warmstart"); plain.ctrl[:] = o.ctrl
    r = emu.tree_step(m, st, 10, precision=precision, outputs=True, dbg=True)
    o.step(10); plain.step(10)
    qpos = st[0, :info["nq"]].astype(np.float64); qvel = st[0, v0:v0 + info["nv"]].astype(np.float64)
    so = o.get("sensordata").reshape(-1, 16); sg = r["sensor"][0].reshape(-1, 16)
    e = dict(qpos=np.abs(qpos - o.qpos).max(), qvel=np.abs(qvel - o.qvel).max() / max(1.0, np.abs(o.qvel).max()),
             force=np.abs(sg[:, 1:10] - so[:, 1:10]).max() / max(1.0, np.abs(so[:, 1:4]).max()), found=np.abs(sg[:, 0] - so[:, 0]).max(),
             gap=np.abs(plain.qvel - o.qvel).max() / max(1.0, np.abs(o.qvel).max()), status=float(st[0, info["s_time"] + 1]))
    print(kw, precision, e)
    assert e["status"] == 0 and e["found"] == 0 and e["gap"] > 1e-3
    if precision == 64:
        assert e["qpos"] < 3e-7 and e["qvel"] < 2e-5 and e["force"] < 2e-5
    else:
        assert e["qpos"] < 3e-6 and e["qvel"] < 1e-3 and e["force"] < 2e-3


def test_tree_kernel_equals_star_kernel_on_the_benchmark_model():
    """The benchmark skeleton stepped by both kernel families (float32, 4 CPG walkers, 15 steps): same physics, different
    factorisation order -> agreement at float32 rounding."""
    from flygym_b200.actions import cpg_table
    m = NMFModel.bench(True)
    n, T = 3, 15
    tab = np.zeros((n, T, m.nu), np.float32); tab[:, :, :42] = cpg_table(m, n, T); tab[:, :, 42:] = 1.0
    key = emu.key_state(m); key[2] = -0.17
    star = np.tile(key, (n, 1)); star[:, 0] = 0.3 * np.arange(n)
    info = emu.tree_info(m)
    tree = np.zeros((n, info["s_stride"]), np.float32)
    tree[:, :73] = star[:, :73]; tree[:, info["s_ctrl"]:info["s_ctrl"] + 48] = star[:, emu.S_CTRL:emu.S_CTRL + 48]
    emu.step(m, star, T, act_table=tab)
    emu.tree_step(m, tree, T, act_table=tab)
    dq = np.abs(star[:, :73] - tree[:, :73]).max(); dv = np.abs(star[:, emu.S_QVEL:emu.S_QVEL + 72] - tree[:, info["s_qvel"]:info["s_qvel"] + 72]).max()
    print(dq, dv)
    assert dq < 5e-6 and dv < 5e-3
    assert tree[0, info["s_time"]] == np.float32(T * m.timestep) and tree[0, info["s_time"] + 1] == 0


def test_tree_host_tables_and_limits():
    """Shared-memory plans fit one SM (227 KB opt-in) in both precisions; what the tree kernels do not handle is refused with a
    message instead of being stepped wrongly."""
    for kw in (dict(joint_preset="all_biological"), dict(joint_preset="all_biological", simplify_geom=False), dict(joint_preset="all_possible", contact_preset="all")):
        info = emu.tree_info(NMFModel.bench(**kw))
        assert info["smem_f32"] < 100 * 1024 and info["smem_f64"] < 227 * 1024, info
        assert info["s_stride"] % 4 == 0 and info["s_qvel"] >= info["nq"] and info["s_time"] >= info["s_ctrl"] + info["nu"]
    m = NMFModel.bench(joint_preset="all_biological")
    assert emu.tree_info(m)["nH"] == sum(len(_anc(m, k)) for k in range(m.nv))
    # a model that asks for the noslip post-solver gets its B_tt / force region (9.8 k reals + one int per contact slot) on top
    for kw in (dict(joint_preset="all_biological"), dict(joint_preset="all_possible", contact_preset="all")):
        plain, ns = emu.tree_info(NMFModel.bench(**kw)), emu.tree_info(NMFModel.bench(**kw).with_options(noslip_iterations=5))
        assert 9500 * 4 < ns["smem_f32"] - plain["smem_f32"] < 10500 * 4 and ns["smem_f64"] < 227 * 1024, (plain, ns)


def _anc(m, k):
    p = m.arrays["dof_parent"]; out = []
    while k >= 0:
        out.append(k); k = int(p[k])
    return out


def test_from_mjmodel_ingests_general_skeletons():
    """The converter no longer insists on the star topology: an MjModel-shaped ALL_BIOLOGICAL world (71 bodies, joint-less ones as
    static children, actuators adhesion-first) comes back as the baked arrays, and the oracle walks identically on both."""
    from flygym_b200.convert import from_mjmodel, mjmodel_like
    m = NMFModel.bench(joint_preset="all_biological")
    m2 = from_mjmodel(mjmodel_like(m))
    assert m2.nv == 132 and m2.names["jointdofs"] == m.names["jointdofs"] and m2.names["legs"] == m.names["legs"]
    for k in ("body_pos", "body_mass", "body_inertia", "dof_axis", "geom_pos", "geom_size", "act_dof", "adh_body", "body_parent", "body_leg", "key_qpos"):
        assert np.abs(np.asarray(m.arrays[k], float).ravel() - np.asarray(m2.arrays[k], float).ravel()).max() < 1e-9, k
    o1, o2 = Oracle(m), Oracle(m2); o1.step(200); o2.step(200)
    assert np.abs(o1.qpos - o2.qpos).max() < 1e-12


def test_tethered_world_with_the_full_skeleton():
    """TetheredWorld (reference world.py:334-366) with the ALL_BIOLOGICAL skeleton: the six weld rows sit on the root body, so they
    enter the tree kernels like a contact on the root (wrench + augmentation of the root's spatial inertia).  f64 source vs oracle."""
    m = NMFModel.tethered(joint_preset="all_biological")
    info = emu.tree_info(m)
    assert info["nv"] == 132
    st = emu.tree_key_state(m)[None].copy()
    o = Oracle(m); o.reset()
    emu.tree_step(m, st, 10, precision=64)
    o.step(10)
    q = st[0, :info["nq"]].astype(np.float64)
    print(np.abs(q - o.qpos).max(), np.abs(o.qpos - m.arrays["key_qpos"]).max())
    assert np.abs(o.qpos - m.arrays["key_qpos"]).max() > 1e-3            # the weld really pulls
    assert np.abs(q - o.qpos).max() < 5e-6 * max(1.0, np.abs(o.qpos).max())
    # the benchmark skeleton through both kernel families
    mb = NMFModel.tethered()
    a = emu.key_state(mb)[None].copy(); emu.step(mb, a, 10)
    ib = emu.tree_info(mb); b = emu.tree_key_state(mb)[None].copy(); emu.tree_step(mb, b, 10)
    assert np.abs(a[0, :73] - b[0, :73]).max() < 2e-4 * np.abs(a[0, :73]).max()


def test_tethered_full_skeleton_with_noslip():
    """The weld rows of the tethered ALL_BIOLOGICAL world under noslip (equality rows, swept unclamped): f64 source vs the oracle, and
    far from the plain solve (the soft weld becomes nearly hard)."""
    base = NMFModel.tethered(joint_preset="all_biological"); m = base.with_options(noslip_iterations=5)
    info = emu.tree_info(m)
    st = emu.tree_key_state(m)[None].copy()
    o = Oracle(m); o.reset(); plain = Oracle(base); plain.reset()
    emu.tree_step(m, st, 10, precision=64)
    o.step(10); plain.step(10)
    q = st[0, :info["nq"]].astype(np.float64)
    print(np.abs(q - o.qpos).max(), np.abs(plain.qpos - o.qpos).max())
    assert np.abs(plain.qpos - o.qpos).max() > 1e-3
    assert np.abs(q - o.qpos).max() < 5e-6 * max(1.0, np.abs(o.qpos).max()) and st[0, info["s_time"] + 1] == 0
