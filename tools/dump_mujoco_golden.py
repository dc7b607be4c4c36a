"""Pin the oracle against REAL MuJoCo (tier T-C).

This cannot run in the build container or on the GPU boxes (mujoco 3.6, dm_control and flygym's dependencies are not installable
there).  A maintainer with the reference installed runs

    python tools/dump_mujoco_golden.py --out tests/golden/mujoco_golden.npz

which composes the reference benchmark model exactly as src/flygym_demo/benchmark/time_gpu_simulation.py:21-64 does, compiles it
with MuJoCo and stores (i) the compiled constants the baker restates (body masses / inertias / inverse weights, geom sizes, the
keyframe, solver options) and (ii) mj_step trajectories of BASELINE configs 1 and 2 (hold-neutral from the keyframe, zero actions,
standing, CPG walking) for `noslip_iterations = 0` (what GPUSimulation runs, warp/simulation.py:427-448) and for the CPU default.
It also stores every public MjModel field that flygym_b200.convert.from_mjmodel reads.  With the file present, ONE command pins baker,
converter, oracle and kernels:  tests/test_cpu_suite.py::test_mujoco_golden_if_present  replays the converter on the stored MjModel
(vs the baked model), and the oracle on every scenario (qpos, qvel, contact sensors; noslip 0 and 5);
tests/test_gpu_step.py::test_mujoco_golden_on_the_gpu_if_present  does the same with the CUDA kernels (f32 and f64).  Without the file
both report "parity vs MuJoCo: not run (golden file absent)".
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


MJMODEL_FIELDS = ["body_parentid", "body_pos", "body_quat", "body_mass", "body_ipos", "body_iquat", "body_inertia", "body_invweight0", "body_jntadr",
                  "body_jntnum", "body_dofadr", "body_dofnum", "jnt_type", "jnt_bodyid", "jnt_axis", "jnt_pos", "jnt_qposadr", "jnt_dofadr", "jnt_stiffness",
                  "qpos_spring", "qpos0", "dof_damping", "dof_armature", "actuator_trntype", "actuator_trnid", "actuator_gainprm", "actuator_biasprm",
                  "actuator_forcerange", "actuator_forcelimited", "actuator_ctrlrange", "geom_type", "geom_bodyid", "geom_pos", "geom_quat", "geom_size",
                  "geom_dataid", "mesh_vert", "mesh_vertadr", "mesh_vertnum", "mesh_graphadr", "mesh_graph", "pair_geom1", "pair_geom2", "pair_friction",
                  "pair_solref", "pair_solimp", "pair_margin", "pair_gap", "site_bodyid", "site_pos", "key_qpos", "key_ctrl", "eq_type", "eq_obj1id",
                  "eq_obj2id", "eq_data", "eq_solref", "eq_solimp"]


def mjmodel_from_golden(z, tag):
    """Rebuild the duck-typed MjModel that flygym_b200.convert.from_mjmodel ingests from a golden file written by this tool."""
    from types import SimpleNamespace
    m = SimpleNamespace()
    for key in MJMODEL_FIELDS:
        if f"{tag}/mjmodel/{key}" in z:
            setattr(m, key, z[f"{tag}/mjmodel/{key}"])
    m.nbody, m.njnt, m.nq, m.nv, m.nu, m.ngeom, m.npair, m.nsite, m.nkey, m.neq, flags = (int(x) for x in z[f"{tag}/mjmodel/sizes"])
    o = z[f"{tag}/opt"]
    m.opt = SimpleNamespace(timestep=float(o[0]), gravity=o[1:4], iterations=int(o[4]), tolerance=float(o[5]), ls_iterations=int(o[6]), ls_tolerance=float(o[7]),
                            noslip_iterations=int(o[8]), impratio=float(o[10]), enableflags=flags)
    m.stat = SimpleNamespace(meaninertia=float(o[9]))
    m.names = {k: [str(x) for x in z[f"{tag}/mjmodel/names_{k}"]] for k in ("body", "joint", "actuator", "geom", "site")}
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/golden/mujoco_golden.npz")
    ap.add_argument("--steps", type=int, default=1000)
    args = ap.parse_args()
    try:
        import mujoco as mj
        from flygym.anatomy import ActuatedDOFPreset, AxisOrder, JointPreset, Skeleton
        from flygym.compose import ActuatorType, FlatGroundWorld, Fly, GeomFittingOption, KinematicPosePreset
        from flygym.simulation import Simulation
        from flygym.utils.math import Rotation3D
    except Exception as e:   # pragma: no cover - needs the reference's environment
        sys.exit(f"this tool needs mujoco and the reference flygym package importable: {e!r}")
    from flygym_b200.actions import cpg_table
    from flygym_b200.model import NMFModel

    out = {}
    for simplify in (True, False):
        tag = "capsule" if simplify else "mesh"
        fly = Fly(geom_fitting_option=GeomFittingOption.ALL_TO_CAPSULES if simplify else GeomFittingOption.UNMODIFIED)
        skeleton = Skeleton(axis_order=AxisOrder.YAW_PITCH_ROLL, joint_preset=JointPreset.LEGS_ONLY)
        pose = KinematicPosePreset.NEUTRAL
        fly.add_joints(skeleton, neutral_pose=pose)
        fly.add_actuators(fly.skeleton.get_actuated_dofs_from_preset(ActuatedDOFPreset.LEGS_ACTIVE_ONLY),
                          actuator_type=ActuatorType.POSITION, kp=50.0, neutral_input=pose)
        fly.add_leg_adhesion()
        world = FlatGroundWorld()
        world.add_fly(fly, (0, 0, 0.8), Rotation3D("quat", (1, 0, 0, 0)))
        sim = Simulation(world)
        m, d = sim.mj_model, sim.mj_data
        # ---- (i) compiled constants, by name so that body fusing / renumbering does not matter
        names = [mj.mj_id2name(m, mj.mjtObj.mjOBJ_BODY, b) for b in range(m.nbody)]
        out[f"{tag}/body_names"] = np.array(names)
        for key in ("body_mass", "body_inertia", "body_ipos", "body_iquat", "body_invweight0", "body_pos", "body_quat", "geom_size", "geom_pos",
                    "geom_quat", "geom_type", "geom_bodyid", "dof_armature", "dof_damping", "dof_invweight0", "jnt_stiffness", "qpos_spring",
                    "actuator_gainprm", "actuator_biasprm", "actuator_forcerange", "actuator_ctrlrange", "key_qpos", "key_ctrl",
                    "pair_solref", "pair_solimp", "pair_friction", "pair_margin", "pair_gap", "eq_data", "eq_solref", "eq_solimp"):
            if hasattr(m, key):
                out[f"{tag}/{key}"] = np.array(getattr(m, key))
        out[f"{tag}/opt"] = np.array([m.opt.timestep, *m.opt.gravity, m.opt.iterations, m.opt.tolerance, m.opt.ls_iterations, m.opt.ls_tolerance,
                                      m.opt.noslip_iterations, m.stat.meaninertia, m.opt.impratio, m.opt.cone, m.opt.integrator, m.opt.solver])
        out[f"{tag}/jointdofs_order"] = np.array([str(x.name) for x in fly.get_jointdofs_order()])
        # ---- every public MjModel field flygym_b200.convert.from_mjmodel reads, so that the converter (and through it the baker)
        # can be replayed against the real compiler's output wherever the golden file is available
        for key in MJMODEL_FIELDS:
            if hasattr(m, key):
                out[f"{tag}/mjmodel/{key}"] = np.array(getattr(m, key))
        out[f"{tag}/mjmodel/sizes"] = np.array([m.nbody, m.njnt, m.nq, m.nv, m.nu, m.ngeom, m.npair, m.nsite, m.nkey, m.neq, m.opt.enableflags])
        for kind, obj, cnt in (("body", mj.mjtObj.mjOBJ_BODY, m.nbody), ("joint", mj.mjtObj.mjOBJ_JOINT, m.njnt), ("actuator", mj.mjtObj.mjOBJ_ACTUATOR, m.nu),
                               ("geom", mj.mjtObj.mjOBJ_GEOM, m.ngeom), ("site", mj.mjtObj.mjOBJ_SITE, m.nsite)):
            out[f"{tag}/mjmodel/names_{kind}"] = np.array([mj.mj_id2name(m, obj, i) or "" for i in range(cnt)])
        # ---- (ii) trajectories
        ours = NMFModel.bench(simplify_geom=simplify)
        cpg = cpg_table(ours, 4, args.steps).astype(np.float64)
        n_act = cpg.shape[2]
        key_ctrl = np.array(m.key_ctrl[0][:n_act])
        scenarios = {"1a_hold_neutral": (None, np.tile(key_ctrl, (args.steps, 1)), 0.0), "1b_zero_actions": (None, np.zeros((args.steps, n_act)), 0.0),
                     "stand": (-0.17, np.tile(key_ctrl, (args.steps, 1)), 1.0)}
        for k in range(4):
            scenarios[f"2_cpg_fly{k}"] = (-0.17, cpg[k], 1.0)
        for noslip in (0, int(m.opt.noslip_iterations)):
            m.opt.noslip_iterations = noslip
            for sname, (z0, table, adh) in scenarios.items():
                sim.reset()
                if z0 is not None:
                    d.qpos[2] = z0
                traj_q, traj_v, traj_f = [], [], []
                for t in range(args.steps):
                    sim.set_actuator_inputs(fly.name, ActuatorType.POSITION, table[t])
                    sim.set_leg_adhesion_states(fly.name, np.full(6, adh))
                    sim.step()
                    if (t + 1) in (1, 10, 100, 300, args.steps):
                        traj_q.append(d.qpos.copy()); traj_v.append(d.qvel.copy())
                        traj_f.append(np.concatenate([np.ravel(x) for x in sim.get_ground_contact_info(fly.name)]))
                out[f"{tag}/noslip{noslip}/{sname}/qpos"] = np.array(traj_q)
                out[f"{tag}/noslip{noslip}/{sname}/qvel"] = np.array(traj_v)
                out[f"{tag}/noslip{noslip}/{sname}/contact_info"] = np.array(traj_f)
        out[f"{tag}/checkpoints"] = np.array([1, 10, 100, 300, args.steps])
    out["mujoco_version"] = np.array(mj.__version__)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "with", len(out), "arrays (mujoco", mj.__version__ + ")")


if __name__ == "__main__":
    main()
