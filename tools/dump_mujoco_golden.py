"""Pin the oracle against REAL MuJoCo (tier T-C).

This cannot run in the build container or on the GPU boxes (mujoco 3.6, dm_control and flygym's dependencies are not installable
there).  A maintainer with the reference installed runs

    python tools/dump_mujoco_golden.py --out tests/golden/mujoco_golden.npz

which composes the reference benchmark model exactly as src/flygym_demo/benchmark/time_gpu_simulation.py:21-64 does, compiles it
with MuJoCo and stores (i) the compiled constants the baker restates (body masses / inertias / inverse weights, geom sizes, the
keyframe, solver options) and (ii) mj_step trajectories of BASELINE configs 1 and 2 (hold-neutral from the keyframe, zero actions,
standing, CPG walking) for `noslip_iterations = 0` (what GPUSimulation runs, warp/simulation.py:427-448) and for the CPU default.
tests/test_cpu_suite.py::test_mujoco_golden_if_present then compares the baked model and the oracle with it; without the file that
test reports "parity vs MuJoCo: not run (golden file absent)".
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


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
