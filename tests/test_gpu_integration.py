"""Executes INTEGRATION.md's reference-side subclass VERBATIM against a stand-in ``flygym`` package.

The real ``flygym`` imports ``mujoco`` / ``dm_control`` (absent here).  The stand-in reproduces exactly what the stub relies on:
``Simulation.__init__`` compiles the world and builds the name -> MuJoCo-id maps of reference ``simulation.py:32-57,311-448``
(here from a duck-typed ``MjModel``, flygym_b200.convert.mjmodel_like: world + attachment body + 69 segment bodies, actuators
declared adhesion-first, so MuJoCo ids do NOT coincide with the record layout)."""
import enum
import re
import sys
import types
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _install_flygym_stand_in():
    class ActuatorType(enum.Enum):
        POSITION = "position"; MOTOR = "motor"; ADHESION = "adhesion"

    class Simulation:
        """the part of reference simulation.py:32-57,311-448 a backend subclass inherits"""

        def __init__(self, world):
            if len(world.fly_lookup) == 0:
                raise ValueError("The world must contain at least one fly.")
            self.renderer, self.world = None, world
            self.mj_model, self.mj_data = world.compile()
            m = self.mj_model
            name2id = {kind: {n: i for i, n in enumerate(m.names[kind])} for kind in ("body", "joint", "actuator")}
            self._internal_bodyids_by_fly, self._intern_qposadrs_by_fly, self._intern_qveladrs_by_fly = {}, {}, {}
            self._intern_actuatorids_by_type_by_fly = {ActuatorType.POSITION: {}}
            self._intern_adhesionactuatorids_by_fly = {}
            for fname, fly in world.fly_lookup.items():
                self._internal_bodyids_by_fly[fname] = np.array([name2id["body"][f"{fname}/{s}"] for s in fly.get_bodysegs_order()], np.int32)
                jids = [name2id["joint"][f"{fname}/{d}"] for d in fly.get_jointdofs_order()]
                self._intern_qposadrs_by_fly[fname] = np.array([m.jnt_qposadr[j] for j in jids], np.int32)
                self._intern_qveladrs_by_fly[fname] = np.array([m.jnt_dofadr[j] for j in jids], np.int32)
                self._intern_actuatorids_by_type_by_fly[ActuatorType.POSITION][fname] = np.array(
                    [name2id["actuator"][f"{fname}/{d}-position"] for d in fly.get_actuated_jointdofs_order("position")], np.int32)
                self._intern_adhesionactuatorids_by_fly[fname] = np.array(
                    [name2id["actuator"][f"{fname}/{leg}_tarsus5-adhesion"] for leg in fly.get_legs_order()], np.int32)

    pkg = types.ModuleType("flygym"); sim = types.ModuleType("flygym.simulation"); comp = types.ModuleType("flygym.compose")
    fly = types.ModuleType("flygym.compose.fly")
    sim.Simulation = Simulation; fly.ActuatorType = ActuatorType
    pkg.simulation, pkg.compose, comp.fly = sim, comp, fly
    sys.modules.update({"flygym": pkg, "flygym.simulation": sim, "flygym.compose": comp, "flygym.compose.fly": fly})
    return ActuatorType


class _World:
    """stand-in for a composed FlatGroundWorld holding one fly named 'nmf'"""

    def __init__(self, model):
        from flygym_b200.convert import mjmodel_like
        from flygym_b200.simulation import FlyView
        self._mj = mjmodel_like(model, prefix="nmf/")
        self.fly_lookup = {"nmf": FlyView(model, "nmf")}

    def compile(self):
        return self._mj, None


@pytest.mark.parametrize("skeleton", ["legs_only", "all_biological"])
def test_integration_md_subclass_runs_and_matches_the_package_class(monkeypatch, skeleton):
    import torch
    from flygym_b200 import B200Simulation as Packaged, NMFModel, _lib
    from flygym_b200.actions import cpg_table
    ActuatorType = _install_flygym_stand_in()
    monkeypatch.setenv("NMF_LIB_PATH", str(_lib.build()))
    code = re.search(r"```python\n(.*?)```", (ROOT / "INTEGRATION.md").read_text(), flags=re.S).group(1)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    Stub = ns["B200Simulation"]
    model = NMFModel.bench(True, joint_preset=skeleton)         # star kernels / general-topology kernels behind the same stub
    n = 3
    a = Stub(_World(model), n)                                   # reference-side subclass over the C ABI, MuJoCo numbering
    b = Packaged(model, n_worlds=n)                              # the package's own class, record numbering
    c = Packaged(_World(model), n_worlds=n)                      # the package class fed with the flygym-style world
    tab = cpg_table(model, n, 40)
    for sim in (a, b, c):
        sim.set_leg_adhesion_states("nmf", np.ones(6, np.float32))
    for s in range(40):
        for sim in (a, b, c):
            sim.set_actuator_inputs("nmf", ActuatorType.POSITION if sim is a else "position", tab[:, s])
            sim.step()
    torch.cuda.synchronize()
    for sim in (a, c):
        assert torch.equal(sim.get_joint_angles("nmf"), b.get_joint_angles("nmf"))
        assert torch.equal(sim.get_joint_velocities("nmf"), b.get_joint_velocities("nmf"))
        assert torch.equal(sim.get_body_positions("nmf"), b.get_body_positions("nmf"))
        assert torch.equal(sim.get_body_rotations("nmf"), b.get_body_rotations("nmf"))
        for x, y in zip(sim.get_ground_contact_info("nmf"), b.get_ground_contact_info("nmf")):
            assert torch.equal(x, y)
    assert torch.equal(a.get_actuator_forces("nmf", ActuatorType.POSITION), b.get_actuator_forces("nmf", "position"))
    assert abs(a.time - 40 * 1e-4) < 1e-7 and float(b.get_joint_angles("nmf").abs().max()) > 0.1
    with pytest.raises(ValueError):
        a.set_actuator_inputs("nmf", ActuatorType.POSITION, np.zeros(41, np.float32))
    for k in [k for k in sys.modules if k == "flygym" or k.startswith("flygym.")]:
        del sys.modules[k]
