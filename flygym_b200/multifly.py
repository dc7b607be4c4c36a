"""Several flies per world (reference ``BaseWorld.add_fly`` called more than once: ``compose/world.py:95-150``,
exercised by the reference's ``tests/core/test_compose.py`` "dupworld" cases).

In the reference's worlds flies only ever collide with the ground (contact pairs are geom x ground plane,
``world.py:292-309``), never with each other, so a world holding F flies is F independent copies of the same dynamics
that differ in their spawn pose.  The step path therefore needs nothing new: world w's fly f is state record
``w * F + f`` of one :class:`B200Simulation`, and this class only keeps the name -> slot bookkeeping, the per-fly
spawn poses of the "neutral" keyframe (``world.py:151-207``) and the per-fly views of the batched getters / setters.
All flies of a world share one baked model (same skeleton, actuators, contact preset)."""
from __future__ import annotations

import numpy as np
import torch

from .anatomy import ActuatorType
from .model import NMFModel
from .simulation import B200Simulation, FlyView


class B200World:
    """``add_fly(name, spawn_position, spawn_quat)`` in the order the reference registers flies in ``fly_lookup``."""

    def __init__(self, model: NMFModel | None = None):
        self.model = model if model is not None else NMFModel.bench(simplify_geom=True)
        self.fly_lookup: dict[str, FlyView] = {}
        self.spawn: dict[str, np.ndarray] = {}

    def add_fly(self, name: str, spawn_position=(0.0, 0.0, 0.8), spawn_quat=(1.0, 0.0, 0.0, 0.0)) -> None:
        if name in self.fly_lookup:
            raise ValueError(f"Fly with name '{name}' already exists in the world.")      # world.py:122-123
        q = np.asarray(spawn_quat, dtype=np.float64)
        if q.shape != (4,) or abs(np.linalg.norm(q) - 1.0) > 1e-6:
            raise ValueError("spawn_quat must be a unit quaternion (w, x, y, z)")           # world.py:119 (quaternion format only)
        self.fly_lookup[name] = FlyView(self.model, name)
        self.spawn[name] = np.r_[np.asarray(spawn_position, dtype=np.float64), q]


class B200MultiFlySimulation:
    """The ``Simulation`` method set for worlds with several flies; getters return ``(n_worlds, ...)`` for the named fly."""

    def __init__(self, world: B200World, n_worlds: int = 1, *, device=None):
        if len(world.fly_lookup) == 0:
            raise ValueError("The world must contain at least one fly.")
        self.world = world
        self.n_worlds = int(n_worlds)
        self._names = list(world.fly_lookup)
        self.F = len(self._names)
        self._slot = {n: i for i, n in enumerate(self._names)}
        self._sim = B200Simulation(world.model, n_worlds=self.n_worlds * self.F, device=device, fly_name="_record")
        self.device = self._sim.device
        self._spawn = torch.as_tensor(np.stack([world.spawn[n] for n in self._names]), dtype=torch.float32, device=self.device)
        self.reset()

    def _f(self, fly_name: str) -> int:
        if fly_name not in self._slot:
            raise KeyError(fly_name)
        return self._slot[fly_name]

    def _per_fly(self, t: torch.Tensor, f: int) -> torch.Tensor:
        return t.view(self.n_worlds, self.F, *t.shape[1:])[:, f]

    # ---- Simulation API ---------------------------------------------------------------------------------------------
    def reset(self) -> None:
        self._sim.reset()
        self._sim.qpos.view(self.n_worlds, self.F, -1)[:, :, 0:7] = self._spawn[None]

    def step(self, n: int = 1) -> None:
        self._sim.step(n)

    def warmup(self, duration_s: float = 0.05) -> None:
        self._sim.warmup(duration_s)

    @property
    def time(self) -> float:
        return self._sim.time

    @property
    def timestep(self) -> float:
        return self._sim.timestep

    def _set(self, f: int, col0: int, ncols: int, values, what: str) -> None:
        if isinstance(values, (list, tuple)):
            values = np.asarray(values, dtype=np.float32)
        if values.shape[-1] != ncols:
            raise ValueError(f"Expected {ncols} {what}, but got {values.shape[-1]}")
        v = torch.as_tensor(np.asarray(values, dtype=np.float32) if not isinstance(values, torch.Tensor) else values).to(self.device, torch.float32)
        if v.ndim == 2 and v.shape[0] != self.n_worlds:
            raise ValueError(f"Expected leading dimension n_worlds={self.n_worlds}, got {v.shape[0]}")
        self._sim.ctrl.view(self.n_worlds, self.F, -1)[:, f, col0:col0 + ncols] = v

    def set_actuator_inputs(self, fly_name: str, actuator_type, inputs) -> None:
        f = self._f(fly_name)
        if ActuatorType(actuator_type) != ActuatorType.POSITION:
            if np.shape(inputs)[-1] != 0:
                raise ValueError(f"Expected 0 inputs for actuator type '{ActuatorType(actuator_type).name}'")
            return
        self._set(f, 0, self._sim.info.nu_pos, inputs, f"inputs for actuator type 'POSITION'")

    def set_leg_adhesion_states(self, fly_name: str, leg_to_adhesion_state) -> None:
        self._set(self._f(fly_name), self._sim.info.nu_pos, self._sim.info.nu_adh, leg_to_adhesion_state, "adhesion states")

    def get_joint_angles(self, fly_name: str) -> torch.Tensor:
        return self._per_fly(self._sim.get_joint_angles("_record"), self._f(fly_name))

    def get_joint_velocities(self, fly_name: str) -> torch.Tensor:
        return self._per_fly(self._sim.get_joint_velocities("_record"), self._f(fly_name))

    def get_body_positions(self, fly_name: str) -> torch.Tensor:
        return self._per_fly(self._sim.seg_xpos, self._f(fly_name)).clone()

    def get_body_rotations(self, fly_name: str) -> torch.Tensor:
        return self._per_fly(self._sim.seg_xquat, self._f(fly_name)).clone()

    def get_site_positions(self, fly_name: str) -> torch.Tensor:
        return self._per_fly(self._sim.get_site_positions("_record"), self._f(fly_name))

    def get_actuator_forces(self, fly_name: str, actuator_type) -> torch.Tensor:
        return self._per_fly(self._sim.get_actuator_forces("_record", actuator_type), self._f(fly_name))

    def get_ground_contact_info(self, fly_name: str):
        f = self._f(fly_name)
        return tuple(self._per_fly(t, f) for t in self._sim.get_ground_contact_info("_record"))
