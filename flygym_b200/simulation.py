"""``B200Simulation`` — batched NeuroMechFly simulation on hand-written sm_100a kernels.

Presents the method set of the reference's ``flygym.Simulation`` as batched by
``flygym.warp.GPUSimulation`` (reference ``src/flygym/simulation.py:59-480``,
``src/flygym/warp/simulation.py:28-453``): same names, argument meaning, orders and
error behaviour.  Differences, all forced by the environment:

* the model comes from a baked :class:`~flygym_b200.model.NMFModel` instead of a
  ``dm_control``/MuJoCo-compiled ``world`` (neither is installable here); the benchmark skeleton (hub + 6 x 8 leg links)
  runs on the star-topology kernels, every other skeleton (``JointPreset.ALL_BIOLOGICAL`` / ``ALL_POSSIBLE``,
  ``ContactBodiesPreset.ALL``) on the general-topology kernels -- the native library picks, the API is the same;
* batched arrays are ``torch.Tensor`` (float32, ``(n_worlds, ...)``) where the
  reference returns ``wp.array``; numpy or torch inputs are accepted where the
  reference accepts numpy or warp;
* ``get_ground_contact_info`` is implemented on the device (the reference's
  ``GPUSimulation`` silently returns stale CPU values).

PyTorch only owns memory and streams; all arithmetic is in ``csrc/*.cu`` behind
the C ABI of ``include/nmf_b200.h``.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes
from time import perf_counter_ns

import numpy as np
import torch

from . import _lib
from .anatomy import ActuatorType
from .model import NMFModel


class FlyView:
    """The slice of the reference ``Fly`` object that ``Simulation`` users touch:
    the ordering getters (reference ``compose/fly.py:189-219``)."""

    def __init__(self, model: NMFModel, name: str = "nmf"):
        self.name = name
        self._names = model.names

    def get_bodysegs_order(self): return list(self._names["segments"])
    def get_jointdofs_order(self): return list(self._names["jointdofs"])
    def get_legs_order(self): return list(self._names["legs"])
    def get_sites_order(self): return list(self._names["sites"])

    def get_actuated_jointdofs_order(self, actuator_type):
        actuator_type = ActuatorType(actuator_type)
        return list(self._names["actuated_position"]) if actuator_type == ActuatorType.POSITION else []


class WorldView:
    """Stand-in for the reference ``BaseWorld`` (``fly_lookup`` only)."""

    def __init__(self, model: NMFModel, fly_name: str = "nmf"):
        self.model = model
        self.fly_lookup = {fly_name: FlyView(model, fly_name)}


class B200Simulation:
    """GPU-resident parallel simulation of ``n_worlds`` independent flies.

    Args:
        world: a baked :class:`NMFModel`, a :class:`WorldView`, a MuJoCo-compiled ``MjModel`` (anything exposing its public
            fields) or a flygym world with ``compile()`` / ``fly_lookup`` -- both converted by
            :func:`flygym_b200.convert.from_mjmodel` -- or ``None`` for the reference benchmark model (capsule geoms).
        n_worlds: number of parallel flies on this GPU.
        device: CUDA device (default: current).
        outputs: allocate the observation buffers (body poses, actuator forces,
            contact sensors) that the kernel fills every launch.
        debug: also allocate the solver-internals dump used by the parity tests.
    """

    def __init__(self, world=None, n_worlds: int = 1, *, device=None, outputs: bool = True, debug: bool = False,
                 fly_name: str = "nmf") -> None:
        if world is None:
            world = NMFModel.bench(simplify_geom=True)
        if hasattr(world, "compile") and hasattr(world, "fly_lookup") and not hasattr(world, "model"):
            # a flygym world (reference compose/base.py:21-27): ingest its MuJoCo-compiled model like Simulation.__init__ does
            if len(world.fly_lookup) == 0:
                raise ValueError("The world must contain at least one fly.")
            from .convert import from_mjmodel
            fly_name = next(iter(world.fly_lookup))
            mj_model = world.compile()
            mj_model = mj_model[0] if isinstance(mj_model, tuple) else mj_model
            world = WorldView(from_mjmodel(mj_model), fly_name)
        elif hasattr(world, "nbody") and hasattr(world, "jnt_type"):          # a (duck-typed) MjModel
            from .convert import from_mjmodel
            world = from_mjmodel(world)
        if isinstance(world, NMFModel):
            world = WorldView(world, fly_name)
        if len(world.fly_lookup) == 0:
            raise ValueError("The world must contain at least one fly.")
        if not torch.cuda.is_available():
            raise RuntimeError("B200Simulation needs a CUDA device (there is no CPU fallback).")
        self.world = world
        self.model: NMFModel = world.model
        self.renderer = None
        self.n_worlds = int(n_worlds)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError(f"B200Simulation needs a CUDA device, got {self.device}")
        if self.device.index is None:       # plain "cuda": the current device, where the tensors below are allocated
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.load()
        blob = self.model.to_blob()
        h = ctypes.c_void_p()
        rc = self._lib.nmf_create(blob, len(blob), self.n_worlds, int(self.device.index), ctypes.byref(h))
        self._h = h
        if rc != 0:
            msg = self._lib.nmf_last_error(h).decode() if h else "allocation failed"
            raise RuntimeError(f"nmf_create failed: {msg}")
        self.info = _lib.NmfInfo()
        self._check(self._lib.nmf_model_info(self._h, ctypes.byref(self.info)))
        i = self.info
        n, dev = self.n_worlds, self.device
        self.state = torch.zeros((n, i.state_stride), dtype=torch.float32, device=dev)
        self.seg_xpos = torch.zeros((n, i.nseg, 3), dtype=torch.float32, device=dev) if outputs else None
        self.seg_xquat = torch.zeros((n, i.nseg, 4), dtype=torch.float32, device=dev) if outputs else None
        self.act_force = torch.zeros((n, i.nu_pos + i.nu_adh), dtype=torch.float32, device=dev) if outputs else None
        self.sensordata = torch.zeros((n, i.nleg * 16), dtype=torch.float32, device=dev) if outputs else None
        self.debug = torch.zeros((n, i.dbg_stride), dtype=torch.float32, device=dev) if debug else None
        self.energy = torch.zeros((n, 2), dtype=torch.float32, device=dev) if outputs else None
        self._bind()
        self._build_index_maps()
        self._curr_step = 0
        self._frames_rendered = 0
        self._total_physics_time_ns = 0
        self._total_render_time_ns = 0
        self.reset()

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"libnmf_b200: {self._lib.nmf_last_error(self._h).decode()} (status {rc})")

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def _bind(self) -> None:
        b = _lib.NmfBuffers(self._ptr(self.state), self._ptr(self.seg_xpos), self._ptr(self.seg_xquat),
                            self._ptr(self.act_force), self._ptr(self.sensordata), self._ptr(self.debug), self._ptr(self.energy))
        self._check(self._lib.nmf_bind(self._h, ctypes.byref(b)))

    def _build_index_maps(self) -> None:
        """Name -> index tables (reference ``_map_internal_*``, simulation.py:311-448)."""
        m, dev = self.model, self.device
        names = m.names
        self._fly_names = list(self.world.fly_lookup.keys())
        # hinge DoF j of the kernel layout lives at qpos[7 + j] / qvel[6 + j]; locked DoFs (joint presets with fewer DoFs) are not exposed
        exposed = torch.as_tensor(m.exposed_hinge_dofs(), dtype=torch.int32, device=dev)
        assert exposed.numel() == len(names["jointdofs"])
        self._qpos_cols = (7 + exposed).contiguous()
        self._qvel_cols = (6 + exposed).contiguous()
        self._act_cols = {ActuatorType.POSITION: torch.arange(0, m.dim("nu_pos"), dtype=torch.int32, device=dev)}
        self._adh_cols = torch.arange(m.dim("nu_pos"), m.nu, dtype=torch.int32, device=dev)
        seg_index = {s: k for k, s in enumerate(names["segments"])}
        self._site_segs = torch.tensor([seg_index[s.split("-")[1]] for s in names["sites"]], dtype=torch.long, device=dev)

    def _fly(self, fly_name: str) -> None:
        if fly_name not in self.world.fly_lookup:
            raise KeyError(fly_name)

    def _as_device(self, x, ncols: int, what: str) -> torch.Tensor:
        """numpy / torch / sequence -> contiguous float32 device tensor of shape (n_worlds, ncols)."""
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x, dtype=np.float32))
        x = x.to(device=self.device, dtype=torch.float32)
        if x.ndim == 1:
            if x.shape[0] != ncols:
                raise ValueError(f"Expected {ncols} {what}, but got {x.shape[0]}")
            x = x.unsqueeze(0).expand(self.n_worlds, ncols)
        if x.shape[-1] != ncols:
            raise ValueError(f"Expected {ncols} {what}, but got {x.shape[-1]}")
        if x.shape[0] != self.n_worlds:
            raise ValueError(f"Expected leading dimension n_worlds={self.n_worlds}, got {x.shape[0]}")
        return x.contiguous()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.nmf_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ reference API
    def reset(self, mask=None) -> None:
        """All worlds (or those selected by the boolean ``mask``) <- keyframe "neutral"."""
        mptr = None
        if mask is not None:
            mask = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = ctypes.c_void_p(mask.data_ptr())
        self._check(self._lib.nmf_reset(self._h, mptr, self._stream()))
        if self.renderer is not None:
            self.renderer.reset()
        self._curr_step = 0
        self._frames_rendered = 0
        self._total_physics_time_ns = 0
        self._total_render_time_ns = 0

    def step(self, n: int = 1, action_table: torch.Tensor | None = None, table_t0: int = 0) -> None:
        """Advance all worlds by ``n`` timesteps inside one kernel launch.

        ``action_table`` (device float32 ``(n_worlds, T, n_position_actuators)``), when
        given, supplies the position-actuator inputs of step ``s`` from row
        ``(table_t0 + s) % T`` (the reference benchmark's replay protocol); with
        ``n_position_actuators + 6`` columns the trailing six are the leg adhesion inputs."""
        if action_table is None:
            self._check(self._lib.nmf_step(self._h, int(n), None, 0, 0, 0, self._stream()))
        else:
            if action_table.dtype != torch.float32 or not action_table.is_cuda or not action_table.is_contiguous():
                raise ValueError("action_table must be a contiguous float32 CUDA tensor")
            if action_table.ndim != 3 or action_table.shape[0] != self.n_worlds or \
                    action_table.shape[2] not in (self.info.nu_pos, self.info.nu_pos + self.info.nu_adh):
                raise ValueError("action_table must have shape (n_worlds, T, n_position_actuators [+ 6 adhesion inputs])")
            self._check(self._lib.nmf_step(self._h, int(n), ctypes.c_void_p(action_table.data_ptr()),
                                           int(action_table.shape[1]), int(table_t0), int(action_table.shape[2]), self._stream()))

    def forward(self) -> None:
        """``mj_forward`` for every world: refresh body poses, actuator forces and contact sensors from the current state
        without advancing it (e.g. right after ``reset`` or after writing ``qpos`` directly)."""
        self._check(self._lib.nmf_forward(self._h, self._stream()))

    def step_with_profile(self) -> None:
        t0 = perf_counter_ns()
        self.step()
        torch.cuda.synchronize(self.device)
        self._total_physics_time_ns += perf_counter_ns() - t0
        self._curr_step += 1

    def warmup(self, duration_s: float = 0.05) -> None:
        n_steps = int(duration_s / self.timestep)
        self.step(n_steps)

    def set_actuator_inputs(self, fly_name: str, actuator_type, inputs) -> None:
        self._fly(fly_name)
        actuator_type = ActuatorType(actuator_type)
        cols = self._act_cols.get(actuator_type)
        ncols = 0 if cols is None else int(cols.numel())
        if isinstance(inputs, (list, tuple)):
            inputs = np.asarray(inputs, dtype=np.float32)
        if inputs.shape[-1] != ncols:
            raise ValueError(f"Expected {ncols} inputs for actuator type '{actuator_type.name}', but got {inputs.shape[-1]}")
        if ncols == 0:
            return
        src = self._as_device(inputs, ncols, "inputs")
        self._check(self._lib.nmf_scatter_ctrl(self._h, self._ptr(src), self._ptr(cols), ncols, self._stream()))

    def set_leg_adhesion_states(self, fly_name: str, leg_to_adhesion_state) -> None:
        self._fly(fly_name)
        ncols = int(self._adh_cols.numel())
        if isinstance(leg_to_adhesion_state, (list, tuple)):
            leg_to_adhesion_state = np.asarray(leg_to_adhesion_state, dtype=np.float32)
        if leg_to_adhesion_state.shape[-1] != ncols:
            raise ValueError(f"Unexpected number of adhesion states: expected {ncols}, got {leg_to_adhesion_state.shape[-1]}")
        src = self._as_device(leg_to_adhesion_state, ncols, "adhesion states")
        self._check(self._lib.nmf_scatter_ctrl(self._h, self._ptr(src), self._ptr(self._adh_cols), ncols, self._stream()))

    def _gather(self, off: int, cols: torch.Tensor) -> torch.Tensor:
        dst = torch.empty((self.n_worlds, cols.numel()), dtype=torch.float32, device=self.device)
        self._check(self._lib.nmf_gather_state(self._h, off, self._ptr(cols), int(cols.numel()), self._ptr(dst), self._stream()))
        return dst

    def get_joint_angles(self, fly_name: str) -> torch.Tensor:
        self._fly(fly_name)
        return self._gather(self.info.off_qpos, self._qpos_cols)

    def get_joint_velocities(self, fly_name: str) -> torch.Tensor:
        self._fly(fly_name)
        return self._gather(self.info.off_qvel, self._qvel_cols)

    def _need_outputs(self):
        if self.seg_xpos is None:
            raise RuntimeError("this simulation was created with outputs=False")

    def get_body_positions(self, fly_name: str) -> torch.Tensor:
        self._fly(fly_name); self._need_outputs()
        return self.seg_xpos.clone()

    def get_body_rotations(self, fly_name: str) -> torch.Tensor:
        self._fly(fly_name); self._need_outputs()
        return self.seg_xquat.clone()

    def get_site_positions(self, fly_name: str) -> torch.Tensor:
        self._fly(fly_name); self._need_outputs()
        return self.seg_xpos[:, self._site_segs, :]

    def get_actuator_forces(self, fly_name: str, actuator_type) -> torch.Tensor:
        self._fly(fly_name); self._need_outputs()
        actuator_type = ActuatorType(actuator_type)
        if actuator_type == ActuatorType.ADHESION:
            return self.act_force[:, self.info.nu_pos:].clone()
        cols = self._act_cols.get(actuator_type)
        if cols is None:
            return torch.zeros((self.n_worlds, 0), dtype=torch.float32, device=self.device)
        return self.act_force[:, : self.info.nu_pos].clone()

    def get_ground_contact_info(self, fly_name: str):
        self._fly(fly_name); self._need_outputs()
        s = self.sensordata.view(self.n_worlds, self.info.nleg, 16)
        return (s[:, :, 0].clone(), s[:, :, 1:4].clone(), s[:, :, 4:7].clone(), s[:, :, 7:10].clone(),
                s[:, :, 10:13].clone(), s[:, :, 13:16].clone())

    # ---- raw state views (extension; zero-copy) -----------------------------
    @property
    def qpos(self) -> torch.Tensor:
        return self.state[:, self.info.off_qpos: self.info.off_qpos + self.info.nq]

    @property
    def qvel(self) -> torch.Tensor:
        return self.state[:, self.info.off_qvel: self.info.off_qvel + self.info.nv]

    @property
    def ctrl(self) -> torch.Tensor:
        return self.state[:, self.info.off_ctrl: self.info.off_ctrl + self.info.nu_pos + self.info.nu_adh]

    @property
    def qacc_warmstart(self) -> torch.Tensor:
        return self.state[:, self.info.off_qacc_warmstart: self.info.off_qacc_warmstart + self.info.nv]

    # bits of the per-fly status word (include/nmf_b200.h, enum nmf_fly_status)
    ST_NONFINITE, ST_NEWTON_CAP, ST_LS_CAP, ST_NOSLIP_SKIP = 1, 2, 4, 8

    @property
    def status(self) -> torch.Tensor:
        """Per-world status word ``(n_worlds,)`` int32: OR of ``ST_NONFINITE`` (a velocity became NaN / infinite),
        ``ST_NEWTON_CAP`` (the solver hit the model's ``iterations`` with the active set still changing) and ``ST_LS_CAP`` (a
        line search used up its evaluations).  Sticky until the world is reset.  Device-side faults are reported here, never by
        trapping (the reference's MuJoCo emits ``mju_warning`` / resets the data instead)."""
        return self.state[:, self.info.off_status].to(torch.int32)

    def get_energy(self) -> torch.Tensor:
        """``(n_worlds, 2)`` potential and kinetic energy of the state the last step started from (``mjData.energy`` with the
        reference model's ``energy`` flag, ``mujoco_globals.yaml:19``)."""
        self._need_outputs()
        return self.energy.clone()

    @property
    def time(self) -> float:
        """Current simulation time in seconds (from world 0; forces a device sync like the reference)."""
        return float(self.state[0, self.info.off_time].item())

    @property
    def timestep(self) -> float:
        return float(self.model.timestep)

    def export_state(self, world_id: int = 0) -> dict:
        """``qpos`` / ``qvel`` / ``time`` of one world as host arrays: what the reference's CPU renderer pulls out of the batched
        data before drawing a frame (``mjw.get_data_into``, reference ``warp/rendering.py:357-359``)."""
        i = self.info
        rec = self.state[int(world_id)].cpu().numpy().astype(np.float64)
        return {"qpos": rec[i.off_qpos:i.off_qpos + i.nq].copy(), "qvel": rec[i.off_qvel:i.off_qvel + i.nv].copy(),
                "time": float(rec[i.off_time])}

    def step_host(self, actions_host: np.ndarray, nsteps: int, qpos_host: np.ndarray) -> None:
        """End-to-end call with HOST buffers (H2D actions, ``nsteps`` steps, D2H qpos); synchronous.  With page-locked buffers
        (``torch.Tensor.pin_memory().numpy()``) the library replays the whole pipeline as one CUDA graph."""
        a, q, info = actions_host, qpos_host, self.info
        if (a.dtype != np.float32 or a.ndim != 2 or a.shape[0] != self.n_worlds or not a.flags.c_contiguous
                or a.shape[1] not in (info.nu_pos, info.nu_pos + info.nu_adh)):
            raise ValueError(f"actions_host must be a C-contiguous float32 array of shape ({self.n_worlds}, {info.nu_pos}) or "
                             f"({self.n_worlds}, {info.nu_pos + info.nu_adh})")
        if q.dtype != np.float32 or q.shape != (self.n_worlds, info.nq) or not q.flags.c_contiguous:
            raise ValueError(f"qpos_host must be a C-contiguous float32 array of shape ({self.n_worlds}, {info.nq})")
        # (plain integers for the pointer arguments: this call is made once per step, every microsecond of wrapper shows end to end)
        rc = self._lib.nmf_step_host(self._h, a.ctypes.data, a.shape[1], nsteps, q.ctypes.data, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            self._check(rc)

    def set_solver(self, max_newton: int = 8, max_linesearch: int = 8) -> None:
        self._check(self._lib.nmf_set_solver(self._h, int(max_newton), int(max_linesearch)))

    def set_precision(self, bits: int = 32) -> None:
        """Arithmetic of the step kernel: 32 (default, what the reference's ``GPUSimulation`` computes in) or 64 (the same kernel
        source in double precision, what the reference's CPU ``Simulation`` computes in: a validation path that shadows the fp64
        oracle).  The tensors of this class stay float32; the library keeps full-precision records between launches and picks
        up anything written to ``state`` / ``qpos`` / ``ctrl`` in between."""
        self._check(self._lib.nmf_set_precision(self._h, int(bits)))

    def set_schedule(self, sub_steps: int = -1) -> None:
        """Steps per work item of multi-step launches (0 = one block per fly for the whole launch, -1 = automatic)."""
        self._check(self._lib.nmf_set_schedule(self._h, int(sub_steps)))

    def set_flies_per_block(self, fpb: int = 0) -> None:
        """Fly slots per thread block of the float32 kernels (1, 2, 4, 8; 0 = chosen per launch); results do not depend on it."""
        self._check(self._lib.nmf_set_flies_per_block(self._h, int(fpb)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.nmf_launch_count(self._h))

    # ---- rendering hooks kept for API compatibility ---------------------------
    def set_renderer(self, *args, **kwargs):
        raise NotImplementedError("rendering is outside the step path (SURVEY.md section 2, rows 13-14); export_state(world_id) returns the "
                                  "qpos / qvel / time a MuJoCo renderer needs (what WarpCPURenderer pulls with mjw.get_data_into)")

    def render_as_needed(self) -> bool:
        return False if self.renderer is None else self.renderer.render_as_needed(self)

    def render_as_needed_with_profile(self) -> bool:
        t0 = perf_counter_ns()
        done = self.render_as_needed()
        self._total_render_time_ns += perf_counter_ns() - t0
        self._frames_rendered += int(done)
        return done

    def print_performance_report(self) -> None:
        n = max(1, self._curr_step)
        per = self._total_physics_time_ns / n / 1e3
        print(f"physics: {self._curr_step} steps x {self.n_worlds} worlds, {per:.1f} us/step, "
              f"{self.n_worlds * 1e6 / max(per, 1e-9):.0f} env-steps/s")
