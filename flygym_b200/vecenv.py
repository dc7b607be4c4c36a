"""Vector-environment wrapper around :class:`B200Simulation`.

FlyGym 2.x dropped the Gymnasium interface of v1 (reference ``docs/migration.md:25``) and names large-batch RL as
the motivation of its GPU backend (``tutorials/3_gpu_accelerated_simulation.ipynb`` cell 0).  This is the thin
caller that turns the step path into that loop: ``n_envs`` flies advance in lock-step, an action is held for
``physics_steps_per_action`` physics steps fused into ONE kernel launch, observations stay on the device, and
flies that terminate are reset individually (masked reset; the reference can only re-upload every world).
The method set follows ``gymnasium.vector.VectorEnv`` (``reset`` / ``step`` returning
``obs, reward, terminated, truncated, info``) without importing it.
"""
from __future__ import annotations

import numpy as np
import torch

from .anatomy import ActuatorType
from .model import NMFModel
from .simulation import B200Simulation


class NMFVectorEnv:
    """Args:
        model: baked model (default: the reference benchmark model on flat ground).
        n_envs: flies on this GPU.
        physics_steps_per_action: physics steps per ``step()`` (decimation); the action is held.
        episode_steps: ``truncated`` after this many ``step()`` calls (0 = never).
        odor: optional ``(source_positions, peak_intensities)`` to add an ``"odor"`` observation.
        vision: add an ``"vision"`` observation ``(n, 2, 721, 2)`` (fused eye-camera + Retina kernel).
        auto_reset: reset terminated / truncated flies at the end of ``step()`` (masked, on the device, no host sync);
            ``step`` still returns their terminal observation and ``info["reset_mask"]`` marks them.
    Action: ``(n_envs, 42)`` position targets, or ``(n_envs, 48)`` with the six leg-adhesion inputs appended.
    Reward (default, override ``compute_reward``): forward displacement of the thorax along +x in mm."""

    def __init__(self, model: NMFModel | None = None, n_envs: int = 4096, *, physics_steps_per_action: int = 10,
                 episode_steps: int = 0, odor=None, vision: bool = False, auto_reset: bool = True, device=None):
        self.sim = B200Simulation(model, n_worlds=n_envs, device=device, outputs=True)
        self.num_envs = int(n_envs)
        self.k = int(physics_steps_per_action)
        if self.k < 1:
            raise ValueError("physics_steps_per_action must be >= 1")
        self.episode_steps = int(episode_steps)
        self.auto_reset = bool(auto_reset)
        self.fly_name = next(iter(self.sim.world.fly_lookup))
        info = self.sim.info
        self.action_dim = info.nu_pos
        self._all_cols = torch.arange(0, info.nu_pos + info.nu_adh, dtype=torch.int32, device=self.sim.device)
        self._thorax = self.sim.model.names["segments"].index("c_thorax")
        self._elapsed = torch.zeros(self.num_envs, dtype=torch.int32, device=self.sim.device)
        self._last_x = torch.zeros(self.num_envs, dtype=torch.float32, device=self.sim.device)
        # thorax x at the keyframe (c_thorax sits inside the free-joint frame: reference assets rigging.yaml:1-4; keyframe quat = identity)
        m = self.sim.model
        self._key_thorax_x = float(m.arrays["key_qpos"][0] + m.arrays["seg_pos"].reshape(-1, 3)[self._thorax][0])
        self._odor = self._eyes = None
        if odor is not None:
            from .retina import OdorSensor
            self._odor = OdorSensor(self.sim, *odor)
        if vision:
            from .retina import EyeCameras
            self._eyes = EyeCameras(self.sim)

    # ------------------------------------------------------------------ observations
    def _observe(self) -> dict:
        sim, name = self.sim, self.fly_name
        found, force, _, _, _, _ = sim.get_ground_contact_info(name)
        obs = {
            "joint_angles": sim.get_joint_angles(name),                 # (n, 66)
            "joint_velocities": sim.get_joint_velocities(name),         # (n, 66)
            "thorax_position": sim.seg_xpos[:, self._thorax].clone(),   # (n, 3)
            "thorax_rotation": sim.seg_xquat[:, self._thorax].clone(),  # (n, 4) wxyz
            "contact_active": found > 0,                                # (n, 6)
            "contact_forces": force,                                    # (n, 6, 3)
        }
        if self._odor is not None:
            obs["odor"] = self._odor()
        if self._eyes is not None:
            obs["vision"] = self._eyes.retina()
        return obs

    def compute_reward(self, obs: dict) -> torch.Tensor:
        x = obs["thorax_position"][:, 0]
        r = x - self._last_x
        self._last_x = x.clone()
        return r

    def compute_terminated(self, obs: dict) -> torch.Tensor:
        """Non-finite state, or the thorax has rolled onto its back (its z axis points down)."""
        q = obs["thorax_rotation"]
        up_z = 1.0 - 2.0 * (q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2])
        bad = ~torch.isfinite(self.sim.state).all(dim=1)
        return bad | (up_z < 0.0)

    # ------------------------------------------------------------------ gymnasium.vector-style API
    def reset(self, mask=None):
        """Reset every fly (or those selected by the boolean ``mask``) to the neutral keyframe; returns ``(obs, info)``."""
        dev = self.sim.device
        m = None if mask is None else torch.as_tensor(mask, device=dev).bool()
        self.sim.reset(m)
        self.sim.forward()                       # mj_forward: poses / sensors of the new state, nothing advances
        obs = self._observe()
        x = obs["thorax_position"][:, 0]
        if m is None:
            self._elapsed.zero_(); self._last_x = x.clone()
        else:
            self._elapsed[m] = 0; self._last_x = torch.where(m, x, self._last_x)
        return obs, {}

    def step(self, actions):
        sim = self.sim
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float32))
        if actions.ndim != 2 or actions.shape[0] != self.num_envs or actions.shape[1] not in (self.action_dim, self.action_dim + sim.info.nu_adh):
            raise ValueError(f"actions must have shape ({self.num_envs}, {self.action_dim}) or ({self.num_envs}, {self.action_dim + sim.info.nu_adh})")
        if actions.shape[1] == self.action_dim:
            sim.set_actuator_inputs(self.fly_name, ActuatorType.POSITION, actions)
        else:
            src = sim._as_device(actions, actions.shape[1], "inputs")
            sim._check(sim._lib.nmf_scatter_ctrl(sim._h, sim._ptr(src), sim._ptr(self._all_cols), int(actions.shape[1]), sim._stream()))
        sim.step(self.k)
        self._elapsed += 1
        obs = self._observe()
        reward = self.compute_reward(obs)
        terminated = self.compute_terminated(obs)
        truncated = (self._elapsed >= self.episode_steps) if self.episode_steps > 0 else torch.zeros_like(terminated)
        info = {}
        done = terminated | truncated
        if self.auto_reset:
            info["reset_mask"] = done
            # masked reset is a no-op kernel for flies whose byte is 0; no host sync is needed to decide
            sim.reset(done)
            self._elapsed = torch.where(done, torch.zeros_like(self._elapsed), self._elapsed)
            self._last_x = torch.where(done, torch.full_like(self._last_x, self._key_thorax_x), self._last_x)
        return obs, reward, terminated, truncated, info

    def close(self):
        self.sim = None
