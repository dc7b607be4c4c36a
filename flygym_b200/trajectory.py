"""Trajectory recording for batched runs: snapshots stay on the device while the simulation runs and
are gathered over the process group (NCCL on GPUs, gloo in the CPU tests) only when the recording is closed."""
from __future__ import annotations

import numpy as np
import torch

from .dist import gather_to_rank0


class TrajectoryRecorder:
    """Keeps ``qpos`` (and optionally ``qvel``) of every fly of this rank every ``every`` calls to :meth:`record`.

    ``gather()`` returns, on rank 0, ``{"time": (S,), "qpos": (S, n_flies_total, nq)[, "qvel": ...]}`` with the flies of all ranks
    concatenated in rank order (= global fly order with ``dist.shard_range``); other ranks get ``None``."""

    def __init__(self, sim, capacity: int, *, every: int = 1, with_qvel: bool = False):
        self.sim, self.every, self.with_qvel = sim, max(1, int(every)), bool(with_qvel)
        n, i = sim.n_worlds, sim.info
        self.qpos = torch.empty((capacity, n, i.nq), dtype=torch.float32, device=sim.device)
        self.qvel = torch.empty((capacity, n, i.nv), dtype=torch.float32, device=sim.device) if with_qvel else None
        self.time = torch.empty((capacity,), dtype=torch.float32, device=sim.device)
        self.count = 0
        self._calls = 0

    def record(self) -> bool:
        self._calls += 1
        if (self._calls - 1) % self.every or self.count >= self.qpos.shape[0]:
            return False
        k = self.count
        self.qpos[k].copy_(self.sim.qpos)              # device-to-device, no host sync
        if self.qvel is not None:
            self.qvel[k].copy_(self.sim.qvel)
        self.time[k].copy_(self.sim.state[0, self.sim.info.off_time])
        self.count += 1
        return True

    def gather(self):
        k = self.count
        out = {"time": self.time[:k].cpu().numpy()}
        root = True
        for name, buf in (("qpos", self.qpos), ("qvel", self.qvel)):
            if buf is None:
                continue
            full = gather_to_rank0(buf[:k].transpose(0, 1).contiguous())       # (n_local, S, d) slabs, fly-major for the gather
            if full is None:
                root = False            # every rank takes part in every collective; only rank 0 keeps the result
                continue
            out[name] = full.transpose(0, 1).cpu().numpy()
        return out if root else None

    def save(self, path) -> bool:
        data = self.gather()
        if data is None:
            return False
        np.savez_compressed(path, **data)
        return True
