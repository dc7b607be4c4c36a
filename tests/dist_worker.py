"""world_size-2 gloo worker: shard bookkeeping + gather used by bench.py / flygym_b200.dist (CPU only)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist
from flygym_b200.dist import shard_range, gather_to_rank0, max_over_ranks

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_total = 10
lo, hi = shard_range(n_total, rank, world)
assert (lo, hi) == ((0, 5) if rank == 0 else (5, 10))
local = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, 3)
full = gather_to_rank0(local)
if rank == 0:
    assert full.shape == (n_total, 3) and torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float32))
t = max_over_ranks(float(rank + 1))
assert t == float(world)
# TrajectoryRecorder over the process group (a stand-in simulation with CPU tensors: the recorder only touches these attributes)
from types import SimpleNamespace
from flygym_b200.trajectory import TrajectoryRecorder
n_local = hi - lo
info = SimpleNamespace(nq=73, nv=72, off_time=300)
state = torch.zeros((n_local, 304))
fake = SimpleNamespace(n_worlds=n_local, info=info, device=torch.device("cpu"), state=state,
                       qpos=state[:, 0:73], qvel=state[:, 76:148])
rec = TrajectoryRecorder(fake, capacity=3, every=1, with_qvel=True)
for step in range(3):
    state[:, 0] = torch.arange(lo, hi, dtype=torch.float32) + 100 * step      # qpos[:, 0] = global fly id + 100 * snapshot
    state[:, 76] = -state[:, 0]
    state[:, 300] = 1e-4 * (step + 1)
    assert rec.record()
data = rec.gather()
if rank == 0:
    assert data["qpos"].shape == (3, n_total, 73) and data["qvel"].shape == (3, n_total, 72)
    for step in range(3):
        assert np.array_equal(data["qpos"][step, :, 0], np.arange(n_total) + 100 * step)
        assert np.array_equal(data["qvel"][step, :, 0], -(np.arange(n_total) + 100 * step))
    assert np.allclose(data["time"], [1e-4, 2e-4, 3e-4])
else:
    assert data is None
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
