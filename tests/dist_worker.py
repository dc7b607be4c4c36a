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
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
