"""One short run of the general-topology kernel for ncu (tools/tree_bench.py measures; a number taken under a profiler is not a bench value)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
preset = sys.argv[3] if len(sys.argv) > 3 else "all_biological"
m = NMFModel.bench(True, joint_preset=preset)
sim = B200Simulation(m, n_worlds=n, outputs=False)
tab = torch.as_tensor(cpg_table(m, n, steps)).cuda().contiguous()
q = sim.qpos.clone(); q[:, 2] = -0.17; sim.qpos.copy_(q)
sim.step(300)
for _ in range(3):
    sim.step(steps, tab, 0)
torch.cuda.synchronize()
