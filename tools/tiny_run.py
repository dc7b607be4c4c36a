import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from flygym_b200 import B200Simulation, NMFModel
np.set_printoptions(linewidth=250, precision=4)
m = NMFModel.bench(True)
sim = B200Simulation(m, n_worlds=2, outputs=False, debug=True)
q = m.arrays["key_qpos"].copy(); q[2] = -0.15
sim.qpos.copy_(torch.as_tensor(np.tile(q, (2, 1)), dtype=torch.float32))
sim.step(1)
torch.cuda.synchronize()
s = sim.state[0].cpu().numpy(); d = sim.debug[0].cpu().numpy()
regions = dict(FS=4, QACC=76, FC=148, QACCE=220, CON=292, XPOS=292+768, CDOF=292+768+192, HROWS=292+768+192+432)
bad = np.where(~np.isfinite(d))[0]
for name, start in regions.items():
    ends = sorted(v for v in regions.values() if v > start)
    end = ends[0] if ends else len(d)
    idx = bad[(bad >= start) & (bad < end)] - start
    print(name, 'nonfinite rel idx', idx[:60], 'n', len(idx))
print('head', d[:4])
print('qacc', d[76:76+72])
print('qacce', d[220:220+72])
print('fs', d[4:76])
print('fc', d[148:220])
h = d[292+768+192+432:]
print('hrows leg0', h[:177].reshape(-1)[:176].reshape(11,16))
print('hbb', h[6*177:6*177+21])
