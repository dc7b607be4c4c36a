"""Config 4 kernel measurement: Retina transform for 1024 flies (two 512x450 RGB eye buffers each)."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200.retina import Retina

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
ret = Retina()
n = args.n
g = torch.Generator(device="cuda").manual_seed(0)
# two distinct input sets (2 x 1.4 GB > 126 MB L2) alternated so that no launch finds its input in L2
imgs = [torch.randint(0, 256, (n, 2, ret.H, ret.W, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)]
out = torch.empty((n, 2, ret.n_ommatidia, 2), dtype=torch.float32, device="cuda")
for i in range(3):
    ret(imgs[i % 2], out)
torch.cuda.synchronize()
ev = []
for i in range(args.iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ret(imgs[i % 2], out); b.record(); ev.append((a, b))
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in ev])
alg = n * (2 * ret.H * ret.W * 3 + 2 * ret.n_ommatidia * 2 * 4)
peak = 6574.5
p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
if p.exists():
    peak = json.loads(p.read_text())["hbm_gbs"]
res = {"kernel": "nmf_retina_kernel", "n_flies": n, "ms_per_launch_mean": float(ms.mean()), "ms_min": float(ms.min()),
       "fly_frames_per_s": n / (ms.mean() * 1e-3), "alg_bytes": alg, "achieved_GBps": alg / (ms.mean() * 1e-3) / 1e9,
       "peak_GBps": peak, "frac": alg / (ms.mean() * 1e-3) / 1e9 / peak}
if args.cpu:
    from oracle.retina_oracle import retina_oracle
    h = imgs[0][:8].cpu().numpy()
    t = time.perf_counter(); retina_oracle(h, ret.id_map, ret.pale); dt = time.perf_counter() - t
    res["cpu_oracle_fly_frames_per_s_1core"] = 8 / dt
print(json.dumps(res))
