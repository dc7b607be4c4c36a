"""Long-run stability: 4096 flies walking for many steps (fp32 kernel); reports finiteness and how the population looks at the end."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from flygym_b200 import B200Simulation, NMFModel
from flygym_b200.actions import cpg_table

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
out = {}
for name, model in (("flat", NMFModel.bench(True)), ("blocks", NMFModel.bench(True, terrain="blocks")), ("gapped", NMFModel.bench(True, terrain="gapped"))):
    n, T = 4096, 2500
    sim = B200Simulation(model, n_worlds=n, outputs=True)
    table = torch.from_numpy(cpg_table(model, n, T)).cuda()
    sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))
    sim.warmup()
    t0 = 0
    for _ in range(steps // 500):
        sim.step(500, table, t0); t0 = (t0 + 500) % T
    torch.cuda.synchronize()
    q = sim.qpos
    finite = bool(torch.isfinite(sim.state).all())
    thorax = model.names["segments"].index("c_thorax")
    z = sim.seg_xpos[:, thorax, 2]
    quat = sim.seg_xquat[:, thorax]
    up = 1 - 2 * (quat[:, 1] ** 2 + quat[:, 2] ** 2)
    out[name] = {"steps": steps + 500, "finite": finite, "thorax_z_mm_min_med_max": [float(z.min()), float(z.median()), float(z.max())],
                 "upright_fraction": float((up > 0.5).float().mean()), "x_displacement_mm_med": float(q[:, 0].median()),
                 "max_abs_qvel": float(sim.qvel.abs().max()), "time_s": sim.time}
    print(name, out[name], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/soak.json").write_text(json.dumps(out, indent=1))
