#!/usr/bin/env python
"""bench.py — env-steps/s of the batched NeuroMechFly step path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one physics timestep of every fly of the batch.  Workload at N=1 =
BASELINE.json configs[1]: 4096 parallel flies, flat terrain, sinusoidal CPG tripod
actions (flygym_b200/actions.py), adhesion on, no vision.  Multi-GPU (torchrun, one
rank per GPU): every rank steps its own 4096 flies, no data-path collective (flies
are independent); NCCL is used for the barrier, the max-over-ranks time and one
gather of a per-fly metrics slab -> "scaling": "weak".

`--impl reference` times the CPU restatement of the reference's mj_step path
(oracle/, kind "port": real MuJoCo cannot be installed here) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_FLIES_PER_GPU = 4096
ALG_BYTES_CORE = 1928          # read qpos+qvel+ctrl+warm, write qpos+qvel+warm (f32), per fly-step
METRIC = "env-steps/sec (batched flies)"
UNIT = "env-steps/s"
TABLE_T = 2500                 # 3 exact CPG periods at 12 Hz, dt = 1e-4


def workload_config(n_flies, chunk, simplify):
    return {
        "workload": f"{n_flies} NeuroMechFly per GPU, flat terrain, CPG tripod gait (12 Hz sinusoids), adhesion on, "
                    f"{'capsule' if simplify else 'mesh-hull'} collision geoms, no vision",
        "n_flies_per_gpu": n_flies, "nv": 72, "nu": 48, "timestep": 1e-4,
        "steps_per_launch": chunk, "l2": "flushed (256 MiB write) before every timed launch; action table 1.7 GB > L2",
    }


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------- CPU arm
def cpu_arm(model, table64, steps_per_thread, threads):
    """Oracle (CPU restatement of mj_step) on `threads` host threads, one fly each, `steps_per_thread` steps
    after a 500-step warm-up (mirrors Simulation.warmup, simulation.py:298-309).  Returns env-steps/s."""
    from oracle.oracle import Oracle
    oracles = [Oracle(model) for _ in range(threads)]
    for k, o in enumerate(oracles):
        o.ctrl[model.dim("nu_pos"):] = 1.0
        o.step(500)
    def run(k):
        oracles[k].step_table(table64[k % table64.shape[0], :steps_per_thread])
    ths = [threading.Thread(target=run, args=(k,)) for k in range(threads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    return threads * steps_per_thread / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    from flygym_b200.actions import cpg_table
    from flygym_b200.model import NMFModel
    model = NMFModel.bench(simplify_geom=not args.mesh)
    cores = os.cpu_count() or 1
    sample_steps = 200                      # per thread per "step" of this arm: bounded sample of the workload
    table = cpg_table(model, cores, sample_steps, n_flies_total=N_FLIES_PER_GPU).astype(np.float64)
    for _ in range(args.warmup):
        cpu_arm(model, table, 50, cores)
    vals, tot = [], 0.0
    for _ in range(args.steps):
        v, dt = cpu_arm(model, table, sample_steps, cores)
        vals.append(v); tot += dt
    value = float(np.mean(vals))
    sample = f"{cores} threads x 1 fly x {sample_steps} steps per timed step (CPG actions, after 500 warm-up steps)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(N_FLIES_PER_GPU, 1, not args.mesh),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU fp64 restatement of the reference's mujoco.mj_step path (oracle/nmf_oracle.c); real MuJoCo 3.6.0 is not installable here",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from flygym_b200 import B200Simulation, NMFModel
    from flygym_b200.actions import cpg_table

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n_flies
    simplify = not args.mesh
    model = NMFModel.bench(simplify_geom=simplify)
    sim = B200Simulation(model, n_worlds=n, device=dev, outputs=False)
    if args.actions == "replay":
        from flygym_b200.actions import replay_table
        table_np = replay_table(model, n, 1000, fly_offset=rank * n)      # sim_steps = 1000 as run_gpu_benchmark.py
    else:
        table_np = cpg_table(model, n, TABLE_T, fly_offset=rank * n, n_flies_total=world * n)
    table_T = table_np.shape[1]
    table = torch.from_numpy(table_np).to(dev)
    sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))     # as the reference benchmark (time_gpu_simulation.py:130)
    sim.warmup()                                                        # 500 steps at the neutral pose
    chunk = max(1, min(args.chunk, args.steps))
    t0 = 0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(max(3, args.warmup)):
        sim.step(chunk, table, t0); t0 = (t0 + chunk) % table_T
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- timed region: exactly K steps, CUDA events on the launching stream around every launch
    launches0 = sim.launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    ev = []
    done = 0
    while done < args.steps:
        c = min(chunk, args.steps - done)
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sim.step(c, table, t0); b.record()
        ev.append((a, b, c)); t0 = (t0 + c) % table_T; done += c
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    n_launch = sim.launch_count - launches0
    t = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the public API with HOST buffers, every step: H2D actions, step, D2H qpos
    e2e_steps = min(args.steps, 200)
    act_host = torch.from_numpy(np.ascontiguousarray(table_np[:, :e2e_steps].transpose(1, 0, 2))).pin_memory()
    qpos_host = torch.empty((n, model.nq), dtype=torch.float32).pin_memory()
    for s in range(3):
        sim.step_host(act_host[s].numpy(), 1, qpos_host.numpy())
    barrier()
    w0 = time.perf_counter()
    for s in range(e2e_steps):
        sim.step_host(act_host[s].numpy(), 1, qpos_host.numpy())
    barrier()
    e2e_s = time.perf_counter() - w0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())

    # ---- metrics slab gathered over NCCL (the only collective of the path)
    slab = torch.stack([sim.qpos[:, 0], sim.qpos[:, 1], sim.qpos[:, 2], sim.qvel[:, 0]], dim=1).contiguous()
    if world > 1:
        gathered = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
        dist.gather(slab, gathered, dst=0)
        if rank == 0:
            slab = torch.cat(gathered)
    finite = bool(torch.isfinite(slab).all().item())

    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        per_launch_steps = chunk
        avg_launch_ms = kernel_ms / max(1, len(ev))
        achieved = ALG_BYTES_CORE * n * per_launch_steps / (avg_launch_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            cs = 10000
            tb = cpg_table(model, cores, cs, n_flies_total=n).astype(np.float64)
            v, dt = cpu_arm(model, tb, cs, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cores} threads x 1 fly x {cs} CPG steps (oracle/nmf_oracle.c, fp64), {dt:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload_config(n, chunk, simplify), actions=args.actions),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * model.dim("nu_pos") * 4),
                    "d2h_bytes_per_step": int(n * model.nq * 4), "steps": e2e_steps},
            "gpu_launches": int(n_launch),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": peak_kind,
                         "note": "the fused step is FP32-issue/latency bound, not HBM bound (SURVEY.md 8d); algorithmic bytes "
                                 f"= {ALG_BYTES_CORE} B per fly-step"},
            "cpu_baseline": cpu,
            "wall_s_timed_region": wall, "state_finite": finite,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-flies", type=int, default=N_FLIES_PER_GPU)
    ap.add_argument("--chunk", type=int, default=100, help="physics steps fused per kernel launch")
    ap.add_argument("--mesh", action="store_true", help="mesh-hull collision geoms (simplify_geom=False)")
    ap.add_argument("--actions", default="cpg", choices=["cpg", "replay"], help="cpg = BASELINE config 2; replay = the reference benchmark's kinematic-replay clip")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
