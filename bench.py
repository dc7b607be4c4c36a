#!/usr/bin/env python
"""bench.py — env-steps/s of the batched NeuroMechFly step path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one physics timestep of every fly of the batch.  Default workload
(`--workload flat`) = BASELINE.json configs[1]: 4096 parallel flies, flat terrain,
sinusoidal CPG tripod actions (flygym_b200/actions.py), adhesion on, no vision.
The other BASELINE configs are selectable and print the same JSON line:
  --workload terrain    configs[2]: 4096 flies on blocks / gapped terrain, stance-phase adhesion
  --workload vision     configs[3]: 1024 flies, two eye-camera renders -> Retina every step
  --workload olfaction  configs[4]: 32768 flies per GPU, flat + odor sensors every step
Multi-GPU (torchrun, one rank per GPU): every rank steps its own flies, no data-path
collective (flies are independent); NCCL is used for the barrier, the max-over-ranks
time and the gathers of a per-fly metrics slab -> "scaling": "weak".

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


ALG_BYTES_OBS = 4940           # the same + the full observation set written every step
RETINA_ALG_BYTES = 2 * 512 * 450 * 3 + 2 * 721 * 2 * 4   # 1 393 936 B per fly-frame
# DRAM bytes of the dominant kernel per fly-step (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture on
# B200 divided by the fly-steps of the captured launch; profiles/ncu_*_summary.txt).  roofline.traffic = this x the fly-steps of one
# launch of the run that prints it, and traffic_source names the capture it was scaled from.
NCU_TRAFFIC_PER_FLY_STEP = {
    "flat": ((0.1738e9 + 3.2041e9) / (4096 * 100), "profiles/ncu_step_r02f_summary.txt (nmf_step_x8_kernel, 4096 flies x 100 steps: 0.174 GB read + 3.20 GB written, "
                                                   "the writes being local-memory spill lines evicted from L2, vs 0.79 GB algorithmic)"),
    "terrain": ((0.0896e9 + 0.5587e9) / (4096 * 100), "profiles/ncu_step_terrain_r01_summary.txt (4096 flies x 100 steps, 80-register build)"),
    "olfaction": ((0.0774e9 + 0.6436e9) / 32768, "profiles/ncu_step_olfaction_r01_summary.txt (one 1-step launch of 32768 flies with outputs)"),
}
VISION_TRAFFIC = (2.92e6 / 1024, "profiles/ncu_eye_body_r02c_summary.txt: the fused kernel reads 2.9 MB per 1024 flies (run table, poses, body capsules) and never materialises the eye buffers")
RETINA_BUFFERS_TRAFFIC = (1.0096e9 + 6.9e6, "profiles/ncu_retina_r02_summary.txt: 1.010 GB read + 6.9 MB written = 0.71 x algorithmic (chunks outside the hexagon skipped)")
# what actually bounds the step kernel (same captures): issue-slot utilisation and the dominant stall reason
NCU_LIMITER = {
    "flat": {"issue_slots_busy": 0.542, "top_stall": "barrier 24.5 % (lockstep passes of the 8 flies of a block + the fly's own barriers), short scoreboard 24.5 % "
             "(shuffles / shared memory), long scoreboard 16.3 % (register spills); no_inst 0.8 % (46 % before the flies of a block shared their fetches)",
             "warp_instructions_per_fly_step": 23330, "issue_ceiling_env_steps_per_s": 148 * 4 * 1.965e9 / 23330,
             # executed FP32 work of the same capture: (2 FFMA + FADD + FMUL) warp instructions x 27.85 active lanes / 409600 fly-steps
             "fp32_flop_per_fly_step": 453000, "fp32_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12,
             "source": "profiles/ncu_step_r02f_summary.txt"},
    "terrain": {"issue_slots_busy": 0.314, "top_stall": "no_inst (instruction fetch) 56 % of samples", "source": "profiles/ncu_step_terrain_r01_summary.txt"},
    "olfaction": {"issue_slots_busy": 0.509, "top_stall": "no_inst (instruction fetch)", "source": "profiles/ncu_step_olfaction_r01_summary.txt"},
    "vision": {"issue_slots_busy": 0.698, "top_stall": "issue-bound: 1.11 G warp instructions per 1024 flies (fused eye + Retina kernel: shading and body raster in about "
               "equal parts, the explicitly rounded FADD / FMUL of the capsule hit test being the largest single item), 3 blocks per SM",
               "source": "profiles/ncu_eye_body_r02c_summary.txt"},
}
TREE_TRAFFIC_PER_FLY_STEP = (8.4e6 / (1480 * 20), "profiles/ncu_tree_r02_summary.txt (nmf_tree_step_kernel, ALL_BIOLOGICAL, 1480 flies x 20 steps: 8.4 MB read + 2 KB written = "
                                    "the records and the model tables once; nothing spills)")
CONFIG5_TOTAL_FLIES = 262144   # BASELINE configs[4]: 262144 flies over 8 GPUs
ODOR_SOURCES = [[12.0, 4.0, 1.5], [12.0, -4.0, 1.5]]     # config 5: 2 sources x 2 odor dimensions, fixed constants
ODOR_PEAKS = [[1.0, 0.0], [0.0, 1.0]]


SKELETON_DIMS = {"legs_only": (72, 48), "all_biological": (132, 48), "all_possible": (210, 78)}      # nv, nu of the baked models


def workload_config(args, n_flies, chunk):
    geoms = "mesh-hull" if (args.mesh and args.workload != "terrain") else "capsule"
    what = {
        "flat": "flat terrain, CPG tripod gait (12 Hz sinusoids), adhesion on, no vision",
        "terrain": f"{args.terrain} terrain (box columns), CPG tripod gait, adhesion 100 in stance / 1 in swing, no vision",
        "vision": "flat terrain, CPG tripod gait, adhesion on, two 512x450 eye-camera renders (checker ground, sky" + (", the fly's own body" if args.eye_body == "on" else "") + ") -> 721-ommatidia Retina " + ("after EVERY step" if getattr(args, "vision_every", 1) <= 1 else f"after every {args.vision_every}th step (vision refresh at {1e4 / args.vision_every:.0f} Hz, the v1 default)"),
        "olfaction": "flat terrain, CPG tripod gait, adhesion on, 4 odor sensors x 2 sources x 2 odor dims after every step",
    }[args.workload]
    cfg = {
        "workload": f"{n_flies} NeuroMechFly per GPU, {what}, {geoms} collision geoms",
        "baseline_config": {"flat": 1, "terrain": 2, "vision": 3, "olfaction": 4}[args.workload],
        "n_flies_per_gpu": n_flies, "nv": SKELETON_DIMS[args.skeleton][0], "nu": SKELETON_DIMS[args.skeleton][1], "timestep": 1e-4,
        "l2": "flushed (256 MiB write) before every timed launch group; action table > L2",
    }
    if args.skeleton != "legs_only":
        cfg["workload"] += f", JointPreset.{args.skeleton.upper()} skeleton ({SKELETON_DIMS[args.skeleton][0] - 6} hinge DoFs; general-topology kernels)"
        cfg["skeleton"] = args.skeleton
    if args.workload in ("flat", "terrain"):
        cfg["steps_per_launch"] = chunk
    else:
        ve = getattr(args, "vision_every", 1) if args.workload == "vision" else 1
        cfg["steps_per_timed_group"] = chunk; cfg["launches_per_step"] = 2 / ve
        if ve > 1:
            cfg["physics_steps_per_vision_frame"] = ve
    return cfg


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class StdoutToStderr:
    """NCCL prints its version banner to stdout when the communicator is created; the contract is ONE JSON line on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Rows from here on count (the GPU is under load from this point to stop())."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        rows = self.rows[self.first:] or self.rows[-1:]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------- CPU arm
class CpuArm:
    """Oracle (CPU restatement of mj_step) on `threads` host threads, one fly each.  Built once (500-step warm-up per fly,
    mirroring Simulation.warmup, simulation.py:298-309); every run() advances each fly by `steps_per_thread` more steps."""

    def __init__(self, model, threads):
        from oracle.oracle import Oracle
        self.model, self.threads = model, threads
        self.oracles = [Oracle(model, native=True) for _ in range(threads)]    # -O3 -march=native build made on this host
        for o in self.oracles:
            o.ctrl[model.dim("nu_pos"):] = 1.0
            o.step(500)

    def run(self, table64, steps_per_thread, adhesion=None):
        """Returns (env-steps/s, seconds)."""
        def work(k):
            tb = table64[k % table64.shape[0], :steps_per_thread]
            if adhesion is None:
                self.oracles[k].step_table(tb)
            else:   # per-step adhesion inputs ride in the table's trailing columns
                self.oracles[k].step_table_full(np.concatenate([tb, adhesion[k % adhesion.shape[0], :steps_per_thread]], axis=1))
        ths = [threading.Thread(target=work, args=(k,)) for k in range(self.threads)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.perf_counter() - t0
        return self.threads * steps_per_thread / dt, dt


def host_adhesion_table(model, n_flies, n_steps, n_total):
    """stance-phase adhesion inputs of the terrain workload (same definition as device_cpg_table)"""
    from flygym_b200.actions import TRIPOD_PHASE
    t = np.arange(n_steps) * model.timestep
    psi = 2 * np.pi * np.arange(n_flies) / n_total
    legph = np.array([TRIPOD_PHASE[l] for l in model.names["legs"]])
    return np.where(np.sin(2 * np.pi * 12.0 * t[None, :, None] + psi[:, None, None] + legph[None, None, :]) < 0, 100.0, 1.0)


def mujoco_probe():
    """use real MuJoCo for the CPU arm if it is ever importable.  (It has never been: not in the image, not in
    /opt/wheelhouse; composing the reference's model additionally needs dm_control.)"""
    try:
        import mujoco  # noqa: F401
        return True
    except Exception:
        return False


CPU_KIND_NOTE = ("oracle/nmf_oracle.c (fp64 restatement of mj_step: dense M, dense Cholesky per Newton iteration) built -O3 -march=native on "
                 "this host; real MuJoCo (sparse L'DL) is not importable here")


def cpu_baseline_leg(model, n_total, target_s=16.0, stance_adhesion=False):
    """cpu_baseline: the oracle on every host core, sized to ~target_s seconds of CPU work from a short pilot run, plus a
    single-thread figure (BASELINE.md 2.1 quotes the reference per core)."""
    from flygym_b200.actions import cpg_table
    cores = os.cpu_count() or 1
    pilot = 200
    adh = (lambda k, c=cores: host_adhesion_table(model, c, k, n_total)) if stance_adhesion else (lambda k, c=cores: None)
    arm = CpuArm(model, cores)
    tb = cpg_table(model, cores, pilot, n_flies_total=n_total).astype(np.float64)
    v0, _ = arm.run(tb, pilot, adh(pilot))
    cs = int(min(100000, max(500, target_s * v0 / cores)))
    tb = cpg_table(model, cores, cs, n_flies_total=n_total).astype(np.float64)
    v, dt = arm.run(tb, cs, adh(cs))
    one = CpuArm(model, 1)
    c1 = int(min(100000, max(500, 4.0 * v0 / cores)))
    v1, dt1 = one.run(tb[:1, :c1], c1, adh(c1, 1))
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} threads x 1 fly x {cs} CPG steps after a 500-step warm-up, {dt:.1f} s; {CPU_KIND_NOTE}",
            "single_thread": {"value": v1, "unit": UNIT, "sample": f"1 thread x 1 fly x {c1} CPG steps, {dt1:.1f} s"},
            "mujoco_importable": mujoco_probe()}


def bench_model(args):
    from flygym_b200.model import NMFModel
    if args.workload == "terrain":
        return NMFModel.bench(simplify_geom=True, terrain=args.terrain, joint_preset=args.skeleton)
    return NMFModel.bench(simplify_geom=not args.mesh, joint_preset=args.skeleton)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from flygym_b200.actions import cpg_table
    model = bench_model(args)
    cores = os.cpu_count() or 1
    sample_steps = 200                      # per thread per "step" of this arm: bounded sample of the workload
    n = args.n_flies
    table = cpg_table(model, cores, sample_steps, n_flies_total=n).astype(np.float64)
    adh = host_adhesion_table(model, cores, sample_steps, n) if args.workload == "terrain" else None
    arm = CpuArm(model, cores)
    for _ in range(args.warmup):
        arm.run(table, 50, adh)
    vals, tot = [], 0.0
    for _ in range(args.steps):
        v, dt = arm.run(table, sample_steps, adh)
        vals.append(v); tot += dt
    value = float(np.mean(vals))
    sample = f"{cores} threads x 1 fly x {sample_steps} steps per timed step (CPG actions, after 500 warm-up steps)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + "; " + CPU_KIND_NOTE,
                         "mujoco_importable": mujoco_probe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU fp64 restatement of the reference's mujoco.mj_step path (oracle/nmf_oracle.c); real MuJoCo 3.6.0 is not installable "
                "here.  Physics only: the CPU arm has no vision / olfaction leg",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- GPU arm
def device_cpg_table(torch, model, n, T, dev, fly_offset, n_total, adhesion_stance=False, fly_stride=1):
    """cpg_table (flygym_b200/actions.py) evaluated on the device in fp64, in slabs (the 32768-fly table is 13.8 GB).
    With adhesion_stance the table gets 6 extra columns: adhesion ctrl 100 during the stance half-cycle of each leg
    (sin < 0 of the leg's coxa-pitch phase), 1 during swing.
    Local fly i is global fly fly_offset + fly_stride * i (its gait phase is 2 pi * global id / n_total)."""
    from flygym_b200.actions import cpg_parameters, TRIPOD_PHASE
    neutral, amp, phase = cpg_parameters(model)
    ncol = len(neutral) + (6 if adhesion_stance else 0)
    out = torch.empty((n, T, ncol), dtype=torch.float32, device=dev)
    tt = torch.arange(T, dtype=torch.float64, device=dev) * model.timestep
    ne, am, ph = (torch.as_tensor(a, dtype=torch.float64, device=dev) for a in (neutral, amp, phase))
    legph = torch.as_tensor([TRIPOD_PHASE[l] for l in model.names["legs"]], dtype=torch.float64, device=dev)
    slab = max(1, (1 << 26) // (T * ncol))
    for lo in range(0, n, slab):
        hi = min(n, lo + slab)
        psi = 2 * np.pi * (torch.arange(lo, hi, dtype=torch.float64, device=dev) * fly_stride + fly_offset) / n_total
        base = 2 * np.pi * 12.0 * tt[None, :, None] + psi[:, None, None]
        out[lo:hi, :, :len(neutral)] = (ne + am * torch.sin(base + ph)).float()
        if adhesion_stance:
            out[lo:hi, :, len(neutral):] = torch.where(torch.sin(base + legph) < 0, 100.0, 1.0).float()
    return out


class Ctx:
    """what one measurement needs to know about the process: rank layout, device, torch / dist modules"""

    def __init__(self, torch, dist, rank, world, dev):
        self.torch, self.dist, self.rank, self.world, self.dev = torch, dist, rank, world, dev
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(ctx, args, *, steps, warmup, sample_clocks, e2e_cap=200, dominant=True, table_rows=TABLE_T):
    """One workload (args.workload / mesh / precision / n_flies / chunk ...) measured three ways: device-timed `value` over
    exactly `steps` steps, the dominant kernel alone (roofline), and `e2e` through the host-buffer API.  Returns a dict."""
    torch, dist, rank, world, dev = ctx.torch, ctx.dist, ctx.rank, ctx.world, ctx.dev
    from flygym_b200 import B200Simulation
    from flygym_b200.anatomy import ActuatorType
    wl = args.workload
    n = args.n_flies
    model = bench_model(args)
    nu_pos = model.dim("nu_pos")
    per_step = wl in ("vision", "olfaction")          # sensors are evaluated after every physics step
    sim = B200Simulation(model, n_worlds=n, device=dev, outputs=per_step)
    sim.set_precision(args.precision)
    if args.actions == "replay":
        from flygym_b200.actions import replay_table_device
        table = replay_table_device(model, n, 1000, dev, fly_offset=rank * n)                 # sim_steps = 1000 as run_gpu_benchmark.py
    else:
        # rank r steps the global flies r, r + N, r + 2N, ...: the gait phase of a fly is 2 pi * global id / total, so every rank holds
        # the whole gait cycle.  (With contiguous blocks each of 8 ranks held one eighth of the cycle, i.e. all of its flies in the
        # same stance / swing transition at the same time, and the ranks' step times differed by up to 16 % from each other.)
        table = device_cpg_table(torch, model, n, table_rows, dev, rank, world * n, adhesion_stance=(wl == "terrain"), fly_stride=world)
    table_T = table.shape[1]
    sim.set_leg_adhesion_states("nmf", np.ones((n, 6), np.float32))     # as the reference benchmark (time_gpu_simulation.py:130)
    # nvidia-smi takes ~0.1 s to initialise NVML: it is started here and its rows count from the physics warm-up on, so that the timed
    # region (3.6 ms with the driver's --steps 20) is neither disturbed by that start-up nor over before the first sample
    sampler = ClockSampler(dev.index) if (sample_clocks and rank == 0) else None
    if sampler is not None:
        sampler.start(); time.sleep(0.25); sampler.mark()
    sim.warmup()                                                        # 500 steps at the neutral pose
    chunk = max(1, min(args.chunk, steps))

    # ---- the sensors of the workload
    eyes = odor = sens_out = None
    if wl == "vision":
        from flygym_b200.retina import EyeCameras
        eyes = EyeCameras(sim, body=args.eye_body == "on")
        sens_out = torch.empty((n, 2, eyes.ret.n_ommatidia, 2), dtype=torch.float32, device=dev)
    elif wl == "olfaction":
        from flygym_b200.retina import OdorSensor
        odor = OdorSensor(sim, ODOR_SOURCES, ODOR_PEAKS)

    state = {"t0": 0, "launches": 0}
    vis_every = max(1, getattr(args, "vision_every", 1))

    def advance(c):
        """c physics steps (+ the workload's sensors after every step); returns nothing, counts our kernel launches"""
        if not per_step:
            sim.step(c, table, state["t0"]); state["launches"] += 1
            state["t0"] = (state["t0"] + c) % table_T
            return
        k = vis_every if eyes is not None else 1       # physics steps fused into one launch between two sensor evaluations
        left = c
        while left > 0:
            kk = min(k, left)
            sim.step(kk, table, state["t0"]); state["launches"] += 1
            state["t0"] = (state["t0"] + kk) % table_T
            left -= kk
            if eyes is not None:
                eyes.retina(sens_out)
            else:
                state["odor"] = odor()
            state["launches"] += 1

    for _ in range(max(3, warmup)):
        advance(chunk)
    if wl == "olfaction" and world > 1:      # NCCL sets its all-gather channels up on first use (tens of ms): part of the warm-up
        slab = torch.cat([sim.qpos[:, :3], sim.qvel[:, :1], state["odor"].reshape(n, -1)[:, :4]], dim=1).contiguous()
        dist.all_gather([torch.empty_like(slab) for _ in range(world)], slab)
    torch.cuda.synchronize(dev)

    # ---- timed region: exactly K steps, CUDA events on the launching stream around every launch group
    state["launches"] = 0
    ctx.barrier()
    wall0 = time.perf_counter()
    ev = []
    done = 0
    gathered_slabs = 0
    gather_ms = 0.0
    pending = []
    gstream = torch.cuda.Stream(device=dev) if (wl == "olfaction" and world > 1) else None
    while done < steps:
        c = min(chunk, steps - done)
        ctx.flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); advance(c); b.record()
        ev.append((a, b, c)); done += c
        if wl == "olfaction" and world > 1 and done % 100 == 0:      # config 5: metrics slab over NCCL every 100 steps
            # on a side stream: the collective starts when the slowest rank arrives, and the stepping stream must not sit through
            # that skew (at N = 8 it cost 28 ms per gather when the gather was issued on the stepping stream) -- it only joins at the end
            slab = torch.cat([sim.qpos[:, :3], sim.qvel[:, :1], state["odor"].reshape(n, -1)[:, :4]], dim=1).contiguous()
            outl = [torch.empty_like(slab) for _ in range(world)]
            ready = torch.cuda.Event(); ready.record()
            with torch.cuda.stream(gstream):
                gstream.wait_event(ready)
                ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ga.record(); dist.all_gather(outl, slab); gb.record()
            pending.append((ga, gb, outl, slab)); gathered_slabs += 1
    if pending:        # the stepping stream joins the last gather: whatever of it is still exposed is part of the timed region
        ja, jb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ja.record(); torch.cuda.current_stream(dev).wait_stream(gstream); jb.record(); ev.append((ja, jb, 0))
    ctx.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler is not None else None
    kernel_ms = sum(a.elapsed_time(b) for a, b, c in ev if c > 0)
    gather_ms = sum(a.elapsed_time(b) for a, b, c in ev if c == 0)            # what the stepping stream waited for the gathers
    gather_dev_ms = sum(ga.elapsed_time(gb) for ga, gb, _, _ in pending)      # their own duration on the side stream (incl. rank skew)
    n_launch = state["launches"]
    # the all-gathers of config 5 overlap the following launch groups; the stepping stream's wait for the last one is part of the step
    ms_total = ctx.max_ranks(kernel_ms + gather_ms)
    value = world * n * steps / (ms_total * 1e-3)

    # ---- dominant-kernel time for the roofline (per launch, CUDA events around that kernel alone)
    roof = {}
    if not dominant:
        pass
    elif wl == "vision":
        imgs = [eyes.render() for _ in range(2)]       # two 1.4 GB buffer sets alternated: no launch finds its input in L2
        for i in range(3):
            eyes.ret(imgs[i % 2], sens_out)
        tms = []
        for i in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eyes.ret(imgs[i % 2], sens_out); b.record(); tms.append((a, b))
        fms = []
        for i in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eyes.retina(sens_out); b.record(); fms.append((a, b))
        torch.cuda.synchronize(dev)
        ret_ms = float(np.mean([a.elapsed_time(b) for a, b in tms])); fused_ms = float(np.mean([a.elapsed_time(b) for a, b in fms]))
        alg = n * RETINA_ALG_BYTES
        roof = {"kernel": "nmf_eye_retina_kernel (fused eye-camera image formation + Retina)", "ms": fused_ms, "alg_bytes": alg, "fly_steps": n,
                "retina_over_buffers": {"kernel": "nmf_retina_kernel", "ms_per_launch": ret_ms, "achieved": alg / (ret_ms * 1e-3) / 1e9,
                                        "note": "the HBM-bound form of the operator: eye buffers materialised in HBM (two 1.4 GB sets alternated); a pure "
                                                "read stream that skips chunks outside the ommatidia hexagon, hence above the copy-measured peak"},
                "note": f"algorithmic bytes = the Retina operator's {RETINA_ALG_BYTES} B per fly-frame (SURVEY.md 8d); the fused kernel shades "
                        "the pixels in registers and never materialises the eye buffers, so it is bound by instruction issue (70 % busy, 3 blocks per SM), not HBM"}
        del imgs
    else:
        per_launch_steps = 1 if per_step else chunk
        if per_step:
            tms = []
            for i in range(20):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); sim.step(1, table, state["t0"]); b.record(); tms.append((a, b))
            torch.cuda.synchronize(dev)
            step_ms = float(np.mean([a.elapsed_time(b) for a, b in tms]))
        else:
            step_ms = kernel_ms / max(1, len([1 for _, _, c in ev if c > 0]))
        per_fly = ALG_BYTES_OBS if per_step else ALG_BYTES_CORE
        kname = "nmf_step_terrain" if wl == "terrain" else "nmf_step"
        kname += "_f64_kernel" if args.precision == 64 else ("_x8_kernel" if n >= 8 * 148 else "_kernel")
        if args.skeleton != "legs_only":       # general formula of SURVEY.md 8d: 4 (2 nq + 4 nv + nu) bytes per fly-step
            per_fly = 4 * (2 * model.nq + 4 * model.nv + model.nu) + (ALG_BYTES_OBS - ALG_BYTES_CORE if per_step else 0)
            kname = "nmf_tree_step_f64_kernel" if args.precision == 64 else "nmf_tree_step_kernel"
        roof = {"kernel": kname, "ms": step_ms, "alg_bytes": per_fly * n * per_launch_steps, "fly_steps": n * per_launch_steps,
                "note": "the fused step is bound by instruction fetch / FP32 issue, not by HBM (SURVEY.md 8d, DESIGN.md 4.1); algorithmic bytes "
                        f"= {per_fly} B per fly-step"}

    # ---- end-to-end through the public API with HOST buffers, every step: H2D actions, step (+ sensors), D2H result
    e2e_steps = min(max(steps, 100), e2e_cap)     # a steady-state figure: at least 100 calls (the driver's --steps 20 would time 5 ms)
    act_cols = table.shape[2] if wl == "terrain" else nu_pos        # terrain: the six adhesion inputs travel with the position targets
    act_host = table[:, :e2e_steps, :act_cols].permute(1, 0, 2).contiguous().cpu().pin_memory()
    if not per_step:
        res_host = torch.empty((n, model.nq), dtype=torch.float32).pin_memory()
        act_rows, res_np = [act_host[s].numpy() for s in range(act_host.shape[0])], res_host.numpy()     # views of the pinned buffers
        def e2e_step(s):
            sim.step_host(act_rows[s], 1, res_np)
    else:
        res_dev = sens_out if eyes is not None else state["odor"]
        res_host = torch.empty(res_dev.shape, dtype=torch.float32).pin_memory()
        def e2e_step(s):
            sim.set_actuator_inputs("nmf", ActuatorType.POSITION, act_host[s])     # H2D inside
            sim.step()
            if eyes is not None and (s + 1) % vis_every:
                return                                                              # no vision frame after this step: nothing to read back
            r = eyes.retina(sens_out) if eyes is not None else odor()
            res_host.copy_(r, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
    for s in range(10):
        e2e_step(s)
    ctx.barrier()
    w0 = time.perf_counter()
    for s in range(e2e_steps):
        e2e_step(s)
    ctx.barrier()
    e2e_value = world * n * e2e_steps / ctx.max_ranks(time.perf_counter() - w0)

    # ---- metrics slab gathered over NCCL (the only collective of the path)
    slab = torch.stack([sim.qpos[:, 0], sim.qpos[:, 1], sim.qpos[:, 2], sim.qvel[:, 0]], dim=1).contiguous()
    if world > 1:
        gathered = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
        dist.gather(slab, gathered, dst=0)
        if rank == 0:
            slab = torch.cat(gathered)
    finite = bool(torch.isfinite(slab).all().item())
    out = {"value": value, "ms_total": ms_total, "ms_per_step": ms_total / steps, "steps": steps, "chunk": chunk, "n": n, "model": model,
           "clocks": clocks, "launches": int(n_launch), "roof": roof, "wall": wall, "finite": finite, "gathered_slabs": gathered_slabs,
           "gather_ms": gather_ms, "gather_dev_ms": gather_dev_ms, "kernel_ms": kernel_ms,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * act_cols * 4),
                   "d2h_bytes_per_step": int(res_host.numel() * 4) // (vis_every if eyes is not None else 1), "steps": e2e_steps}}
    del sim, table
    torch.cuda.empty_cache()
    return out


def sub_record(args, m, note):
    """A secondary measurement printed inside the main JSON line (same metric and unit)."""
    rec = {"value": m["value"], "unit": UNIT, "ms_per_step": m["ms_per_step"], "steps": m["steps"], "dtype": f"f{args.precision}",
           "config": dict(workload_config(args, m["n"], m["chunk"]), actions=args.actions), "e2e": m["e2e"], "gpu_launches": m["launches"],
           "state_finite": m["finite"], "note": note}
    if m["gathered_slabs"]:
        rec["nccl_all_gathers_in_timed_region"] = m["gathered_slabs"]
        rec["nccl_all_gather_exposed_ms_total"] = m["gather_ms"]; rec["nccl_all_gather_side_stream_ms_total"] = m["gather_dev_ms"]
        rec["step_kernels_ms_total"] = m["kernel_ms"]
    return rec


def run_ours(args, rank, world, local_rank):
    import copy
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with StdoutToStderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()                  # creates the communicator (and prints NCCL's banner) now
    ctx = Ctx(torch, dist, rank, world, dev)
    wl = args.workload
    m = measure(ctx, args, steps=args.steps, warmup=args.warmup, sample_clocks=True)
    n, chunk, model, roof = m["n"], m["chunk"], m["model"], m["roof"]

    # ---- secondary records next to the headline (default flat workload only): the parity-qualified fp64 build and the
    # reference's default mesh-hull geometry at N = 1; BASELINE config 5 (32768 flies per GPU + olfaction + NCCL all-gather of a
    # metrics slab every 100 steps) when several GPUs take part
    extras = {}
    if wl == "flat" and args.extras and args.actions == "cpg" and args.precision == 32 and not args.mesh and args.n_flies == DEFAULT_FLIES["flat"] \
            and args.skeleton == "legs_only":
        if world == 1:
            a2 = copy.copy(args); a2.mesh = True
            extras["mesh"] = sub_record(a2, measure(ctx, a2, steps=300, warmup=3, sample_clocks=False, e2e_cap=100, dominant=False),
                                        "simplify_geom=False: mesh convex hulls (tarsus5 capsules), up to 4 plane-hull contacts per geom")
            a3 = copy.copy(args); a3.precision = 64
            extras["f64"] = sub_record(a3, measure(ctx, a3, steps=200, warmup=3, sample_clocks=False, e2e_cap=100, dominant=False),
                                       "the same kernel source in double precision: the build that stays within 1e-4 of the fp64 oracle for every walking fly")
            a4 = copy.copy(args); a4.workload = "terrain"; a4.chunk = DEFAULT_CHUNK["terrain"]
            extras["config3_terrain"] = sub_record(a4, measure(ctx, a4, steps=300, warmup=3, sample_clocks=False, e2e_cap=100, dominant=False),
                                                   "BASELINE config 3: 4096 flies on the blocks terrain, stance-phase adhesion")
            a7 = copy.copy(args); a7.skeleton = "all_biological"
            extras["all_biological"] = sub_record(a7, measure(ctx, a7, steps=200, warmup=3, sample_clocks=False, e2e_cap=50, dominant=False),
                                                  "JointPreset.ALL_BIOLOGICAL (anatomy.py:418-436): head, proboscis, antennae, eyes, abdomen, wings and halteres "
                                                  "articulated too, 126 hinge DoFs; stepped by the general-topology kernels (csrc/nmf_tree.cuh)")
            a6 = copy.copy(args); a6.workload = "vision"; a6.n_flies = DEFAULT_FLIES["vision"]; a6.chunk = DEFAULT_CHUNK["vision"]
            extras["config4_vision"] = sub_record(a6, measure(ctx, a6, steps=100, warmup=3, sample_clocks=False, e2e_cap=50, dominant=False),
                                                  "BASELINE config 4: 1024 flies, two eye-camera renders (ground, sky, the fly's own body) -> Retina after every step")
            a8 = copy.copy(a6); a8.vision_every = 20; a8.chunk = 20
            extras["config4_vision_every20"] = sub_record(a8, measure(ctx, a8, steps=400, warmup=3, sample_clocks=False, e2e_cap=100, dominant=False),
                                                          "config 4 with the v1 vision refresh (SURVEY.md 8d): 20 physics steps fused per launch, one vision frame "
                                                          "(render + Retina) after every 20th; end to end the frame is read back every 20th step")
        # BASELINE config 5 at every N: weak (32768 flies per GPU) and strong (262144 flies in total) scaling records
        a5 = copy.copy(args); a5.workload = "olfaction"; a5.n_flies = DEFAULT_FLIES["olfaction"]; a5.chunk = DEFAULT_CHUNK["olfaction"]
        extras["config5"] = sub_record(a5, measure(ctx, a5, steps=200, warmup=3, sample_clocks=False, e2e_cap=50, dominant=False),
                                       "BASELINE config 5 (weak scaling): 32768 flies per GPU, odor sensors after every step, all_gather of a (n, 8) metrics slab "
                                       "every 100 steps on a side stream (the stepping stream joins it at the end; that wait is inside ms_per_step)")
        a9 = copy.copy(a5); a9.n_flies = CONFIG5_TOTAL_FLIES // world
        extras["config5_strong"] = sub_record(a9, measure(ctx, a9, steps=100, warmup=3, sample_clocks=False, e2e_cap=20, dominant=False, table_rows=250),   # (a 2500-row table of 262144 flies would be 110 GB)
                                              f"BASELINE config 5 with the TOTAL fixed (strong scaling): {CONFIG5_TOTAL_FLIES} flies over {world} GPU(s) = "
                                              f"{CONFIG5_TOTAL_FLIES // world} per GPU, same sensors and all_gather")
        extras["config5_strong"]["scaling"] = "strong"

    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        achieved = roof["alg_bytes"] / (roof["ms"] * 1e-3) / 1e9
        cpu = cpu_baseline_leg(model, n, stance_adhesion=(wl == "terrain")) if (world == 1 and not args.no_cpu) else None
        cfg = dict(workload_config(args, n, chunk), actions=args.actions)
        if wl == "vision":
            tr = (VISION_TRAFFIC[0] * n, VISION_TRAFFIC[1])
        else:
            per, src = NCU_TRAFFIC_PER_FLY_STEP[wl]
            tr = (per * roof["fly_steps"], f"{per:.0f} B per fly-step x {roof['fly_steps']} fly-steps of this launch shape; scaled from {src}")
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": f"f{args.precision}", "data": "synthetic", "config": cfg,
            "clocks": m["clocks"],
            "e2e": m["e2e"],
            "gpu_launches": m["launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": tr[0], "traffic_source": tr[1], "peak_kind": peak_kind, "kernel": roof["kernel"], "kernel_ms_per_launch": roof["ms"],
                         "note": roof["note"]},
            "cpu_baseline": cpu,
            "wall_s_timed_region": m["wall"], "state_finite": m["finite"],
        }
        if args.skeleton != "legs_only":
            line["roofline"]["traffic"] = TREE_TRAFFIC_PER_FLY_STEP[0] * roof["fly_steps"]
            line["roofline"]["traffic_source"] = TREE_TRAFFIC_PER_FLY_STEP[1]
            line["roofline"]["note"] = "general-topology kernel: the fly lives in shared memory, bound by instruction issue (DESIGN.md 4.5); algorithmic bytes = 4 (2 nq + 4 nv + nu) per fly-step"
        if wl in NCU_LIMITER and args.actions == "cpg" and not args.mesh and args.precision == 32 and args.skeleton == "legs_only":
            line["roofline"]["limiter"] = dict(NCU_LIMITER[wl])
            if "fp32_flop_per_fly_step" in NCU_LIMITER[wl]:       # achieved FP32 rate of the fused step, per GPU
                lim = line["roofline"]["limiter"]
                lim["fp32_tflops_achieved"] = m["value"] / world * lim["fp32_flop_per_fly_step"] / 1e12
                lim["fp32_frac_of_cuda_core_peak"] = lim["fp32_tflops_achieved"] / lim["fp32_peak_tflops"]
        if "retina_over_buffers" in roof:
            rb = dict(roof["retina_over_buffers"]); rb["frac"] = rb["achieved"] / peak
            if n == 1024:
                rb["traffic"], rb["traffic_source"] = RETINA_BUFFERS_TRAFFIC
            line["roofline"]["retina_over_buffers"] = rb
        if m["gathered_slabs"]:
            line["nccl_all_gathers_in_timed_region"] = m["gathered_slabs"]
        if extras:
            line["extras"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


DEFAULT_FLIES = {"flat": 4096, "terrain": 4096, "vision": 1024, "olfaction": 32768}
DEFAULT_CHUNK = {"flat": 100, "terrain": 100, "vision": 10, "olfaction": 10}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="flat", choices=list(DEFAULT_FLIES), help="BASELINE.json configs[1..4]")
    ap.add_argument("--terrain", default="blocks", choices=["blocks", "gapped"], help="terrain of --workload terrain")
    ap.add_argument("--n-flies", type=int, default=None, help="flies per GPU (default: the BASELINE config's)")
    ap.add_argument("--chunk", type=int, default=None, help="physics steps per timed launch group (fused into one launch when no sensors run)")
    ap.add_argument("--mesh", action="store_true", help="mesh-hull collision geoms (simplify_geom=False)")
    ap.add_argument("--skeleton", default="legs_only", choices=list(SKELETON_DIMS), help="JointPreset of the fly: legs_only = the reference benchmark "
                    "model (star kernels); all_biological / all_possible = the full skeletons (general-topology kernels)")
    ap.add_argument("--actions", default="cpg", choices=["cpg", "replay"], help="cpg = BASELINE config 2; replay = the reference benchmark's kinematic-replay clip")
    ap.add_argument("--vision-every", type=int, default=1, help="vision workload: physics steps per vision frame (1 = BASELINE's 'per step'; 20 = v1's 500 Hz refresh)")
    ap.add_argument("--eye-body", default="on", choices=["on", "off"], help="vision workload: the eye cameras also see the fly's own body (capsule proxies)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the secondary records (mesh / f64 at N = 1, config 5 at N > 1)")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64], help="arithmetic of the step kernel: 32 = product path; 64 = the same "
                    "kernel source in double precision (validation build that shadows the fp64 oracle)")
    args = ap.parse_args()
    if args.n_flies is None:
        args.n_flies = DEFAULT_FLIES[args.workload]
    if args.chunk is None:
        args.chunk = DEFAULT_CHUNK[args.workload]
    if args.steps is None:
        args.steps = 1000 if args.impl == "ours" and args.workload in ("flat", "terrain") else (200 if args.impl == "ours" else 5)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
