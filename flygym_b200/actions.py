"""Synthetic action sources for the benchmark configurations.

The reference ships no CPG controller (its only action source is the kinematic
replay clip, ``src/flygym_demo/spotlight_data/preprocessing.py``); BASELINE.json's
config 2 asks for "sinusoidal CPG tripod-gait actions", defined here
deterministically: for actuated DoF i of leg l,

    target_i(t) = neutral_i + A_i * sin(2 pi f t + phi_l + delta_i + psi_k)

f = 12 Hz; tripod phases phi = 0 for {lf, rm, lh}, pi for {rf, lm, rh}; amplitudes
coxa-pitch 0.35, trochanterfemur-pitch 0.30, tibia-pitch 0.40, tarsus1-pitch 0.15 rad
(the last three with delta = pi/2), roll / yaw 0; per-fly phase psi_k = 2 pi k / n.
"""
from __future__ import annotations

import numpy as np

TRIPOD_PHASE = {"lf": 0.0, "rm": 0.0, "lh": 0.0, "rf": np.pi, "lm": np.pi, "rh": np.pi}
AMPLITUDE = {"coxa-pitch": (0.35, 0.0), "trochanterfemur-pitch": (0.30, np.pi / 2),
             "tibia-pitch": (0.40, np.pi / 2), "tarsus1-pitch": (0.15, np.pi / 2)}


def cpg_parameters(model):
    """Per-actuator (neutral, amplitude, phase) for the model's position actuators."""
    names = model.names["actuated_position"]
    neutral = model.arrays["key_ctrl"][: len(names)].astype(np.float64)
    amp = np.zeros(len(names))
    phase = np.zeros(len(names))
    for i, nm in enumerate(names):
        _, child, axis = nm.split("-")
        leg, link = child.split("_")
        a, d = AMPLITUDE.get(f"{link}-{axis}", (0.0, 0.0))
        amp[i] = a
        phase[i] = TRIPOD_PHASE[leg] + d
    return neutral, amp, phase


def cpg_table(model, n_flies: int, n_steps: int, *, freq_hz: float = 12.0, fly_offset: int = 0,
              n_flies_total: int | None = None, dtype=np.float32) -> np.ndarray:
    """Action table ``(n_flies, n_steps, n_position_actuators)``; fly k of the *global* batch gets
    phase offset 2 pi k / n_flies_total, so results do not depend on how flies are sharded over ranks."""
    neutral, amp, phase = cpg_parameters(model)
    total = n_flies if n_flies_total is None else n_flies_total
    t = np.arange(n_steps) * model.timestep
    psi = 2 * np.pi * (np.arange(n_flies) + fly_offset) / total
    arg = 2 * np.pi * freq_hz * t[None, :, None] + phase[None, None, :] + psi[:, None, None]
    return (neutral[None, None, :] + amp[None, None, :] * np.sin(arg)).astype(dtype)


def replay_angles(model, timestep: float | None = None) -> np.ndarray:
    """The reference's only shipped action source: the recorded walking clip, Savitzky-Golay filtered (done offline,
    ``assets/replay_clip_filtered.npz``; generator in ``tests/golden/make_golden.py``) and cubic-interpolated onto the
    simulation time grid exactly as ``MotionSnippet.get_joint_angles`` does
    (reference ``src/flygym_demo/spotlight_data/preprocessing.py:80-142``).  Returns ``(n_steps, n_position_actuators)``."""
    from scipy.interpolate import interp1d
    from .model import ASSETS_DIR
    with np.load(ASSETS_DIR / "replay_clip_filtered.npz") as z:
        filt, fps, names = z["angles"].astype(np.float64), int(z["fps"]), [str(s) for s in z["actuators"]]
    if names != list(model.names["actuated_position"]):
        raise ValueError("replay clip was baked for a different actuator order")
    dt = model.timestep if timestep is None else timestep
    n = filt.shape[0]
    src = np.arange(n) / fps
    out_t = np.arange(0, n / fps, dt)
    f = interp1d(src, filt, kind="cubic", axis=0, bounds_error=False, fill_value=(filt[0], filt[-1]))
    return f(out_t)


def replay_table(model, n_flies: int, n_steps: int, *, fly_offset: int = 0, dtype=np.float32) -> np.ndarray:
    """``ReplayTargetData.make_target_angles_all_worlds`` (reference ``time_gpu_simulation.py:73-86``): world k replays
    partition ``k % n_partitions`` of the clip."""
    ang = replay_angles(model)
    n_part = ang.shape[0] // n_steps
    if n_part < 1:
        raise ValueError("clip shorter than the requested number of steps")
    out = np.empty((n_flies, n_steps, ang.shape[1]), dtype=dtype)
    for k in range(n_flies):
        p = (k + fly_offset) % n_part
        out[k] = ang[p * n_steps:(p + 1) * n_steps]
    return out


def replay_spline(model, timestep: float | None = None):
    """The cubic spline ``MotionSnippet.get_joint_angles`` evaluates (``interp1d(kind="cubic")`` = not-a-knot B-spline through the
    filtered samples) as piecewise-cubic coefficients on the source grid: ``(coef[4, n_int, A], last[A], fps, n_out_steps)``."""
    from scipy.interpolate import CubicSpline
    from .model import ASSETS_DIR
    with np.load(ASSETS_DIR / "replay_clip_filtered.npz") as z:
        filt, fps, names = z["angles"].astype(np.float64), int(z["fps"]), [str(s) for s in z["actuators"]]
    if names != list(model.names["actuated_position"]):
        raise ValueError("replay clip was baked for a different actuator order")
    n = filt.shape[0]
    # the not-a-knot cubic interpolant is unique, so CubicSpline gives the same function as interp1d's B-spline, already as
    # per-interval polynomials c[4, n - 1, A] (highest power first) in the local variable t - x_i
    pp = CubicSpline(np.arange(n) / fps, filt, axis=0, bc_type="not-a-knot")
    dt = model.timestep if timestep is None else timestep
    n_out = len(np.arange(0, n / fps, dt))
    return np.ascontiguousarray(pp.c), np.ascontiguousarray(filt[-1]), float(fps), n_out


def replay_table_device(model, n_flies: int, n_steps: int, device, *, fly_offset: int = 0):
    """``replay_table`` evaluated on the GPU (``nmf_replay_table``): only the spline coefficients are uploaded; returns a float32
    CUDA tensor ``(n_flies, n_steps, n_position_actuators)``."""
    import ctypes
    import torch
    from . import _lib
    coef, last, fps, n_out = replay_spline(model)
    n_part = n_out // n_steps
    if n_part < 1:
        raise ValueError("clip shorter than the requested number of steps")
    A = coef.shape[2]
    out = torch.empty((n_flies, n_steps, A), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        rc = _lib.load().nmf_replay_table(coef.ctypes.data_as(ctypes.c_void_p), last.ctypes.data_as(ctypes.c_void_p), coef.shape[1], A, fps,
                                          float(model.timestep), n_part, n_steps, n_flies, int(fly_offset), ctypes.c_void_p(out.data_ptr()),
                                          ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"nmf_replay_table failed (status {rc})")
    return out
