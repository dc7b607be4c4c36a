"""Build + ctypes binding of ``libnmf_b200.so`` (the sm_100a product library).

There is deliberately no CPU fallback: if the CUDA library cannot be built or
loaded, importing the simulation fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
SO_PATH = CSRC / "libnmf_b200.so"
SOURCES = ["nmf_capi.cu", "nmf_retina.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str | None:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build() -> bool:
    if not SO_PATH.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                 + [CSRC.parent.parent / "include" / "nmf_b200.h"])
    return SO_PATH.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source of the package for sm_100a into one shared library."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("flygym_b200: nvcc not found and libnmf_b200.so is missing/out of date")
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    extra = os.environ.get("NMF_NVCC_EXTRA", "").split()
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", str(SO_PATH), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return SO_PATH


class NmfInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n_flies", "nq", "nv", "nu_pos", "nu_adh", "nseg", "nleg", "state_stride", "off_qpos", "off_qvel",
        "off_qacc_warmstart", "off_ctrl", "off_time", "dbg_stride")] + [("timestep", ctypes.c_float), ("off_status", ctypes.c_int32)]


class NmfBuffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("state", "seg_xpos", "seg_xquat", "act_force", "sensordata", "debug", "energy")]


class NmfEyeParams(ctypes.Structure):
    _fields_ = [("eye_seg", ctypes.c_int32 * 2), ("rel_pos", ctypes.c_float * 6), ("R_local", ctypes.c_float * 18),
                ("cx", ctypes.c_float), ("cy", ctypes.c_float), ("inv_f", ctypes.c_float), ("inv_check", ctypes.c_float),
                ("ground_lo", ctypes.c_uint32), ("ground_hi", ctypes.c_uint32), ("sky_g", ctypes.c_uint32), ("sky_b", ctypes.c_uint32),
                ("body_g", ctypes.c_uint32), ("body_b", ctypes.c_uint32)]


_LIB = None


def load() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    override = os.environ.get("NMF_LIB_PATH")      # kernel-tuning experiments: load an alternative build of the same sources
    if override is None and needs_build():
        build()
    lib = ctypes.CDLL(override or str(SO_PATH))
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.nmf_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ci, ci, ctypes.POINTER(vp)]
    lib.nmf_destroy.argtypes = [vp]
    lib.nmf_model_info.argtypes = [vp, ctypes.POINTER(NmfInfo)]
    lib.nmf_last_error.argtypes = [vp]
    lib.nmf_last_error.restype = ctypes.c_char_p
    lib.nmf_bind.argtypes = [vp, ctypes.POINTER(NmfBuffers)]
    lib.nmf_reset.argtypes = [vp, vp, vp]
    lib.nmf_step.argtypes = [vp, ci, vp, ci, ci, ci, vp]
    lib.nmf_set_schedule.argtypes = [vp, ci]
    lib.nmf_set_precision.argtypes = [vp, ci]
    lib.nmf_set_flies_per_block.argtypes = [vp, ci]
    lib.nmf_forward.argtypes = [vp, vp]
    lib.nmf_replay_table.argtypes = [vp, vp, ci, ci, ctypes.c_double, ctypes.c_double, ci, ci, ci, ci, vp, vp]
    lib.nmf_replay_table.restype = ci
    lib.nmf_scatter_ctrl.argtypes = [vp, vp, vp, ci, vp]
    lib.nmf_gather_state.argtypes = [vp, ci, vp, ci, vp, vp]
    lib.nmf_step_host.argtypes = [vp, vp, ci, ci, vp, vp]
    lib.nmf_set_solver.argtypes = [vp, ci, ci]
    lib.nmf_launch_count.argtypes = [vp]
    lib.nmf_launch_count.restype = ctypes.c_int64
    for fn in ("nmf_create", "nmf_destroy", "nmf_model_info", "nmf_bind", "nmf_reset", "nmf_step", "nmf_scatter_ctrl",
               "nmf_gather_state", "nmf_step_host", "nmf_set_solver", "nmf_set_schedule", "nmf_set_precision", "nmf_forward", "nmf_set_flies_per_block"):
        getattr(lib, fn).restype = ci
    lib.nmf_retina_create.argtypes = [vp, vp, ci, ci, ci, ci, ctypes.POINTER(vp)]
    lib.nmf_retina_destroy.argtypes = [vp]
    lib.nmf_retina_last_error.argtypes = [vp]
    lib.nmf_retina_last_error.restype = ctypes.c_char_p
    lib.nmf_retina_launch_count.argtypes = [vp]
    lib.nmf_retina_launch_count.restype = ctypes.c_int64
    lib.nmf_retina_forward.argtypes = [vp, vp, ci, vp, vp]
    lib.nmf_retina_forward_host.argtypes = [vp, vp, ci, vp, vp]
    lib.nmf_odor_intensity.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, ci, ci, vp, vp]
    lib.nmf_eye_render.argtypes = [vp, ctypes.POINTER(NmfEyeParams), vp, vp, ci, ci, vp, vp]
    lib.nmf_eye_retina.argtypes = [vp, ctypes.POINTER(NmfEyeParams), vp, vp, ci, ci, vp, vp]
    lib.nmf_eye_set_body.argtypes = [vp, vp, vp, vp, vp, ci]
    lib.nmf_eye_set_body.restype = ci
    lib.nmf_eye_render.restype = ci
    lib.nmf_eye_retina.restype = ci
    for fn in ("nmf_retina_create", "nmf_retina_destroy", "nmf_retina_forward", "nmf_retina_forward_host", "nmf_odor_intensity"):
        getattr(lib, fn).restype = ci
    _LIB = lib
    return lib


# symbols declared in include/nmf_b200.h (checked by the CPU test-suite)
DECLARED_SYMBOLS = [
    "nmf_create", "nmf_destroy", "nmf_model_info", "nmf_last_error", "nmf_bind", "nmf_reset", "nmf_step",
    "nmf_scatter_ctrl", "nmf_gather_state", "nmf_step_host", "nmf_set_solver", "nmf_set_schedule", "nmf_set_precision", "nmf_set_flies_per_block", "nmf_forward", "nmf_replay_table", "nmf_launch_count",
    "nmf_retina_create", "nmf_retina_destroy", "nmf_retina_last_error", "nmf_retina_launch_count", "nmf_retina_forward",
    "nmf_retina_forward_host", "nmf_odor_intensity", "nmf_eye_render", "nmf_eye_retina", "nmf_eye_set_body",
]
