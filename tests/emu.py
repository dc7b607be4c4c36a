"""ctypes front-end of the SIMT-emulated kernel build (tests only)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent / "simt_emu"
ROOT = Path(__file__).resolve().parent.parent
S_STRIDE, S_QPOS, S_QVEL, S_WARM, S_CTRL, S_TIME = 304, 0, 76, 148, 220, 300
NV = 72
DBG_FS = 4
DBG_QACC, DBG_FC, DBG_QACCE = DBG_FS + NV, DBG_FS + 2 * NV, DBG_FS + 3 * NV
DBG_CON = DBG_FS + 4 * NV
DBG_XPOS = DBG_CON + 64 * 24
DBG_CDOF = DBG_XPOS + 64 * 3
DBG_HROWS = DBG_CDOF + NV * 6
DBG_STRIDE = DBG_HROWS + 6 * 177 + 21 + 3
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = HERE / "libnmf_emu.so"
        srcs = [HERE / "emu_driver.cpp", HERE / "simt_emu.h"] + list((ROOT / "flygym_b200/csrc").glob("nmf_*.*h"))
        if not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-DNMF_SIMT_EMU", "-fPIC", "-shared", "-o", str(so),
                                   str(HERE / "emu_driver.cpp")])
        _lib = ctypes.CDLL(str(so))
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def key_state(model):
    blob = model.to_blob()
    out = np.zeros(S_STRIDE, np.float32)
    assert lib().emu_key_state(blob, ctypes.c_size_t(len(blob)), _p(out)) == 0
    return out


def step(model, state, nsteps=1, dbg=False, outputs=False, act_table=None, t0=0, max_newton=0, max_ls=0, precision=32, fpb=1, sub_steps=0):
    """state: float32 [n, S_STRIDE], updated in place. Returns dict of optional outputs."""
    blob = model.to_blob()
    n = state.shape[0]
    res = {}
    d = np.zeros((n, DBG_STRIDE), np.float32) if dbg else None
    nseg, nu = model.dim("nseg"), model.nu
    ox = np.zeros((n, nseg, 3), np.float32) if outputs else None
    oq = np.zeros((n, nseg, 4), np.float32) if outputs else None
    oa = np.zeros((n, nu), np.float32) if outputs else None
    os_ = np.zeros((n, 96), np.float32) if outputs else None
    oe = np.zeros((n, 2), np.float32) if outputs else None
    T = 0 if act_table is None else act_table.shape[1]
    cols = 0 if act_table is None else act_table.shape[2]
    rc = lib().emu_step(blob, ctypes.c_size_t(len(blob)), _p(state), n, nsteps, _p(d), _p(ox), _p(oq), _p(oa), _p(os_),
                        _p(act_table), T, t0, cols, max_newton, max_ls, precision, fpb, sub_steps, _p(oe))
    assert rc == 0
    res.update(dbg=d, xpos=ox, xquat=oq, actf=oa, sensor=os_, energy=oe)
    return res


# ------------------------------------------------------------------ general-topology (tree) kernels
TREE_INFO_FIELDS = ["s_stride", "s_qpos", "s_qvel", "s_warm", "s_ctrl", "s_time", "nq", "nv", "nu", "nseg", "nleg", "smem_f32", "smem_f64", "nH"]


def tree_info(model):
    blob = model.to_blob()
    out = np.zeros(len(TREE_INFO_FIELDS), np.int32)
    assert lib().emu_tree_info(blob, ctypes.c_size_t(len(blob)), _p(out)) == 0
    return dict(zip(TREE_INFO_FIELDS, (int(v) for v in out)))


def tree_key_state(model):
    info = tree_info(model)
    blob = model.to_blob()
    out = np.zeros(info["s_stride"], np.float32)
    assert lib().emu_tree_key_state(blob, ctypes.c_size_t(len(blob)), _p(out)) == 0
    return out


def tree_step(model, state, nsteps=1, outputs=False, act_table=None, t0=0, max_newton=0, max_ls=0, precision=32, forward_only=False, dbg=False):
    """state: float32 [n, s_stride] of the tree layout, updated in place."""
    blob = model.to_blob()
    info = tree_info(model)
    n = state.shape[0]
    assert state.dtype == np.float32 and state.shape[1] == info["s_stride"] and state.flags.c_contiguous
    nseg, nu, nleg = info["nseg"], info["nu"], info["nleg"]
    ox = np.zeros((n, nseg, 3), np.float32) if outputs else None
    oq = np.zeros((n, nseg, 4), np.float32) if outputs else None
    oa = np.zeros((n, nu), np.float32) if outputs else None
    os_ = np.zeros((n, nleg * 16), np.float32) if outputs else None
    oe = np.zeros((n, 2), np.float32) if outputs else None
    d = np.zeros((n, 4), np.float32) if dbg else None
    T = 0 if act_table is None else act_table.shape[1]
    cols = 0 if act_table is None else act_table.shape[2]
    rc = lib().emu_tree_step(blob, ctypes.c_size_t(len(blob)), _p(state), n, nsteps, _p(d), _p(ox), _p(oq), _p(oa), _p(os_),
                             _p(act_table), T, t0, cols, max_newton, max_ls, precision, _p(oe), int(forward_only))
    assert rc == 0
    return dict(xpos=ox, xquat=oq, actf=oa, sensor=os_, energy=oe, dbg=d, info=info)
