"""ctypes front-end of the CPU fp64 oracle (``oracle/nmf_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py; never by ``flygym_b200``.
PARITY UNPINNED (see the C file's header): MuJoCo 3.6.0, the reference's real
arithmetic, is unavailable; this restates its documented pipeline.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libnmf_oracle.so"
    src = _HERE / "nmf_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "-B", "libnmf_oracle.so"])
    return so


def build_native() -> Path | None:
    """The same source built ``-O3 -march=native`` ON THE MACHINE THAT RUNS IT (bench.py's CPU arm only: a build made for
    this container's CPU must not travel to another host).  Returns None when no compiler is available."""
    so = _HERE / "_native" / "libnmf_oracle_native.so"
    src = _HERE / "nmf_oracle.c"
    try:
        so.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=c11", "-shared", "-o", str(so), str(src), "-lm"],
                              stderr=subprocess.DEVNULL)
        return so
    except Exception:
        return None


_LIB_NATIVE = None


def _lib(native: bool = False):
    global _LIB, _LIB_NATIVE
    if native:
        if _LIB_NATIVE is None:
            so = build_native()
            _LIB_NATIVE = _bind(ctypes.CDLL(str(so))) if so is not None else _lib()
        return _LIB_NATIVE
    if _LIB is None:
        _LIB = _bind(ctypes.CDLL(str(build())))
    return _LIB


def _bind(lib):
    lib.nmfo_create.restype = ctypes.c_void_p
    lib.nmfo_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    lib.nmfo_destroy.argtypes = [ctypes.c_void_p]
    lib.nmfo_reset.argtypes = [ctypes.c_void_p]
    lib.nmfo_forward.argtypes = [ctypes.c_void_p]
    lib.nmfo_step.argtypes = [ctypes.c_void_p]
    lib.nmfo_step_n.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.nmfo_step_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.nmfo_step_table_cols.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.nmfo_dim.restype = ctypes.c_int
    lib.nmfo_dim.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.nmfo_array.restype = ctypes.POINTER(ctypes.c_double)
    lib.nmfo_array.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
    lib.nmfo_con_geom.restype = ctypes.c_int
    lib.nmfo_con_geom.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.nmfo_last_error.restype = ctypes.c_char_p
    lib.nmfo_last_error.argtypes = [ctypes.c_void_p]
    return lib


class Oracle:
    """One fly, fp64.  ``get(name)`` returns a *view* (numpy) into oracle memory."""

    def __init__(self, model, native: bool = False):
        self._lib = _lib(native)
        blob = model.to_blob()
        self._h = self._lib.nmfo_create(blob, len(blob))
        if not self._h:
            raise RuntimeError("oracle: bad model blob")
        self.model = model
        self.reset()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.nmfo_destroy(self._h)
            self._h = None

    def dim(self, name: str) -> int:
        return self._lib.nmfo_dim(self._h, name.encode())

    def get(self, name: str) -> np.ndarray:
        n = ctypes.c_int(0)
        p = self._lib.nmfo_array(self._h, name.encode(), ctypes.byref(n))
        if not p:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    @property
    def qpos(self): return self.get("qpos")
    @property
    def qvel(self): return self.get("qvel")
    @property
    def ctrl(self): return self.get("ctrl")
    @property
    def time(self): return float(self.get("time")[0])

    def con_geom(self) -> np.ndarray:
        buf = (ctypes.c_int * 512)()
        n = self._lib.nmfo_con_geom(self._h, buf, 512)
        return np.array(buf[:n], dtype=np.int32)

    def reset(self): self._lib.nmfo_reset(self._h)
    def forward(self): self._lib.nmfo_forward(self._h)

    def step(self, n: int = 1):
        self._lib.nmfo_step_n(self._h, int(n))

    def step_table(self, table):
        """One step per row of ``table`` (float64 ``[n][nu_pos]``) used as position-actuator inputs."""
        table = np.ascontiguousarray(table, dtype=np.float64)
        self._lib.nmfo_step_table(self._h, table.ctypes.data_as(ctypes.c_void_p), int(table.shape[0]))

    def step_table_full(self, table):
        """One step per row of ``table`` (float64 ``[n][cols]``) copied into ``ctrl[0:cols]`` (position targets, then adhesion)."""
        table = np.ascontiguousarray(table, dtype=np.float64)
        self._lib.nmfo_step_table_cols(self._h, table.ctypes.data_as(ctypes.c_void_p), int(table.shape[0]), int(table.shape[1]))

    def error(self) -> str:
        return self._lib.nmfo_last_error(self._h).decode()
