"""Retina transform and odor-intensity sensor on top of ``libnmf_b200`` (sm_100a kernels).

FlyGym 2.0.1 ships neither component: only the v1 parameter
block survives at ``/root/reference/src/flygym/assets/model/legacy/flygym1_config.yaml:141-200``
(512 x 450 px per eye, 721 ommatidia per eye, fisheye coefficient 3.8 / zoom 2.72; four odor
sensors: L/R maxillary palp on the rostrum, L/R antenna on the funiculi) and the id-map /
pale-mask assets it points to are absent.  PARITY UNPINNED: the operator is therefore defined by
the deterministic generator below, mirroring the v1 semantics as far as they are known [PRIOR]:

* every eye-camera pixel belongs to at most one of 721 hexagonal ommatidia (radius-15 hexagon on
  a hex lattice laid over the fisheye-corrected image plane);
* each ommatidium is "yellow" (reads the green channel) or "pale" (reads blue, ~30 %);
* readout ``(2 eyes, 721, 2)``: mean of that channel over the ommatidium's pixels / 255, written to
  slot 0 (yellow) or 1 (pale), the other slot is 0.

The id map is part of the result ("ommatidia index map bit-exact" in BASELINE.json): the CUDA
kernel consumes exactly the table produced here and is checked bit-for-bit against the numpy
restatement in ``oracle/retina_oracle.py``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

RAW_H, RAW_W = 512, 450          # flygym1_config.yaml:143-144
N_RINGS = 15                     # 1 + 3*15*16 = 721 ommatidia (flygym1_config.yaml:145)
FISHEYE_K, FISHEYE_ZOOM = 3.8, 2.72   # flygym1_config.yaml:146-147
PALE_FRACTION = 0.3

# odor sensors (flygym1_config.yaml:175-192), v1 -> v2 segment names via utils/api1to2.py:6-45
ODOR_SENSORS = [("c_rostrum", (-0.15, 0.15, -0.15)), ("c_rostrum", (-0.15, -0.15, -0.15)),
                ("l_funiculus", (0.02, 0.0, -0.10)), ("r_funiculus", (0.02, 0.0, -0.10))]


def hex_cells(n_rings: int = N_RINGS):
    """Axial coordinates (q, r) of the radius-n hexagon, canonical order (r, then q)."""
    cells = [(q, r) for r in range(-n_rings, n_rings + 1) for q in range(-n_rings, n_rings + 1)
             if max(abs(q), abs(r), abs(q + r)) <= n_rings]
    return np.array(cells, dtype=np.int64)


def pale_mask(n_omm: int) -> np.ndarray:
    """Deterministic ~30 % pale-type mask (Knuth multiplicative hash of the ommatidium index)."""
    o = np.arange(n_omm, dtype=np.uint64)
    h = (o * np.uint64(2654435761)) % np.uint64(2**32)
    return (h.astype(np.float64) / 2**32) < PALE_FRACTION


def ommatidia_id_map(H: int = RAW_H, W: int = RAW_W, n_rings: int = N_RINGS, k: float = FISHEYE_K, zoom: float = FISHEYE_ZOOM):
    """``(2, H, W)`` int16 map: 0 = no ommatidium, 1..721 = ommatidium id; eye 1 is the mirror image of eye 0."""
    cells = hex_cells(n_rings)
    lut = -np.ones((2 * n_rings + 1, 2 * n_rings + 1), dtype=np.int64)
    lut[cells[:, 1] + n_rings, cells[:, 0] + n_rings] = np.arange(len(cells))
    rows, cols = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    half = max(H, W) / 2.0
    u, v = (cols - (W - 1) / 2.0) / half, (rows - (H - 1) / 2.0) / half
    f = (1.0 + k * (u * u + v * v)) / zoom          # fisheye correction of the image plane
    x, y = u * f, v * f
    u_edge = ((W - 1) / 2.0) / half
    r_in = u_edge * (1.0 + k * u_edge * u_edge) / zoom   # corrected radius at the edge of the shorter image axis
    s = r_in / ((n_rings + 0.5) * np.sqrt(3.0))         # cell size: the hexagon's inradius spans that radius
    qf = (np.sqrt(3.0) / 3.0 * x - 1.0 / 3.0 * y) / s   # pointy-top axial coordinates + cube rounding
    rf = (2.0 / 3.0 * y) / s
    xf, zf = qf, rf
    yf = -xf - zf
    rx, ry, rz = np.round(xf), np.round(yf), np.round(zf)
    dx, dy, dz = np.abs(rx - xf), np.abs(ry - yf), np.abs(rz - zf)
    fix_x = (dx > dy) & (dx > dz)
    fix_z = ~fix_x & ~(dy > dz)
    rx = np.where(fix_x, -ry - rz, rx)
    rz = np.where(fix_z, -rx - ry, rz)
    q, r = rx.astype(np.int64), rz.astype(np.int64)
    inside = np.maximum(np.maximum(np.abs(q), np.abs(r)), np.abs(q + r)) <= n_rings
    ids = np.zeros((H, W), dtype=np.int64)
    ids[inside] = lut[r[inside] + n_rings, q[inside] + n_rings] + 1
    return np.stack([ids, ids[:, ::-1]]).astype(np.int16)


def retina_tables(id_map: np.ndarray, pale: np.ndarray):
    """Kernel tables: ``pixcode (2, H*W) int16`` and ``inv_norm (2, n_omm+1, 2) float32``."""
    n_omm = len(pale)
    flat = id_map.reshape(2, -1).astype(np.int64)
    is_pale = np.r_[False, pale][flat]
    pixcode = (2 * flat + is_pale).astype(np.int16)
    pixcode[flat == 0] = 0
    inv = np.zeros((2, n_omm + 1, 2), dtype=np.float32)
    for e in range(2):
        cnt = np.bincount(flat[e], minlength=n_omm + 1).astype(np.float32)
        with np.errstate(divide="ignore"):
            w = np.float32(1.0) / (np.float32(255.0) * cnt)
        w[~np.isfinite(w)] = 0.0
        inv[e, 1:, 0] = np.where(pale, 0.0, w[1:])
        inv[e, 1:, 1] = np.where(pale, w[1:], 0.0)
    return pixcode, inv


class Retina:
    """Batched Retina transform: ``images (n, 2, H, W, 3) uint8`` -> ``(n, 2, 721, 2) float32``."""

    def __init__(self, device=None, H: int = RAW_H, W: int = RAW_W, n_rings: int = N_RINGS):
        if not torch.cuda.is_available():
            raise RuntimeError("Retina needs a CUDA device (there is no CPU fallback).")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:       # plain "cuda": the current device
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.H, self.W = H, W
        self.id_map = ommatidia_id_map(H, W, n_rings)
        self.n_ommatidia = int(self.id_map.max())
        self.pale = pale_mask(self.n_ommatidia)
        self.pixcode, self.inv_norm = retina_tables(self.id_map, self.pale)
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.nmf_retina_create(self.pixcode.ctypes.data_as(ctypes.c_void_p), self.inv_norm.ctypes.data_as(ctypes.c_void_p),
                                         H, W, self.n_ommatidia, int(self.device.index), ctypes.byref(h))
        self._h = h
        if rc != 0:
            raise RuntimeError("nmf_retina_create failed: " + (self._lib.nmf_retina_last_error(h).decode() if h else "alloc"))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.nmf_retina_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"libnmf_b200 retina: {self._lib.nmf_retina_last_error(self._h).decode()} (status {rc})")

    def __call__(self, images: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if images.dtype != torch.uint8 or not images.is_cuda or not images.is_contiguous():
            raise ValueError("images must be a contiguous uint8 CUDA tensor")
        if images.ndim != 5 or tuple(images.shape[1:]) != (2, self.H, self.W, 3):
            raise ValueError(f"images must have shape (n, 2, {self.H}, {self.W}, 3)")
        n = images.shape[0]
        if out is None:
            out = torch.empty((n, 2, self.n_ommatidia, 2), dtype=torch.float32, device=images.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)
        self._check(self._lib.nmf_retina_forward(self._h, ctypes.c_void_p(images.data_ptr()), n, ctypes.c_void_p(out.data_ptr()), stream))
        return out

    def forward_host(self, images: np.ndarray, out: np.ndarray) -> None:
        """HOST buffers in and out (H2D, kernel, D2H); synchronous."""
        n = images.shape[0]
        assert images.dtype == np.uint8 and out.dtype == np.float32 and out.shape == (n, 2, self.n_ommatidia, 2)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._check(self._lib.nmf_retina_forward_host(self._h, images.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p), stream))

    @property
    def launch_count(self) -> int:
        return int(self._lib.nmf_retina_launch_count(self._h))


class OdorSensor:
    """Odor intensity at the four olfactory sensor sites, ``(n_worlds, odor_dim, 4)``:
    ``I[d, s] = sum_src peak[src, d] / |x_sensor[s] - x_src|^2``  (v1 semantics, [PRIOR])."""

    def __init__(self, sim, source_positions, peak_intensities):
        self.sim = sim
        dev = sim.device
        segs = sim.model.names["segments"]
        self.sensor_seg = torch.tensor([segs.index(s) for s, _ in ODOR_SENSORS], dtype=torch.int32, device=dev)
        self.sensor_rel = torch.tensor([p for _, p in ODOR_SENSORS], dtype=torch.float32, device=dev).contiguous()
        self.src_pos = torch.as_tensor(np.asarray(source_positions, dtype=np.float32)).to(dev).contiguous()
        self.src_peak = torch.as_tensor(np.asarray(peak_intensities, dtype=np.float32)).to(dev).contiguous()
        if self.src_pos.ndim != 2 or self.src_pos.shape[1] != 3 or self.src_peak.shape[0] != self.src_pos.shape[0]:
            raise ValueError("source_positions must be (n_src, 3) and peak_intensities (n_src, odor_dim)")
        self.odor_dim = int(self.src_peak.shape[1])
        self._lib = _lib.load()

    def __call__(self) -> torch.Tensor:
        sim = self.sim
        if sim.seg_xpos is None:
            raise RuntimeError("the simulation was created with outputs=False")
        out = torch.empty((sim.n_worlds, self.odor_dim, 4), dtype=torch.float32, device=sim.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        rc = self._lib.nmf_odor_intensity(p(sim.seg_xpos), p(sim.seg_xquat), sim.n_worlds, sim.info.nseg, p(self.sensor_seg),
                                          p(self.sensor_rel), p(self.src_pos), p(self.src_peak), int(self.src_pos.shape[0]),
                                          self.odor_dim, p(out), ctypes.c_void_p(torch.cuda.current_stream(sim.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"nmf_odor_intensity failed (status {rc})")
        return out


# ------------------------------------------------------------------------------------------------ eye cameras
FOVY_DEG = 157.0                                   # flygym1_config.yaml:142
EYE_CAMERAS = [("l_eye", (-0.03, 0.38, 0.0), (1.57, 0.00, -0.47)),      # flygym1_config.yaml:163-174 (parent, rel_pos, euler)
               ("r_eye", (-0.03, -0.38, 0.0), (-1.57, 3.14, 0.47))]
CHECKER_MM = 4.0                                   # world.py:233-250: 2000 mm plane, texrepeat 250, 2x2 builtin checker
GROUND_U8, SKY_GB_U8 = (77, 102), (178, 229)       # rgb1 = 0.3, rgb2 = 0.4 (world.py:240-241); uniform light-blue sky


BODY_GB_U8 = (64, 38)                              # the fly's own body in its eyes' view: one dark brown tone (green, blue)
# segments the eye cameras do not see (flygym1_config.yaml:148-162, v1 names -> v2 via utils/api1to2.py:6-45)
HIDDEN_SEGMENTS = ("lf_coxa", "l_eye", "l_arista", "l_funiculus", "l_pedicel", "rf_coxa", "r_eye", "r_arista", "r_funiculus", "r_pedicel",
                   "c_head", "c_rostrum", "c_haustellum", "c_thorax")


def body_capsules(model) -> dict | None:
    """Capsule proxies (float32) of the segments the eye cameras see, from the baked model's ``viscap_*`` arrays (the capsule the
    baker fits to every segment's mesh): segment index, the two end points in the segment frame, radius.  ``None`` for models
    without them (e.g. converted from an ``MjModel``)."""
    a = model.arrays
    if "viscap_pos" not in a:
        return None
    segs = model.names["segments"]
    keep = [i for i, s in enumerate(segs) if s not in HIDDEN_SEGMENTS]
    pos, axis, size = (np.asarray(a[k], dtype=np.float64) for k in ("viscap_pos", "viscap_axis", "viscap_size"))
    seg = np.array(keep, dtype=np.int32)
    A = (pos[keep] + axis[keep] * size[keep, 1:2]).astype(np.float32)
    B = (pos[keep] - axis[keep] * size[keep, 1:2]).astype(np.float32)
    return dict(seg=seg, a=np.ascontiguousarray(A), b=np.ascontiguousarray(B), rad=size[keep, 0].astype(np.float32), colour=BODY_GB_U8)


def euler_xyz_to_mat(e) -> np.ndarray:
    """Intrinsic x-y-z Euler angles -> rotation matrix (own convention, documented in DESIGN.md)."""
    a, b, c = e
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def eye_params(segments: list[str], H: int = RAW_H, W: int = RAW_W) -> dict:
    """Plain-number camera description shared by the CUDA kernels and the numpy oracle (all float32)."""
    f = (H / 2.0) / np.tan(np.deg2rad(FOVY_DEG) / 2.0)
    return dict(
        eye_seg=np.array([segments.index(s) for s, _, _ in EYE_CAMERAS], dtype=np.int32),
        rel_pos=np.array([p for _, p, _ in EYE_CAMERAS], dtype=np.float32),
        R_local=np.array([euler_xyz_to_mat(e) for _, _, e in EYE_CAMERAS]).astype(np.float32),
        cx=np.float32((W - 1) / 2.0), cy=np.float32((H - 1) / 2.0), inv_f=np.float32(1.0 / f),
        inv_check=np.float32(1.0 / CHECKER_MM), ground=GROUND_U8, sky=SKY_GB_U8)


class EyeCameras:
    """Image formation for the two compound-eye cameras on top of a :class:`B200Simulation` (segment poses of the last
    step): ``render()`` gives the raw buffers ``(n, 2, 512, 450, 3) uint8``; ``retina()`` is the fused
    render + Retina path that never materialises them.  The two are bit-identical by construction
    (``retina() == Retina()(render())``).  The cameras see the checker ground, the sky and -- with ``body`` -- the fly's own
    body: every segment outside the v1 hidden list as the capsule the baker fits to its mesh."""

    def __init__(self, sim, retina: Retina | None = None, body: bool = True):
        self.sim = sim
        self.ret = retina if retina is not None else Retina(device=sim.device)
        self.params = eye_params(sim.model.names["segments"], self.ret.H, self.ret.W)
        p = self.params
        c = _lib.NmfEyeParams()
        c.eye_seg[:] = [int(v) for v in p["eye_seg"]]
        c.rel_pos[:] = [float(v) for v in p["rel_pos"].ravel()]
        c.R_local[:] = [float(v) for v in p["R_local"].ravel()]
        c.cx, c.cy, c.inv_f, c.inv_check = float(p["cx"]), float(p["cy"]), float(p["inv_f"]), float(p["inv_check"])
        c.ground_lo, c.ground_hi = p["ground"]
        c.sky_g, c.sky_b = p["sky"]
        c.body_g, c.body_b = BODY_GB_U8
        self._c = c
        self._lib = _lib.load()
        self.body = body_capsules(sim.model) if body else None
        if self.body is not None:
            b = self.body
            ptr = lambda x: x.ctypes.data_as(ctypes.c_void_p)
            self.ret._check(self._lib.nmf_eye_set_body(self.ret._h, ptr(b["seg"]), ptr(b["a"]), ptr(b["b"]), ptr(b["rad"]), len(b["seg"])))
        else:
            self.ret._check(self._lib.nmf_eye_set_body(self.ret._h, None, None, None, None, 0))

    def _args(self):
        sim = self.sim
        if sim.seg_xpos is None:
            raise RuntimeError("the simulation was created with outputs=False")
        return (self.ret._h, ctypes.byref(self._c), ctypes.c_void_p(sim.seg_xpos.data_ptr()), ctypes.c_void_p(sim.seg_xquat.data_ptr()),
                sim.n_worlds, sim.info.nseg)

    def render(self, out: torch.Tensor | None = None) -> torch.Tensor:
        sim = self.sim
        if out is None:
            out = torch.empty((sim.n_worlds, 2, self.ret.H, self.ret.W, 3), dtype=torch.uint8, device=sim.device)
        rc = self._lib.nmf_eye_render(*self._args(), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(sim.device).cuda_stream))
        self.ret._check(rc)
        return out

    def retina(self, out: torch.Tensor | None = None) -> torch.Tensor:
        sim = self.sim
        if out is None:
            out = torch.empty((sim.n_worlds, 2, self.ret.n_ommatidia, 2), dtype=torch.float32, device=sim.device)
        rc = self._lib.nmf_eye_retina(*self._args(), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(sim.device).cuda_stream))
        self.ret._check(rc)
        return out
