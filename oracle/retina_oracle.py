"""numpy restatement of the Retina transform and the odor-intensity sensor.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU legs).  PARITY UNPINNED: FlyGym 2.0.1 has no Retina /
olfaction code (only parameters at src/flygym/assets/model/legacy/flygym1_config.yaml:141-200, assets absent);
this restates the v1 semantics [PRIOR] on the deterministic id map of flygym_b200/retina.py.
"""
import numpy as np


def retina_oracle(images: np.ndarray, id_map: np.ndarray, pale: np.ndarray) -> np.ndarray:
    """images (n, 2, H, W, 3) uint8; id_map (2, H, W) int16 (0 = none); pale (n_omm,) bool -> (n, 2, n_omm, 2) float32."""
    n = images.shape[0]
    n_omm = len(pale)
    out = np.zeros((n, 2, n_omm, 2), dtype=np.float32)
    for e in range(2):
        ids = id_map[e].reshape(-1).astype(np.int64)
        cnt = np.bincount(ids, minlength=n_omm + 1).astype(np.float32)
        with np.errstate(divide="ignore"):
            w = np.float32(1.0) / (np.float32(255.0) * cnt)
        for i in range(n):
            px = images[i, e].reshape(-1, 3).astype(np.int64)
            green = np.bincount(ids, weights=px[:, 1], minlength=n_omm + 1)
            blue = np.bincount(ids, weights=px[:, 2], minlength=n_omm + 1)
            s = np.where(pale, blue[1:], green[1:]).astype(np.float32)     # exact integers < 2**24
            val = s * w[1:]
            out[i, e, :, 0] = np.where(pale, 0.0, val)
            out[i, e, :, 1] = np.where(pale, val, 0.0)
    return out


def quat_rotate(q, v):
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R @ v


def odor_oracle(seg_xpos, seg_xquat, sensor_seg, sensor_rel, src_pos, src_peak):
    """seg_xpos (n, nseg, 3), seg_xquat (n, nseg, 4) -> (n, D, 4) float64."""
    n, D = seg_xpos.shape[0], src_peak.shape[1]
    out = np.zeros((n, D, 4))
    for i in range(n):
        for s in range(4):
            p = seg_xpos[i, sensor_seg[s]] + quat_rotate(seg_xquat[i, sensor_seg[s]], np.asarray(sensor_rel[s], dtype=np.float64))
            d2 = ((p[None, :] - src_pos) ** 2).sum(axis=1)
            out[i, :, s] = (src_peak / d2[:, None]).sum(axis=0)
    return out


def eye_render_oracle(seg_xpos, seg_xquat, prm, H, W):
    """float32 restatement of csrc/nmf_retina.cu::eye_pixel (same operation order, every op individually rounded):
    raw eye images (n, 2, H, W, 3) uint8 from float32 segment poses."""
    f32 = np.float32
    n = seg_xpos.shape[0]
    img = np.zeros((n, 2, H, W, 3), dtype=np.uint8)
    rows, cols = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    dx = (cols - prm["cx"]) * prm["inv_f"]
    dy = (prm["cy"] - rows) * prm["inv_f"]
    for i in range(n):
        for e in range(2):
            seg = int(prm["eye_seg"][e])
            xp = seg_xpos[i, seg].astype(f32)
            w, x, y, z = (f32(v) for v in seg_xquat[i, seg])
            two, one = f32(2), f32(1)
            S = np.array([[one - two * (y * y + z * z), two * (x * y - w * z), two * (x * z + w * y)],
                          [two * (x * y + w * z), one - two * (x * x + z * z), two * (y * z - w * x)],
                          [two * (x * z - w * y), two * (y * z + w * x), one - two * (x * x + y * y)]], dtype=f32)
            rel, Rl = prm["rel_pos"][e].astype(f32), prm["R_local"][e].astype(f32)
            pos = np.array([xp[k] + ((S[k, 0] * rel[0] + S[k, 1] * rel[1]) + S[k, 2] * rel[2]) for k in range(3)], dtype=f32)
            R = np.array([[(S[k, 0] * Rl[0, j] + S[k, 1] * Rl[1, j]) + S[k, 2] * Rl[2, j] for j in range(3)] for k in range(3)], dtype=f32)
            wz = (R[2, 0] * dx + R[2, 1] * dy) - R[2, 2]
            wx = (R[0, 0] * dx + R[0, 1] * dy) - R[0, 2]
            wy = (R[1, 0] * dx + R[1, 1] * dy) - R[1, 2]
            hit = (wz < 0) & (pos[2] > 0)
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                t = pos[2] * (f32(1) / (-wz))                     # correctly rounded reciprocal, then one multiplication
                hx, hy = pos[0] + t * wx, pos[1] + t * wy
                sat = lambda v: np.clip(np.nan_to_num(np.floor(v), nan=0.0), -2.0 ** 31, 2.0 ** 31 - 1).astype(np.int64)   # float -> int32, round down, saturating
                ix, iy = sat(hx * prm["inv_check"]), sat(hy * prm["inv_check"])
                ix, iy = ix.astype(np.int32), iy.astype(np.int32)
            v = np.where(((ix.astype(np.int64) + iy) & 1) == 1, prm["ground"][1], prm["ground"][0])
            g = np.where(hit, v, prm["sky"][0]).astype(np.uint8)
            b = np.where(hit, v, prm["sky"][1]).astype(np.uint8)
            img[i, e, :, :, 0] = g
            img[i, e, :, :, 1] = g
            img[i, e, :, :, 2] = b
    return img
