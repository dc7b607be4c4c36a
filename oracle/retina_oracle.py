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
