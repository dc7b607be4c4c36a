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


def _seg_matrix(q, f32=np.float32):
    w, x, y, z = (f32(v) for v in q)
    two, one = f32(2), f32(1)
    return np.array([[one - two * (y * y + z * z), two * (x * y - w * z), two * (x * z + w * y)],
                     [two * (x * y + w * z), one - two * (x * x + z * z), two * (y * z - w * x)],
                     [two * (x * z - w * y), two * (y * z + w * x), one - two * (x * x + y * y)]], dtype=f32)


def body_mask_oracle(seg_xpos_i, seg_xquat_i, body, pos, wx, wy, wz):
    """float32 restatement of csrc/nmf_retina.cu::eye_body_hit for every pixel and every visible capsule (no culling here: the
    kernel's culling is conservative): True where the ray from the camera at `pos` with direction (wx, wy, wz) passes within the
    radius of a capsule's axis segment."""
    f32 = np.float32
    dot = lambda a0, a1, a2, b0, b1, b2: (a0 * b0 + a1 * b1) + a2 * b2
    mask = np.zeros(wx.shape, dtype=bool)
    cc = dot(wx, wy, wz, wx, wy, wz)
    for k in range(len(body["seg"])):
        seg = int(body["seg"][k])
        xp = seg_xpos_i[seg].astype(f32); S = _seg_matrix(seg_xquat_i[seg])
        a_, b_ = body["a"][k].astype(f32), body["b"][k].astype(f32)
        A = np.array([xp[i] + dot(S[i, 0], S[i, 1], S[i, 2], a_[0], a_[1], a_[2]) for i in range(3)], dtype=f32)
        B = np.array([xp[i] + dot(S[i, 0], S[i, 1], S[i, 2], b_[0], b_[1], b_[2]) for i in range(3)], dtype=f32)
        W0 = (A - pos).astype(f32); U = (B - A).astype(f32)
        a = dot(U[0], U[1], U[2], U[0], U[1], U[2]); d = dot(U[0], U[1], U[2], W0[0], W0[1], W0[2])
        r2 = f32(body["rad"][k]) * f32(body["rad"][k])
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            b = dot(U[0], U[1], U[2], wx, wy, wz)
            e = dot(wx, wy, wz, W0[0], W0[1], W0[2])
            D = a * cc - b * b
            ok = D > f32(1e-12)
            s = np.where(ok, np.minimum(np.maximum((b * e - cc * d) * (f32(1) / np.where(ok, D, f32(1))), f32(0)), f32(1)), f32(0)).astype(f32)
            t = ((b * s + e) * (f32(1) / cc)).astype(f32)
            neg = t < 0
            s_alt = np.minimum(np.maximum(-d / a, f32(0)), f32(1)) if a > 0 else f32(0)
            s = np.where(neg, s_alt, s).astype(f32); t = np.where(neg, f32(0), t).astype(f32)
            px = (W0[0] + s * U[0]) - t * wx; py = (W0[1] + s * U[1]) - t * wy; pz = (W0[2] + s * U[2]) - t * wz
            n0, n1, n2 = wy * U[2] - wz * U[1], wz * U[0] - wx * U[2], wx * U[1] - wy * U[0]      # quick reject: the infinite cylinder about the axis
            h = dot(W0[0], W0[1], W0[2], n0, n1, n2)
            keep = ~((h * h) > (r2 * f32(1.01)) * dot(n0, n1, n2, n0, n1, n2))
            mask |= keep & (dot(px, py, pz, px, py, pz) <= r2)
    return mask


def eye_render_oracle(seg_xpos, seg_xquat, prm, H, W, body=None):
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
            if body is not None:
                m = body_mask_oracle(seg_xpos[i], seg_xquat[i], body, pos, wx.astype(f32), wy.astype(f32), wz.astype(f32))
                g = np.where(m, np.uint8(body["colour"][0]), g); b = np.where(m, np.uint8(body["colour"][1]), b)
            img[i, e, :, :, 0] = g
            img[i, e, :, :, 1] = g
            img[i, e, :, :, 2] = b
    return img
