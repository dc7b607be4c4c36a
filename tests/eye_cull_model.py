"""numpy model of the CULLING in csrc/nmf_retina.cu (eye_body_setup's 16-row band intervals + eye_body_strip / eye_strip_row): which
(capsule, pixel) pairs the body raster hands to the exact hit test.  Test infrastructure: the CPU suite checks that this candidate set
contains every pixel the un-culled numpy restatement (oracle/retina_oracle.py::body_mask_oracle) marks, on poses where the strip is an
interval, a pair of half-lines, or switched off; the GPU suite checks the kernel's images bit for bit."""
import numpy as np

f32 = np.float32


def camera(prm, seg_xpos, seg_xquat, e):
    from oracle.retina_oracle import _seg_matrix
    seg = int(prm["eye_seg"][e]); xp = seg_xpos[seg].astype(f32); S = _seg_matrix(seg_xquat[seg])
    rel, Rl = prm["rel_pos"][e].astype(f32), prm["R_local"][e].astype(f32)
    pos = np.array([xp[k] + ((S[k, 0] * rel[0] + S[k, 1] * rel[1]) + S[k, 2] * rel[2]) for k in range(3)], dtype=f32)
    R = np.array([[(S[k, 0] * Rl[0, j] + S[k, 1] * Rl[1, j]) + S[k, 2] * Rl[2, j] for j in range(3)] for k in range(3)], dtype=f32)
    return pos, R


def capsule_in_camera(body, k, seg_xpos, seg_xquat, pos):
    from oracle.retina_oracle import _seg_matrix
    dot = lambda a0, a1, a2, b0, b1, b2: (a0 * b0 + a1 * b1) + a2 * b2
    sg = int(body["seg"][k]); xp = seg_xpos[sg].astype(f32); S = _seg_matrix(seg_xquat[sg])
    a_, b_ = body["a"][k].astype(f32), body["b"][k].astype(f32)
    A = np.array([xp[i] + dot(S[i, 0], S[i, 1], S[i, 2], a_[0], a_[1], a_[2]) for i in range(3)], dtype=f32)
    B = np.array([xp[i] + dot(S[i, 0], S[i, 1], S[i, 2], b_[0], b_[1], b_[2]) for i in range(3)], dtype=f32)
    rad = f32(body["rad"][k])
    return (A - pos).astype(f32), (B - A).astype(f32), rad * rad


def _axis_bounds(u, z, rs, tmax):
    h = np.hypot(u, z)
    if h <= rs:
        return -tmax, tmax
    phi = np.arctan2(u, z); dl = np.arcsin(min(rs / h, 1.0)) + 1e-4; amax = np.arctan(tmax)
    a0, a1 = max(phi - dl, -amax), min(phi + dl, amax)
    if a0 > a1:
        return None
    return np.tan(a0) - 1e-3, np.tan(a1) + 1e-3


def band_intervals(prm, R, W0, U, r2, H, W, sub=16):
    """eye_body_setup: column interval per 16-row band from `sub` padded spheres along the axis."""
    Rd = R.astype(np.float64); fpx = 1.0 / float(prm["inv_f"]); cx, cy = float(prm["cx"]), float(prm["cy"])
    tx, ty = (0.5 * W + 2) * float(prm["inv_f"]), (0.5 * H + 2) * float(prm["inv_f"])
    nb = 512 // 16 + 2
    c0 = np.full(nb, 1 << 20); c1 = np.full(nb, -1)
    rs = 1.02 * (np.sqrt(float(r2)) + 0.5 * np.sqrt(float(U.astype(np.float64) @ U.astype(np.float64))) / sub) + 1e-4
    for i in range(sub):
        d = W0.astype(np.float64) + (i + 0.5) / sub * U.astype(np.float64)
        u, v, z = Rd[:, 0] @ d, Rd[:, 1] @ d, -(Rd[:, 2] @ d)
        bx, by = _axis_bounds(u, z, rs, tx), _axis_bounds(v, z, rs, ty)
        if bx is None or by is None:
            continue
        a0, a1 = max(0, int(np.floor(cx + bx[0] * fpx)) - 1), min(W - 1, int(np.ceil(cx + bx[1] * fpx)) + 1)
        r0, r1 = max(0, int(np.floor(cy - by[1] * fpx)) - 1), min(H - 1, int(np.ceil(cy - by[0] * fpx)) + 1)
        if r0 > r1 or a0 > a1:
            continue
        for bnd in range(r0 >> 4, (r1 >> 4) + 1):
            c0[bnd] = min(c0[bnd], a0); c1[bnd] = max(c1[bnd], a1)
    return c0, c1


def strip_coefficients(R, W0, U, r2):
    """eye_body_strip: None when the strip is switched off for this capsule."""
    Rd = R.astype(np.float64); W0d, Ud = W0.astype(np.float64), U.astype(np.float64)
    m = np.cross(Ud, W0d); uu = Ud @ Ud; rho2 = 1.03 * float(r2) + 1e-9
    R0, R1, R2 = Rd[:, 0], Rd[:, 1], Rd[:, 2]
    zb, db = 0.0, None
    for i in range(3):
        d = W0d + 0.5 * i * Ud; z = -(R2 @ d)
        if z > zb:
            zb, db = z, d
    if not zb > 1e-3:
        return None
    dxc, dyc = (R0 @ db) / zb, (R1 @ db) / zb
    if not (abs(dxc) < 16 and abs(dyc) < 16):
        return None
    wc = dxc * R0 + dyc * R1 - R2
    qf = lambda a, b: (a @ m) * (b @ m) - rho2 * (uu * (a @ b) - (a @ Ud) * (b @ Ud))
    co = np.array([qf(R0, R0), qf(R0, wc), qf(R0, R1), qf(wc, wc), 2 * qf(R1, wc), qf(R1, R1)])
    big = np.abs(co).max()
    if not (1e-30 < big < 1e30):
        return None
    return np.concatenate([co / big, [dxc, dyc]]).astype(f32)


def strip_row(prm, s, row, c0, c1):
    """eye_strip_row: list of (lo, hi) column intervals of `row` inside [c0, c1]."""
    if s is None:
        return [(c0, c1)]
    cx, cy, inv_f = f32(prm["cx"]), f32(prm["cy"]), f32(prm["inv_f"])
    y = (cy - f32(row)) * inv_f - s[7]
    A, B, C = s[0], s[1] + s[2] * y, s[3] + (s[4] + s[5] * y) * y
    bb, ac = B * B, A * C
    disc = bb - ac; tol = f32(1e-4) * (bb + abs(ac)) + f32(1e-12)
    f = f32(1) / inv_f
    clamp = lambda x: min(max(x, f32(-16)), f32(16))
    if A > 1e-6:
        if disc < -tol:
            return []
        sq = np.sqrt(max(disc, f32(0)) + tol)
        x1, x2 = clamp((-B - sq) / A + s[6]), clamp((-B + sq) / A + s[6])
        return [(max(c0, int(np.floor(cx + x1 * f)) - 2), min(c1, int(np.ceil(cx + x2 * f)) + 2))]
    if A < -1e-6:
        if disc < tol:
            return [(c0, c1)]
        sq = np.sqrt(disc - tol)
        x1, x2 = clamp((-B + sq) / A + s[6]), clamp((-B - sq) / A + s[6])
        h0 = min(c1, int(np.ceil(cx + x1 * f)) + 2)
        return [(c0, h0), (max(c0, int(np.floor(cx + x2 * f)) - 2, h0 + 1), c1)]
    return [(c0, c1)]


def candidates(prm, R, W0, U, r2, H, W, strip=True):
    """boolean (H, W) mask of the pixels the raster tests for this capsule, and the number of tests with / without the strip."""
    c0, c1 = band_intervals(prm, R, W0, U, r2, H, W)
    s = strip_coefficients(R, W0, U, r2) if strip else None
    cand = np.zeros((H, W), dtype=bool)
    n_band = 0
    for bnd in range(len(c0)):
        if c0[bnd] > c1[bnd]:
            continue
        for row in range(bnd * 16, min(H, bnd * 16 + 16)):
            n_band += c1[bnd] - c0[bnd] + 1
            for lo, hi in strip_row(prm, s, row, int(c0[bnd]), int(c1[bnd])):
                if lo <= hi:
                    cand[row, lo:hi + 1] = True
    return cand, int(cand.sum()), int(n_band), s
