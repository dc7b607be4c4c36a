"""Mesh / rigid-body geometry helpers for the offline model baker (NumPy, fp64).

These restate, from MuJoCo's public documentation, what the reference obtains
from the MuJoCo *compiler* when it calls ``world.compile()``
(reference ``src/flygym/compose/base.py:21-27``): mesh volume integration,
principal-axis inertial frames, the equivalent-inertia-box capsule fit used for
``type="capsule"`` geoms that carry a mesh (``fly.py:585-589``), and capsule
mass properties.  MuJoCo itself is not available in this environment, so every
formula here is marked [PRIOR] in DESIGN.md.
"""
from __future__ import annotations

import struct

import numpy as np


# ----------------------------------------------------------------------------
# quaternions (w, x, y, z)
# ----------------------------------------------------------------------------
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w>=0)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q = q / np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def axis_angle_quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n < 1e-300:
        return np.array([1.0, 0, 0, 0])
    s = np.sin(angle / 2) / n
    return np.array([np.cos(angle / 2), axis[0] * s, axis[1] * s, axis[2] * s])


# ----------------------------------------------------------------------------
# STL + mass properties
# ----------------------------------------------------------------------------
def read_binary_stl(path) -> np.ndarray:
    """Return triangles as float64 array (n, 3 vertices, 3 xyz)."""
    with open(path, "rb") as f:
        f.read(80)
        (n,) = struct.unpack("<I", f.read(4))
        rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
        data = np.frombuffer(f.read(n * rec.itemsize), dtype=rec, count=n)
    return data["v"].astype(np.float64)


def mesh_mass_properties(tris: np.ndarray):
    """Volume, centre of mass and unit-density inertia about the COM of a closed
    triangle mesh by signed-tetrahedron integration (the "exact" mesh-inertia
    convention; identical to the legacy one for outward-oriented star-shaped
    meshes).  Orientation-agnostic: a mirrored mesh (negative scale) gives the
    mirrored result."""
    ref = tris.reshape(-1, 3).mean(axis=0)
    a, b, c = (tris[:, i] - ref for i in range(3))
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))
    V = vol6.sum() / 6.0
    sign = 1.0 if V >= 0 else -1.0
    V *= sign
    w = sign * vol6 / 6.0
    com = (w[:, None] * (a + b + c) / 4.0).sum(axis=0) / V
    s = a + b + c
    C = (
        np.einsum("n,ni,nj->ij", w, a, a)
        + np.einsum("n,ni,nj->ij", w, b, b)
        + np.einsum("n,ni,nj->ij", w, c, c)
        + np.einsum("n,ni,nj->ij", w, s, s)
    ) / 20.0
    Cc = C - V * np.outer(com, com)
    I = np.trace(Cc) * np.eye(3) - Cc
    return V, com + ref, I


def principal_axes(I: np.ndarray):
    """Diagonalise a symmetric inertia; eigenvalues in DECREASING order
    (MuJoCo's ``mju_eig3`` convention), right-handed eigenvector frame."""
    w, v = np.linalg.eigh((I + I.T) / 2)
    order = np.argsort(-w)
    w, v = w[order], v[:, order]
    # fix sign deterministically: largest-magnitude component of each axis positive
    for k in range(2):
        j = int(np.argmax(np.abs(v[:, k])))
        if v[j, k] < 0:
            v[:, k] = -v[:, k]
    v[:, 2] = np.cross(v[:, 0], v[:, 1])
    return w, v


def inertia_box_halfsizes(mass: float, diag_inertia: np.ndarray) -> np.ndarray:
    """Half-sizes of the uniform box with the given mass and principal inertia.
    I0 = m/3 (b1^2 + b2^2) etc."""
    I0, I1, I2 = diag_inertia
    return np.array([
        np.sqrt(max(0.0, 6 * (I1 + I2 - I0) / mass)) / 2,
        np.sqrt(max(0.0, 6 * (I0 + I2 - I1) / mass)) / 2,
        np.sqrt(max(0.0, 6 * (I0 + I1 - I2) / mass)) / 2,
    ])


def fit_capsule(box: np.ndarray):
    """[PRIOR] MuJoCo mesh->capsule fit with ``fitaabb=false``: radius = mean of
    the two smaller equivalent-inertia-box half-sizes, half-length = largest
    half-size minus radius/2.  Capsule axis = local z (smallest-inertia axis)."""
    radius = 0.5 * (box[0] + box[1])
    half = max(0.0, box[2] - radius / 2)
    return radius, half


def capsule_inertia(mass: float, radius: float, half: float) -> np.ndarray:
    """Principal inertia (Ixx=Iyy, Izz) of a solid capsule of given total mass."""
    h = 2 * half
    vc = np.pi * radius**2 * h
    vs = 4.0 / 3.0 * np.pi * radius**3
    mc = mass * vc / (vc + vs)
    ms = mass * vs / (vc + vs)
    izz = mc * radius**2 / 2 + ms * 2 * radius**2 / 5
    ixx = mc * (3 * radius**2 + h**2) / 12 + ms * (2 * radius**2 / 5 + h**2 / 4 + 3 * h * radius / 8)
    return np.array([ixx, ixx, izz])


def convex_hull_vertices(points: np.ndarray):
    """Hull vertices (sorted by original index) and their adjacency lists (edges of the triangulated hull facets)."""
    from scipy.spatial import ConvexHull

    up = np.unique(points, axis=0)
    hull = ConvexHull(up)
    order = np.sort(hull.vertices)
    remap = {int(v): i for i, v in enumerate(order)}
    nbrs = [set() for _ in order]
    for tri in hull.simplices:
        a, b, c = (remap[int(t)] for t in tri)
        nbrs[a].update((b, c)); nbrs[b].update((a, c)); nbrs[c].update((a, b))
    return up[order], [sorted(n) for n in nbrs]
