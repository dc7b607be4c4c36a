"""Offline model baker: reference assets -> flat ``NMFModel`` (no MuJoCo needed).

Re-creates what the reference obtains from the MuJoCo compiler when it builds
the benchmark world (``src/flygym_demo/benchmark/time_gpu_simulation.py:21-64``,
identical to ``tests/warp/conftest.py:25-72``):

* bodies/geoms   <- ``compose/fly.py:545-612`` + ``assets/model/rigging.yaml``
* meshes         <- ``compose/fly.py:507-543`` (scale 1000, right side = mirrored left)
* joints         <- ``compose/fly.py:221-299`` (hinges; stiffness 10, damping 0.5,
                    armature 1e-6, springref = neutral angle; R-side roll/yaw axes negated)
* actuators      <- ``compose/fly.py:301-369`` (position, forcerange +-30) and
                    ``:407-441`` (adhesion on tarsus5, ctrlrange (1,100))
* world/contacts <- ``compose/world.py:263-331`` + ``compose/physics.py:60-111``
* options        <- ``assets/model/mujoco_globals.yaml``
* keyframe       <- ``compose/world.py:151-207``, ``compose/fly.py:653-678``

Run as ``python -m flygym_b200.baker.bake --reference /root/reference`` to
regenerate ``flygym_b200/assets/nmf_bench_{capsule,mesh}.npz``.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import yaml

from .. import anatomy as A
from ..model import (ASSETS_DIR, CONTACT_FIELDS, DIM_FIELDS, GEOM_CAPSULE, GEOM_HULL, OPT_FIELDS,
                     NMFModel)
from . import geometry as G

MESH_SCALE = 1000.0  # metres -> mm  (fly.py:510-512)


# ----------------------------------------------------------------------------
def _load_neutral_pose(assets: Path, axis_order: A.AxisOrder) -> dict[str, float]:
    """``KinematicPosePreset.NEUTRAL`` (pose.py:80-128,147-161): degrees -> rad,
    left side copied to the right when the right entry is missing."""
    path = assets / "model/pose/neutral" / ("_".join(axis_order.value) + ".yaml")
    with open(path) as f:
        data = yaml.safe_load(f)
    angles = {k: float(v) for k, v in data["joint_angles"].items()}
    if data.get("angle_unit") == "degree":
        angles = {k: float(np.deg2rad(v)) for k, v in angles.items()}
    for name, ang in list(angles.items()):
        parent, child, axis = name.split("-")
        if child[0] != "l":
            continue
        mparent = ("r" + parent[1:]) if parent[0] == "l" else parent
        mname = f"{mparent}-r{child[1:]}-{axis}"
        angles.setdefault(mname, ang)
    return angles


def _segment_mesh(assets: Path, seg: str) -> np.ndarray:
    """Triangles of a segment in its body frame, mm (fly.py:507-543)."""
    mesh_dir = assets / "model/meshes/simplified_max2000faces"
    fallback = assets / "model/meshes/fullsize"
    src, ysign = (("l" + seg[1:]), -1.0) if seg[0] == "r" else (seg, 1.0)
    path = mesh_dir / f"{src}.stl"
    if not path.exists():
        path = fallback / f"{src}.stl"
    tris = G.read_binary_stl(path) * MESH_SCALE
    tris[..., 1] *= ysign
    return tris


class _Frame:
    """Rigid transform (pos, quat) with composition."""

    def __init__(self, pos=(0, 0, 0), quat=(1, 0, 0, 0)):
        self.pos = np.asarray(pos, dtype=np.float64)
        self.quat = G.quat_normalize(quat)

    @property
    def mat(self):
        return G.quat_to_mat(self.quat)

    def __matmul__(self, other: "_Frame") -> "_Frame":
        return _Frame(self.pos + self.mat @ other.pos, G.quat_mul(self.quat, other.quat))

    def apply(self, p):
        return self.pos + (self.mat @ np.asarray(p).T).T


# ----------------------------------------------------------------------------
def bake(
    reference_root: str | Path,
    *,
    joint_preset: str = "legs_only",
    axis_order: A.AxisOrder = A.AxisOrder.YAW_PITCH_ROLL,
    actuated_preset: str = "legs_active_only",
    position_gain: float = 50.0,
    spawn_position=(0.0, 0.0, 0.8),
    spawn_quat=(1.0, 0.0, 0.0, 0.0),
    simplify_geom: bool = False,
    contact_preset: str = "legs_thorax_abdomen_head",
    stiffness: float = 10.0,
    damping: float = 0.5,
    armature: float = 1e-6,
    adhesion_gain: float = 1.0,
    # ContactParams defaults (physics.py:60-77); note the reference passes a
    # 4-tuple solimp (physics.py:103-111) so MuJoCo reads it as
    # (dmin, dmax, width=0.5, midpoint=3.0) with power left at its default 2.
    sliding_friction: float = 1.0,
    solref=(2e-4, 1.0),
    solimp=(0.98, 0.99, 0.5, 3.0, 2.0),
    margin: float = 1e-3,
    noslip_iterations: int = 0,
) -> NMFModel:
    ref = Path(reference_root)
    assets = ref / "src/flygym/assets"
    with open(assets / "model/rigging.yaml") as f:
        rigging = yaml.safe_load(f)
    with open(assets / "model/mujoco_globals.yaml") as f:
        glob = yaml.safe_load(f)
    boundmass = float(glob["compiler"]["boundmass"])
    boundinertia = float(glob["compiler"]["boundinertia"])
    neutral = _load_neutral_pose(assets, axis_order)

    segs = A.bodysegs_order()
    seg_parent = {c: p for p, c in A.ALL_CONNECTED_SEGMENT_PAIRS}
    dofs = A.jointdofs_order(joint_preset, axis_order)
    dofs_by_child: dict[str, list[A.JointDOF]] = {}
    for d in dofs:
        dofs_by_child.setdefault(d.child, []).append(d)

    # ---- movable bodies: hub (free joint) + every segment that owns hinge DoFs
    movable = [s for s in segs if s in dofs_by_child]
    body_names = ["hub"] + movable
    body_id = {n: i for i, n in enumerate(body_names)}
    nbody = len(body_names)

    def owner(seg):  # nearest movable ancestor-or-self ("hub" if none)
        while seg is not None and seg not in dofs_by_child:
            seg = seg_parent.get(seg)
        return "hub" if seg is None else seg

    # segment frame relative to parent segment (rigging); the root c_thorax sits
    # at rigging pos inside the attachment frame that carries the free joint
    # (world.py:276-279) -> that attachment frame is our hub frame.
    seg_local = {s: _Frame(rigging[s]["pos"], rigging[s]["quat"]) for s in segs}
    seg_in_owner: dict[str, _Frame] = {}
    for s in segs:  # DFS order => parents first
        if s in dofs_by_child:
            seg_in_owner[s] = _Frame()
        else:
            p = seg_parent.get(s)
            base = _Frame() if p is None else seg_in_owner[p] if owner(p) == owner(s) else _Frame()
            seg_in_owner[s] = base @ seg_local[s]

    body_parent = np.full(nbody, -1, np.int32)
    body_pos = np.zeros((nbody, 3))
    body_quat = np.zeros((nbody, 4))
    body_pos[0], body_quat[0] = spawn_position, G.quat_normalize(spawn_quat)
    for s in movable:
        p = seg_parent[s]
        fr = seg_in_owner[p] @ seg_local[s] if p not in dofs_by_child else seg_local[s]
        b = body_id[s]
        body_parent[b] = body_id[owner(p)]
        body_pos[b], body_quat[b] = fr.pos, fr.quat

    # ---- per-segment geometry & mass properties ---------------------------
    contact_segs = A.contact_bodies(contact_preset)
    seg_mass, seg_com, seg_inertia = {}, {}, {}
    seg_geom = {}
    viscap = {}
    for s in segs:
        tris = _segment_mesh(assets, s)
        mass = float(rigging[s]["mass"])
        V, com, I_unit = G.mesh_mass_properties(tris)
        w, ax = G.principal_axes(I_unit)
        is_capsule = simplify_geom or (A.is_leg(s) and A.seg_link(s) == "tarsus5")  # fly.py:585-589
        box = G.inertia_box_halfsizes(V, w)
        radius, half = G.fit_capsule(box)
        viscap[s] = (com, ax[:, 2].copy(), radius, half)      # capsule proxy of EVERY segment: what the eye cameras see of the body
        if is_capsule:
            I_seg = ax @ np.diag(G.capsule_inertia(mass, radius, half)) @ ax.T
            seg_geom[s] = dict(type=GEOM_CAPSULE, pos=com, quat=G.mat_to_quat(ax), size=(radius, half),
                               verts=np.zeros((0, 3)), nbrs=[])
        else:
            I_seg = I_unit * (mass / V)
            hv, hn = G.convex_hull_vertices(tris.reshape(-1, 3))
            seg_geom[s] = dict(type=GEOM_HULL, pos=com, quat=G.mat_to_quat(ax), size=(0.0, 0.0), verts=hv, nbrs=hn)
        seg_mass[s], seg_com[s], seg_inertia[s] = mass, com, I_seg

    # ---- fuse static segments into their owner, then bound mass/inertia ----
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.zeros((nbody, 4))
    body_inertia = np.zeros((nbody, 3))
    for bname in body_names:
        members = [s for s in segs if owner(s) == bname]
        m = sum(seg_mass[s] for s in members)
        com = sum(seg_mass[s] * seg_in_owner[s].apply(seg_com[s]) for s in members) / m
        I = np.zeros((3, 3))
        for s in members:
            R = seg_in_owner[s].mat
            d = seg_in_owner[s].apply(seg_com[s]) - com
            I += R @ seg_inertia[s] @ R.T + seg_mass[s] * (d @ d * np.eye(3) - np.outer(d, d))
        w, ax = G.principal_axes(I)
        b = body_id[bname]
        body_mass[b] = max(m, boundmass)
        body_inertia[b] = np.maximum(w, boundinertia)
        body_ipos[b], body_iquat[b] = com, G.mat_to_quat(ax)

    # ---- DoFs ---------------------------------------------------------------
    nv = 6 + len(dofs)
    nq = 7 + len(dofs)
    dof_body = np.zeros(nv, np.int32)
    dof_parent = np.full(nv, -1, np.int32)
    dof_axis = np.zeros((nv, 3))
    body_dofadr = np.zeros(nbody, np.int32)
    body_dofnum = np.zeros(nbody, np.int32)
    body_dofnum[0] = 6
    dof_parent[1:6] = np.arange(5)
    last_dof_of_body = {0: 5}
    springref = np.zeros(nv)
    k = 6
    dof_index: dict[str, int] = {}
    for s in movable:
        b = body_id[s]
        body_dofadr[b] = k
        prev = last_dof_of_body[int(body_parent[b])]
        for d in dofs_by_child[s]:
            vec = np.array(A.AXIS_VECTOR[d.axis])
            if s[0] == "r" and d.axis != "pitch":  # fly.py:279-283
                vec = -vec
            dof_axis[k] = vec
            dof_body[k] = b
            dof_parent[k] = prev
            springref[k] = neutral.get(d.name, 0.0)
            dof_index[d.name] = k
            prev = k
            k += 1
        body_dofnum[b] = k - body_dofadr[b]
        last_dof_of_body[b] = k - 1
    dof_stiffness = np.r_[np.zeros(6), np.full(nv - 6, stiffness)]
    dof_damping = np.r_[np.zeros(6), np.full(nv - 6, damping)]
    dof_armature = np.r_[np.zeros(6), np.full(nv - 6, armature)]

    # leg bookkeeping (contact sensors: subtree of the most proximal contact
    # segment of each leg, world.py:311-331)
    body_leg = np.full(nbody, -1, np.int32)
    for s in movable:
        if A.is_leg(s):
            body_leg[body_id[s]] = A.LEGS.index(A.seg_pos(s))
    leg_rootbody = np.full(len(A.LEGS), -1, np.int32)
    for li, leg in enumerate(A.LEGS):
        cs = [s for s in contact_segs if A.seg_pos(s) == leg]
        if cs:
            cs.sort(key=lambda s: A.LEG_LINKS.index(A.seg_link(s)))
            leg_rootbody[li] = body_id[owner(cs[0])]

    # ---- actuators ----------------------------------------------------------
    act = A.actuated_dofs(dofs, actuated_preset)
    act_dof = np.array([dof_index[d.name] for d in act], np.int32)
    act_kp = np.full(len(act), position_gain)
    act_kv = np.zeros(len(act))
    act_frcrange = np.tile([-30.0, 30.0], (len(act), 1))
    adh_body = np.array([body_id[owner(f"{leg}_tarsus5")] for leg in A.LEGS], np.int32)
    adh_gain = np.full(len(A.LEGS), adhesion_gain)
    adh_ctrlrange = np.tile([1.0, 100.0], (len(A.LEGS), 1))

    # ---- contact geoms ------------------------------------------------------
    geom_body, geom_type, geom_pos, geom_quat, geom_size = [], [], [], [], []
    geom_vertadr, geom_vertnum, hull = [], [], []
    nbr_adr, nbr = [0], []          # CSR adjacency of the hull vertices (indices local to the geom)
    nvert = 0
    for s in contact_segs:
        g = seg_geom[s]
        fr = seg_in_owner[s] @ _Frame(g["pos"], g["quat"])
        geom_body.append(body_id[owner(s)])
        geom_type.append(g["type"])
        geom_pos.append(fr.pos)
        geom_quat.append(fr.quat)
        geom_size.append(g["size"])
        v = seg_in_owner[s].apply(g["verts"]) if len(g["verts"]) else np.zeros((0, 3))
        geom_vertadr.append(nvert)
        geom_vertnum.append(len(v))
        hull.append(v)
        for lst in g["nbrs"]:
            nbr.extend(lst)
            nbr_adr.append(len(nbr))
        nvert += len(v)
    hull_vert = np.concatenate(hull) if nvert else np.zeros((0, 3))

    # ---- sites (child-segment origin of every anatomical joint, fly.py:371-405)
    site_pairs = list(A.ALL_CONNECTED_SEGMENT_PAIRS)
    site_body = np.array([body_id[owner(c)] for _, c in site_pairs], np.int32)
    site_pos = np.array([seg_in_owner[c].pos for _, c in site_pairs])

    seg_body = np.array([body_id[owner(s)] for s in segs], np.int32)
    seg_pos = np.array([seg_in_owner[s].pos for s in segs])
    seg_quat = np.array([seg_in_owner[s].quat for s in segs])

    # ---- keyframe "neutral" (world.py:151-207) ------------------------------
    key_qpos = np.zeros(nq)
    key_qpos[:3], key_qpos[3:7] = spawn_position, G.quat_normalize(spawn_quat)
    for d in dofs:
        key_qpos[dof_index[d.name] + 1] = neutral.get(d.name, 0.0)
    key_ctrl = np.r_[[neutral.get(d.name, 0.0) for d in act], np.zeros(len(A.LEGS))]

    o = glob["option"]
    opt = dict(timestep=float(o["timestep"]), gx=float(o["gravity"][0]), gy=float(o["gravity"][1]),
               gz=float(o["gravity"][2]), iterations=float(o["iterations"]), tolerance=1e-8,
               ls_iterations=50.0, ls_tolerance=0.01, noslip_iterations=float(noslip_iterations),
               meaninertia=0.0, impratio=1.0,
               multiccd=1.0 if str(glob.get("option", {}).get("flag", {}).get("multiccd", "disable")) == "enable" else 0.0)
    contact = dict(mu=sliding_friction, solref0=solref[0], solref1=solref[1], solimp0=solimp[0],
                   solimp1=solimp[1], solimp2=solimp[2], solimp3=solimp[3], solimp4=solimp[4],
                   margin=margin, gap=0.0)
    dims = dict(nbody=nbody, nq=nq, nv=nv, nu_pos=len(act), nu_adh=len(A.LEGS), ngeom=len(geom_body),
                nsite=len(site_pairs), nseg=len(segs), nleg=len(A.LEGS), nhullvert=nvert)

    arrays = dict(
        dims=np.array([dims[k] for k in DIM_FIELDS], np.int32),
        opt=np.array([opt[k] for k in OPT_FIELDS]),
        contact=np.array([contact[k] for k in CONTACT_FIELDS], dtype=np.float64),
        body_parent=body_parent, body_pos=body_pos, body_quat=body_quat, body_mass=body_mass,
        body_ipos=body_ipos, body_iquat=body_iquat, body_inertia=body_inertia,
        body_invweight0=np.zeros((nbody, 2)), body_dofadr=body_dofadr, body_dofnum=body_dofnum,
        body_leg=body_leg,
        dof_body=dof_body, dof_parent=dof_parent, dof_axis=dof_axis, dof_stiffness=dof_stiffness,
        dof_damping=dof_damping, dof_armature=dof_armature, dof_springref=springref,
        act_dof=act_dof, act_kp=act_kp, act_kv=act_kv, act_frcrange=act_frcrange,
        adh_body=adh_body, adh_gain=adh_gain, adh_ctrlrange=adh_ctrlrange,
        geom_body=np.array(geom_body, np.int32), geom_type=np.array(geom_type, np.int32),
        geom_pos=np.array(geom_pos), geom_quat=np.array(geom_quat), geom_size=np.array(geom_size),
        geom_vertadr=np.array(geom_vertadr, np.int32), geom_vertnum=np.array(geom_vertnum, np.int32),
        hull_vert=hull_vert, hull_nbr_adr=np.array(nbr_adr, np.int32), hull_nbr=np.array(nbr if nbr else [0], np.int32),
        site_body=site_body, site_pos=site_pos, seg_body=seg_body, seg_pos=seg_pos, seg_quat=seg_quat,
        leg_rootbody=leg_rootbody, key_qpos=key_qpos, key_ctrl=key_ctrl,
        # not part of the blob: capsule proxy of every segment in its own frame (centre, axis, radius, half length), used by the
        # eye cameras to draw the fly's own body (flygym_b200/retina.py)
        viscap_pos=np.array([viscap[s][0] for s in segs]), viscap_axis=np.array([viscap[s][1] for s in segs]),
        viscap_size=np.array([[viscap[s][2], viscap[s][3]] for s in segs]),
    )
    names = dict(
        bodies=body_names, segments=segs, jointdofs=[d.name for d in dofs],
        actuated_position=[d.name for d in act], legs=list(A.LEGS),
        contact_geoms=contact_segs, sites=[f"{p}-{c}" for p, c in site_pairs],
    )
    meta = dict(joint_preset=joint_preset, axis_order="_".join(axis_order.value),
                actuated_preset=actuated_preset, position_gain=position_gain,
                simplify_geom=bool(simplify_geom), contact_preset=contact_preset,
                units="mm, g, s (forces in uN)", source="flygym 2.0.1 assets")
    model = NMFModel(arrays, names, meta)
    set_const(model)
    return model


# ----------------------------------------------------------------------------
def kinematics_qpos0(model: NMFModel):
    """World pose of every movable body at qpos0 (hinges at 0, hub at spawn)."""
    a = model.arrays
    nb = model.nbody
    xpos = np.zeros((nb, 3))
    xquat = np.zeros((nb, 4))
    for b in range(nb):
        p = a["body_parent"][b]
        if p < 0:
            xpos[b], xquat[b] = a["body_pos"][b], a["body_quat"][b]
        else:
            xpos[b] = xpos[p] + G.quat_to_mat(xquat[p]) @ a["body_pos"][b]
            xquat[b] = G.quat_mul(xquat[p], a["body_quat"][b])
    return xpos, xquat


def set_const(model: NMFModel) -> None:
    """[PRIOR] ``mj_setConst`` subset: ``body_invweight0`` (mean translational /
    rotational diagonal of J M^-1 J^T at each body's COM, at qpos0) and
    ``stat.meaninertia`` (mean diagonal of M at qpos0).

    Locked DoFs (``model.arrays['locked_dofs']``, see ``NMFModel.with_locked_dofs``) stand for rigid attachments: the
    bodies they connect are treated as ONE body, as the MuJoCo compiler would see them after fusing -- the inverse weight is
    evaluated at the group's common COM and shared by its members, and ``meaninertia`` averages over the free DoFs only."""
    a = model.arrays
    nb, nv = model.nbody, model.nv
    xpos, xquat = kinematics_qpos0(model)
    xmat = np.array([G.quat_to_mat(q) for q in xquat])
    xipos = np.array([xpos[b] + xmat[b] @ a["body_ipos"][b] for b in range(nb)])

    def jac(b, point):
        """6 x nv: rows 0-2 translational, 3-5 rotational (world frame)."""
        J = np.zeros((6, nv))
        J[0:3, 0:3] = np.eye(3)
        R0 = xmat[0]
        for k in range(3):
            J[3:6, 3 + k] = R0[:, k]
            J[0:3, 3 + k] = np.cross(R0[:, k], point - xpos[0])
        d = a["body_dofadr"][b] + a["body_dofnum"][b] - 1 if b > 0 else -1
        while d >= 6:
            bd = a["dof_body"][d]
            axis = xmat[bd] @ a["dof_axis"][d]
            J[3:6, d] = axis
            J[0:3, d] = np.cross(axis, point - xpos[bd])  # hinge anchors sit at body origins
            d = a["dof_parent"][d]
        return J

    M = np.diag(a["dof_armature"]).astype(np.float64)
    for b in range(nb):
        J = jac(b, xipos[b])
        Rw = xmat[b] @ G.quat_to_mat(a["body_iquat"][b])
        Iw = Rw @ np.diag(a["body_inertia"][b]) @ Rw.T
        M += a["body_mass"][b] * J[0:3].T @ J[0:3] + J[3:6].T @ Iw @ J[3:6]
    Minv = np.linalg.inv(M)
    locked = set(int(d) for d in a.get("locked_dofs", []))
    group = list(range(nb))                 # representative (most proximal member) of each rigid group
    for b in range(1, nb):
        adr, num = int(a["body_dofadr"][b]), int(a["body_dofnum"][b])
        if num > 0 and all(d in locked for d in range(adr, adr + num)):
            group[b] = group[int(a["body_parent"][b])]
    inv = np.zeros((nb, 2))
    for g in sorted(set(group)):
        members = [b for b in range(nb) if group[b] == g]
        mass = a["body_mass"][members]
        com = xipos[members[0]] if len(members) == 1 else (mass[:, None] * xipos[members]).sum(0) / mass.sum()
        J = jac(g, com)
        Ab = J @ Minv @ J.T
        inv[members, 0] = np.trace(Ab[0:3, 0:3]) / 3
        inv[members, 1] = np.trace(Ab[3:6, 3:6]) / 3
    a["body_invweight0"] = inv
    free = [d for d in range(nv) if d not in locked]
    opt = a["opt"].copy(); opt[OPT_FIELDS.index("meaninertia")] = np.trace(M[np.ix_(free, free)]) / len(free); a["opt"] = opt


LOCK_ARMATURE = NMFModel.LOCK_ARMATURE


def bake_kernel_layout(reference_root, joint_preset: str = "legs_active_only", **kw) -> NMFModel:
    """A joint preset with fewer leg DoFs, laid out for the sm_100a kernels (hub + 6 chains of 8 links, 11 DoFs each).

    The MuJoCo compiler fuses links without joints into their parent (and applies ``boundmass`` / ``boundinertia`` to the
    FUSED body, ``mujoco_globals.yaml:6-7``), so the reduced model is baked for real first; its bodies are then spread back
    over the LEGS_ONLY lanes: a link that kept its joint carries the fused body's inertia, the links fused into it become
    massless lanes whose DoFs are locked (armature ``LOCK_ARMATURE``, no spring / damper, keyframe angle 0 -- a rigid
    attachment is the infinite-armature limit) and share its inverse weight.  Geoms keep their own lanes."""
    true = bake(reference_root, joint_preset=joint_preset, **kw)
    full = bake(reference_root, joint_preset="legs_only", **kw)
    a, t = dict(full.arrays), true.arrays
    tb = {n: i for i, n in enumerate(true.names["bodies"])}
    nb = full.nbody
    mass, ipos, iquat, inertia, invw = (a[k].copy() for k in ("body_mass", "body_ipos", "body_iquat", "body_inertia", "body_invweight0"))
    ipos, iquat, inertia, invw = ipos.reshape(nb, 3), iquat.reshape(nb, 4), inertia.reshape(nb, 3), invw.reshape(nb, 2)
    tt = {k: t[k].reshape(true.nbody, -1) for k in ("body_ipos", "body_iquat", "body_inertia", "body_invweight0")}
    locked = []
    for b, name in enumerate(full.names["bodies"]):
        if name in tb:
            i = tb[name]
            mass[b] = t["body_mass"][i]; ipos[b] = tt["body_ipos"][i]; iquat[b] = tt["body_iquat"][i]
            inertia[b] = tt["body_inertia"][i]; invw[b] = tt["body_invweight0"][i]
        else:                                   # fused into the nearest ancestor that kept a joint
            anc = int(a["body_parent"][b])
            while full.names["bodies"][anc] not in tb:
                anc = int(a["body_parent"][anc])
            mass[b] = 0.0; ipos[b] = 0.0; iquat[b] = (1.0, 0.0, 0.0, 0.0); inertia[b] = 0.0
            invw[b] = tt["body_invweight0"][tb[full.names["bodies"][anc]]]
            locked += list(range(int(a["body_dofadr"][b]), int(a["body_dofadr"][b]) + int(a["body_dofnum"][b])))
    a.update(body_mass=mass, body_ipos=ipos, body_iquat=iquat, body_inertia=inertia, body_invweight0=invw)
    locked = np.array(sorted(locked), dtype=np.int32)
    for key, v in (("dof_stiffness", 0.0), ("dof_damping", 0.0), ("dof_armature", LOCK_ARMATURE), ("dof_springref", 0.0)):
        arr = a[key].copy(); arr[locked] = v; a[key] = arr
    key = a["key_qpos"].copy(); key[locked + 1] = 0.0; a["key_qpos"] = key
    a["locked_dofs"] = locked
    opt = a["opt"].copy(); opt[OPT_FIELDS.index("meaninertia")] = true.opt("meaninertia"); a["opt"] = opt
    if true.names["actuated_position"] != full.names["actuated_position"]:
        raise ValueError("the reduced preset changes the actuator set; not representable in the kernel layout")
    lockset = set(int(d) - 6 for d in locked)
    names = dict(full.names, jointdofs=list(true.names["jointdofs"]),
                 locked_jointdofs=[nm for j, nm in enumerate(full.names["jointdofs"]) if j in lockset])
    assert names["jointdofs"] == [nm for j, nm in enumerate(full.names["jointdofs"]) if j not in lockset]
    return NMFModel(a, names, dict(true.meta, kernel_layout="legs_only lanes, fused links massless + locked"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=str(ASSETS_DIR))
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    for simplify, name in ((True, "nmf_bench_capsule.npz"), (False, "nmf_bench_mesh.npz")):
        m = bake(args.reference, simplify_geom=simplify)
        m.save(out / name)
        print(name, {k: m.dim(k) for k in DIM_FIELDS}, "mass", m.arrays["body_mass"].sum(),
              "meaninertia", m.opt("meaninertia"))
    m = bake_kernel_layout(args.reference, "legs_active_only", simplify_geom=True)
    m.save(out / "nmf_bench_capsule_legs_active_only.npz")
    print("nmf_bench_capsule_legs_active_only.npz", {k: m.dim(k) for k in DIM_FIELDS}, "free hinge DoFs", len(m.names["jointdofs"]))
    # general-topology models (stepped by the tree kernels): every joint preset with every segment as a contact body
    # (ContactBodiesPreset.ALL; NMFModel.with_contact_bodies narrows it), capsule geoms; ALL_BIOLOGICAL also with mesh hulls
    for preset, simplify, name in (("legs_only", True, "nmf_legs_only_allcontacts_capsule.npz"), ("all_biological", True, "nmf_all_biological_capsule.npz"),
                                   ("all_possible", True, "nmf_all_possible_capsule.npz"), ("all_biological", False, "nmf_all_biological_mesh.npz")):
        m = bake(args.reference, joint_preset=preset, simplify_geom=simplify, contact_preset="all")
        m.save(out / name)
        print(name, {k: m.dim(k) for k in DIM_FIELDS})


if __name__ == "__main__":
    main()
