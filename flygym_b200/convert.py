"""``MjModel`` -> baked :class:`~flygym_b200.model.NMFModel` (the blob the sm_100a kernels ingest).

The reference hands its backends a MuJoCo-compiled model: ``Simulation.__init__`` calls ``world.compile()``
(reference ``src/flygym/simulation.py:32-57``, ``compose/base.py:21-27``) and ``GPUSimulation`` uploads that
``mj_model`` (``warp/simulation.py:49-62,416-425``).  :func:`from_mjmodel` is the matching ingestion step of this
backend: it reads ONLY public ``MjModel`` fields (``body_*``, ``jnt_*``, ``dof_*``, ``geom_*``, ``mesh_*``,
``pair_*``, ``actuator_*``, ``eq_*``, ``opt.*``, ``key_*``, ``stat.meaninertia``), so it works on a real
``mujoco.MjModel`` and on any object that exposes the same attributes (MuJoCo is not installable in this
environment; the tests drive it with a duck-typed model synthesised by :func:`mjmodel_like`).

What it does: finds the free body, fuses joint-less bodies into the nearest ancestor that has joints (mass, centre of
mass and inertia combined exactly; geoms, sites and segment frames re-expressed in the owner's frame), checks that
the result is something the kernels step -- a free hub carrying a tree of bodies with up to three hinges each: the hub + 6
chains of 8 links with 3, 2, 1, 1, 1, 1, 1, 1 hinges of ``JointPreset.LEGS_ONLY`` runs on the star kernels, anything else
(``ALL_BIOLOGICAL``, ``ALL_POSSIBLE``) on the general-topology kernels -- and lays DoFs, actuators and contact geoms out in
the kernels' order.  Index maps from MuJoCo addresses to the
kernel layout are returned in ``model.meta['mj_maps']`` -- a ``Simulation`` subclass indexes the state record through
them, since MuJoCo's ``jnt_qposadr`` / actuator ids need not coincide with the record layout.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from .baker import geometry as G
from .model import (CONTACT_FIELDS, DIM_FIELDS, GEOM_CAPSULE, GEOM_HULL, OPT_FIELDS, WELD_FIELDS, NMFModel)

# MuJoCo enums (mjtJoint, mjtGeom, mjtTrn, mjtEq, mjtEnableBit)
JNT_FREE, JNT_HINGE = 0, 3
GEOM_PLANE, GEOM_MJ_CAPSULE, GEOM_MESH = 0, 3, 7
TRN_JOINT, TRN_BODY = 0, 5
EQ_WELD = 1
ENBL_MULTICCD = 1 << 4
LEG_DOFS = (3, 2, 1, 1, 1, 1, 1, 1)


class ConversionError(ValueError):
    """The MuJoCo model is outside what the sm_100a kernels handle (topology, joint or actuator kinds ...)."""


def _names(m, kind: str, n: int) -> list[str]:
    """Element names: ``m.names[kind]`` on duck-typed models, ``m.<kind>(i).name`` on a real ``mujoco.MjModel``."""
    tab = getattr(m, "names", None)
    if isinstance(tab, dict) and kind in tab:
        return list(tab[kind])
    acc = getattr(m, kind, None)
    if callable(acc):
        return [acc(i).name for i in range(n)]
    raise ConversionError(f"cannot read {kind} names from this model object")


def _short(name: str) -> str:
    """dm_control prefixes attached elements with the fly's name: 'nmf/l_eye' -> 'l_eye'."""
    return name.split("/")[-1]


class _T:
    """rigid transform"""

    def __init__(self, pos=(0.0, 0.0, 0.0), quat=(1.0, 0.0, 0.0, 0.0)):
        self.pos = np.asarray(pos, dtype=np.float64); self.quat = G.quat_normalize(quat)

    @property
    def mat(self): return G.quat_to_mat(self.quat)

    def __matmul__(self, o): return _T(self.pos + self.mat @ o.pos, G.quat_mul(self.quat, o.quat))

    def apply(self, p): return self.pos + self.mat @ np.asarray(p, dtype=np.float64)


def from_mjmodel(m, *, fly_root: str | None = None, segments: list[str] | None = None) -> NMFModel:
    """Convert a compiled MuJoCo model holding ONE fly (``FlatGroundWorld`` or ``TetheredWorld``) into an ``NMFModel``.

    ``fly_root``: name of the body that carries the free joint (default: the only body with one).  ``segments``: anatomical
    body-segment names in Fly order (default: ``flygym_b200.anatomy.bodysegs_order()``); every one that exists as a body of the
    model becomes a row of the segment table behind ``get_body_positions`` / ``get_body_rotations``."""
    from . import anatomy as A
    nbody, njnt, nv_mj, nq_mj = int(m.nbody), int(m.njnt), int(m.nv), int(m.nq)
    body_names = [_short(s) for s in _names(m, "body", nbody)]
    jnt_names = [_short(s) for s in _names(m, "joint", njnt)]
    jnt_type = np.asarray(m.jnt_type); jnt_bodyid = np.asarray(m.jnt_bodyid)
    parent = np.asarray(m.body_parentid)

    # ---- the free body
    free = [j for j in range(njnt) if jnt_type[j] == JNT_FREE]
    if fly_root is not None:
        free = [j for j in free if body_names[jnt_bodyid[j]] == _short(fly_root)]
    if len(free) != 1:
        raise ConversionError(f"expected exactly one free joint (one fly per world record), found {len(free)}")
    jfree = free[0]; hub_b = int(jnt_bodyid[jfree])
    if any(t not in (JNT_FREE, JNT_HINGE) for t in jnt_type):
        raise ConversionError("only free + hinge joints are supported")
    if np.abs(np.asarray(m.jnt_pos)[[j for j in range(njnt) if jnt_type[j] == JNT_HINGE]]).max(initial=0.0) > 1e-12:
        raise ConversionError("hinge anchors must sit at the child-body origin (jnt_pos = 0), as flygym composes them (fly.py:285-295)")

    # ---- subtree of the free body; owner = nearest ancestor-or-self that has joints
    in_fly = np.zeros(nbody, bool); in_fly[hub_b] = True
    for b in range(hub_b + 1, nbody):                     # MuJoCo numbers parents before children
        in_fly[b] = in_fly[parent[b]]
    has_jnt = np.asarray(m.body_jntnum) > 0
    owner = np.full(nbody, -1)
    rel = {}                                              # body -> transform in its owner's frame
    for b in range(nbody):
        if not in_fly[b]:
            continue
        if has_jnt[b]:
            owner[b] = b; rel[b] = _T()
        else:
            owner[b] = owner[parent[b]]
            rel[b] = rel[parent[b]] @ _T(m.body_pos[b], m.body_quat[b])
    movable = [b for b in range(nbody) if in_fly[b] and has_jnt[b] and b != hub_b]

    # ---- kernel body order = hub, then the jointed bodies in MuJoCo order (parents before children; a chain stays contiguous).
    # A hub + 6 chains of 8 links with (3, 2, 1, 1, 1, 1, 1, 1) hinges (JointPreset.LEGS_ONLY) is what the star kernels step;
    # any other tree of hinge-jointed bodies (ALL_BIOLOGICAL, ALL_POSSIBLE ...) goes to the general-topology kernels.  The
    # native library decides (nmf_create); here only what neither handles is rejected.
    mov_parent = {b: int(owner[parent[b]]) for b in movable}
    dofnum_mj = np.asarray(m.body_dofnum)
    if any(int(dofnum_mj[b]) > 3 for b in movable):
        raise ConversionError("a body carries at most three hinge joints (one anatomical joint, anatomy.py:411-416)")
    leg_names = list(A.LEGS)
    def leg_of(b):               # leg index of a jointed body by its segment name ('lf_coxa' -> 0), -1 off the legs
        nm = body_names[b]
        return leg_names.index(nm.split("_")[0]) if A.is_leg(nm) and nm.split("_")[0] in leg_names else -1
    kbodies = [hub_b] + movable                           # kernel body index -> MuJoCo body id
    kidx = {b: i for i, b in enumerate(kbodies)}
    nb = len(kbodies)

    # ---- fuse static bodies into their owner: mass, COM, inertia
    body_mass = np.zeros(nb); body_ipos = np.zeros((nb, 3)); body_iquat = np.zeros((nb, 4)); body_inertia = np.zeros((nb, 3))
    members = {o: [b for b in range(nbody) if in_fly[b] and owner[b] == o] for o in kbodies}
    for o in kbodies:
        ms = np.array([float(m.body_mass[b]) for b in members[o]])
        cs = np.array([rel[b].apply(m.body_ipos[b]) for b in members[o]])
        M = ms.sum()
        if M <= 0:
            raise ConversionError(f"body {body_names[o]} has no mass")
        com = (ms[:, None] * cs).sum(0) / M
        I = np.zeros((3, 3))
        for b, mb, cb in zip(members[o], ms, cs):
            R = rel[b].mat @ G.quat_to_mat(G.quat_normalize(m.body_iquat[b]))
            d = cb - com
            I += R @ np.diag(np.asarray(m.body_inertia[b], dtype=np.float64)) @ R.T + mb * (d @ d * np.eye(3) - np.outer(d, d))
        if len(members[o]) == 1:                         # nothing fused: keep MuJoCo's own inertial frame bit for bit
            body_iquat[kidx[o]] = G.quat_normalize(m.body_iquat[o]); body_inertia[kidx[o]] = m.body_inertia[o]
        else:
            w, ax = G.principal_axes(I)
            body_iquat[kidx[o]] = G.mat_to_quat(ax); body_inertia[kidx[o]] = w
        body_mass[kidx[o]] = M; body_ipos[kidx[o]] = com

    # ---- body frames relative to the owner of the parent
    body_parent = np.full(nb, -1, np.int32); body_pos = np.zeros((nb, 3)); body_quat = np.zeros((nb, 4))
    body_leg = np.full(nb, -1, np.int32)
    for b in movable:
        fr = rel[parent[b]] @ _T(m.body_pos[b], m.body_quat[b])
        body_parent[kidx[b]] = kidx[mov_parent[b]]; body_pos[kidx[b]] = fr.pos; body_quat[kidx[b]] = fr.quat
        body_leg[kidx[b]] = leg_of(b)

    # ---- DoFs in kernel order (6 free + chain hinges) and the MuJoCo <-> kernel address maps
    jnt_qposadr, jnt_dofadr, body_jntadr, body_jntnum = (np.asarray(getattr(m, k)) for k in ("jnt_qposadr", "jnt_dofadr", "body_jntadr", "body_jntnum"))
    nv = 6 + int(sum(int(body_jntnum[b]) for b in movable)); nq = nv + 1
    dof_body = np.zeros(nv, np.int32); dof_parent = np.full(nv, -1, np.int32); dof_parent[1:6] = np.arange(5)
    dof_axis = np.zeros((nv, 3)); stiff = np.zeros(nv); damp = np.zeros(nv); arm = np.zeros(nv); sref = np.zeros(nv)
    body_dofadr = np.zeros(nb, np.int32); body_dofnum = np.zeros(nb, np.int32); body_dofnum[0] = 6
    mjdof_of_k = np.zeros(nv, np.int64); mjqpos_of_k = np.zeros(nq, np.int64)
    mjdof_of_k[:6] = jnt_dofadr[jfree] + np.arange(6); mjqpos_of_k[:7] = jnt_qposadr[jfree] + np.arange(7)
    kdof_names = []
    last = {0: 5}; k = 6
    for b in movable:
        kb = kidx[b]; body_dofadr[kb] = k; prev = last[int(body_parent[kb])]
        for j in range(body_jntadr[b], body_jntadr[b] + body_jntnum[b]):
            dof_axis[k] = m.jnt_axis[j]; dof_body[k] = kb; dof_parent[k] = prev
            stiff[k] = m.jnt_stiffness[j]; sref[k] = m.qpos_spring[jnt_qposadr[j]]
            damp[k] = m.dof_damping[jnt_dofadr[j]]; arm[k] = m.dof_armature[jnt_dofadr[j]]
            mjdof_of_k[k] = jnt_dofadr[j]; mjqpos_of_k[k + 1] = jnt_qposadr[j]
            kdof_names.append(jnt_names[j]); prev = k; k += 1
        body_dofnum[kb] = k - body_dofadr[kb]; last[kb] = k - 1
    k_of_mjdof = np.full(nv_mj, -1, np.int64); k_of_mjdof[mjdof_of_k] = np.arange(nv)
    k_of_mjqpos = np.full(nq_mj, -1, np.int64); k_of_mjqpos[mjqpos_of_k] = np.arange(nq)

    # ---- actuators: position (joint transmission, affine bias -kp q - kv qdot) then adhesion (body transmission)
    nu_mj = int(m.nu)
    trntype = np.asarray(m.actuator_trntype); trnid = np.asarray(m.actuator_trnid).reshape(nu_mj, 2)
    gainprm = np.asarray(m.actuator_gainprm).reshape(nu_mj, -1); biasprm = np.asarray(m.actuator_biasprm).reshape(nu_mj, -1)
    act_names = [_short(s) for s in _names(m, "actuator", nu_mj)]
    pos_ids = [a for a in range(nu_mj) if trntype[a] == TRN_JOINT]
    adh_ids = [a for a in range(nu_mj) if trntype[a] == TRN_BODY]
    if len(pos_ids) + len(adh_ids) != nu_mj:
        raise ConversionError("only joint (position) and body (adhesion) transmissions are supported")
    act_dof, act_kp, act_kv, act_frc = [], [], [], []
    for a in pos_ids:
        j = int(trnid[a, 0])
        if jnt_type[j] != JNT_HINGE:
            raise ConversionError("actuators on the free joint are not supported")
        kp = float(gainprm[a, 0])
        if abs(biasprm[a, 0]) > 0 or abs(biasprm[a, 1] + kp) > 1e-12 * max(1.0, kp):
            raise ConversionError(f"actuator {act_names[a]} is not a position actuator (bias must be (0, -kp, -kv))")
        if k_of_mjdof[jnt_dofadr[j]] < 0:
            raise ConversionError(f"actuator {act_names[a]} drives a joint that is not part of the fly's articulated tree")
        act_dof.append(int(k_of_mjdof[jnt_dofadr[j]])); act_kp.append(kp); act_kv.append(-float(biasprm[a, 2]))
        fr = np.asarray(m.actuator_forcerange).reshape(nu_mj, 2)[a]
        limited = bool(np.asarray(getattr(m, "actuator_forcelimited", np.ones(nu_mj)))[a]) and fr[0] < fr[1]
        act_frc.append(fr if limited else (-1e30, 1e30))
    adh_body, adh_gain, adh_ctrl = [], [], []
    for a in adh_ids:
        b = int(trnid[a, 0])
        if not in_fly[b] or owner[b] == hub_b:
            raise ConversionError("adhesion actuators must sit on a jointed body, not on the free hub")
        adh_body.append(kidx[int(owner[b])]); adh_gain.append(float(gainprm[a, 0]))
        adh_ctrl.append(np.asarray(m.actuator_ctrlrange).reshape(nu_mj, 2)[a])
    k_of_mjact = np.full(nu_mj, -1, np.int64)
    k_of_mjact[pos_ids] = np.arange(len(pos_ids)); k_of_mjact[adh_ids] = len(pos_ids) + np.arange(len(adh_ids))

    # ---- contact pairs: fly geom vs ground plane (world.py:292-309)
    geom_type_mj = np.asarray(m.geom_type); geom_bodyid = np.asarray(m.geom_bodyid)
    npair = int(getattr(m, "npair", 0))
    geom_names = [_short(s) for s in _names(m, "geom", int(m.ngeom))]
    g_body, g_type, g_pos, g_quat, g_size, g_vadr, g_vnum, hull, nbr_adr, nbr, cnames = [], [], [], [], [], [], [], [], [0], [], []
    params = None; nvert = 0
    for p in range(npair):
        g1, g2 = int(m.pair_geom1[p]), int(m.pair_geom2[p])
        if geom_type_mj[g1] == GEOM_PLANE: g1, g2 = g2, g1
        if geom_type_mj[g2] != GEOM_PLANE or not in_fly[geom_bodyid[g1]]:
            raise ConversionError("only fly-geom vs ground-plane contact pairs are supported (world.py:292-309)")
        if abs(float(np.asarray(m.geom_pos)[g2][2])) > 1e-12 or np.abs(G.quat_to_mat(G.quat_normalize(m.geom_quat[g2]))[:, 2] - [0, 0, 1]).max() > 1e-12:
            raise ConversionError("the ground plane must be z = 0 with normal +z")
        pp = (float(m.pair_friction[p][0]), *map(float, m.pair_solref[p][:2]), *map(float, m.pair_solimp[p][:5]), float(m.pair_margin[p]), float(m.pair_gap[p]))
        if params is None: params = pp
        elif not np.allclose(pp, params, rtol=1e-12, atol=0):
            raise ConversionError("contact pairs with different parameters are not supported (one ContactParams per world)")
        b = int(geom_bodyid[g1]); fr = rel[b] @ _T(m.geom_pos[g1], m.geom_quat[g1])
        g_body.append(kidx[int(owner[b])]); g_pos.append(fr.pos); g_quat.append(fr.quat); cnames.append(geom_names[g1])
        if geom_type_mj[g1] == GEOM_MJ_CAPSULE:
            g_type.append(GEOM_CAPSULE); g_size.append((float(m.geom_size[g1][0]), float(m.geom_size[g1][1])))
            g_vadr.append(nvert); g_vnum.append(0)
        elif geom_type_mj[g1] == GEOM_MESH:
            verts, adj = _mesh_hull(m, int(m.geom_dataid[g1]))
            v = np.array([fr.apply(x) for x in verts])
            g_type.append(GEOM_HULL); g_size.append((0.0, 0.0)); g_vadr.append(nvert); g_vnum.append(len(v)); hull.append(v)
            for lst in adj:
                nbr.extend(lst); nbr_adr.append(len(nbr))
            nvert += len(v)
        else:
            raise ConversionError(f"geom {geom_names[g1]}: only capsule and mesh geoms collide with the ground")
    if params is None:
        params = (1.0, 2e-4, 1.0, 0.98, 0.99, 0.5, 3.0, 2.0, 1e-3, 0.0)     # no pairs (tethered): unused

    # ---- segments and sites
    segments = segments or A.bodysegs_order()
    seg_rows = [(s, body_names.index(s)) for s in segments if s in body_names and in_fly[body_names.index(s)]]
    seg_body = np.array([kidx[int(owner[b])] for _, b in seg_rows], np.int32)
    seg_pos = np.array([rel[b].pos for _, b in seg_rows]).reshape(-1, 3); seg_quat = np.array([rel[b].quat for _, b in seg_rows]).reshape(-1, 4)
    k_of_mjbody = np.full(nbody, -1, np.int64)
    for i, (_, b) in enumerate(seg_rows): k_of_mjbody[b] = i
    nsite = int(getattr(m, "nsite", 0))
    site_names = [_short(s) for s in _names(m, "site", nsite)] if nsite else []
    site_rows = [s for s in range(nsite) if in_fly[int(m.site_bodyid[s])]]
    site_body = np.array([kidx[int(owner[int(m.site_bodyid[s])])] for s in site_rows], np.int32)
    site_pos = np.array([rel[int(m.site_bodyid[s])].apply(m.site_pos[s]) for s in site_rows]).reshape(-1, 3)

    # ---- leg sensors: subtree of the most proximal contact segment of each leg (world.py:311-331)
    leg_root = np.full(6, -1, np.int32)
    for kb in sorted(set(g_body)):
        l = int(body_leg[kb])
        if l >= 0 and leg_root[l] < 0: leg_root[l] = kb

    # ---- keyframe (world.py:151-207), options, weld
    key_qpos = np.zeros(nq); key_ctrl = np.zeros(nu_mj)
    if int(getattr(m, "nkey", 0)) > 0:
        kq = np.asarray(m.key_qpos).reshape(-1, nq_mj)[0]; kc = np.asarray(m.key_ctrl).reshape(-1, nu_mj)[0]
        key_qpos = kq[mjqpos_of_k]; key_ctrl[k_of_mjact] = kc
    else:
        key_qpos = np.asarray(m.qpos0)[mjqpos_of_k]
    body_pos[0], body_quat[0] = key_qpos[:3], G.quat_normalize(key_qpos[3:7])
    o = m.opt
    opt = dict(timestep=float(o.timestep), gx=float(o.gravity[0]), gy=float(o.gravity[1]), gz=float(o.gravity[2]), iterations=float(o.iterations),
               tolerance=float(o.tolerance), ls_iterations=float(o.ls_iterations), ls_tolerance=float(o.ls_tolerance),
               noslip_iterations=float(o.noslip_iterations), meaninertia=float(m.stat.meaninertia), impratio=float(o.impratio),
               multiccd=1.0 if int(getattr(o, "enableflags", 0)) & ENBL_MULTICCD else 0.0)
    invw = np.asarray(m.body_invweight0).reshape(nbody, 2)[kbodies]
    arrays = dict(
        dims=np.array([nb, nq, nv, len(pos_ids), len(adh_ids), len(g_body), len(site_rows), len(seg_rows), 6, nvert], np.int32),
        opt=np.array([opt[k] for k in OPT_FIELDS]), contact=np.array(params, dtype=np.float64),
        body_parent=body_parent, body_pos=body_pos, body_quat=body_quat, body_mass=body_mass, body_ipos=body_ipos, body_iquat=body_iquat,
        body_inertia=body_inertia, body_invweight0=invw, body_dofadr=body_dofadr, body_dofnum=body_dofnum, body_leg=body_leg,
        dof_body=dof_body, dof_parent=dof_parent, dof_axis=dof_axis, dof_stiffness=stiff, dof_damping=damp, dof_armature=arm, dof_springref=sref,
        act_dof=np.array(act_dof, np.int32), act_kp=np.array(act_kp), act_kv=np.array(act_kv), act_frcrange=np.array(act_frc, dtype=np.float64).reshape(-1, 2),
        adh_body=np.array(adh_body, np.int32), adh_gain=np.array(adh_gain), adh_ctrlrange=np.array(adh_ctrl, dtype=np.float64).reshape(-1, 2),
        geom_body=np.array(g_body, np.int32), geom_type=np.array(g_type, np.int32), geom_pos=np.array(g_pos).reshape(-1, 3),
        geom_quat=np.array(g_quat).reshape(-1, 4), geom_size=np.array(g_size, dtype=np.float64).reshape(-1, 2),
        geom_vertadr=np.array(g_vadr, np.int32), geom_vertnum=np.array(g_vnum, np.int32),
        hull_vert=np.concatenate(hull) if nvert else np.zeros((0, 3)), hull_nbr_adr=np.array(nbr_adr, np.int32), hull_nbr=np.array(nbr if nbr else [0], np.int32),
        site_body=site_body, site_pos=site_pos, seg_body=seg_body, seg_pos=seg_pos, seg_quat=seg_quat, leg_rootbody=leg_root,
        key_qpos=key_qpos, key_ctrl=key_ctrl,
    )
    neq = int(getattr(m, "neq", 0))
    if neq:
        if neq != 1 or int(m.eq_type[0]) != EQ_WELD:
            raise ConversionError("only the TetheredWorld weld equality is supported")
        b1, b2 = int(m.eq_obj1id[0]), int(m.eq_obj2id[0])
        if b2 != 0 or not in_fly[b1] or owner[b1] != hub_b:
            raise ConversionError("the weld must tie a body of the free hub to the world (world.py:350-366)")
        d = np.asarray(m.eq_data).reshape(neq, -1)[0]       # anchor2(3) relpos... : [PRIOR] mjEQ_WELD data = anchor(3), relpose pos(3), quat(4), torquescale
        relpos, relquat, ts = d[3:6], d[6:10], float(d[10])
        anchor = rel[b1].apply(relpos); quat = G.quat_mul(rel[b1].quat, G.quat_normalize(relquat))
        arrays["weld"] = np.array([1.0, *anchor, *quat, *map(float, m.eq_solref[0][:2]), *map(float, m.eq_solimp[0][:5]), ts, invw[0, 0], invw[0, 1]])
        assert len(arrays["weld"]) == len(WELD_FIELDS)
    legs = [l for l in leg_names if any(leg_of(b) == leg_names.index(l) for b in movable)]
    if legs != leg_names[:len(legs)] or len(legs) != 6:
        raise ConversionError(f"expected the six legs {leg_names}; found {legs}")
    names = dict(bodies=["hub"] + [body_names[b] for b in movable], segments=[s for s, _ in seg_rows], jointdofs=kdof_names,
                 actuated_position=[jnt_names[int(trnid[a, 0])] for a in pos_ids], legs=legs, contact_geoms=cnames,
                 sites=[site_names[s] for s in site_rows])
    meta = dict(source="MjModel", units="as the MuJoCo model", simplify_geom=bool(nvert == 0),
                mj_maps=dict(qpos=k_of_mjqpos.tolist(), dof=k_of_mjdof.tolist(), actuator=k_of_mjact.tolist(), body_to_segment=k_of_mjbody.tolist()))
    assert len(arrays["contact"]) == len(CONTACT_FIELDS) and len(arrays["dims"]) == len(DIM_FIELDS)
    return NMFModel(arrays, names, meta)


def _mesh_hull(m, mesh_id: int):
    """Convex-hull vertices (mesh frame) and their adjacency lists from MuJoCo's ``mesh_graph``:
    ``[numvert, numface, vert_edgeadr[numvert], vert_globalid[numvert], edge_localid[numvert + 3 numface], face_globalid[3 numface]]``,
    the neighbour list of hull vertex i starts at ``edge_localid[vert_edgeadr[i]]`` and ends with -1."""
    gadr = int(m.mesh_graphadr[mesh_id])
    if gadr < 0:
        raise ConversionError("mesh geoms need their convex hull (mesh_graph); compile with convexhull enabled")
    g = np.asarray(m.mesh_graph)
    nvert = int(g[gadr]); ea = g[gadr + 2: gadr + 2 + nvert]; gid = g[gadr + 2 + nvert: gadr + 2 + 2 * nvert]
    el = g[gadr + 2 + 2 * nvert:]
    vadr = int(m.mesh_vertadr[mesh_id]); mv = np.asarray(m.mesh_vert).reshape(-1, 3)
    verts = mv[vadr + gid]
    adj = []
    for i in range(nvert):
        lst = []; e = int(ea[i])
        while el[e] >= 0:
            lst.append(int(el[e])); e += 1
        adj.append(lst)
    return verts, adj


# ----------------------------------------------------------------------------------------------------------------------
def mjmodel_like(model: NMFModel, *, prefix: str = "nmf/", unfused: bool = True, shuffle_actuators: bool = True) -> SimpleNamespace:
    """A duck-typed ``MjModel`` of the world a baked model stands for, as the reference composes it: world body, the dm_control
    attachment body with the free joint (world.py:276-279), and -- with ``unfused`` -- one body per anatomical segment with the
    jointless ones (head, eyes, antennae, abdomen, wings, halteres ...) as static children carrying their own geoms, so that
    :func:`from_mjmodel` has real fusing and re-indexing to do.  Actuators are declared adhesion-first when ``shuffle_actuators``
    so that MuJoCo actuator ids differ from the record's ctrl layout.  TEST / DOCUMENTATION helper (MuJoCo itself is absent)."""
    a = model.arrays
    nbk = model.nbody; segs = model.names["segments"]; seg_body = a["seg_body"]
    seg_pos = a["seg_pos"].reshape(-1, 3); seg_quat = a["seg_quat"].reshape(-1, 4)
    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), mass=0.0, ipos=np.zeros(3), iquat=np.array([1.0, 0, 0, 0]), inertia=np.zeros(3), joints=[])]
    kb_to_mj = {}

    def add_body(name, parent, pos, quat, mass, ipos, iquat, inertia):
        bodies.append(dict(name=prefix + name, parent=parent, pos=np.asarray(pos, float), quat=np.asarray(quat, float), mass=float(mass),
                           ipos=np.asarray(ipos, float), iquat=np.asarray(iquat, float), inertia=np.asarray(inertia, float), joints=[]))
        return len(bodies) - 1

    # kernel body k is named after the segment that coincides with its frame (identity offset); the hub is the attachment body
    frame_seg = {}
    for s, (kb, p, q) in enumerate(zip(seg_body, seg_pos, seg_quat)):
        if kb > 0 and np.abs(p).max() == 0 and abs(abs(q[0]) - 1) < 1e-15: frame_seg[int(kb)] = segs[s]
    static_of = {kb: [s for s in range(len(segs)) if seg_body[s] == kb and segs[s] != frame_seg.get(kb)] for kb in range(nbk)}
    order = [0]
    children = {kb: [c for c in range(1, nbk) if a["body_parent"][c] == kb] for kb in range(nbk)}
    seg_mj = {}

    def visit(kb, mj_parent):
        name = "" if kb == 0 else frame_seg[kb]
        pos, quat = (a["key_qpos"][:3], a["key_qpos"][3:7]) if kb == 0 else (a["body_pos"][kb], a["body_quat"][kb])
        b = add_body(name, mj_parent, pos, quat, a["body_mass"][kb], a["body_ipos"][kb], a["body_iquat"][kb], a["body_inertia"][kb])
        kb_to_mj[kb] = b
        if kb > 0: seg_mj[frame_seg[kb]] = b
        if unfused:
            for s in static_of[kb]:                      # massless static children (MuJoCo would carry their share of the mass; the sum is what matters)
                seg_mj[segs[s]] = add_body(segs[s], b, seg_pos[s], seg_quat[s], 0.0, np.zeros(3), [1.0, 0, 0, 0], np.zeros(3))
        for c in children[kb]:
            visit(c, b)

    visit(0, 0)
    nbody = len(bodies)
    # joints / dofs in MuJoCo (depth-first body) order
    jnt = []; qadr = 0; dadr = 0
    mj_of_kdof = {}
    for b, bd in enumerate(bodies):
        kb = next((k for k, v in kb_to_mj.items() if v == b), None)
        bd["jntadr"] = len(jnt) if kb is not None and (kb == 0 or a["body_dofnum"][kb] > 0) else -1
        bd["dofadr"] = dadr if bd["jntadr"] >= 0 else -1
        if kb == 0:
            jnt.append(dict(name=prefix, type=JNT_FREE, body=b, axis=[0, 0, 1], qposadr=qadr, dofadr=dadr, stiffness=0.0, kdof=0)); qadr += 7; dadr += 6
        elif kb is not None:
            for j in range(int(a["body_dofnum"][kb])):
                d = int(a["body_dofadr"][kb]) + j
                jnt.append(dict(name=prefix + model.names.get("all_jointdofs", _all_dof_names(model))[d - 6], type=JNT_HINGE, body=b, axis=a["dof_axis"].reshape(-1, 3)[d], qposadr=qadr, dofadr=dadr,
                                stiffness=a["dof_stiffness"][d], kdof=d)); mj_of_kdof[d] = dadr; qadr += 1; dadr += 1
        bd["jntnum"] = len(jnt) - bd["jntadr"] if bd["jntadr"] >= 0 else 0
        bd["dofnum"] = dadr - bd["dofadr"] if bd["dofadr"] >= 0 else 0
    nq, nv = qadr, dadr
    m = SimpleNamespace()
    m.nbody, m.njnt, m.nq, m.nv = nbody, len(jnt), nq, nv
    m.body_parentid = np.array([b["parent"] for b in bodies]); m.body_pos = np.array([b["pos"] for b in bodies]); m.body_quat = np.array([b["quat"] for b in bodies])
    m.body_mass = np.array([b["mass"] for b in bodies]); m.body_ipos = np.array([b["ipos"] for b in bodies]); m.body_iquat = np.array([b["iquat"] for b in bodies])
    m.body_inertia = np.array([b["inertia"] for b in bodies])
    m.body_jntadr = np.array([b["jntadr"] for b in bodies]); m.body_jntnum = np.array([b["jntnum"] for b in bodies])
    m.body_dofadr = np.array([b["dofadr"] for b in bodies]); m.body_dofnum = np.array([b["dofnum"] for b in bodies])
    invw = np.zeros((nbody, 2))
    for kb, b in kb_to_mj.items(): invw[b] = a["body_invweight0"].reshape(-1, 2)[kb]
    m.body_invweight0 = invw
    m.jnt_type = np.array([j["type"] for j in jnt]); m.jnt_bodyid = np.array([j["body"] for j in jnt]); m.jnt_axis = np.array([j["axis"] for j in jnt], dtype=float)
    m.jnt_pos = np.zeros((len(jnt), 3)); m.jnt_qposadr = np.array([j["qposadr"] for j in jnt]); m.jnt_dofadr = np.array([j["dofadr"] for j in jnt])
    m.jnt_stiffness = np.array([j["stiffness"] for j in jnt], dtype=float)
    m.qpos_spring = np.zeros(nq); m.qpos0 = np.zeros(nq); m.dof_damping = np.zeros(nv); m.dof_armature = np.zeros(nv)
    key_qpos = np.zeros(nq)
    for j in jnt:
        if j["type"] == JNT_FREE:
            key_qpos[j["qposadr"]:j["qposadr"] + 7] = a["key_qpos"][:7]; m.qpos0[j["qposadr"]:j["qposadr"] + 7] = a["key_qpos"][:7]
        else:
            d = j["kdof"]; m.qpos_spring[j["qposadr"]] = a["dof_springref"][d]; key_qpos[j["qposadr"]] = a["key_qpos"][d + 1]
            m.dof_damping[j["dofadr"]] = a["dof_damping"][d]; m.dof_armature[j["dofadr"]] = a["dof_armature"][d]
    # actuators: adhesion first when shuffled
    nup, nua = model.dim("nu_pos"), model.dim("nu_adh")
    acts = [("pos", i) for i in range(nup)] + [("adh", i) for i in range(nua)]
    if shuffle_actuators: acts = acts[nup:] + acts[:nup]
    nu = len(acts)
    m.nu = nu; m.actuator_trntype = np.zeros(nu, int); m.actuator_trnid = np.full((nu, 2), -1); m.actuator_gainprm = np.zeros((nu, 10)); m.actuator_biasprm = np.zeros((nu, 10))
    m.actuator_forcerange = np.zeros((nu, 2)); m.actuator_forcelimited = np.zeros(nu, int); m.actuator_ctrlrange = np.zeros((nu, 2)); key_ctrl = np.zeros(nu); anames = []
    jnt_of_kdof = {j["kdof"]: i for i, j in enumerate(jnt) if j["type"] == JNT_HINGE}
    for i, (kind, k) in enumerate(acts):
        if kind == "pos":
            d = int(a["act_dof"][k]); m.actuator_trntype[i] = TRN_JOINT; m.actuator_trnid[i, 0] = jnt_of_kdof[d]
            m.actuator_gainprm[i, 0] = a["act_kp"][k]; m.actuator_biasprm[i, 1] = -a["act_kp"][k]; m.actuator_biasprm[i, 2] = -a["act_kv"][k]
            m.actuator_forcerange[i] = a["act_frcrange"].reshape(-1, 2)[k]; m.actuator_forcelimited[i] = 1; key_ctrl[i] = a["key_ctrl"][k]
            anames.append(prefix + jnt[jnt_of_kdof[d]]["name"].split("/")[-1] + "-position")
        else:
            kb = int(a["adh_body"][k]); m.actuator_trntype[i] = TRN_BODY; m.actuator_trnid[i, 0] = kb_to_mj[kb]
            m.actuator_gainprm[i, 0] = a["adh_gain"][k]; m.actuator_ctrlrange[i] = a["adh_ctrlrange"].reshape(-1, 2)[k]; key_ctrl[i] = a["key_ctrl"][nup + k]
            anames.append(prefix + bodies[kb_to_mj[kb]]["name"].split("/")[-1] + "-adhesion")
    # geoms: ground plane + one per contact pair, attached to their own segment body when unfused
    ng = model.dim("ngeom"); cg = model.names["contact_geoms"]
    gtype = [GEOM_PLANE]; gbody = [0]; gpos = [np.zeros(3)]; gquat = [np.array([1.0, 0, 0, 0])]; gsize = [np.array([1000.0, 1000.0, 1.0])]; gdata = [-1]; gnames = ["ground_plane"]
    mesh_vert, mesh_vertadr, mesh_vertnum, mesh_graphadr, mesh_graph = [], [], [], [], []
    hv = a["hull_vert"].reshape(-1, 3); nadr = a["hull_nbr_adr"]; nb_ = a["hull_nbr"]
    for g in range(ng):
        kb = int(a["geom_body"][g]); seg = cg[g]
        fr_owner = _T(a["geom_pos"].reshape(-1, 3)[g], a["geom_quat"].reshape(-1, 4)[g])
        if unfused and seg in seg_mj and seg_mj[seg] != kb_to_mj[kb]:
            s = segs.index(seg); Ts = _T(seg_pos[s], seg_quat[s])
            inv = _T(-(Ts.mat.T @ Ts.pos), G.quat_conj(Ts.quat)); fr = inv @ fr_owner; body = seg_mj[seg]
        else:
            inv = _T(); fr = fr_owner; body = kb_to_mj[kb]
        gbody.append(body); gpos.append(fr.pos); gquat.append(fr.quat); gnames.append(prefix + seg)
        if int(a["geom_type"][g]) == GEOM_CAPSULE:
            gtype.append(GEOM_MJ_CAPSULE); gsize.append(np.array([*a["geom_size"].reshape(-1, 2)[g], 0.0])); gdata.append(-1)
        else:
            adr, num = int(a["geom_vertadr"][g]), int(a["geom_vertnum"][g])
            # mesh vertices are stored in the geom's own frame: v_mesh = fr^-1 (v_owner)
            Fo = fr_owner; vm = np.array([Fo.mat.T @ (x - Fo.pos) for x in hv[adr:adr + num]])
            pad = np.zeros((2, 3))                       # two non-hull vertices in front, so that vert_globalid is not the identity
            mesh_vertadr.append(sum(len(x) for x in mesh_vert)); mesh_vert.append(np.concatenate([pad, vm])); mesh_vertnum.append(num + 2)
            lists = [[int(x) for x in nb_[nadr[adr + i]:nadr[adr + i + 1]]] for i in range(num)]
            ea, el = [], []
            for lst in lists:
                ea.append(len(el)); el.extend(lst + [-1])
            graph = [num, 0] + ea + [i + 2 for i in range(num)] + el
            mesh_graphadr.append(sum(len(x) for x in mesh_graph)); mesh_graph.append(graph)
            gtype.append(GEOM_MESH); gsize.append(np.zeros(3)); gdata.append(len(mesh_vertadr) - 1)
    m.ngeom = len(gtype); m.geom_type = np.array(gtype); m.geom_bodyid = np.array(gbody); m.geom_pos = np.array(gpos); m.geom_quat = np.array(gquat)
    m.geom_size = np.array(gsize); m.geom_dataid = np.array(gdata)
    m.mesh_vert = np.concatenate(mesh_vert) if mesh_vert else np.zeros((0, 3)); m.mesh_vertadr = np.array(mesh_vertadr, int); m.mesh_vertnum = np.array(mesh_vertnum, int)
    m.mesh_graphadr = np.array(mesh_graphadr, int); m.mesh_graph = np.array([x for gph in mesh_graph for x in gph], int)
    c = a["contact"]
    m.npair = ng; m.pair_geom1 = np.arange(1, ng + 1); m.pair_geom2 = np.zeros(ng, int)       # (fly geom, ground) as world.py:300-303 declares them
    m.pair_friction = np.tile([c[0], c[0], 0.02, 1e-4, 1e-4], (ng, 1)); m.pair_solref = np.tile(c[1:3], (ng, 1)); m.pair_solimp = np.tile(c[3:8], (ng, 1))
    m.pair_margin = np.full(ng, c[8]); m.pair_gap = np.full(ng, c[9])
    # sites
    sb = a["site_body"]; sp = a["site_pos"].reshape(-1, 3)
    m.nsite = len(sb); m.site_bodyid = np.array([kb_to_mj[int(k)] for k in sb]); m.site_pos = sp.copy()
    m.nkey = 1; m.key_qpos = key_qpos[None]; m.key_ctrl = key_ctrl[None]
    o = model._opt_full()
    m.opt = SimpleNamespace(timestep=o[0], gravity=np.array(o[1:4]), iterations=int(o[4]), tolerance=o[5], ls_iterations=int(o[6]), ls_tolerance=o[7],
                            noslip_iterations=int(o[8]), impratio=o[10], enableflags=(ENBL_MULTICCD if o[11] else 0) | 2)
    m.stat = SimpleNamespace(meaninertia=o[9])
    m.neq = 0
    w = a.get("weld")
    if w is not None and w[0] != 0:
        m.neq = 1; m.eq_type = np.array([EQ_WELD]); m.eq_obj1id = np.array([kb_to_mj[0]]); m.eq_obj2id = np.array([0])
        m.eq_data = np.array([[0, 0, 0, *w[1:4], *w[4:8], w[15]]]); m.eq_solref = np.array([w[8:10]]); m.eq_solimp = np.array([w[10:15]])
    m.names = dict(body=[b["name"] for b in bodies], joint=[j["name"] for j in jnt], actuator=anames, geom=gnames,
                   site=[prefix + s for s in model.names["sites"]])
    return m


def _all_dof_names(model: NMFModel) -> list[str]:
    """hinge-DoF names of the kernel layout (66), locked ones included, in layout order"""
    ex = list(model.exposed_hinge_dofs()); names = [None] * (model.nv - 6)
    for j, nm in zip(ex, model.names["jointdofs"]): names[int(j)] = nm
    locked = iter(model.names.get("locked_jointdofs", []))
    return [n if n is not None else next(locked) for n in names]
