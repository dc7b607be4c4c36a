"""Flat, compiled description of one NeuroMechFly world ("baked model").

This is what the reference gets from ``world.compile()`` (MuJoCo ``MjModel``,
reference ``src/flygym/compose/base.py:21-27``) reduced to the arrays the step
path reads.  It is produced offline by :mod:`flygym_b200.baker.bake` from the
reference's assets, stored as ``.npz`` under ``flygym_b200/assets`` and handed to
the native libraries as one self-describing binary blob (``to_blob``).

Blob layout (little-endian), parsed identically by ``flygym_b200/csrc`` and by
``oracle/nmf_oracle.c``::

    char   magic[8] = "NMFB200\\0"
    int32  version, nsections
    repeat nsections:  char name[24]; int32 dtype(0=f64,1=i32); int32 count; int64 offset
    payload (8-byte aligned)
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

BLOB_MAGIC = b"NMFB200\0"
BLOB_VERSION = 1
ASSETS_DIR = Path(__file__).resolve().parent / "assets"

# name -> dtype ; the order is the blob section order
_F64 = [
    "body_pos", "body_quat", "body_mass", "body_ipos", "body_iquat", "body_inertia",
    "body_invweight0",
    "dof_axis", "dof_stiffness", "dof_damping", "dof_armature", "dof_springref",
    "act_kp", "act_kv", "act_frcrange",
    "adh_gain", "adh_ctrlrange",
    "geom_pos", "geom_quat", "geom_size", "hull_vert",
    "site_pos", "seg_pos", "seg_quat",
    "key_qpos", "key_ctrl",
    "opt",  # see OPT_FIELDS
    "contact",  # see CONTACT_FIELDS
    "terrain",  # see TERRAIN_FIELDS (optional: all zeros = the reference's flat ground plane)
    "weld",  # see WELD_FIELDS (optional: all zeros = no weld; TetheredWorld sets it)
]
_I32 = [
    "dims",  # see DIM_FIELDS
    "body_parent", "body_dofadr", "body_dofnum", "body_leg",
    "dof_body", "dof_parent",
    "act_dof", "adh_body",
    "geom_body", "geom_type", "geom_vertadr", "geom_vertnum",
    "site_body", "seg_body", "leg_rootbody", "hull_nbr_adr", "hull_nbr",
]
DIM_FIELDS = ["nbody", "nq", "nv", "nu_pos", "nu_adh", "ngeom", "nsite", "nseg", "nleg", "nhullvert"]
OPT_FIELDS = ["timestep", "gx", "gy", "gz", "iterations", "tolerance", "ls_iterations",
              "ls_tolerance", "noslip_iterations", "meaninertia", "impratio", "multiccd"]
CONTACT_FIELDS = ["mu", "solref0", "solref1", "solimp0", "solimp1", "solimp2", "solimp3",
                  "solimp4", "margin", "gap"]
OPT_DEFAULTS = {"multiccd": 1.0}
TERRAIN_FIELDS = ["type", "period_x", "period_y", "half_x", "half_y", "top_even", "top_odd", "z_floor"]
# Terrain worlds.  FlyGym 2.0.1 ships only FlatGroundWorld / TetheredWorld (reference src/flygym/compose/world.py:229-366);
# BASELINE.json config 3 asks for the v1-style "blocks" and "gapped" arenas, defined here as a floor plane plus a grid of
# axis-aligned box columns ([PRIOR] v1 defaults): column (i, j) covers |x - i*period_x| <= half_x,
# |y - j*period_y| <= half_y, z <= top_even / top_odd by the parity of i + j.
TERRAINS = {
    # 1.0 mm blocks separated by 0.4 mm gaps that are 2 mm deep, running across the walking direction (infinite in y)
    "gapped": [1.0, 1.4, 1.0e6, 0.5, 0.5e6, 0.0, 0.0, -2.0],
    # 1.3 mm square tiles, every other one raised by 0.2 mm (checkerboard)
    "blocks": [1.0, 1.3, 1.3, 0.65, 0.65, 0.0, 0.2, -1.0],
}
# TetheredWorld's weld(body1=c_thorax, body2=world) (reference src/flygym/compose/world.py:350-366), reduced to what the
# step needs: the body-1 anchor point and the orientation offset in the frame of the free ("hub") body, the constraint's
# own solref / solimp, torquescale, and the hub's translational / rotational inverse weights (diagApprox of the six rows).
WELD_FIELDS = ["enabled", "anchor_x", "anchor_y", "anchor_z", "quat_w", "quat_x", "quat_y", "quat_z", "solref0", "solref1",
               "solimp0", "solimp1", "solimp2", "solimp3", "solimp4", "torquescale", "invweight_tran", "invweight_rot"]
GEOM_CAPSULE = 0
GEOM_HULL = 1


def _quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def _quat_rotate(q, v):
    t = 2.0 * np.cross(q[1:], v)
    return v + q[0] * t + np.cross(q[1:], t)


@dataclass
class NMFModel:
    arrays: dict[str, np.ndarray]
    names: dict[str, list[str]] = field(default_factory=dict)
    meta: dict = field(default_factory=dict)

    # ---- convenience ----------------------------------------------------
    def dim(self, key: str) -> int:
        return int(self.arrays["dims"][DIM_FIELDS.index(key)])

    def opt(self, key: str) -> float:
        return float(self._opt_full()[OPT_FIELDS.index(key)])

    def _opt_full(self) -> np.ndarray:
        """``opt`` padded to the current OPT_FIELDS (models baked before a field existed get the reference's setting:
        ``multiccd`` is enabled in ``mujoco_globals.yaml:18``)."""
        o = np.asarray(self.arrays["opt"], dtype=np.float64)
        if len(o) < len(OPT_FIELDS):
            o = np.r_[o, [OPT_DEFAULTS[k] for k in OPT_FIELDS[len(o):]]]
        return o

    def with_options(self, **kw) -> "NMFModel":
        """Copy with solver / collision options changed: ``noslip_iterations`` (5 = the reference's CPU ``Simulation``,
        0 = its ``GPUSimulation``, which strips noslip: ``warp/simulation.py:427-448``), ``multiccd`` (0 / 1), ``iterations`` ..."""
        o = self._opt_full().copy()
        for k, v in kw.items():
            o[OPT_FIELDS.index(k)] = float(v)
        return NMFModel(dict(self.arrays, opt=o), self.names, dict(self.meta))

    @property
    def nq(self): return self.dim("nq")
    @property
    def nv(self): return self.dim("nv")
    @property
    def nu(self): return self.dim("nu_pos") + self.dim("nu_adh")
    @property
    def nbody(self): return self.dim("nbody")
    @property
    def timestep(self): return self.opt("timestep")

    # ---- persistence ----------------------------------------------------
    def save(self, path) -> None:
        payload = {k: v for k, v in self.arrays.items()}
        payload["__names__"] = np.frombuffer(json.dumps(self.names).encode(), dtype=np.uint8)
        payload["__meta__"] = np.frombuffer(json.dumps(self.meta).encode(), dtype=np.uint8)
        np.savez_compressed(path, **payload)

    @classmethod
    def load(cls, path) -> "NMFModel":
        path = Path(path)
        if not path.exists() and not path.is_absolute():
            path = ASSETS_DIR / path
        with np.load(path) as z:
            arrays = {k: z[k] for k in z.files if not k.startswith("__")}
            names = json.loads(bytes(z["__names__"]).decode())
            meta = json.loads(bytes(z["__meta__"]).decode())
        return cls(arrays, names, meta)

    @classmethod
    def bench(cls, simplify_geom: bool = True, terrain: str | None = None, joint_preset: str = "legs_only",
              contact_preset: str | None = None) -> "NMFModel":
        """The reference benchmark model (``time_gpu_simulation.py:21-64``); ``terrain`` = ``"blocks"`` / ``"gapped"``
        replaces the flat ground plane by a box-column terrain (capsule geoms only); ``joint_preset="legs_active_only"``
        (reference ``anatomy.py:402-409``: no passive tarsal joints, 42 hinge DoFs) is baked in the kernels' chain layout with
        the tarsus2-5 links massless and their DoFs locked (``baker.bake.bake_kernel_layout``).

        ``joint_preset="all_biological"`` (126 hinge DoFs: head, proboscis, antennae, eyes, abdomen, wings and halteres
        articulated as well, reference ``anatomy.py:418-436``) and ``"all_possible"`` (204: three DoFs at every anatomical joint,
        ``anatomy.py:411-416``) are general trees, stepped by the general-topology kernels (``csrc/nmf_tree.cuh``); so is
        ``contact_preset="all"`` (``ContactBodiesPreset.ALL``, ``anatomy.py:519-526``: all 69 segments collide with the
        ground) on any skeleton.  ``contact_preset`` defaults to the reference's ``legs_thorax_abdomen_head``."""
        if joint_preset == "legs_active_only":
            if not simplify_geom:
                raise ValueError("the LEGS_ACTIVE_ONLY preset is baked with capsule geoms only")
            if contact_preset == "all":
                raise ValueError("ContactBodiesPreset.ALL is baked for the LEGS_ONLY / ALL_BIOLOGICAL / ALL_POSSIBLE skeletons")
            m = cls.load(ASSETS_DIR / "nmf_bench_capsule_legs_active_only.npz")    # flygym_b200.baker.bake.bake_kernel_layout
        elif joint_preset == "legs_only" and contact_preset != "all":
            m = cls.load(ASSETS_DIR / ("nmf_bench_capsule.npz" if simplify_geom else "nmf_bench_mesh.npz"))
        else:
            files = {("legs_only", True): "nmf_legs_only_allcontacts_capsule.npz", ("all_biological", True): "nmf_all_biological_capsule.npz",
                     ("all_possible", True): "nmf_all_possible_capsule.npz", ("all_biological", False): "nmf_all_biological_mesh.npz"}
            if (joint_preset, bool(simplify_geom)) not in files:
                raise ValueError(f"no baked model for joint_preset={joint_preset!r}, simplify_geom={simplify_geom} "
                                 "(bake it with flygym_b200.baker.bake.bake where the reference's assets are available)")
            m = cls.load(ASSETS_DIR / files[(joint_preset, bool(simplify_geom))])      # baked with every segment as a contact body
            if contact_preset != "all":
                m = m.with_contact_bodies(contact_preset or "legs_thorax_abdomen_head")
            contact_preset = None
        if contact_preset not in (None, "legs_thorax_abdomen_head"):
            m = m.with_contact_bodies(contact_preset)
        return m if terrain in (None, "flat") else m.with_terrain(terrain)

    @classmethod
    def tethered(cls, spawn_position=(0.0, 0.0, 1.5), spawn_quat=(1.0, 0.0, 0.0, 0.0), joint_preset: str = "legs_only") -> "NMFModel":
        """The same fly in the reference's ``TetheredWorld`` (``world.py:334-366``, the world behind the reference's own
        ``tests/core/test_simulation.py`` fixtures): no ground, no contact pairs, and a soft weld
        ``weld(body1=c_thorax, body2=world, relpose=(*spawn_position, *spawn_rotation), solref=(2e-4, 1),
        solimp=(0.98, 0.99, 1e-5, 0.5, 3))``.  MuJoCo reads ``relpose`` as the pose of body 2 in the frame of body 1, so the
        constraint pulls the thorax-frame point ``spawn_position`` onto the world origin and ``q_thorax * spawn_quat`` onto
        the identity ([PRIOR] mjEQ_WELD semantics); that effective behaviour is what is reproduced."""
        m = cls.bench(True, joint_preset=joint_preset)      # (any skeleton: the full ones run on the general-topology kernels)
        a = dict(m.arrays)
        dims = a["dims"].copy(); dims[DIM_FIELDS.index("ngeom")] = 0; dims[DIM_FIELDS.index("nhullvert")] = 0
        a["dims"] = dims
        for k, w in (("geom_pos", 3), ("geom_quat", 4), ("geom_size", 2)):
            a[k] = np.zeros((0, w))
        for k in ("geom_body", "geom_type", "geom_vertadr", "geom_vertnum"):
            a[k] = np.zeros(0, np.int32)
        a["hull_vert"] = np.zeros((0, 3)); a["hull_nbr_adr"] = np.zeros(1, np.int32); a["hull_nbr"] = np.zeros(0, np.int32)
        key = a["key_qpos"].copy(); key[0:3] = spawn_position; key[3:7] = spawn_quat
        a["key_qpos"] = key
        seg = m.names["segments"].index("c_thorax")
        if int(a["seg_body"][seg]) != 0:
            raise ValueError("c_thorax is not part of the free body")
        sp, sq = a["seg_pos"].reshape(-1, 3)[seg], a["seg_quat"].reshape(-1, 4)[seg]
        anchor = sp + _quat_rotate(sq, np.asarray(spawn_position, dtype=np.float64))
        quat = _quat_mul(sq, np.asarray(spawn_quat, dtype=np.float64))
        invw = a["body_invweight0"].reshape(-1, 2)[0]
        a["weld"] = np.array([1.0, *anchor, *quat, 2e-4, 1.0, 0.98, 0.99, 1e-5, 0.5, 3.0, 1.0, invw[0], invw[1]])
        names = dict(m.names, contact_geoms=[])
        return NMFModel(a, names, dict(m.meta, world="tethered", contact_preset=None))

    # ---- model variants the reference composes through Fly / World keyword arguments ----------------------------------
    def with_contact_bodies(self, preset: str) -> "NMFModel":
        """Keep only the ground-contact pairs of a ``ContactBodiesPreset`` (reference ``anatomy.py:501-562``, applied in
        ``world.py:292-309``): ``"legs_thorax_abdomen_head"`` (the baked default), ``"legs_only"``, ``"tibia_tarsus_only"``.
        The per-leg contact sensor follows: its subtree starts at the most proximal remaining contact segment of the leg
        (``world.py:311-331``)."""
        from . import anatomy as A
        keep_names = set(A.contact_bodies(preset))
        geoms = self.names["contact_geoms"]
        keep = np.array([g in keep_names for g in geoms], dtype=bool)
        if keep.sum() == 0 or not keep_names.issubset(set(geoms)):
            raise ValueError(f"contact preset {preset!r} is not a subset of the baked contact geoms")
        a = dict(self.arrays)
        for k, w in (("geom_pos", 3), ("geom_quat", 4), ("geom_size", 2)):
            a[k] = a[k].reshape(-1, w)[keep]
        for k in ("geom_body", "geom_type", "geom_vertadr", "geom_vertnum"):
            a[k] = a[k][keep]            # hull vertex storage is shared and indexed by (adr, num): nothing to repack
        dims = a["dims"].copy(); dims[DIM_FIELDS.index("ngeom")] = int(keep.sum()); a["dims"] = dims
        root = np.full(self.dim("nleg"), -1, np.int32)
        body_leg = a["body_leg"]
        for b in sorted(set(int(x) for x in a["geom_body"])):
            l = int(body_leg[b])
            if l >= 0 and root[l] < 0:
                root[l] = b                   # bodies of a leg are numbered proximal -> distal
        a["leg_rootbody"] = root
        names = dict(self.names, contact_geoms=[g for g, k in zip(geoms, keep) if k])
        return NMFModel(a, names, dict(self.meta, contact_preset=preset))

    def with_actuator_gains(self, kp: float | None = None, kv: float | None = None, forcerange=None) -> "NMFModel":
        """``Fly.add_actuators(..., kp=, kv=, forcerange=)`` (reference ``fly.py:301-369``; tutorial 2 uses kp = 150)."""
        a = dict(self.arrays)
        n = self.dim("nu_pos")
        if kp is not None: a["act_kp"] = np.full(n, float(kp))
        if kv is not None: a["act_kv"] = np.full(n, float(kv))
        if forcerange is not None: a["act_frcrange"] = np.tile(np.asarray(forcerange, dtype=np.float64), (n, 1))
        return NMFModel(a, self.names, dict(self.meta, position_gain=float(a["act_kp"][0])))

    def with_actuated_dofs(self, preset: str) -> "NMFModel":
        """Position actuators on the DoFs of an ``ActuatedDOFPreset`` (reference ``anatomy.py:463-498``, applied by
        ``Fly.add_actuators``, ``fly.py:301-369``): ``"legs_active_only"`` (the baked default, 42 actuators) or ``"legs_only"`` /
        ``"all"`` (every leg hinge incl. the passive tarsal joints, 66 actuators in the LEGS_ONLY skeleton).  Gains and force range
        are those of the existing actuators; the neutral input of an actuator is its DoF's neutral (keyframe) angle."""
        from . import anatomy as A
        dofs = [A.JointDOF(*nm.split("-")) for nm in self.names["jointdofs"]]
        chosen = {d.name for d in A.actuated_dofs(dofs, preset)}
        idx = [j for j, d in enumerate(dofs) if d.name in chosen]
        if not idx:
            raise ValueError(f"actuated-DoF preset {preset!r} selects nothing in this skeleton")
        ex = self.exposed_hinge_dofs()
        a = dict(self.arrays)
        n = len(idx)
        a["act_dof"] = np.array([6 + int(ex[j]) for j in idx], dtype=np.int32)
        a["act_kp"] = np.full(n, float(self.arrays["act_kp"][0])); a["act_kv"] = np.full(n, float(self.arrays["act_kv"][0]))
        a["act_frcrange"] = np.tile(self.arrays["act_frcrange"].reshape(-1, 2)[0], (n, 1))
        a["key_ctrl"] = np.r_[self.arrays["key_qpos"][1 + a["act_dof"]], self.arrays["key_ctrl"][self.dim("nu_pos"):]]
        dims = a["dims"].copy(); dims[DIM_FIELDS.index("nu_pos")] = n; a["dims"] = dims
        names = dict(self.names, actuated_position=[dofs[j].name for j in idx])
        return NMFModel(a, names, dict(self.meta, actuated_preset=preset))

    def with_joint_params(self, stiffness: float | None = None, damping: float | None = None, armature: float | None = None) -> "NMFModel":
        """``Fly.add_joints(..., stiffness=, damping=, armature=)`` (reference ``fly.py:221-299``) for every hinge DoF that is
        not locked."""
        a = dict(self.arrays)
        free = np.ones(self.nv, bool); free[:6] = False
        if "locked_dofs" in a: free[a["locked_dofs"]] = False
        for key, v in (("dof_stiffness", stiffness), ("dof_damping", damping), ("dof_armature", armature)):
            if v is not None:
                arr = a[key].copy(); arr[free] = float(v); a[key] = arr
        return NMFModel(a, self.names, dict(self.meta))

    LOCK_ARMATURE = 1.0e6     # g mm^2: ~1e12 x the tarsal inertias; a locked DoF accelerates by < 1e-5 rad/s^2 under walking loads

    def with_locked_dofs(self, locked_names) -> "NMFModel":
        """Hold the named hinge DoFs at angle 0 (they disappear from every ordering the API exposes).  This is how joint
        presets with fewer DoFs run on kernels written for the LEGS_ONLY chain layout: a rigid attachment is the limit of a
        DoF with infinite armature, so the DoF keeps its lane but gets armature ``LOCK_ARMATURE``, no spring, no damper, no
        actuator and a zero keyframe angle."""
        locked_names = set(locked_names)
        all_dofs = self.names["jointdofs"]
        ex = self.exposed_hinge_dofs()          # position j of names['jointdofs'] is hinge DoF ex[j] of the kernel layout
        idx = np.array([6 + int(ex[j]) for j, nm in enumerate(all_dofs) if nm in locked_names], dtype=np.int32)
        if len(idx) != len(locked_names):
            raise ValueError("unknown DoF names: " + ", ".join(sorted(locked_names - set(all_dofs))))
        act = set(int(d) for d in self.arrays["act_dof"])
        if act & set(int(i) for i in idx):
            raise ValueError("cannot lock an actuated DoF")
        a = dict(self.arrays)
        for key, v in (("dof_stiffness", 0.0), ("dof_damping", 0.0), ("dof_armature", self.LOCK_ARMATURE), ("dof_springref", 0.0)):
            arr = a[key].copy(); arr[idx] = v; a[key] = arr
        key = a["key_qpos"].copy(); key[idx + 1] = 0.0; a["key_qpos"] = key
        a["locked_dofs"] = np.union1d(a.get("locked_dofs", np.zeros(0, np.int32)), idx).astype(np.int32)
        names = dict(self.names, jointdofs=[nm for nm in all_dofs if nm not in locked_names],
                     locked_jointdofs=self.names.get("locked_jointdofs", []) + [nm for nm in all_dofs if nm in locked_names])
        out = NMFModel(a, names, dict(self.meta))
        from .baker.bake import set_const       # inverse weights / meaninertia of the model with the rigid groups fused
        set_const(out)
        return out

    def exposed_hinge_dofs(self) -> np.ndarray:
        """Indices (0-based among the hinge DoFs of the kernel layout) of the DoFs in ``names['jointdofs']`` order."""
        n_h = self.nv - 6
        locked = set(int(i) - 6 for i in self.arrays.get("locked_dofs", []))
        return np.array([j for j in range(n_h) if j not in locked], dtype=np.int64)

    def with_terrain(self, terrain) -> "NMFModel":
        """Copy of the model standing on a box-column terrain: a name from ``TERRAINS`` or the 8 ``TERRAIN_FIELDS`` values."""
        spec = TERRAINS[terrain] if isinstance(terrain, str) else list(terrain)
        if len(spec) != len(TERRAIN_FIELDS):
            raise ValueError(f"terrain needs {len(TERRAIN_FIELDS)} values: {TERRAIN_FIELDS}")
        if (self.arrays["geom_type"] != GEOM_CAPSULE).any():
            raise ValueError("terrain worlds need capsule collision geoms (simplify_geom=True)")
        arrays = dict(self.arrays)
        arrays["terrain"] = np.asarray(spec, dtype=np.float64)
        return NMFModel(arrays, self.names, dict(self.meta, terrain=terrain if isinstance(terrain, str) else "custom"))

    def to_blob(self) -> bytes:
        arrays = dict(self.arrays)
        arrays["opt"] = self._opt_full()
        arrays.setdefault("terrain", np.zeros(len(TERRAIN_FIELDS)))
        arrays.setdefault("weld", np.zeros(len(WELD_FIELDS)))
        secs = [(n, 0, np.ascontiguousarray(arrays[n], dtype=np.float64).ravel()) for n in _F64]
        secs += [(n, 1, np.ascontiguousarray(self.arrays[n], dtype=np.int32).ravel()) for n in _I32]
        header_size = 8 + 8 + len(secs) * (24 + 4 + 4 + 8)
        off = (header_size + 7) // 8 * 8
        table = b""
        payload = b""
        for name, dt, arr in secs:
            raw = arr.tobytes()
            table += struct.pack("<24sii q", name.encode(), dt, arr.size, off + len(payload))
            payload += raw + b"\0" * ((-len(raw)) % 8)
        head = BLOB_MAGIC + struct.pack("<ii", BLOB_VERSION, len(secs)) + table
        head += b"\0" * (off - len(head))
        return head + payload
