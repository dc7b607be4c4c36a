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
              "ls_tolerance", "noslip_iterations", "meaninertia", "impratio"]
CONTACT_FIELDS = ["mu", "solref0", "solref1", "solimp0", "solimp1", "solimp2", "solimp3",
                  "solimp4", "margin", "gap"]
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
GEOM_CAPSULE = 0
GEOM_HULL = 1


@dataclass
class NMFModel:
    arrays: dict[str, np.ndarray]
    names: dict[str, list[str]] = field(default_factory=dict)
    meta: dict = field(default_factory=dict)

    # ---- convenience ----------------------------------------------------
    def dim(self, key: str) -> int:
        return int(self.arrays["dims"][DIM_FIELDS.index(key)])

    def opt(self, key: str) -> float:
        return float(self.arrays["opt"][OPT_FIELDS.index(key)])

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
    def bench(cls, simplify_geom: bool = True, terrain: str | None = None) -> "NMFModel":
        """The reference benchmark model (``time_gpu_simulation.py:21-64``); ``terrain`` = ``"blocks"`` / ``"gapped"``
        replaces the flat ground plane by a box-column terrain (capsule geoms only)."""
        m = cls.load(ASSETS_DIR / ("nmf_bench_capsule.npz" if simplify_geom else "nmf_bench_mesh.npz"))
        return m if terrain in (None, "flat") else m.with_terrain(terrain)

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
        arrays = self.arrays if "terrain" in self.arrays else dict(self.arrays, terrain=np.zeros(len(TERRAIN_FIELDS)))
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
