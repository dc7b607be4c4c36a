"""Naming tables and orderings of the NeuroMechFly body plan.

Restates (does not import) the parts of the reference's ``flygym.anatomy`` that
fix the ordering contract of the Simulation API:

* segment names / tree   -> reference ``src/flygym/anatomy.py:192-227``
* DoF naming + per-joint axis sets (``JointPreset.LEGS_ONLY`` etc.)
                          -> ``anatomy.py:388-460``
* actuated DoF presets    -> ``anatomy.py:463-498``
* contact-body presets    -> ``anatomy.py:501-562``
* DFS DoF iteration       -> ``anatomy.py:615-626`` + ``utils/math.py:92-105``

Everything here is plain strings/tuples so the host code never needs
``dm_control``/``mujoco``.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

SIDES = ["l", "r"]
LEGS = [f"{s}{p}" for s in SIDES for p in "fmh"]  # lf lm lh rf rm rh
LEG_LINKS = ["coxa", "trochanterfemur", "tibia"] + [f"tarsus{i}" for i in "12345"]
ANTENNA_LINKS = ["pedicel", "funiculus", "arista"]
PROBOSCIS_LINKS = ["rostrum", "haustellum"]
ABDOMEN_LINKS = ["abdomen12"] + [f"abdomen{i}" for i in "3456"]
PASSIVE_TARSAL_LINKS = [f"tarsus{i}" for i in "2345"]

AXIS_VECTOR = {"pitch": (0.0, 1.0, 0.0), "roll": (0.0, 0.0, 1.0), "yaw": (1.0, 0.0, 0.0)}


def _chain(*names):
    return [(names[i], names[i + 1]) for i in range(len(names) - 1)]


ALL_CONNECTED_SEGMENT_PAIRS: list[tuple[str, str]] = (
    [("c_thorax", "c_head")]
    + _chain("c_head", *(f"c_{k}" for k in PROBOSCIS_LINKS))
    + _chain("c_thorax", *(f"c_{k}" for k in ABDOMEN_LINKS))
    + [("c_head", f"{s}_eye") for s in SIDES]
    + [e for s in SIDES for e in _chain("c_head", *(f"{s}_{k}" for k in ANTENNA_LINKS))]
    + [("c_thorax", f"{s}_wing") for s in SIDES]
    + [("c_thorax", f"{s}_haltere") for s in SIDES]
    + [e for leg in LEGS for e in _chain("c_thorax", *(f"{leg}_{k}" for k in LEG_LINKS))]
)
ALL_SEGMENT_NAMES: list[str] = list(
    dict.fromkeys(s for pair in ALL_CONNECTED_SEGMENT_PAIRS for s in pair)
)


def seg_pos(name: str) -> str:
    return name.split("_")[0]


def seg_link(name: str) -> str:
    return name.split("_")[1]


def is_leg(name: str) -> bool:
    return seg_pos(name) in LEGS


class AxisOrder(Enum):
    PITCH_ROLL_YAW = ("pitch", "roll", "yaw")
    PITCH_YAW_ROLL = ("pitch", "yaw", "roll")
    ROLL_PITCH_YAW = ("roll", "pitch", "yaw")
    ROLL_YAW_PITCH = ("roll", "yaw", "pitch")
    YAW_PITCH_ROLL = ("yaw", "pitch", "roll")
    YAW_ROLL_PITCH = ("yaw", "roll", "pitch")


class ActuatorType(Enum):
    """Same members/values as the reference's ``compose/fly.py:64-77``."""

    MOTOR = "motor"
    POSITION = "position"
    VELOCITY = "velocity"
    INTVELOCITY = "intvelocity"
    DAMPER = "damper"
    CYLINDER = "cylinder"
    MUSCLE = "muscle"
    ADHESION = "adhesion"


@dataclass(frozen=True)
class JointDOF:
    parent: str
    child: str
    axis: str

    @property
    def name(self) -> str:
        return f"{self.parent}-{self.child}-{self.axis}"


def dfs_edges(pairs: list[tuple[str, str]], root: str = "c_thorax"):
    """Pre-order DFS visiting children in edge-insertion order
    (same traversal as the reference ``Tree.dfs_edges``)."""
    children: dict[str, list[str]] = {}
    for p, c in pairs:
        children.setdefault(p, []).append(c)
        children.setdefault(c, [])
    out = []
    stack = [(None, root)]
    while stack:
        parent, node = stack.pop()
        if parent is not None:
            out.append((parent, node))
        stack.extend((node, ch) for ch in reversed(children[node]))
    return out


def bodysegs_order(root: str = "c_thorax") -> list[str]:
    """Body order exposed by ``Fly.get_bodysegs_order`` (``fly.py:545-582``)."""
    return [root] + [c for _, c in dfs_edges(ALL_CONNECTED_SEGMENT_PAIRS, root)]


def _axes_for_joint(child: str, preset: str) -> set[str]:
    axes = {"pitch", "roll", "yaw"}
    if preset == "all_possible":
        return axes
    if is_leg(child):
        link = seg_link(child)
        if link == "coxa":
            pass
        elif link == "trochanterfemur":
            axes.discard("yaw")
        else:
            axes = {"pitch"}
    return axes


def joint_pairs(preset: str) -> list[tuple[str, str]]:
    """Anatomical joints kept by a ``JointPreset`` value (``anatomy.py:411-460``)."""
    if preset in ("all_possible", "all_biological"):
        return list(ALL_CONNECTED_SEGMENT_PAIRS)
    if preset == "legs_only":
        return [(p, c) for p, c in ALL_CONNECTED_SEGMENT_PAIRS if is_leg(c)]
    if preset == "legs_active_only":
        return [
            (p, c)
            for p, c in ALL_CONNECTED_SEGMENT_PAIRS
            if is_leg(c) and seg_link(c) not in PASSIVE_TARSAL_LINKS
        ]
    raise ValueError(f"unknown joint preset {preset!r}")


def jointdofs_order(preset: str = "legs_only",
                    axis_order: AxisOrder = AxisOrder.YAW_PITCH_ROLL,
                    root: str = "c_thorax") -> list[JointDOF]:
    """DoF order of ``Skeleton.iter_jointdofs`` (``anatomy.py:615-626``)."""
    pairs = joint_pairs(preset)
    out = []
    for p, c in dfs_edges(pairs, root):
        axes = _axes_for_joint(c, preset)
        for ax in axis_order.value:
            if ax in axes:
                out.append(JointDOF(p, c, ax))
    return out


def actuated_dofs(dofs: list[JointDOF], preset: str = "legs_active_only") -> list[JointDOF]:
    """``ActuatedDOFPreset.filter`` (``anatomy.py:480-498``)."""
    if preset == "all":
        return list(dofs)
    legs = [d for d in dofs if is_leg(d.child)]
    if preset == "legs_only":
        return legs
    if preset == "legs_active_only":
        return [d for d in legs if seg_link(d.child) not in PASSIVE_TARSAL_LINKS]
    raise ValueError(f"unknown actuated-dof preset {preset!r}")


def contact_bodies(preset: str = "legs_thorax_abdomen_head") -> list[str]:
    """``ContactBodiesPreset.to_body_segments_list`` (``anatomy.py:524-562``)."""
    if preset == "all":
        return list(ALL_SEGMENT_NAMES)
    if preset == "legs_thorax_abdomen_head":
        return [
            s for s in ALL_SEGMENT_NAMES
            if is_leg(s) or s == "c_thorax" or seg_link(s) in ABDOMEN_LINKS or s == "c_head"
        ]
    if preset == "legs_only":
        return [s for s in ALL_SEGMENT_NAMES if is_leg(s)]
    if preset == "tibia_tarsus_only":
        return [s for s in ALL_SEGMENT_NAMES
                if is_leg(s) and (seg_link(s) == "tibia" or seg_link(s).startswith("tarsus"))]
    raise ValueError(f"unknown contact preset {preset!r}")
