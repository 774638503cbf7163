"""Minimal URDF reader: links (inertial origin, primitive collision geometry) and joints.

Only what the edge-validation path needs.  Link indices follow PyBullet's convention
(link index == index of the joint whose child the link is, joints numbered by a pre-order
walk of the tree with children in file order; the root link is -1), because the reference
addresses links that way (e.g. ``getLinkState(robot, 9)`` = ``iiwa_link_ee`` in
plant_motion_planning/pybullet_tools/kuka_primitives.py:478).
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .transforms import Frame


def _floats(text: Optional[str], n: int, default: float = 0.0) -> np.ndarray:
    if text is None:
        return np.full(n, default, dtype=np.float64)
    vals = [float(v) for v in text.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} floats, got {text!r}")
    return np.asarray(vals, dtype=np.float64)


def _origin(elem) -> Frame:
    if elem is None:
        return Frame()
    o = elem.find("origin")
    if o is None:
        return Frame()
    return Frame.from_xyz_rpy(_floats(o.get("xyz"), 3), _floats(o.get("rpy"), 3))


@dataclass
class Geometry:
    kind: str                 # "sphere" | "box" | "cylinder" | "mesh"
    origin: Frame             # pose in the link frame
    size: np.ndarray          # sphere: [r]; box: full extents [x,y,z]; cylinder: [r, length]; mesh: []
    filename: str = ""


@dataclass
class Link:
    name: str
    inertial_origin: Frame
    mass: float
    collisions: List[Geometry] = field(default_factory=list)


@dataclass
class Joint:
    name: str
    kind: str                 # revolute | continuous | prismatic | fixed
    parent: str
    child: str
    origin: Frame
    axis: np.ndarray
    lower: float
    upper: float


@dataclass
class UrdfTree:
    name: str
    links: Dict[str, Link]
    joints: List[Joint]              # PyBullet order (index == child link index)
    root: str

    def link_index(self, name: str) -> int:
        if name == self.root:
            return -1
        for i, j in enumerate(self.joints):
            if j.child == name:
                return i
        raise KeyError(name)

    def parent_index(self, link_idx: int) -> int:
        """PyBullet link index of the parent link of ``link_idx``."""
        return self.link_index(self.joints[link_idx].parent)


def _parse_geometry(c) -> Optional[Geometry]:
    g = c.find("geometry")
    if g is None:
        return None
    origin = _origin(c)
    for child in g:
        tag = child.tag
        if tag == "sphere":
            return Geometry("sphere", origin, np.array([float(child.get("radius"))]))
        if tag == "box":
            return Geometry("box", origin, _floats(child.get("size"), 3))
        if tag == "cylinder":
            return Geometry("cylinder", origin, np.array([float(child.get("radius")), float(child.get("length"))]))
        if tag == "mesh":
            return Geometry("mesh", origin, np.zeros(0), child.get("filename", ""))
    return None


def parse_urdf(source: str, is_text: bool = False) -> UrdfTree:
    root_el = ET.fromstring(source) if is_text else ET.parse(source).getroot()
    links: Dict[str, Link] = {}
    for le in root_el.findall("link"):
        inertial = le.find("inertial")
        mass = 0.0
        if inertial is not None and inertial.find("mass") is not None:
            mass = float(inertial.find("mass").get("value", "0"))
        link = Link(le.get("name"), _origin(inertial), mass)
        for c in le.findall("collision"):
            geo = _parse_geometry(c)
            if geo is not None:
                link.collisions.append(geo)
        links[link.name] = link

    raw: List[Joint] = []
    for je in root_el.findall("joint"):
        if je.find("parent") is None or je.find("child") is None:
            continue  # <transmission><joint name=.../></transmission> never reaches here, but be safe
        axis_el = je.find("axis")
        axis = _floats(axis_el.get("xyz"), 3) if axis_el is not None else np.array([1.0, 0.0, 0.0])
        lim = je.find("limit")
        lower = float(lim.get("lower", "0")) if lim is not None else 0.0
        upper = float(lim.get("upper", "0")) if lim is not None else 0.0
        kind = je.get("type")
        if kind == "continuous":
            # PyBullet reports lower=0, upper=-1 for unlimited joints, which the reference
            # reads as "circular" (pybullet_tools/utils.py:2050-2054: upper < lower).
            lower, upper = 0.0, -1.0
        raw.append(Joint(je.get("name"), kind, je.find("parent").get("link"), je.find("child").get("link"),
                         _origin(je), axis, lower, upper))

    children = {j.child for j in raw}
    roots = [n for n in links if n not in children]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")
    root = roots[0]

    by_parent: Dict[str, List[Joint]] = {}
    for j in raw:
        by_parent.setdefault(j.parent, []).append(j)
    ordered: List[Joint] = []

    def walk(link_name: str) -> None:
        for j in by_parent.get(link_name, []):
            ordered.append(j)
            walk(j.child)

    walk(root)
    return UrdfTree(root_el.get("name", ""), links, ordered, root)
