"""Robot model compiler: URDF -> the SoA tables the edge-validation kernels consume.

Restates the *setup-time* part of the reference's collision closure:
  * moving links / self-collision link pairs:
        pybullet_tools/utils.py:3718-3751 (get_moving_links, get_moving_pairs, get_self_link_pairs)
  * joint limits and the "circular" rule:  pybullet_tools/utils.py:2050-2054, 2128-2137
  * link frames / COM frames as PyBullet reports them (getLinkState [4:6] and [0:2]):
        pybullet_tools/utils.py:2197-2203

Collision geometry is the primitive model the URDF carries (spheres on the moving links,
cylinder/box/sphere on links welded to the base).  The vendored primitive model of the
reference's arm is models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf:61-334.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from itertools import combinations, product
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from .transforms import Frame, axis_angle_matrix, skew
from .urdf import UrdfTree, parse_urdf

# Kernel-side capacity limits (mirrored in csrc/pmp_tables.h).
MAX_DOF = 16
MAX_SPHERES = 32
MAX_LINKS = 64
MAX_FIXED_PRIMS = 16

PRIM_BOX = 0
PRIM_CYLINDER = 1
PRIM_SPHERE = 2


@dataclass
class FixedPrim:
    """A static collision primitive in the world frame."""
    kind: int                      # PRIM_*
    R: np.ndarray                  # 3x3 world rotation of the primitive frame
    t: np.ndarray                  # world position of the primitive centre
    half: np.ndarray               # box: half extents; cylinder: (r, half_len, 0); sphere: (r, 0, 0)
    sphere_mask: int = 0xFFFFFFFF  # bit a set => arm sphere a is tested against it
    name: str = ""


@dataclass
class HullModel:
    """Convex hulls of the arm's collision meshes, for the exact rigid narrow phase (pmp_set_robot_hulls).
    PyBullet turns every URDF collision mesh into a btConvexHullShape of its vertices; collision between two links
    (or a link and a static box) is GJK distance of the margin-free shapes minus both margins <= 0
    (pybullet_tools/utils.py:3390-3441)."""
    link_names: List[str]
    joint: np.ndarray               # [H] joint the hull rides on
    verts: List[np.ndarray]         # H arrays [n_h, 3], in the joint frame
    margin: np.ndarray              # [H] collision margin outside the hull
    pairs: np.ndarray               # [NP, 2] hull index pairs tested against each other
    sph_hull: np.ndarray            # [A] hull each arm sphere approximates (-1 = none)
    sph_radius_outer: np.ndarray    # [A] radii with which a hull's spheres cover hull (+) margin (broad phase)
    fixed_margin: float = 1.0e-3    # margin carved out of static boxes
    sph_radius_inner: Optional[np.ndarray] = None   # [A] radius of the ball around a sphere's centre that lies inside
                                                    # hull (+) margin (0 = unknown): certain hits without GJK


@dataclass
class RobotModel:
    name: str
    dof: int
    joint_names: List[str]
    joint_R0: np.ndarray           # [D,3,3] parent-joint-frame -> joint frame at q=0
    joint_t0: np.ndarray           # [D,3]
    joint_axis: np.ndarray         # [D,3] unit axis in the joint frame
    lower: np.ndarray              # [D]
    upper: np.ndarray              # [D]
    circular: np.ndarray           # [D] bool
    base: Frame                    # world pose of the chain root link frame
    link_names: List[str]          # PyBullet order (index 0 .. L-1)
    link_joint: np.ndarray         # [L] int: active joint the link rides on, -1 = welded to base
    link_R: np.ndarray             # [L,3,3] link frame in its joint frame (or in the root frame)
    link_t: np.ndarray             # [L,3]
    link_com: np.ndarray           # [L,3] inertial origin in the link frame
    sph_joint: np.ndarray          # [A] int
    sph_center: np.ndarray         # [A,3] in the joint frame
    sph_radius: np.ndarray         # [A]
    sph_link: np.ndarray           # [A] PyBullet link index
    self_pair_mask: np.ndarray     # [A,A] bool, upper triangle used
    base_prims: List[FixedPrim] = field(default_factory=list)  # robot's own welded geometry (world frame)
    ee_joint: int = -1
    ee_local: np.ndarray = field(default_factory=lambda: np.zeros(3))
    ee_link: int = -1
    hulls: Optional[HullModel] = None

    # ---------------------------------------------------------------- io
    def to_json(self) -> str:
        def arr(a):
            return np.asarray(a).tolist()
        d = {
            "name": self.name, "dof": self.dof, "joint_names": self.joint_names,
            "joint_R0": arr(self.joint_R0), "joint_t0": arr(self.joint_t0), "joint_axis": arr(self.joint_axis),
            "lower": arr(self.lower), "upper": arr(self.upper), "circular": arr(self.circular.astype(int)),
            "base_R": arr(self.base.R), "base_t": arr(self.base.t),
            "link_names": self.link_names, "link_joint": arr(self.link_joint),
            "link_R": arr(self.link_R), "link_t": arr(self.link_t), "link_com": arr(self.link_com),
            "sph_joint": arr(self.sph_joint), "sph_center": arr(self.sph_center),
            "sph_radius": arr(self.sph_radius), "sph_link": arr(self.sph_link),
            "self_pair_mask": arr(self.self_pair_mask.astype(int)),
            "base_prims": [
                {"kind": p.kind, "R": arr(p.R), "t": arr(p.t), "half": arr(p.half),
                 "sphere_mask": int(p.sphere_mask), "name": p.name} for p in self.base_prims],
            "ee_joint": int(self.ee_joint), "ee_local": arr(self.ee_local), "ee_link": int(self.ee_link),
        }
        if self.hulls is not None:
            h = self.hulls
            d["hulls"] = {"link_names": h.link_names, "joint": arr(h.joint), "margin": arr(h.margin), "pairs": arr(h.pairs),
                          "sph_hull": arr(h.sph_hull), "sph_radius_outer": arr(h.sph_radius_outer),
                          "fixed_margin": float(h.fixed_margin)}     # vertices travel in a .npz next to the JSON
            if h.sph_radius_inner is not None:
                d["hulls"]["sph_radius_inner"] = arr(h.sph_radius_inner)
        return json.dumps(d, indent=1)

    @staticmethod
    def from_json(text: str, hull_verts: Optional[Dict[str, np.ndarray]] = None) -> "RobotModel":
        """``hull_verts``: link name -> [n, 3] hull vertices, required when the JSON carries a "hulls" section."""
        d = json.loads(text)
        f = lambda k: np.asarray(d[k], dtype=np.float64)
        hulls = None
        if "hulls" in d:
            if hull_verts is None:
                raise ValueError("this model has hulls: pass their vertices")
            h = d["hulls"]
            hulls = HullModel(list(h["link_names"]), np.asarray(h["joint"], dtype=np.int64),
                              [np.asarray(hull_verts[n], dtype=np.float64) for n in h["link_names"]],
                              np.asarray(h["margin"], dtype=np.float64), np.asarray(h["pairs"], dtype=np.int64).reshape(-1, 2),
                              np.asarray(h["sph_hull"], dtype=np.int64), np.asarray(h["sph_radius_outer"], dtype=np.float64),
                              float(h["fixed_margin"]),
                              np.asarray(h["sph_radius_inner"], dtype=np.float64) if "sph_radius_inner" in h else None)
        return RobotModel(
            name=d["name"], dof=int(d["dof"]), joint_names=list(d["joint_names"]),
            joint_R0=f("joint_R0"), joint_t0=f("joint_t0"), joint_axis=f("joint_axis"),
            lower=f("lower"), upper=f("upper"), circular=np.asarray(d["circular"], dtype=bool),
            base=Frame(f("base_R"), f("base_t")),
            link_names=list(d["link_names"]), link_joint=np.asarray(d["link_joint"], dtype=np.int64),
            link_R=f("link_R"), link_t=f("link_t"), link_com=f("link_com"),
            sph_joint=np.asarray(d["sph_joint"], dtype=np.int64), sph_center=f("sph_center"),
            sph_radius=f("sph_radius"), sph_link=np.asarray(d["sph_link"], dtype=np.int64),
            self_pair_mask=np.asarray(d["self_pair_mask"], dtype=bool),
            base_prims=[FixedPrim(int(p["kind"]), np.asarray(p["R"], float), np.asarray(p["t"], float),
                                  np.asarray(p["half"], float), int(p["sphere_mask"]), p.get("name", ""))
                        for p in d["base_prims"]],
            ee_joint=int(d["ee_joint"]), ee_local=f("ee_local"), ee_link=int(d["ee_link"]), hulls=hulls,
        )

    def with_base(self, base: Frame) -> "RobotModel":
        """Same robot placed at another world pose (base link frame)."""
        rel = base @ self.base.inverse()
        prims = [FixedPrim(p.kind, rel.R @ p.R, rel.apply(p.t), p.half.copy(), p.sphere_mask, p.name)
                 for p in self.base_prims]
        import copy
        out = copy.copy(self)
        out.base = base.copy()
        out.base_prims = prims
        return out

    @property
    def num_spheres(self) -> int:
        return int(self.sph_radius.shape[0])

    @property
    def num_links(self) -> int:
        return len(self.link_names)


def _link_subtree(tree: UrdfTree, link_idx: int) -> List[int]:
    out = [link_idx]
    for i, j in enumerate(tree.joints):
        if tree.link_index(j.parent) in out and i not in out:
            out.append(i)  # joints are in pre-order, so one pass suffices
    return out


def compile_robot(urdf: str | UrdfTree,
                  active_joints: Optional[Sequence[str]] = None,
                  frozen_positions: Optional[Dict[str, float]] = None,
                  base: Optional[Frame] = None,
                  ee_link: Optional[str] = None,
                  disabled_collisions: Iterable[Tuple[str, str]] = (),
                  name: Optional[str] = None,
                  mesh_spheres: Optional[Dict[str, Sequence[Tuple[Sequence[float], float, float]]]] = None,
                  mesh_hulls: Optional[Dict[str, np.ndarray]] = None,
                  hull_margin: float = 1.0e-3) -> RobotModel:
    """Compile a URDF into kernel tables.

    ``active_joints``: names of the planned joints in planner order (default: every movable
    joint, the reference's ``get_movable_joints``).  They must form one serial chain.
    Other movable joints are frozen at ``frozen_positions`` (default 0).
    ``disabled_collisions``: link-name pairs to skip (the SRDF list of
    pybullet_tools/utils.py:873-880).
    ``mesh_spheres`` / ``mesh_hulls``: for links whose collision geometry is a MESH (which PyBullet loads as the
    convex hull of its vertices): link name -> fitted spheres (centre, radius, covering radius) in the link frame, and
    link name -> hull vertices [n, 3] in the link frame.  The spheres carry the plant contact model and the broad
    phase, the hulls the exact narrow phase (HullModel).
    """
    tree = parse_urdf(urdf) if isinstance(urdf, str) else urdf
    frozen_positions = dict(frozen_positions or {})
    movable = [j.name for j in tree.joints if j.kind in ("revolute", "continuous")]
    if any(j.kind == "prismatic" for j in tree.joints):
        raise NotImplementedError("prismatic joints are not on the reference's arm path")
    if active_joints is None:
        active_joints = movable
    active_joints = list(active_joints)
    D = len(active_joints)
    if not (1 <= D <= MAX_DOF):
        raise ValueError(f"dof {D} outside 1..{MAX_DOF}")
    joint_by_name = {j.name: (i, j) for i, j in enumerate(tree.joints)}
    for n in active_joints:
        if n not in joint_by_name or n not in movable:
            raise ValueError(f"{n!r} is not a movable joint of the URDF")

    L = len(tree.joints)
    # Walk the tree once; for each link record (active joint it rides on, transform from that
    # joint's post-rotation frame -- or from the root frame -- to the link frame).
    link_joint = np.full(L, -1, dtype=np.int64)
    link_rel: List[Frame] = [Frame() for _ in range(L)]
    joint_R0 = np.zeros((D, 3, 3))
    joint_t0 = np.zeros((D, 3))
    joint_axis = np.zeros((D, 3))
    lower = np.zeros(D)
    upper = np.zeros(D)
    circular = np.zeros(D, dtype=bool)
    chain_prev = -1  # active index of the previous chain joint
    for i, j in enumerate(tree.joints):
        pidx = tree.link_index(j.parent)
        p_joint = -1 if pidx < 0 else int(link_joint[pidx])
        p_rel = Frame() if pidx < 0 else link_rel[pidx]
        to_joint = p_rel @ j.origin              # riding-frame -> this joint's frame at q=0
        if j.name in active_joints:
            k = active_joints.index(j.name)
            if k != chain_prev + 1 or p_joint != chain_prev:
                raise ValueError("active joints must form a serial chain listed root-to-tip; "
                                 f"{j.name!r} breaks it")
            chain_prev = k
            joint_R0[k], joint_t0[k] = to_joint.R, to_joint.t
            a = j.axis / np.linalg.norm(j.axis)
            joint_axis[k] = a
            lower[k], upper[k] = j.lower, j.upper
            circular[k] = j.upper < j.lower
            link_joint[i] = k
            link_rel[i] = Frame()                # link frame == joint frame after rotation
        else:
            q = 0.0
            if j.kind in ("revolute", "continuous"):
                q = float(frozen_positions.get(j.name, 0.0))
            rot = Frame(axis_angle_matrix(j.axis, q) if q != 0.0 else np.eye(3))
            link_joint[i] = p_joint
            link_rel[i] = to_joint @ rot
    if chain_prev != D - 1:
        raise ValueError("not every active joint was found in the tree")

    link_names = [j.child for j in tree.joints]
    link_R = np.stack([f.R for f in link_rel])
    link_t = np.stack([f.t for f in link_rel])
    link_com = np.stack([tree.links[n].inertial_origin.t for n in link_names])

    # ---- can_collide / moving links / self pairs (pybullet_tools/utils.py:3718-3751) ----
    def can_collide(idx: int) -> bool:
        nm = tree.root if idx < 0 else link_names[idx]
        return len(tree.links[nm].collisions) > 0

    active_idx = [joint_by_name[n][0] for n in active_joints]
    moving: List[int] = []
    for ji in active_idx:
        for l in _link_subtree(tree, ji):
            if l not in moving:
                moving.append(l)
    moving_c = [l for l in moving if can_collide(l)]
    fixed_c = [l for l in ([-1] + list(range(L))) if l not in moving and can_collide(l)]

    def moving_ancestors(l: int) -> frozenset:
        out = set()
        while l >= 0:
            if l in active_idx:
                out.add(l)
            l = tree.parent_index(l)
        return frozenset(out)

    def adjacent(a: int, b: int) -> bool:
        pa = tree.parent_index(a) if a >= 0 else None
        pb = tree.parent_index(b) if b >= 0 else None
        return pa == b or pb == a

    disabled: Set[Tuple[int, int]] = set()
    for n1, n2 in disabled_collisions:
        try:
            disabled.add((tree.link_index(n1), tree.link_index(n2)))
        except KeyError:
            continue
    pairs = list(product(moving_c, fixed_c))
    pairs += [(a, b) for a, b in combinations(moving_c, 2) if moving_ancestors(a) != moving_ancestors(b)]
    pairs = [p for p in pairs if not adjacent(*p)]
    pairs = [p for p in pairs if p not in disabled and p[::-1] not in disabled]

    # ---- primitives ----
    sph_joint: List[int] = []
    sph_center: List[np.ndarray] = []
    sph_radius: List[float] = []
    sph_link: List[int] = []
    base_prims: List[FixedPrim] = []
    base_prim_link: List[int] = []
    sph_outer: List[float] = []
    sph_inner: List[float] = []
    hull_links: List[int] = []
    hull_vert_list: List[np.ndarray] = []
    root_frame = base if base is not None else Frame()
    for l in ([-1] + list(range(L))):
        nm = tree.root if l < 0 else link_names[l]
        rel = Frame() if l < 0 else link_rel[l]
        jn = -1 if l < 0 else int(link_joint[l])
        for g in tree.links[nm].collisions:
            if g.kind == "mesh":
                if jn < 0 or mesh_spheres is None or nm not in mesh_spheres:
                    raise NotImplementedError(
                        f"link {nm!r} carries a mesh; pass its fitted spheres (mesh_spheres) and hull (mesh_hulls), "
                        "or use a URDF with primitive collision geometry")
                for entry in mesh_spheres[nm]:
                    cen, rad, rad_out = entry[:3]
                    sph_inner.append(float(entry[3]) if len(entry) > 3 else 0.0)
                    sph_joint.append(jn)
                    sph_center.append((rel @ g.origin).apply(np.asarray(cen, dtype=np.float64)))
                    sph_radius.append(float(rad))
                    sph_link.append(l)
                    sph_outer.append(float(rad_out))
                if mesh_hulls is not None and nm in mesh_hulls:
                    fr = rel @ g.origin
                    hull_links.append(l)
                    hull_vert_list.append(np.asarray(mesh_hulls[nm], dtype=np.float64) @ fr.R.T + fr.t[None])
                continue
            pose = rel @ g.origin
            if jn >= 0:
                if g.kind != "sphere":
                    raise NotImplementedError("moving links must carry spheres")
                sph_joint.append(jn)
                sph_center.append(pose.t)
                sph_radius.append(float(g.size[0]))
                sph_link.append(l)
                sph_outer.append(float(g.size[0]))
                sph_inner.append(0.0)
            else:
                w = root_frame @ pose
                if g.kind == "box":
                    base_prims.append(FixedPrim(PRIM_BOX, w.R, w.t, 0.5 * g.size, 0, nm))
                elif g.kind == "cylinder":
                    base_prims.append(FixedPrim(PRIM_CYLINDER, w.R, w.t,
                                                np.array([g.size[0], 0.5 * g.size[1], 0.0]), 0, nm))
                else:
                    base_prims.append(FixedPrim(PRIM_SPHERE, w.R, w.t, np.array([g.size[0], 0.0, 0.0]), 0, nm))
                base_prim_link.append(l)
    A = len(sph_radius)
    if A > MAX_SPHERES:
        raise ValueError(f"{A} spheres > {MAX_SPHERES}")
    order = np.argsort(np.asarray(sph_joint), kind="stable")  # root-to-tip, file order inside a link
    sph_joint_a = np.asarray(sph_joint, dtype=np.int64)[order]
    sph_center_a = np.asarray(sph_center, dtype=np.float64).reshape(-1, 3)[order]
    sph_radius_a = np.asarray(sph_radius, dtype=np.float64)[order]
    sph_link_a = np.asarray(sph_link, dtype=np.int64)[order]
    sph_outer_a = np.asarray(sph_outer, dtype=np.float64)[order]
    sph_inner_a = np.asarray(sph_inner, dtype=np.float64)[order]

    pair_set = {frozenset(p) for p in pairs}
    mask = np.zeros((A, A), dtype=bool)
    for a, b in combinations(range(A), 2):
        if frozenset((int(sph_link_a[a]), int(sph_link_a[b]))) in pair_set and sph_link_a[a] != sph_link_a[b]:
            mask[a, b] = True
    for prim, pl in zip(base_prims, base_prim_link):
        m = 0
        for a in range(A):
            if frozenset((int(sph_link_a[a]), pl)) in pair_set:
                m |= (1 << a)
        prim.sphere_mask = m

    if ee_link is None:
        ee_idx = L - 1
    else:
        ee_idx = tree.link_index(ee_link)
    ee_frame = link_rel[ee_idx]
    ee_local = ee_frame.apply(link_com[ee_idx])

    hulls = None
    if hull_links:
        hp = [(i, j) for i, j in combinations(range(len(hull_links)), 2)
              if frozenset((hull_links[i], hull_links[j])) in pair_set]
        sph_hull = np.array([hull_links.index(int(l)) if int(l) in hull_links else -1 for l in sph_link_a], dtype=np.int64)
        hulls = HullModel([link_names[l] for l in hull_links], np.array([int(link_joint[l]) for l in hull_links]),
                          hull_vert_list, np.full(len(hull_links), float(hull_margin)),
                          np.asarray(hp, dtype=np.int64).reshape(-1, 2), sph_hull, sph_outer_a,
                          sph_radius_inner=sph_inner_a if np.any(sph_inner_a > 0) else None)

    return RobotModel(
        name=name or tree.name, dof=D, joint_names=active_joints,
        joint_R0=joint_R0, joint_t0=joint_t0, joint_axis=joint_axis,
        lower=lower, upper=upper, circular=circular, base=root_frame.copy(),
        link_names=link_names, link_joint=link_joint, link_R=link_R, link_t=link_t, link_com=link_com,
        sph_joint=sph_joint_a, sph_center=sph_center_a, sph_radius=sph_radius_a, sph_link=sph_link_a,
        self_pair_mask=mask, base_prims=base_prims,
        ee_joint=int(link_joint[ee_idx]), ee_local=ee_local, ee_link=int(ee_idx), hulls=hulls,
    )


def _joint_frames_np(model: RobotModel, Q: np.ndarray):
    B, D = Q.shape
    R = np.zeros((B, D, 3, 3))
    t = np.zeros((B, D, 3))
    Rc = np.broadcast_to(model.base.R, (B, 3, 3)).copy()
    tc = np.broadcast_to(model.base.t, (B, 3)).copy()
    for j in range(D):
        a = model.joint_axis[j]
        K = skew(a)
        s = np.sin(Q[:, j])[:, None, None]
        c = np.cos(Q[:, j])[:, None, None]
        rot = np.eye(3)[None] + s * K[None] + (1 - c) * (K @ K)[None]
        tc = tc + np.einsum("bij,j->bi", Rc, model.joint_t0[j])
        Rc = Rc @ model.joint_R0[j][None] @ rot
        R[:, j], t[:, j] = Rc, tc
    return R, t


def prune_always_colliding(model: RobotModel, samples: int = 4096, seed: int = 0) -> List[Tuple[int, int]]:
    """Disable self-collision sphere pairs that overlap at the zero configuration *and* at every one
    of ``samples`` uniform configurations (fixed seed) -- the primitive model's analogue of the SRDF
    ``disable_collisions`` list the reference loads (pybullet_tools/utils.py:873-880; env.py:84-91).
    For the vendored iiwa14 sphere model this removes exactly one pair: the distal sphere of link 5
    vs the sphere of link 7 (always 14-15 mm deep; the meshes the reference tests do not touch there).
    Returns the removed (a, b) sphere index pairs; ``model`` is modified in place."""
    rng = np.random.RandomState(seed)
    lo = np.where(model.circular, -np.pi, model.lower)
    hi = np.where(model.circular, np.pi, model.upper)
    Q = rng.uniform(lo, hi, size=(samples, model.dof))
    Q[0] = np.clip(0.0, lo, hi)
    R, t = _joint_frames_np(model, Q)
    A = model.num_spheres
    c = np.zeros((samples, A, 3))
    for a in range(A):
        j = int(model.sph_joint[a])
        c[:, a] = np.einsum("bij,j->bi", R[:, j], model.sph_center[a]) + t[:, j]
    removed = []
    for a in range(A):
        for b in range(a + 1, A):
            if model.self_pair_mask[a, b]:
                rr = model.sph_radius[a] + model.sph_radius[b]
                if np.all(np.sum((c[:, a] - c[:, b]) ** 2, axis=1) <= rr * rr):
                    model.self_pair_mask[a, b] = False
                    removed.append((a, b))
    return removed


def joint_constant_matrices(model: RobotModel) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per joint, the three constant matrices with  R0 @ Rot(axis, q) = C + cos(q) A + sin(q) B.

    (Rodrigues: Rot = cos I + sin [a]x + (1-cos) a a^T.)  The kernels evaluate the joint
    rotation with 18 FMAs from these instead of building Rot(axis, q) and multiplying.
    """
    D = model.dof
    Cm = np.zeros((D, 3, 3))
    Am = np.zeros((D, 3, 3))
    Bm = np.zeros((D, 3, 3))
    for k in range(D):
        a = model.joint_axis[k]
        aa = np.outer(a, a)
        Cm[k] = model.joint_R0[k] @ aa
        Am[k] = model.joint_R0[k] @ (np.eye(3) - aa)
        Bm[k] = model.joint_R0[k] @ skew(a)
    return Cm, Am, Bm
