"""Convex-hull collision oracle for the arm the reference scripts load.        *** TEST INFRASTRUCTURE ***

Only tests/, tools/ (fitting and reports run in the build container), __graft_entry__.smoke() and bench.py may import
this module; the product path never does.

What it restates (paths relative to /root/reference/plant_motion_planning/ unless noted):
  * the rigid part of the reference's collision closure, pybullet_tools/utils.py:3965-4016
    (get_collision_fn_with_controls_v3): joint limits (:3753-3762), self link pairs (:3740-3751, :3996-4006), moving
    links x fixed bodies (:4008-4014), each pair decided by p.getClosestPoints(distance=0) non-empty (:3390-3441);
  * ON THE GEOMETRY THE SCRIPTS LOAD: DRAKE_IIWA_URDF_EDIT (pybullet_tools/utils.py:78;
    /root/reference/models/drake/iiwa_description/urdf/iiwa14_polytope_collision_edit.urdf:89-307) = convex hulls of
    meshes/collision/link_{1..6}.obj and link_7_polytope.obj (iiwa_link_0 and the flange links carry no collision
    element), the floor box of models/short_floor.urdf, the block of
    models/drake/objects/block_for_pick_and_place_mid_size.urdf (box 0.059 x 0.089 x 0.199 plus eight 1e-7 m corner
    spheres) posed by scripts/our_method_multi_branch.py:113-115, and the plant stems as the 0.05 x 0.05 x 2h boxes of
    rand_gen_plants_programmatically_main.py:30-51.
The narrow phase lives in PyBullet 3.2.0 (Bullet 3.x C++, not under /root/reference; restated from Bullet's public
source, to be re-verified if PyBullet ever becomes installable): btGjkPairDetector on the margin-free convex shapes,
signed distance = GJK distance - margin_A - margin_B, a point is reported iff that is <= the query distance 0.
Margins: URDF and createCollisionShape shapes get 0.001 m; a btConvexHullShape's margin lies OUTSIDE its vertices'
hull, a btBoxShape's margin is carved out of its half extents (core box = half - margin, faces stay where the URDF
puts them, edges and corners are rounded by the margin).  PARITY STATUS: unpinned against PyBullet itself (cannot run
here); pinned against an independent QP solution of the same distance problem (tests/test_hull_oracle.py).

The forward kinematics here are composed from the joint records parsed out of the URDF text by
tools/extract_iiwa_hulls.py (tests/golden/iiwa14_collision_hulls.npz), NOT from the product's model compiler, so that
a mistake in the product's tables shows up as a disagreement.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libhull_gjk.so")
FIXTURE = os.path.join(HERE, "..", "tests", "golden", "iiwa14_collision_hulls.npz")
BULLET_MARGIN = 1.0e-3     # gUrdfDefaultCollisionMargin / m_defaultCollisionMargin (Bullet public source)

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "hull_gjk.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-std=gnu11",
                        "-fno-fast-math", "-o", LIB, src, "-lm"], check=True, capture_output=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(LIB)
        except OSError:
            build(force=True)
            _lib = C.CDLL(LIB)
        vp = C.c_void_p
        _lib.hull_gjk_pair.restype = C.c_double
        _lib.hull_gjk_pair.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
        _lib.hull_gjk_batch.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, C.c_double, vp]
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def gjk_pair(VA: np.ndarray, FA: np.ndarray, VB: np.ndarray, FB: np.ndarray) -> float:
    """Margin-free distance between conv(R_A VA + t_A) and conv(R_B VB + t_B); F = (R row-major, t) as 12 numbers."""
    VA = np.ascontiguousarray(VA, dtype=np.float64); VB = np.ascontiguousarray(VB, dtype=np.float64)
    FA = np.ascontiguousarray(FA, dtype=np.float64).reshape(12); FB = np.ascontiguousarray(FB, dtype=np.float64).reshape(12)
    return float(_load().hull_gjk_pair(_ptr(VA), len(VA), _ptr(FA), _ptr(VB), len(VB), _ptr(FB)))


# ----------------------------------------------------------------------------- URDF conventions (own restatement)
def rpy_matrix(r: float, p: float, y: float) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(y) Ry(p) Rx(r)."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def box_vertices(half: Sequence[float]) -> np.ndarray:
    h = np.asarray(half, dtype=np.float64)
    s = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64)
    return s * h[None]


class Shape:
    """A convex shape = vertex set (local frame) + Bullet margin.  kind 'hull': margin outside the vertices' hull;
    kind 'box': vertices are the corners of the core box (half extents - margin)."""

    def __init__(self, verts: np.ndarray, margin: float, name: str = ""):
        self.verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
        self.margin = float(margin)
        self.name = name
        c = 0.5 * (self.verts.min(0) + self.verts.max(0))
        self.bs_center = c
        self.bs_radius = float(np.sqrt(((self.verts - c) ** 2).sum(1).max()))

    @staticmethod
    def bullet_box(half: Sequence[float], margin: float = BULLET_MARGIN, name: str = "") -> "Shape":
        h = np.asarray(half, dtype=np.float64)
        return Shape(box_vertices(np.maximum(h - margin, 0.0)), margin, name)


def frame12(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    return np.concatenate([np.asarray(R, dtype=np.float64).reshape(9), np.asarray(t, dtype=np.float64).reshape(3)])


class HullArm:
    """iiwa14 as the scripts load it: kinematic chain + per-link convex hulls from the fixture."""

    def __init__(self, fixture: str = FIXTURE, base_R: Optional[np.ndarray] = None, base_t: Optional[np.ndarray] = None,
                 margin: float = BULLET_MARGIN):
        d = np.load(fixture, allow_pickle=False)
        self.link_names = [str(n) for n in d["link_names"]]
        self.hulls = [Shape(d[f"verts_{n}"], margin, n) for n in self.link_names]
        self.faces = [d[f"faces_{n}"] for n in self.link_names]
        names = [str(n) for n in d["joint_names"]]
        kinds = [str(n) for n in d["joint_types"]]
        child = [str(n) for n in d["joint_child"]]
        parent = [str(n) for n in d["joint_parent"]]
        act = [i for i, k in enumerate(kinds) if k in ("revolute", "continuous")]
        self.dof = len(act)
        self.joint_names = [names[i] for i in act]
        self.origin_R = [rpy_matrix(*d["joint_rpy"][i]) for i in act]
        self.origin_t = [np.asarray(d["joint_xyz"][i]) for i in act]
        self.axis = [np.asarray(d["joint_axis"][i]) / np.linalg.norm(d["joint_axis"][i]) for i in act]
        self.lower = np.array([d["joint_limits"][i][0] for i in act])
        self.upper = np.array([d["joint_limits"][i][1] for i in act])
        # serial chain check: each active joint's parent is the previous active joint's child
        for k in range(1, self.dof):
            assert parent[act[k]] == child[act[k - 1]], "not a serial chain"
        # fixed joints between the world root and the first active joint must be identity here
        for i in range(act[0]):
            assert np.allclose(d["joint_xyz"][i], 0) and np.allclose(d["joint_rpy"][i], 0)
        self.link_of_joint = [child[i] for i in act]
        # hull k rides on active joint hull_joint[k] with identity offset (collision origins are identity)
        self.hull_joint = [self.link_of_joint.index(n) for n in self.link_names]
        self.base_R = np.eye(3) if base_R is None else np.asarray(base_R, dtype=np.float64)
        self.base_t = np.zeros(3) if base_t is None else np.asarray(base_t, dtype=np.float64)
        # self pairs, restating get_self_link_pairs (pybullet_tools/utils.py:3740-3751) on this tree: every pair of
        # moving links that can collide (have a collision element) and are not adjacent (parent/child); moving x fixed
        # pairs vanish because iiwa_link_0 / base carry no collision element.
        self.self_pairs = [(a, b) for a in range(len(self.hulls)) for b in range(a + 1, len(self.hulls))
                           if abs(self.hull_joint[a] - self.hull_joint[b]) > 1]

    def joint_frames(self, Q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """World frames of the 7 moving links: T_j = T_{j-1} . origin_j . Rot(axis_j, q_j)  (URDF semantics)."""
        Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
        B = Q.shape[0]
        R = np.zeros((B, self.dof, 3, 3)); t = np.zeros((B, self.dof, 3))
        Rc = np.broadcast_to(self.base_R, (B, 3, 3)).copy(); tc = np.broadcast_to(self.base_t, (B, 3)).copy()
        for j in range(self.dof):
            a = self.axis[j]
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            s = np.sin(Q[:, j])[:, None, None]; c = np.cos(Q[:, j])[:, None, None]
            rot = np.eye(3)[None] + s * K[None] + (1 - c) * (K @ K)[None]
            tc = tc + np.einsum("bij,j->bi", Rc, self.origin_t[j])
            Rc = Rc @ self.origin_R[j][None] @ rot
            R[:, j], t[:, j] = Rc, tc
        return R, t

    def hull_frames(self, Q: np.ndarray) -> np.ndarray:
        R, t = self.joint_frames(Q)
        B = R.shape[0]
        out = np.zeros((B, len(self.hulls), 12))
        for k, j in enumerate(self.hull_joint):
            out[:, k, :9] = R[:, j].reshape(B, 9)
            out[:, k, 9:] = t[:, j]
        return out

    def limits_violated(self, Q: np.ndarray) -> np.ndarray:
        Q = np.atleast_2d(Q)
        return ~(np.all(self.lower <= Q, axis=1) & np.all(Q <= self.upper, axis=1))


def default_fixed_shapes(offset_xy=(0.0, 0.0)) -> List[Tuple[Shape, np.ndarray]]:
    """(shape, frame12) of the fixed bodies every script loads: floor and block
    (scripts/our_method_multi_branch.py:107-115; env.py:38-47)."""
    ox, oy = offset_xy
    out = []
    out.append((Shape.bullet_box([1.0, 1.0, 0.05], name="floor"), frame12(np.eye(3), [ox, oy, -0.05])))
    yaw = 1.57
    Rb = rpy_matrix(0.0, 0.0, yaw)
    tb = np.array([0.4 + ox, -0.4 + oy, 0.45])
    out.append((Shape.bullet_box([0.0295, 0.0445, 0.0995], name="block"), frame12(Rb, tb)))
    # the block's eight 1e-7 m contact spheres at the corners of its visual box
    corners = np.array([[sx * 0.03, sy * 0.0375, sz * 0.1] for sz in (-1, 1) for sx in (-1, 1) for sy in (-1, 1)])
    for i, c in enumerate(corners):
        out.append((Shape(np.zeros((1, 3)), 1e-7, f"block_corner{i}"), frame12(Rb, Rb @ c + tb)))
    return out


def stem_box_shapes(world, margin: float = BULLET_MARGIN) -> List[Tuple[Shape, np.ndarray]]:
    """The stems of one PlantWorld at rest as Bullet boxes (rand_gen_plants_programmatically_main.py:30-51: half
    extents (w, w, h), centred h above the stem frame origin + z_offset), plus the plant's base boxes."""
    from plant_motion_planning_b200.model.tables import stem_rest_frames   # rest frames of the pinned sampler
    rest = stem_rest_frames(world)
    out = []
    for s, fr in zip(world.stems, rest):
        c = fr.apply(np.array([0.0, 0.0, s.z_offset + s.half_height]))
        out.append((Shape.bullet_box([s.half_width, s.half_width, s.half_height], margin, "stem"), frame12(fr.R, c)))
    return out


def batch_distances(shapes: Sequence[Shape], frames: np.ndarray, pairs: Sequence[Tuple[int, int]],
                    far: float = 0.05) -> np.ndarray:
    """Signed Bullet distance (GJK - both margins) [B, NP] of shape pairs posed by frames [B, S, 12]; pairs further
    apart than `far` report a lower bound; intersecting cores report -(margin_a + margin_b)."""
    S = len(shapes)
    off = np.zeros(S + 1, dtype=np.int32)
    for i, s in enumerate(shapes):
        off[i + 1] = off[i] + len(s.verts)
    verts = np.ascontiguousarray(np.concatenate([s.verts for s in shapes], axis=0))
    bs_c = np.ascontiguousarray(np.stack([s.bs_center for s in shapes]))
    bs_r = np.ascontiguousarray(np.array([s.bs_radius for s in shapes]))
    margins = np.array([s.margin for s in shapes])
    frames = np.ascontiguousarray(frames, dtype=np.float64)
    B = frames.shape[0]
    pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    out = np.zeros((B, len(pr)))
    if len(pr) and B:
        _load().hull_gjk_batch(B, S, _ptr(off), _ptr(verts), _ptr(bs_c), _ptr(bs_r), _ptr(frames), len(pr), _ptr(pr),
                               float(far), _ptr(out))
    return out - (margins[pr[:, 0]] + margins[pr[:, 1]])[None]


class HullScene:
    """Arm hulls + fixed shapes; batched signed distances through the C GJK."""

    def __init__(self, arm: Optional[HullArm] = None, fixed: Optional[List[Tuple[Shape, np.ndarray]]] = None):
        self.arm = arm or HullArm()
        self.fixed = default_fixed_shapes() if fixed is None else list(fixed)
        self._pack()

    def _pack(self):
        shapes = self.arm.hulls + [s for s, _ in self.fixed]
        self.shapes = shapes
        self.off = np.zeros(len(shapes) + 1, dtype=np.int32)
        for i, s in enumerate(shapes):
            self.off[i + 1] = self.off[i] + len(s.verts)
        self.verts = np.ascontiguousarray(np.concatenate([s.verts for s in shapes], axis=0))
        self.bs_c = np.ascontiguousarray(np.stack([s.bs_center for s in shapes]))
        self.bs_r = np.ascontiguousarray(np.array([s.bs_radius for s in shapes]))
        self.margins = np.array([s.margin for s in shapes])
        nA = len(self.arm.hulls)
        self.self_pairs = list(self.arm.self_pairs)
        self.fixed_pairs = [(a, nA + f) for f in range(len(self.fixed)) for a in range(nA)]

    def distances(self, Q: np.ndarray, pairs: Sequence[Tuple[int, int]], far: float = 0.05) -> np.ndarray:
        """Signed Bullet distance (GJK - both margins) of each pair [B, NP]; pairs further than `far` report a lower
        bound.  Intersecting cores report -(margin_A + margin_B) (the depth is not computed: no EPA)."""
        Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
        B = Q.shape[0]
        S = len(self.shapes)
        frames = np.zeros((B, S, 12))
        frames[:, :len(self.arm.hulls)] = self.arm.hull_frames(Q)
        for f, (_, fr) in enumerate(self.fixed):
            frames[:, len(self.arm.hulls) + f] = fr[None]
        frames = np.ascontiguousarray(frames)
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        out = np.zeros((B, len(pr)))
        _load().hull_gjk_batch(B, S, _ptr(self.off), _ptr(self.verts), _ptr(self.bs_c), _ptr(self.bs_r), _ptr(frames),
                               len(pr), _ptr(pr), float(far), _ptr(out))
        m = self.margins[pr[:, 0]] + self.margins[pr[:, 1]]
        return out - m[None]

    def rigid(self, Q: np.ndarray) -> dict:
        """limits | self pairs | moving links x fixed bodies  (pybullet_tools/utils.py:3979-4016)."""
        Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
        lim = self.arm.limits_violated(Q)
        sd = self.distances(Q, self.self_pairs)
        fd = self.distances(Q, self.fixed_pairs)
        hit = lim | np.any(sd <= 0.0, axis=1) | np.any(fd <= 0.0, axis=1)
        clear = np.minimum(sd.min(axis=1), fd.min(axis=1))
        return {"hit": hit, "limits": lim, "self_dist": sd, "fixed_dist": fd, "clearance": clear}
