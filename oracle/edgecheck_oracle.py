"""CPU oracle (float64 numpy) for the edge-validation hot path.   *** TEST INFRASTRUCTURE ***

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (plant_motion_planning_b200/) never does.

PARITY STATUS: **partially pinned**.
  pinned by reference artefacts / reference code executed in the build container:
    * plant sampler + RNG order ........ models/plant.urdf + models/plant_params.pkl (tests/golden/seed1_plant.json)
    * interpolation / metric / limits .. the reference's own get_extend_fn / get_distance_fn /
      get_difference_fn / get_sample_fn / get_limits_fn, imported with a stub ``pybullet``
      (tests/golden/gen_reference_vectors.py -> tests/golden/reference_planner_fns.json)
    * observer arithmetic .............. the reference's representation.py classes run on link poses
      supplied by the stub (same script -> tests/golden/reference_observers.json)
    * quaternion / Euler conventions ... reference's pybullet_tools/transformations.py (same script)
  NOT PINNED AGAINST THE ENGINE (no reference artefact carries these values and PyBullet 3.2.0 cannot be installed or
  run here):
    * collision booleans: decided on the reference's own geometry -- the convex hulls of the iiwa collision meshes,
      floor, block, Bullet's margins -- by GJK distance - margins <= 0 (hull_rigid_hit below, oracle/hull_oracle.py with
      its own forward kinematics, oracle/hull_gjk.c); the procedure is pinned against an independent QP / LP solution
      (tests/test_hull_oracle.py), not against Bullet's btGjkPairDetector itself
    * plant response (50 Featherstone sub-steps per step) -> restated as the quasi-static model below; its distance
      from a time-stepped restatement (oracle/dynamic_oracle.py) is MEASURED and enforced (tests/test_dynamic_gap.py):
      <= 0.05 rad and identical decisions on contact-free paths, median 0.05 / p90 0.2-0.3 rad one-sided in contact
  See DESIGN.md "Oracle" for the full statement.

Each function cites the reference lines it restates (paths relative to
/root/reference/plant_motion_planning/ unless noted).

Quasi-static plant model (replaces env.py:127-163 -- spring torques + 50 x stepSimulation):
  every stem s has two angles th = (tx, ty), |th| <= limit, pose  T_s = T_parent . origin_s . Rx(tx) . Ry(ty);
  its collision body is the reference's box (rand_gen_plants_programmatically_main.py:30-51): half extents (w, w, h)
  centred zc = z_offset + h up the stem's z axis, with Bullet's 1 mm margin carved out of it (core half extents
  w - m, h - m; the margin is added back as a radius, so faces stay where the reference puts them and edges are
  rounded by m).  Stems are visited parents first.  A stem starts from its REST pose th0 -- the static balance of
  the reference's springs and gravity without contact (stems are 10 kg; the reference's observers capture their
  reference pose at th = 0, so the sag counts as deflection) -- computed by oracle/dynamic_oracle.static_rest_pose.
  If no arm sphere overlaps the box in that pose the stem stays there; otherwise K damped Gauss-Newton steps on the
  arm-sphere penetrations are applied.  In the stem's link frame (R_s = R_v Rx Ry, joint origin o):
     for every arm sphere a:  pf = R_s^T (c_a - o)  (relative to the joint),  p = pf - zc e_z  (relative to the box
         centre),  q = clamp(p, +-half),  d = p - q,  dist = |d|,  pen_a = max(r_a + m - dist, 0),
         J_a = d(dist)/d(th) = ((cy, 0, sy) . (d x pf), (0, 1, 0) . (d x pf)) / dist
     H = sum_a pen_a J_a J_a^T,  g = sum_a pen_a^2 J_a,  S = sum_a pen_a,  lam = eps S
     dth = adj(H + lam I) g / max(det(H + lam I), lam (lam + tr H))   (skipped when S = 0; the max is an identity in
     exact arithmetic, it guards the float32 cancellation), each component clamped to +-max_step,
     th <- clamp(th + dth, +-limit)
  i.e. a least-squares removal of the penetrations weighted by their depth: for a single contact it is the
  minimum-norm joint motion that removes the penetration to first order, and it is continuous in the arm
  configuration (no arg-max over spheres), which is what makes float32-vs-float64 parity meaningful.  Children see
  their parent's final pose; a pushed child does not load its parent (measured gap to the time-stepped restatement:
  DESIGN.md section 4, tools/dynamic_gap_report.py).

Arm lag (optional, OracleParams.arm_lag_gain > 0): the reference drives the arm with a position motor that closes
1 % of the joint error per sub-step (env.py:154,170) for 50 sub-steps per interpolation step, after teleporting it to
the first step of an edge (env.py:222-227, primitives.py:220), so the collision / deflection tests of step i see
    a_0 = q_0,  a_i = a_{i-1} + c (q_i - a_{i-1}),  c = 1 - (1 - gain)^substeps
while the joint-limit test sees q_i itself (pybullet_tools/utils.py:3980).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from plant_motion_planning_b200.model.plants import OBS_ANGLE, OBS_MOD3, PlantWorld
from plant_motion_planning_b200.model.robot import PRIM_BOX, PRIM_CYLINDER, PRIM_SPHERE, FixedPrim, RobotModel
from plant_motion_planning_b200.model.tables import default_obstacles
from plant_motion_planning_b200.model.transforms import matrix_to_quat

MODE_OUR_METHOD, MODE_AVOID_ALL, MODE_IGNORE_ALL = 0, 1, 2


@dataclass
class OracleParams:
    resolution: float = math.radians(3)      # DEFAULT_RESOLUTION, pybullet_tools/utils.py:3660
    deflection_limit: float = 0.30
    iterations: int = 4
    max_step: float = 0.5
    reg_eps: float = 1e-4
    stem_margin: float = 1.0e-3               # Bullet's collision margin of the stem boxes (carved out of the box)
    gravity_sag: bool = True                  # stems rest at their spring/gravity balance, not at th = 0
    arm_lag_gain: float = 0.0                 # per-sub-step position gain of the arm motor (reference: 0.01); 0 = no lag
    arm_lag_substeps: int = 50                # env.py:157, 216
    alpha: float = 0.1                        # kuka_primitives.py:476
    mode: int = MODE_OUR_METHOD


def stem_box(stem, margin: float):
    """(core half width, core half height, zc): the stem's box with Bullet's margin carved out
    (rand_gen_plants_programmatically_main.py:30-37: halfExtents (w, w, h), collisionFramePosition z = spacing + h)."""
    return (max(stem.half_width - margin, 0.0), max(stem.half_height - margin, 0.0), stem.z_offset + stem.half_height)


def rest_theta(world: PlantWorld, prm: "OracleParams") -> np.ndarray:
    """Rest pose [P, 2] of the world's stems (zeros without gravity sag).  Cached on the world object."""
    if not prm.gravity_sag or world.num_stems == 0:
        return np.zeros((world.num_stems, 2))
    cached = getattr(world, "_oracle_rest_theta", None)
    if cached is None:
        from oracle.dynamic_oracle import static_rest_pose
        cached = static_rest_pose(world)
        world._oracle_rest_theta = cached
    return cached


def lag_closure(gain: float, substeps: int) -> float:
    """Fraction of the joint error the arm motor closes during one interpolation step: 1 - (1 - gain)^substeps."""
    return 1.0 - (1.0 - gain) ** substeps if gain > 0.0 else 1.0


def lagged_configs(Q: np.ndarray, prm: "OracleParams") -> np.ndarray:
    """Configurations the arm actually has at the steps of one edge: a_0 = q_0, a_i = a_{i-1} + c (q_i - a_{i-1})."""
    Q = np.asarray(Q, dtype=np.float64)
    c = lag_closure(prm.arm_lag_gain, prm.arm_lag_substeps)
    if c >= 1.0 or Q.shape[0] == 0:
        return Q.copy()
    out = Q.copy()
    for i in range(1, Q.shape[0]):
        out[i] = out[i - 1] + c * (Q[i] - out[i - 1])
    return out


# ----------------------------------------------------------------------------- planner callables
def wrap_pi(x):
    """wrap_interval(x, [-pi, pi))  (pybullet_tools/utils.py:1666-1669, 1679-1686)."""
    return (x + math.pi) % (2.0 * math.pi) - math.pi


def difference(model: RobotModel, q2, q1) -> np.ndarray:
    """get_difference_fn (pybullet_tools/utils.py:3604-3610): q2 - q1, wrapped on circular joints."""
    q2 = np.asarray(q2, dtype=np.float64)
    q1 = np.asarray(q1, dtype=np.float64)
    d = q2 - q1
    return np.where(model.circular, wrap_pi(d), d)


def distance(model: RobotModel, q1, q2) -> float:
    """get_distance_fn, weights 1, norm 2 (pybullet_tools/utils.py:3619-3627)."""
    d = difference(model, q2, q1)
    return float(np.sqrt(np.dot(np.ones_like(d), d * d)))


def num_steps(model: RobotModel, q1, q2, resolution: float) -> int:
    """steps + 1 of get_extend_fn (pybullet_tools/utils.py:3667-3676)."""
    d = difference(model, q2, q1) / resolution
    s = 0.0
    for v in d:                     # fixed summation order (the CUDA path reproduces it bit for bit)
        s = s + v * v
    return int(math.sqrt(s)) + 1


def extend(model: RobotModel, q1, q2, resolution: float) -> List[Tuple[float, ...]]:
    """get_extend_fn + get_refine_fn (pybullet_tools/utils.py:3641-3651, 3667-3676):
    n = int(|dq / res|_2) + 1 configurations, q <- q + (q2 - q)/(n - i); q1 excluded, last == q2."""
    n = num_steps(model, q1, q2, resolution)
    out = []
    q = tuple(float(v) for v in q1)
    for i in range(n):
        q = tuple((1.0 / (n - i)) * difference(model, q2, q) + np.asarray(q))
        out.append(tuple(float(v) for v in q))
    return out


def asymmetric_extend(model: RobotModel, q1, q2, resolution: float, backward: bool):
    """motion_planners/motion_planners/primitives.py:16-19."""
    if backward:
        return list(reversed(extend(model, q2, q1, resolution)))
    return extend(model, q1, q2, resolution)


def edge_configs(model: RobotModel, q1, q2, resolution: float, backward: bool = False) -> np.ndarray:
    """Closed form of the two sequences above (what the kernels evaluate):
    forward  q_i = q1 + (i+1)/n * d,  i = 0..n-1 ;  backward  q_i = q1 + i/n * d  (starts at q1, stops short of q2)."""
    n = num_steps(model, q1, q2, resolution)
    d = difference(model, q2, q1)
    off = 0 if backward else 1
    return np.stack([np.asarray(q1, dtype=np.float64) + d * ((i + off) / n) for i in range(n)])


def limits_violated(model: RobotModel, Q: np.ndarray, inflate: float = 0.0) -> np.ndarray:
    """get_limits_fn (pybullet_tools/utils.py:3753-3762); circular joints are unbounded (:2128-2137).
    ``inflate`` (tests only) moves the decision boundary: > 0 makes violations more likely."""
    Q = np.atleast_2d(Q)
    lo = np.where(model.circular, -np.inf, model.lower + inflate)
    hi = np.where(model.circular, np.inf, model.upper - inflate)
    return ~(np.all(lo <= Q, axis=1) & np.all(Q <= hi, axis=1))


def sample_fn(model: RobotModel, np_random=np.random):
    """get_sample_fn (pybullet_tools/utils.py:3549-3551, 3586-3598): one uniform(size=D) draw per sample."""
    lo = np.where(model.circular, -math.pi, model.lower)
    hi = np.where(model.circular, math.pi, model.upper)

    def fn():
        w = np_random.uniform(size=model.dof)
        return tuple((1 - w) * lo + w * hi)
    return fn


# ----------------------------------------------------------------------------- forward kinematics
def _rodrigues(axis: np.ndarray, ang: np.ndarray) -> np.ndarray:
    a = axis / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    s = np.sin(ang)[:, None, None]
    c = np.cos(ang)[:, None, None]
    return np.eye(3)[None] + s * K[None] + (1 - c) * (K @ K)[None]


def joint_frames(model: RobotModel, Q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """World frames of the D joint (== riding link) frames:  T_j = T_{j-1} . origin_j . Rot(axis_j, q_j).
    (Implicit in PyBullet's resetJointState/getLinkState; URDF semantics)"""
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    B, D = Q.shape
    R = np.zeros((B, D, 3, 3))
    t = np.zeros((B, D, 3))
    Rc = np.broadcast_to(model.base.R, (B, 3, 3)).copy()
    tc = np.broadcast_to(model.base.t, (B, 3)).copy()
    for j in range(D):
        tc = tc + np.einsum("bij,j->bi", Rc, model.joint_t0[j])
        Rc = Rc @ model.joint_R0[j][None] @ _rodrigues(model.joint_axis[j], Q[:, j])
        R[:, j], t[:, j] = Rc, tc
    return R, t


def link_poses(model: RobotModel, Q: np.ndarray):
    """(link frame position [B,L,3], link frame quaternion xyzw [B,L,4], COM position [B,L,3]) --
    getLinkState [4],[5] and [0] (pybullet_tools/utils.py:2197-2203)."""
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    R, t = joint_frames(model, Q)
    B = Q.shape[0]
    L = model.num_links
    pos = np.zeros((B, L, 3))
    quat = np.zeros((B, L, 4))
    com = np.zeros((B, L, 3))
    for l in range(L):
        j = int(model.link_joint[l])
        if j < 0:
            Rj = np.broadcast_to(model.base.R, (B, 3, 3))
            tj = np.broadcast_to(model.base.t, (B, 3))
        else:
            Rj, tj = R[:, j], t[:, j]
        Rl = Rj @ model.link_R[l][None]
        pl = np.einsum("bij,j->bi", Rj, model.link_t[l]) + tj
        pos[:, l] = pl
        com[:, l] = np.einsum("bij,j->bi", Rl, model.link_com[l]) + pl
        for b in range(B):
            quat[b, l] = matrix_to_quat(Rl[b])
    return pos, quat, com


def sphere_centers(model: RobotModel, Q: np.ndarray) -> np.ndarray:
    R, t = joint_frames(model, Q)
    A = model.num_spheres
    out = np.zeros((R.shape[0], A, 3))
    for a in range(A):
        j = int(model.sph_joint[a])
        out[:, a] = np.einsum("bij,j->bi", R[:, j], model.sph_center[a]) + t[:, j]
    return out


def ee_positions(model: RobotModel, Q: np.ndarray) -> np.ndarray:
    """COM of the end-effector link (getLinkState(robot, 9)[0], kuka_primitives.py:478,504)."""
    R, t = joint_frames(model, Q)
    j = model.ee_joint
    return np.einsum("bij,j->bi", R[:, j], model.ee_local) + t[:, j]


# ----------------------------------------------------------------------------- rigid collision
def _prim_hit(prim: FixedPrim, c: np.ndarray, r) -> np.ndarray:
    """Sphere (centre c [B,3], radius r) vs static primitive: signed distance <= 0
    (getClosestPoints(distance=0) non-empty, pybullet_tools/utils.py:3390-3415)."""
    pl = (c - prim.t[None]) @ prim.R          # R^T (c - t)
    if prim.kind == PRIM_BOX:
        q = np.clip(pl, -prim.half, prim.half)
        d2 = np.sum((pl - q) ** 2, axis=1)
    elif prim.kind == PRIM_CYLINDER:
        rho = np.sqrt(pl[:, 0] ** 2 + pl[:, 1] ** 2)
        dr = np.maximum(rho - prim.half[0], 0.0)
        dz = np.maximum(np.abs(pl[:, 2]) - prim.half[1], 0.0)
        d2 = dr * dr + dz * dz
    else:
        rr = max(r + prim.half[0], 0.0)
        return np.sum(pl * pl, axis=1) <= rr * rr
    r = max(r, 0.0)
    return d2 <= r * r


def rigid_hit(model: RobotModel, obstacles: Sequence[FixedPrim], Q: np.ndarray,
              centers: Optional[np.ndarray] = None, inflate: float = 0.0,
              Q_limits: Optional[np.ndarray] = None) -> np.ndarray:
    """get_collision_fn_with_controls_v3 (pybullet_tools/utils.py:3965-4016), world-independent part:
    joint limits, self pairs, moving links x fixed bodies.
    ``inflate`` (tests only): every sphere radius grows by it and the limits shrink by it, so
    rigid_hit(+eps) != rigid_hit(-eps) marks configurations within eps of a decision boundary."""
    Q = np.atleast_2d(Q)
    if getattr(model, "hulls", None) is not None and inflate == 0.0:
        return limits_violated(model, Q if Q_limits is None else np.atleast_2d(Q_limits)) | \
            hull_rigid_hit(model, obstacles, Q)
    c = sphere_centers(model, Q) if centers is None else centers
    # the limit test sees the commanded configuration, the geometry the (possibly lagged) simulated one
    hit = limits_violated(model, Q if Q_limits is None else np.atleast_2d(Q_limits), inflate)
    A = model.num_spheres
    rad = model.sph_radius + inflate
    for a in range(A):
        for b in range(a + 1, A):
            if model.self_pair_mask[a, b]:
                rr = rad[a] + rad[b]
                hit |= np.sum((c[:, a] - c[:, b]) ** 2, axis=1) <= rr * rr
    for prim in model.base_prims:
        for a in range(A):
            if (prim.sphere_mask >> a) & 1:
                hit |= _prim_hit(prim, c[:, a], rad[a])
    for prim in obstacles:
        for a in range(A):
            hit |= _prim_hit(prim, c[:, a], rad[a])
    return hit


def hull_clearance(model: RobotModel, obstacles: Sequence[FixedPrim], Q: np.ndarray) -> np.ndarray:
    """Smallest signed Bullet distance [B] over the self pairs and the (hull, obstacle) pairs of a model that carries
    the convex hulls of its collision meshes: GJK distance of the margin-free shapes minus both margins
    (pybullet_tools/utils.py:3390-3441 on btConvexHullShape / btBoxShape / btSphereShape; see oracle/hull_oracle.py).
    Intersecting cores report -(margin_a + margin_b)."""
    from oracle import hull_oracle as H
    hm = model.hulls
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    R, t = joint_frames(model, Q)
    B = Q.shape[0]
    shapes = [H.Shape(v, m, n) for v, m, n in zip(hm.verts, hm.margin, hm.link_names)]
    nH = len(shapes)
    frames = np.zeros((B, nH + len(obstacles), 12))
    for k in range(nH):
        j = int(hm.joint[k])
        frames[:, k, :9] = R[:, j].reshape(B, 9)
        frames[:, k, 9:] = t[:, j]
    for f, prim in enumerate(obstacles):
        if prim.kind == PRIM_BOX:
            shapes.append(H.Shape.bullet_box(prim.half, hm.fixed_margin, prim.name))
        elif prim.kind == PRIM_SPHERE:
            shapes.append(H.Shape(np.zeros((1, 3)), float(prim.half[0]), prim.name))
        else:
            raise NotImplementedError("exact hull test supports box and sphere obstacles")
        frames[:, nH + f] = H.frame12(prim.R, prim.t)[None]
    pairs = [tuple(int(v) for v in p) for p in hm.pairs] + [(h, nH + f) for f in range(len(obstacles)) for h in range(nH)]
    d = H.batch_distances(shapes, frames, pairs)
    return d.min(axis=1) if d.shape[1] else np.full(B, np.inf)


def hull_rigid_hit(model: RobotModel, obstacles: Sequence[FixedPrim], Q: np.ndarray) -> np.ndarray:
    return hull_clearance(model, obstacles, Q) <= 0.0


# ----------------------------------------------------------------------------- plants
def _rx_ry(tx: np.ndarray, ty: np.ndarray) -> np.ndarray:
    cx, sx, cy, sy = np.cos(tx), np.sin(tx), np.cos(ty), np.sin(ty)
    M = np.zeros(tx.shape + (3, 3))
    M[..., 0, 0], M[..., 0, 1], M[..., 0, 2] = cy, 0.0, sy
    M[..., 1, 0], M[..., 1, 1], M[..., 1, 2] = sx * sy, cx, -sx * cy
    M[..., 2, 0], M[..., 2, 1], M[..., 2, 2] = -cx * sy, sx, cx * cy
    return M


def plant_equilibrium(world: PlantWorld, centers: np.ndarray, radii: np.ndarray, prm: OracleParams,
                      solve: bool = True):
    """Quasi-static stand-in for MultiPlantWorld.step (env.py:127-163, 212-219); see module docstring.

    Returns theta [B,P,2], R [B,P,3,3], t [B,P,3] (final stem link frames) and rest_pen [B] =
    deepest sphere-vs-box penetration in the rest pose (> 0 <=> the arm touches the undeflected plant).
    """
    B = centers.shape[0]
    P = world.num_stems
    theta = np.zeros((B, P, 2))
    Rs = np.zeros((B, P, 3, 3))
    ts = np.zeros((B, P, 3))
    rest_pen = np.full(B, -np.inf)
    th0 = rest_theta(world, prm)
    m = prm.stem_margin
    for s, stem in enumerate(world.stems):
        if stem.parent < 0:
            Rp = np.broadcast_to(np.eye(3), (B, 3, 3))
            tp = np.zeros((B, 3))
        else:
            Rp, tp = Rs[:, stem.parent], ts[:, stem.parent]
        Rv = Rp @ stem.origin.R[None]
        tv = np.einsum("bij,j->bi", Rp, stem.origin.t) + tp
        hw, hh, zc = stem_box(stem, m)
        half = np.array([hw, hw, hh])
        tx = np.full(B, th0[s, 0])
        ty = np.full(B, th0[s, 1])
        touch = np.zeros(B, dtype=bool)
        iters = prm.iterations if solve else 1
        for it in range(iters):
            cy, sy = np.cos(ty), np.sin(ty)
            Rl = Rv @ _rx_ry(tx, ty)
            Hxx = np.zeros(B)
            Hxy = np.zeros(B)
            Hyy = np.zeros(B)
            gx = np.zeros(B)
            gy = np.zeros(B)
            S = np.zeros(B)
            raw_max = np.full(B, -np.inf)
            for a in range(centers.shape[1]):
                pf = np.einsum("bji,bj->bi", Rl, centers[:, a] - tv)        # R_l^T (c - o)
                p = pf.copy()
                p[:, 2] -= zc
                q = np.clip(p, -half, half)
                d = p - q
                dist = np.sqrt(np.sum(d * d, axis=1))
                raw = radii[a] + m - dist
                raw_max = np.maximum(raw_max, raw)
                pen = np.maximum(raw, 0.0)
                mm = np.cross(d, pf)
                inv = 1.0 / np.maximum(dist, 1e-12)
                Jx = inv * (cy * mm[:, 0] + sy * mm[:, 2])
                Jy = inv * mm[:, 1]
                Hxx += pen * Jx * Jx
                Hxy += pen * Jx * Jy
                Hyy += pen * Jy * Jy
                gx += pen * pen * Jx
                gy += pen * pen * Jy
                S += pen
            if it == 0:
                # rest_pen is only used by the avoid_all test, where no stem is ever deflected
                rest_pen = np.maximum(rest_pen, raw_max)
                touch = raw_max > 0.0       # a stem that nothing touches in its initial pose is never pushed
            if not solve:
                break
            lam = prm.reg_eps * S
            a11 = Hxx + lam
            a22 = Hyy + lam
            act = touch & (S > 0.0)
            # det(H + lam I) = lam^2 + lam tr(H) + det(H) >= lam (lam + tr H); the bound only matters in float32
            det = np.where(act, np.maximum(a11 * a22 - Hxy * Hxy, lam * (lam + Hxx + Hyy)), 1.0)
            dx = np.where(act, (a22 * gx - Hxy * gy) / det, 0.0)
            dy = np.where(act, (a11 * gy - Hxy * gx) / det, 0.0)
            dx = np.clip(dx, -prm.max_step, prm.max_step)
            dy = np.clip(dy, -prm.max_step, prm.max_step)
            tx = np.where(act, np.clip(tx + dx, -stem.limit, stem.limit), tx)
            ty = np.where(act, np.clip(ty + dy, -stem.limit, stem.limit), ty)
        theta[:, s, 0], theta[:, s, 1] = tx, ty
        Rs[:, s] = Rv @ _rx_ry(tx, ty)
        ts[:, s] = tv
    return theta, Rs, ts, rest_pen


def frames_from_theta(world: PlantWorld, theta: np.ndarray):
    """Stem link frames (R [B,P,3,3], t [B,P,3]) for given stem angles theta [B,P,2] (URDF/createMultiBody
    composition: child joint frame = parent link frame . origin; joints Rx then Ry)."""
    theta = np.asarray(theta, dtype=np.float64)
    B, P = theta.shape[0], world.num_stems
    Rs = np.zeros((B, P, 3, 3))
    ts = np.zeros((B, P, 3))
    for s, stem in enumerate(world.stems):
        if stem.parent < 0:
            Rp = np.broadcast_to(np.eye(3), (B, 3, 3))
            tp = np.zeros((B, 3))
        else:
            Rp, tp = Rs[:, stem.parent], ts[:, stem.parent]
        Rs[:, s] = Rp @ stem.origin.R[None] @ _rx_ry(theta[:, s, 0], theta[:, s, 1])
        ts[:, s] = np.einsum("bij,j->bi", Rp, stem.origin.t) + tp
    return Rs, ts


def quat_difference(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """pybullet.getDifferenceQuaternion(a, b) = nearest(b) * inverse(a)   (Bullet b3GetQuaternionDifference)."""
    b = np.asarray(b, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    if np.dot(a - b, a - b) >= np.dot(a + b, a + b):
        b = -b
    ai = np.array([-a[0], -a[1], -a[2], a[3]])
    return quat_mul(b, ai)


def quat_mul(p: np.ndarray, q: np.ndarray) -> np.ndarray:
    px, py, pz, pw = p
    qx, qy, qz, qw = q
    return np.array([
        pw * qx + px * qw + py * qz - pz * qy,
        pw * qy + py * qw + pz * qx - px * qz,
        pw * qz + pz * qw + px * qy - py * qx,
        pw * qw - px * qx - py * qy - pz * qz,
    ])


def euler_from_quat(q: np.ndarray) -> Tuple[float, float, float]:
    """pybullet.getEulerFromQuaternion (Bullet btQuaternion::getEulerZYX with the +-0.99999 gimbal branch)."""
    x, y, z, w = (float(v) for v in q)
    sqx, sqy, sqz, sqw = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return 0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y)
    if sarg >= 0.99999:
        return 0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y)
    return (math.atan2(2.0 * (y * z + w * x), sqw - sqx - sqy + sqz), math.asin(sarg),
            math.atan2(2.0 * (x * y + w * z), sqw + sqx - sqy - sqz))


def _rest_frames(world: PlantWorld):
    R, t = [], []
    for s in world.stems:
        if s.parent < 0:
            R.append(s.origin.R.copy())
            t.append(s.origin.t.copy())
        else:
            R.append(R[s.parent] @ s.origin.R)
            t.append(R[s.parent] @ s.origin.t + t[s.parent])
    return R, t


def raw_deflections(world: PlantWorld, Rs: np.ndarray, ts: np.ndarray) -> np.ndarray:
    """Per-stem observer value before the chain rule.

    OBS_MOD3  : TwoAngleRepresentation_mod3 (representation.py:191-235)
    OBS_ANGLE : TwoAngleRepresentation (:27-51, vertical reference + 1e-5 snap) and
                TwoAngleRepresentation_mod (:60-123, rest direction reference)
    Link orientation/position are those of the COM frame (getLinkState [0],[1]); the stem's inertial
    frame has identity rotation and sits com_z up the stem (rand_gen...:265-266).
    """
    B, P = Rs.shape[0], world.num_stems
    out = np.zeros((B, P))
    R0, t0 = _rest_frames(world)
    for s, stem in enumerate(world.stems):
        if stem.obs_kind == OBS_MOD3:
            if stem.parent < 0:
                init = matrix_to_quat(R0[s])
            else:
                init = quat_difference(matrix_to_quat(R0[s]), matrix_to_quat(R0[stem.parent]))
            for b in range(B):
                ql = matrix_to_quat(Rs[b, s])
                if stem.parent >= 0:
                    ql = quat_difference(ql, matrix_to_quat(Rs[b, stem.parent]))
                d = quat_difference(ql, init)
                r, p, _ = euler_from_quat(d / np.linalg.norm(d))
                out[b, s] = math.sqrt(r * r + p * p)
        elif stem.obs_kind == OBS_ANGLE:
            com0 = R0[s] @ np.array([0.0, 0.0, stem.com_z]) + t0[s]
            com = np.einsum("bij,j->bi", Rs[:, s], np.array([0.0, 0.0, stem.com_z])) + ts[:, s]
            cur = com - stem.obs_ref_point[None]
            if stem.obs_vertical:
                x, y, z = cur[:, 0].copy(), cur[:, 1].copy(), cur[:, 2]
                if stem.obs_snap:
                    x[np.abs(x) < 1e-5] = 0.0
                    y[np.abs(y) < 1e-5] = 0.0
                r = np.sqrt(x * x + y * y + z * z)
                out[:, s] = np.arccos(z / r)
            else:
                init = com0 - stem.obs_ref_point
                cosang = (cur @ init) / (np.linalg.norm(init) * np.linalg.norm(cur, axis=1))
                out[:, s] = np.arccos(np.clip(cosang, -1.0, 1.0))
        else:
            raise ValueError(stem.obs_kind)
    return out


def chain_deflections(world: PlantWorld, raw: np.ndarray) -> np.ndarray:
    """CharacterizePlant(2).observe_all (representation.py:268-294, 354-380): subtract the last
    main-stem value, clamp at 0, main stems update the running value (with the *subtracted* value)."""
    out = np.zeros_like(raw)
    last = np.zeros(raw.shape[0])
    for s, stem in enumerate(world.stems):
        if stem.plant_start:
            last = np.zeros(raw.shape[0])
        d = np.maximum(raw[:, s] - last, 0.0)
        out[:, s] = d
        if stem.is_main:
            last = d
    return out


# ----------------------------------------------------------------------------- the hot path
@dataclass
class ConfigResult:
    rigid: np.ndarray          # bool [B,W]   limits | self | fixed (| plant at rest in avoid_all)
    exceeded: np.ndarray       # bool [B,W]   any stem deflection > limit
    deflection: np.ndarray     # [B,W,Pmax]   chained deflections (0 padded)
    theta: np.ndarray          # [B,W,Pmax,2]
    ee: np.ndarray             # [B,3]


def check_configs(model: RobotModel, worlds: Sequence[PlantWorld], Q: np.ndarray, prm: OracleParams,
                  obstacles: Optional[Sequence[FixedPrim]] = None, Q_limits: Optional[np.ndarray] = None) -> ConfigResult:
    """One (configuration x world) evaluation = collision_fn after env.step
    (env.py:212-238; pybullet_tools/utils.py:3827-3857) without the 5 % random pass-through.
    ``Q_limits``: the COMMANDED configurations for the joint-limit test when ``Q`` holds the lagged ones the
    geometry is in (pybullet_tools/utils.py:3980 tests q, everything else the simulation state)."""
    obstacles = default_obstacles() if obstacles is None else obstacles
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    B, W = Q.shape[0], len(worlds)
    Pmax = max(1, max(w.num_stems for w in worlds))
    c = sphere_centers(model, Q)
    base_hit = rigid_hit(model, obstacles, Q, c, Q_limits=Q_limits)
    rigid = np.repeat(base_hit[:, None], W, axis=1)
    exceeded = np.zeros((B, W), dtype=bool)
    defl = np.zeros((B, W, Pmax))
    theta = np.zeros((B, W, Pmax, 2))
    for wi, world in enumerate(worlds):
        if prm.mode == MODE_IGNORE_ALL or world.num_stems == 0:
            continue
        solve = prm.mode == MODE_OUR_METHOD
        th, Rs, ts, rest_pen = plant_equilibrium(world, c, model.sph_radius, prm, solve=solve)
        if prm.mode == MODE_AVOID_ALL:
            # plant body joins `fixed` (env.py:75-77): stems at rest + the base box
            hit = rest_pen > 0.0
            for fr, half in world.base_boxes:
                prim = FixedPrim(PRIM_BOX, fr.R, fr.t, half)
                for a in range(model.num_spheres):
                    hit |= _prim_hit(prim, c[:, a], model.sph_radius[a])
            rigid[:, wi] |= hit
        d = chain_deflections(world, raw_deflections(world, Rs, ts))
        P = world.num_stems
        defl[:, wi, :P] = d
        theta[:, wi, :P] = th
        exceeded[:, wi] = np.any(d > prm.deflection_limit, axis=1)
    return ConfigResult(rigid, exceeded, defl, theta, ee_positions(model, Q))


@dataclass
class EdgeResult:
    n_steps: int
    first_hit: int             # first step with any (rigid | exceeded) over worlds; n_steps if none
    rigid: np.ndarray          # bool [N,W]
    exceeded: np.ndarray       # bool [N,W]
    max_deflection: np.ndarray # [W]  max over steps and stems
    over_limit: np.ndarray     # [W]  sum over steps and stems of (deflection - limit)+
    ee_len: float              # sum |p_ee(q_i) - p_ee(q_{i-1})|, q_{-1} = q1
    cost: float                # ee_len + alpha * mean_w over_limit[w]


def validate_edge(model: RobotModel, worlds: Sequence[PlantWorld], q1, q2, prm: OracleParams,
                  backward: bool = False, obstacles=None) -> EdgeResult:
    """One tree extension / shortcut check (primitives.py:196-255; rrt_connect_..._v2.py:143-198) and
    its cost terms (kuka_primitives.py:467-521), evaluated for *all* steps (no early exit)."""
    q1f = np.asarray(q1, dtype=np.float64)
    q2f = np.asarray(q2, dtype=np.float64)
    Q = edge_configs(model, q1f, q2f, prm.resolution, backward)
    if prm.arm_lag_gain > 0.0:
        res = check_configs(model, worlds, lagged_configs(Q, prm), prm, obstacles, Q_limits=Q)
    else:
        res = check_configs(model, worlds, Q, prm, obstacles)
    n = Q.shape[0]
    viol = np.any(res.rigid | res.exceeded, axis=1)
    first = int(np.argmax(viol)) if viol.any() else n
    ee0 = ee_positions(model, q1f[None])[0]
    pts = np.concatenate([ee0[None], res.ee], axis=0)
    ee_len = float(np.sum(np.linalg.norm(pts[1:] - pts[:-1], axis=1)))
    over = np.sum(np.maximum(res.deflection - prm.deflection_limit, 0.0), axis=(0, 2))
    mx = np.max(res.deflection, axis=(0, 2))
    return EdgeResult(n, first, res.rigid, res.exceeded, mx, over, ee_len,
                      ee_len + prm.alpha * float(np.mean(over)))


# ----------------------------------------------------------------------------- batched tree extension
def nearest(model: RobotModel, nodes: np.ndarray, targets: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """argmin over tree nodes of distance_fn(node, target), first minimum wins
    (motion_planners/motion_planners/utils.py:47-53 called from primitives.py:207; metric
    pybullet_tools/utils.py:3604-3627).  Returns (index [B], distance [B]).  The per-joint terms are summed in
    joint order (the CUDA kernel reproduces the float64 arithmetic bit for bit).  A NaN distance is never selected;
    if every distance is NaN the answer is node 0."""
    nodes = np.asarray(nodes, dtype=np.float64)
    targets = np.asarray(targets, dtype=np.float64)
    idx = np.zeros(targets.shape[0], dtype=np.int32)
    dist = np.full(targets.shape[0], np.inf)
    circ = np.asarray(model.circular, dtype=bool)
    for b in range(targets.shape[0]):
        d = targets[b][None, :] - nodes
        if circ.any():
            d = np.where(circ[None, :], wrap_pi(d), d)
        s = np.zeros(nodes.shape[0])
        for j in range(nodes.shape[1]):
            s = s + d[:, j] * d[:, j]
        r = np.sqrt(s)
        r = np.where(np.isnan(r), np.inf, r)
        k = int(np.argmin(r))
        if np.isfinite(r[k]):
            idx[b], dist[b] = k, r[k]
    return idx, dist


def edge_point(model: RobotModel, q1, q2, resolution: float, step: int, backward: bool = False) -> np.ndarray:
    """Configuration number `step` of the edge (0-based): edge_configs(...)[step] without building the sequence."""
    n = num_steps(model, q1, q2, resolution)
    d = difference(model, q2, q1)
    num = step + (0 if backward else 1)
    return np.asarray(q1, dtype=np.float64) + d * (num / n)


def tree_append(model: RobotModel, tree: dict, from_index: np.ndarray, targets: np.ndarray, first_hit: np.ndarray,
                n_steps: np.ndarray, resolution: float, max_steps: int) -> dict:
    """The bookkeeping of one batched extension given the per-edge results (first_hit, n_steps) of the edges
    tree.nodes[from_index[e]] -> targets[e]: every extension with at least one safe step appends its last safe
    configuration, in batch order (the batched form of primitives.py:243-254: `safe = steps[:first_hit]`, last
    node parented on the nearest node).  ``tree``: dict(nodes [n, D], parent [n], via [n, D], via_steps [n]) plus
    'capacity'; returns the outputs of pmp_tree_extend and mutates ``tree``."""
    B = targets.shape[0]
    new_index = np.full(B, -1, dtype=np.int32)
    reached = np.zeros(B, dtype=np.int32)
    new_cfg = np.full(targets.shape, np.nan, dtype=np.float64)
    appended = overflow = 0
    for e in range(B):
        if first_hit[e] < 1:
            continue
        slot = tree["nodes"].shape[0]
        if slot >= tree["capacity"]:
            overflow = 1
            continue
        q0 = tree["nodes"][from_index[e]]
        cfg = edge_point(model, q0, targets[e], resolution, int(first_hit[e]) - 1)
        if n_steps[e] <= max_steps and first_hit[e] >= n_steps[e]:
            reached[e] = 1
            cfg = np.where(model.circular, cfg, np.asarray(targets[e], dtype=np.float64))
        tree["nodes"] = np.concatenate([tree["nodes"], cfg[None]], axis=0)
        tree["parent"] = np.append(tree["parent"], np.int32(from_index[e]))
        tree["via"] = np.concatenate([tree["via"], np.asarray(targets[e], dtype=np.float64)[None]], axis=0)
        tree["via_steps"] = np.append(tree["via_steps"], np.int32(first_hit[e]))
        new_index[e] = slot
        new_cfg[e] = cfg
        appended += 1
    hits = np.nonzero(reached)[0]
    stats = np.array([appended, len(hits), hits[0] if len(hits) else B, overflow], dtype=np.int32)
    return {"new_index": new_index, "reached": reached, "new_config": new_cfg, "stats": stats}
