"""The convex-hull collision oracle (oracle/hull_oracle.py, oracle/hull_gjk.c) and the sphere model fitted to the hulls.

  * GJK distance vs an independent QP solution of the same problem (random polytopes) and vs an LP feasibility test of
    hull intersection on the real link hulls;
  * invariances of the distance (rigid motion of the world, vertex order, float32 rounding of the vertices);
  * the product's compiled tables (18 fitted spheres + hulls, model/robot.py) against the independent hull oracle, whose
    forward kinematics are composed from the URDF text: same clearances, covering spheres really cover (no hull contact
    outside the broad phase), stated +-20 mm deviation of the fitted spheres;
  * the script's own, unmodified goal configuration is free on the hulls (0.39 mm of Bullet clearance).
"""
import numpy as np
import pytest
from scipy.optimize import linprog, minimize

from oracle import edgecheck_oracle as O
from oracle import hull_oracle as H
from plant_motion_planning_b200 import load_iiwa14
from plant_motion_planning_b200.model.tables import default_obstacles
from plant_motion_planning_b200.scenarios import REFERENCE_GOAL, START

I12 = np.concatenate([np.eye(3).ravel(), np.zeros(3)])


def _qp_distance(A, B, rng):
    nA, nB = len(A), len(B)

    def f(x):
        d = x[:nA] @ A - x[nA:] @ B
        return d @ d

    def g(x):
        d = x[:nA] @ A - x[nA:] @ B
        return np.concatenate([2 * A @ d, -2 * B @ d])
    cons = [{"type": "eq", "fun": lambda x: x[:nA].sum() - 1}, {"type": "eq", "fun": lambda x: x[nA:].sum() - 1}]
    best = np.inf
    for _ in range(3):
        x0 = rng.random(nA + nB)
        x0[:nA] /= x0[:nA].sum()
        x0[nA:] /= x0[nA:].sum()
        r = minimize(f, x0, jac=g, bounds=[(0, 1)] * (nA + nB), constraints=cons, method="SLSQP",
                     options={"ftol": 1e-16, "maxiter": 500})
        best = min(best, np.sqrt(max(r.fun, 0.0)))
    return best


def test_gjk_equals_qp_on_random_polytopes():
    rng = np.random.default_rng(0)
    for _ in range(60):
        A = rng.normal(size=(rng.integers(1, 12), 3)) * 0.1
        B = rng.normal(size=(rng.integers(1, 12), 3)) * 0.1 + rng.normal(size=3) * 0.15
        assert abs(H.gjk_pair(A, I12, B, I12) - _qp_distance(A, B, rng)) < 1e-7


def _hulls_intersect_lp(VA, VB):
    """conv(VA) and conv(VB) share a point <=> the LP {lam, mu >= 0, sum = 1, lam VA = mu VB} is feasible."""
    nA, nB = len(VA), len(VB)
    Aeq = np.zeros((5, nA + nB))
    Aeq[:3, :nA] = VA.T
    Aeq[:3, nA:] = -VB.T
    Aeq[3, :nA] = 1.0
    Aeq[4, nA:] = 1.0
    r = linprog(np.zeros(nA + nB), A_eq=Aeq, b_eq=[0, 0, 0, 1, 1], bounds=(0, None), method="highs")
    return r.status == 0


def test_gjk_intersection_agrees_with_lp_on_link_hulls():
    arm = H.HullArm()
    rng = np.random.default_rng(5)
    Q = rng.uniform(arm.lower, arm.upper, size=(4000, 7))
    frames = arm.hull_frames(Q)
    sc = H.HullScene(arm)
    d = sc.distances(Q, arm.self_pairs, far=10.0) + 2 * H.BULLET_MARGIN          # margin-free distances
    # pick pairs near and inside contact and check the sign of the distance against the LP
    idx = np.argwhere(d < 0.01)
    pick = idx[rng.choice(len(idx), size=min(120, len(idx)), replace=False)]
    n_touch = 0
    for b, p in pick:
        h1, h2 = arm.self_pairs[p]
        def posed(h):
            F = frames[b, h]
            return arm.hulls[h].verts @ F[:9].reshape(3, 3).T + F[9:]
        lp = _hulls_intersect_lp(posed(h1), posed(h2))
        if d[b, p] > 1e-7:
            assert not lp
        elif d[b, p] == 0.0:
            n_touch += 1
            assert lp
    assert n_touch >= 10, "the sample must contain intersecting hull pairs"


def test_gjk_invariances():
    arm = H.HullArm()
    rng = np.random.default_rng(1)
    Q = rng.uniform(arm.lower, arm.upper, size=(3000, 7))
    sc = H.HullScene(arm)
    pairs = sc.self_pairs + sc.fixed_pairs
    d0 = sc.distances(Q, pairs, far=10.0)
    # float32 rounding of the vertices (what the product ships) moves distances by <= 1e-7
    for s in sc.arm.hulls:
        s.verts = np.ascontiguousarray(s.verts.astype(np.float32).astype(np.float64))
    sc._pack()
    assert np.abs(sc.distances(Q, pairs, far=10.0) - d0).max() < 2e-7
    # reversed vertex order
    for s in sc.arm.hulls:
        s.verts = np.ascontiguousarray(s.verts[::-1])
    sc._pack()
    assert np.abs(sc.distances(Q, pairs, far=10.0) - d0).max() < 2e-7


def test_product_tables_agree_with_the_independent_hull_oracle():
    m = load_iiwa14()
    rng = np.random.default_rng(3)
    Q = rng.uniform(m.lower, m.upper, size=(20000, 7))
    ours = O.hull_clearance(m, default_obstacles(), Q)           # product's compiled chain + hull copies
    ref = H.HullScene().rigid(Q)["clearance"]                    # URDF text + extracted hulls
    assert np.abs(np.minimum(ours, 0.05) - np.minimum(ref, 0.05)).max() < 2e-7
    assert np.array_equal(ours <= 0.0, ref <= 0.0)
    assert 0.15 < float(np.mean(ref <= 0.0)) < 0.25


def test_covering_spheres_cover_and_fitted_spheres_stay_within_20mm():
    m = load_iiwa14()
    rng = np.random.default_rng(4)
    Q = rng.uniform(m.lower, m.upper, size=(20000, 7))
    c = O.sphere_centers(m, Q)
    hm = m.hulls
    near = np.zeros(len(Q), dtype=bool)
    hit_nominal = np.zeros(len(Q), dtype=bool)
    A = m.num_spheres
    for a in range(A):
        for b in range(a + 1, A):
            if m.self_pair_mask[a, b]:
                dist = np.linalg.norm(c[:, a] - c[:, b], axis=1)
                near |= dist <= hm.sph_radius_outer[a] + hm.sph_radius_outer[b]
                hit_nominal |= dist <= m.sph_radius[a] + m.sph_radius[b]
    for prim in default_obstacles():
        for a in range(A):
            near |= O._prim_hit(prim, c[:, a], hm.sph_radius_outer[a])
            hit_nominal |= O._prim_hit(prim, c[:, a], m.sph_radius[a])
    clear = O.hull_clearance(m, default_obstacles(), Q)
    hull_hit = clear <= 0.0
    assert not np.any(hull_hit & ~near), "a hull contact outside the covering spheres: the broad phase would miss it"
    # the nominal spheres (the plant contact model) decide differently from the hulls only inside a +-41 mm band of
    # hull clearance (two surfaces, each fitted to +-20.5 mm)
    diff = hull_hit != hit_nominal
    assert np.all(np.abs(clear[diff & (clear > -0.0019)]) < 0.042)
    assert float(np.mean(diff)) < 0.12
    assert np.all(hm.sph_radius_outer - m.sph_radius < 0.0225)


def test_reference_goal_is_free_on_the_hulls():
    m = load_iiwa14()
    clear = O.hull_clearance(m, default_obstacles(), np.array([REFERENCE_GOAL, START]))
    assert 3.0e-4 < clear[0] < 4.5e-4          # 0.39 mm: flange against the block
    assert clear[1] > 0.02
    sc = H.HullScene()
    r = sc.rigid(np.array([REFERENCE_GOAL]))
    assert not r["hit"][0] and abs(r["clearance"][0] - clear[0]) < 1e-7
    assert abs(r["self_dist"][0].min() - 5.2e-4) < 5e-5      # links 5 and 7


def test_script_seed19_start_touches_the_fourth_plant():
    """multi_world_avoid_all.py's own seed: at the arm's zero configuration a branch of the fourth plant is 1.3 mm from
    link 4 -- inside the two 1 mm Bullet margins -- on the exact geometry (hull vs stem box), so with avoid_all the
    start is a collision; the sphere model of the kernels agrees (it is not an artefact of the fitted spheres)."""
    from oracle.c_oracle import COracle
    from plant_motion_planning_b200.scenarios import script_worlds
    worlds, cfg = script_worlds("multi_world_avoid_all")
    m = load_iiwa14()
    fl = COracle(m, worlds, mode="avoid_all").check_configs(np.array([START], dtype=np.float32))["flags"][0]
    assert fl.tolist() == [0, 0, 0, 1]
    w = worlds[3]
    th0 = O.rest_theta(w, O.OracleParams(gravity_sag=False))
    R, t = O.frames_from_theta(w, th0[None])
    boxes = []
    for s, st in enumerate(w.stems):
        c = R[0, s] @ np.array([0.0, 0.0, st.z_offset + st.half_height]) + t[0, s]
        boxes.append((H.Shape.bullet_box([st.half_width, st.half_width, st.half_height]), H.frame12(R[0, s], c)))
    sc = H.HullScene(H.HullArm(), boxes)
    d = sc.distances(np.array([START]), sc.fixed_pairs, far=1.0)
    assert -0.001 < d.min() < -0.0005          # 0.7 mm inside the margins, even before gravity sag


@pytest.mark.gpu
def test_gpu_rigid_flags_equal_the_independent_hull_oracle_at_scale():
    """50 000 random configurations: the CUDA hull pre-pass (float32 GJK, clearance tables, covering / inscribed balls)
    against the independent float64 oracle with its own forward kinematics -- identical outside the oracle's 1e-5 m
    decision band (tools/rigid_parity_report.py runs the same comparison on 2 x 10^6 configurations:
    profiles/r02_rigid_parity.json, one mismatch, 3e-7 m from the boundary)."""
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.workloads import synthetic_worlds
    model = load_iiwa14()
    chk = EdgeChecker(model, synthetic_worlds(1), EdgeCheckConfig(mode="ignore_all"))
    Q = np.random.RandomState(7).uniform(model.lower, model.upper, size=(50000, 7)).astype(np.float32)
    gpu = (chk.check_configs(torch.from_numpy(Q), detail=False).flags.cpu().numpy()[:, 0] & 1) > 0
    ref = H.HullScene().rigid(Q.astype(np.float64))
    mis = gpu != ref["hit"]
    assert not np.any(mis & (np.abs(ref["clearance"]) > 1e-5))
    assert mis.sum() <= 5 and 0.1 < ref["hit"].mean() < 0.3
