"""Properties the quasi-static plant model of DESIGN.md section 4 claims, checked on the float64 oracle (the CUDA path
is held to the oracle by tests/test_gpu_parity.py):

  * a stem that nothing touches stays exactly at rest;
  * one contact: the K-step solve removes the penetration (to a small residual) and the angles it finds are, to a
    few per cent, the SMALLEST joint motion that does so -- the frictionless minimum-spring-energy response (the
    reference's springs are isotropic in (theta_x, theta_y): kp 500 on both, env.py:127-144) -- found here
    independently by a brute-force search over the angle plane;
  * the response is continuous in the arm position, through contact onset included;
  * mirror symmetry.
"""
import numpy as np
import pytest

from oracle import edgecheck_oracle as O
from plant_motion_planning_b200.model.plants import single_stem_world
MARGIN = 1.0e-3


@pytest.fixture(scope="module")
def stem():
    w = single_stem_world([(0.0, 0.0)])
    s = w.stems[0]
    hw, hh, zc = O.stem_box(s, MARGIN)
    rs = s.half_width                       # distance of a face from the stem axis (core + margin)
    return w, s, (hw, hh, zc), None, rs


def _penetration(world, stem_rec, theta, centre, radius):
    """Penetration depth of one sphere into the stem's box (core + margin) at given angles, independent of the solver
    code path: the box is sampled as a dense point cloud on the surface of its core."""
    s, (hw, hh, zc), _, rs = stem_rec
    R, t = O.frames_from_theta(world, np.asarray(theta, dtype=np.float64).reshape(1, 1, 2))
    local = R[0, 0].T @ (centre - t[0, 0]) - np.array([0.0, 0.0, zc])
    # nearest point of the core box to `local`, searched over a surface grid around the sphere's height
    zs = np.clip(local[2] + np.linspace(-0.08, 0.08, 321), -hh, hh)
    us = np.linspace(-hw, hw, 49)
    best = np.inf
    for face in ((0, hw), (0, -hw), (1, hw), (1, -hw)):
        g = np.zeros((len(us), len(zs), 3))
        g[:, :, 2] = zs[None, :]
        g[:, :, face[0]] = face[1]
        g[:, :, 1 - face[0]] = us[:, None]
        best = min(best, float(np.min(np.linalg.norm(g - local[None, None], axis=2))))
    return radius + MARGIN - best


def _solve(world, centre, radius, iterations=4):
    prm = O.OracleParams(iterations=iterations)
    theta, _, _, rest = O.plant_equilibrium(world, centre.reshape(1, 1, 3), np.array([radius]), prm)
    return theta[0, 0], rest[0]


def test_untouched_stem_stays_at_rest(stem):
    w, s, z0, z1, rs = stem
    base = s.origin.t
    for off in (0.2, 0.09, rs + 0.05 + 1e-6):
        th, rest = _solve(w, base + np.array([off, 0.0, 0.4]), 0.05)
        assert np.all(th == 0.0) and rest < 0.0


@pytest.mark.parametrize("height", [0.25, 0.5, 0.9])
@pytest.mark.parametrize("depth", [0.005, 0.02, 0.05])
@pytest.mark.parametrize("angle", [0.0, 0.7, 2.3, 4.0])
def test_single_contact_is_resolved_with_minimum_joint_motion(stem, height, depth, angle):
    w, s, z0, z1, rs = stem
    rec = (s, z0, z1, rs)
    r = 0.05
    # centre distance from the stem axis at which the sphere is `depth` deep in the box at rest (direction dependent:
    # faces at 25 mm, edges further out), by bisection on the oracle's own rest penetration
    lo_d, hi_d = 0.0, 0.2
    for _ in range(60):
        d = 0.5 * (lo_d + hi_d)
        c_try = s.origin.t + np.array([d * np.cos(angle), d * np.sin(angle), height])
        if _solve(w, c_try, r)[1] > depth:
            lo_d = d
        else:
            hi_d = d
    centre = s.origin.t + np.array([d * np.cos(angle), d * np.sin(angle), height])
    th, rest = _solve(w, centre, r)
    assert abs(rest - depth) < 1e-9
    assert abs(_penetration(w, rec, np.zeros(2), centre, r) - depth) < 2e-4      # the two penetration codes agree
    residual = _penetration(w, rec, th, centre, r)
    assert residual <= max(0.06 * depth, 4e-4)                 # K = 4 damped steps leave at most a few per cent
    # brute force: smallest |theta| on a polar grid whose penetration is <= the residual the solver left
    best = np.inf
    for phi in np.linspace(0.0, 2 * np.pi, 91)[:-1]:
        lo, hi = 0.0, 0.6
        dirv = np.array([np.cos(phi), np.sin(phi)])
        if _penetration(w, rec, hi * dirv, centre, r) > residual:
            continue                                           # pushing this way never clears the sphere
        for _ in range(20):
            mid = 0.5 * (lo + hi)
            if _penetration(w, rec, mid * dirv, centre, r) > residual:
                lo = mid
            else:
                hi = mid
        best = min(best, hi)
    assert np.isfinite(best)
    norm = float(np.linalg.norm(th))
    assert best * 0.97 - 2e-4 <= norm <= best * 1.06 + 2e-4, (norm, best)


def test_response_is_continuous_through_contact_onset(stem):
    w, s, z0, z1, rs = stem
    r = 0.05
    xs = np.linspace(r + rs + 0.004, r + rs - 0.03, 400)        # sphere sliding into the stem
    th = np.array([_solve(w, s.origin.t + np.array([x, 0.01, 0.6]), r)[0] for x in xs])
    assert np.all(th[xs >= r + rs + 1e-9] == 0.0)
    steps = np.linalg.norm(np.diff(th, axis=0), axis=1)
    assert steps.max() <= 6.0 * np.abs(np.diff(xs)).max() / 0.6 + 1e-6      # ~ (d theta / d x) = 1 / lever arm, no jumps
    assert np.linalg.norm(th[-1]) > 0.03


def test_mirror_symmetry(stem):
    w, s, z0, z1, rs = stem
    r = 0.05
    c = np.array([0.06, 0.025, 0.7])
    a, _ = _solve(w, s.origin.t + c, r)
    b, _ = _solve(w, s.origin.t + c * np.array([1.0, -1.0, 1.0]), r)     # mirrored in the x-z plane
    Ra, _ = O.frames_from_theta(w, a.reshape(1, 1, 2))
    Rb, _ = O.frames_from_theta(w, b.reshape(1, 1, 2))
    ua, ub = Ra[0, 0][:, 2], Rb[0, 0][:, 2]
    assert np.allclose(ua * np.array([1.0, -1.0, 1.0]), ub, atol=1e-12)  # the stem axes are mirror images


def test_two_contacts_blend_continuously(stem):
    """No arg-max over spheres: as a second sphere is withdrawn its influence fades without a jump, and once it no
    longer touches the stem the answer is the single-contact one."""
    w, s, z0, z1, rs = stem
    r = 0.05
    prm = O.OracleParams()
    c1 = s.origin.t + np.array([r + rs - 0.02, 0.0, 0.7])
    ys = np.linspace(r + rs - 0.02, r + rs + 0.01, 300)
    th = []
    for y in ys:
        c2 = s.origin.t + np.array([0.0, y, 0.5])
        t, _, _, _ = O.plant_equilibrium(w, np.stack([c1, c2])[None], np.array([r, r]), prm)
        th.append(t[0, 0])
    th = np.array(th)
    single, _ = _solve(w, c1, r)
    assert np.allclose(th[ys > r + rs + 1e-9], single, atol=1e-12)
    assert np.linalg.norm(np.diff(th, axis=0), axis=1).max() <= 2e-3          # 0.1 mm steps of the second sphere
    assert abs(th[0, 0]) > 0.01 and abs(th[0, 1]) > 0.01                      # both contacts bend it at the start


def test_pushed_child_does_not_load_its_parent_but_a_pushed_parent_carries_its_children():
    import random
    from plant_motion_planning_b200.model.plants import sample_world
    np.random.seed(1)
    random.seed(1)
    w = sample_world(1, 2, 1, ((0.3, 0.3), (0.0, 0.0)))
    prm = O.OracleParams(gravity_sag=False)        # rest pose = zero angles: isolates the contact response
    R0, t0 = O.frames_from_theta(w, np.zeros((1, w.num_stems, 2)))
    child = next(k for k, st in enumerate(w.stems) if st.parent == 0)
    rs = w.stems[child].half_width
    # a sphere pressed into the middle of the child stem, from the side
    mid = t0[0, child] + R0[0, child] @ np.array([0.0, 0.0, w.stems[child].half_height])
    side = R0[0, child] @ np.array([1.0, 0.0, 0.0])
    centre = mid + side * (0.05 + rs - 0.015)
    far = np.linalg.norm(centre - (t0[0, 0] + R0[0, 0] @ np.array([0.0, 0.0, w.stems[0].half_height])))
    theta, R, t, _ = O.plant_equilibrium(w, centre.reshape(1, 1, 3), np.array([0.05]), prm)
    if far > 0.05 + rs + 0.01:                       # the sphere does not also touch the main stem
        assert np.all(theta[0, 0] == 0.0)
    assert np.linalg.norm(theta[0, child]) > 1e-3
    # push the main stem instead: its children keep theta = 0 but their world frames move with it
    base_mid = t0[0, 0] + R0[0, 0] @ np.array([0.0, 0.0, 0.6 * w.stems[0].half_height])
    c2 = base_mid + np.array([0.05 + rs - 0.01, 0.0, 0.0])
    theta2, R2, t2, _ = O.plant_equilibrium(w, c2.reshape(1, 1, 3), np.array([0.05]), prm)
    assert np.linalg.norm(theta2[0, 0]) > 1e-3
    moved = [k for k, st in enumerate(w.stems) if st.parent == 0 and np.all(theta2[0, k] == 0.0)]
    assert moved and all(np.abs(R2[0, k] - R0[0, k]).max() > 1e-4 for k in moved)
