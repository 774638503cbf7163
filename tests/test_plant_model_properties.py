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
from plant_motion_planning_b200.model.tables import stem_capsule


@pytest.fixture(scope="module")
def stem():
    w = single_stem_world([(0.0, 0.0)])
    s = w.stems[0]
    z0, z1, rs = stem_capsule(s, None)
    return w, s, z0, z1, rs


def _penetration(world, stem_rec, theta, centre, radius):
    """Penetration depth of one sphere into the stem capsule at given angles (independent of the solver code path:
    brute-force closest point on the segment by dense sampling)."""
    s, z0, z1, rs = stem_rec
    R, t = O.frames_from_theta(world, np.asarray(theta, dtype=np.float64).reshape(1, 1, 2))
    u = R[0, 0][:, 2]
    zs = np.linspace(z0, z1, 1201)
    pts = t[0, 0][None] + zs[:, None] * u[None]
    return radius + rs - np.min(np.linalg.norm(pts - centre[None], axis=1))


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
    d = r + rs - depth                                         # centre distance from the stem axis at rest
    centre = s.origin.t + np.array([d * np.cos(angle), d * np.sin(angle), height])
    th, rest = _solve(w, centre, r)
    assert abs(rest - depth) < 1e-9
    residual = _penetration(w, rec, th, centre, r)
    assert residual <= max(0.06 * depth, 2e-4)                 # K = 4 damped steps leave at most a few per cent
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
    assert best * 0.98 - 1e-4 <= norm <= best * 1.05 + 1e-4, (norm, best)


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
