"""The C twin of the oracle (float64 C on the packed float32 tables) against the numpy oracle (float64 on the
float64 model).  Differences come only from the float32 rounding of the tables."""
import numpy as np
import pytest

from oracle import edgecheck_oracle as O
from oracle.c_oracle import COracle
from plant_motion_planning_b200.model.plants import single_stem_world
from plant_motion_planning_b200.robots import load_iiwa14
from plant_motion_planning_b200.workloads import random_edges, synthetic_worlds

MODES = {"our_method": 0, "avoid_all": 1, "ignore_all": 2}


@pytest.fixture(scope="module")
def setup():
    model = load_iiwa14()
    worlds = synthetic_worlds(6)
    rng = np.random.RandomState(0)
    Q = rng.uniform(model.lower, model.upper, size=(768, 7)).astype(np.float32)
    return model, worlds, Q


@pytest.mark.parametrize("mode", ["our_method", "avoid_all", "ignore_all"])
def test_configs_agree(setup, mode):
    model, worlds, Q = setup
    c = COracle(model, worlds, mode=mode).check_configs(Q)
    prm = O.OracleParams(mode=MODES[mode])
    r = O.check_configs(model, worlds, Q.astype(np.float64), prm)
    rigid, exc = (c["flags"] & 1) > 0, (c["flags"] & 2) > 0
    # decisions may only differ inside an eps band of a boundary
    eps = 1e-5
    c_lo = O.rigid_hit(model, O.default_obstacles(), Q.astype(np.float64), inflate=-eps)
    c_hi = O.rigid_hit(model, O.default_obstacles(), Q.astype(np.float64), inflate=+eps)
    robust = (c_lo == c_hi)
    if mode != "avoid_all":
        assert np.array_equal(rigid[robust], r.rigid[robust])
    assert (rigid != r.rigid).mean() < 2e-3
    near = np.any(np.abs(r.deflection - prm.deflection_limit) < 1e-3, axis=2)
    assert np.array_equal(exc[~near], r.exceeded[~near])
    err = np.abs(c["deflection"] - r.deflection)
    assert np.percentile(err, 99.9) < 1e-4
    assert (err > 1e-3).mean() < 1e-3
    assert np.abs(c["ee"] - r.ee).max() < 1e-6


def test_edges_agree(setup):
    model, worlds, _ = setup
    q0, q1 = random_edges(model, 24, seed=5)
    co = COracle(model, worlds)
    for backward in (False, True):
        c = co.validate_edges(q0, q1, backward=backward, want_masks=True)
        for e in range(q0.shape[0]):
            r = O.validate_edge(model, worlds, q0[e].astype(np.float64), q1[e].astype(np.float64), O.OracleParams(),
                                backward=backward)
            assert r.n_steps == c["n_steps"][e]
            assert r.first_hit == c["first_hit"][e]
            assert np.abs(r.max_deflection - c["max_deflection"][e]).max() < 2e-3
            assert abs(r.ee_len - c["ee_len"][e]) < 1e-5
            for w in range(len(worlds)):
                for s in range(r.n_steps):
                    assert bool(c["rigid_mask"][e, w, s >> 5] >> (s & 31) & 1) == bool(r.rigid[s, w])


def test_single_stem_worlds(setup):
    """Forest of root stems (y-first joint order, vertical-reference observer with the 1e-5 snap)."""
    model, _, Q = setup
    worlds = [single_stem_world([[0.4, 0.1], [0.12, -0.15], [-0.25, 0.60], [-0.25, 0.45], [-0.05, 0.50]]),
              single_stem_world([[0.15, -0.1], [-0.08, -0.15], [0.25, -0.60]])]
    c = COracle(model, worlds).check_configs(Q[:256])
    r = O.check_configs(model, worlds, Q[:256].astype(np.float64), O.OracleParams())
    assert np.array_equal((c["flags"] & 1) > 0, r.rigid)
    err = np.abs(c["deflection"] - r.deflection)
    assert np.percentile(err, 99.5) < 1e-4
    assert (r.deflection > 0).any()


def test_threads_do_not_change_results(setup):
    model, worlds, _ = setup
    q0, q1 = random_edges(model, 64, seed=9)
    co = COracle(model, worlds)
    a = co.validate_edges(q0, q1, n_threads=1, want_masks=True)
    b = co.validate_edges(q0, q1, n_threads=4, want_masks=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
