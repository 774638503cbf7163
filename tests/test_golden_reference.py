"""Oracle and host mirror vs golden vectors produced by running the REFERENCE'S OWN CODE
(tests/golden/gen_reference_vectors.py) and vs the reference's saved seed-1 plant."""
import json
import os
import random

import numpy as np
import pytest

from oracle import edgecheck_oracle as O
from plant_motion_planning_b200 import planner_fns as PF
from plant_motion_planning_b200.model.plants import (plant_from_parsed, sample_plant, sample_world,
                                                     single_stem_world)
from plant_motion_planning_b200.model.urdf import parse_urdf
from plant_motion_planning_b200.robots import load_iiwa14

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    with open(os.path.join(G, name)) as fp:
        return json.load(fp)


class _Robot:
    def __init__(self, d):
        self.lower = np.array([v if np.isfinite(v) else 0.0 for v in d["lower"]])
        self.upper = np.array([v if np.isfinite(v) else 0.0 for v in d["upper"]])
        self.circular = np.array(d["circular"], dtype=bool)
        self.dof = len(d["lower"])


@pytest.fixture(scope="module")
def planner_golden():
    return _load("reference_planner_fns.json")


@pytest.mark.parametrize("which", ["iiwa14", "circular"])
def test_planner_fns_match_reference(planner_golden, which):
    d = planner_golden[which]
    rob = _Robot(d)
    model = load_iiwa14()
    if which == "iiwa14":
        assert np.allclose(model.lower, d["lower"]) and np.allclose(model.upper, d["upper"])
        assert not model.circular.any()
    else:
        model.circular = rob.circular.copy()
    diff, dist, ext = PF.get_difference_fn(rob), PF.get_distance_fn(rob), PF.get_extend_fn(rob)
    res = d["resolution"]
    assert res == PF.DEFAULT_RESOLUTION
    for c in d["cases"]:
        q1, q2 = tuple(c["q1"]), tuple(c["q2"])
        # host mirror: bit-identical to the reference (same float64 operations)
        assert list(diff(q2, q1)) == c["difference"]
        assert float(dist(q1, q2)) == c["distance"]
        fwd = [list(q) for q in ext(q1, q2)]
        bwd = [list(q) for q in PF.asymmetric_extend(q1, q2, ext, backward=True)]
        for seq, ref in ((fwd, c["extend_compact"]), (bwd, c["extend_backward_compact"])):
            assert len(seq) == ref["n"]
            assert seq[:2] == ref["head"] and seq[-2:] == ref["tail"]
            assert np.allclose(np.sum(np.asarray(seq), axis=0), ref["sum"], rtol=0, atol=1e-9)
        if c["extend"] is not None:
            assert fwd == c["extend"]
        # oracle restatement
        assert np.allclose(O.difference(model, q2, q1), c["difference"], atol=1e-15)
        assert abs(O.distance(model, q1, q2) - c["distance"]) < 1e-14
        assert O.num_steps(model, q1, q2, res) == c["extend_compact"]["n"]
        of = O.extend(model, q1, q2, res)
        assert np.allclose(np.asarray(of), np.asarray(fwd), atol=1e-14)
        ob = O.asymmetric_extend(model, q1, q2, res, backward=True)
        assert np.allclose(np.asarray(ob), np.asarray(bwd), atol=1e-14)
        # closed forms the kernels evaluate
        assert np.allclose(O.edge_configs(model, q1, q2, res, False), np.asarray(fwd), atol=1e-12)
        # backward on a circular joint: the reference walks from q2 by the wrapped difference, so its values can
        # differ from the closed form by a multiple of 2 pi (the same joint angle)
        eb = O.edge_configs(model, q1, q2, res, True) - np.asarray(bwd)
        eb = np.where(model.circular[None], (eb + np.pi) % (2 * np.pi) - np.pi, eb)
        assert np.abs(eb).max() < 1e-12
    lim = PF.get_limits_fn(rob)
    for c in d["limits"]:
        assert bool(lim(tuple(c["q"]))) == c["violated"]
        assert bool(O.limits_violated(model, np.asarray(c["q"]))[0]) == c["violated"]
    np.random.seed(d["sample_seed"])
    sfn = PF.get_sample_fn(rob)
    assert [list(sfn()) for _ in range(5)] == d["samples"]
    np.random.seed(d["sample_seed"])
    ofn = O.sample_fn(model)
    assert np.allclose([list(ofn()) for _ in range(5)], d["samples"], atol=1e-15)


def test_self_collision_link_pairs_match_reference(planner_golden):
    """compile_robot's pair table vs get_self_link_pairs / get_moving_links run on the same URDF tree."""
    d = planner_golden["iiwa14"]
    model = load_iiwa14(collision="urdf_spheres")
    assert d["moving_links"] == [1, 2, 3, 4, 5, 6, 7, 8, 9]
    pairs = set()
    A = model.num_spheres
    for a in range(A):
        for b in range(a + 1, A):
            if model.self_pair_mask[a, b]:
                pairs.add(tuple(sorted((int(model.sph_link[a]), int(model.sph_link[b])))))
    for prim in model.base_prims:
        for a in range(A):
            if (prim.sphere_mask >> a) & 1:
                pairs.add(tuple(sorted((0, int(model.sph_link[a])))))
    ref = {tuple(p) for p in d["self_link_pairs"]}
    # links 8, 9 (ee frames) carry no geometry -> never in the reference list either
    assert pairs == ref
    # one *sphere* pair is pruned at compile time (always 14-15 mm deep in the primitive model): link 5's distal
    # sphere vs link 7's sphere; the link pair itself stays checked through link 5's other sphere
    assert not model.self_pair_mask[8, 11] and model.self_pair_mask[9, 11]
    assert int(model.self_pair_mask.sum()) == 38
    # the URDF the scripts load (mesh collision on links 1-7, none on link 0): the reference's get_self_link_pairs gives
    # the 15 non-adjacent pairs of links 1..7, and so do the hull pairs / sphere masks of the default model
    e = planner_golden["iiwa14_polytope_edit"]
    hm = load_iiwa14()
    ref = {tuple(p) for p in e["self_link_pairs"]}
    assert e["moving_links"] == [1, 2, 3, 4, 5, 6, 7, 8, 9] and len(ref) == 15
    link_of_hull = [hm.link_names.index(n) for n in hm.hulls.link_names]
    assert {tuple(sorted((link_of_hull[a], link_of_hull[b]))) for a, b in hm.hulls.pairs} == ref
    sp = set()
    for a in range(hm.num_spheres):
        for b in range(a + 1, hm.num_spheres):
            if hm.self_pair_mask[a, b]:
                sp.add(tuple(sorted((int(hm.sph_link[a]), int(hm.sph_link[b])))))
    assert sp == ref and not hm.base_prims


def test_sampler_matches_reference_run():
    for c in _load("reference_sampler.json"):
        nr, pr = np.random.RandomState(c["seed"]), random.Random(c["seed"])
        w = sample_plant(*c["args"], c["base_pos_xy"], np_random=nr, py_random=pr)
        assert w.base_pos == c["base_pos"]
        assert w.joint_list == c["link_parent_indices_arg"]
        assert w.main_stem_indices == c["main_stem_indices"]
        assert {str(k): v for k, v in w.base_points.items()} == c["base_points"]
        assert {str(k): v for k, v in w.link_parent_indices.items()} == c["link_parent_indices"]
        hh = [c["stem_half_height"][str(2 * k + 2)] for k in range(w.num_stems)]
        assert [s.half_height for s in w.stems] == hh
        base = np.array(c["base_pos"])
        for k, s in enumerate(w.stems):
            pos = np.array(c["link_positions"][2 * k]) + (base if k == 0 else 0.0)
            assert np.array_equal(s.origin.t, pos)
            from plant_motion_planning_b200.model.transforms import quat_to_matrix
            assert np.allclose(s.origin.R, quat_to_matrix(c["link_orientations"][2 * k]), atol=1e-15)
        # the RNG streams are left exactly where the reference leaves them
        assert float(nr.uniform()) == c["next_np_uniform"]
        assert pr.random() == c["next_py_random"]


def test_sampler_reproduces_saved_seed1_plant():
    """models/plant.urdf + plant_params.pkl == np.random.seed(1); random.seed(1), limits of scripts/steps_test.py:55,
    stem_base_spacing 0.05."""
    d = _load("seed1_plant.json")
    w = sample_world(1, 2, 1, ((0, 0.35), (0.15, 0.50)), stem_base_spacing=0.05,
                     np_random=np.random.RandomState(1), py_random=random.Random(1))
    assert w.joint_list == d["joint_list"] and w.main_stem_indices == d["main_stem_indices"]
    assert w.base_pos == d["base_pos"]
    assert {str(k): v for k, v in w.base_points.items()} == d["base_points"]
    tree = parse_urdf(d["urdf"], is_text=True)
    assert len(tree.joints) == 2 * w.num_stems
    for k, s in enumerate(w.stems):
        j = tree.joints[2 * k]
        if k > 0:
            assert np.abs(s.origin.t - j.origin.t).max() < 6e-6       # URDF prints 5 decimals
            assert np.abs(s.origin.R - j.origin.R).max() < 2e-5
        box = [g for g in tree.links[tree.joints[2 * k + 1].child].collisions if g.kind == "box"][0]
        assert abs(2 * s.half_height - box.size[2]) < 6e-6
        assert abs(box.origin.t[2] - (0.05 + s.half_height)) < 6e-6


def _observe(world, theta):
    Rs, ts = O.frames_from_theta(world, np.asarray(theta)[None])
    return O.chain_deflections(world, O.raw_deflections(world, Rs, ts))[0]


def test_observers_match_reference_classes():
    d = _load("reference_observers.json")
    for g in d["mod3"]:
        w = sample_world(*g["args"], g["xy_limits"], np_random=np.random.RandomState(g["seed"]),
                         py_random=random.Random(g["seed"]))
        for c in g["cases"]:
            got = _observe(w, c["theta"])
            assert np.allclose(got, c["deflections"], atol=1e-9), (g["seed"], got, c["deflections"])
            assert bool((got > 0.30).any()) == c["exceeded_0.30"]
    fx = _load("plant_fixtures.json")
    for g in d["mod"]:
        f = fx[g["env"]]
        w = plant_from_parsed(parse_urdf(f["urdf"], is_text=True), f["joint_list"], f["main_stem_indices"],
                              {int(k): v for k, v in f["base_points"].items()}, f["base_pos"])
        for c in g["cases"]:
            got = _observe(w, c["theta"])
            assert np.allclose(got, c["deflections"], atol=1e-7), (g["env"], got, c["deflections"])
            assert bool((got > 0.30).any()) == c["exceeded_0.30"]
    for g in d["single"]:
        w = single_stem_world(g["positions"])
        for c in g["cases"]:
            Rs, ts = O.frames_from_theta(w, np.asarray(c["theta"])[None])
            got = O.raw_deflections(w, Rs, ts)[0]
            assert np.allclose(got, c["deflections"], atol=1e-9)


def test_gate_replay_matches_reference_closure():
    """The 5 % pass-through (pybullet_tools/utils.py:3843): same results and same number of draws."""
    d = _load("reference_gate.json")
    for c in d["cases"]:
        np.random.seed(c["seed"])
        for s in c["sequence"]:
            rigid = np.array([[1 if (s["limit"] or s["rigid"]) else 0]], dtype=np.uint32)
            exc = np.array([[1 if s["exceeded"] else 0]], dtype=np.uint32)
            hit = PF.replay_gate(1, rigid, exc, d["gate_probability"]) == 0
            assert hit == s["result"]
        assert float(np.random.rand()) == c["next_uniform_after"]


def test_fk_matches_urdf_composition_with_reference_transformations():
    """iiwa14 link frames composed with the reference's pybullet_tools/transformations.py (no shared code)."""
    d = _load("reference_fk.json")
    model = load_iiwa14()
    assert d["links"] == model.link_names
    Q = np.array([c["q"] for c in d["cases"]])
    pos, quat, com = O.link_poses(model, Q)
    rp = np.array([c["link_pos"] for c in d["cases"]])
    rq = np.array([c["link_quat"] for c in d["cases"]])
    rc = np.array([c["com_pos"] for c in d["cases"]])
    sg = np.sign((quat * rq).sum(-1, keepdims=True))
    assert np.abs(pos - rp).max() < 1e-12 and np.abs(quat * sg - rq).max() < 1e-12 and np.abs(com - rc).max() < 1e-12
    # end-effector point of the cost term = COM of PyBullet link 9 (iiwa_link_ee)
    assert np.abs(O.ee_positions(model, Q) - rc[:, 9]).max() < 1e-12
