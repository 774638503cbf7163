"""Host-side logic that needs no GPU: gate replay, waypoint extraction, first_violation semantics."""
import numpy as np

from plant_motion_planning_b200 import birrt
from plant_motion_planning_b200 import planner_fns as PF


def test_waypoints_from_path_keeps_corners_only():
    path = [(0., 0.), (0.5, 0.), (0.5, 0.), (1., 0.), (1., 0.5), (1., 1.), (2., 2.)]
    assert birrt.waypoints_from_path(path) == [(0., 0.), (1., 0.), (1., 1.), (2., 2.)]
    assert birrt.waypoints_from_path([(0., 0.), (1., 1.)]) == [(0., 0.), (1., 1.)]


def test_first_violation_order_and_draws():
    flags = np.zeros((4, 3), np.uint8)
    flags[1, 2] = 2          # exceeded in the last world at step 1
    flags[2, 0] = 1          # rigid in the first world at step 2
    draws = iter([0.99, 0.5])
    assert birrt.first_violation(flags, True, rand=lambda: next(draws)) == 2     # step 1 passes the gate (0.99 >= 0.95)
    assert birrt.first_violation(flags, False) == 2                              # backward checker ignores plants
    flags[1, 0] = 1
    used = []
    assert birrt.first_violation(flags, True, rand=lambda: used.append(1) or 0.0) == 1 and not used   # rigid short-circuits


def test_replay_gate_equals_first_violation():
    rng = np.random.RandomState(0)
    for _ in range(50):
        n, W = int(rng.randint(1, 70)), int(rng.randint(1, 6))
        fl = (rng.rand(n, W) < 0.1).astype(np.uint8) | ((rng.rand(n, W) < 0.2).astype(np.uint8) << 1)
        C = (n + 31) // 32
        rm = np.zeros((W, C), np.uint32)
        xm = np.zeros((W, C), np.uint32)
        for s in range(n):
            for w in range(W):
                if fl[s, w] & 1:
                    rm[w, s >> 5] |= np.uint32(1) << np.uint32(s & 31)
                if fl[s, w] & 2:
                    xm[w, s >> 5] |= np.uint32(1) << np.uint32(s & 31)
        for fwd in (True, False):
            np.random.seed(5)
            a = PF.replay_gate(n, rm, xm, forward=fwd)
            ra = np.random.rand()
            np.random.seed(5)
            b = birrt.first_violation(fl, fwd)
            assert a == b and ra == np.random.rand()


def test_config_tree_nearest_equals_reference_argmin():
    from plant_motion_planning_b200.robots import load_iiwa14
    robot = load_iiwa14()
    dist = PF.get_distance_fn(robot)
    rng = np.random.RandomState(1)
    tree = birrt.ConfigTree(circular=robot.circular)
    plain = []
    for _ in range(300):
        n = birrt.TreeNode(tuple(rng.uniform(robot.lower, robot.upper)))
        tree.append(n)
        plain.append(n)
    assert len(tree) == 300 and tree[137] is plain[137]
    for _ in range(50):
        t = tuple(rng.uniform(robot.lower, robot.upper))
        scores = [dist(n.config, t) for n in plain]
        assert tree.nearest(t) is plain[scores.index(min(scores))]
    # circular joints wrap
    class R:  # noqa: D401
        lower = robot.lower; upper = robot.upper; dof = 7
        circular = np.array([False, False, True, False, False, False, False])
    d2 = PF.get_distance_fn(R)
    tree2 = birrt.ConfigTree(circular=R.circular)
    for n in plain:
        tree2.append(n)
    for _ in range(50):
        t = tuple(rng.uniform(robot.lower, robot.upper))
        scores = [d2(n.config, t) for n in plain]
        assert tree2.nearest(t) is plain[scores.index(min(scores))]


def test_plant_writer_reproduces_the_saved_fixtures(tmp_path):
    """save_plant writes what generate_random_plant(save_plant=True) left in environments/multi_branch/envK
    (plant_motion_planning/utils.py:270-279): every origin / box / axis / parent / child / mass line of the fixture
    URDFs is reproduced to the 5 printed decimals, the pickle carries the same four entries, and the pair loads back
    to bit-identical world tables."""
    import json
    import os
    import pickle
    import re
    from plant_motion_planning_b200.model.plants import load_plant, plant_from_parsed, save_plant
    from plant_motion_planning_b200.model.tables import pack_worlds
    from plant_motion_planning_b200.model.urdf import parse_urdf
    with open(os.path.join(os.path.dirname(__file__), "golden", "plant_fixtures.json")) as fp:
        fx = json.load(fp)
    num = lambda txt: np.array([float(x) for x in re.findall(r"-?\d+\.\d+", txt)])
    keep = lambda txt: [l.strip() for l in txt.splitlines()
                        if any(k in l for k in ("<origin", "<box", "<axis", "<parent", "<child", "<mass"))]
    assert len(fx) == 4
    for env, f in fx.items():
        bp = {int(k): v for k, v in f["base_points"].items()}
        w = plant_from_parsed(parse_urdf(f["urdf"], is_text=True), f["joint_list"], f["main_stem_indices"], bp, f["base_pos"])
        up, pp = str(tmp_path / f"{env}.urdf"), str(tmp_path / f"{env}.pkl")
        save_plant(w, up, pp)
        a, b = keep(f["urdf"]), keep(open(up).read())
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert re.sub(r"-?\d+\.\d+", "#", x) == re.sub(r"-?\d+\.\d+", "#", y)
            assert np.abs(num(x) - num(y)).max(initial=0.0) <= 1.01e-5, (env, x, y)
        with open(pp, "rb") as fp:
            jl, ms, pts, pos = pickle.load(fp)
        assert jl == f["joint_list"] and ms == f["main_stem_indices"] and pts == bp and pos == f["base_pos"]
        t1, t2 = pack_worlds([w]), pack_worlds([load_plant(up, pp)])
        assert np.array_equal(t1.stems, t2.stems, equal_nan=True) and np.array_equal(t1.n_stems, t2.n_stems)


def test_sampled_plant_survives_a_save_load_cycle(tmp_path):
    import random
    from plant_motion_planning_b200.model.plants import load_plant, sample_world, save_plant
    from plant_motion_planning_b200.model.tables import pack_worlds
    np.random.seed(4)
    random.seed(4)
    w = sample_world(2, 2, 1, ((0, 0.35), (-0.5, 0)), stem_base_spacing=0.05)
    save_plant(w, str(tmp_path / "p.urdf"), str(tmp_path / "p.pkl"))
    w2 = load_plant(str(tmp_path / "p.urdf"), str(tmp_path / "p.pkl"))
    assert w2.num_stems == w.num_stems and w2.joint_list == w.joint_list
    a, b = pack_worlds([w]).stems, pack_worlds([w2]).stems
    # frames, sizes, limits and rest world frames agree to the URDF's 5 printed decimals; the observer columns
    # (16..21) differ by design: a loaded plant is observed by CharacterizePlant / TwoAngleRepresentation_mod
    # (utils.py:323), a generated one by CharacterizePlant2 / _mod3 (utils.py:282)
    for geo in (slice(0, 16), slice(32, 44)):
        assert np.nanmax(np.abs(a[..., geo] - b[..., geo])) <= 2e-5
