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
