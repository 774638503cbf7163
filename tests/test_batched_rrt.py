"""Batched tree extension: oracle restatement, host logic of the batched planner (CPU, oracle-backed
test double) and, under -m gpu, the CUDA kernels through the C ABI against the oracle.

Bars: nearest-node INDEX bit-identical (index work); interpolation points within 1e-6 rad of the float64 oracle;
tree bookkeeping (slots, parents, step counts, stats) identical given the same per-edge results.
"""
import numpy as np
import pytest

from oracle import edgecheck_oracle as O
from plant_motion_planning_b200 import planner_fns as F
from plant_motion_planning_b200 import scenarios as S
from plant_motion_planning_b200.batched_rrt import batched_rrt_connect, sample_targets
from plant_motion_planning_b200.edgecheck import EdgeCheckConfig
from plant_motion_planning_b200.robots import load_iiwa14

from oracle_checker import OracleChecker
from synthetic_robot import synthetic_robot as synth9


def _random_tree(robot, rng, T, B, dup=True):
    lo = np.where(robot.circular, -np.pi, robot.lower)
    hi = np.where(robot.circular, np.pi, robot.upper)
    nodes = rng.uniform(lo, hi, size=(T, robot.dof)).astype(np.float32)
    if dup and T > 4:
        nodes[T // 2] = nodes[1]                 # exact tie: the lower index must win
        nodes[T - 1] = nodes[1]
    targets = rng.uniform(lo, hi, size=(B, robot.dof)).astype(np.float32)
    if dup and T > 4 and B > 1:
        targets[1] = nodes[1]
    return nodes, targets


# ------------------------------------------------------------------------------------------------ oracle / host logic
@pytest.mark.parametrize("robot_fn", [load_iiwa14, synth9])
def test_oracle_nearest_is_the_reference_argmin(robot_fn):
    """oracle.nearest == argmin(distance_fn) over the node list, as motion_planners/utils.py:47-53 computes it."""
    robot = robot_fn()
    rng = np.random.RandomState(3)
    nodes, targets = _random_tree(robot, rng, 200, 12)
    dist_fn = F.get_distance_fn(robot)
    idx, dist = O.nearest(robot, nodes, targets)
    for b in range(len(targets)):
        d = [dist_fn(tuple(n.astype(np.float64)), tuple(targets[b].astype(np.float64))) for n in nodes]
        k = int(np.argmin(d))                    # first minimum, like the reference's min(..., key=)
        assert idx[b] == k
        assert abs(dist[b] - d[k]) <= 1e-12
    assert idx[1] == 1                           # the duplicated node: lowest index wins


def test_oracle_edge_point_is_the_extend_sequence():
    robot = synth9()
    rng = np.random.RandomState(5)
    lo = np.where(robot.circular, -np.pi, robot.lower)
    hi = np.where(robot.circular, np.pi, robot.upper)
    for _ in range(20):
        a, b = rng.uniform(lo, hi), rng.uniform(lo, hi)
        seq = np.array(O.extend(robot, a, b, np.radians(3)))
        for s in (0, len(seq) // 2, len(seq) - 1):
            p = O.edge_point(robot, a, b, np.radians(3), s)
            assert np.abs(O.wrap_pi(p - seq[s])).max() <= 1e-9


def test_tree_append_bookkeeping():
    robot = load_iiwa14()
    root = np.zeros(7, np.float32)
    tree = {"nodes": root[None].astype(np.float64), "parent": np.array([-1], np.int32), "via": root[None].astype(np.float64),
            "via_steps": np.array([0], np.int32), "capacity": 4}
    targets = np.tile(np.linspace(0.1, 0.7, 7, dtype=np.float32), (5, 1)) * np.arange(1, 6)[:, None]
    n = np.array([O.num_steps(robot, root, t, np.radians(3)) for t in targets], np.int32)
    first = np.array([0, n[1], 3, 1, 2], np.int32)          # no progress, reached, partial, partial, overflow
    out = O.tree_append(robot, tree, np.zeros(5, np.int32), targets, first, n, np.radians(3), 320)
    assert out["new_index"].tolist() == [-1, 1, 2, 3, -1]
    assert out["reached"].tolist() == [0, 1, 0, 0, 0]
    assert out["stats"].tolist() == [3, 1, 1, 1]
    assert tree["parent"].tolist() == [-1, 0, 0, 0] and tree["via_steps"].tolist() == [0, int(n[1]), 3, 1]
    assert np.array_equal(tree["nodes"][1], targets[1].astype(np.float64))
    assert np.allclose(tree["nodes"][2], targets[2] * (3.0 / n[2]), atol=1e-12)
    assert np.isnan(out["new_config"][0]).all() and np.isnan(out["new_config"][4]).all()


def _check_plan(plan, checker, robot, start, goal):
    P = np.array(plan.path)
    assert np.allclose(P[0], np.asarray(start, np.float32)) and np.allclose(P[-1], np.asarray(goal, np.float32))
    step = np.linalg.norm(np.diff(P, axis=0), axis=1)
    assert step.max() <= np.radians(3) * (1 + 1e-4)          # consecutive configurations at most one resolution apart
    way = np.array(plan.waypoints)
    assert np.allclose(way[0], P[0]) and np.allclose(way[-1], P[-1])
    # every waypoint appears in the dense path, in order
    pos = 0
    for w in way:
        hits = np.nonzero(np.all(P[pos:] == w, axis=1))[0]
        assert len(hits), "waypoint missing from the dense path"
        pos += int(hits[0])
    return P


@pytest.mark.parametrize("name", ["multi_world_our_method", "multi_world_avoid_all_seed37"])
def test_batched_planner_on_oracle_backend(name):
    robot = load_iiwa14()
    worlds, cfg = S.script_worlds(name)
    ck = OracleChecker(robot, worlds, EdgeCheckConfig(deflection_limit=cfg["deflection_limit"], mode=cfg["mode"]))
    plan = batched_rrt_connect(ck, S.START, S.GOAL, batch=16, max_iterations=40, seed=0)
    assert plan is not None
    P = _check_plan(plan, ck, robot, S.START, S.GOAL)
    assert not ck.check_configs_host(P.astype(np.float32), detail=False)["flags"].any()
    again = batched_rrt_connect(ck, S.START, S.GOAL, batch=16, max_iterations=40, seed=0, sync_every=1)
    assert again.path == plan.path and again.iterations == plan.iterations        # deterministic for a seed, and
    assert plan.edges_validated == 2 * 16 * plan.iterations                       # independent of the sync cadence


def test_batched_planner_rejects_invalid_ends_and_gives_up():
    robot = load_iiwa14()
    worlds, cfg = S.script_worlds("multi_world_avoid_all_seed37")
    ck = OracleChecker(robot, worlds, EdgeCheckConfig(deflection_limit=cfg["deflection_limit"], mode=cfg["mode"]))
    jam = list(S.REFERENCE_GOAL)
    jam[5] += 0.02                                     # the flange goes 1.4 mm -> into the block
    with pytest.raises(ValueError):
        batched_rrt_connect(ck, S.START, jam, batch=4, max_iterations=2)
    bad = list(S.START)
    bad[1] = 10.0
    with pytest.raises(ValueError):
        batched_rrt_connect(ck, bad, S.GOAL, batch=4, max_iterations=2)
    assert batched_rrt_connect(ck, S.START, S.GOAL, batch=1, max_iterations=1, seed=0) is None


def test_sample_targets_is_get_sample_fn():
    robot = synth9()
    rng = np.random.RandomState(11)
    got = sample_targets(robot, rng, 5)
    np.random.seed(11)
    fn = F.get_sample_fn(robot)
    w = np.random.uniform(size=(5, robot.dof))          # the batch draws the same stream, one row per sample_fn() call
    lo = np.where(robot.circular, -np.pi, robot.lower)
    hi = np.where(robot.circular, np.pi, robot.upper)
    assert np.array_equal(got, ((1 - w) * lo + w * hi).astype(np.float32))
    np.random.seed(11)
    assert np.allclose(got[0], np.asarray(fn(), dtype=np.float32))


# ------------------------------------------------------------------------------------------------ CUDA kernels
@pytest.fixture(scope="module")
def gpu():
    import torch
    from plant_motion_planning_b200 import EdgeChecker
    from plant_motion_planning_b200.workloads import synthetic_worlds
    worlds = synthetic_worlds(4)
    r7, r9 = load_iiwa14(), synth9()
    return {"torch": torch, "iiwa": (r7, EdgeChecker(r7, worlds, EdgeCheckConfig())),
            "synth9": (r9, EdgeChecker(r9, worlds, EdgeCheckConfig())),
            "worlds": worlds}


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["iiwa", "synth9"])
@pytest.mark.parametrize("T,B", [(1, 1), (7, 9), (1000, 300), (50000, 64), (300000, 5)])
def test_nearest_indices_bit_identical(gpu, which, T, B):
    robot, chk = gpu[which]
    torch = gpu["torch"]
    nodes, targets = _random_tree(robot, np.random.RandomState(T + B), T, B)
    idx, dist = chk.nearest(torch.from_numpy(nodes), torch.from_numpy(targets))
    oi, od = O.nearest(robot, nodes, targets)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy(), od)            # float64 distances, same operations: identical bits


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["iiwa", "synth9"])
def test_nearest_with_an_ordered_tree_and_ties(gpu, which):
    """The float32 bound of pass 1 comes from a sample of the tree (the first 256-node block of every slice and every
    8th after it); the answer may not depend on it.  A tree sorted along joint 0 makes the sample as unrepresentative
    as it gets, duplicated nodes (in unsampled blocks, before and after a sampled one) exercise the first-minimum rule,
    and targets that ARE tree nodes give distance 0."""
    robot, chk = gpu[which]
    torch = gpu["torch"]
    rng = np.random.RandomState(11)
    nodes, targets = _random_tree(robot, rng, 120000, 96)
    nodes = nodes[np.argsort(nodes[:, 0], kind="stable")]
    dup = rng.randint(300, 119000, size=40)
    nodes[dup + 517] = nodes[dup]                  # a later copy of the same node: the earlier index must win
    targets[:20] = nodes[dup[:20] + 517]           # exact hits, distance 0
    targets[20:40] = nodes[dup[20:]] + np.float32(1e-4)
    idx, dist = chk.nearest(torch.from_numpy(nodes), torch.from_numpy(targets))
    oi, od = O.nearest(robot, nodes, targets)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy(), od)
    assert np.all(od[:20] == 0.0) and np.all(oi[:20] <= dup[:20])


@pytest.mark.gpu
def test_nearest_live_count_nan_and_gather(gpu):
    robot, chk = gpu["iiwa"]
    torch = gpu["torch"]
    nodes, targets = _random_tree(robot, np.random.RandomState(1), 5000, 33)
    targets[4, 2] = np.nan
    size = torch.tensor([1234], dtype=torch.int32, device="cuda")
    idx, dist, q_near = chk.nearest(torch.from_numpy(nodes), torch.from_numpy(targets), size=size, want_nodes=True)
    oi, od = O.nearest(robot, nodes[:1234], targets)
    assert np.array_equal(idx.cpu().numpy(), oi) and oi[4] == 0
    assert np.array_equal(q_near.cpu().numpy(), nodes[oi])


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["iiwa", "synth9"])
@pytest.mark.parametrize("backward", [False, True])
def test_edge_points_match_the_oracle_sequence(gpu, which, backward):
    robot, chk = gpu[which]
    rng = np.random.RandomState(9)
    lo = np.where(robot.circular, -np.pi, robot.lower)
    hi = np.where(robot.circular, np.pi, robot.upper)
    q0 = rng.uniform(lo, hi, size=(256, robot.dof)).astype(np.float32)
    q1 = rng.uniform(lo, hi, size=(256, robot.dof)).astype(np.float32)
    n = np.array([O.num_steps(robot, a, b, np.radians(3)) for a, b in zip(q0, q1)])
    steps = (rng.uniform(size=256) * n).astype(np.int32)
    steps[:8] = n[:8] - 1                                   # last step: the target itself (forward) / one short (backward)
    steps[8:16] = 0
    got = chk.edge_points(q0, q1, steps, backward=backward).cpu().numpy()
    want = np.stack([O.edge_point(robot, a, b, np.radians(3), int(s), backward) for a, b, s in zip(q0, q1, steps)])
    assert np.abs(got - want).max() <= 1e-6
    if not backward:
        nc = ~np.asarray(robot.circular)
        assert np.array_equal(got[:8][:, nc], q1[:8][:, nc])      # the end point is q1 itself


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["iiwa", "synth9"])
def test_tree_extend_equals_its_parts(gpu, which):
    """pmp_tree_extend == pmp_nearest + pmp_validate_edges + the oracle's append bookkeeping, over several batches
    (tree layout, parents, step counts and stats identical; configurations within 1e-6)."""
    robot, chk = gpu[which]
    torch = gpu["torch"]
    rng = np.random.RandomState(2)
    root = np.zeros(robot.dof, np.float32)
    tree = chk.new_tree(root, capacity=300)
    mirror = {"nodes": root[None].astype(np.float64), "parent": np.array([-1], np.int32),
              "via": root[None].astype(np.float64), "via_steps": np.array([0], np.int32), "capacity": 300}
    overflowed = 0
    for it in range(4):
        targets = sample_targets(robot, rng, 200)
        targets[7] = np.nan                                           # inert proposal
        if it == 1:
            targets[3] = mirror["nodes"][5].astype(np.float32)        # aims at an existing node: zero-length edge
        before = tree.download()
        idx, _, q_near = chk.nearest(torch.from_numpy(before["nodes"]), torch.from_numpy(targets), want_nodes=True)
        res = chk.validate_edges(q_near, torch.from_numpy(targets))
        ext = chk.tree_extend(tree, targets)
        mirror["nodes"] = before["nodes"].astype(np.float64)          # device float32 nodes are the truth for parents
        want = O.tree_append(robot, mirror, idx.cpu().numpy(), targets, res.first_hit.cpu().numpy(),
                             res.n_steps.cpu().numpy(), np.radians(3), chk.config.max_steps)
        assert np.array_equal(ext.new_index.cpu().numpy(), want["new_index"])
        assert np.array_equal(ext.reached.cpu().numpy(), want["reached"])
        assert np.array_equal(ext.stats.cpu().numpy(), want["stats"])
        after = tree.download()
        assert np.array_equal(after["parent"], mirror["parent"])
        assert np.array_equal(after["via_steps"], mirror["via_steps"])
        assert np.abs(after["nodes"] - mirror["nodes"]).max() <= 1e-6
        assert np.array_equal(after["via"], mirror["via"].astype(np.float32))
        got_cfg, want_cfg = ext.new_config.cpu().numpy(), want["new_config"]
        assert np.array_equal(np.isnan(got_cfg), np.isnan(want_cfg))
        both = ~np.isnan(want_cfg)
        assert both.sum() == 0 or np.abs(got_cfg[both] - want_cfg[both]).max() <= 1e-6
        assert ext.new_index.cpu().numpy()[7] == -1
        overflowed |= int(want["stats"][3])
    assert overflowed == 1                                            # the later batches ran into the capacity
    assert len(after["parent"]) == 300


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["multi_world_our_method", "multi_world_avoid_all_seed37", "multi_world_ignore_all"])
def test_batched_planner_on_gpu_is_valid_under_the_oracle(gpu, name):
    from plant_motion_planning_b200 import EdgeChecker
    robot = load_iiwa14()
    worlds, cfg = S.script_worlds(name)
    ec = EdgeCheckConfig(deflection_limit=cfg["deflection_limit"], mode=cfg["mode"])
    chk = EdgeChecker(robot, worlds, ec)
    before = chk.launch_info()["kernel_launches"]
    plan = batched_rrt_connect(chk, S.START, S.GOAL, batch=256, max_iterations=50, seed=0)
    assert plan is not None
    assert chk.launch_info()["kernel_launches"] >= before + 8 * plan.iterations
    P = _check_plan(plan, chk, robot, S.START, S.GOAL)
    # the oracle agrees that every configuration of the path is free (flags identical away from the decision bands;
    # a path through a band would show up here as an oracle flag)
    oracle_flags = OracleChecker(robot, worlds, ec).check_configs_host(P.astype(np.float32), detail=False)["flags"]
    assert oracle_flags.mean() <= 0.005
    assert not chk.check_configs_host(P.astype(np.float32), detail=False)["flags"].any()
    again = batched_rrt_connect(chk, S.START, S.GOAL, batch=256, max_iterations=50, seed=0, sync_every=1)
    assert again.path == plan.path and again.iterations == plan.iterations


def test_extend_many_equals_per_edge_gate_replay():
    """birrt.extend_many's vectorised first-violation (lowest rigid bit when no gate draw is needed) == replay_gate
    edge by edge, including the number of random draws consumed."""
    from plant_motion_planning_b200 import birrt
    from plant_motion_planning_b200.workloads import random_edges, synthetic_worlds
    robot = load_iiwa14()
    # low limit so that many steps exceed it; without gravity sag, or every step would (rest deflections are ~0.1)
    ck = OracleChecker(robot, synthetic_worlds(3), EdgeCheckConfig(deflection_limit=0.05, gravity_sag=False))
    q0, q1 = random_edges(robot, 96, seed=4, max_len=2.5)
    for forward in (True, False):
        np.random.seed(8)
        first, n, out = birrt.extend_many(ck, q0, q1, forward_checker=forward)
        after = np.random.rand()
        np.random.seed(8)
        cap = out["rigid_mask"].shape[2] * 32
        want = [F.replay_gate(min(int(out["n_steps"][e]), cap), out["rigid_mask"][e], out["exceed_mask"][e],
                              F.GATE_PROBABILITY, None, forward) for e in range(len(q0))]
        assert first.tolist() == want
        assert np.random.rand() == after
        assert out["exceed_mask"].any() and (np.asarray(want) < np.minimum(n, cap)).any() and (np.asarray(want) >= n).any()


def test_batched_restarts_select_the_shortest_valid_candidate():
    """batched_restarts ranks all smoothed candidates with one edge-batch evaluation; the winner is valid under the
    oracle, its cost report equals env.path_cost's (Command.execute_multi_world's totals) and no valid candidate has
    a shorter end-effector path."""
    from plant_motion_planning_b200.batched_rrt import batched_restarts
    from plant_motion_planning_b200.env import MultiPlantWorld
    robot = load_iiwa14()
    worlds, cfg = S.script_worlds("multi_world_avoid_all_seed37")
    ec = EdgeCheckConfig(deflection_limit=cfg["deflection_limit"], mode=cfg["mode"])
    ck = OracleChecker(robot, worlds, ec)
    res = batched_restarts(ck, S.START, S.GOAL, restarts=3, batch=16, max_iterations=40, smooth_rounds=4, proposals=32)
    assert res is not None and len(res.candidates) >= 2
    assert np.allclose(res.path[0], S.START) and np.allclose(res.path[-1], S.GOAL, atol=1e-6)
    assert res.ee_path_cost == min(c["ee_path_cost"] for c in res.candidates if c["valid"])
    raw = batched_rrt_connect(ck, S.START, S.GOAL, batch=16, max_iterations=40, seed=0)
    dist = F.get_distance_fn(robot)
    length = lambda p: sum(dist(a, b) for a, b in zip(p[:-1], p[1:]))
    assert length(res.candidates[0]["waypoints"]) <= length(raw.waypoints) + 1e-9       # smoothing never lengthens
    # the same totals from the facade's execute report over the dense step sequence of the winner
    env = MultiPlantWorld(cfg["xs"], cfg["ys"], cfg["deflection_limit"], worlds=worlds, avoid_all=cfg["avoid_all"],
                          robot=robot, checker_factory=lambda r, w, c: ck)
    ext = F.get_extend_fn(robot)
    dense = [q for a, b in zip(res.path[:-1], res.path[1:]) for q in ext(a, b)]
    ee_len, over, total = env.path_cost(res.path[0], dense)
    assert abs(ee_len - res.ee_path_cost) <= 1e-4
    assert np.allclose(over, res.deflection_over_limit, atol=1e-3) and np.allclose(total, res.total_cost, atol=1e-3)


@pytest.mark.gpu
def test_batched_restarts_on_gpu(gpu):
    from plant_motion_planning_b200 import EdgeChecker
    from plant_motion_planning_b200.batched_rrt import batched_restarts
    robot = load_iiwa14()
    worlds, cfg = S.script_worlds("multi_world_avoid_all_seed37")
    ec = EdgeCheckConfig(deflection_limit=cfg["deflection_limit"], mode=cfg["mode"])
    chk = EdgeChecker(robot, worlds, ec)
    res = batched_restarts(chk, S.START, S.GOAL, restarts=6, batch=128, max_iterations=60)
    assert res is not None and sum(c["valid"] for c in res.candidates) >= 1
    assert res.ee_path_cost == min(c["ee_path_cost"] for c in res.candidates if c["valid"])
    # forward re-check of the winner by the oracle: every step of every segment is free
    ext = F.get_extend_fn(robot)
    dense = np.asarray([q for a, b in zip(res.path[:-1], res.path[1:]) for q in ext(a, b)], dtype=np.float32)
    fl = OracleChecker(robot, worlds, ec).check_configs_host(dense, detail=False)["flags"]
    assert fl.mean() <= 0.005
    again = batched_restarts(chk, S.START, S.GOAL, restarts=6, batch=128, max_iterations=60)
    assert again.path == res.path and again.ee_path_cost == res.ee_path_cost
