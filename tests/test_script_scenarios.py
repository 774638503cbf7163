"""Decision parity on the reference scripts' default-seed environments.   -m gpu

The same BiRRT driver is run twice from the same seeds, once on flags from the CUDA path and once on flags from
the oracle (C twin, float64).  Every edge either planner proposes is logged as (kind, n, first violation); the two
logs -- hence all accept/reject decisions, the RNG consumption and the returned path -- must be identical.
"""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


class OracleBackend:
    def __init__(self, robot, worlds, **kw):
        from oracle.c_oracle import COracle
        self.o = COracle(robot, worlds, **kw)

    def flags(self, configs):
        return self.o.check_configs(np.asarray(configs, dtype=np.float32))["flags"]


def _run(robot, worlds, mode, limit, backend, seed=1, max_iterations=400):
    from plant_motion_planning_b200 import birrt, planner_fns as F
    from plant_motion_planning_b200.scenarios import GOAL, START
    np.random.seed(seed)
    random.seed(seed)
    log = birrt.PlanLog()
    res = birrt.rrt_connect_multi_world(START, GOAL, F.get_distance_fn(robot), F.get_sample_fn(robot),
                                        F.get_extend_fn(robot), backend, max_iterations=max_iterations, log=log)
    return res, log, float(np.random.rand())


def _compare(robot, worlds, mode, limit):
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.birrt import GpuBackend
    chk = EdgeChecker(robot, worlds, EdgeCheckConfig(mode=mode, deflection_limit=limit))
    g_res, g_log, g_rng = _run(robot, worlds, mode, limit, GpuBackend(chk))
    o_res, o_log, o_rng = _run(robot, worlds, mode, limit, OracleBackend(robot, worlds, mode=mode, deflection_limit=limit))
    assert g_log.edges == o_log.edges, "accept/reject decisions diverged"
    assert g_rng == o_rng, "RNG streams diverged"
    assert (g_res is None) == (o_res is None)
    if g_res is not None and g_res[0] is not None:
        assert g_res[0] == o_res[0]
    return g_res, g_log


@pytest.mark.parametrize("script", ["multi_world_our_method", "multi_world_avoid_all_seed37", "multi_world_ignore_all"])
def test_multi_world_scripts(script):
    from plant_motion_planning_b200 import load_iiwa14
    from plant_motion_planning_b200.scenarios import script_worlds
    worlds, cfg = script_worlds(script)
    res, log = _compare(load_iiwa14(), worlds, cfg["mode"], cfg["deflection_limit"])
    assert len(log.edges) >= 3
    if script != "multi_world_avoid_all_seed37":      # avoiding every plant in all 4 worlds may need more than 400 iterations
        assert res is not None and res[0] is not None, "no plan found on the default seed"


@pytest.mark.parametrize("env", ["env1", "env2", "env3", "env4"])
def test_our_method_multi_branch(env):
    from plant_motion_planning_b200 import load_iiwa14
    from plant_motion_planning_b200.model.plants import plant_from_parsed
    from plant_motion_planning_b200.model.urdf import parse_urdf
    with open(os.path.join(G, "plant_fixtures.json")) as fp:
        f = json.load(fp)[env]
    world = plant_from_parsed(parse_urdf(f["urdf"], is_text=True), f["joint_list"], f["main_stem_indices"],
                              {int(k): v for k, v in f["base_points"].items()}, f["base_pos"])
    res, log = _compare(load_iiwa14(), [world], "our_method", 0.30)
    assert len(log.edges) >= 3


@pytest.mark.parametrize("env", ["env1", "env2", "env5"])
def test_our_method_single_branch(env):
    from plant_motion_planning_b200 import load_iiwa14
    from plant_motion_planning_b200.scenarios import single_branch_world
    res, log = _compare(load_iiwa14(), [single_branch_world(env)], "our_method", 0.30)
    assert len(log.edges) >= 3


def test_batched_extension_equals_sequential():
    """extend_many (one launch for a batch of candidate extensions) gives the per-edge result of the sequential
    explicit-configuration path, including the gate's RNG use in batch order."""
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, birrt, load_iiwa14, planner_fns as F
    from plant_motion_planning_b200.scenarios import script_worlds
    worlds, cfg = script_worlds("multi_world_avoid_all_seed37")
    robot = load_iiwa14()
    chk = EdgeChecker(robot, worlds, EdgeCheckConfig(mode="our_method", deflection_limit=0.30))
    rng = np.random.RandomState(4)
    E = 64
    q0 = rng.uniform(robot.lower, robot.upper, size=(E, 7)).astype(np.float32)
    q1 = (q0 + rng.uniform(-0.6, 0.6, size=(E, 7))).clip(robot.lower, robot.upper).astype(np.float32)
    np.random.seed(3)
    first, n, _ = birrt.extend_many(chk, q0, q1)
    after_batched = np.random.rand()
    np.random.seed(3)
    ext = F.get_extend_fn(robot)
    be = birrt.GpuBackend(chk)
    seq = []
    for e in range(E):
        cfgs = list(ext(tuple(q0[e].astype(np.float64)), tuple(q1[e].astype(np.float64))))
        assert len(cfgs) == n[e]
        seq.append(birrt.first_violation(be.flags(cfgs), True))
    assert after_batched == np.random.rand()
    assert list(first) == seq


def test_smoothing_sequential_parity_and_batched_validity():
    """Shortcut smoothing after a plan: the sequential smoother makes identical decisions on GPU and oracle flags;
    the batched smoother (one launch per round) returns a shorter path that the oracle also accepts."""
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, birrt, load_iiwa14, planner_fns as F
    from plant_motion_planning_b200.scenarios import GOAL, START, script_worlds
    worlds, cfg = script_worlds("multi_world_our_method")
    robot = load_iiwa14()
    chk = EdgeChecker(robot, worlds, EdgeCheckConfig(mode=cfg["mode"], deflection_limit=cfg["deflection_limit"]))
    dist, ext, samp = F.get_distance_fn(robot), F.get_extend_fn(robot), F.get_sample_fn(robot)
    np.random.seed(1)
    random.seed(1)
    path, _ = birrt.rrt_connect_multi_world(START, GOAL, dist, samp, ext, birrt.GpuBackend(chk), max_iterations=400)
    assert path is not None
    logs = []
    outs = []
    for backend in (birrt.GpuBackend(chk), OracleBackend(robot, worlds, mode=cfg["mode"], deflection_limit=cfg["deflection_limit"])):
        np.random.seed(7)
        random.seed(7)
        log = birrt.PlanLog()
        outs.append(birrt.smooth_path_multi_world(path, ext, dist, backend, max_iterations=20, log=log))
        logs.append(log.edges)
    assert logs[0] == logs[1] and outs[0] == outs[1]
    length = lambda p: sum(dist(a, b) for a, b in zip(p[:-1], p[1:]))
    assert length(outs[0]) <= length(path) + 1e-9
    np.random.seed(3)
    way = birrt.smooth_path_batched(chk, path, dist, rounds=6, proposals=128, seed=2)
    assert way[0] == tuple(path[0]) and np.allclose(way[-1], path[-1])
    assert length(way) <= length(birrt.waypoints_from_path(path)) + 1e-9
    ob = OracleBackend(robot, worlds, mode=cfg["mode"], deflection_limit=cfg["deflection_limit"])
    for a, b in zip(way[:-1], way[1:]):
        cfgs = list(ext(a, b))
        fl = ob.flags(cfgs)
        assert not (fl & 1).any(), "batched smoother produced a segment the oracle rejects"
    # ungated variant (first-violation-only launches): every segment is free of rigid AND deflection violations
    way2 = birrt.smooth_path_batched(chk, path, dist, rounds=6, proposals=128, seed=2, gate=False)
    assert way2[0] == tuple(path[0]) and np.allclose(way2[-1], path[-1])
    assert length(way2) <= length(birrt.waypoints_from_path(path)) + 1e-9
