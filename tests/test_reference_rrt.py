"""The reference's own RRT loop vs the restated driver.

tests/golden/reference_rrt.json holds the paths returned by the reference's UNCHANGED rrt_connect_multi_world
(motion_planners/.../rrt_connect_with_angle_constraints_v2.py:202-280) when it is given this repository's
MultiPlantWorld facade and the reference's own sample/distance/extend callables, from the scripts' seeds.
birrt.rrt_connect_multi_world must return the same paths and leave np.random in the same state:
  * on oracle flags (CPU, runs everywhere)   -> pins the restated control flow against the reference's code
  * on GPU flags (-m gpu)                    -> identical accept/reject decisions of the CUDA path along whole plans
"""
import json
import os
import random

import numpy as np
import pytest

from oracle_checker import OracleChecker
from plant_motion_planning_b200 import birrt
from plant_motion_planning_b200 import planner_fns as F
from plant_motion_planning_b200.robots import load_iiwa14
from plant_motion_planning_b200.scenarios import GOAL, SCRIPTS, START, script_worlds

G = os.path.join(os.path.dirname(__file__), "golden")


class _Backend(birrt.FlagBackend):
    def __init__(self, checker, plant_aware):
        self.checker, self.plant_aware = checker, plant_aware

    def flags(self, configs):
        fl = self.checker.check_configs_host(np.asarray(configs, np.float32), detail=False)["flags"]
        return fl if self.plant_aware else (fl & 1)     # *_benchmark checkers: rigid obstacles only (env.py:249-255)


def _plan(script, checker_cls):
    from plant_motion_planning_b200.edgecheck import EdgeCheckConfig
    with open(os.path.join(G, "reference_rrt.json")) as fp:
        gold = json.load(fp)[script]
    worlds, cfg = script_worlds(script)
    robot = load_iiwa14()
    mode = "avoid_all" if cfg["avoid_all"] else "our_method"
    chk = checker_cls(robot, worlds, EdgeCheckConfig(mode=mode, deflection_limit=cfg["deflection_limit"]))
    np.random.seed(gold["planner_seed"])
    random.seed(gold["planner_seed"])
    backend = _Backend(chk, gold["plant_aware"])
    res = birrt.rrt_connect_multi_world(START, GOAL, F.get_distance_fn(robot), F.get_sample_fn(robot),
                                        F.get_extend_fn(robot), backend)
    if gold["path"] is None:            # the script's own seed: the start touches a plant (scenarios.py)
        assert res is None or res[0] is None
        return None, None, gold
    path, _ = res
    nxt = float(np.random.rand())
    np.random.seed(gold["smoothing_seed"])
    random.seed(gold["smoothing_seed"])
    sm = birrt.smooth_path_multi_world(path, F.get_extend_fn(robot), F.get_distance_fn(robot), backend, max_iterations=20)
    gold["_smoothed"] = ([list(q) for q in sm], float(np.random.rand()), float(random.random()))
    return path, nxt, gold


@pytest.mark.parametrize("script", sorted(SCRIPTS))
def test_restated_driver_equals_reference_rrt_on_oracle_flags(script):
    path, nxt, gold = _plan(script, OracleChecker)
    if gold["path"] is None:
        return
    assert path is not None and len(path) == gold["path_len"]
    assert [list(q) for q in path] == gold["path"]
    assert nxt == gold["next_np_rand"]
    sm, np_next, py_next = gold["_smoothed"]
    assert sm == gold["smoothed"] and np_next == gold["smoothed_next_np_rand"] and py_next == gold["smoothed_next_py_random"]


@pytest.mark.gpu
@pytest.mark.parametrize("script", sorted(SCRIPTS))
def test_gpu_plans_equal_reference_rrt(script):
    from plant_motion_planning_b200 import EdgeChecker
    path, nxt, gold = _plan(script, EdgeChecker)
    if gold["path"] is None:
        return
    assert path is not None and [list(q) for q in path] == gold["path"]
    assert nxt == gold["next_np_rand"]
    sm, np_next, py_next = gold["_smoothed"]
    assert sm == gold["smoothed"] and np_next == gold["smoothed_next_np_rand"] and py_next == gold["smoothed_next_py_random"]


def _cost(script, factory):
    from plant_motion_planning_b200.env import MultiPlantWorld
    with open(os.path.join(G, "reference_rrt.json")) as fp:
        gold = json.load(fp)[script]
    cfg = SCRIPTS[script]
    np.random.seed(cfg["seed"])
    random.seed(cfg["seed"])
    env = MultiPlantWorld(cfg["xs"], cfg["ys"], cfg["deflection_limit"], *cfg["args"], cfg["limits"],
                          avoid_all=cfg["avoid_all"], checker_factory=factory)
    ee_len, over, total = env.path_cost(START, [tuple(q) for q in gold["smoothed"]])
    # the batched execute report carries the same totals plus the per-world deflection traces
    import io
    buf = io.StringIO()
    rep = env.execute_multi_world(START, [tuple(q) for q in gold["smoothed"]], out=buf)
    assert abs(rep["ee_path_cost"] - ee_len) < 1e-9 and np.allclose(rep["deflection_over_limit"], over, atol=1e-9)
    assert rep["deflection"].shape[:2] == (len(gold["smoothed"]), len(env.envs))
    text = buf.getvalue()
    assert text.startswith("Executing path...") and "total Deflection over limit: " in text
    assert text.count("Environment: ") == len(gold["smoothed"]) * len(env.envs)
    assert (text.count("Error! Deflection limit exceeded!") > 0) == bool(np.any(over > 0))
    return ee_len, over, total, gold["execute"]


PLANNED = sorted(s for s in SCRIPTS if s != "multi_world_avoid_all")      # seed 19: the start touches a plant


@pytest.mark.parametrize("script", PLANNED)
def test_path_cost_equals_reference_execute_report(script):
    """env.path_cost vs the totals printed by the reference's unchanged Command.execute_multi_world
    (pybullet_tools/kuka_primitives.py:467-521) on the same smoothed path (oracle-backed facade)."""
    ee_len, over, total, ex = _cost(script, OracleChecker)
    assert abs(ee_len - ex["ee_len"]) < 1e-5
    assert np.allclose(over, ex["over"], atol=1e-4) and np.allclose(total, ex["total"], atol=1e-4)
    if script == "multi_world_ignore_all":
        assert max(ex["over"]) > 1.0          # the ignore-all plan really does bend plants past the limit


@pytest.mark.gpu
@pytest.mark.parametrize("script", PLANNED)
def test_gpu_path_cost_equals_reference_execute_report(script):
    ee_len, over, total, ex = _cost(script, None)
    assert abs(ee_len - ex["ee_len"]) < 1e-4
    assert np.allclose(over, ex["over"], atol=5e-3, rtol=1e-3) and np.allclose(total, ex["total"], atol=1e-3, rtol=1e-3)


def _planner_case(factory):
    from plant_motion_planning_b200.env import MultiPlantWorld
    from synthetic_robot import val_like_robot
    with open(os.path.join(G, "reference_planner.json")) as fp:
        gold = json.load(fp)
    model, _ = val_like_robot()
    np.random.seed(gold["world_seed"])
    random.seed(gold["world_seed"])
    env = MultiPlantWorld((0, 5), (0, 5), 0.30, 1, 2, 1, ((0.2, 0.35), (-0.5, 0.0)), robot=model, checker_factory=factory)
    backend = _Backend(env.checker, True)
    np.random.seed(gold["planner_seed"])
    random.seed(gold["planner_seed"])
    path = birrt.move_arm_conf2conf_multi_world(tuple(gold["start"]), tuple(gold["goal"]), F.get_distance_fn(model),
                                                F.get_sample_fn(model), F.get_extend_fn(model), backend)
    nxt = (float(np.random.rand()), float(random.random()))
    ee_len, over, _ = env.path_cost(tuple(gold["start"]), path)
    return path, nxt, ee_len, over, gold


def test_restated_planner_equals_reference_planner_on_oracle_flags():
    """The reference's unchanged Planner.move_arm_conf2conf_multi_world (whole chain: free-motion generator, start/goal
    checks, random restarts, RRT-connect, 10 x 20 smoothing, forward re-checks; 32 018 checker calls) on a robot with
    Val's joint layout, vs birrt.move_arm_conf2conf_multi_world: same path, same RNG state, same cost report."""
    path, nxt, ee_len, over, gold = _planner_case(OracleChecker)
    assert [list(q) for q in path] == gold["path"]
    assert nxt == (gold["next_np_rand"], gold["next_py_random"])
    assert abs(ee_len - gold["execute"]["ee_len"]) < 1e-5 and np.allclose(over, gold["execute"]["over"], atol=1e-4)


@pytest.mark.gpu
def test_gpu_planner_equals_reference_planner():
    path, nxt, ee_len, over, gold = _planner_case(None)
    assert [list(q) for q in path] == gold["path"]
    assert nxt == (gold["next_np_rand"], gold["next_py_random"])
    assert abs(ee_len - gold["execute"]["ee_len"]) < 1e-4 and np.allclose(over, gold["execute"]["over"], atol=5e-3)


def _benchmark_case(factory):
    from plant_motion_planning_b200.env import MultiPlantWorld
    from synthetic_robot import val_like_robot
    with open(os.path.join(G, "reference_planner.json")) as fp:
        gold = json.load(fp)
    model, _ = val_like_robot()
    np.random.seed(gold["world_seed"])
    random.seed(gold["world_seed"])
    env = MultiPlantWorld((0, 5), (0, 5), 0.30, 1, 2, 1, ((0.2, 0.35), (-0.5, 0.0)), robot=model, avoid_all=True,
                          checker_factory=factory)
    np.random.seed(gold["planner_seed"])
    random.seed(gold["planner_seed"])
    b = gold["benchmark"]
    path = birrt.move_arm_conf2conf_multi_world_benchmark(tuple(gold["start"]), tuple(b["goal"]), F.get_distance_fn(model),
                                                          F.get_sample_fn(model), F.get_extend_fn(model),
                                                          _Backend(env.checker, True))
    return path, (float(np.random.rand()), float(random.random())), b


def test_restated_benchmark_planner_equals_reference_on_oracle_flags():
    """Planner.move_arm_conf2conf_multi_world_benchmark (planner.py:53-78; what multi_world_avoid_all / ignore_all run),
    unchanged reference code vs birrt.move_arm_conf2conf_multi_world_benchmark, avoid_all worlds."""
    path, nxt, b = _benchmark_case(OracleChecker)
    assert [list(q) for q in path] == b["path"]
    assert nxt == (b["next_np_rand"], b["next_py_random"])


@pytest.mark.gpu
def test_gpu_benchmark_planner_equals_reference():
    path, nxt, b = _benchmark_case(None)
    assert [list(q) for q in path] == b["path"]
    assert nxt == (b["next_np_rand"], b["next_py_random"])
