"""Where the wall time of a one-iteration batched plan goes on the host (cProfile) and per stage (perf_counter around
the stages with a device synchronisation after each): the script scene of multi_world_avoid_all, seed 37.
usage: python tools/planner_host_profile.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.batched_rrt import batched_rrt_connect, sample_targets  # noqa: E402
from plant_motion_planning_b200.scenarios import REFERENCE_GOAL, START, script_worlds  # noqa: E402

worlds, cfg = script_worlds("multi_world_avoid_all_seed37")
robot = load_iiwa14()
chk = EdgeChecker(robot, worlds, EdgeCheckConfig(deflection_limit=cfg["deflection_limit"], mode=cfg["mode"]))
for _ in range(5):
    batched_rrt_connect(chk, START, REFERENCE_GOAL, batch=256)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    plan = batched_rrt_connect(chk, START, REFERENCE_GOAL, batch=256)
    ts.append(time.perf_counter() - t0)
print("plan: %.3f ms median, %.3f ms min, %d iterations" % (np.median(ts) * 1e3, np.min(ts) * 1e3, plan.iterations))


def stage(name, fn, n=50):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    print("  %-34s %8.1f us" % (name, (time.perf_counter() - t0) / n * 1e6))
    return r


q2 = np.stack([np.asarray(START, np.float32), np.asarray(REFERENCE_GOAL, np.float32)])
stage("check_configs_host(2)", lambda: chk.check_configs_host(q2, detail=False))
tree = stage("new_tree(capacity 102402)", lambda: chk.new_tree(q2[0], 102402))
rng = np.random.RandomState(0)
tg = stage("sample_targets(256)", lambda: sample_targets(robot, rng, 256))
ext = stage("tree_extend(256) + sync", lambda: chk.tree_extend(chk.new_tree(q2[0], 4096), tg), n=30)
stage("stats D2H", lambda: ext.stats.cpu().numpy())
stage("tree.download()", lambda: tree.download())
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    batched_rrt_connect(chk, START, REFERENCE_GOAL, batch=256)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
