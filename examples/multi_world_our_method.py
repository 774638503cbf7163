#!/usr/bin/env python
"""scripts/multi_world_our_method.py of the reference, on the B200 edge checker (needs a GPU).

    python examples/multi_world_our_method.py [--script multi_world_our_method|multi_world_avoid_all|multi_world_ignore_all]
                                              [--deflection_limit X] [--iterations N] [--smooth]

Same seeds, plant sampler and planner control flow as the reference script (seed 1, limits ((0,0),(0.35,0.35)),
plant (1,2,1), deflection limit 2.0, one world); the robot is the vendored iiwa14 instead of the non-vendored Val.
Prints the path length, the end-effector path cost and the per-world deflection-over-limit
(Command.execute_multi_world's report, pybullet_tools/kuka_primitives.py:511-521).
"""
import argparse
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from plant_motion_planning_b200 import birrt, planner_fns as F  # noqa: E402
from plant_motion_planning_b200.env import MultiPlantWorld  # noqa: E402
from plant_motion_planning_b200.scenarios import GOAL, SCRIPTS, START  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", default="multi_world_our_method", choices=sorted(SCRIPTS))
    ap.add_argument("--deflection_limit", type=float, default=None)
    ap.add_argument("--iterations", type=int, default=2000)
    ap.add_argument("--smooth", action="store_true")
    ap.add_argument("--batched", action="store_true", help="batched RRT-Connect on device-resident trees (pmp_tree_extend)")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--plant_dir", default=None, help="our_method_multi_branch.py: directory with plant.urdf + "
                    "plant_params.pkl (e.g. the reference's environments/multi_branch/env1); deflection limit 0.30")
    ap.add_argument("--single_branch", default=None, metavar="ENV", help="our_method_single_branch.py: one of the "
                    "fixed single-stem plant sets env0..env8 (plant_motion_planning/utils.py:120-230)")
    args = ap.parse_args()
    cfg = dict(SCRIPTS[args.script])
    limit = cfg["deflection_limit"] if args.deflection_limit is None else args.deflection_limit

    np.random.seed(cfg["seed"])
    random.seed(cfg["seed"])
    if args.plant_dir or args.single_branch:
        # the two single-world scripts: one saved multi-branch plant (scripts/our_method_multi_branch.py:82-133) or a
        # fixed set of single-stem plants (scripts/our_method_single_branch.py); seeds fixed here as SURVEY 8d says
        from plant_motion_planning_b200.scenarios import single_branch_world
        if args.plant_dir:
            limit = 0.30 if args.deflection_limit is None else args.deflection_limit
            env = MultiPlantWorld([0], [0], limit, loadPath=args.plant_dir)
        else:
            env = MultiPlantWorld([0], [0], limit, worlds=[single_branch_world(args.single_branch)])
        cfg.update(mode="our_method", avoid_all=False, seed=1)
        np.random.seed(1)
        random.seed(1)
    else:
        env = MultiPlantWorld(cfg["xs"], cfg["ys"], limit, *cfg["args"], cfg["limits"], avoid_all=cfg["avoid_all"])
    if cfg["mode"] == "ignore_all":
        env.checker.set_config(mode="ignore_all")
    robot = env.sample_robot
    dist, samp, ext = F.get_distance_fn(robot), F.get_sample_fn(robot), F.get_extend_fn(robot)
    backend = birrt.GpuBackend(env.checker)
    log = birrt.PlanLog()
    t0 = time.time()
    res = None
    if args.batched:
        from plant_motion_planning_b200.batched_rrt import batched_rrt_connect
        plan = batched_rrt_connect(env.checker, START, GOAL, batch=args.batch, max_iterations=args.iterations, seed=cfg["seed"])
        if plan is None:
            print("Unable to find a plan!")
            return 1
        print(f"batched plan: {len(plan.path)} configurations through {len(plan.waypoints)} tree nodes, "
              f"{plan.iterations} iterations x {2 * args.batch} extensions, trees {plan.tree_sizes}, {time.time() - t0:.3f} s")
        path = plan.path[1:]
        if args.smooth:
            path = birrt.smooth_path_batched(env.checker, [START] + list(path), dist, rounds=10, proposals=256)[1:]
            print(f"smoothed: {len(path)} waypoints")
        ee_len, over, _ = env.path_cost(START, path)
        res = (path, ee_len, over)
        args.smooth = False
    for attempt in range(0 if args.batched else 200):                      # planner.py:36-48
        res = birrt.plan(env, START, GOAL, dist, samp, ext, backend, max_iterations=args.iterations, log=log)
        if res is not None:
            break
    if res is None:
        print("Unable to find a plan!")
        return 1
    path, ee_len, over = res
    if not args.batched:
        print(f"plan: {len(path)} configurations, {len(log.edges)} edges checked in {backend.launches} kernel launches, "
              f"{time.time() - t0:.2f} s")
    if args.smooth:
        path = birrt.smooth_path_multi_world(path, ext, dist, backend, max_iterations=20)
        ee_len, over, _ = env.path_cost(START, path)
        print(f"smoothed: {len(path)} configurations")
    print("====================================================")
    print("total path cost: ", ee_len)
    print("total Deflection over limit: ", over)
    print("alpha: ", env.config.alpha)
    print("====================================================")
    print("Total cost: ", ee_len + env.config.alpha * over)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
