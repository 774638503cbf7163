#!/usr/bin/env python
"""Planner-level timing on a B200: the reference-order sequential BiRRT driver (birrt.plan, one
kernel launch per extension) against the batched device-resident planner (batched_rrt.batched_rrt_connect) on the
reference scripts' scenarios and on a crowded many-world scene, plus the nearest-node kernel alone.

    python tools/planner_bench.py > gpurun_out/planner_bench.json
"""
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, birrt, load_iiwa14, planner_fns as F  # noqa: E402
from plant_motion_planning_b200.batched_rrt import batched_rrt_connect  # noqa: E402
from plant_motion_planning_b200.env import MultiPlantWorld  # noqa: E402
from plant_motion_planning_b200.scenarios import GOAL, SCRIPTS, START, script_worlds  # noqa: E402
from plant_motion_planning_b200.workloads import synthetic_worlds  # noqa: E402


def timed(fn, repeats=3):
    best, out = 1e30, None
    for _ in range(repeats):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    robot = load_iiwa14()
    report = {"scenarios": [], "nearest": []}
    # crowded scenes: sampled worlds in which the scripts' start and goal are valid at the given deflection limit
    cand = synthetic_worlds(64)
    scenes = [(name, None, None) for name in SCRIPTS]
    for limit, n in ((0.6, 4), (1.0, 16), (1.0, 48)):
        probe = EdgeChecker(robot, cand, EdgeCheckConfig(deflection_limit=limit))
        ok = ~probe.check_configs_host(np.asarray([START, GOAL], dtype=np.float32), detail=False)["flags"].any(axis=0)
        probe.close()
        crowd = [w for w, k in zip(cand, ok) if k][:n]
        scenes.append((f"synthetic_{len(crowd)}_worlds_limit_{limit}", crowd, limit))
    if "--nearest-only" in sys.argv:
        scenes = []
    for name, worlds, limit in scenes:
        if worlds is None:
            worlds, cfg = script_worlds(name)
            limit, mode, avoid = cfg["deflection_limit"], cfg["mode"], cfg["avoid_all"]
        else:
            mode, avoid = "our_method", False
        ec = EdgeCheckConfig(deflection_limit=limit, mode=mode)
        chk = EdgeChecker(robot, worlds, ec)
        row = {"scenario": name, "worlds": len(worlds), "mode": mode}
        ends = chk.check_configs_host(np.asarray([START, GOAL], dtype=np.float32), detail=False)["flags"]
        if ends.any():
            # e.g. the seed-19 worlds of multi_world_avoid_all: a branch sits inside the Bullet margins of link 4 at
            # the zero configuration (scenarios.py), the reference would exit() there
            row["skipped"] = "start or goal is invalid in this scene"
            report["scenarios"].append(row)
            chk.close()
            continue
        for batch in (64, 256, 1024):
            t1, _ = timed(lambda: batched_rrt_connect(chk, START, GOAL, batch=batch, max_iterations=300, seed=0, sync_every=1))
            t, plan = timed(lambda: batched_rrt_connect(chk, START, GOAL, batch=batch, max_iterations=300, seed=0))
            row[f"batched_{batch}"] = {"seconds": t, "seconds_sync_every_iteration": t1, "iterations": plan.iterations if plan else None,
                                       "edges_validated": plan.edges_validated if plan else None,
                                       "path_configs": len(plan.path) if plan else None,
                                       "waypoints": len(plan.waypoints) if plan else None}
        from plant_motion_planning_b200.batched_rrt import batched_restarts
        t, ranked = timed(lambda: batched_restarts(chk, START, GOAL, restarts=8, batch=256, max_iterations=300), repeats=2)
        row["batched_restarts_8"] = {"seconds": t, "candidates": len(ranked.candidates) if ranked else 0,
                                     "valid": sum(c["valid"] for c in ranked.candidates) if ranked else 0,
                                     "ee_path_cost": ranked.ee_path_cost if ranked else None,
                                     "waypoints": len(ranked.path) if ranked else None}
        # the reference-order driver on the same checker (sequential, one launch per extension; gate replayed on the host)
        xs = list(range(len(worlds)))
        env = MultiPlantWorld(xs, [0], limit, worlds=worlds, avoid_all=avoid, checker_factory=lambda r, w, c: chk,
                              robot=robot)
        dist, samp, ext = F.get_distance_fn(robot), F.get_sample_fn(robot), F.get_extend_fn(robot)
        backend = birrt.GpuBackend(chk)

        def sequential():
            np.random.seed(1); random.seed(1)
            for _ in range(5):
                res = birrt.plan(env, START, GOAL, dist, samp, ext, backend, max_iterations=2000)
                if res is not None:
                    return res
            return None
        l0 = backend.launches
        t, res = timed(sequential, repeats=1)
        row["sequential_reference_order"] = {"seconds": t, "kernel_launches": backend.launches - l0,
                                             "path_configs": len(res[0]) if res else None}
        report["scenarios"].append(row)
        chk.close()
    # the planner's accept/reject question: full evaluation vs first-violation-only calls on random tree-like edges
    from plant_motion_planning_b200.workloads import random_edges
    report["first_hit_only"] = []
    for W in (4, 64):
        chk = EdgeChecker(robot, synthetic_worlds(W), EdgeCheckConfig())
        q0, q1 = random_edges(robot, 1 << 14, seed=5)
        a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
        row = {"worlds": W, "edges": len(q0)}
        for name, kw in (("full", {}), ("first_hit_only", {"first_hit_only": True})):
            for _ in range(3):
                out = chk.validate_edges(a, b, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                out = chk.validate_edges(a, b, **kw)
            e1.record()
            torch.cuda.synchronize()
            row[name + "_ms"] = e0.elapsed_time(e1) / 5
            row["steps"] = int(out.n_steps.sum().item())
            row["free_fraction"] = float((out.first_hit >= out.n_steps).float().mean().item())
        report["first_hit_only"].append(row)
        chk.close()
    # nearest-node kernel alone
    chk = EdgeChecker(robot, synthetic_worlds(1), EdgeCheckConfig())
    rng = np.random.RandomState(0)
    for T, B in ((1 << 10, 256), (1 << 14, 1024), (1 << 17, 1024), (1 << 20, 1024), (1 << 20, 8192)):
        nodes = torch.from_numpy(rng.uniform(robot.lower, robot.upper, size=(T, 7)).astype(np.float32)).cuda()
        targets = torch.from_numpy(rng.uniform(robot.lower, robot.upper, size=(B, 7)).astype(np.float32)).cuda()
        for _ in range(3):
            chk.nearest(nodes, targets)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            chk.nearest(nodes, targets)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        report["nearest"].append({"T": T, "B": B, "ms": ms, "pairs_per_s": T * B / (ms * 1e-3),
                                  "node_bytes_per_s": T * 28 * ((B + 7) // 8) / (ms * 1e-3)})
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
