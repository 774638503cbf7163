"""Per-call latency of the scalar adapter path (what the reference's unchanged RRT loop calls once per step)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.env import MultiPlantWorld  # noqa: E402
from plant_motion_planning_b200.workloads import synthetic_worlds  # noqa: E402

m = load_iiwa14()
for W in (1, 4, 64):
    chk = EdgeChecker(m, synthetic_worlds(W), EdgeCheckConfig())
    rng = np.random.RandomState(0)
    for B in (1, 32, 256):
        Q = rng.uniform(m.lower, m.upper, size=(B, 7)).astype(np.float32)
        for detail in (False, True):
            for _ in range(20):
                chk.check_configs_host(Q, detail=detail)
            t0 = time.perf_counter()
            n = 300
            for _ in range(n):
                chk.check_configs_host(Q, detail=detail)
            dt = (time.perf_counter() - t0) / n
            print(f"W={W:3d} B={B:4d} detail={detail}: {dt * 1e6:8.1f} us/call  {B * W / dt / 1e6:8.3f} M units/s")
env = MultiPlantWorld((0, 5), (0, 5), 0.30, 1, 2, 1, ((0.2, 0.35), (-0.5, 0.0)))
q = tuple(np.zeros(7))
t0 = time.perf_counter()
for i in range(2000):
    qq = tuple(np.array(q) + 1e-4 * i)
    env.step(qq)
    env.net_constraint_violation_checker_forward(qq)
print("facade step+check: %.1f us" % ((time.perf_counter() - t0) / 2000 * 1e6))
