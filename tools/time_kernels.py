"""Per-kernel device times of one validate_edges call on a sweep slice (pmp_set_kernel_timing): hull pre-pass and edge kernel.
usage: python tools/time_kernels.py [log2_edges]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
m = load_iiwa14()
chk = EdgeChecker(m, synthetic_worlds(64), EdgeCheckConfig(max_steps=32))
q0, q1 = synthetic_edges(m, n)
a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
chk.set_kernel_timing(True)
for dense in (False, True):
    chk.set_config(dense=dense)
    chk.validate_edges(a, b)
    r = [], []
    for _ in range(5):
        chk.validate_edges(a, b)
        x, y = chk.kernel_times()
        r[0].append(x); r[1].append(y)
    print("dense" if dense else "prod ", "pre-pass %.3f ms  edge kernel %.3f ms" % (np.median(r[0]), np.median(r[1])), flush=True)
