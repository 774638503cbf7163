"""Two launches of the edge kernel for ncu: production (dense=0) then dense=1.  usage: python tools/profile_modes.py [log2_edges]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 15)
m = load_iiwa14()
chk = EdgeChecker(m, synthetic_worlds(64), EdgeCheckConfig(max_steps=32))
q0, q1 = synthetic_edges(m, n)
a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
for dense in (False, True):
    chk.set_config(dense=dense)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    chk.validate_edges(a, b)
    e0.record()
    r = chk.validate_edges(a, b)
    e1.record()
    torch.cuda.synchronize()
    print("dense" if dense else "prod", e0.elapsed_time(e1), "ms", n * 32 * 64 / e0.elapsed_time(e1) / 1e6, "G units/s", chk.launch_info())
