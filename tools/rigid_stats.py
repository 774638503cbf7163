"""Narrow-phase counters of the hull pre-pass on a sweep slice (experiment build with -DPMP_RIGID_STATS):
    bash tools/exp_variants.sh stats "-DPMP_RIGID_STATS"; PMP_LIB=.../variants/libpmp_stats.so python tools/rigid_stats.py [log2_edges]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, _capi, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.workloads import random_edges, synthetic_edges, synthetic_worlds  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 14)
m = load_iiwa14()
chk = EdgeChecker(m, synthetic_worlds(4), EdgeCheckConfig(max_steps=32))
lib = _capi.load()
buf = (C.c_ulonglong * 128)()
for name, (q0, q1) in (("sweep", synthetic_edges(m, n)), ("ragged", random_edges(m, n, seed=3, max_len=0.6))):
    a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
    lib.pmp_debug_rigid_stats(buf)
    chk.validate_edges(a, b, first_hit_only=True)
    torch.cuda.synchronize()
    lib.pmp_debug_rigid_stats(buf)
    s = np.array(list(buf), dtype=np.int64)
    print(name, "items", s[0], "narrow steps/item %.2f" % (s[1] / max(s[0], 1)), "gjk self/step %.2f" % (s[2] / max(s[1], 1)),
          "gjk fixed/step %.2f" % (s[3] / max(s[1], 1)), "iters/gjk %.2f" % (s[4] / max(s[2] + s[3], 1)),
          "hit fraction of narrow steps %.2f" % (s[6] / max(s[1], 1)), "early-decided steps/item %.2f" % (s[5] / max(s[0], 1)))
    print("  self pairs:", {i: int(v) for i, v in enumerate(s[16:48]) if v})
    print("  obstacle x hull:", {(i // 8, i % 8): int(v) for i, v in enumerate(s[48:112]) if v})
