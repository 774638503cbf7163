"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck).  usage: compute-sanitizer python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.model.plants import SINGLE_STEM_ENVS, single_stem_world  # noqa: E402
from plant_motion_planning_b200.workloads import random_edges, synthetic_worlds  # noqa: E402

m = load_iiwa14()
for worlds in (synthetic_worlds(5), [single_stem_world(SINGLE_STEM_ENVS["env3"])], synthetic_worlds(70)):
    for mode in ("our_method", "avoid_all", "ignore_all"):
        chk = EdgeChecker(m, worlds, EdgeCheckConfig(mode=mode))
        q0, q1 = random_edges(m, 97, seed=1)
        for bw in (False, True):
            r = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), backward=bw, want_masks=True, max_steps=70)
        chk.validate_edges(torch.from_numpy(q0[:3]), torch.from_numpy(q1[:3]))
        h = chk.validate_edges_host(q0, q1, want_masks=True)
        c = chk.check_configs(torch.from_numpy(q0[:33]))
        chk.check_configs_host(q0[:5])
        chk.fk(torch.from_numpy(q0[:50]))
        torch.cuda.synchronize()
        chk.close()
print("sanitize_small ok")
