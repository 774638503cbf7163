#!/usr/bin/env python
"""Three batched tree extensions on a 2^17-node tree, 1024 targets, 8 worlds -- the launch sequence profiled for
profiles/r01_tree_kernels_ncu.csv (nearest min/exact/finish, first-violation-only edge kernel, append)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.batched_rrt import sample_targets  # noqa: E402
from plant_motion_planning_b200.workloads import synthetic_worlds  # noqa: E402

robot = load_iiwa14()
chk = EdgeChecker(robot, synthetic_worlds(8), EdgeCheckConfig(deflection_limit=2.0))
rng = np.random.RandomState(0)
T = 1 << 17
tree = chk.new_tree(np.zeros(7, np.float32), capacity=T + 8192)
tree.nodes[:T] = torch.from_numpy(sample_targets(robot, rng, T)).cuda()
tree.size.fill_(T)
tree.size_bound = T
for _ in range(3):
    ext = chk.tree_extend(tree, sample_targets(robot, rng, 1024))
torch.cuda.synchronize()
print("appended", ext.stats.cpu().numpy(), "tree size", int(tree.size.item()), chk.launch_info())
