"""plant_motion_planning_b200 -- B200-native edge validation for the UM-ARM-Lab plant planner.

Only the hot path of the reference is here: batched arm FK, arm-vs-world
collision, quasi-static plant deflection and the per-edge first-violation / deflection / cost
reduction, as hand-written sm_100a CUDA behind a C ABI (include/pmp_edgecheck.h), plus the host-side
mirror of the reference's callables (planner_fns.py, env.py).
"""
from .edgecheck import (FLAG_EXCEEDED, FLAG_RIGID, ConfigBatchResult, EdgeBatchResult, EdgeCheckConfig,
                        EdgeChecker)
from .robots import load_iiwa14

__all__ = ["EdgeChecker", "EdgeCheckConfig", "EdgeBatchResult", "ConfigBatchResult", "FLAG_RIGID",
           "FLAG_EXCEEDED", "load_iiwa14"]
__version__ = "0.1.0"
