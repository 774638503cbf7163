"""Test double with EdgeChecker's host API, backed by the C twin of the oracle.   *** TEST INFRASTRUCTURE ***
Lets the host-side logic (facade, BiRRT driver, the reference's own RRT loop) run in the CPU-only build container."""
import numpy as np

from oracle.c_oracle import COracle
from plant_motion_planning_b200.edgecheck import EdgeCheckConfig


class OracleChecker:
    def __init__(self, robot, worlds, config=None):
        self.config = config or EdgeCheckConfig()
        self.robot = robot
        self.dof = robot.dof
        self.worlds = list(worlds)
        c = self.config
        self.o = COracle(robot, self.worlds, resolution=c.resolution, deflection_limit=c.deflection_limit, alpha=c.alpha,
                         iterations=c.iterations, max_step=c.max_step, reg_eps=c.reg_eps, stem_radius=c.stem_radius,
                         mode=c.mode)
        self.num_worlds = len(self.worlds)

    def check_configs_host(self, q, detail=True):
        r = self.o.check_configs(np.asarray(q, dtype=np.float32).reshape(-1, self.dof))
        out = {"flags": r["flags"], "ee": r["ee"].astype(np.float32)}
        if detail:
            out["deflection"] = r["deflection"].astype(np.float32)
            out["theta"] = r["theta"].astype(np.float32)
        return out

    def validate_edges_host(self, q0, q1, backward=False, want_masks=False, max_steps=None, out=None):
        return self.o.validate_edges(q0, q1, backward=backward, want_masks=want_masks,
                                     max_steps=int(max_steps or self.config.max_steps))
