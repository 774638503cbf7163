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
                         iterations=c.iterations, max_step=c.max_step, reg_eps=c.reg_eps, stem_margin=c.stem_margin,
                         gravity_sag=c.gravity_sag, arm_lag_gain=c.arm_lag_gain, arm_lag_substeps=c.arm_lag_substeps,
                         use_hulls=c.use_hulls,
                         mode=c.mode)
        self.num_worlds = len(self.worlds)

    def check_configs_host(self, q, detail=True):
        r = self.o.check_configs(np.asarray(q, dtype=np.float32).reshape(-1, self.dof))
        out = {"flags": r["flags"], "ee": r["ee"].astype(np.float32)}
        if detail:
            out["deflection"] = r["deflection"].astype(np.float32)
            out["theta"] = r["theta"].astype(np.float32)
        return out

    def validate_edges_host(self, q0, q1, backward=False, want_masks=False, max_steps=None, out=None,
                            first_hit_only=False):
        return self.o.validate_edges(q0, q1, backward=backward, want_masks=want_masks,
                                     max_steps=int(max_steps or self.config.max_steps))

    # ---- batched tree extension (mirrors EdgeChecker.new_tree / tree_extend / edge_points with numpy buffers) ----
    def new_tree(self, root, capacity=1 << 16):
        return OracleTree(root, capacity)

    def tree_extend(self, tree, targets, max_steps=None):
        from oracle import edgecheck_oracle as O
        targets = np.asarray(targets, dtype=np.float32)
        ms = int(max_steps or self.config.max_steps)
        idx, _ = O.nearest(self.robot, tree.t["nodes"], targets)
        q_from = tree.t["nodes"][idx].astype(np.float32)
        inert = np.isnan(targets).any(axis=1)
        first = np.zeros(len(targets), np.int32)
        n = np.ones(len(targets), np.int32)
        if (~inert).any():
            r = self.o.validate_edges(q_from[~inert], targets[~inert], backward=False, want_masks=False, max_steps=ms)
            first[~inert], n[~inert] = r["first_hit"], r["n_steps"]
        out = O.tree_append(self.robot, tree.t, idx, targets, first, n, self.config.resolution, ms)
        tree.t["nodes"] = tree.t["nodes"].astype(np.float32)
        out["new_config"] = out["new_config"].astype(np.float32)
        return _Ext(out)

    def edge_points(self, q0, q1, steps, backward=False):
        from oracle import edgecheck_oracle as O
        return np.stack([O.edge_point(self.robot, a, b, self.config.resolution, int(s), backward)
                         for a, b, s in zip(np.asarray(q0), np.asarray(q1), np.asarray(steps))]).astype(np.float32)


class _Ext:
    def __init__(self, d):
        self.new_index, self.reached, self.new_config, self.stats = d["new_index"], d["reached"], d["new_config"], d["stats"]


class OracleTree:
    def __init__(self, root, capacity):
        root = np.asarray(root, dtype=np.float32)
        self.capacity = int(capacity)
        self.t = {"nodes": root[None].copy(), "parent": np.array([-1], np.int32), "via": root[None].copy(),
                  "via_steps": np.array([0], np.int32), "capacity": self.capacity}

    def download(self):
        return {k: np.asarray(v) for k, v in self.t.items() if k != "capacity"}
