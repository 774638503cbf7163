"""Batched bidirectional RRT-Connect on device-resident trees.

The reference grows its two trees one edge at a time: `argmin(distance_fn)` over the tree in a Python loop, then one
`env.step` + `collision_fn` per interpolation step (motion_planners/motion_planners/primitives.py:196-255;
rrt_connect_with_angle_constraints_v2.py:240-290).  Here every iteration proposes ``batch`` targets at once and the
whole extension -- nearest node, multi-world edge validation, append of the last safe configuration -- is one
`pmp_tree_extend` call (five kernels, trees never leave HBM).  The connect step aims the other tree at the nodes
just added.  The host reads back 16 bytes per iteration (the `stats` vector) to learn whether the trees met.

Semantics kept from the reference: extension stops at the first violating step and keeps the safe prefix
(`safe = steps[:first_hit]`), targets are sampled with get_sample_fn's formula from one RandomState, start and goal
must be valid.  Differences (this is the additive batched surface, not the bit-for-bit replay -- that is birrt.py):
one node per extension instead of one per step (the dense step sequence is recovered exactly by `pmp_edge_points`),
both trees use the full plant-aware check, the 5 % gate is off unless the checker's config enables the on-device gate.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

Conf = Tuple[float, ...]


def _np(x) -> np.ndarray:
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


@dataclass
class BatchedPlan:
    path: List[Conf]                   # dense sequence start .. goal: exactly the configurations that were validated
    waypoints: List[Conf]              # the tree nodes along it
    iterations: int
    tree_sizes: Tuple[int, int]
    edges_validated: int
    log: List[dict] = field(default_factory=list)


def sample_targets(robot, rng: np.random.RandomState, batch: int) -> np.ndarray:
    """``batch`` draws of get_sample_fn (pybullet_tools/utils.py:3593-3599): (1 - w) * lower + w * upper with
    continuous joints sampled in [-pi, pi]."""
    lo = np.where(robot.circular, -np.pi, robot.lower).astype(np.float64)
    hi = np.where(robot.circular, np.pi, robot.upper).astype(np.float64)
    w = rng.uniform(size=(batch, robot.dof))
    return ((1.0 - w) * lo + w * hi).astype(np.float32)


def _drop_repeats(Q: np.ndarray) -> np.ndarray:
    keep = np.ones(len(Q), dtype=bool)
    keep[1:] = np.any(Q[1:] != Q[:-1], axis=1)
    return Q[keep]


def _chain(parent: np.ndarray, idx: int) -> List[int]:
    seq = []
    while idx >= 0:
        seq.append(int(idx))
        idx = int(parent[idx])
    return seq[::-1]


def dense_chain(checker, tree: dict, chain: Sequence[int]) -> Tuple[np.ndarray, List[int]]:
    """All validated configurations along root -> chain[-1]: for every node its edge's steps 0 .. via_steps-1,
    recomputed on the device with the edge kernel's arithmetic.  Returns (configs [n, D] incl. the root, index of
    each chain node in it)."""
    q0, q1, st, marks = [], [], [], [0]
    for node in chain[1:]:
        p, k = int(tree["parent"][node]), int(tree["via_steps"][node])
        for s in range(k):
            q0.append(tree["nodes"][p]); q1.append(tree["via"][node]); st.append(s)
        marks.append(len(st))
    root = np.asarray(tree["nodes"][chain[0]], dtype=np.float32)[None]
    if not st:
        return root, marks
    pts = _np(checker.edge_points(np.asarray(q0, dtype=np.float32), np.asarray(q1, dtype=np.float32),
                                  np.asarray(st, dtype=np.int32)))
    return np.concatenate([root, pts.astype(np.float32)], axis=0), marks


def batched_rrt_connect(checker, start: Sequence[float], goal: Sequence[float], *, batch: int = 256,
                        max_iterations: int = 200, seed: int = 0, capacity: Optional[int] = None,
                        max_steps: Optional[int] = None, sync_every: int = 4) -> Optional[BatchedPlan]:
    """Plan start -> goal; returns None if ``max_iterations`` batches did not connect the trees.
    Raises ValueError if start or goal is invalid (the reference exits there, pybullet_tools/utils.py:4364-4380).
    ``sync_every``: the host reads the 16-byte termination records of up to this many iterations at once (1, 1, 2, 4,
    ... iterations per read: easy scenes connect in the first batch and must not pay for queued ones), so the launches
    of the next iterations are queued while the current ones run.  The trees meet at the first connecting iteration
    either way; at most ``sync_every - 1`` further batches are appended to the trees before the host notices, and
    the returned path does not depend on it."""
    robot = checker.robot
    start32 = np.asarray(start, dtype=np.float32)
    goal32 = np.asarray(goal, dtype=np.float32)
    ends = checker.check_configs_host(np.stack([start32, goal32]), detail=False)["flags"]
    lo = np.where(robot.circular, -np.inf, robot.lower)
    hi = np.where(robot.circular, np.inf, robot.upper)
    for name, q, fl in (("start", start32, ends[0]), ("goal", goal32, ends[1])):
        if fl.any() or (q < lo).any() or (q > hi).any():
            raise ValueError(f"{name} configuration is invalid (collision, deflection limit or joint limits)")
    capacity = int(capacity or (2 * batch * max_iterations + 2))
    rng = np.random.RandomState(seed)
    t_start = checker.new_tree(start32, capacity)
    t_goal = checker.new_tree(goal32, capacity)
    roots = {id(t_start): start32, id(t_goal): goal32}
    ta, tb = t_start, t_goal
    log: List[dict] = []
    hit = None
    it = 0
    sync_every = max(1, int(sync_every))
    chunk, reads = 1, 0                                 # iterations per host read: 1, 1, 2, 4, ... <= sync_every
    pending = []                                        # (iteration, ext_a, ext_b, tree a, tree b) not yet looked at
    for nxt in range(1, max_iterations + 1):
        targets = sample_targets(robot, rng, batch)
        targets[0] = roots[id(tb)]                      # one proposal per batch aims at the other root
        ext_a = checker.tree_extend(ta, targets, max_steps)
        ext_b = checker.tree_extend(tb, ext_a.new_config, max_steps)     # rows without a new node are NaN: inert
        pending.append((nxt, ext_a, ext_b, ta, tb))
        ta, tb = tb, ta
        if len(pending) < chunk and nxt < max_iterations:
            continue
        reads += 1
        if reads >= 2:
            chunk = min(2 * chunk, sync_every)
        # the only synchronisation: 2 x 16 bytes per pending iteration in one copy
        if hasattr(pending[0][1].stats, "detach"):
            import torch
            stats = _np(torch.stack([torch.stack([p[1].stats, p[2].stats]) for p in pending]))
        else:
            stats = np.stack([np.stack([np.asarray(p[1].stats), np.asarray(p[2].stats)]) for p in pending])
        for (j, ea, eb, tja, tjb), (sa, sb) in zip(pending, stats):
            log.append({"iteration": j, "appended": (int(sa[0]), int(sb[0])), "connected": int(sb[1])})
            if sa[3] or sb[3]:
                raise RuntimeError("tree capacity exhausted")
            if sb[1] > 0:
                e = int(sb[2])
                hit = (int(_np(ea.new_index)[e]), int(_np(eb.new_index)[e]))
                it, ta, tb = j, tja, tjb
                break
        pending = []
        if hit is not None:
            break
    else:
        it = max_iterations
    if hit is None:
        return None
    da, db = ta.download(), tb.download()
    pa, ma = dense_chain(checker, da, _chain(da["parent"], hit[0]))
    pb, mb = dense_chain(checker, db, _chain(db["parent"], hit[1]))
    # ta's chain runs root_a .. meeting node, tb's root_b .. meeting node (same configuration): join them
    if ta is t_start:
        first, second, m1, m2 = pa, pb, ma, mb
    else:
        first, second, m1, m2 = pb, pa, mb, ma
    dense = _drop_repeats(np.concatenate([first, second[::-1][1:]], axis=0))
    way = _drop_repeats(np.concatenate([first[m1], second[m2][::-1][1:]], axis=0))
    to_conf = lambda q: tuple(float(v) for v in q)
    sizes = (len(da["parent"]), len(db["parent"])) if ta is t_start else (len(db["parent"]), len(da["parent"]))
    return BatchedPlan([to_conf(q) for q in dense], [to_conf(q) for q in way], it, sizes, 2 * batch * it, log)


@dataclass
class RankedPlan:
    path: List[Conf]                   # waypoints of the selected candidate (straight joint-space segments)
    ee_path_cost: float                # sum of end-effector segment lengths (the reference's ranking key)
    deflection_over_limit: np.ndarray  # [W]
    total_cost: np.ndarray             # [W] ee_path_cost + alpha * over
    candidates: List[dict]             # per restart: waypoints, valid, ee_path_cost, over, seed


def batched_restarts(checker, start: Sequence[float], goal: Sequence[float], *, restarts: int = 8, batch: int = 256,
                     max_iterations: int = 200, smooth_rounds: int = 10, proposals: int = 256, seed: int = 0,
                     alpha: Optional[float] = None) -> Optional[RankedPlan]:
    """random_restarts + smoothing + forward re-check + end-effector-length selection
    (motion_planners/motion_planners/meta.py:188-267; cost terms of Command.execute_multi_world,
    pybullet_tools/kuka_primitives.py:467-521) in batched form: ``restarts`` independent batched RRT-Connect runs,
    each shortened by the batched smoother, then ONE pmp_validate_edges launch over every segment of every candidate
    gives the forward re-check (first_hit >= n_steps on all segments), the end-effector path length and the
    deflection-over-limit of all candidates at once.  The valid candidate with the shortest end-effector path wins
    (the reference's ranking key)."""
    from . import birrt, planner_fns as F
    distance_fn = F.get_distance_fn(checker.robot)
    cands = []
    for r in range(restarts):
        plan = batched_rrt_connect(checker, start, goal, batch=batch, max_iterations=max_iterations, seed=seed + r)
        if plan is None:
            continue
        way = birrt.smooth_path_batched(checker, plan.waypoints, distance_fn, rounds=smooth_rounds, proposals=proposals,
                                        seed=seed + r, gate=False)
        cands.append({"seed": seed + r, "waypoints": way, "iterations": plan.iterations})
    if not cands:
        return None
    q0 = np.concatenate([np.asarray(c["waypoints"][:-1], dtype=np.float32) for c in cands])
    q1 = np.concatenate([np.asarray(c["waypoints"][1:], dtype=np.float32) for c in cands])
    owner = np.concatenate([np.full(len(c["waypoints"]) - 1, k) for k, c in enumerate(cands)])
    out = checker.validate_edges_host(q0, q1)
    ok_seg = out["first_hit"] >= out["n_steps"]
    a = float(checker.config.alpha if alpha is None else alpha)
    for k, c in enumerate(cands):
        m = owner == k
        c["valid"] = bool(ok_seg[m].all())
        c["ee_path_cost"] = float(out["ee_len"][m].astype(np.float64).sum())
        c["over"] = out["over_limit"][m].astype(np.float64).sum(axis=0)
    valid = [c for c in cands if c["valid"]]
    if not valid:
        return None
    best = min(valid, key=lambda c: c["ee_path_cost"])
    return RankedPlan(best["waypoints"], best["ee_path_cost"], best["over"], best["ee_path_cost"] + a * best["over"], cands)
