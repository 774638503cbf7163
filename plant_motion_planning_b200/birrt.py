"""Multi-world BiRRT driver on top of the GPU edge checker  (the callers of the hot path).

Control flow restated from the lab's fork so that decisions can be compared edge by edge:
  extend_towards ............ motion_planners/motion_planners/primitives.py:196-255 (extend_towards_with_angle_constraint_multi_world)
  check_forward_path ........ motion_planners/motion_planners/rrt_connect_with_angle_constraints_v2.py:143-198
  rrt_connect_multi_world ... same file :202-280 (swap alternation, max_iterations forced to 2000)
  forward re-check + EE length  motion_planners/motion_planners/meta.py:222-248
What changes is the cost of one extension: the reference spends N x (50 stepSimulation + W collision closures);
here one extension is ONE kernel launch over its N configurations x W worlds, and `extend_many` validates a whole
batch of candidate extensions in one launch (additive API).

The random 5 % pass-through of exceeded deflections (pybullet_tools/utils.py:3843) is replayed on the host from the
raw per-(step, world) flags, drawing np.random.rand() in the reference's order, so runs are reproducible from the
same seeds regardless of the backend that produced the flags.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .edgecheck import FLAG_EXCEEDED, FLAG_RIGID
from .planner_fns import GATE_PROBABILITY, asymmetric_extend

Conf = Tuple[float, ...]


class TreeNode:
    """TreeNode_v2 without the Bullet snapshot (motion_planners/motion_planners/rrt.py:7-35)."""
    __slots__ = ("config", "parent")

    def __init__(self, config: Conf, parent: Optional["TreeNode"] = None):
        self.config = config
        self.parent = parent

    def retrace(self) -> List["TreeNode"]:
        seq, node = [], self
        while node is not None:
            seq.append(node)
            node = node.parent
        return seq[::-1]


class ConfigTree(list):
    """A list of TreeNode that also keeps the configurations in one float64 array, so that the nearest-node search
    of every extension (`argmin(distance_fn(n.config, target))`, primitives.py:207 -- an O(|tree|) Python loop in the
    reference) is one vectorised pass.  Ties resolve to the first minimum, like the reference's argmin
    (motion_planners/motion_planners/utils.py:47-53)."""

    def __init__(self, nodes=(), circular=None):
        super().__init__()
        self._buf = None
        self._circ = None if circular is None else np.asarray(circular, dtype=bool)
        for n in nodes:
            self.append(n)

    def append(self, node: TreeNode) -> None:
        k = len(self)
        if self._buf is None:
            self._buf = np.empty((64, len(node.config)), dtype=np.float64)
        elif k == self._buf.shape[0]:
            self._buf = np.concatenate([self._buf, np.empty_like(self._buf)], axis=0)
        self._buf[k] = node.config
        super().append(node)

    def extend(self, nodes) -> None:
        for n in nodes:
            self.append(n)

    def clear(self) -> None:
        super().clear()

    def nearest(self, target: Conf) -> TreeNode:
        diff = np.asarray(target, dtype=np.float64)[None] - self._buf[:len(self)]
        if self._circ is not None and self._circ.any():
            diff = np.where(self._circ[None], (diff + np.pi) % (2 * np.pi) - np.pi, diff)
        return self[int(np.argmin(np.einsum("ij,ij->i", diff, diff)))]


class FlagBackend:
    """Anything that maps a list of configurations to raw flags uint8 [n, W] (bit0 rigid, bit1 exceeded)."""

    def flags(self, configs: Sequence[Conf]) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError


class GpuBackend(FlagBackend):
    def __init__(self, checker):
        self.checker = checker
        self.launches = 0

    def flags(self, configs: Sequence[Conf]) -> np.ndarray:
        self.launches += 1
        q = np.asarray(configs, dtype=np.float32).reshape(-1, self.checker.dof)
        return self.checker.check_configs_host(q, detail=False)["flags"]


def first_violation(flags: np.ndarray, forward: bool, probability: float = GATE_PROBABILITY,
                    rand: Optional[Callable[[], float]] = None) -> int:
    """Index of the first configuration the reference's checker would reject (len(flags) if none):
    worlds in construction order with early exit (env.py:230-247), forward checker = rigid | (exceeded & gate),
    backward checker = rigid only."""
    rand = np.random.rand if rand is None else rand
    n, W = flags.shape
    for i in range(n):
        for w in range(W):
            f = flags[i, w]
            if f & FLAG_RIGID:
                return i
            if forward and (f & FLAG_EXCEEDED) and rand() < probability:
                return i
    return n


@dataclass
class PlanLog:
    """Every edge the planner proposed: (kind, n configurations, first violation)."""
    edges: List[Tuple[str, int, int]] = field(default_factory=list)
    iterations: int = 0


def extend_towards(tree: List[TreeNode], target: Conf, distance_fn, extend_fn, backend: FlagBackend, forward: bool,
                   swap: bool, log: Optional[PlanLog] = None) -> Tuple[TreeNode, bool]:
    if isinstance(tree, ConfigTree):
        last = tree.nearest(target)              # unit weights, 2-norm: the reference's default distance_fn
    else:
        scores = [distance_fn(n.config, target) for n in tree]
        last = tree[scores.index(min(scores))]
    extend = list(asymmetric_extend(last.config, target, extend_fn, backward=swap))
    k = first_violation(backend.flags(extend), forward) if extend else 0
    if log is not None:
        log.edges.append(("fwd" if forward else "back", len(extend), k))
    for q in extend[:k]:
        last = TreeNode(q, parent=last)
        tree.append(last)
    return last, k == len(extend)


def check_forward_path(path2: List[TreeNode], backend: FlagBackend, nodes1: List[TreeNode], nodes2: List[TreeNode],
                       log: Optional[PlanLog] = None) -> bool:
    """Re-validate the backward tree's half with the forward (plant-aware) checker."""
    last = nodes1[-1]
    nodes2.clear()
    tail = [n.config for n in path2[1:]]
    k = first_violation(backend.flags(tail), True) if tail else 0
    if log is not None:
        log.edges.append(("recheck", len(tail), k))
    for q in tail[:k]:
        last = TreeNode(q, parent=last)
        nodes1.append(last)
    if k < len(tail):
        nodes2.extend(path2[k + 1:])
        return False
    return True


def is_default_metric(distance_fn) -> bool:
    """True when `distance_fn` is known to be planner_fns.get_distance_fn with unit weights and the 2-norm (it carries
    a `default_metric` attribute); anything else is evaluated node by node, like the reference's argmin."""
    return bool(getattr(distance_fn, "default_metric", False))


def euclidean_distance(q1: Conf, q2: Conf) -> float:
    """motion_planners/motion_planners/utils.py:159-160 get_distance: the plain 2-norm of q2 - q1 (no wrap, no weights).
    The reference's restart driver calls smooth_path_multi_world WITHOUT a distance_fn (meta.py:213-215), so the
    smoother falls back to this (smoothing.py:66-67) -- it sets the np.random.choice probabilities."""
    return float(np.linalg.norm(np.asarray(q2, dtype=np.float64) - np.asarray(q1, dtype=np.float64)))


def rrt_connect_multi_world(start: Conf, goal: Conf, distance_fn, sample_fn, extend_fn, backend: FlagBackend,
                            max_iterations: int = 2000, log: Optional[PlanLog] = None, circular=None,
                            vectorized_nearest: Optional[bool] = None):
    """Returns (path of configurations, node path) or (None, None).  `vectorized_nearest` assumes the default metric
    (unit weights, 2-norm, wrap on `circular` joints); pass False to call `distance_fn` node by node."""
    if first_violation(backend.flags([start]), True) == 0:
        return None
    if vectorized_nearest is None:
        vectorized_nearest = is_default_metric(distance_fn)
    if circular is None:
        circular = getattr(distance_fn, "circular", None)
    if vectorized_nearest:
        nodes1, nodes2 = ConfigTree([TreeNode(start)], circular), ConfigTree([TreeNode(goal)], circular)
    else:
        nodes1, nodes2 = [TreeNode(start)], [TreeNode(goal)]
    swap = True
    for it in range(max_iterations):
        if log is not None:
            log.iterations = it + 1
        swap = not swap
        target = sample_fn()
        last1, _ = extend_towards(nodes1, target, distance_fn, extend_fn, backend, True, swap, log)
        last2, success = extend_towards(nodes2, last1.config, distance_fn, extend_fn, backend, False, not swap, log)
        if success:
            path1, path2 = last1.retrace(), last2.retrace()[::-1]
            if check_forward_path(path2, backend, nodes1, nodes2, log):
                node_path = path1[:-1] + path2
                return [n.config for n in node_path], node_path
            if not nodes2:
                # the reference leaves tree2 with the unreached remainder; if nothing remains, regrow from the goal
                nodes2.append(TreeNode(goal))
            swap = True
    return None, None


def plan(env, start: Conf, goal: Conf, distance_fn, sample_fn, extend_fn, backend: Optional[FlagBackend] = None,
         max_iterations: int = 2000, log: Optional[PlanLog] = None):
    """check_initial_end_multi_world + rrt_connect + forward re-check with EE path length
    (pybullet_tools/utils.py:4358-4380; meta.py:222-248).  Returns (path, ee_len, over[W]) or None."""
    backend = backend or GpuBackend(env.checker)
    for name, q in (("Initial", start), ("End", goal)):
        if first_violation(backend.flags([q]), True) == 0:
            raise ValueError(f"{name} configuration is in collision")       # the reference exit()s here
    # the vectorised nearest-node search is the DEFAULT metric (unit weights, 2-norm, wrap on continuous joints);
    # it needs the robot's circular mask, else a wrapped joint would pick a different nearest node than distance_fn
    circular = getattr(getattr(env, "sample_robot", None), "circular", None)
    res = rrt_connect_multi_world(start, goal, distance_fn, sample_fn, extend_fn, backend, max_iterations, log,
                                  circular=circular, vectorized_nearest=is_default_metric(distance_fn))
    if res is None or res[0] is None:
        return None
    path = res[0]
    k = first_violation(backend.flags(path), True)
    if log is not None:
        log.edges.append(("final", len(path), k))
    if k < len(path):
        return None
    ee_len, over, _ = env.path_cost(start, path)
    return path, ee_len, over


# ------------------------------------------------------------------------------------------------ shortcut smoothing
def _path_length(waypoints: Sequence[Conf], distance_fn) -> float:
    return sum(distance_fn(a, b) for a, b in zip(waypoints[:-1], waypoints[1:]))


def waypoints_from_path(path: Sequence[Conf], tolerance: float = 1e-3) -> List[Conf]:
    """Corner points of a dense path (motion_planners/motion_planners/utils.py:229-260, waypoints_from_path_v4):
    consecutive duplicates dropped, a waypoint kept wherever the unit direction changes by more than `tolerance`."""
    pts = [tuple(path[0])]
    for q in path[1:]:
        if not np.allclose(q, pts[-1], atol=tolerance, rtol=0):
            pts.append(tuple(q))
    if len(pts) < 3:
        return pts
    unit = lambda a, b: (np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(a) - np.asarray(b)), 1e-300)
    way = [pts[0]]
    last_conf = pts[1]
    last_dir = unit(last_conf, way[-1])
    for conf in pts[2:]:
        d = unit(conf, way[-1])
        if not np.allclose(last_dir, d, atol=tolerance, rtol=0):
            way.append(last_conf)
            d = unit(conf, way[-1])
        last_conf, last_dir = conf, d
    way.append(last_conf)
    return way


def smooth_path_multi_world(path: Sequence[Conf], extend_fn, distance_fn, backend: FlagBackend, max_iterations: int = 20,
                            log: Optional[PlanLog] = None) -> List[Conf]:
    """Shortcut smoother of the lab's fork (motion_planners/motion_planners/smoothing.py:55-141): per iteration two
    segments are drawn with probability proportional to their length (np.random.choice), a point is drawn on each
    (random.random), and the shortcut between them is kept if it shortens the path and passes the forward checker.
    One kernel launch per tried shortcut."""
    return _smooth(path, extend_fn, distance_fn, backend, max_iterations, True, log)


def _smooth(path: Sequence[Conf], extend_fn, distance_fn, backend: FlagBackend, max_iterations: int, forward: bool,
            log: Optional[PlanLog]) -> List[Conf]:
    import random as _random
    way = waypoints_from_path(path)
    for _ in range(max_iterations):
        if len(way) <= 2:
            break
        segs = list(zip(range(len(way) - 1), range(1, len(way))))
        dists = [distance_fn(way[i], way[j]) for i, j in segs]
        total = sum(dists)
        i1, i2 = np.random.choice(list(range(len(segs))), size=2, replace=True, p=np.array(dists) / total)
        if i1 == i2:
            continue
        if i2 < i1:
            i1, i2 = i2, i1
        pts = []
        for (i, j) in (segs[i1], segs[i2]):
            w = _random.random()
            pts.append(tuple((1 - w) * np.array(way[i]) + w * np.array(way[j])))
        p1, p2 = pts
        new_way = way[:segs[i1][0] + 1] + [p1, p2] + way[segs[i2][1]:]
        if _path_length(new_way, distance_fn) >= total:
            continue
        cfgs = list(extend_fn(p1, p2))
        k = first_violation(backend.flags(cfgs), forward) if cfgs else 0
        if log is not None:
            log.edges.append(("shortcut", len(cfgs), k))
        if k == len(cfgs):
            way = new_way
    out: List[Conf] = []          # refine_waypoints (smoothing.py:49-52): the first waypoint itself is not emitted
    for a, b in zip(way[:-1], way[1:]):
        out.extend(extend_fn(a, b))
    return out


def smooth_path_batched(checker, path: Sequence[Conf], distance_fn, rounds: int = 10, proposals: int = 256,
                        seed: int = 0, gate: bool = True) -> List[Conf]:
    """Additive batched smoother: every round proposes `proposals` random shortcuts between points of the current
    waypoint polyline, validates ALL of them in one pmp_validate_edges launch and applies the valid one that shortens
    the path most.  ``gate=True`` replays the reference's 5 % pass-through per edge in batch order (needs the step
    masks); ``gate=False`` treats every exceeded deflection as a violation and uses the first-violation-only call."""
    rng = np.random.RandomState(seed)
    way = [np.asarray(w, dtype=np.float64) for w in waypoints_from_path(path)]
    circ = np.asarray(getattr(getattr(checker, "robot", None), "circular", np.zeros(len(way[0]), dtype=bool)), dtype=bool)

    def dist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """get_distance_fn (weights 1, 2-norm of difference_fn) for arrays of configurations."""
        d = b - a
        if circ.any():
            d = np.where(circ, (d + np.pi) % (2 * np.pi) - np.pi, d)
        return np.sqrt(np.sum(d * d, axis=-1))

    for _ in range(rounds):
        if len(way) <= 2:
            break
        W = np.stack(way)
        seg_len = dist(W[:-1], W[1:])
        cum = np.concatenate([[0.0], np.cumsum(seg_len)])
        s = np.sort(rng.uniform(0.0, cum[-1], size=(proposals, 2)), axis=1)
        seg = np.clip(np.searchsorted(cum, s, side="right") - 1, 0, len(seg_len) - 1)
        frac = (s - cum[seg]) / np.maximum(seg_len[seg], 1e-300)
        P = W[seg] * (1 - frac[..., None]) + W[seg + 1] * frac[..., None]          # [proposals, 2, D]
        direct = dist(P[:, 0], P[:, 1])
        gain = (s[:, 1] - s[:, 0]) - direct
        ok = (seg[:, 0] != seg[:, 1]) & (gain > 1e-9)
        if not ok.any():
            continue
        idx = np.nonzero(ok)[0]
        if gate:
            first, n, _ = extend_many(checker, P[idx, 0].astype(np.float32), P[idx, 1].astype(np.float32))
        else:
            r = checker.validate_edges_host(P[idx, 0].astype(np.float32), P[idx, 1].astype(np.float32),
                                            first_hit_only=True)
            first, n = r["first_hit"], r["n_steps"]
        valid = first >= n
        if not valid.any():
            continue
        best = idx[valid][np.argmax(gain[idx][valid])]
        way = way[:seg[best, 0] + 1] + [P[best, 0], P[best, 1]] + way[seg[best, 1] + 1:]
    return [tuple(float(v) for v in w) for w in way]


# ------------------------------------------------------------------------------------------------ restarts / Planner
def random_restarts_multi_world(start: Conf, goal: Conf, distance_fn, sample_fn, extend_fn, backend: FlagBackend,
                                max_solutions: int = 1, smooth_attempts: int = 10, smoothing_iterations: int = 20,
                                rrt_iterations: int = 2000, circular=None, log: Optional[PlanLog] = None):
    """motion_planners/motion_planners/meta.py:156-273: solve, then up to 10 rounds of (20-iteration shortcut smoothing +
    forward re-check of the smoothed path); the smoothed path replaces the raw one only if its re-check passes.
    With max_solutions = 1 (birrt_multi_world, rrt_connect_..._v2.py:291-292) the end-effector length the reference
    accumulates for ranking never changes the answer, so it is not computed here."""
    solutions: List[List[Conf]] = []
    while len(solutions) < max_solutions:
        res = rrt_connect_multi_world(start, goal, distance_fn, sample_fn, extend_fn, backend, rrt_iterations, log,
                                      circular=circular)
        if res is None or res[0] is None:
            continue
        path = res[0]
        accepted = None
        for _ in range(smooth_attempts):
            # the reference passes no distance_fn here (meta.py:213-215): plain Euclidean get_distance
            sm = smooth_path_multi_world(path, extend_fn, euclidean_distance, backend, max_iterations=smoothing_iterations, log=log)
            k = first_violation(backend.flags(sm), True) if sm else 0
            if log is not None:
                log.edges.append(("resmooth-check", len(sm), k))
            if k == len(sm):
                accepted = sm
                break
        solutions.append(accepted if accepted is not None else path)
    return solutions[0]


def move_arm_conf2conf_multi_world(start: Conf, goal: Conf, distance_fn, sample_fn, extend_fn, backend: FlagBackend,
                                   num_attempts: int = 200, circular=None, log: Optional[PlanLog] = None):
    """Planner.move_arm_conf2conf_multi_world (planner.py:18-51) -> get_free_motion_gen_with_angle_constraints_multi_world
    (kuka_primitives.py:1069-1107) -> plan_joint_motion_with_angle_constraints_multi_world (pybullet_tools/utils.py:4506-4567):
    per attempt the start and goal are re-checked (the reference exit()s if either is rejected; here ValueError), then
    birrt_multi_world = random restarts with max_solutions 1.  Returns the path (list of configurations) or None."""
    for _ in range(num_attempts):
        for name, q in (("Initial", start), ("End", goal)):
            if first_violation(backend.flags([q]), True) == 0:
                raise ValueError(f"Error! {name} configuration is in collision")
        path = random_restarts_multi_world(start, goal, distance_fn, sample_fn, extend_fn, backend, circular=circular, log=log)
        if path is not None:
            return path
    return None


# ------------------------------------------------------------------------------------------------ benchmark chain
def rrt_connect_benchmark(start: Conf, goal: Conf, distance_fn, sample_fn, extend_fn, backend: FlagBackend,
                          max_iterations: int = 20, log: Optional[PlanLog] = None, circular=None):
    """rrt_connect_multiworld_benchmark (motion_planners/motion_planners/rrt_connect.py:62-91): classic RRT-connect --
    the smaller tree is extended first, both extensions use the rigid-obstacle checker
    (net_constraint_violation_checker_benchmark), RRT_ITERATIONS = 20 iterations per call."""
    nodes1, nodes2 = ConfigTree([TreeNode(start)], circular), ConfigTree([TreeNode(goal)], circular)
    for it in range(max_iterations):
        if log is not None:
            log.iterations += 1
        swap = len(nodes1) > len(nodes2)
        tree1, tree2 = (nodes2, nodes1) if swap else (nodes1, nodes2)
        target = sample_fn()
        last1, _ = extend_towards(tree1, target, distance_fn, extend_fn, backend, False, swap, log)
        last2, success = extend_towards(tree2, last1.config, distance_fn, extend_fn, backend, False, not swap, log)
        if success:
            path1, path2 = last1.retrace(), last2.retrace()
            if swap:
                path1, path2 = path2, path1
            return [n.config for n in path1[:-1] + path2[::-1]]
    return None


def smooth_path_benchmark(path: Sequence[Conf], extend_fn, backend: FlagBackend, max_iterations: int = 5,
                          log: Optional[PlanLog] = None) -> List[Conf]:
    """smooth_path_multiworld_benchmark (smoothing.py:651-707): as smooth_path_multi_world but with the Euclidean
    metric, the rigid checker (no gate, so the bisect visiting order of the reference does not matter) and
    RRT_SMOOTHING = 5 iterations."""
    euclid = lambda a, b: float(np.linalg.norm(np.asarray(b) - np.asarray(a)))
    return _smooth(path, extend_fn, euclid, backend, max_iterations, False, log)


def random_restarts_benchmark(start: Conf, goal: Conf, distance_fn, sample_fn, extend_fn, backend: FlagBackend,
                              restarts: int = 2, smooth: int = 5, rrt_iterations: int = 20, circular=None,
                              log: Optional[PlanLog] = None):
    """random_restarts_multiworld_benchmark (meta.py:89-118) with max_solutions = 1: up to restarts + 1 solves, the first
    solution is smoothed and returned."""
    for _ in range(restarts + 1):
        path = rrt_connect_benchmark(start, goal, distance_fn, sample_fn, extend_fn, backend, rrt_iterations, log, circular)
        if path is None:
            continue
        return smooth_path_benchmark(path, extend_fn, backend, smooth, log)
    return None


def move_arm_conf2conf_multi_world_benchmark(start: Conf, goal: Conf, distance_fn, sample_fn, extend_fn,
                                             backend: FlagBackend, num_attempts: int = 200, circular=None,
                                             log: Optional[PlanLog] = None):
    """Planner.move_arm_conf2conf_multi_world_benchmark (planner.py:53-78) -> get_free_motion_gen_multiworld_benchmark
    (kuka_primitives.py:1019-1039) -> plan_joint_motion_multiworld_benchmark (pybullet_tools/utils.py:4483-4503): what
    the avoid_all / ignore_all scripts run.  Rigid checker everywhere; start/goal are re-checked every attempt."""
    for _ in range(num_attempts):
        for name, q in (("Initial", start), ("End", goal)):
            if first_violation(backend.flags([q]), False) == 0:
                raise ValueError(f"Error! {name} configuration is in collision")
        path = random_restarts_benchmark(start, goal, distance_fn, sample_fn, extend_fn, backend, circular=circular, log=log)
        if path is not None:
            return path
    return None


# ------------------------------------------------------------------------------------------------ batched proposals
def extend_many(checker, q_from: np.ndarray, q_to: np.ndarray, backward: bool = False, forward_checker: bool = True,
                probability: float = GATE_PROBABILITY, rand: Optional[Callable[[], float]] = None):
    """Validate a batch of candidate extensions / shortcuts in ONE launch (pmp_validate_edges_host) and apply the
    gate per edge in batch order.  Returns (first_violation[E], n_steps[E], result dict)."""
    from .planner_fns import replay_gate
    ms = int(checker.config.max_steps)
    out = checker.validate_edges_host(q_from, q_to, backward=backward, want_masks=True, max_steps=ms)
    E = out["n_steps"].shape[0]
    cap = ms          # steps >= max_steps are not evaluated (their mask bits are 0): never count them as safe
    n_all = np.minimum(out["n_steps"], cap).astype(np.int32)
    # edges without any "deflection exceeded" bit draw no random number: their first violation is the lowest set bit
    # of the rigid masks OR-ed over the worlds (vectorised); only the others need the sequential gate replay
    rig = np.bitwise_or.reduce(out["rigid_mask"].astype(np.uint32), axis=1)                  # [E, C]
    exc_any = out["exceed_mask"].any(axis=(1, 2)) if forward_checker else np.zeros(E, dtype=bool)
    word_has = rig != 0
    first_word = np.argmax(word_has, axis=1)
    w = rig[np.arange(E), first_word]
    low = (w & (~w + np.uint32(1))).astype(np.float64)                                       # isolated lowest set bit
    bit = np.where(w != 0, np.log2(np.maximum(low, 1.0)), 0).astype(np.int32)
    first = np.where(word_has.any(axis=1), np.minimum(first_word * 32 + bit, n_all), n_all).astype(np.int32)
    for e in np.nonzero(exc_any)[0]:
        first[e] = replay_gate(int(n_all[e]), out["rigid_mask"][e], out["exceed_mask"][e], probability, rand,
                               forward_checker)
    return first, out["n_steps"], out
