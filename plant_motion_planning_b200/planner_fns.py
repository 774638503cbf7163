"""Host-side mirror of the reference's planner callables (same names, argument meaning and results).

The reference builds them from a PyBullet body id (pybullet_tools/utils.py:3593-3676, 3753-3762); here the
first argument is a RobotModel (or anything with .lower/.upper/.circular/.dof).  They are float64 host
arithmetic on tuples, exactly like the reference: the RRT loop calls them once per node and they are not
the hot path -- the hot path is what collision_fn / MultiPlantWorld.step dispatch to the GPU.

    sample_fn()            -> tuple            get_sample_fn            :3593-3599
    difference_fn(q2, q1)  -> tuple            get_difference_fn        :3604-3610
    distance_fn(q1, q2)    -> float            get_distance_fn          :3619-3627
    refine_fn(q1, q2)      -> iterator[tuple]  get_refine_fn            :3641-3651
    extend_fn(q1, q2)      -> iterator[tuple]  get_extend_fn            :3667-3676
    limits_fn(q)           -> bool             get_limits_fn            :3753-3762
    asymmetric_extend(...)                     motion_planners/motion_planners/primitives.py:16-19
"""
from __future__ import annotations

import math
from typing import Callable, Iterator, Optional, Sequence, Tuple

import numpy as np

PI = math.pi
DEFAULT_RESOLUTION = math.radians(3)      # pybullet_tools/utils.py:3660
MAX_DISTANCE = 0.0                        # pybullet_tools/utils.py:3317
GATE_PROBABILITY = 0.95                   # pybullet_tools/utils.py:3843


def wrap_interval(value: float, lower: float = -PI, upper: float = PI) -> float:
    return (value - lower) % (upper - lower) + lower


def circular_difference(theta2: float, theta1: float) -> float:
    return wrap_interval(theta2 - theta1)


def _circular(robot) -> Sequence[bool]:
    return [bool(c) for c in robot.circular]


def get_custom_limits(robot, custom_limits=None, circular_limits=(-np.inf, np.inf)):
    """(lower, upper) with circular joints replaced by ``circular_limits`` (utils.py:2128-2137)."""
    custom_limits = custom_limits or {}
    lower, upper = [], []
    for j in range(robot.dof):
        if j in custom_limits:
            lo, hi = custom_limits[j]
        elif robot.circular[j]:
            lo, hi = circular_limits
        else:
            lo, hi = float(robot.lower[j]), float(robot.upper[j])
        lower.append(lo)
        upper.append(hi)
    return tuple(lower), tuple(upper)


def get_sample_fn(robot, custom_limits=None) -> Callable[[], Tuple[float, ...]]:
    lower, upper = get_custom_limits(robot, custom_limits, circular_limits=(-PI, PI))
    lower, upper = np.array(lower), np.array(upper)

    def fn():
        w = np.random.uniform(size=len(lower))          # one global-stream draw per sample (utils.py:3549-3551)
        return tuple((1 - w) * lower + w * upper)
    return fn


def get_difference_fn(robot) -> Callable:
    circ = _circular(robot)

    def fn(q2, q1):
        return tuple(circular_difference(v2, v1) if c else (v2 - v1) for c, v2, v1 in zip(circ, q2, q1))
    return fn


def get_distance_fn(robot, weights=None, norm=2) -> Callable:
    default_metric = weights is None and norm == 2
    weights = 1 * np.ones(robot.dof) if weights is None else weights
    difference_fn = get_difference_fn(robot)

    def fn(q1, q2):
        diff = np.array(difference_fn(q2, q1))
        if norm == 2:
            return np.sqrt(np.dot(weights, diff * diff))
        return np.linalg.norm(np.multiply(weights, diff), ord=norm)
    # lets the batched / vectorised nearest-node searches recognise the metric they implement (unit weights, 2-norm,
    # wrap on continuous joints); any other distance_fn is evaluated node by node
    fn.default_metric = default_metric
    fn.circular = np.asarray(_circular(robot), dtype=bool)
    return fn


def get_refine_fn(robot, num_steps=0) -> Callable:
    difference_fn = get_difference_fn(robot)
    any_circular = any(_circular(robot))
    num_steps = num_steps + 1

    def fn(q1, q2) -> Iterator[Tuple[float, ...]]:
        # the reference's `(1. / (num_steps - i)) * np.array(difference_fn(q2, q)) + q` element by element in Python
        # floats: the same IEEE operations in the same order (bit-identical results, pinned by
        # tests/golden/reference_planner_fns.json) without a numpy round trip per step
        q = tuple(float(v) for v in q1)
        if not any_circular:                  # difference_fn is a plain subtraction: skip the call per step
            target = tuple(q2)
            for i in range(num_steps):
                c = 1. / (num_steps - i)
                q = tuple(c * (b - a) + a for a, b in zip(q, target))
                yield q
            return
        for i in range(num_steps):
            c = 1. / (num_steps - i)
            q = tuple(c * d + v for d, v in zip(difference_fn(q2, q), q))
            yield q
    return fn


def get_extend_fn(robot, resolutions=None, norm=2) -> Callable:
    resolutions = DEFAULT_RESOLUTION * np.ones(robot.dof) if resolutions is None else resolutions
    difference_fn = get_difference_fn(robot)

    def fn(q1, q2):
        steps = int(np.linalg.norm(np.divide(difference_fn(q2, q1), resolutions), ord=norm))
        return get_refine_fn(robot, num_steps=steps)(q1, q2)
    return fn


def get_limits_fn(robot, custom_limits=None) -> Callable:
    lower, upper = get_custom_limits(robot, custom_limits)

    def limits_fn(q):
        return not (np.less_equal(lower, q).all() and np.less_equal(q, upper).all())
    return limits_fn


def asymmetric_extend(q1, q2, extend_fn, backward=False):
    if backward:
        return reversed(list(extend_fn(q2, q1)))
    return extend_fn(q1, q2)


# ------------------------------------------------------------------------------------------------ gate replay
def replay_gate(n_steps: int, rigid_mask: np.ndarray, exceed_mask: np.ndarray, probability: float = GATE_PROBABILITY,
                rand: Optional[Callable[[], float]] = None, forward: bool = True) -> int:
    """First violating step of one edge under the reference's *randomised* deflection gate.

    The reference reports an exceeded deflection only if ``np.random.rand() < 0.95`` (utils.py:3843), drawing
    one number per reached (step, world) whose rigid test passed and whose deflection test failed, worlds in
    construction order with early exit (env.py:230-238), steps in order with early exit (primitives.py:225-241).
    ``rigid_mask`` / ``exceed_mask``: uint32 [W, C] step bitmasks from validate_edges(want_masks=True).
    ``forward=False`` = the backward-tree checker, which ignores plants (env.py:241-247).
    Returns the first violating step, or n_steps.
    """
    rand = np.random.rand if rand is None else rand
    W = rigid_mask.shape[0]
    for s in range(n_steps):
        word, bit = s >> 5, np.uint32(1) << np.uint32(s & 31)
        for w in range(W):
            if rigid_mask[w, word] & bit:
                return s
            if forward and (exceed_mask[w, word] & bit) and rand() < probability:
                return s
    return n_steps


def gate_hash(seed: int, e, s, w):
    """include/pmp_edgecheck.h:pmp_gate_hash -- the counter-based generator of the on-device gate (numpy, vectorised)."""
    u = np.uint32
    with np.errstate(over="ignore"):
        h = u(seed) ^ (np.asarray(e, dtype=np.uint32) * u(0x9E3779B1)) ^ (np.asarray(s, dtype=np.uint32) * u(0x85EBCA77)) \
            ^ (np.asarray(w, dtype=np.uint32) * u(0xC2B2AE3D))
        h = h ^ (h >> u(16))
        h = h * u(0x85EBCA6B)
        h = h ^ (h >> u(13))
        h = h * u(0xC2B2AE35)
        h = h ^ (h >> u(16))
    return h


def device_gate_first_hit(n_steps: np.ndarray, rigid_mask: np.ndarray, exceed_mask: np.ndarray, seed: int,
                          probability: float = GATE_PROBABILITY, max_steps: Optional[int] = None) -> np.ndarray:
    """What validate_edges(device_gate_seed=seed) reports as first_hit, recomputed on the host from the raw masks."""
    E, W, C = rigid_mask.shape
    cap = C * 32 if max_steps is None else max_steps
    thr = 0xFFFFFFFF if probability >= 1.0 else int(probability * 4294967296.0)
    steps = np.arange(C * 32, dtype=np.uint32)
    bit = lambda m: ((m[..., steps >> 5] >> (steps & 31).astype(np.uint32)) & np.uint32(1)).astype(bool)   # [E, W, S]
    r, x = bit(rigid_mask.astype(np.uint32)), bit(exceed_mask.astype(np.uint32))
    h = gate_hash(seed, np.arange(E, dtype=np.uint32)[:, None, None], steps[None, None, :],
                  np.arange(W, dtype=np.uint32)[None, :, None])
    viol = (r | (x & (h < np.uint32(thr)))).any(axis=1)                                                   # [E, S]
    out = np.minimum(n_steps, cap).astype(np.int32)
    has = viol.any(axis=1)
    out[has] = np.argmax(viol[has], axis=1)
    return out
