"""Synthetic workloads of BASELINE.json: seeded edges and sampled worlds."""
from __future__ import annotations

import math
import random
from typing import List, Tuple

import numpy as np

from .model.plants import PlantWorld, sample_world
from .model.robot import RobotModel

SWEEP_XY_LIMITS = ((0.0, 0.35), (-0.5, 0.0))    # scripts/multi_world_ignore_all.py:70


def synthetic_worlds(n_worlds: int, seed_base: int = 1000, branches: int = 1, vert_stems: int = 2,
                     extensions: int = 1, xy_limits=SWEEP_XY_LIMITS, stem_base_spacing: float = 0.0) -> List[PlantWorld]:
    """World w = the reference's sampler seeded with np.random.seed(seed_base+w); random.seed(seed_base+w)."""
    out = []
    for w in range(n_worlds):
        out.append(sample_world(branches, vert_stems, extensions, xy_limits, stem_base_spacing=stem_base_spacing,
                                np_random=np.random.RandomState(seed_base + w), py_random=random.Random(seed_base + w)))
    return out


def synthetic_edges(model: RobotModel, n_edges: int, seed: int = 0, steps: int = 32,
                    resolution: float = math.radians(3)) -> Tuple[np.ndarray, np.ndarray]:
    """q_from ~ U(limits); direction = normalised N(0, I); length (steps - 0.5) * resolution so that every
    edge has exactly ``steps`` interpolation steps; edges whose end leaves the limits are redrawn."""
    rng = np.random.RandomState(seed)
    lo, hi = model.lower, model.upper
    D = model.dof
    q0 = np.empty((n_edges, D))
    q1 = np.empty((n_edges, D))
    todo = np.arange(n_edges)
    length = (steps - 0.5) * resolution
    while todo.size:
        a = rng.uniform(lo, hi, size=(todo.size, D))
        d = rng.normal(size=(todo.size, D))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        b = a + length * d
        ok = np.all((b >= lo) & (b <= hi), axis=1)
        q0[todo[ok]] = a[ok]
        q1[todo[ok]] = b[ok]
        todo = todo[~ok]
    return q0.astype(np.float32), q1.astype(np.float32)


def random_edges(model: RobotModel, n_edges: int, seed: int = 0, max_len: float = 2.5) -> Tuple[np.ndarray, np.ndarray]:
    """Ragged edges (1 .. ~48 steps) for parity tests: both ends uniform in the limits, length capped."""
    rng = np.random.RandomState(seed)
    lo, hi = model.lower, model.upper
    a = rng.uniform(lo, hi, size=(n_edges, model.dof))
    b = rng.uniform(lo, hi, size=(n_edges, model.dof))
    d = b - a
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    scale = np.minimum(1.0, rng.uniform(0.0, max_len, size=(n_edges, 1)) / np.maximum(nrm, 1e-9))
    return a.astype(np.float32), (a + d * scale).astype(np.float32)
