"""Vectorised multi-branch plant sampler for large world counts.

``sample_world_tables(n, ...)`` draws n plants of one topology at once and writes the kernel's world tables directly,
without building n PlantWorld objects (the sequential sampler of model/plants.py -- the bit-for-bit restatement of
create_plant_params, rand_gen_plants_programmatically_main.py:71-277 -- manages about 5 k worlds/s; this one
about 120 k worlds/s).  Every quantity has the distribution the reference draws it from, the branch-placement rejection
rule included; what is NOT reproduced is the order of draws in the two global streams, so a seed here does not
correspond to a seed of the scripts (use model/plants.sample_world for that).

``worlds_from_draws`` turns the same draws into PlantWorld objects; tests pack those with the scalar packer and require
the vectorised tables to be identical.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .plants import OBS_MOD3, PLANT_JOINT_LIMIT, SCALING, STEM_COM_Z, STEM_HALF_WIDTH, PlantWorld, Stem
from .tables import (BOX_STRIDE, F_HAS_PARENT, F_IS_MAIN, F_OBS_ROOT, F_PLANT_START, MAX_SLOTS, MAX_STEMS, STEM_MARGIN,
                     WorldTables, assign_slots, build_stem_records)
from .transforms import Frame


@dataclass
class PlantDraws:
    """The random numbers of n plants of one topology (everything the tables are a function of)."""
    parent: np.ndarray        # int [P] parent stem, -1 for the first
    is_main: np.ndarray       # bool [P]
    base_xy: np.ndarray       # [n, 2]
    base_half: np.ndarray     # [n, 3] base box half extents
    half_height: np.ndarray   # [n, P]
    xyz: np.ndarray           # [n, P, 3] joint origin in the parent's link frame (base frame for the first stem)
    rpy: np.ndarray           # [n, P, 3]
    spacing: float


def _topology(branches: int, verts: int, exts: int) -> Tuple[np.ndarray, np.ndarray, List[Tuple[str, int]]]:
    """Creation order of create_plant_params: per vertical stem, the main stem then `branches` x (branch + `exts`
    extensions).  Returns (parent [P], is_main [P], kind list) with kind in {'first', 'stack', 'branch', 'ext'}."""
    parent, is_main, kinds = [], [], []
    main = -1
    for v in range(verts):
        k = len(parent)
        parent.append(main)
        is_main.append(True)
        kinds.append(("first" if v == 0 else "stack", main))
        main = k
        for _ in range(branches):
            parent.append(main)
            is_main.append(False)
            kinds.append(("branch", main))
            for _ in range(exts):
                parent.append(len(parent) - 1)
                is_main.append(False)
                kinds.append(("ext", len(parent) - 2))
    return np.asarray(parent), np.asarray(is_main, dtype=bool), kinds


def draw_plants(n: int, branches: int, verts: int, exts: int, plant_pos_xy_limits, seed: int = 0,
                stem_base_spacing: float = 0.0) -> PlantDraws:
    if exts < 1 or branches < 0 or verts < 1:
        raise ValueError("need total_num_extensions >= 1 and total_num_vert_stems >= 1")
    rng = np.random.default_rng(seed)
    sf = SCALING
    parent, is_main, kinds = _topology(branches, verts, exts)
    P = len(parent)
    if P > MAX_STEMS:
        raise ValueError(f"{P} stems > {MAX_STEMS}")
    (x0, x1), (y0, y1) = plant_pos_xy_limits
    base_xy = np.stack([rng.uniform(x0, x1, n), rng.uniform(y0, y1, n)], axis=1)
    base_half = rng.uniform(sf * 0.07, sf * 0.15, size=(n, 3))
    hh = rng.uniform(sf * 0.7, sf * 1.0, size=(n, P))                       # create_stem_element (rand_gen...:32)
    xyz = np.zeros((n, P, 3))
    rpy = np.zeros((n, P, 3))
    taken: List[np.ndarray] = []                                            # branch (x, y) already placed on this main stem
    for k, (kind, ref) in enumerate(kinds):
        if kind == "first":
            xyz[:, k, 2] = base_half[:, 2]
            taken = []
        elif kind == "stack":                                               # rand_gen...:237-244
            taken = []
            xyz[:, k, 2] = stem_base_spacing + 2.0 * hh[:, ref]
            ang = rng.uniform(-0.4, 0.4, n)
            pitch_side = rng.random(n) < 0.5
            rpy[:, k, 1] = np.where(pitch_side, ang, 0.0)
            rpy[:, k, 0] = np.where(pitch_side, 0.0, ang)
            rpy[:, k, 2] = rng.uniform(-0.4, 0.4, n)
        elif kind == "ext":                                                 # rand_gen...:183-188
            xyz[:, k, 2] = stem_base_spacing + 2.0 * hh[:, ref]
            ang = rng.uniform(sf * -0.5, sf * 0.5, n)
            roll_side = rng.random(n) < 0.5
            rpy[:, k, 0] = np.where(roll_side, ang, 0.0)
            rpy[:, k, 1] = np.where(roll_side, 0.0, ang)
        else:                                                               # branch, rand_gen...:158-162, 196-213
            x = np.empty(n)
            y = np.empty(n)
            todo = np.ones(n, dtype=bool)
            tol = sf * 0.2
            while todo.any():
                m = int(todo.sum())
                side = rng.random(m) < 0.5
                fixed = np.where(rng.random(m) < 0.5, sf * 0.2, sf * -0.2)
                free = rng.uniform(0.0, sf * 0.2, m)
                x[todo] = np.where(side, fixed, free)
                y[todo] = np.where(side, free, fixed)
                clash = np.zeros(n, dtype=bool)
                for (px, py) in taken:
                    clash |= (np.abs(x - px) < tol) & (np.abs(y - py) < tol)
                todo &= clash
            taken.append((x.copy(), y.copy()))
            yaw = rng.uniform(-0.5, 0.5, n)
            mag = rng.uniform(0.7, 1.5, n)
            along_x = np.abs(x) > np.abs(y)
            rpy[:, k, 1] = np.where(along_x, np.where(x > 0, mag, -mag), 0.0)
            rpy[:, k, 0] = np.where(along_x, 0.0, np.where(y > 0, -mag, mag))
            rpy[:, k, 2] = yaw
            xyz[:, k, 0], xyz[:, k, 1] = x, y
            xyz[:, k, 2] = stem_base_spacing + rng.uniform(0.0, 1.0, n) * 2.0 * hh[:, ref]
    return PlantDraws(parent, is_main, base_xy, base_half, hh, xyz, rpy, float(stem_base_spacing))


# ---- the device sampler's draw layout (csrc/pmp_sampler.cuh; include/pmp_edgecheck.h PMP_SAMPLER_*) -----------------
SAMPLER_MAX_STEMS, SAMPLER_TRIES, SAMPLER_HEAD = 8, 16, 5
SAMPLER_PER_STEM = 8 + 3 * SAMPLER_TRIES
KIND_CODE = {"first": 0, "stack": 1, "branch": 2, "ext": 3}


def uniforms_per_world(n_stems: int) -> int:
    return SAMPLER_HEAD + n_stems * SAMPLER_PER_STEM


def splitmix_uniforms(seed: int, n: int, n_uniforms: int) -> np.ndarray:
    """The counter-based generator of the device sampler (splitmix64 of (seed, world * n_uniforms + draw)) -> [n, n_uniforms]
    doubles in [0, 1)."""
    with np.errstate(over="ignore"):
        ctr = np.arange(n * n_uniforms, dtype=np.uint64) + np.uint64(1)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + ctr * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)).reshape(n, n_uniforms)


def draws_from_uniforms(U: np.ndarray, branches: int, verts: int, exts: int, plant_pos_xy_limits,
                        stem_base_spacing: float = 0.0) -> PlantDraws:
    """Host mirror of the device sampler: the same fixed block of uniforms per world -> the same plants
    (distributions and placement rule of draw_plants; at most SAMPLER_TRIES placement attempts per branch)."""
    sf = SCALING
    parent, is_main, kinds = _topology(branches, verts, exts)
    P = len(parent)
    n = U.shape[0]
    if U.shape[1] < uniforms_per_world(P):
        raise ValueError("too few uniforms per world")
    lerp = lambda a, b, x: a + (b - a) * x
    (x0, x1), (y0, y1) = plant_pos_xy_limits
    base_xy = np.stack([lerp(x0, x1, U[:, 0]), lerp(y0, y1, U[:, 1])], axis=1)
    base_half = lerp(sf * 0.07, sf * 0.15, U[:, 2:5])
    hh = np.zeros((n, P)); xyz = np.zeros((n, P, 3)); rpy = np.zeros((n, P, 3))
    for k, (kind, ref) in enumerate(kinds):
        B = SAMPLER_HEAD + k * SAMPLER_PER_STEM
        hh[:, k] = lerp(sf * 0.7, sf * 1.0, U[:, B])
        if kind == "first":
            xyz[:, k, 2] = base_half[:, 2]
        elif kind == "stack":
            xyz[:, k, 2] = stem_base_spacing + 2.0 * hh[:, ref]
            ang = lerp(-0.4, 0.4, U[:, B + 1])
            pitch_side = U[:, B + 2] < 0.5
            rpy[:, k, 1] = np.where(pitch_side, ang, 0.0)
            rpy[:, k, 0] = np.where(pitch_side, 0.0, ang)
            rpy[:, k, 2] = lerp(-0.4, 0.4, U[:, B + 3])
        elif kind == "ext":
            xyz[:, k, 2] = stem_base_spacing + 2.0 * hh[:, ref]
            ang = lerp(sf * -0.5, sf * 0.5, U[:, B + 1])
            roll_side = U[:, B + 2] < 0.5
            rpy[:, k, 0] = np.where(roll_side, ang, 0.0)
            rpy[:, k, 1] = np.where(roll_side, 0.0, ang)
        else:
            tol = sf * 0.2
            x = np.zeros(n); y = np.zeros(n)
            todo = np.ones(n, dtype=bool)
            earlier = [j for j in range(k) if kinds[j][0] == "branch" and kinds[j][1] == ref]
            for t in range(SAMPLER_TRIES):
                side = U[:, B + 8 + 3 * t] < 0.5
                fixed = np.where(U[:, B + 9 + 3 * t] < 0.5, sf * 0.2, sf * -0.2)
                free = lerp(0.0, sf * 0.2, U[:, B + 10 + 3 * t])
                x = np.where(todo, np.where(side, fixed, free), x)
                y = np.where(todo, np.where(side, free, fixed), y)
                clash = np.zeros(n, dtype=bool)
                for j in earlier:
                    clash |= (np.abs(x - xyz[:, j, 0]) < tol) & (np.abs(y - xyz[:, j, 1]) < tol)
                todo &= clash
            yaw = lerp(-0.5, 0.5, U[:, B + 1])
            mag = lerp(0.7, 1.5, U[:, B + 2])
            along_x = np.abs(x) > np.abs(y)
            rpy[:, k, 1] = np.where(along_x, np.where(x > 0, mag, -mag), 0.0)
            rpy[:, k, 0] = np.where(along_x, 0.0, np.where(y > 0, -mag, mag))
            rpy[:, k, 2] = yaw
            xyz[:, k, 0], xyz[:, k, 1] = x, y
            xyz[:, k, 2] = stem_base_spacing + U[:, B + 3] * 2.0 * hh[:, ref]
    return PlantDraws(parent, is_main, base_xy, base_half, hh, xyz, rpy, float(stem_base_spacing))


def sampler_topology(branches: int, verts: int, exts: int) -> dict:
    """Topology arrays of the device sampler's descriptor (pmp_sampler_desc): parent, kind, ref, flags, frame slots."""
    parent, is_main, kinds = _topology(branches, verts, exts)
    P = len(parent)
    if P > SAMPLER_MAX_STEMS:
        raise ValueError(f"{P} stems > {SAMPLER_MAX_STEMS}: the device sampler handles small topologies; use "
                         "sample_world_tables (host) for this one")
    template = PlantWorld(stems=[Stem(parent=int(p), origin=Frame(), half_height=0.1, z_offset=0.0) for p in parent],
                          base_boxes=[])
    par, own, n_slots = assign_slots(template)
    flags = [(F_PLANT_START if k == 0 else 0) | (F_IS_MAIN if is_main[k] else 0) |
             (F_HAS_PARENT if parent[k] >= 0 else F_OBS_ROOT) for k in range(P)]
    return {"n_stems": P, "parent": [int(p) for p in parent], "kind": [KIND_CODE[k] for k, _ in kinds],
            "ref": [max(int(r), 0) for _, r in kinds], "flags": flags, "pslot": par, "oslot": own, "n_slots": n_slots}


def _rpy_matrices(rpy: np.ndarray) -> np.ndarray:
    """Rz(yaw) Ry(pitch) Rx(roll) for arrays [..., 3] -> [..., 3, 3] (model/transforms.rpy_to_matrix)."""
    r, p, y = rpy[..., 0], rpy[..., 1], rpy[..., 2]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.empty(rpy.shape[:-1] + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def tables_from_draws(d: PlantDraws, stem_margin: float = STEM_MARGIN, gravity_sag: bool = True) -> WorldTables:
    """The kernel's world tables of all n plants (what model/tables.pack_worlds writes, vectorised over worlds: both go
    through tables.build_stem_records, so the records are identical)."""
    n, P = d.half_height.shape
    base_pos = np.concatenate([d.base_xy, d.base_half[:, 2:3]], axis=1)
    R = _rpy_matrices(d.rpy)                                               # [n, P, 3, 3]
    t = d.xyz.copy()
    t[:, 0] += base_pos                                                    # first stem's origin is a world frame
    # slot assignment, flags: functions of the topology only -- take them from a template object
    template = PlantWorld(stems=[Stem(parent=int(p), origin=Frame(), half_height=0.1, z_offset=0.0) for p in d.parent],
                          base_boxes=[])
    par, own, n_slots = assign_slots(template)
    if n_slots > MAX_SLOTS:
        raise ValueError(f"plant needs {n_slots} frame slots > {MAX_SLOTS}")
    flags = [(F_PLANT_START if k == 0 else 0) | (F_IS_MAIN if d.is_main[k] else 0) |
             (F_HAS_PARENT if d.parent[k] >= 0 else F_OBS_ROOT) for k in range(P)]
    full = lambda v: np.full((n, P), float(v))
    rec = build_stem_records([int(p) for p in d.parent], R, t, d.half_height, full(d.spacing), full(STEM_HALF_WIDTH),
                             full(STEM_COM_Z), full(PLANT_JOINT_LIMIT), flags, np.zeros((n, P, 3)), [False] * P,
                             par, own, stem_margin, gravity_sag)
    boxes = np.zeros((n, 1, BOX_STRIDE), dtype=np.float32)
    boxes[:, 0, 0:9] = np.eye(3).reshape(-1)
    boxes[:, 0, 9:12] = base_pos
    boxes[:, 0, 12:15] = d.base_half
    return WorldTables(n, P, 1, n_slots, np.full(n, P, dtype=np.int32), rec, np.ones(n, dtype=np.int32), boxes)


def worlds_from_draws(d: PlantDraws, index: Sequence[int]) -> List[PlantWorld]:
    """PlantWorld objects of selected plants (the scalar representation of the same draws)."""
    from .transforms import rpy_to_matrix
    out = []
    for i in index:
        base_pos = np.array([d.base_xy[i, 0], d.base_xy[i, 1], d.base_half[i, 2]])
        stems = []
        for k in range(len(d.parent)):
            t = d.xyz[i, k] + (base_pos if d.parent[k] < 0 else 0.0)
            stems.append(Stem(parent=int(d.parent[k]), origin=Frame(rpy_to_matrix(d.rpy[i, k]), t),
                              half_height=float(d.half_height[i, k]), z_offset=d.spacing, plant_start=(k == 0),
                              is_main=bool(d.is_main[k]), obs_kind=OBS_MOD3))
        out.append(PlantWorld(stems=stems, base_boxes=[(Frame(np.eye(3), base_pos), d.base_half[i].copy())]))
    return out


def sample_world_tables(n: int, branches: int, verts: int, exts: int, plant_pos_xy_limits, seed: int = 0,
                        stem_base_spacing: float = 0.0, stem_margin: float = STEM_MARGIN,
                        gravity_sag: bool = True) -> WorldTables:
    return tables_from_draws(draw_plants(n, branches, verts, exts, plant_pos_xy_limits, seed, stem_base_spacing),
                             stem_margin, gravity_sag)
