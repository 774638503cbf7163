"""Pack robot / obstacle / plant-world models into the flat float32 SoA tables of the C-ABI
(include/pmp_edgecheck.h).  Layout constants are mirrored in csrc/pmp_tables.h.

Static obstacles of the reference's worlds (env.py:38-46; identical in every world because all
bodies of a world share its xy offset):
  floor  models/short_floor.urdf: box 2 x 2 x 0.1 centred z = -0.05
  block  models/drake/objects/block_for_pick_and_place_mid_size.urdf: box 0.059 x 0.089 x 0.199
         posed at (0.4, -0.4, 0.45), yaw 1.57 (its eight 1e-7 m corner spheres are dropped)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .plants import OBS_ANGLE, OBS_MOD3, PlantWorld, Stem
from .robot import (MAX_FIXED_PRIMS, PRIM_BOX, FixedPrim, RobotModel, joint_constant_matrices)
from .transforms import Frame, rpy_to_matrix

STEM_STRIDE = 48          # floats per stem record
BOX_STRIDE = 16           # floats per world box record
MAX_STEMS = 32            # per world
MAX_WORLD_BOXES = 32
MAX_SLOTS = 8

# stem record field offsets (floats)
S_RO, S_TO, S_Z0, S_Z1, S_RAD, S_LIM, S_OBSV, S_OBSP, S_COMZ = 0, 9, 12, 13, 14, 15, 16, 19, 22
S_PSLOT, S_OSLOT, S_FLAGS = 24, 25, 26
S_RW0, S_TW0 = 32, 41
F_PLANT_START, F_IS_MAIN, F_OBS_ANGLE, F_OBS_ROOT, F_SNAP, F_HAS_PARENT = 1, 2, 4, 8, 16, 32

MODE_OUR_METHOD, MODE_AVOID_ALL, MODE_IGNORE_ALL = 0, 1, 2
MODES = {"our_method": MODE_OUR_METHOD, "avoid_all": MODE_AVOID_ALL, "ignore_all": MODE_IGNORE_ALL}


def default_obstacles() -> List[FixedPrim]:
    floor = FixedPrim(PRIM_BOX, np.eye(3), np.array([0.0, 0.0, -0.05]), np.array([1.0, 1.0, 0.05]),
                      0xFFFFFFFF, "floor")
    block = FixedPrim(PRIM_BOX, rpy_to_matrix((0.0, 0.0, 1.57)), np.array([0.4, -0.4, 0.45]),
                      0.5 * np.array([0.059, 0.089, 0.199]), 0xFFFFFFFF, "block")
    return [floor, block]


@dataclass
class RobotTables:
    dof: int
    n_spheres: int
    n_links: int
    n_fixed: int
    joint_C: np.ndarray
    joint_A: np.ndarray
    joint_B: np.ndarray
    joint_t: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    circular: np.ndarray
    base_R: np.ndarray
    base_t: np.ndarray
    sph_joint: np.ndarray
    sph_center: np.ndarray
    sph_radius: np.ndarray
    self_mask: np.ndarray
    link_joint: np.ndarray
    link_R: np.ndarray
    link_t: np.ndarray
    link_com: np.ndarray
    ee_joint: int
    ee_local: np.ndarray
    fixed_kind: np.ndarray
    fixed_R: np.ndarray
    fixed_t: np.ndarray
    fixed_half: np.ndarray
    fixed_mask: np.ndarray


def pack_robot(model: RobotModel, obstacles: Optional[Sequence[FixedPrim]] = None) -> RobotTables:
    """Robot + static obstacle tables.  ``obstacles`` default = the reference's floor and block."""
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32))
    Cm, Am, Bm = joint_constant_matrices(model)
    A = model.num_spheres
    self_mask = np.zeros(A, dtype=np.uint32)
    for a in range(A):
        m = 0
        for b in range(a + 1, A):
            if model.self_pair_mask[a, b]:
                m |= (1 << b)
        self_mask[a] = m
    obstacles = default_obstacles() if obstacles is None else list(obstacles)
    fixed = list(obstacles) + list(model.base_prims)
    # obstacles are tested against every arm sphere (product(moving links, fixed bodies),
    # pybullet_tools/utils.py:4008-4014); the robot's own welded geometry only against the
    # self-collision link pairs.
    all_mask = (1 << A) - 1
    masks = [all_mask & 0xFFFFFFFF] * len(obstacles) + [p.sphere_mask for p in model.base_prims]
    if len(fixed) > MAX_FIXED_PRIMS:
        raise ValueError(f"{len(fixed)} static primitives > {MAX_FIXED_PRIMS}")
    F = len(fixed)
    return RobotTables(
        dof=model.dof, n_spheres=A, n_links=model.num_links, n_fixed=F,
        joint_C=f32(Cm).reshape(-1), joint_A=f32(Am).reshape(-1), joint_B=f32(Bm).reshape(-1),
        joint_t=f32(model.joint_t0).reshape(-1), lower=f32(model.lower), upper=f32(model.upper),
        circular=np.ascontiguousarray(model.circular.astype(np.uint8)),
        base_R=f32(model.base.R).reshape(-1), base_t=f32(model.base.t),
        sph_joint=np.ascontiguousarray(model.sph_joint.astype(np.int32)),
        sph_center=f32(model.sph_center).reshape(-1), sph_radius=f32(model.sph_radius),
        self_mask=self_mask,
        link_joint=np.ascontiguousarray(model.link_joint.astype(np.int32)),
        link_R=f32(model.link_R).reshape(-1), link_t=f32(model.link_t).reshape(-1),
        link_com=f32(model.link_com).reshape(-1),
        ee_joint=int(model.ee_joint), ee_local=f32(model.ee_local),
        fixed_kind=np.asarray([p.kind for p in fixed], dtype=np.int32),
        fixed_R=f32(np.stack([p.R for p in fixed]) if F else np.zeros((0, 3, 3))).reshape(-1),
        fixed_t=f32(np.stack([p.t for p in fixed]) if F else np.zeros((0, 3))).reshape(-1),
        fixed_half=f32(np.stack([p.half for p in fixed]) if F else np.zeros((0, 3))).reshape(-1),
        fixed_mask=np.asarray(masks, dtype=np.uint32),
    )


# --------------------------------------------------------------------------- plant worlds
def stem_rest_frames(world: PlantWorld) -> List[Frame]:
    """World frames of the stem links at zero deflection (float64)."""
    frames: List[Frame] = []
    for s in world.stems:
        parent = Frame() if s.parent < 0 else frames[s.parent]
        frames.append(parent @ s.origin)
    return frames


def assign_slots(world: PlantWorld) -> Tuple[List[int], List[int], int]:
    """Register-slot allocation for parent frames.

    The kernels walk stems in link order and keep the frames of stems that still have unvisited
    children in a small register file.  Returns (parent_slot[s], own_slot[s], n_slots);
    own_slot = -1 for leaves, parent_slot = -1 for roots.
    """
    n = world.num_stems
    last_child = [-1] * n
    for i, s in enumerate(world.stems):
        if s.parent >= i:
            raise ValueError("stems must be listed parents-first")
        if s.parent >= 0:
            last_child[s.parent] = i
    own = [-1] * n
    free: List[int] = []
    n_slots = 0
    busy_until: dict = {}
    for i in range(n):
        for sl, until in list(busy_until.items()):
            if until < i:          # parent no longer needed by stems >= i
                free.append(sl)
                del busy_until[sl]
        if last_child[i] >= 0:
            if free:
                sl = min(free)
                free.remove(sl)
            else:
                sl = n_slots
                n_slots += 1
            own[i] = sl
            busy_until[sl] = last_child[i]
    par = [(-1 if s.parent < 0 else own[s.parent]) for s in world.stems]
    return par, own, max(n_slots, 1)


def stem_capsule(s: Stem, stem_radius: Optional[float]) -> Tuple[float, float, float]:
    """(z0, z1, r): capsule standing in for the stem's square-section box (rand_gen...:34-37).

    Default radius gives the capsule the box's cross-section area (r = w / sqrt(pi), 28.2 mm for
    the 50 mm stems); the segment is shortened by r at both ends so the capsule's length equals
    the box's.
    """
    r = float(stem_radius) if stem_radius is not None else 2.0 * s.half_width / np.sqrt(np.pi)
    z0 = s.z_offset + r
    z1 = s.z_offset + 2.0 * s.half_height - r
    if z1 < z0:
        z0 = z1 = s.z_offset + s.half_height
    return z0, z1, r


def pack_stem_records(world: PlantWorld, stem_radius: Optional[float] = None) -> Tuple[np.ndarray, int]:
    """float32 [n_stems, STEM_STRIDE] records + slot count for one world."""
    n = world.num_stems
    if n > MAX_STEMS:
        raise ValueError(f"{n} stems > {MAX_STEMS}")
    par, own, n_slots = assign_slots(world)
    if n_slots > MAX_SLOTS:
        raise ValueError(f"plant needs {n_slots} frame slots > {MAX_SLOTS}")
    rest = stem_rest_frames(world)
    rec = np.zeros((n, STEM_STRIDE), dtype=np.float32)
    irec = rec.view(np.int32)
    ez = np.array([0.0, 0.0, 1.0])
    for i, s in enumerate(world.stems):
        z0, z1, r = stem_capsule(s, stem_radius)
        rec[i, S_RO:S_RO + 9] = s.origin.R.reshape(-1)
        rec[i, S_TO:S_TO + 3] = s.origin.t
        rec[i, S_Z0], rec[i, S_Z1], rec[i, S_RAD], rec[i, S_LIM] = z0, z1, r, s.limit
        rec[i, S_COMZ] = s.com_z
        rec[i, S_RW0:S_RW0 + 9] = rest[i].R.reshape(-1)
        rec[i, S_TW0:S_TW0 + 3] = rest[i].t
        flags = 0
        if s.plant_start:
            flags |= F_PLANT_START
        if s.is_main:
            flags |= F_IS_MAIN
        if s.parent >= 0:
            flags |= F_HAS_PARENT
        Rl0 = rest[i].R
        if s.obs_kind == OBS_MOD3:
            if s.parent < 0:
                flags |= F_OBS_ROOT
                c = Rl0.T @ ez                       # v = R_l c
            else:
                Rp0 = rest[s.parent].R
                c = Rl0 @ Rp0.T @ ez                 # v = R_p (R_l^T c)
            rec[i, S_OBSV:S_OBSV + 3] = c
        elif s.obs_kind == OBS_ANGLE:
            flags |= F_OBS_ANGLE
            if s.obs_snap:
                flags |= F_SNAP
            com0 = rest[i].apply(np.array([0.0, 0.0, s.com_z]))
            d = ez if s.obs_vertical else (com0 - s.obs_ref_point)
            nrm = np.linalg.norm(d)
            rec[i, S_OBSV:S_OBSV + 3] = d / nrm if nrm > 0 else ez
            rec[i, S_OBSP:S_OBSP + 3] = s.obs_ref_point
        else:
            raise ValueError(f"unknown observer kind {s.obs_kind}")
        irec[i, S_PSLOT] = par[i]
        irec[i, S_OSLOT] = own[i]
        irec[i, S_FLAGS] = flags
    return rec, n_slots


@dataclass
class WorldTables:
    n_worlds: int
    max_stems: int
    max_boxes: int
    n_slots: int
    n_stems: np.ndarray     # int32 [W]
    stems: np.ndarray       # float32 [W, max_stems, STEM_STRIDE]
    n_boxes: np.ndarray     # int32 [W]
    boxes: np.ndarray       # float32 [W, max_boxes, BOX_STRIDE]   (R 9, t 3, half 3, pad)


def pack_worlds(worlds: Sequence[PlantWorld], stem_radius: Optional[float] = None) -> WorldTables:
    W = len(worlds)
    if W < 1:
        raise ValueError("need at least one world")
    P = max(1, max(w.num_stems for w in worlds))
    NB = max(1, max(len(w.base_boxes) for w in worlds))
    if NB > MAX_WORLD_BOXES:
        raise ValueError(f"{NB} plant-base boxes in a world > {MAX_WORLD_BOXES}")
    stems = np.zeros((W, P, STEM_STRIDE), dtype=np.float32)
    boxes = np.zeros((W, NB, BOX_STRIDE), dtype=np.float32)
    n_stems = np.zeros(W, dtype=np.int32)
    n_boxes = np.zeros(W, dtype=np.int32)
    n_slots = 1
    for w, world in enumerate(worlds):
        rec, ns = pack_stem_records(world, stem_radius)
        n_slots = max(n_slots, ns)
        stems[w, :rec.shape[0]] = rec
        n_stems[w] = rec.shape[0]
        n_boxes[w] = len(world.base_boxes)
        for b, (fr, half) in enumerate(world.base_boxes):
            boxes[w, b, 0:9] = fr.R.reshape(-1)
            boxes[w, b, 9:12] = fr.t
            boxes[w, b, 12:15] = half
    return WorldTables(W, P, NB, n_slots, n_stems, stems, n_boxes, boxes)
