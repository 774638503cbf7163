"""Pack robot / obstacle / plant-world models into the flat float32 SoA tables of the C-ABI
(include/pmp_edgecheck.h).  Layout constants are mirrored in csrc/pmp_tables.h.

Static obstacles of the reference's worlds (env.py:38-46; identical in every world because all
bodies of a world share its xy offset):
  floor  models/short_floor.urdf: box 2 x 2 x 0.1 centred z = -0.05
  block  models/drake/objects/block_for_pick_and_place_mid_size.urdf: box 0.059 x 0.089 x 0.199
         posed at (0.4, -0.4, 0.45), yaw 1.57, plus its eight 1e-7 m contact spheres at the corners of the visual box
         (0.06 x 0.075 x 0.2: they stick 0.5 mm out of the collision box in x and z)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .plants import OBS_ANGLE, OBS_MOD3, PlantWorld, Stem
from .robot import (MAX_FIXED_PRIMS, PRIM_BOX, PRIM_SPHERE, FixedPrim, RobotModel, joint_constant_matrices)
from .transforms import Frame, rpy_to_matrix

STEM_STRIDE = 48          # floats per stem record
BOX_STRIDE = 16           # floats per world box record
MAX_STEMS = 32            # per world
MAX_WORLD_BOXES = 32
MAX_SLOTS = 4          # parent-frame slots compiled into the kernels (PMP_MAX_SLOTS)

# stem record field offsets (floats); mirrored in csrc/pmp_tables.h
S_RS0, S_B0, S_HW, S_HH, S_ZC, S_MARGIN = 0, 9, 12, 13, 14, 15
S_OBSV, S_OBSP, S_COMZ, S_RAW0 = 16, 19, 22, 23
S_PSLOT, S_OSLOT, S_FLAGS, S_LIM = 24, 25, 26, 27
S_TH0, S_BRAD = 28, 30
S_RO, S_TO, S_TV0 = 32, 41, 44
F_PLANT_START, F_IS_MAIN, F_OBS_ANGLE, F_OBS_ROOT, F_SNAP, F_HAS_PARENT = 1, 2, 4, 8, 16, 32
STEM_MARGIN = 1.0e-3      # Bullet's collision margin of the stem boxes (m_defaultCollisionMargin)

MODE_OUR_METHOD, MODE_AVOID_ALL, MODE_IGNORE_ALL = 0, 1, 2
MODES = {"our_method": MODE_OUR_METHOD, "avoid_all": MODE_AVOID_ALL, "ignore_all": MODE_IGNORE_ALL}


def default_obstacles() -> List[FixedPrim]:
    floor = FixedPrim(PRIM_BOX, np.eye(3), np.array([0.0, 0.0, -0.05]), np.array([1.0, 1.0, 0.05]),
                      0xFFFFFFFF, "floor")
    Rb, tb = rpy_to_matrix((0.0, 0.0, 1.57)), np.array([0.4, -0.4, 0.45])
    block = FixedPrim(PRIM_BOX, Rb, tb, 0.5 * np.array([0.059, 0.089, 0.199]), 0xFFFFFFFF, "block")
    corners = [FixedPrim(PRIM_SPHERE, Rb, Rb @ np.array([sx * 0.03, sy * 0.0375, sz * 0.1]) + tb,
                         np.array([1e-7, 0.0, 0.0]), 0xFFFFFFFF, "block_corner")
               for sz in (-1, 1) for sx in (-1, 1) for sy in (-1, 1)]
    return [floor, block] + corners


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


from .statics import _dot, _mm, _mtv, _mv  # noqa: E402  (fixed-order 3x3 products)


def _rx_ry(tx: np.ndarray, ty: np.ndarray) -> np.ndarray:
    cx, sx, cy, sy = np.cos(tx), np.sin(tx), np.cos(ty), np.sin(ty)
    M = np.zeros(tx.shape + (3, 3))
    M[..., 0, 0], M[..., 0, 2] = cy, sy
    M[..., 1, 0], M[..., 1, 1], M[..., 1, 2] = sx * sy, cx, -sx * cy
    M[..., 2, 0], M[..., 2, 1], M[..., 2, 2] = -cx * sy, sx, cx * cy
    return M


def _observer_raw(flags: int, c: np.ndarray, refp: np.ndarray, com_z: np.ndarray, Rs: np.ndarray, tv: np.ndarray,
                  Rp: Optional[np.ndarray]) -> np.ndarray:
    """The kernels' observer arithmetic in float64, vectorised over worlds (csrc/pmp_kernels.cuh observer block;
    representation.py:27-51, 84-123, 220-235).  c / refp [W,3], Rs / Rp [W,3,3], tv [W,3]."""
    if flags & F_OBS_ANGLE:
        v = com_z[:, None] * Rs[:, :, 2] + tv - refp
        if flags & F_SNAP:
            v = v.copy()
            v[:, 0][np.abs(v[:, 0]) < 1e-5] = 0.0
            v[:, 1][np.abs(v[:, 1]) < 1e-5] = 0.0
        k = np.cross(v, c)
        return np.arctan2(np.sqrt(_dot(k, k)), _dot(v, c))
    if flags & F_OBS_ROOT:
        v = _mv(Rs, c)
    else:
        v = _mv(Rp, _mtv(Rs, c))
    sarg = -v[:, 0]
    gim = np.abs(sarg) >= 0.99999
    roll = np.where(gim, 0.0, np.arctan2(v[:, 1], v[:, 2]))
    pitch = np.where(gim, np.sign(sarg) * 0.5 * np.pi, np.arcsin(np.clip(sarg, -1.0, 1.0)))
    return np.sqrt(roll * roll + pitch * pitch)


def build_stem_records(parent: Sequence[int], origin_R: np.ndarray, origin_t: np.ndarray, half_height: np.ndarray,
                       z_offset: np.ndarray, half_width: np.ndarray, com_z: np.ndarray, limit: np.ndarray,
                       flags: Sequence[int], obs_ref_point: np.ndarray, obs_vertical: Sequence[bool],
                       pslot: Sequence[int], oslot: Sequence[int], margin: float = STEM_MARGIN,
                       gravity_sag: bool = True, rest_theta: Optional[np.ndarray] = None) -> np.ndarray:
    """float32 [W, P, STEM_STRIDE] stem records of W worlds sharing one topology (`parent`, `flags`, slots).
    Per-world arrays: origin_R [W,P,3,3], origin_t [W,P,3], the rest [W,P] (obs_ref_point [W,P,3]).

    The record holds the stem's REST pose (springs and gravity in balance, model/statics.py), in which the kernels
    test it first, the box the reference gives it (rand_gen_plants_programmatically_main.py:30-51) with Bullet's
    margin carved out, the observer constants (captured at zero deflection, like the reference's observers) and the
    observer value of the rest pose."""
    from .statics import rest_pose_batch
    W, P = half_height.shape
    th0 = rest_theta
    if th0 is None:
        th0 = (rest_pose_batch(list(parent), origin_R, origin_t, com_z, limit=float(np.min(limit)) if P else 1.5)
               if gravity_sag else np.zeros((W, P, 2)))
    rec = np.zeros((W, P, STEM_STRIDE), dtype=np.float32)
    irec = rec.view(np.int32)
    ez = np.array([0.0, 0.0, 1.0])
    R0 = [None] * P; T0 = [None] * P      # zero-deflection link frames (observer references)
    Rr = [None] * P; Tr = [None] * P      # rest link frames / joint origins
    for s in range(P):
        p = int(parent[s])
        if p < 0:
            R0[s], T0[s] = origin_R[:, s], origin_t[:, s]
            Rv, Tr[s] = origin_R[:, s], origin_t[:, s]
        else:
            R0[s] = _mm(R0[p], origin_R[:, s])
            T0[s] = _mv(R0[p], origin_t[:, s]) + T0[p]
            Rv = _mm(Rr[p], origin_R[:, s])
            Tr[s] = _mv(Rr[p], origin_t[:, s]) + Tr[p]
        Rr[s] = _mm(Rv, _rx_ry(th0[:, s, 0], th0[:, s, 1]))
        zc = z_offset[:, s] + half_height[:, s]
        rec[:, s, S_RS0:S_RS0 + 9] = Rr[s].reshape(W, 9)
        rec[:, s, S_B0:S_B0 + 3] = _mtv(Rr[s], Tr[s]) + zc[:, None] * ez[None]
        hw = np.maximum(half_width[:, s] - margin, 0.0)
        hh = np.maximum(half_height[:, s] - margin, 0.0)
        rec[:, s, S_HW], rec[:, s, S_HH], rec[:, s, S_ZC], rec[:, s, S_MARGIN] = hw, hh, zc, margin
        rec[:, s, S_COMZ], rec[:, s, S_LIM] = com_z[:, s], limit[:, s]
        rec[:, s, S_TH0], rec[:, s, S_TH0 + 1] = th0[:, s, 0], th0[:, s, 1]
        # radius of the box's cross-section about its axis, rounded up: the swept-bound cull may use the capsule
        rec[:, s, S_BRAD] = (np.sqrt(2.0) * hw + margin) * 1.000001
        rec[:, s, S_RO:S_RO + 9] = origin_R[:, s].reshape(W, 9)
        rec[:, s, S_TO:S_TO + 3] = origin_t[:, s]
        rec[:, s, S_TV0:S_TV0 + 3] = Tr[s]
        f = int(flags[s])
        if f & F_OBS_ANGLE:
            com0 = R0[s][:, :, 2] * com_z[:, s, None] + T0[s]
            d = np.broadcast_to(ez, (W, 3)).copy() if obs_vertical[s] else com0 - obs_ref_point[:, s]
            nrm = np.linalg.norm(d, axis=1, keepdims=True)
            c = np.where(nrm > 0, d / np.where(nrm > 0, nrm, 1.0), ez[None])
            Rp_rest = None
        elif f & F_OBS_ROOT:
            c = R0[s][:, 2, :]                                   # R_l0^T ez = third row of R_l0
            Rp_rest = None
        else:
            c = _mv(R0[s], R0[p][:, 2, :])                       # R_l0 R_p0^T ez
            Rp_rest = Rr[p]
        rec[:, s, S_OBSV:S_OBSV + 3] = c
        rec[:, s, S_OBSP:S_OBSP + 3] = obs_ref_point[:, s]
        rec[:, s, S_RAW0] = _observer_raw(f, c, obs_ref_point[:, s], com_z[:, s], Rr[s], Tr[s], Rp_rest)
        irec[:, s, S_PSLOT] = int(pslot[s])
        irec[:, s, S_OSLOT] = int(oslot[s])
        irec[:, s, S_FLAGS] = f
    return rec


def stem_flags(s: Stem) -> int:
    f = 0
    if s.plant_start:
        f |= F_PLANT_START
    if s.is_main:
        f |= F_IS_MAIN
    if s.parent >= 0:
        f |= F_HAS_PARENT
    if s.obs_kind == OBS_MOD3:
        if s.parent < 0:
            f |= F_OBS_ROOT
    elif s.obs_kind == OBS_ANGLE:
        f |= F_OBS_ANGLE
        if s.obs_snap:
            f |= F_SNAP
    else:
        raise ValueError(f"unknown observer kind {s.obs_kind}")
    return f


def pack_stem_records(world: PlantWorld, margin: float = STEM_MARGIN, gravity_sag: bool = True) -> Tuple[np.ndarray, int]:
    """float32 [n_stems, STEM_STRIDE] records + slot count for one world."""
    n = world.num_stems
    if n > MAX_STEMS:
        raise ValueError(f"{n} stems > {MAX_STEMS}")
    par, own, n_slots = assign_slots(world)
    if n_slots > MAX_SLOTS:
        raise ValueError(f"plant needs {n_slots} frame slots > {MAX_SLOTS}")
    if n == 0:
        return np.zeros((0, STEM_STRIDE), dtype=np.float32), n_slots
    st = world.stems
    arr = lambda f: np.asarray([f(s) for s in st], dtype=np.float64)[None]
    rec = build_stem_records(
        [s.parent for s in st], np.stack([s.origin.R for s in st])[None], np.stack([s.origin.t for s in st])[None],
        arr(lambda s: s.half_height), arr(lambda s: s.z_offset), arr(lambda s: s.half_width), arr(lambda s: s.com_z),
        arr(lambda s: s.limit), [stem_flags(s) for s in st], np.stack([s.obs_ref_point for s in st])[None],
        [s.obs_vertical for s in st], par, own, margin, gravity_sag)
    return rec[0], n_slots


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


def pack_worlds(worlds: Sequence[PlantWorld], stem_margin: float = STEM_MARGIN, gravity_sag: bool = True) -> WorldTables:
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
    # worlds of one topology (parents, flags, observer wiring) are packed in one vectorised call
    groups: dict = {}
    for w, world in enumerate(worlds):
        if world.num_stems > MAX_STEMS:
            raise ValueError(f"{world.num_stems} stems > {MAX_STEMS}")
        key = (tuple(s.parent for s in world.stems), tuple(stem_flags(s) for s in world.stems),
               tuple(bool(s.obs_vertical) for s in world.stems))
        groups.setdefault(key, []).append(w)
    for (parents, flags, vertical), idx in groups.items():
        n = len(parents)
        if n == 0:
            continue
        par, own, ns = assign_slots(worlds[idx[0]])
        if ns > MAX_SLOTS:
            raise ValueError(f"plant needs {ns} frame slots > {MAX_SLOTS}")
        n_slots = max(n_slots, ns)
        arr = lambda f: np.asarray([[f(s) for s in worlds[w].stems] for w in idx], dtype=np.float64)
        rec = build_stem_records(
            list(parents), arr(lambda s: s.origin.R), arr(lambda s: s.origin.t), arr(lambda s: s.half_height),
            arr(lambda s: s.z_offset), arr(lambda s: s.half_width), arr(lambda s: s.com_z), arr(lambda s: s.limit),
            list(flags), arr(lambda s: s.obs_ref_point), list(vertical), par, own, stem_margin, gravity_sag)
        stems[idx, :n] = rec
        n_stems[idx] = n
    for w, world in enumerate(worlds):
        n_boxes[w] = len(world.base_boxes)
        for b, (fr, half) in enumerate(world.base_boxes):
            boxes[w, b, 0:9] = fr.R.reshape(-1)
            boxes[w, b, 9:12] = fr.t
            boxes[w, b, 12:15] = half
    return WorldTables(W, P, NB, n_slots, n_stems, stems, n_boxes, boxes)
