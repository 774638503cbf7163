"""Plant worlds: the reference's random multi-branch sampler, its on-disk formats, and the
single-stem plant sets -- restated as kinematic *stem tables* (no physics engine).

Reference behaviour followed here:
  * sampler and RNG draw order ....... rand_gen_plants_programmatically_main.py:30-51, 71-277, 280-309
  * plant xy draw per world ........... env.py:63-64
  * saved plants (URDF + pickle) ...... plant_motion_planning/utils.py:270-279, 287-325
  * single-stem plant sets ............ plant_motion_planning/utils.py:97-117, 120-230; urdf/tall_plant_multi_dof2.urdf
  * observers wired per source ........ utils.py:282 (generated -> CharacterizePlant2 / _mod3),
                                        utils.py:323 (loaded -> CharacterizePlant / _mod),
                                        utils.py:112 (single stems -> TwoAngleRepresentation)

A *stem* is one 2-DOF bending element: a massless ``v1`` link on a revolute-x joint followed by the
stem link on a revolute-y joint (rand_gen...:257,271-272).  Frames compose as in PyBullet's
createMultiBody / URDF: child joint frame = parent *link* frame . origin(xyz, rpy).
"""
from __future__ import annotations

import pickle
import random as _pyrandom
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .transforms import Frame, rot_z, rpy_to_matrix
from .urdf import UrdfTree, parse_urdf

OBS_MOD3 = 0       # representation.py:191-235  (quaternion difference -> roll/pitch norm)
OBS_ANGLE = 1      # representation.py:27-51 and 60-123 (angle between COM direction and a reference direction)

SCALING = 0.25                     # rand_gen...:18
STEM_HALF_WIDTH = SCALING * 0.1    # rand_gen...:74-75  (0.05 m square section)
STEM_COM_Z = SCALING * 1.0         # rand_gen...:265
PLANT_JOINT_LIMIT = 1.5            # rand_gen...:307, utils.py:298
RAISED_HEIGHT = 0.25               # representation.py:12
DEFAULT_STEM_RADIUS = 2.0 * STEM_HALF_WIDTH / np.sqrt(np.pi)   # capsule with the box's cross-section area


@dataclass
class Stem:
    parent: int                    # index of the parent stem in the same world, -1 = rooted in the world
    origin: Frame                  # joint frame in the parent's link frame (world frame for roots)
    half_height: float             # box half length along +z
    z_offset: float                # box starts at this z in the stem frame (stem_base_spacing)
    com_z: float = STEM_COM_Z
    limit: float = PLANT_JOINT_LIMIT
    half_width: float = STEM_HALF_WIDTH
    plant_start: bool = False      # first stem of a plant: the observer chain restarts (representation.py:270)
    is_main: bool = False          # (2e+2) in main_stem_indices (representation.py:289)
    obs_kind: int = OBS_MOD3
    obs_ref_point: np.ndarray = field(default_factory=lambda: np.zeros(3))   # OBS_ANGLE: base point
    obs_snap: bool = False         # OBS_ANGLE: zero |x|,|y| < 1e-5 (representation.py:41-44)
    obs_vertical: bool = False     # OBS_ANGLE: reference direction is +z (TwoAngleRepresentation) else rest direction
    # joint order bookkeeping: the kernel model is always Rx(tx) Ry(ty).  A URDF stem whose first
    # joint is about y (urdf/tall_plant_multi_dof2.urdf:62-82) is mapped onto it with a constant
    # yaw of +90deg folded into ``origin`` (Ry(a)Rx(b) = Rz90 Rx(a) Ry(-b) Rz90^T); then
    # joint angles (first, second) = (tx, -ty).
    y_first: bool = False


@dataclass
class PlantWorld:
    """Everything that differs between sampled worlds."""
    stems: List[Stem]
    base_boxes: List[Tuple[Frame, np.ndarray]]          # (pose, half extents) of immovable plant bases
    # reference bookkeeping, kept for interchange with plant_params.pkl (utils.py:274)
    joint_list: List[int] = field(default_factory=list)
    main_stem_indices: List[int] = field(default_factory=list)
    base_points: Dict[int, List[float]] = field(default_factory=dict)
    base_pos: List[float] = field(default_factory=list)
    link_parent_indices: Dict[int, int] = field(default_factory=dict)

    @property
    def num_stems(self) -> int:
        return len(self.stems)

    def shifted(self, dx: float, dy: float) -> "PlantWorld":
        """The same plant translated in xy (utils.py:306-317 applies base_pos_offset_xy)."""
        import copy
        out = copy.deepcopy(self)
        d = np.array([dx, dy, 0.0])
        for s in out.stems:
            if s.parent < 0:
                s.origin = Frame(s.origin.R, s.origin.t + d)
            s.obs_ref_point = s.obs_ref_point + d
        out.base_boxes = [(Frame(f.R, f.t + d), h) for f, h in out.base_boxes]
        for k in out.base_points:
            out.base_points[k] = [out.base_points[k][0] + dx, out.base_points[k][1] + dy, out.base_points[k][2]]
        if out.base_pos:
            out.base_pos = [out.base_pos[0] + dx, out.base_pos[1] + dy, out.base_pos[2]]
        return out


# --------------------------------------------------------------------------- sampler
class _Rng:
    """The two global streams the reference draws from (numpy.random and random)."""

    def __init__(self, np_state=None, py_state=None):
        self.np = np.random if np_state is None else np_state
        self.py = _pyrandom if py_state is None else py_state

    def uniform(self, low, high):
        return float(self.np.uniform(low=low, high=high))

    def coin(self):
        return self.py.random()

    def choice(self, seq):
        return self.py.choice(seq)


def sample_plant(num_branches_per_stem: int, total_num_vert_stems: int, total_num_extensions: int,
                 base_pos_xy: Sequence[float], stem_base_spacing: float = 0.0,
                 np_random=None, py_random=None) -> PlantWorld:
    """One random multi-branch plant; consumes the RNG streams in the reference's order.

    Draw order (np = numpy.random.uniform, py = random.random / random.choice), per
    rand_gen_plants_programmatically_main.py:84-86, 32, 158-162, 183-188, 196-213, 237-244:
      np x3 (base half dims), then per created stem one np (half height) *after* its placement draws;
      branch: [py, py-choice, np] repeated until clear of earlier branches, np (yaw), np (pitch|roll), np (z);
      extension: py, np (roll|pitch);   stacked stem: py, np (pitch|roll), np (yaw).
    ``np_random`` / ``py_random`` default to the global modules (what the reference uses), so
    ``np.random.seed(s); random.seed(s)`` reproduces its plants.
    """
    if total_num_extensions < 1 or num_branches_per_stem < 0 or total_num_vert_stems < 1:
        # with zero extensions the reference's loop never terminates (rand_gen...:128-151)
        raise ValueError("need total_num_extensions >= 1 and total_num_vert_stems >= 1")
    rng = _Rng(np_random, py_random)
    sf = SCALING
    bx, by = float(base_pos_xy[0]), float(base_pos_xy[1])
    half_w = rng.uniform(sf * 0.07, sf * 0.15)
    half_l = rng.uniform(sf * 0.07, sf * 0.15)
    half_h = rng.uniform(sf * 0.07, sf * 0.15)
    base_pos = [bx, by, half_h]

    stems: List[Stem] = []
    heights: List[float] = []          # half height per stem, creation order
    link_parents: List[int] = []       # createMultiBody linkParentIndices (1-based, 0 = base)
    base_points: Dict[int, List[float]] = {}
    link_parent_idx: Dict[int, int] = {}
    mains: List[int] = []              # 1-based link numbers of main stems ("main_stem_indices")

    main = -1                          # stem index of the current main stem
    taken: List[Tuple[float, float]] = []
    branches_done = 0                  # branches completed on the current main stem
    ext_left = 0                       # extensions still owed to the current branch
    verts = 0

    def add(parent: int, xyz, rpy, is_main: bool) -> int:
        k = len(stems)
        v1_link, stem_link = 2 * k + 1, 2 * k + 2             # 1-based link numbers
        link_parents.extend([0 if parent < 0 else 2 * parent + 2, v1_link])
        base_points[stem_link - 1] = [bx, by, float(xyz[2]) if parent >= 0 else 2.0 * float(xyz[2])]
        link_parent_idx[stem_link - 1] = -1 if parent < 0 else 2 * parent + 1
        h = rng.uniform(sf * 0.7, sf * 1.0)                   # create_stem_element (rand_gen...:32)
        heights.append(h)
        stems.append(Stem(parent=parent, origin=Frame(rpy_to_matrix(rpy), np.asarray(xyz, float)),
                          half_height=h, z_offset=stem_base_spacing, plant_start=(k == 0), is_main=is_main,
                          obs_kind=OBS_MOD3))
        if is_main:
            mains.append(stem_link)
        return k

    while True:
        if not stems:                                          # first stem, centred on the base
            main = add(-1, (0.0, 0.0, half_h), (0.0, 0.0, 0.0), True)
            verts = 1
            ext_left = 0
            branches_done = 0
        elif branches_done < num_branches_per_stem:
            if ext_left > 0:                                   # extend the last stem of this branch
                prev = len(stems) - 1
                z = stem_base_spacing + 2.0 * heights[prev]
                if rng.coin() < 0.5:
                    rpy = (rng.uniform(sf * -0.5, sf * 0.5), 0.0, 0.0)
                else:
                    rpy = (0.0, rng.uniform(sf * -0.5, sf * 0.5), 0.0)
                add(prev, (0.0, 0.0, z), rpy, False)
                ext_left -= 1
                if ext_left == 0:
                    branches_done += 1
            else:                                              # start a new side branch on the main stem
                while True:
                    if rng.coin() < 0.5:
                        x = rng.choice([sf * 0.2, sf * -0.2])
                        y = rng.uniform(0.0, sf * 0.2)
                    else:
                        y = rng.choice([sf * 0.2, sf * -0.2])
                        x = rng.uniform(0.0, sf * 0.2)
                    tol = sf * 0.2
                    if any(abs(x - px) < tol and abs(y - py_) < tol for px, py_ in taken):
                        continue
                    taken.append((x, y))
                    break
                yaw = rng.uniform(-0.5, 0.5)
                if abs(x) > abs(y):
                    pitch = rng.uniform(0.7, 1.5) if x > 0 else rng.uniform(-1.5, -0.7)
                    roll = 0.0
                else:
                    roll = rng.uniform(-1.5, -0.7) if y > 0 else rng.uniform(0.7, 1.5)
                    pitch = 0.0
                z = rng.uniform(stem_base_spacing, stem_base_spacing + 2.0 * heights[main])
                add(main, (x, y, z), (roll, pitch, yaw), False)
                ext_left = total_num_extensions
        elif verts == total_num_vert_stems:
            break
        else:                                                  # stack the next main stem on top
            verts += 1
            branches_done = 0
            taken = []
            z = stem_base_spacing + 2.0 * heights[main]
            if rng.coin() < 0.5:
                pitch, roll = rng.uniform(-0.4, 0.4), 0.0
            else:
                roll, pitch = rng.uniform(-0.4, 0.4), 0.0
            yaw = rng.uniform(-0.4, 0.4)
            main = add(main, (0.0, 0.0, z), (roll, pitch, yaw), True)

    # root stem origin is expressed in the world: base link frame (at base_pos) . (0,0,half_h)
    stems[0].origin = Frame(stems[0].origin.R, stems[0].origin.t + np.array(base_pos))
    world = PlantWorld(
        stems=stems,
        base_boxes=[(Frame(np.eye(3), np.array(base_pos)), np.array([half_w, half_l, half_h]))],
        joint_list=link_parents, main_stem_indices=mains, base_points=base_points, base_pos=base_pos,
        link_parent_indices=link_parent_idx)
    return world


def sample_world(num_branches_per_stem: int, total_num_vert_stems: int, total_num_extensions: int,
                 plant_pos_xy_limits, base_offset_xy=(0.0, 0.0), stem_base_spacing: float = 0.0,
                 np_random=None, py_random=None) -> PlantWorld:
    """SinglePlantEnv's world draw: plant xy first (two uniforms, drawn even when low == high,
    env.py:63-64), then the plant.  Coordinates are kept *relative to the world's base offset*
    because every body of a world is shifted by the same offset (env.py:38-50,66-70)."""
    rs = np.random if np_random is None else np_random
    (x0, x1), (y0, y1) = plant_pos_xy_limits
    px = float(rs.uniform(low=x0, high=x1))
    py = float(rs.uniform(low=y0, high=y1))
    del base_offset_xy  # relative coordinates: the offset cancels
    return sample_plant(num_branches_per_stem, total_num_vert_stems, total_num_extensions, (px, py),
                        stem_base_spacing, np_random, py_random)


# --------------------------------------------------------------------------- loaders
def plant_from_urdf_tree(tree: UrdfTree, base_frame: Frame) -> Tuple[List[Stem], List[Tuple[Frame, np.ndarray]]]:
    """Read a plant URDF whose links alternate (massless v1, stem) below a base link."""
    L = len(tree.joints)
    if L % 2 != 0:
        raise ValueError("plant URDF must have an even number of joints (two per stem)")
    stems: List[Stem] = []
    link_to_stem: Dict[int, int] = {}
    for k in range(L // 2):
        j1, j2 = tree.joints[2 * k], tree.joints[2 * k + 1]
        if j2.parent != j1.child:
            raise ValueError("expected (v1, stem) joint pairs in link order")
        a1, a2 = np.round(j1.axis), np.round(j2.axis)
        pidx = tree.link_index(j1.parent)
        parent = -1 if pidx < 0 else link_to_stem[pidx]
        origin = j1.origin @ Frame(j2.origin.R, j2.origin.t)  # second joint origin is identity in all fixtures
        if abs(a1[0]) == 1 and abs(a2[1]) == 1:
            y_first = False
        elif abs(a1[1]) == 1 and abs(a2[0]) == 1:
            y_first = True
            origin = origin @ Frame(rot_z(np.pi / 2))
        else:
            raise ValueError(f"unsupported stem joint axes {j1.axis} / {j2.axis}")
        link = tree.links[j2.child]
        boxes = [g for g in link.collisions if g.kind == "box"]
        if len(boxes) != 1:
            raise ValueError("stem link must carry one box")
        size = boxes[0].size
        zc = float(boxes[0].origin.t[2])
        hh = 0.5 * float(size[2])
        lim = PLANT_JOINT_LIMIT
        if parent < 0:
            origin = base_frame @ origin
        elif stems[parent].y_first:
            origin = Frame(rot_z(-np.pi / 2)) @ origin
        stems.append(Stem(parent=parent, origin=origin, half_height=hh, z_offset=zc - hh,
                          com_z=float(link.inertial_origin.t[2]), limit=lim,
                          half_width=0.5 * float(size[0]), plant_start=(k == 0), y_first=y_first))
        link_to_stem[2 * k + 1] = k
    base_boxes = []
    for g in tree.links[tree.root].collisions:
        if g.kind == "box" and float(np.max(g.size)) > 0:
            base_boxes.append((base_frame @ g.origin, 0.5 * g.size))
    return stems, base_boxes


def load_plant(urdf_path: str, params_path: str, base_offset_xy=(0.0, 0.0)) -> PlantWorld:
    """``load_plant_from_urdf`` (plant_motion_planning/utils.py:287-325): URDF + plant_params.pkl.

    The base is placed at ``base_pos`` from the pickle (set_pose on a base whose inertial frame
    coincides with its link frame), observers are CharacterizePlant / TwoAngleRepresentation_mod
    with ``base_points`` from the pickle.
    """
    with open(params_path, "rb") as fp:
        joint_list, main_stem_indices, base_points, base_pos = pickle.load(fp)
    tree = parse_urdf(urdf_path)
    return plant_from_parsed(tree, joint_list, main_stem_indices, base_points, base_pos, base_offset_xy)


def plant_from_parsed(tree: UrdfTree, joint_list, main_stem_indices, base_points, base_pos,
                      base_offset_xy=(0.0, 0.0)) -> PlantWorld:
    ox, oy = float(base_offset_xy[0]), float(base_offset_xy[1])
    del ox, oy  # world-relative coordinates (see sample_world)
    base_frame = Frame(np.eye(3), np.array([base_pos[0], base_pos[1], base_pos[2]], dtype=np.float64))
    stems, base_boxes = plant_from_urdf_tree(tree, base_frame)
    for k, s in enumerate(stems):
        s.obs_kind = OBS_ANGLE
        s.obs_ref_point = np.asarray(base_points[2 * k + 1], dtype=np.float64)
        s.obs_vertical = False
        s.obs_snap = False
        s.is_main = (2 * k + 2) in main_stem_indices
    return PlantWorld(stems=stems, base_boxes=base_boxes, joint_list=list(joint_list),
                      main_stem_indices=list(main_stem_indices),
                      base_points={int(k): list(v) for k, v in base_points.items()},
                      base_pos=list(base_pos))


# Single-stem plant: urdf/tall_plant_multi_dof2.urdf restated as numbers (base box 0.07x0.07x0.1,
# joint rot_y then rot_x at z=0.10 above the base link frame, stem box 0.05x0.05x1.2 centred z=0.6,
# COM z=0.3, limits +-1.571).  short_plant_multi_dof2.urdf differs only in the stem length.
TALL_STEM = dict(base_half=(0.035, 0.035, 0.05), joint_z=0.10, box_len=1.2, com_z=0.3, limit=1.571)


def single_stem_world(positions: Sequence[Sequence[float]], spec: Optional[dict] = None,
                      floor_z: float = 0.0) -> PlantWorld:
    """``generate_tall_plants`` (utils.py:97-117): one single-stem plant per xy position.

    ``set_random_pose`` rests each base on the floor (stable_z => base link frame at
    floor_z + base half height).  Each plant is observed by TwoAngleRepresentation
    (representation.py:27-51): angle from vertical of COM - base_pos - (0,0,RAISED_HEIGHT).
    """
    spec = dict(TALL_STEM if spec is None else spec)
    hb = np.asarray(spec["base_half"], dtype=np.float64)
    stems: List[Stem] = []
    boxes = []
    for (px, py) in positions:
        base = Frame(np.eye(3), np.array([float(px), float(py), floor_z + hb[2]]))
        origin = base @ Frame(np.eye(3), np.array([0.0, 0.0, spec["joint_z"]])) @ Frame(rot_z(np.pi / 2))
        stems.append(Stem(parent=-1, origin=origin, half_height=0.5 * spec["box_len"], z_offset=0.0,
                          com_z=spec["com_z"], limit=spec["limit"], plant_start=True, is_main=False,
                          obs_kind=OBS_ANGLE, obs_ref_point=base.t + np.array([0.0, 0.0, RAISED_HEIGHT]),
                          obs_snap=True, obs_vertical=True, y_first=True))
        boxes.append((base, hb.copy()))
    return PlantWorld(stems=stems, base_boxes=boxes)


# plant_motion_planning/utils.py:120-230 -- fixed plant position sets for the single-branch scripts.
SINGLE_STEM_ENVS: Dict[str, List[List[float]]] = {}


def _load_env_sets() -> None:
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "..", "data", "single_stem_envs.json")
    if os.path.exists(path):
        with open(path) as fp:
            SINGLE_STEM_ENVS.update(json.load(fp))


_load_env_sets()


# --------------------------------------------------------------------------- writers
def _rpy_from_matrix(R: np.ndarray) -> Tuple[float, float, float]:
    """Inverse of rpy_to_matrix (Rz(yaw) Ry(pitch) Rx(roll)), pitch in [-pi/2, pi/2]."""
    sp = -float(R[2, 0])
    if abs(sp) < 1.0 - 1e-12:
        return (float(np.arctan2(R[2, 1], R[2, 2])), float(np.arcsin(sp)), float(np.arctan2(R[1, 0], R[0, 0])))
    return (float(np.arctan2(-R[1, 2], R[1, 1])), float(np.copysign(np.pi / 2, sp)), 0.0)


def plant_urdf_text(world: PlantWorld, decimals: int = 5) -> str:
    """The plant of ``world`` as URDF text in the layout of the saved fixtures (environments/multi_branch/envK/
    plant.urdf, written by urdfEditor.saveUrdf from generate_random_plant, plant_motion_planning/utils.py:270-272):
    link0 = base box, then per stem a massless link (joint about x) and the stem link (joint about y, box of
    0.05 x 0.05 x 2h centred z_offset + h, mass 10, COM at com_z), `continuous` joints with damping 1.0.
    Numbers are printed with ``decimals`` places like urdfEditor does (5)."""
    if len(world.base_boxes) != 1:
        raise ValueError("one plant (one base box) per URDF")
    if any(s.y_first for s in world.stems):
        raise ValueError("only x-then-y stems have a generated-plant URDF form")
    base, half = world.base_boxes[0]
    base_inv = Frame(base.R.T, -base.R.T @ base.t)
    f = lambda v: f"{float(v) + 0.0:.{decimals}f}"
    vec = lambda a: " ".join(f(v) for v in a)

    def link(name, mass, com, box_xyz, box_size, inertia):
        o = f'<origin rpy="{vec((0, 0, 0))}" xyz="{vec(box_xyz)}"/>'
        return (f'\t<link name="{name}">\n\t\t<inertial>\n\t\t\t<origin rpy="{vec((0, 0, 0))}" xyz="{vec(com)}"/>\n'
                f'\t\t\t<mass value="{f(mass)}"/>\n\t\t\t<inertia ixx="{f(inertia[0])}" ixy="0" ixz="0" iyy="{f(inertia[1])}" '
                f'iyz="0" izz="{f(inertia[2])}"/>\n\t\t</inertial>\n\t\t<visual>\n\t\t\t{o}\n\t\t\t<geometry>\n'
                f'\t\t\t\t<box size="{vec(box_size)}"/>\n\t\t\t</geometry>\n\t\t</visual>\n\t\t<collision>\n\t\t\t{o}\n'
                f'\t\t\t<geometry>\n\t\t\t\t<box size="{vec(box_size)}"/>\n\t\t\t</geometry>\n\t\t</collision>\n\t</link>\n')

    def joint(k, parent, child, origin: Frame, axis):
        return (f'\t<joint name="joint{k}" type="continuous">\n\t\t<parent link="link{parent}"/>\n\t\t<child link="link{child}"/>\n'
                f'\t\t<dynamics damping="1.0" friction="0.0001"/>\n'
                f'\t\t<origin rpy="{vec(_rpy_from_matrix(origin.R))}" xyz="{vec(origin.t)}"/>\n'
                f'\t\t<axis xyz="{vec(axis)}"/>\n\t</joint>\n')

    def box_inertia(m, sx, sy, sz):
        return (m * (sy * sy + sz * sz) / 12.0, m * (sx * sx + sz * sz) / 12.0, m * (sx * sx + sy * sy) / 12.0)

    size0 = 2.0 * np.asarray(half)
    links = [link("link0", 1e9, (0, 0, 0), (0, 0, 0), size0, box_inertia(1e9, *size0))]
    joints = []
    for k, s in enumerate(world.stems):
        size = (2 * s.half_width, 2 * s.half_width, 2 * s.half_height)
        links.append(link(f"link{2 * k + 1}", 0.0, (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0)))
        links.append(link(f"link{2 * k + 2}", 10.0, (0, 0, s.com_z), (0, 0, s.z_offset + s.half_height), size,
                          box_inertia(10.0, *size)))
        origin = (base_inv @ s.origin) if s.parent < 0 else s.origin
        parent_link = 0 if s.parent < 0 else 2 * s.parent + 2
        joints.append(joint(2 * k + 1, parent_link, 2 * k + 1, origin, (1, 0, 0)))
        joints.append(joint(2 * k + 2, 2 * k + 1, 2 * k + 2, Frame(np.eye(3), np.zeros(3)), (0, 1, 0)))
    return '<?xml version="0.0" ?>\n<robot name="">\n' + "".join(links) + "".join(joints) + "</robot>\n"


def save_plant(world: PlantWorld, urdf_path: str, params_path: str) -> None:
    """Write ``plant.urdf`` + ``plant_params.pkl`` = [joint_list, main_stem_indices, base_points, base_pos], the pair
    generate_random_plant(save_plant=True) leaves behind (plant_motion_planning/utils.py:270-279) and
    load_plant_from_urdf / load_plant read back."""
    if not world.base_pos or not world.base_points:
        raise ValueError("world carries no plant_params bookkeeping (base_pos / base_points)")
    with open(urdf_path, "w") as fp:
        fp.write(plant_urdf_text(world))
    params = [list(world.joint_list), list(world.main_stem_indices),
              {int(k): [float(c) for c in v] for k, v in world.base_points.items()}, [float(c) for c in world.base_pos]]
    with open(params_path, "wb") as fp:
        pickle.dump(params, fp)
