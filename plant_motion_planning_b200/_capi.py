"""ctypes binding of libpmp_edgecheck.so (include/pmp_edgecheck.h).  Thin: structs + prototypes.

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded, importing the
compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _build

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_uint32_p = C.POINTER(C.c_uint32)
c_uint8_p = C.POINTER(C.c_uint8)


class RobotDesc(C.Structure):
    _fields_ = [
        ("dof", C.c_int32), ("n_spheres", C.c_int32), ("n_links", C.c_int32), ("n_fixed", C.c_int32),
        ("joint_C", c_float_p), ("joint_A", c_float_p), ("joint_B", c_float_p), ("joint_t", c_float_p),
        ("lower", c_float_p), ("upper", c_float_p), ("circular", c_uint8_p),
        ("base_R", c_float_p), ("base_t", c_float_p),
        ("sph_joint", c_int32_p), ("sph_center", c_float_p), ("sph_radius", c_float_p), ("self_mask", c_uint32_p),
        ("link_joint", c_int32_p), ("link_R", c_float_p), ("link_t", c_float_p), ("link_com", c_float_p),
        ("ee_joint", C.c_int32), ("ee_local", C.c_float * 3),
        ("fixed_kind", c_int32_p), ("fixed_R", c_float_p), ("fixed_t", c_float_p), ("fixed_half", c_float_p),
        ("fixed_mask", c_uint32_p),
    ]


class WorldDesc(C.Structure):
    _fields_ = [
        ("n_worlds", C.c_int32), ("max_stems", C.c_int32), ("max_boxes", C.c_int32), ("n_slots", C.c_int32),
        ("n_stems", c_int32_p), ("stems", c_float_p), ("n_boxes", c_int32_p), ("boxes", c_float_p),
    ]


class Params(C.Structure):
    _fields_ = [
        ("resolution", C.c_double), ("deflection_limit", C.c_float), ("alpha", C.c_float),
        ("max_step", C.c_float), ("reg_eps", C.c_float), ("iterations", C.c_int32), ("mode", C.c_int32),
        ("dense", C.c_int32), ("gate_seed", C.c_uint32), ("gate_probability", C.c_float),
        ("arm_lag_gain", C.c_float), ("arm_lag_substeps", C.c_int32),
    ]


class HullDesc(C.Structure):
    _fields_ = [
        ("n_hulls", C.c_int32), ("hull_joint", c_int32_p), ("hull_off", c_int32_p), ("hull_verts", c_float_p),
        ("hull_margin", c_float_p), ("n_pairs", C.c_int32), ("pairs", c_int32_p), ("sph_hull", c_int32_p),
        ("sph_radius_outer", c_float_p), ("fixed_margin", C.c_float), ("n_fixed_exact", C.c_int32),
        ("sph_radius_inner", c_float_p),
    ]


SAMPLER_MAX_STEMS, SAMPLER_TRIES, SAMPLER_HEAD = 8, 16, 5
SAMPLER_PER_STEM = 8 + 3 * SAMPLER_TRIES


class SamplerDesc(C.Structure):
    _I8 = C.c_int32 * SAMPLER_MAX_STEMS
    _fields_ = [
        ("n_stems", C.c_int32), ("parent", _I8), ("kind", _I8), ("ref", _I8), ("flags", _I8), ("pslot", _I8), ("oslot", _I8),
        ("n_slots", C.c_int32), ("gravity_sag", C.c_int32), ("xy_lo", C.c_double * 2), ("xy_hi", C.c_double * 2),
        ("spacing", C.c_double), ("scaling", C.c_double), ("half_width", C.c_double), ("com_z", C.c_double),
        ("limit", C.c_double), ("stem_margin", C.c_double), ("kp", C.c_double), ("mass", C.c_double), ("g", C.c_double),
    ]


class LaunchInfo(C.Structure):
    _fields_ = [
        ("grid", C.c_int32), ("block", C.c_int32), ("smem_bytes", C.c_int32), ("regs_per_thread", C.c_int32),
        ("worlds_in_smem", C.c_int32), ("n_sphere_pad", C.c_int32), ("n_slot_pad", C.c_int32),
        ("kernel_launches", C.c_int64),
    ]


EXPORTS = [
    "pmp_create", "pmp_destroy", "pmp_last_error", "pmp_set_robot", "pmp_set_robot_hulls", "pmp_set_worlds", "pmp_set_params",
    "pmp_num_worlds", "pmp_max_stems", "pmp_fk", "pmp_check_configs", "pmp_validate_edges",
    "pmp_validate_edges_host", "pmp_check_configs_host", "pmp_get_launch_info", "pmp_measure_fp32_peak",
    "pmp_set_kernel_timing", "pmp_get_kernel_times", "pmp_sample_worlds", "pmp_set_worlds_dev",
    "pmp_nearest", "pmp_edge_points", "pmp_tree_extend",
]


class Tree(C.Structure):
    """pmp_tree: caller-owned device buffers of one search tree."""
    _fields_ = [("nodes", C.c_void_p), ("parent", C.c_void_p), ("via", C.c_void_p), ("via_steps", C.c_void_p),
                ("size", C.c_void_p), ("capacity", C.c_int32), ("reserved", C.c_int32)]

_lib: Optional[C.CDLL] = None


def library_path() -> str:
    # PMP_LIB: load another build of the same library (kernel tuning experiments only)
    return os.environ.get("PMP_LIB") or _build.LIB_PATH


def load() -> C.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  plant_motion_planning_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.pmp_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.pmp_destroy.argtypes = [vp]
    lib.pmp_last_error.argtypes = [vp]
    lib.pmp_last_error.restype = C.c_char_p
    lib.pmp_set_robot.argtypes = [vp, C.POINTER(RobotDesc)]
    lib.pmp_set_worlds.argtypes = [vp, C.POINTER(WorldDesc)]
    lib.pmp_set_params.argtypes = [vp, C.POINTER(Params)]
    lib.pmp_num_worlds.argtypes = [vp]
    lib.pmp_max_stems.argtypes = [vp]
    lib.pmp_fk.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp]
    lib.pmp_check_configs.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.pmp_set_robot_hulls.argtypes = [vp, C.POINTER(HullDesc)]
    lib.pmp_validate_edges.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32] + [vp] * 9 + [C.c_int32]
    lib.pmp_validate_edges_host.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32] + [vp] * 8
    lib.pmp_check_configs_host.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp]
    lib.pmp_get_launch_info.argtypes = [vp, C.POINTER(LaunchInfo)]
    lib.pmp_measure_fp32_peak.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pmp_nearest.argtypes = [vp, vp, C.c_int32, vp, vp, C.c_int32, vp, vp, vp, vp]
    lib.pmp_edge_points.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int32, vp, vp]
    lib.pmp_tree_extend.argtypes = [vp, C.POINTER(Tree), C.c_int32, vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
    lib.pmp_sample_worlds.argtypes = [vp, C.POINTER(SamplerDesc), C.c_int32, C.c_uint64, vp, C.c_int32, vp, vp, vp]
    lib.pmp_set_worlds_dev.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    lib.pmp_set_kernel_timing.argtypes = [vp, C.c_int32]
    lib.pmp_get_kernel_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for name in EXPORTS:
        if name != "pmp_last_error":
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_float_p)


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


def uptr(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_uint32_p)


def bptr(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_uint8_p)


def robot_desc(t) -> RobotDesc:
    """pmp_robot_desc over the arrays of a model.tables.RobotTables (which must outlive the call it is passed to)."""
    d = RobotDesc()
    d.dof, d.n_spheres, d.n_links, d.n_fixed = t.dof, t.n_spheres, t.n_links, t.n_fixed
    d.joint_C, d.joint_A, d.joint_B = fptr(t.joint_C), fptr(t.joint_A), fptr(t.joint_B)
    d.joint_t, d.lower, d.upper = fptr(t.joint_t), fptr(t.lower), fptr(t.upper)
    d.circular = bptr(t.circular)
    d.base_R, d.base_t = fptr(t.base_R), fptr(t.base_t)
    d.sph_joint, d.sph_center = iptr(t.sph_joint), fptr(t.sph_center)
    d.sph_radius, d.self_mask = fptr(t.sph_radius), uptr(t.self_mask)
    d.link_joint, d.link_R = iptr(t.link_joint), fptr(t.link_R)
    d.link_t, d.link_com = fptr(t.link_t), fptr(t.link_com)
    d.ee_joint = t.ee_joint
    d.ee_local = (C.c_float * 3)(*[float(v) for v in t.ee_local])
    d.fixed_kind, d.fixed_R = iptr(t.fixed_kind), fptr(t.fixed_R)
    d.fixed_t, d.fixed_half = fptr(t.fixed_t), fptr(t.fixed_half)
    d.fixed_mask = uptr(t.fixed_mask)
    return d


def hull_desc(hulls, n_fixed_exact: int):
    """(pmp_hull_desc, keep-alive list) for a model.robot.HullModel; `n_fixed_exact` = number of obstacle primitives
    at the head of the static-primitive table."""
    H = len(hulls.verts)
    off = np.zeros(H + 1, dtype=np.int32)
    for i, v in enumerate(hulls.verts):
        off[i + 1] = off[i] + len(v)
    keep = {
        "joint": np.ascontiguousarray(hulls.joint, dtype=np.int32), "off": off,
        "verts": np.ascontiguousarray(np.concatenate(hulls.verts, axis=0), dtype=np.float32).reshape(-1),
        "margin": np.ascontiguousarray(hulls.margin, dtype=np.float32),
        "pairs": np.ascontiguousarray(hulls.pairs, dtype=np.int32).reshape(-1),
        "sph_hull": np.ascontiguousarray(hulls.sph_hull, dtype=np.int32),
        "outer": np.ascontiguousarray(hulls.sph_radius_outer, dtype=np.float32),
    }
    inner = getattr(hulls, "sph_radius_inner", None)
    if inner is not None:
        keep["inner"] = np.ascontiguousarray(inner, dtype=np.float32)
    d = HullDesc()
    d.n_hulls = H
    d.hull_joint, d.hull_off, d.hull_verts = iptr(keep["joint"]), iptr(keep["off"]), fptr(keep["verts"])
    d.hull_margin = fptr(keep["margin"])
    d.n_pairs = len(hulls.pairs)
    d.pairs, d.sph_hull, d.sph_radius_outer = iptr(keep["pairs"]), iptr(keep["sph_hull"]), fptr(keep["outer"])
    d.fixed_margin = float(hulls.fixed_margin)
    d.n_fixed_exact = int(n_fixed_exact)
    d.sph_radius_inner = fptr(keep["inner"]) if "inner" in keep else None
    return d, keep
