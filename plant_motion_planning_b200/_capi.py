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
        ("dense", C.c_int32), ("gate_seed", C.c_uint32), ("gate_probability", C.c_float), ("reserved", C.c_int32),
    ]


class LaunchInfo(C.Structure):
    _fields_ = [
        ("grid", C.c_int32), ("block", C.c_int32), ("smem_bytes", C.c_int32), ("regs_per_thread", C.c_int32),
        ("worlds_in_smem", C.c_int32), ("n_sphere_pad", C.c_int32), ("n_slot_pad", C.c_int32),
        ("kernel_launches", C.c_int64),
    ]


EXPORTS = [
    "pmp_create", "pmp_destroy", "pmp_last_error", "pmp_set_robot", "pmp_set_worlds", "pmp_set_params",
    "pmp_num_worlds", "pmp_max_stems", "pmp_fk", "pmp_check_configs", "pmp_validate_edges",
    "pmp_validate_edges_host", "pmp_check_configs_host", "pmp_get_launch_info", "pmp_measure_fp32_peak",
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
    lib.pmp_validate_edges.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32] + [vp] * 9
    lib.pmp_validate_edges_host.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32] + [vp] * 8
    lib.pmp_check_configs_host.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp]
    lib.pmp_get_launch_info.argtypes = [vp, C.POINTER(LaunchInfo)]
    lib.pmp_measure_fp32_peak.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pmp_nearest.argtypes = [vp, vp, C.c_int32, vp, vp, C.c_int32, vp, vp, vp, vp]
    lib.pmp_edge_points.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int32, vp, vp]
    lib.pmp_tree_extend.argtypes = [vp, C.POINTER(Tree), C.c_int32, vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
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
