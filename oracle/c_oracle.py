"""ctypes loader for the C twin of the oracle (oracle/edgecheck_oracle.c).   *** TEST INFRASTRUCTURE ***

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from plant_motion_planning_b200 import _capi
from plant_motion_planning_b200.model.tables import MODES, pack_robot, pack_worlds

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_edgecheck.so")


def build(force: bool = False) -> str:
    deps = [os.path.join(HERE, "edgecheck_oracle.c"), os.path.join(HERE, "hull_gjk.c"),
            os.path.join(HERE, "..", "include", "pmp_edgecheck.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        # -march=native binaries are rebuilt on the machine that runs them when possible
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB


_libs = {}
NATIVE_LIB = os.path.join(HERE, "_build", "liboracle_edgecheck_native.so")
SINGLE_LIB = os.path.join(HERE, "_build", "liboracle_edgecheck_f32.so")
NATIVE_SINGLE_LIB = os.path.join(HERE, "_build", "liboracle_edgecheck_f32_native.so")


def build_native() -> bool:
    """Rebuild with -march=native on THIS machine (used by bench.py for the timed CPU baseline)."""
    try:
        subprocess.run(["make", "-C", HERE, "native"], check=True, capture_output=True)
        return os.path.exists(NATIVE_LIB) and os.path.exists(NATIVE_SINGLE_LIB)
    except Exception:
        return False


def load(native: bool = False, single: bool = False):
    """The C twin: double precision (parity tests) or, with ``single``, the float32 build (CPU baseline)."""
    key = (bool(native), bool(single))
    if key not in _libs:
        path = SINGLE_LIB if single else LIB
        if native and build_native():
            path = NATIVE_SINGLE_LIB if single else NATIVE_LIB
        else:
            build()
        try:
            lib = C.CDLL(path)
        except OSError:
            build(force=True)
            lib = C.CDLL(SINGLE_LIB if single else LIB)
        vp = C.c_void_p
        lib.pmp_oracle_validate_edges.argtypes = [C.POINTER(_capi.RobotDesc), C.POINTER(_capi.WorldDesc),
                                                  C.POINTER(_capi.Params), vp, vp, C.c_int32, C.c_int32, C.c_int32] + \
            [vp] * 8 + [C.c_int32, vp]
        lib.pmp_oracle_check_configs.argtypes = [C.POINTER(_capi.RobotDesc), C.POINTER(_capi.WorldDesc),
                                                 C.POINTER(_capi.Params), vp, C.c_int32, vp, vp, vp, vp, C.c_int32, vp]
        lib.pmp_oracle_validate_edges.restype = C.c_int
        lib.pmp_oracle_check_configs.restype = C.c_int
        lib.pmp_oracle_max_threads.restype = C.c_int
        _libs[key] = lib
    return _libs[key]


class COracle:
    """Holds the packed tables (same packer as the CUDA path) and calls the C twin."""

    def __init__(self, robot, worlds, obstacles=None, resolution=np.radians(3), deflection_limit=0.30, alpha=0.1,
                 iterations=4, max_step=0.5, reg_eps=1e-4, stem_margin=1e-3, gravity_sag=True, arm_lag_gain=0.0,
                 arm_lag_substeps=50, mode="our_method", dense=False, native=False, use_hulls=True, single=False):
        """``single``: the float32 build of the twin (the CPU baseline of bench.py); parity tests use the double one."""
        self.lib = load(native, single)
        self.rt = pack_robot(robot, obstacles)
        from plant_motion_planning_b200.model.tables import WorldTables, default_obstacles
        self.wt = worlds if isinstance(worlds, WorldTables) else pack_worlds(list(worlds), stem_margin, gravity_sag)
        self.rd = _capi.robot_desc(self.rt)
        self.hd, self._hull_keep = None, None
        if use_hulls and getattr(robot, "hulls", None) is not None:
            n_obs = len(default_obstacles() if obstacles is None else list(obstacles))
            self.hd, self._hull_keep = _capi.hull_desc(robot.hulls, n_obs)
        w = _capi.WorldDesc()
        self._stems = np.ascontiguousarray(self.wt.stems.reshape(-1))
        self._boxes = np.ascontiguousarray(self.wt.boxes.reshape(-1))
        w.n_worlds, w.max_stems, w.max_boxes, w.n_slots = self.wt.n_worlds, self.wt.max_stems, self.wt.max_boxes, self.wt.n_slots
        w.n_stems, w.stems = _capi.iptr(self.wt.n_stems), _capi.fptr(self._stems)
        w.n_boxes, w.boxes = _capi.iptr(self.wt.n_boxes), _capi.fptr(self._boxes)
        self.wd = w
        self.prm = _capi.Params(float(resolution), deflection_limit, alpha, max_step, reg_eps, iterations, MODES[mode],
                                1 if dense else 0, 0, 0.0, float(arm_lag_gain), int(arm_lag_substeps))

    def _hp(self):
        return C.byref(self.hd) if self.hd is not None else None

    @property
    def max_threads(self) -> int:
        return int(self.lib.pmp_oracle_max_threads())

    def validate_edges(self, q0, q1, backward=False, max_steps=320, want_masks=False, n_threads=0) -> dict:
        q0 = np.ascontiguousarray(q0, dtype=np.float32).reshape(-1, self.rt.dof)
        q1 = np.ascontiguousarray(q1, dtype=np.float32).reshape(-1, self.rt.dof)
        E, W, Cw = q0.shape[0], self.wt.n_worlds, (max_steps + 31) // 32
        out = {"first_hit": np.empty(E, np.int32), "n_steps": np.empty(E, np.int32),
               "max_deflection": np.empty((E, W), np.float32), "over_limit": np.empty((E, W), np.float32),
               "ee_len": np.empty(E, np.float32), "cost": np.empty(E, np.float32)}
        if want_masks:
            out["rigid_mask"] = np.zeros((E, W, Cw), np.uint32)
            out["exceed_mask"] = np.zeros((E, W, Cw), np.uint32)
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p()
        rc = self.lib.pmp_oracle_validate_edges(
            C.byref(self.rd), C.byref(self.wd), C.byref(self.prm), vp(q0), vp(q1), E, max_steps, 1 if backward else 0,
            vp(out["first_hit"]), vp(out["n_steps"]), vp(out["max_deflection"]), vp(out["over_limit"]), vp(out["ee_len"]),
            vp(out["cost"]), vp(out.get("rigid_mask")), vp(out.get("exceed_mask")), int(n_threads), self._hp())
        assert rc == 0
        return out

    def check_configs(self, q, n_threads=0) -> dict:
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.rt.dof)
        B, W, P = q.shape[0], self.wt.n_worlds, self.wt.max_stems
        out = {"flags": np.empty((B, W), np.uint8), "deflection": np.empty((B, W, P), np.float64),
               "theta": np.empty((B, W, P, 2), np.float64), "ee": np.empty((B, 3), np.float64)}
        vp = lambda a: C.c_void_p(a.ctypes.data)
        rc = self.lib.pmp_oracle_check_configs(C.byref(self.rd), C.byref(self.wd), C.byref(self.prm), vp(q), B,
                                               vp(out["flags"]), vp(out["deflection"]), vp(out["theta"]), vp(out["ee"]),
                                               int(n_threads), self._hp())
        assert rc == 0
        return out
