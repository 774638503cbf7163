"""EdgeChecker: the batched edge-validation engine (PyTorch tensors in, PyTorch tensors out).

Host side of the hot path named in BASELINE.json: owns one C-ABI context (robot + obstacle tables,
sampled plant worlds, constants) and launches the sm_100a kernels on torch's current stream.
PyTorch is plumbing only (device memory, streams); all arithmetic is in csrc/.

Additive batched surface:
    validate_edges(q_from[E,D], q_to[E,D]) -> first_hit[E], n_steps[E], max_deflection[E,W],
                                              over_limit[E,W], ee_len[E], cost[E], optional step masks
    check_configs(q[B,D])                  -> flags[B,W], deflection[B,W,P], theta[B,W,P,2], ee[B,3]
    fk(q[B,D])                             -> link positions / quaternions / COM positions
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, replace
from typing import Optional, Sequence

import numpy as np

from . import _capi
from .model.plants import PlantWorld
from .model.robot import FixedPrim, RobotModel
from .model.tables import MODES, RobotTables, WorldTables, pack_robot, pack_worlds

FLAG_RIGID = 1
FLAG_EXCEEDED = 2


@dataclass(frozen=True)
class EdgeCheckConfig:
    """Constants the reference hard-codes ("Config / flags")."""
    resolution: float = math.radians(3)   # DEFAULT_RESOLUTION, pybullet_tools/utils.py:3660
    deflection_limit: float = 0.30        # scripts/*: 0.30; multi_world_our_method.py: 2.0
    alpha: float = 0.1                    # kuka_primitives.py:476
    iterations: int = 4                   # Gauss-Newton steps of the quasi-static plant solve
    max_step: float = 0.5                 # rad per step
    reg_eps: float = 1e-4                 # m^2
    stem_margin: float = 1.0e-3           # Bullet's collision margin of the stem boxes (carved out of the box)
    gravity_sag: bool = True              # stems rest in their spring / gravity balance (model/statics.py), not at 0
    arm_lag_gain: float = 0.0             # 0 = the arm is at the commanded configuration; the reference's position
                                          # motor closes 1 % of the joint error per sub-step (env.py:154,170): 0.01
    arm_lag_substeps: int = 50            # env.py:157, 216
    use_hulls: bool = True                # decide self / obstacle collisions exactly on the convex hulls of the arm's
                                          # collision meshes when the robot model carries them (else on its spheres)
    mode: str = "our_method"              # our_method | avoid_all | ignore_all
    gate_probability: float = 0.95        # pybullet_tools/utils.py:3843 (applied by the host adapters only)
    max_steps: int = 320                  # per-edge step capacity of validate_edges
    dense: bool = False                   # disable data-dependent early exits (roofline measurements)
    device_gate_seed: Optional[int] = None   # not None: validate_edges applies the gate on the device (counter-based
                                             # hash instead of np.random; first_hit / valid then include the 5 % pass-through)


@dataclass
class EdgeBatchResult:
    first_hit: "torch.Tensor"        # int32 [E]
    n_steps: "torch.Tensor"          # int32 [E]
    max_deflection: "torch.Tensor"   # float32 [E, W]
    over_limit: "torch.Tensor"       # float32 [E, W]
    ee_len: "torch.Tensor"           # float32 [E]
    cost: "torch.Tensor"             # float32 [E]
    rigid_mask: Optional["torch.Tensor"] = None    # uint32 as int32 [E, W, C]
    exceed_mask: Optional["torch.Tensor"] = None

    @property
    def valid(self):
        """True where the whole edge is free (success of extend_towards, primitives.py:254)."""
        return self.first_hit >= self.n_steps


@dataclass
class ConfigBatchResult:
    flags: "torch.Tensor"            # uint8 [B, W]  FLAG_RIGID | FLAG_EXCEEDED
    deflection: "torch.Tensor"       # float32 [B, W, P]
    theta: "torch.Tensor"            # float32 [B, W, P, 2]
    ee: "torch.Tensor"               # float32 [B, 3]


class EdgeChecker:
    def __init__(self, robot: RobotModel, worlds: Sequence[PlantWorld],
                 config: Optional[EdgeCheckConfig] = None,
                 obstacles: Optional[Sequence[FixedPrim]] = None, device: int = 0):
        self._lib = _capi.load()
        self._ctx = C.c_void_p()
        self.device_index = int(device)
        rc = self._lib.pmp_create(C.byref(self._ctx), self.device_index)
        if rc != 0:
            msg = self._lib.pmp_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise RuntimeError(f"pmp_create failed ({rc}): {msg}")
        self.robot = robot
        self.obstacles = obstacles
        self.config = config or EdgeCheckConfig()
        self._set_robot(robot, obstacles)
        self.set_worlds(worlds)
        self._push_params()

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._lib.pmp_last_error(self._ctx).decode()}")

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.pmp_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _set_robot(self, robot: RobotModel, obstacles) -> None:
        t: RobotTables = pack_robot(robot, obstacles)
        self.robot_tables = t
        d = _capi.robot_desc(t)
        self._check(self._lib.pmp_set_robot(self._ctx, C.byref(d)), "pmp_set_robot")
        self._set_hulls()

    def _set_hulls(self) -> None:
        hulls = getattr(self.robot, "hulls", None)
        if hulls is not None and self.config.use_hulls:
            from .model.tables import default_obstacles
            n_obs = len(default_obstacles() if self.obstacles is None else list(self.obstacles))
            hd, keep = _capi.hull_desc(hulls, n_obs)
            self._check(self._lib.pmp_set_robot_hulls(self._ctx, C.byref(hd)), "pmp_set_robot_hulls")
            del keep
        else:
            self._check(self._lib.pmp_set_robot_hulls(self._ctx, None), "pmp_set_robot_hulls")

    def set_worlds(self, worlds) -> None:
        """``worlds``: PlantWorld objects, or ready-made WorldTables (model/batch_sampler.sample_world_tables for
        large world counts)."""
        if isinstance(worlds, WorldTables):
            self.worlds = None
            wt = worlds
        else:
            self.worlds = list(worlds)
            wt = pack_worlds(self.worlds, self.config.stem_margin, self.config.gravity_sag)
        self.world_tables = wt
        d = _capi.WorldDesc()
        d.n_worlds, d.max_stems, d.max_boxes, d.n_slots = wt.n_worlds, wt.max_stems, wt.max_boxes, wt.n_slots
        stems = np.ascontiguousarray(wt.stems.reshape(-1))
        boxes = np.ascontiguousarray(wt.boxes.reshape(-1))
        d.n_stems, d.stems = _capi.iptr(wt.n_stems), _capi.fptr(stems)
        d.n_boxes, d.boxes = _capi.iptr(wt.n_boxes), _capi.fptr(boxes)
        self._check(self._lib.pmp_set_worlds(self._ctx, C.byref(d)), "pmp_set_worlds")

    def sample_worlds_on_device(self, n: int, branches: int = 1, vert_stems: int = 2, extensions: int = 1,
                                plant_pos_xy_limits=((0.0, 0.35), (-0.5, 0.0)), seed: int = 0,
                                stem_base_spacing: float = 0.0, uniforms=None) -> "DeviceWorldTables":
        """Draw ``n`` plant worlds of one topology ON THE DEVICE (pmp_sample_worlds) and make them this checker's
        worlds (pmp_set_worlds_dev): the tables never exist on the host.  ``uniforms``: optional [n, U] float64 array /
        tensor of draws in [0, 1) (model/batch_sampler.uniforms_per_world; parity tests), else the counter-based
        generator keyed by ``seed``.  Same plants as model/batch_sampler.draws_from_uniforms on the same draws."""
        torch = self._torch()
        from .model import batch_sampler as bs
        from .model.plants import PLANT_JOINT_LIMIT, SCALING, STEM_COM_Z, STEM_HALF_WIDTH
        from .model.statics import GRAVITY, SPRING_KP, STEM_MASS
        topo = bs.sampler_topology(branches, vert_stems, extensions)
        P = topo["n_stems"]
        d = _capi.SamplerDesc()
        d.n_stems, d.n_slots, d.gravity_sag = P, topo["n_slots"], 1 if self.config.gravity_sag else 0
        for key in ("parent", "kind", "ref", "flags", "pslot", "oslot"):
            arr = getattr(d, key)
            for k, v in enumerate(topo[key]):
                arr[k] = int(v)
        (x0, x1), (y0, y1) = plant_pos_xy_limits
        d.xy_lo[0], d.xy_lo[1], d.xy_hi[0], d.xy_hi[1] = x0, y0, x1, y1
        d.spacing, d.scaling, d.half_width, d.com_z = stem_base_spacing, SCALING, STEM_HALF_WIDTH, STEM_COM_Z
        d.limit, d.stem_margin, d.kp, d.mass, d.g = PLANT_JOINT_LIMIT, self.config.stem_margin, SPRING_KP, STEM_MASS, GRAVITY
        dev = self._dev()
        with torch.cuda.device(dev):
            stems = torch.empty((n, P, 48), device=dev, dtype=torch.float32)
            boxes = torch.empty((n, 1, 16), device=dev, dtype=torch.float32)
            u_t, nu = None, 0
            if uniforms is not None:
                u_t = torch.as_tensor(uniforms).to(device=dev, dtype=torch.float64).contiguous()
                if u_t.dim() != 2 or u_t.shape[0] != n:
                    raise ValueError("uniforms must be [n, U]")
                nu = int(u_t.shape[1])
            self._check(self._lib.pmp_sample_worlds(self._ctx, C.byref(d), int(n), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                    self._p(u_t), nu, self._p(stems), self._p(boxes), self._stream()),
                        "pmp_sample_worlds")
            self._check(self._lib.pmp_set_worlds_dev(self._ctx, int(n), P, 1, topo["n_slots"], self._p(stems),
                                                     self._p(boxes), self._stream()), "pmp_set_worlds_dev")
        self.worlds = None
        self.world_tables = DeviceWorldTables(int(n), P, 1, topo["n_slots"], stems, boxes)
        return self.world_tables

    def _push_params(self) -> None:
        c = self.config
        if c.mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        p = _capi.Params(c.resolution, c.deflection_limit, c.alpha, c.max_step, c.reg_eps, c.iterations,
                         MODES[c.mode], 1 if c.dense else 0,
                         0 if c.device_gate_seed is None else int(c.device_gate_seed) & 0xFFFFFFFF,
                         0.0 if c.device_gate_seed is None else float(c.gate_probability),
                         float(c.arm_lag_gain), int(c.arm_lag_substeps))
        self._check(self._lib.pmp_set_params(self._ctx, C.byref(p)), "pmp_set_params")

    def set_config(self, **changes) -> None:
        """Replace constants (e.g. mode='avoid_all', deflection_limit=2.0)."""
        old = self.config
        self.config = replace(self.config, **changes)
        if (self.config.stem_margin, self.config.gravity_sag) != (old.stem_margin, old.gravity_sag):
            if self.worlds is None:
                raise ValueError("worlds were given as tables: re-pack them with the new stem_margin / gravity_sag and "
                                 "call set_worlds")
            self.set_worlds(self.worlds)
        if self.config.use_hulls != old.use_hulls:
            self._set_hulls()
        self._push_params()

    # ------------------------------------------------------------------ properties
    @property
    def dof(self) -> int:
        return self.robot.dof

    @property
    def num_worlds(self) -> int:
        return self.world_tables.n_worlds

    @property
    def max_stems(self) -> int:
        return self.world_tables.max_stems

    @property
    def num_links(self) -> int:
        return self.robot.num_links

    def launch_info(self) -> dict:
        info = _capi.LaunchInfo()
        self._check(self._lib.pmp_get_launch_info(self._ctx, C.byref(info)), "pmp_get_launch_info")
        return {k: int(getattr(info, k)) for k, _ in info._fields_}

    def set_kernel_timing(self, on: bool) -> None:
        """Record CUDA events around the kernels of every validate_edges call (bench.py's roofline leg)."""
        self._check(self._lib.pmp_set_kernel_timing(self._ctx, 1 if on else 0), "pmp_set_kernel_timing")

    def kernel_times(self):
        """(hull pre-pass ms, edge kernel ms) of the last timed validate_edges call; waits for it."""
        r, e = C.c_float(), C.c_float()
        self._check(self._lib.pmp_get_kernel_times(self._ctx, C.byref(r), C.byref(e)), "pmp_get_kernel_times")
        return r.value, e.value

    def measure_fp32_peak(self, repeats: int = 5):
        tf, ms = C.c_double(), C.c_double()
        self._check(self._lib.pmp_measure_fp32_peak(self._ctx, repeats, C.byref(tf), C.byref(ms)),
                    "pmp_measure_fp32_peak")
        return tf.value, ms.value

    # ------------------------------------------------------------------ device API (torch)
    def _torch(self):
        import torch
        return torch

    def _dev(self):
        return self._torch().device("cuda", self.device_index)

    def _prep(self, q, name: str):
        torch = self._torch()
        if not isinstance(q, torch.Tensor):
            q = torch.as_tensor(np.asarray(q, dtype=np.float32))
        if q.dim() == 1:
            q = q[None]
        if q.dim() != 2 or q.shape[1] != self.dof:
            raise ValueError(f"{name} must be [N, {self.dof}]")
        return q.to(device=self._dev(), dtype=torch.float32).contiguous()

    def _stream(self):
        return C.c_void_p(self._torch().cuda.current_stream(self._dev()).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    def fk(self, q):
        """Link frames for a batch of configurations: (pos [B,L,3], quat xyzw [B,L,4], com [B,L,3])."""
        torch = self._torch()
        q = self._prep(q, "q")
        B, L = q.shape[0], self.num_links
        with torch.cuda.device(self._dev()):
            pos = torch.empty((B, L, 3), device=q.device, dtype=torch.float32)
            quat = torch.empty((B, L, 4), device=q.device, dtype=torch.float32)
            com = torch.empty((B, L, 3), device=q.device, dtype=torch.float32)
            self._check(self._lib.pmp_fk(self._ctx, self._p(q), B, self._p(pos), self._p(quat), self._p(com),
                                         self._stream()), "pmp_fk")
        return pos, quat, com

    def check_configs(self, q, detail: bool = True) -> ConfigBatchResult:
        torch = self._torch()
        q = self._prep(q, "q")
        B, W, P = q.shape[0], self.num_worlds, self.max_stems
        with torch.cuda.device(self._dev()):
            flags = torch.empty((B, W), device=q.device, dtype=torch.uint8)
            defl = torch.empty((B, W, P), device=q.device, dtype=torch.float32) if detail else None
            theta = torch.empty((B, W, P, 2), device=q.device, dtype=torch.float32) if detail else None
            ee = torch.empty((B, 3), device=q.device, dtype=torch.float32)
            self._check(self._lib.pmp_check_configs(self._ctx, self._p(q), B, self._p(flags), self._p(defl),
                                                    self._p(theta), self._p(ee), self._stream()),
                        "pmp_check_configs")
        return ConfigBatchResult(flags, defl, theta, ee)

    def validate_edges(self, q_from, q_to, backward: bool = False, want_masks: bool = False,
                       max_steps: Optional[int] = None, out: Optional[EdgeBatchResult] = None,
                       first_hit_only: bool = False, edge_base: int = 0) -> EdgeBatchResult:
        """``edge_base``: index of edge 0 of this call in the caller's whole batch (enters only the on-device gate's
        hash, so shards / chunks of a batch decide exactly like the unsplit batch).
        ``first_hit_only``: request only first_hit / n_steps.  The kernel then stops like the reference's loops
        (steps after the first violating one and the remaining worlds are not evaluated) -- the planner's
        accept/reject question, several times cheaper on edges that collide early."""
        torch = self._torch()
        q0 = self._prep(q_from, "q_from")
        q1 = self._prep(q_to, "q_to")
        if q0.shape != q1.shape:
            raise ValueError("q_from and q_to must have the same shape")
        E, W = q0.shape[0], self.num_worlds
        ms = int(max_steps or self.config.max_steps)
        Cw = (ms + 31) // 32
        with torch.cuda.device(self._dev()):
            mk = lambda shape, dt: torch.empty(shape, device=q0.device, dtype=dt)
            if out is None and first_hit_only:
                out = EdgeBatchResult(mk((E,), torch.int32), mk((E,), torch.int32), None, None, None, None)
            elif out is None:
                out = EdgeBatchResult(mk((E,), torch.int32), mk((E,), torch.int32), mk((E, W), torch.float32),
                                      mk((E, W), torch.float32), mk((E,), torch.float32), mk((E,), torch.float32))
                if want_masks:
                    out.rigid_mask = mk((E, W, Cw), torch.int32)
                    out.exceed_mask = mk((E, W, Cw), torch.int32)
            self._check(self._lib.pmp_validate_edges(
                self._ctx, self._p(q0), self._p(q1), E, ms, 1 if backward else 0, self._p(out.first_hit),
                self._p(out.n_steps), self._p(out.max_deflection), self._p(out.over_limit), self._p(out.ee_len),
                self._p(out.cost), self._p(out.rigid_mask), self._p(out.exceed_mask), self._stream(), int(edge_base)),
                "pmp_validate_edges")
        return out

    # ------------------------------------------------------------------ batched tree extension (device)
    def nearest(self, nodes, targets, size=None, want_nodes: bool = False):
        """Nearest tree node of each target (pmp_nearest): (index int32 [B], distance float64 [B][, q_near [B, D]]).
        ``nodes`` [T, D] / ``targets`` [B, D] torch CUDA tensors; ``size``: optional device int32 [1] live count."""
        torch = self._torch()
        nodes = self._prep(nodes, "nodes")
        targets = self._prep(targets, "targets")
        T, B = nodes.shape[0], targets.shape[0]
        with torch.cuda.device(self._dev()):
            index = torch.empty((B,), device=nodes.device, dtype=torch.int32)
            dist = torch.empty((B,), device=nodes.device, dtype=torch.float64)
            q_near = torch.empty((B, self.dof), device=nodes.device, dtype=torch.float32) if want_nodes else None
            self._check(self._lib.pmp_nearest(self._ctx, self._p(nodes), T, self._p(size), self._p(targets), B,
                                              self._p(index), self._p(dist), self._p(q_near), self._stream()),
                        "pmp_nearest")
        return (index, dist, q_near) if want_nodes else (index, dist)

    def edge_points(self, q_from, q_to, steps, backward: bool = False):
        """Configuration number steps[e] of every edge, with the arithmetic of the edge kernel (pmp_edge_points)."""
        torch = self._torch()
        q0 = self._prep(q_from, "q_from")
        q1 = self._prep(q_to, "q_to")
        st = torch.as_tensor(steps).to(device=q0.device, dtype=torch.int32).contiguous()
        if q0.shape != q1.shape or st.shape != (q0.shape[0],):
            raise ValueError("q_from, q_to [E, D] and steps [E] must agree")
        with torch.cuda.device(self._dev()):
            out = torch.empty_like(q0)
            self._check(self._lib.pmp_edge_points(self._ctx, self._p(q0), self._p(q1), q0.shape[0], self._p(st),
                                                  1 if backward else 0, self._p(out), self._stream()),
                        "pmp_edge_points")
        return out

    def new_tree(self, root, capacity: int = 1 << 16) -> "DeviceTree":
        return DeviceTree(self, root, capacity)

    def tree_extend(self, tree: "DeviceTree", targets, max_steps: Optional[int] = None) -> "TreeExtension":
        """One batched extension of ``tree`` towards ``targets`` [B, D] (pmp_tree_extend): nearest node, multi-world
        edge check, append of the last safe configuration -- five kernels, no host synchronisation."""
        torch = self._torch()
        targets = self._prep(targets, "targets")
        B = targets.shape[0]
        ms = int(max_steps or self.config.max_steps)
        with torch.cuda.device(self._dev()):
            mk = lambda shape, dt: torch.empty(shape, device=targets.device, dtype=dt)
            ext = TreeExtension(mk((B,), torch.int32), mk((B,), torch.int32), mk((B, self.dof), torch.float32),
                                mk((4,), torch.int32), targets)
            self._check(self._lib.pmp_tree_extend(
                self._ctx, C.byref(tree._desc), min(tree.size_bound, tree.capacity), self._p(targets), B, ms,
                self._p(ext.new_index), self._p(ext.reached), self._p(ext.new_config), self._p(ext.stats),
                self._stream()), "pmp_tree_extend")
        tree.size_bound = min(tree.size_bound + B, tree.capacity)
        return ext

    # ------------------------------------------------------------------ host API (numpy; the C-ABI *_host calls)
    def validate_edges_host(self, q_from: np.ndarray, q_to: np.ndarray, backward: bool = False,
                            want_masks: bool = False, max_steps: Optional[int] = None, out: Optional[dict] = None,
                            first_hit_only: bool = False) -> dict:
        """HOST buffers in, HOST buffers out: H2D, kernel, D2H and the synchronisation happen inside the
        C-ABI call -- the entry a CPU-side planner uses."""
        q0 = np.ascontiguousarray(q_from, dtype=np.float32).reshape(-1, self.dof)
        q1 = np.ascontiguousarray(q_to, dtype=np.float32).reshape(-1, self.dof)
        E, W = q0.shape[0], self.num_worlds
        ms = int(max_steps or self.config.max_steps)
        Cw = (ms + 31) // 32
        if out is None and first_hit_only:
            out = {"first_hit": np.empty(E, np.int32), "n_steps": np.empty(E, np.int32)}
        elif out is None:
            out = {
                "first_hit": np.empty(E, np.int32), "n_steps": np.empty(E, np.int32),
                "max_deflection": np.empty((E, W), np.float32), "over_limit": np.empty((E, W), np.float32),
                "ee_len": np.empty(E, np.float32), "cost": np.empty(E, np.float32),
            }
            if want_masks:
                out["rigid_mask"] = np.empty((E, W, Cw), np.uint32)
                out["exceed_mask"] = np.empty((E, W, Cw), np.uint32)
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p()
        self._check(self._lib.pmp_validate_edges_host(
            self._ctx, vp(q0), vp(q1), E, ms, 1 if backward else 0, vp(out["first_hit"]), vp(out["n_steps"]),
            vp(out.get("max_deflection")), vp(out.get("over_limit")), vp(out.get("ee_len")), vp(out.get("cost")),
            vp(out.get("rigid_mask")), vp(out.get("exceed_mask"))), "pmp_validate_edges_host")
        return out

    def check_configs_host(self, q: np.ndarray, detail: bool = True) -> dict:
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dof)
        B, W, P = q.shape[0], self.num_worlds, self.max_stems
        out = {"flags": np.empty((B, W), np.uint8), "ee": np.empty((B, 3), np.float32)}
        if detail:
            out["deflection"] = np.empty((B, W, P), np.float32)
            out["theta"] = np.empty((B, W, P, 2), np.float32)
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p()
        self._check(self._lib.pmp_check_configs_host(self._ctx, vp(q), B, vp(out["flags"]), vp(out.get("deflection")),
                                                     vp(out.get("theta")), vp(out["ee"])), "pmp_check_configs_host")
        return out


class DeviceWorldTables:
    """World tables that live in device memory (EdgeChecker.sample_worlds_on_device).  Same attribute names as
    model.tables.WorldTables; ``stems`` / ``boxes`` are downloaded on first use (tests, reports)."""

    def __init__(self, n_worlds, max_stems, max_boxes, n_slots, stems_dev, boxes_dev):
        self.n_worlds, self.max_stems, self.max_boxes, self.n_slots = n_worlds, max_stems, max_boxes, n_slots
        self.stems_dev, self.boxes_dev = stems_dev, boxes_dev
        self._stems = self._boxes = None

    @property
    def stems(self) -> np.ndarray:
        if self._stems is None:
            self._stems = self.stems_dev.cpu().numpy()
        return self._stems

    @property
    def boxes(self) -> np.ndarray:
        if self._boxes is None:
            self._boxes = self.boxes_dev.cpu().numpy()
        return self._boxes

    @property
    def n_stems(self) -> np.ndarray:
        return np.full(self.n_worlds, self.max_stems, dtype=np.int32)

    @property
    def n_boxes(self) -> np.ndarray:
        return np.full(self.n_worlds, self.max_boxes, dtype=np.int32)


@dataclass
class TreeExtension:
    """Device outputs of one pmp_tree_extend call."""
    new_index: object      # int32 [B] slot of the appended node, -1 if the extension made no progress
    reached: object        # int32 [B] 1 if the target itself was reached
    new_config: object     # float32 [B, D] appended configuration, NaN where none
    stats: object          # int32 [4] appended, reached, first reached batch index (B if none), capacity overflow
    targets: object        # float32 [B, D] the targets of the call


class DeviceTree:
    """A search tree resident in HBM (pmp_tree): configurations, parent links and, per node, the edge it lies on."""

    def __init__(self, checker: EdgeChecker, root, capacity: int):
        import torch
        self.checker = checker
        self.capacity = int(capacity)
        dev = checker._dev()
        D = checker.dof
        # only rows below `size` are ever read (the append kernel writes every field of a new node): no fills
        self.nodes = torch.empty((self.capacity, D), device=dev, dtype=torch.float32)
        self.parent = torch.empty((self.capacity,), device=dev, dtype=torch.int32)
        self.via = torch.empty((self.capacity, D), device=dev, dtype=torch.float32)
        self.via_steps = torch.empty((self.capacity,), device=dev, dtype=torch.int32)
        self.size = torch.ones((1,), device=dev, dtype=torch.int32)
        root_t = torch.as_tensor(np.asarray(root, dtype=np.float32).reshape(D)).to(dev)
        self.nodes[0] = root_t
        self.via[0] = root_t
        self.parent[:1] = -1
        self.via_steps[:1] = 0
        self.size_bound = 1            # host upper bound of the live node count (no synchronisation needed)
        self._desc = _capi.Tree(self.nodes.data_ptr(), self.parent.data_ptr(), self.via.data_ptr(),
                                self.via_steps.data_ptr(), self.size.data_ptr(), self.capacity, 0)

    def download(self) -> dict:
        """Host copy of the live part of the tree (synchronises)."""
        n = int(self.size.item())
        self.size_bound = n
        return {"nodes": self.nodes[:n].cpu().numpy(), "parent": self.parent[:n].cpu().numpy(),
                "via": self.via[:n].cpu().numpy(), "via_steps": self.via_steps[:n].cpu().numpy()}
