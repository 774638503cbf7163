"""MultiPlantWorld / SinglePlantEnv facade with the reference's surface (plant_motion_planning/env.py:30-255),
backed by the GPU edge checker instead of a Bullet world.

What the reference's RRT fork calls and what it gets here:
  env.step(q, set_joint_pos=False)            evaluates q in ALL worlds with one kernel launch (the quasi-static
                                              stand-in for 50 x stepSimulation) and caches flags + deflections
  env.step_after_restoring(q), step_no_action same evaluation (the model is stateless)
  env.net_constraint_violation_checker_forward(q)    OR over worlds in construction order with early exit;
                                              per world: limits(q) | rigid | (exceeded & rand() < 0.95), drawing
                                              np.random.rand() exactly where the reference does (utils.py:3836-3855)
  ..._backward(q) / ..._benchmark(q)          rigid obstacles only (env.py:96-120, 241-255)
  env.envs[(x, y)].plant_rep.observe_all() / .deflections / .observe_all_and_check(limit)
  save_state() / restore_state(id)            snapshot / restore of the one piece of state the model has: the arm
                                              configuration last stepped to (pybullet_tools/utils.py:1517-1523)
Additive: env.checker (EdgeChecker), env.validate_edges(...) for batched use.

Like the reference's collision closures, the checkers test the state left by the last ``step`` and use their
argument only for the joint-limit test (pybullet_tools/utils.py:3979-4014).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import weakref

import numpy as np

from .edgecheck import FLAG_RIGID, EdgeCheckConfig, EdgeChecker
from .model.plants import PlantWorld, load_plant, sample_world
from .model.robot import RobotModel
from .planner_fns import GATE_PROBABILITY, get_limits_fn
from .robots import load_iiwa14

# State handles (pybullet_tools/utils.py:1517-1523).  The reference's p.saveState() is a module-level call on the one
# global Bullet client; here a snapshot belongs to the MultiPlantWorld that was active when it was taken: it is stored
# IN that environment (and dies with it), and its id carries the environment's serial number, so two environments in
# one process never see each other's snapshots.
_ENVS: "weakref.WeakValueDictionary[int, MultiPlantWorld]" = weakref.WeakValueDictionary()
_ACTIVE: List[Optional["MultiPlantWorld"]] = [None]
_SERIAL = [0]
_ID_BITS = 32


def save_state() -> int:
    """pybullet_tools/utils.py:1517-1519 (p.saveState).  The quasi-static world has one piece of state -- the arm
    configuration it was last stepped to -- so a snapshot is that configuration of the active MultiPlantWorld."""
    env = _ACTIVE[0]
    if env is None:
        return 0
    return env.save_state()


def restore_state(state_id: int) -> None:
    """pybullet_tools/utils.py:1521-1523 (p.restoreState)."""
    env = _ENVS.get(int(state_id) >> _ID_BITS)
    if env is not None:
        env.restore_state(state_id)


class PlantRepresentation:
    """CharacterizePlant / CharacterizePlant2 view of one world (representation.py:248-418)."""

    def __init__(self, owner: "MultiPlantWorld", world_index: int, n_stems: int):
        self._owner = owner
        self._w = world_index
        self._n = n_stems
        self.deflections: List[float] = []

    def observe_all(self) -> None:
        d = self._owner._deflections()
        self.deflections.clear()
        self.deflections.extend(float(v) for v in d[self._w, :self._n])

    def observe_all_and_check(self, deflection_limit: float) -> bool:
        self.observe_all()
        return any(d > deflection_limit for d in self.deflections)


class SinglePlantEnv:
    """One sampled world.  Constructed by MultiPlantWorld (which owns the shared GPU context)."""

    def __init__(self, owner: "MultiPlantWorld", world_index: int, world: PlantWorld, base_offset_xy, avoid_all: bool):
        self._owner = owner
        self._w = world_index
        self.world = world
        self.base_offset_xy = tuple(base_offset_xy)
        self.robot = owner.sample_robot
        self.joints = list(range(owner.sample_robot.dof))
        self.plant_rep = PlantRepresentation(owner, world_index, world.num_stems)
        self.plant_id = ("plant", world_index)
        self.base_id = self.plant_id
        self.joint_list = list(world.joint_list)
        self.fixed = ["floor", "block"] + ([self.plant_id] if avoid_all else [])
        self.movable = [self.plant_rep]
        self.deflection_limit = owner.deflection_limit
        self.avoid_all = avoid_all

    # -- the three closures of env.py:96-120 ---------------------------------------------------
    def collision_fn(self, q, verbose: bool = False) -> bool:
        """get_collision_fn_with_angle_constraints_v4 with movable=[plant_rep] (utils.py:3836-3855)."""
        if self._owner._rigid(self._w, q):
            return True
        for b in self.movable:
            if b.observe_all_and_check(self.deflection_limit) and np.random.rand() < GATE_PROBABILITY:
                return True
        return False

    def collision_fn_back(self, q, verbose: bool = False) -> bool:
        """Same closure with movable=[] (env.py:104-110)."""
        return self._owner._rigid(self._w, q)

    def collision_fn_benchmark(self, q, verbose: bool = False) -> bool:
        """get_collision_fn_with_controls_benchmark_multiworld (utils.py:4019-4062)."""
        return self._owner._rigid(self._w, q)

    def constraint_violation_checker_forward(self, q):
        return self.collision_fn(q)

    def constraint_violation_checker_backward(self, q):
        return self.collision_fn_back(q)

    def constraint_violation_checker_benchmark(self, q):
        return self.collision_fn_benchmark(q)

    def step(self, action, set_joint_pos: bool = False):
        self._owner.step(action, set_joint_pos)

    def step_after_restoring(self, action, set_joint_pos: bool = True):
        self._owner._set_action(action)

    def state(self):
        return self.plant_rep.deflections


class MultiPlantWorld:
    """Multi-world environment (env.py:187-255): one world per (x, y) offset, x outer / y inner."""

    def __init__(self, xs: Sequence[float], ys: Sequence[float], deflection_limit: float,
                 num_branches_per_stem: Optional[int] = None, total_num_vert_stems: Optional[int] = None,
                 total_num_extensions: Optional[int] = None, plant_pos_xy_limits=((0, 1), (0, 1)),
                 loadPath: Optional[str] = None, physicsClientId=None, avoid_all: bool = False, *,
                 robot: Optional[RobotModel] = None, worlds: Optional[Sequence[PlantWorld]] = None,
                 config: Optional[EdgeCheckConfig] = None, device: int = 0, stem_base_spacing: float = 0.0,
                 checker_factory=None):
        self.sample_robot = robot if robot is not None else load_iiwa14()
        self.joints = list(range(self.sample_robot.dof))
        self.deflection_limit = deflection_limit
        self.avoid_all = avoid_all
        offsets = [(x, y) for x in xs for y in ys]
        if worlds is None:
            worlds = []
            for xy in offsets:
                if loadPath is None:
                    # same draws, same order as SinglePlantEnv.__init__ (env.py:63-70)
                    worlds.append(sample_world(num_branches_per_stem, total_num_vert_stems, total_num_extensions,
                                               plant_pos_xy_limits, base_offset_xy=xy,
                                               stem_base_spacing=stem_base_spacing))
                else:
                    worlds.append(load_plant(os.path.join(loadPath, "plant.urdf"),
                                             os.path.join(loadPath, "plant_params.pkl"), xy))
        elif len(worlds) != len(offsets):
            raise ValueError("need one PlantWorld per (x, y) offset")
        base = config or EdgeCheckConfig()
        from dataclasses import replace
        self.config = replace(base, deflection_limit=float(deflection_limit),
                              mode="avoid_all" if avoid_all else "our_method")
        # `checker_factory(robot, worlds, config)`: dependency injection for tests (returns anything with
        # check_configs_host / validate_edges*); the product always builds the CUDA EdgeChecker, which raises
        # without a B200.
        if checker_factory is not None:
            self.checker = checker_factory(self.sample_robot, list(worlds), self.config)
        else:
            self.checker = EdgeChecker(self.sample_robot, list(worlds), self.config, device=device)
        self.plant_worlds = list(worlds)
        self._limits_fn = get_limits_fn(self.sample_robot)
        self.envs: Dict[Tuple[float, float], SinglePlantEnv] = {}
        for i, (xy, w) in enumerate(zip(offsets, worlds)):
            self.envs[xy] = SinglePlantEnv(self, i, w, xy, avoid_all)
        self.fixed = self.envs[offsets[0]].fixed
        self._action: Optional[Tuple[float, ...]] = None
        _SERIAL[0] += 1
        self._serial = _SERIAL[0]
        self._snapshots: Dict[int, Optional[Tuple[float, ...]]] = {}
        _ENVS[self._serial] = self
        _ACTIVE[0] = self
        self._cache_q: Optional[Tuple[float, ...]] = None
        self._cache = None
        self.kernel_evaluations = 0

    # -- state ------------------------------------------------------------------------------------
    def save_state(self) -> int:
        """Snapshot of this environment (the configuration last stepped to); the id is valid for restore_state of this
        module and of this object.  Snapshots live until release_states() or the environment goes away."""
        k = len(self._snapshots) + 1
        self._snapshots[k] = self._action
        return (self._serial << _ID_BITS) | k

    def restore_state(self, state_id: int) -> None:
        if (int(state_id) >> _ID_BITS) != self._serial:
            raise ValueError("state id belongs to another environment")
        self._action = self._snapshots[int(state_id) & ((1 << _ID_BITS) - 1)]
        self._cache_q = None

    def release_states(self) -> None:
        """Forget every snapshot (the reference never frees its Bullet snapshots; a long planning session here can)."""
        self._snapshots.clear()

    def activate(self) -> None:
        """Make this the environment the module-level save_state() acts on (the last one constructed is by default)."""
        _ACTIVE[0] = self

    def _set_action(self, action) -> None:
        self._action = tuple(float(v) for v in action)

    def _evaluate(self):
        if self._action is None:
            raise RuntimeError("step(q) must be called before the checkers (as in the reference)")
        if self._cache_q != self._action:
            out = self.checker.check_configs_host(np.asarray(self._action, dtype=np.float32)[None], detail=True)
            self._cache = (out["flags"][0], out["deflection"][0])
            self._cache_q = self._action
            self.kernel_evaluations += 1
        return self._cache

    def _deflections(self) -> np.ndarray:
        return self._evaluate()[1]

    def _rigid(self, w: int, q) -> bool:
        if self._action is None:
            self._set_action(q)
        if self._limits_fn(q):
            return True
        return bool(self._evaluate()[0][w] & FLAG_RIGID)

    # -- reference surface ------------------------------------------------------------------------
    def step(self, action, set_joint_pos: bool = False) -> None:
        self._set_action(action)

    def step_after_restoring(self, action) -> None:
        self._set_action(action)

    def step_no_action(self) -> None:
        return None

    def net_constraint_violation_checker_forward(self, q) -> bool:
        for env in self.envs.values():
            if env.collision_fn(q):
                return True
        return False

    def net_constraint_violation_checker_backward(self, q) -> bool:
        for env in self.envs.values():
            if env.collision_fn_back(q):
                return True
        return False

    def net_constraint_violation_checker_benchmark(self, q) -> bool:
        for env in self.envs.values():
            if env.collision_fn_benchmark(q):
                return True
        return False

    # -- additive batched surface -----------------------------------------------------------------
    def validate_edges(self, q_from, q_to, **kw):
        return self.checker.validate_edges(q_from, q_to, **kw)

    def path_cost(self, init_conf, path: Sequence[Sequence[float]]):
        """Command.execute_multi_world's cost (kuka_primitives.py:467-521) for a list of waypoints visited by
        env.step: (ee_len, over[W], total[W]) with total = ee_len + alpha * over."""
        Q = np.asarray([init_conf] + [tuple(p) for p in path], dtype=np.float32)
        out = self.checker.check_configs_host(Q, detail=True)
        ee = out["ee"].astype(np.float64)
        ee_len = float(np.sum(np.linalg.norm(ee[1:] - ee[:-1], axis=1)))
        d = out["deflection"][1:].astype(np.float64)
        over = np.sum(np.maximum(d - self.deflection_limit, 0.0), axis=(0, 2))
        return ee_len, over, ee_len + self.config.alpha * over

    def execute_multi_world(self, init_conf, path: Sequence[Sequence[float]], verbose: bool = True, out=None) -> dict:
        """Command.execute_multi_world (pybullet_tools/kuka_primitives.py:467-521) as ONE batched evaluation: the
        whole path is checked in every world by a single launch and the report the reference prints step by step
        is produced from the returned traces.  Returns dict(ee_path_cost, deflection_over_limit [W], total_cost [W],
        alpha, deflection [steps, W, P] (per-world deflection traces), flags [steps, W])."""
        import sys
        out = sys.stdout if out is None else out
        Q = np.asarray([tuple(init_conf)] + [tuple(p) for p in path], dtype=np.float32)
        res = self.checker.check_configs_host(Q, detail=True)
        ee = res["ee"].astype(np.float64)
        defl = res["deflection"][1:].astype(np.float64)
        limit, alpha = float(self.deflection_limit), float(self.config.alpha)
        over = np.zeros(len(self.envs))
        ee_cost = 0.0
        if verbose:
            print("Executing path...", file=out)
        for i in range(defl.shape[0]):
            for w, env in enumerate(self.envs.values()):
                n = env.plant_rep._n
                if verbose:
                    print("Environment: ", w, file=out)
                for e in range(n):
                    d = float(defl[i, w, e])
                    if verbose:
                        print("\tDeflection of plant %d: %f" % (e, d), file=out)
                    if d > limit:
                        if verbose:
                            print("Error! Deflection limit exceeded!", file=out)
                        over[w] += d - limit
                if verbose:
                    print("========================================", file=out)
            ee_cost += float(np.linalg.norm(ee[i + 1] - ee[i]))
        total = ee_cost + alpha * over
        if verbose:
            bar = "===================================================="
            print(bar, file=out)
            print("total path cost: ", ee_cost, file=out)
            print("total Deflection over limit: ", over, file=out)
            print("alpha: ", alpha, file=out)
            print(bar, file=out)
            print("Total cost: ", total, file=out)
            print(bar, file=out)
        self._set_action(Q[-1])
        return {"ee_path_cost": ee_cost, "deflection_over_limit": over, "total_cost": total, "alpha": alpha,
                "deflection": defl, "flags": res["flags"][1:]}
