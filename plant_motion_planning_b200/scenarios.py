"""The reference scripts' environments (seeds, limits, start/goal), as data.   SURVEY 8d configs 1-3.

  multi_world_our_method ... scripts/multi_world_our_method.py:74-105  seed 1, limits ((0,0),(0.35,0.35)), (1,2,1),
                             deflection limit 2.0, one world at offset (0, 0)
  multi_world_avoid_all .... scripts/multi_world_avoid_all.py:73-92    seed 19, limits ((0.2,0.35),(-0.5,0)), limit 0.30,
                             offsets (0,5) x (0,5), avoid_all=True
  multi_world_ignore_all ... scripts/multi_world_ignore_all.py:66-96   seed 1, limits ((0,0.35),(-0.5,0)), limit 0.30
  our_method_multi_branch .. scripts/our_method_multi_branch.py:82-133 saved plant environments/multi_branch/envK, limit 0.30
  our_method_single_branch . scripts/our_method_single_branch.py + plant_motion_planning/utils.py:120-230 (env1..env8)

Robot: the vendored iiwa14 as the scripts load it (DRAKE_IIWA_URDF_EDIT, at the world origin).  Start = zeros
(scripts/our_method_multi_branch.py:82).  Goal: the script's own goal (:84-85), unmodified: it parks the flange
face 1.4 mm from the block (0.39 mm of Bullet clearance after both 1 mm margins) and leaves 0.52 mm between the
hulls of links 5 and 7 -- free on the convex hulls the reference tests, which is what the kernels decide on.
"""
from __future__ import annotations

import random
from typing import List, Tuple

import numpy as np

from .model.plants import SINGLE_STEM_ENVS, PlantWorld, sample_world, single_stem_world

START = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
REFERENCE_GOAL = (-1.3871757013351371, 1.6063773991870438, 2.152853076950719, -1.0638334445672613,
                  -0.1398096715085235, 1.9391079682303638, -1.051507203470595)
GOAL = REFERENCE_GOAL

SCRIPTS = {
    "multi_world_our_method": dict(seed=1, xs=(0,), ys=(0,), limits=((0, 0), (0.35, 0.35)), args=(1, 2, 1),
                                   deflection_limit=2.0, avoid_all=False, mode="our_method"),
    "multi_world_avoid_all": dict(seed=19, xs=(0, 5), ys=(0, 5), limits=((0.2, 0.35), (-0.5, 0.0)), args=(1, 2, 1),
                                  deflection_limit=0.30, avoid_all=True, mode="avoid_all"),
    # The script's seed-19 worlds put a branch of the fourth plant 1.3 mm from link 4 of the iiwa at its zero
    # configuration (0.7 mm inside the two 1 mm Bullet margins: oracle/hull_oracle.py on the hulls and the stem boxes),
    # so with avoid_all the start itself collides and the reference would exit() in check_initial_end_multi_world
    # (pybullet_tools/utils.py:4364-4380).  (The script drives the non-vendored Val robot, not the iiwa.)  The same
    # script with seed 37 leaves start and goal free in all four worlds and is the avoid_all PLANNING scenario.
    "multi_world_avoid_all_seed37": dict(seed=37, xs=(0, 5), ys=(0, 5), limits=((0.2, 0.35), (-0.5, 0.0)), args=(1, 2, 1),
                                         deflection_limit=0.30, avoid_all=True, mode="avoid_all"),
    "multi_world_ignore_all": dict(seed=1, xs=(0, 5), ys=(0, 5), limits=((0, 0.35), (-0.5, 0.0)), args=(1, 2, 1),
                                   deflection_limit=0.30, avoid_all=False, mode="ignore_all"),
}


def script_worlds(name: str) -> Tuple[List[PlantWorld], dict]:
    """Worlds of a multi_world_* script, drawn exactly as MultiPlantWorld.__init__ does (seeds the GLOBAL streams,
    like the script)."""
    cfg = SCRIPTS[name]
    np.random.seed(cfg["seed"])
    random.seed(cfg["seed"])
    worlds = [sample_world(*cfg["args"], cfg["limits"], base_offset_xy=(x, y)) for x in cfg["xs"] for y in cfg["ys"]]
    return worlds, cfg


def single_branch_world(env_name: str = "env1") -> PlantWorld:
    return single_stem_world(SINGLE_STEM_ENVS[env_name])
