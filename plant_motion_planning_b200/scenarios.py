"""The reference scripts' environments (seeds, limits, start/goal), as data.   SURVEY 8d configs 1-3.

  multi_world_our_method ... scripts/multi_world_our_method.py:74-105  seed 1, limits ((0,0),(0.35,0.35)), (1,2,1),
                             deflection limit 2.0, one world at offset (0, 0)
  multi_world_avoid_all .... scripts/multi_world_avoid_all.py:73-92    seed 19, limits ((0.2,0.35),(-0.5,0)), limit 0.30,
                             offsets (0,5) x (0,5), avoid_all=True
  multi_world_ignore_all ... scripts/multi_world_ignore_all.py:66-96   seed 1, limits ((0,0.35),(-0.5,0)), limit 0.30
  our_method_multi_branch .. scripts/our_method_multi_branch.py:82-133 saved plant environments/multi_branch/envK, limit 0.30
  our_method_single_branch . scripts/our_method_single_branch.py + plant_motion_planning/utils.py:120-230 (env1..env8)

Robot: the vendored iiwa14 (the scripts that name it load it at the world origin).  Start = zeros
(scripts/our_method_multi_branch.py:82).  Goal: the script's goal (:84-85) puts the flange 2 cm above the block,
which the link-7 *bounding sphere* of the primitive model overlaps, so joint 4 is backed off by 0.05 rad
(end effector moves 19 mm) -- the nearest goal that is free in the sphere model.
"""
from __future__ import annotations

import random
from typing import List, Tuple

import numpy as np

from .model.plants import SINGLE_STEM_ENVS, PlantWorld, sample_world, single_stem_world

START = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
REFERENCE_GOAL = (-1.3871757013351371, 1.6063773991870438, 2.152853076950719, -1.0638334445672613,
                  -0.1398096715085235, 1.9391079682303638, -1.051507203470595)
GOAL = REFERENCE_GOAL[:3] + (REFERENCE_GOAL[3] + 0.05,) + REFERENCE_GOAL[4:]

SCRIPTS = {
    "multi_world_our_method": dict(seed=1, xs=(0,), ys=(0,), limits=((0, 0), (0.35, 0.35)), args=(1, 2, 1),
                                   deflection_limit=2.0, avoid_all=False, mode="our_method"),
    "multi_world_avoid_all": dict(seed=19, xs=(0, 5), ys=(0, 5), limits=((0.2, 0.35), (-0.5, 0.0)), args=(1, 2, 1),
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
