"""Model compiler: URDF / plant samplers -> kernel tables (host side, float64 until packed)."""
from .plants import (OBS_ANGLE, OBS_MOD3, PlantWorld, Stem, load_plant, plant_from_parsed, sample_plant,
                     sample_world, single_stem_world)
from .robot import FixedPrim, RobotModel, compile_robot, prune_always_colliding
from .tables import default_obstacles, pack_robot, pack_worlds
from .transforms import Frame

__all__ = [
    "OBS_ANGLE", "OBS_MOD3", "PlantWorld", "Stem", "load_plant", "plant_from_parsed", "sample_plant",
    "sample_world", "single_stem_world", "FixedPrim", "RobotModel", "compile_robot", "prune_always_colliding",
    "default_obstacles", "pack_robot", "pack_worlds", "Frame",
]
