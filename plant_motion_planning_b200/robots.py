"""Bundled robot tables.

``data/iiwa14_spheres.json`` is the compiled form of the reference's vendored primitive model
models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf (7 revolute joints, 12 spheres on
links 1-7, cylinder on link 0, end-effector link ``iiwa_link_ee`` = PyBullet link 9), produced by
tools/compile_iiwa14.py.  The reference tree is not available at run time, so the numbers travel
with the package.
"""
from __future__ import annotations

import os
from typing import Optional

from .model.robot import RobotModel
from .model.transforms import Frame

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_iiwa14(base: Optional[Frame] = None) -> RobotModel:
    """iiwa14 with its sphere collision model, base at the world origin unless ``base`` is given
    (scripts/our_method_multi_branch.py:107 loads it at the origin)."""
    with open(os.path.join(_DATA, "iiwa14_spheres.json")) as fp:
        model = RobotModel.from_json(fp.read())
    return model if base is None else model.with_base(base)
