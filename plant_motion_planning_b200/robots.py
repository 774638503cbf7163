"""Bundled robot tables (compiled from the reference tree by tools/compile_iiwa14.py; the reference tree is not
available at run time, so the numbers travel with the package).

``iiwa14_hull_spheres.json`` + ``iiwa14_collision_hulls.npz`` -- the arm every iiwa script of the reference loads,
    DRAKE_IIWA_URDF_EDIT (pybullet_tools/utils.py:78; models/drake/iiwa_description/urdf/
    iiwa14_polytope_collision_edit.urdf): 7 revolute joints, collision geometry = convex hulls of the link meshes (what
    PyBullet builds from them) for the exact self / obstacle test, 18 spheres fitted to those hulls (+-20 mm) for the
    plant contact model and the broad phase; end-effector link ``iiwa_link_ee`` = PyBullet link 9.
``iiwa14_spheres.json`` -- the URDF family's own primitive model iiwa14_spheres_collision.urdf (12 spheres on links
    1-7, cylinder on link 0), decided on the spheres alone.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from .model.robot import RobotModel
from .model.transforms import Frame

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_iiwa14(base: Optional[Frame] = None, collision: str = "hulls") -> RobotModel:
    """iiwa14, base at the world origin unless ``base`` is given (scripts/our_method_multi_branch.py:107 loads it at the
    origin).  ``collision``: "hulls" (default: the meshes' convex hulls + fitted spheres) or "urdf_spheres"."""
    if collision == "hulls":
        with np.load(os.path.join(_DATA, "iiwa14_collision_hulls.npz")) as z:
            verts = {k: z[k].astype(np.float64) for k in z.files}
        with open(os.path.join(_DATA, "iiwa14_hull_spheres.json")) as fp:
            model = RobotModel.from_json(fp.read(), hull_verts=verts)
    elif collision == "urdf_spheres":
        with open(os.path.join(_DATA, "iiwa14_spheres.json")) as fp:
            model = RobotModel.from_json(fp.read())
    else:
        raise ValueError(f"unknown collision model {collision!r}")
    return model if base is None else model.with_base(base)
