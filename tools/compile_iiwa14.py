"""Regenerate plant_motion_planning_b200/data/iiwa14_spheres.json from the reference's URDF.

Run in the build container (needs /root/reference):  python tools/compile_iiwa14.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from plant_motion_planning_b200.model.robot import compile_robot, prune_always_colliding  # noqa: E402

REF = os.environ.get("PMP_REFERENCE", "/root/reference")
URDF = os.path.join(REF, "models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf")
OUT = os.path.join(os.path.dirname(__file__), "..", "plant_motion_planning_b200", "data", "iiwa14_spheres.json")

model = compile_robot(URDF, ee_link="iiwa_link_ee", name="iiwa14")
removed = prune_always_colliding(model)
print("always-colliding sphere pairs removed:", removed)
with open(OUT, "w") as fp:
    fp.write(model.to_json())
print("wrote", os.path.abspath(OUT))
