"""Regenerate the bundled iiwa14 tables from the reference tree.   Run in the build container (needs /root/reference).

  data/iiwa14_hull_spheres.json + data/iiwa14_collision_hulls.npz
      the arm the reference scripts load, DRAKE_IIWA_URDF_EDIT = iiwa14_polytope_collision_edit.urdf: kinematic chain
      and link table from the URDF, collision geometry = convex hulls of meshes/collision/link_{1..6}.obj and
      link_7_polytope.obj (tools/extract_iiwa_hulls.py) for the exact rigid narrow phase, plus the spheres fitted to
      them (tools/fit_arm_spheres.py -> data/iiwa14_hull_spheres_fit.json) for the plant contact model / broad phase.
  data/iiwa14_spheres.json
      the URDF's own 12-sphere primitive model (iiwa14_spheres_collision.urdf), kept as an alternative.

    python tools/extract_iiwa_hulls.py && python tools/fit_arm_spheres.py --budget 18 --kmax 3 && python tools/compile_iiwa14.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from plant_motion_planning_b200.model.robot import compile_robot, prune_always_colliding  # noqa: E402

REF = os.environ.get("PMP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "..", "plant_motion_planning_b200", "data")
URDF_DIR = os.path.join(REF, "models/drake/iiwa_description/urdf")

# ---- the URDF's own sphere model
model = compile_robot(os.path.join(URDF_DIR, "iiwa14_spheres_collision.urdf"), ee_link="iiwa_link_ee", name="iiwa14")
removed = prune_always_colliding(model)
print("iiwa14_spheres: always-colliding sphere pairs removed:", removed)
with open(os.path.join(DATA, "iiwa14_spheres.json"), "w") as fp:
    fp.write(model.to_json())

# ---- hulls + fitted spheres on the URDF the scripts load
fix = np.load(os.path.join(HERE, "..", "tests", "golden", "iiwa14_collision_hulls.npz"))
fit = json.load(open(os.path.join(DATA, "iiwa14_hull_spheres_fit.json")))
hull_names = [str(n) for n in fix["link_names"]]
mesh_hulls = {n: fix[f"verts_{n}"] for n in hull_names}
SLACK = 1.0e-3      # on top of the measured under-cover (it was measured on 20 000 surface samples per link)
mesh_spheres = {}
for l in fit["links"]:
    # (centre, fitted radius, covering radius, inscribed radius = depth of the centre in the hull + margin)
    mesh_spheres[l["link"]] = [(c, r, r + l["under_cover_m"] + SLACK, ri)
                               for c, r, ri in zip(l["center"], l["radius"], l["radius_inscribed"])]
hm = compile_robot(os.path.join(URDF_DIR, "iiwa14_polytope_collision_edit.urdf"), ee_link="iiwa_link_ee", name="iiwa14_hulls",
                   mesh_spheres=mesh_spheres, mesh_hulls=mesh_hulls, hull_margin=fit["margin"])
print("iiwa14_hulls:", hm.num_spheres, "spheres,", len(hm.hulls.verts), "hulls,", len(hm.hulls.pairs), "hull pairs,",
      int(hm.self_pair_mask.sum()), "sphere pairs")
with open(os.path.join(DATA, "iiwa14_hull_spheres.json"), "w") as fp:
    fp.write(hm.to_json())
np.savez_compressed(os.path.join(DATA, "iiwa14_collision_hulls.npz"),
                    **{n: v.astype(np.float32) for n, v in zip(hm.hulls.link_names, hm.hulls.verts)})
print("wrote", os.path.abspath(DATA))
