"""Extract the convex hulls Bullet builds for the arm the reference scripts load.   Run in the build container.

scripts/{our_method,avoid_all,ignore_all}_{single,multi}_branch.py load DRAKE_IIWA_URDF_EDIT
(plant_motion_planning/pybullet_tools/utils.py:78) = models/drake/iiwa_description/urdf/iiwa14_polytope_collision_edit.urdf.
Its <collision> elements (:89-97 ... :299-307) are the meshes meshes/collision/link_{1..6}.obj and
link_7_polytope.obj; iiwa_link_0 and the end-effector links carry none (so they never collide).  PyBullet turns every
URDF collision mesh into a btConvexHullShape of its vertices, so the collision geometry is the convex hull of each
OBJ's vertex set.  This script reads the `v` records, takes scipy.spatial.ConvexHull and writes the hull vertices
(float64, link frame; the collision origins are all identity) plus the link names to
tests/golden/iiwa14_collision_hulls.npz, which travels with the repository (the reference tree does not).

    python tools/extract_iiwa_hulls.py
"""
import os
import re
import sys

import numpy as np
from scipy.spatial import ConvexHull

REF = os.environ.get("PMP_REFERENCE", "/root/reference")
URDF = os.path.join(REF, "models/drake/iiwa_description/urdf/iiwa14_polytope_collision_edit.urdf")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "iiwa14_collision_hulls.npz")


def collision_meshes(urdf_path):
    """(link name, mesh path) of every <collision><geometry><mesh> in the URDF, in file order."""
    text = open(urdf_path).read()
    text = re.sub(r"<!--.*?-->", "", text, flags=re.S)
    out = []
    for m in re.finditer(r'<link name="([^"]+)"\s*>(.*?)</link>', text, flags=re.S):
        name, body = m.group(1), m.group(2)
        for c in re.finditer(r"<collision>(.*?)</collision>", body, flags=re.S):
            origin = re.search(r'<origin rpy="([^"]+)" xyz="([^"]+)"', c.group(1))
            if origin:
                assert [float(v) for v in origin.group(1).split()] == [0, 0, 0], "non-identity collision origin"
                assert [float(v) for v in origin.group(2).split()] == [0, 0, 0], "non-identity collision origin"
            mesh = re.search(r'<mesh filename="([^"]+)"', c.group(1))
            assert mesh, f"{name}: collision geometry is not a mesh"
            out.append((name, os.path.normpath(os.path.join(os.path.dirname(urdf_path), mesh.group(1)))))
    return out


def joint_chain(urdf_path):
    """The URDF's joints in file order: (name, type, parent, child, xyz, rpy, axis, lower, upper) -- parsed from the
    text here so that the hull oracle's forward kinematics share no code with the product's URDF compiler."""
    text = open(urdf_path).read()
    text = re.sub(r"<!--.*?-->", "", text, flags=re.S)
    text = text.split("<transmission")[0]
    out = []
    for m in re.finditer(r'<joint name="([^"]+)" type="([^"]+)">(.*?)</joint>', text, flags=re.S):
        name, kind, body = m.groups()
        parent = re.search(r'<parent link="([^"]+)"', body).group(1)
        child = re.search(r'<child link="([^"]+)"', body).group(1)
        o = re.search(r'<origin rpy="([^"]+)" xyz="([^"]+)"', body)
        rpy = [float(v) for v in o.group(1).split()]
        xyz = [float(v) for v in o.group(2).split()]
        ax = re.search(r'<axis xyz="([^"]+)"', body)
        axis = [float(v) for v in ax.group(1).split()] if ax else [0.0, 0.0, 0.0]
        lim = re.search(r'<limit [^>]*lower="([^"]+)" upper="([^"]+)"', body)
        lo, hi = (float(lim.group(1)), float(lim.group(2))) if lim else (0.0, 0.0)
        out.append((name, kind, parent, child, xyz, rpy, axis, lo, hi))
    return out


def obj_vertices(path):
    v = [[float(x) for x in line.split()[1:4]] for line in open(path) if line.startswith("v ")]
    return np.asarray(v, dtype=np.float64)


def main():
    meshes = collision_meshes(URDF)
    data = {"link_names": np.array([n for n, _ in meshes])}
    for name, path in meshes:
        V = obj_vertices(path)
        hull = ConvexHull(V)
        hv = V[np.sort(hull.vertices)]
        data[f"verts_{name}"] = hv
        data[f"faces_{name}"] = np.searchsorted(np.sort(hull.vertices), hull.simplices).astype(np.int32)
        # outward orientation of the triangles
        tri = hv[data[f"faces_{name}"]]
        nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        flip = np.einsum("ij,ij->i", nrm, tri[:, 0] - hv.mean(0)) < 0
        data[f"faces_{name}"][flip] = data[f"faces_{name}"][flip][:, ::-1]
        print(f"{name}: {os.path.relpath(path, REF)}  {len(V)} vertices -> hull {len(hv)} vertices, "
              f"{len(hull.simplices)} faces, volume {hull.volume:.6f}")
    chain = joint_chain(URDF)
    data["joint_names"] = np.array([c[0] for c in chain])
    data["joint_types"] = np.array([c[1] for c in chain])
    data["joint_parent"] = np.array([c[2] for c in chain])
    data["joint_child"] = np.array([c[3] for c in chain])
    data["joint_xyz"] = np.array([c[4] for c in chain], dtype=np.float64)
    data["joint_rpy"] = np.array([c[5] for c in chain], dtype=np.float64)
    data["joint_axis"] = np.array([c[6] for c in chain], dtype=np.float64)
    data["joint_limits"] = np.array([[c[7], c[8]] for c in chain], dtype=np.float64)
    for c in chain:
        print("joint", c[0], c[1], c[2], "->", c[3], "xyz", c[4], "rpy", c[5], "axis", c[6], "limits", c[7], c[8])
    np.savez_compressed(OUT, **data)
    print("wrote", os.path.abspath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
