"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN PYTHON in the build container.

    python tests/golden/gen_reference_vectors.py        (needs /root/reference; writes tests/golden/*.json)

PyBullet cannot be installed here, so the reference is imported against a stub ``pybullet`` module that
implements only data access (joint info, link states, quaternion helpers with Bullet's published semantics).
Everything that is *arithmetic of the reference* runs unmodified from /root/reference:

  reference_planner_fns.json   get_difference_fn / get_distance_fn / get_extend_fn (+ reversed, i.e.
                               asymmetric_extend backward) / get_limits_fn / get_sample_fn
                               (pybullet_tools/utils.py:3593-3676, 3753-3762; motion_planners/.../primitives.py:16-19)
                               and get_self_link_pairs / get_moving_links (:3718-3751) on the iiwa14 tree
  reference_sampler.json       create_plant_params (rand_gen_plants_programmatically_main.py:71-277) for several seeds
  reference_observers.json     CharacterizePlant2 / CharacterizePlant / TwoAngleRepresentation observe_all +
                               observe_all_and_check (representation.py) on link poses served by the stub
  reference_gate.json          get_collision_fn_with_angle_constraints_v4 (pybullet_tools/utils.py:3827-3857):
                               result and number of np.random draws for scripted (rigid, exceeded) flags
  reference_rrt.json           the reference's rrt_connect_multi_world run unchanged against this repository's
                               MultiPlantWorld facade (oracle-backed here): returned paths + RNG state for the three
                               multi_world_* script scenarios
  reference_planner.json       the reference's Planner.move_arm_conf2conf_multi_world + execute_path_multi_world run
                               unchanged (whole call chain down to the RRT loop) on a robot with Val's joint layout
  reference_fk.json            iiwa14 link frames composed with the reference's pybullet_tools/transformations.py
  seed1_plant.json             models/plant.urdf + models/plant_params.pkl (the one artefact of the reference that
                               carries sampler outputs) as plain numbers
  plant_fixtures.json          environments/multi_branch/env1..4 (URDF numbers + pickle) as input fixtures
  transformations_check        the stub's quaternion/Euler helpers are asserted against the reference's
                               pybullet_tools/transformations.py before anything is written

The reference tree does not travel to the GPU box; the JSON files do.
"""
import collections
import collections.abc
import json
import math
import os
import pickle
import random
import sys
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PMP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

for _n in ("Mapping", "MutableSet", "MutableMapping", "Sequence", "Iterable", "Set", "Callable"):
    if not hasattr(collections, _n):
        setattr(collections, _n, getattr(collections.abc, _n))

from plant_motion_planning_b200.model.transforms import Frame, matrix_to_quat, quat_to_matrix, rpy_to_matrix  # noqa: E402
from plant_motion_planning_b200.model.urdf import parse_urdf  # noqa: E402


# ----------------------------------------------------------------------------- Bullet helper semantics
def bullet_quat_from_euler(e):
    r, p, y = (0.5 * float(v) for v in e)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
            cr * cp * cy + sr * sp * sy)


def bullet_euler_from_quat(q):
    x, y, z, w = (float(v) for v in q)
    sqx, sqy, sqz, sqw = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return (0.0, -0.5 * math.pi, 2 * math.atan2(x, -y))
    if sarg >= 0.99999:
        return (0.0, 0.5 * math.pi, 2 * math.atan2(-x, y))
    return (math.atan2(2 * (y * z + w * x), sqw - sqx - sqy + sqz), math.asin(sarg),
            math.atan2(2 * (x * y + w * z), sqw + sqx - sqy - sqz))


def quat_mul(p, q):
    px, py, pz, pw = p
    qx, qy, qz, qw = q
    return (pw * qx + px * qw + py * qz - pz * qy, pw * qy + py * qw + pz * qx - px * qz,
            pw * qz + pz * qw + px * qy - py * qx, pw * qw - px * qx - py * qy - pz * qz)


def bullet_quat_difference(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    if np.dot(a - b, a - b) >= np.dot(a + b, a + b):
        b = -b
    return quat_mul(tuple(b), (-a[0], -a[1], -a[2], a[3]))


# ----------------------------------------------------------------------------- stub pybullet
class StubBullet(types.ModuleType):
    JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_SPHERICAL, JOINT_PLANAR, JOINT_FIXED = 0, 1, 2, 3, 4
    GEOM_SPHERE, GEOM_BOX, GEOM_CYLINDER, GEOM_MESH, GEOM_PLANE, GEOM_CAPSULE = 2, 3, 4, 5, 6, 7
    DIRECT, GUI = 2, 1
    LINK_FRAME, WORLD_FRAME = 1, 2
    POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL = 2, 0, 1

    def __init__(self):
        super().__init__("pybullet")
        self.bodies = {}
        self.shape_counter = 0
        self.closest_script = None   # callable(bodyA, bodyB, linkA, linkB) -> list

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name.isupper():
            return 0
        def missing(*a, **k):
            raise NotImplementedError(f"stub pybullet has no {name}")
        return missing

    # data access
    def getNumJoints(self, body, physicsClientId=0):
        return len(self.bodies[body]["joints"])

    def getJointInfo(self, body, joint, physicsClientId=0):
        return self.bodies[body]["joints"][joint]

    def getBodyInfo(self, body, physicsClientId=0):
        return (self.bodies[body]["root"].encode(), self.bodies[body]["name"].encode())

    def getCollisionShapeData(self, body, link, physicsClientId=0):
        return [(body, link, 3, (1, 1, 1), b"", (0, 0, 0), (0, 0, 0, 1))] if self.bodies[body]["collides"](link) else []

    def getLinkState(self, body, link, *a, **k):
        return self.bodies[body]["link_state"](link)

    def getBasePositionAndOrientation(self, body, physicsClientId=0):
        return self.bodies[body]["base"]

    def getClosestPoints(self, bodyA, bodyB, distance=0, linkIndexA=-2, linkIndexB=-2, physicsClientId=0):
        return self.closest_script(bodyA, bodyB, linkIndexA, linkIndexB)

    # math helpers (Bullet semantics)
    def getQuaternionFromEuler(self, e):
        return bullet_quat_from_euler(e)

    def getEulerFromQuaternion(self, q):
        return bullet_euler_from_quat(q)

    def getDifferenceQuaternion(self, a, b):
        return bullet_quat_difference(a, b)

    def createCollisionShape(self, *a, **k):
        self.shape_counter += 1
        return self.shape_counter


pb = StubBullet()
sys.modules["pybullet"] = pb
_pd = types.ModuleType("pybullet_data")
_pd.getDataPath = lambda: "/nonexistent"
sys.modules["pybullet_data"] = _pd
_pu = types.ModuleType("pybullet_utils")
_pu.urdfEditor = types.ModuleType("pybullet_utils.urdfEditor")
sys.modules["pybullet_utils"] = _pu
sys.modules["pybullet_utils.urdfEditor"] = _pu.urdfEditor
_rp = types.ModuleType("rospkg")
_rp.RosPack = type("RosPack", (), {"get_path": lambda self, n: "/nonexistent/" + n})
_rp.get_ros_root = lambda: "/nonexistent"
sys.modules["rospkg"] = _rp

import plant_motion_planning.pybullet_tools.transformations as T  # noqa: E402
import plant_motion_planning.pybullet_tools.utils as U  # noqa: E402
import plant_motion_planning.rand_gen_plants_programmatically_main as G  # noqa: E402
import plant_motion_planning.representation as REP  # noqa: E402
from plant_motion_planning.motion_planners.motion_planners.primitives import asymmetric_extend  # noqa: E402


def check_transformations():
    """Stub quaternion helpers vs the reference's transformations.py (its quaternions are x, y, z, w)."""
    rng = np.random.RandomState(5)
    for _ in range(200):
        e = rng.uniform(-1.5, 1.5, 3)
        q = bullet_quat_from_euler(e)
        qt = T.quaternion_from_euler(e[0], e[1], e[2], axes="sxyz")
        assert np.allclose(qt, q, atol=1e-12)
        back = bullet_euler_from_quat(q)
        assert np.allclose(back, T.euler_from_quaternion(qt, axes="sxyz"), atol=1e-9)
        assert np.allclose(quat_to_matrix(q), T.quaternion_matrix(qt)[:3, :3], atol=1e-12)
        assert np.allclose(rpy_to_matrix(e), T.euler_matrix(e[0], e[1], e[2], axes="sxyz")[:3, :3], atol=1e-12)
        e2 = rng.uniform(-1.5, 1.5, 3)
        q2 = bullet_quat_from_euler(e2)
        d = bullet_quat_difference(q, q2)
        want = quat_to_matrix(q2) @ quat_to_matrix(q).T
        assert np.allclose(quat_to_matrix(d), want, atol=1e-12)


# ----------------------------------------------------------------------------- iiwa14 as a stub body
def register_urdf_body(body_id, urdf_path, extra_continuous=False):
    tree = parse_urdf(urdf_path)
    joints = []
    for i, j in enumerate(tree.joints):
        jt = {"revolute": pb.JOINT_REVOLUTE, "continuous": pb.JOINT_REVOLUTE, "fixed": pb.JOINT_FIXED,
              "prismatic": pb.JOINT_PRISMATIC}[j.kind]
        lower, upper = j.lower, j.upper
        if j.kind == "fixed":
            lower, upper = 0.0, -1.0
        if extra_continuous and i == 3:
            lower, upper = 0.0, -1.0        # pretend joint 3 is continuous to pin the circular code path
        joints.append((i, j.name.encode(), jt, 7 + i, 6 + i, 0, 0.0, 0.0, lower, upper, 300.0, 10.0, j.child.encode(),
                       tuple(j.axis), tuple(j.origin.t), tuple(matrix_to_quat(j.origin.R)), tree.link_index(j.parent)))
    names = [tree.root] + [j.child for j in tree.joints]

    def collides(link):
        nm = names[link + 1]
        return len(tree.links[nm].collisions) > 0
    pb.bodies[body_id] = {"joints": joints, "root": tree.root, "name": tree.name, "collides": collides,
                          "link_state": None, "base": ((0, 0, 0), (0, 0, 0, 1))}
    return tree


def gen_planner_fns(out):
    urdf = os.path.join(REF, "models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf")
    res = {}
    for body, circ in ((0, False), (1, True)):
        register_urdf_body(body, urdf, extra_continuous=circ)
        joints = U.get_movable_joints(body)
        assert len(joints) == 7
        diff = U.get_difference_fn(body, joints)
        dist = U.get_distance_fn(body, joints)
        ext = U.get_extend_fn(body, joints)
        lim = U.get_limits_fn(body, joints)
        lower, upper = U.get_custom_limits(body, joints, {}, circular_limits=U.CIRCULAR_LIMITS)
        lower, upper = np.array(list(lower)), np.array(list(upper))
        rng = np.random.RandomState(11 + body)
        cases = []
        for k in range(40):
            q1 = rng.uniform(lower, upper)
            q2 = rng.uniform(lower, upper)
            if k % 4 == 0:
                q2 = q1 + rng.uniform(-0.08, 0.08, 7)       # short edges: 1-3 steps
            if k == 1:
                q2 = q1.copy()                                # zero-length edge
            fwd = [list(map(float, q)) for q in ext(tuple(q1), tuple(q2))]
            bwd = [list(map(float, q)) for q in asymmetric_extend(tuple(q1), tuple(q2), ext, backward=True)]
            def compact(seq):
                # full sequences would make the fixture megabytes: keep length, ends and the per-joint sum
                return {"n": len(seq), "head": seq[:2], "tail": seq[-2:],
                        "sum": [float(v) for v in np.sum(np.asarray(seq, dtype=np.float64), axis=0)]}
            cases.append({"q1": q1.tolist(), "q2": q2.tolist(), "difference": list(map(float, diff(tuple(q2), tuple(q1)))),
                          "distance": float(dist(tuple(q1), tuple(q2))),
                          "extend": fwd if k < 3 else None, "extend_compact": compact(fwd),
                          "extend_backward_compact": compact(bwd)})
        lim_cases = []
        for k in range(20):
            q = rng.uniform(lower - 0.2, upper + 0.2)
            lim_cases.append({"q": q.tolist(), "violated": bool(lim(tuple(q)))})
        lim_cases.append({"q": lower.tolist(), "violated": bool(lim(tuple(lower)))})
        lim_cases.append({"q": upper.tolist(), "violated": bool(lim(tuple(upper)))})
        np.random.seed(77)
        sfn = U.get_sample_fn(body, joints)
        samples = [list(map(float, sfn())) for _ in range(5)]
        res["circular" if circ else "iiwa14"] = {
            "joints": list(map(int, joints)), "lower": lower.tolist(), "upper": upper.tolist(),
            "circular": [bool(U.is_circular(body, j)) for j in joints], "resolution": float(U.DEFAULT_RESOLUTION),
            "cases": cases, "limits": lim_cases, "sample_seed": 77, "samples": samples,
        }
    joints = U.get_movable_joints(0)
    res["iiwa14"]["moving_links"] = sorted(map(int, U.get_moving_links(0, joints)))
    res["iiwa14"]["self_link_pairs"] = sorted([sorted(map(int, p)) for p in U.get_self_link_pairs(0, joints)])
    # the URDF the scripts actually load (DRAKE_IIWA_URDF_EDIT): mesh collision on links 1-7, none on link 0
    register_urdf_body(2, os.path.join(REF, "models/drake/iiwa_description/urdf/iiwa14_polytope_collision_edit.urdf"))
    joints2 = U.get_movable_joints(2)
    res["iiwa14_polytope_edit"] = {
        "joints": list(map(int, joints2)),
        "moving_links": sorted(map(int, U.get_moving_links(2, joints2))),
        "self_link_pairs": sorted([sorted(map(int, p)) for p in U.get_self_link_pairs(2, joints2)]),
    }
    out["reference_planner_fns.json"] = res


def gen_sampler(out):
    res = []
    for seed, args, xy in ((1, (1, 2, 1), (0.3, -0.2)), (7, (2, 2, 1), (0.1, -0.4)), (42, (1, 3, 2), (0.0, 0.0)),
                           (1003, (1, 2, 1), (0.2, -0.1)), (5, (2, 1, 2), (0.25, -0.3))):
        np.random.seed(seed)
        random.seed(seed)
        G.current_index = 0
        G.main_stem_index = 0
        G.base_points.clear()
        G.main_stem_indices.clear()
        G.link_parent_indices.clear()
        G.stem_half_height.clear()
        base, stems, mains = G.create_plant_params(args[0], args[1], args[2], list(xy))
        res.append({
            "seed": seed, "args": list(args), "base_pos_xy": list(xy), "base_pos": [float(v) for v in base[3]],
            "link_positions": [[float(v) for v in p_] for p_ in stems[3]],
            "link_orientations": [[float(v) for v in q] for q in stems[4]],
            "link_parent_indices_arg": [int(v) for v in stems[7]],
            "main_stem_indices": [int(v) for v in mains],
            "base_points": {str(k): [float(x) for x in v] for k, v in G.base_points.items()},
            "link_parent_indices": {str(k): int(v) for k, v in G.link_parent_indices.items()},
            "stem_half_height": {str(k): float(v) for k, v in G.stem_half_height.items()},
            "next_np_uniform": float(np.random.uniform()), "next_py_random": float(random.random()),
        })
    out["reference_sampler.json"] = res


def plant_link_states(world, theta):
    """COM-frame pose of every link (v1 and stem links alternate) of a PlantWorld at stem angles theta."""
    from plant_motion_planning_b200.model.transforms import rot_x, rot_y
    frames = []
    states = {}
    for k, s in enumerate(world.stems):
        parent = Frame() if s.parent < 0 else frames[s.parent]
        v1 = parent @ s.origin @ Frame(rot_x(theta[k][0]))
        st = v1 @ Frame(rot_y(theta[k][1]))
        frames.append(st)
        qv, qs = matrix_to_quat(v1.R), matrix_to_quat(st.R)
        com = st.apply([0, 0, s.com_z])
        states[2 * k] = (tuple(v1.t), tuple(qv), (0, 0, 0), (0, 0, 0, 1), tuple(v1.t), tuple(qv))
        states[2 * k + 1] = (tuple(com), tuple(qs), (0, 0, s.com_z), (0, 0, 0, 1), tuple(st.t), tuple(qs))
    return states


def gen_observers(out):
    from plant_motion_planning_b200.model.plants import load_plant, sample_world, single_stem_world
    res = {"mod3": [], "mod": [], "single": []}
    rng = np.random.RandomState(3)
    # generated plants -> CharacterizePlant2 / TwoAngleRepresentation_mod3
    for seed, args in ((1000, (1, 2, 1)), (1001, (1, 2, 1)), (7, (2, 2, 1))):
        world = sample_world(args[0], args[1], args[2], ((0, 0.35), (-0.5, 0.0)),
                             np_random=np.random.RandomState(seed), py_random=random.Random(seed))
        P = world.num_stems
        body = 10 + seed
        zero = [[0.0, 0.0]] * P
        cur = {"st": plant_link_states(world, zero)}
        pb.bodies[body] = {"joints": [None] * (2 * P), "link_state": lambda l, c=cur: c["st"][l],
                           "base": (tuple(world.base_pos), (0, 0, 0, 1)), "root": "b", "name": "p", "collides": lambda l: True}
        rep = REP.CharacterizePlant2(body, world.base_points, world.main_stem_indices, world.link_parent_indices)
        cases = []
        for k in range(12):
            th = rng.uniform(-1.0, 1.0, size=(P, 2)) * (rng.uniform(size=(P, 1)) < 0.6)
            if k == 0:
                th[:] = 0
            cur["st"] = plant_link_states(world, th)
            exceeded = bool(rep.observe_all_and_check(0.30))
            cases.append({"theta": th.tolist(), "deflections": [float(d) for d in rep.deflections], "exceeded_0.30": exceeded})
        res["mod3"].append({"seed": seed, "args": list(args), "xy_limits": [[0, 0.35], [-0.5, 0.0]], "cases": cases})
    # saved plants -> CharacterizePlant / TwoAngleRepresentation_mod
    for env in ("env1", "env3"):
        d = os.path.join(REF, "environments/multi_branch", env)
        world = load_plant(os.path.join(d, "plant.urdf"), os.path.join(d, "plant_params.pkl"))
        P = world.num_stems
        body = 50 + int(env[-1])
        cur = {"st": plant_link_states(world, [[0.0, 0.0]] * P)}
        pb.bodies[body] = {"joints": [None] * (2 * P), "link_state": lambda l, c=cur: c["st"][l],
                           "base": (tuple(world.base_pos), (0, 0, 0, 1)), "root": "b", "name": "p", "collides": lambda l: True}
        rep = REP.CharacterizePlant(body, world.base_points, world.main_stem_indices)
        cases = []
        for k in range(12):
            th = rng.uniform(-1.0, 1.0, size=(P, 2)) * (rng.uniform(size=(P, 1)) < 0.6)
            if k == 0:
                th[:] = 0
            cur["st"] = plant_link_states(world, th)
            exceeded = bool(rep.observe_all_and_check(0.30))
            cases.append({"theta": th.tolist(), "deflections": [float(d) for d in rep.deflections], "exceeded_0.30": exceeded})
        res["mod"].append({"env": env, "cases": cases})
    # single stems -> TwoAngleRepresentation (body-level observe)
    world = single_stem_world([[0.4, 0.1], [0.12, -0.15], [-0.25, 0.6]])
    cases = []
    for k in range(10):
        th = rng.uniform(-1.2, 1.2, size=(world.num_stems, 2))
        if k == 0:
            th[:] = 0
        if k == 1:
            th[:] = [[3e-6, 0.0]] * world.num_stems      # exercises the 1e-5 snap
        # y_first stems: kernel angles (tx, ty) = (first joint about y, -second joint about x)
        states = plant_link_states(world, th)
        defl = []
        for i in range(world.num_stems):
            body = 70 + i
            base_t = world.base_boxes[i][0].t
            pb.bodies[body] = {"joints": [None, None], "link_state": lambda l, s=states, i=i: s[2 * i + l],
                               "base": (tuple(base_t), (0, 0, 0, 1)), "root": "b", "name": "p", "collides": lambda l: True}
            r = REP.TwoAngleRepresentation(body, 1)
            r.observe()
            defl.append(float(r.deflection))
        cases.append({"theta": th.tolist(), "deflections": defl})
    res["single"].append({"positions": [[0.4, 0.1], [0.12, -0.15], [-0.25, 0.6]], "cases": cases})
    out["reference_observers.json"] = res


def gen_gate(out):
    """Run the reference's v4 collision closure on scripted flags and count its np.random draws."""
    urdf = os.path.join(REF, "models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf")
    register_urdf_body(0, urdf)
    joints = U.get_movable_joints(0)

    class FakeRep:
        def __init__(self):
            self.flag = False
            self.deflections = []

        def observe_all_and_check(self, limit):
            return self.flag

    rep = FakeRep()
    state = {"rigid": False}
    pb.closest_script = lambda a, b, la, lb: ([(0,) * 14] if state["rigid"] else [])
    pb.bodies[5] = {"joints": [], "root": "floor", "name": "floor", "collides": lambda l: True, "link_state": None,
                    "base": ((0, 0, 0), (0, 0, 0, 1))}
    fn = U.get_collision_fn_with_angle_constraints_v4(0, [5], joints, [rep], 0.3, [], True, set(), custom_limits={},
                                                      max_distance=U.MAX_DISTANCE, use_aabb=False, cache=True)
    q_ok = tuple([0.1] * 7)
    q_bad = tuple([9.0] * 7)
    cases = []
    import contextlib
    import io
    for seed in range(6):
        np.random.seed(seed)
        rng = np.random.RandomState(100 + seed)
        seq = []
        for k in range(40):
            rigid, exc, lim = bool(rng.rand() < 0.2), bool(rng.rand() < 0.5), bool(rng.rand() < 0.1)
            state["rigid"], rep.flag = rigid, exc
            with contextlib.redirect_stdout(io.StringIO()):
                r = bool(fn(q_bad if lim else q_ok))
            seq.append({"limit": lim, "rigid": rigid, "exceeded": exc, "result": r})
        # the next draw of the global stream pins how many gate draws the reference consumed
        cases.append({"seed": seed, "sequence": seq, "next_uniform_after": float(np.random.rand())})
    out["reference_gate.json"] = {"gate_probability": 0.95, "cases": cases}


def gen_fk(out):
    """Link frames of the iiwa14 URDF composed with the REFERENCE's transformations.py (euler_matrix /
    rotation_matrix / translation_matrix / quaternion_from_matrix): an implementation of URDF forward kinematics
    that shares no code with the oracle or the kernels."""
    urdf = os.path.join(REF, "models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf")
    tree = parse_urdf(urdf)
    rng = np.random.RandomState(21)
    movable = [j for j in tree.joints if j.kind == "revolute"]
    lower = np.array([j.lower for j in movable])
    upper = np.array([j.upper for j in movable])
    cases = []
    import xml.etree.ElementTree as ET
    root = ET.parse(urdf).getroot()
    origin = {}
    for je in root.findall("joint"):
        if je.find("parent") is None:
            continue
        o = je.find("origin")
        xyz = [float(v) for v in o.get("xyz").split()]
        rpy = [float(v) for v in o.get("rpy").split()]
        ax = je.find("axis")
        axis = [float(v) for v in ax.get("xyz").split()] if ax is not None else [1.0, 0.0, 0.0]
        origin[je.get("name")] = (xyz, rpy, axis)
    inertial = {}
    for le in root.findall("link"):
        ine = le.find("inertial")
        xyz = [0.0, 0.0, 0.0]
        if ine is not None and ine.find("origin") is not None:
            xyz = [float(v) for v in ine.find("origin").get("xyz").split()]
        inertial[le.get("name")] = xyz
    for k in range(24):
        q = rng.uniform(lower, upper) if k else np.zeros(7)
        qmap = {j.name: float(v) for j, v in zip(movable, q)}
        world = {tree.root: np.eye(4)}
        pos, quat, com = [], [], []
        for j in tree.joints:
            xyz, rpy, axis = origin[j.name]
            M = T.concatenate_matrices(T.translation_matrix(xyz), T.euler_matrix(rpy[0], rpy[1], rpy[2], axes="sxyz"))
            if j.kind == "revolute":
                M = T.concatenate_matrices(M, T.rotation_matrix(qmap[j.name], axis))
            W = world[j.parent] @ M
            world[j.child] = W
            pos.append([float(v) for v in W[:3, 3]])
            quat.append([float(v) for v in T.quaternion_from_matrix(W)])       # x, y, z, w in this fork
            com.append([float(v) for v in (W @ np.array(inertial[j.child] + [1.0]))[:3]])
        cases.append({"q": [float(v) for v in q], "link_pos": pos, "link_quat": quat, "com_pos": com})
    out["reference_fk.json"] = {"urdf": "models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf",
                                "links": [j.child for j in tree.joints], "cases": cases}


def gen_rrt(out):
    """Run the reference's rrt_connect_multi_world (rrt_connect_with_angle_constraints_v2.py:202-280) UNCHANGED against
    the MultiPlantWorld facade of this repository.  No GPU exists in the build container, so the facade is given an
    oracle-backed checker (tests/oracle_checker.py); what is pinned is the planner-facing behaviour: the facade drops
    into the reference's RRT loop, and birrt.py (the restated driver) must return exactly these paths and leave the RNG
    streams in exactly this state."""
    import contextlib
    import io
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_checker import OracleChecker
    from plant_motion_planning.motion_planners.motion_planners.rrt_connect_with_angle_constraints_v2 import \
        rrt_connect_multi_world
    from plant_motion_planning_b200.env import MultiPlantWorld
    from plant_motion_planning_b200.scenarios import GOAL, SCRIPTS, START
    pb.saveState = lambda *a, **k: 1
    pb.restoreState = lambda *a, **k: None
    urdf = os.path.join(REF, "models/drake/iiwa_description/urdf/iiwa14_spheres_collision.urdf")
    register_urdf_body(0, urdf)
    joints = U.get_movable_joints(0)
    res = {}
    for script, cfg in SCRIPTS.items():
        np.random.seed(cfg["seed"])
        random.seed(cfg["seed"])
        env = MultiPlantWorld(cfg["xs"], cfg["ys"], cfg["deflection_limit"], *cfg["args"], cfg["limits"],
                              avoid_all=cfg["avoid_all"], checker_factory=OracleChecker)
        plant_aware = script == "multi_world_our_method"
        if script == "multi_world_avoid_all":
            # start in collision with the fourth plant (see scenarios.py): the reference's check_initial_end_multi_world
            # (pybullet_tools/utils.py:4358-4380) would exit(); its RRT loop returns None at once
            env.step(START)
            assert env.net_constraint_violation_checker_benchmark(START)
            res[script] = {"planner_seed": 1, "plant_aware": False, "path": None, "start_in_collision": True}
            continue
        fwd = env.net_constraint_violation_checker_forward if plant_aware else env.net_constraint_violation_checker_benchmark
        back = env.net_constraint_violation_checker_backward if plant_aware else env.net_constraint_violation_checker_benchmark
        np.random.seed(1)
        random.seed(1)
        with contextlib.redirect_stdout(io.StringIO()):
            r = rrt_connect_multi_world(START, 1, GOAL, U.get_distance_fn(0, joints), U.get_sample_fn(0, joints),
                                        U.get_extend_fn(0, joints), [fwd, back], env)
        path = r[0]
        res[script] = {"planner_seed": 1, "plant_aware": plant_aware, "path_len": len(path),
                       "path": [[float(v) for v in q] for q in path],
                       "next_np_rand": float(np.random.rand()), "kernel_evaluations": int(env.kernel_evaluations)}
        # the reference's shortcut smoother (smoothing.py:55-141), unchanged, on that path
        from plant_motion_planning.motion_planners.motion_planners.smoothing import smooth_path_multi_world
        np.random.seed(7)
        random.seed(7)
        with contextlib.redirect_stdout(io.StringIO()):
            sm = smooth_path_multi_world(path, r[1], U.get_extend_fn(0, joints), fwd, env, max_iterations=20)
        res[script]["smoothing_seed"] = 7
        res[script]["smoothed"] = [[float(v) for v in q] for q in sm]
        res[script]["smoothed_next_np_rand"] = float(np.random.rand())
        res[script]["smoothed_next_py_random"] = float(random.random())
        # the reference's cost report: Command.execute_multi_world (kuka_primitives.py:467-521), unchanged, on the
        # smoothed path.  Its getLinkState(sample_robot, 9) is served by the stub from the oracle FK at the
        # configuration the facade was last stepped to; the totals are parsed from what it prints.
        from oracle import edgecheck_oracle as ORC
        from plant_motion_planning.pybullet_tools.kuka_primitives import BodyPath, Command
        robot_model = env.sample_robot

        def ee_state(link, _env=env, _m=robot_model):
            pos, quat, com = ORC.link_poses(_m, np.asarray(_env._action, dtype=np.float64)[None])
            return (tuple(com[0, link]), tuple(quat[0, link]), (0, 0, 0), (0, 0, 0, 1), tuple(pos[0, link]), tuple(quat[0, link]))
        pb.bodies[0]["link_state"] = ee_state

        class _Proxy:
            sample_robot = 0
            def __init__(self, e):
                self._e = e
            def __getattr__(self, k):
                return getattr(self._e, k)
        env.checker.config = env.config
        env.step(START)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            Command([BodyPath(0, sm, joints)]).execute_multi_world(START, _Proxy(env), num_worlds=len(env.envs))
        txt = buf.getvalue()
        import re
        ee_len = float(re.search(r"total path cost:\s+([-0-9.e+]+)", txt).group(1))
        over = [float(v) for v in re.search(r"total Deflection over limit:\s+\[([^\]]*)\]", txt).group(1).split()]
        total = [float(v) for v in re.search(r"Total cost:\s+\[([^\]]*)\]", txt, flags=re.S).group(1).split()]
        res[script]["execute"] = {"ee_len": ee_len, "over": over, "total": total, "alpha": 0.1}
    out["reference_rrt.json"] = res


def gen_planner(out):
    """The reference's whole planning entry point, UNCHANGED: Planner.move_arm_conf2conf_multi_world (planner.py:18-51) ->
    get_free_motion_gen_with_angle_constraints_multi_world -> plan_joint_motion_with_angle_constraints_multi_world ->
    birrt_multi_world -> random_restarts_multi_world -> rrt_connect_multi_world / smooth_path_multi_world, followed by
    Planner.execute_path_multi_world, run against this repository's MultiPlantWorld facade (oracle-backed here).
    BodyConf hard-codes Val's joint layout (val_utils.py:12-23: the arm is movable joints 6..14), so the robot is a
    synthetic URDF with that layout (tests/synthetic_robot.py: 4 wheels, 2 torso joints, a 9-joint arm)."""
    import contextlib
    import io
    import re
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import edgecheck_oracle as ORC
    from oracle_checker import OracleChecker
    from plant_motion_planning.planner import Planner
    from plant_motion_planning.pybullet_tools.val_utils import get_arm_joints
    from plant_motion_planning_b200.env import MultiPlantWorld
    from synthetic_robot import val_like_robot, val_like_urdf_text
    pb.saveState = lambda *a, **k: 1
    pb.restoreState = lambda *a, **k: None
    pb.resetJointState = lambda *a, **k: None
    model, tree = val_like_robot()
    with tempfile.NamedTemporaryFile("w", suffix=".urdf", delete=False) as fp:
        fp.write(val_like_urdf_text())
        path_urdf = fp.name
    register_urdf_body(0, path_urdf)
    os.unlink(path_urdf)
    joints = get_arm_joints(0, is_left=True, include_torso=False)
    assert list(joints) == [6, 7, 8, 9, 10, 11, 12, 13, 14]
    np.random.seed(19)
    random.seed(19)
    env = MultiPlantWorld((0, 5), (0, 5), 0.30, 1, 2, 1, ((0.2, 0.35), (-0.5, 0.0)), robot=model,
                          checker_factory=OracleChecker)
    # goal: the accepted configuration (no flag in any world) that bends the plants most -- so the gate is exercised
    rng = np.random.RandomState(2)
    cands = rng.uniform(model.lower, model.upper, size=(6000, 9))
    r = env.checker.check_configs_host(cands.astype(np.float32), detail=True)
    ok = (r["flags"] == 0).all(axis=1)
    score = np.where(ok, r["deflection"].sum(axis=(1, 2)), -1.0)
    goal = tuple(float(v) for v in cands[int(np.argmax(score))])
    start = tuple([0.0] * 9)

    def ee_state(link, _env=env, _m=model):
        pos, quat, com = ORC.link_poses(_m, np.asarray(_env._action, dtype=np.float64)[None])
        return (tuple(com[0, link]), tuple(quat[0, link]), (0, 0, 0), (0, 0, 0, 1), tuple(pos[0, link]), tuple(quat[0, link]))
    pb.bodies[0]["link_state"] = ee_state

    class _Proxy:
        sample_robot = 0
        def __init__(self, e):
            self._e = e
            self.joints = list(joints)
        def __getattr__(self, k):
            return getattr(self._e, k)
    penv = _Proxy(env)
    from plant_motion_planning_b200 import env as facade
    pb.saveState = lambda *a, **k: facade.save_state()
    pb.restoreState = lambda sid=None, *a, **k: facade.restore_state(sid if sid is not None else k.get("stateId"))
    env.step(start)
    saved_world = pb.saveState()                       # scripts/multi_world_our_method.py:111
    np.random.seed(1)
    random.seed(1)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        planner = Planner()
        cmd = planner.move_arm_conf2conf_multi_world(start, goal, penv)
        nxt_np, nxt_py = float(np.random.rand()), float(random.random())
        pb.restoreState(saved_world)                   # scripts/multi_world_our_method.py:118
        planner.execute_path_multi_world(penv)
    txt = buf.getvalue()
    path = [[float(v) for v in q] for q in cmd.body_paths[0].path]
    ee_len = float(re.findall(r"total path cost:\s+([-0-9.e+]+)", txt)[-1])
    over = [float(v) for v in re.findall(r"total Deflection over limit:\s+\[([^\]]*)\]", txt)[-1].split()]
    result = {
        "robot": "tests/synthetic_robot.py:val_like_robot()", "world_seed": 19, "planner_seed": 1,
        "start": list(start), "goal": list(goal), "goal_deflection_sum": float(score.max()),
        "path": path, "next_np_rand": nxt_np, "next_py_random": nxt_py,
        "kernel_evaluations": int(env.kernel_evaluations),
        "execute": {"ee_len": ee_len, "over": over},
    }
    # ---- the benchmark entry the avoid_all / ignore_all scripts use (planner.py:53-78), same worlds with avoid_all ----
    np.random.seed(19)
    random.seed(19)
    env2 = MultiPlantWorld((0, 5), (0, 5), 0.30, 1, 2, 1, ((0.2, 0.35), (-0.5, 0.0)), robot=model, avoid_all=True,
                           checker_factory=OracleChecker)
    r2 = env2.checker.check_configs_host(cands.astype(np.float32), detail=False)
    ok2 = (r2["flags"] & 1 == 0).all(axis=1)
    ee_c = ORC.ee_positions(model, cands)
    score2 = np.where(ok2, -np.linalg.norm(ee_c - np.array([0.27, -0.25, 0.45]), axis=1), -1e9)
    goal2 = tuple(float(v) for v in cands[int(np.argmax(score2))])
    pb.bodies[0]["link_state"] = lambda link, _env=env2, _m=model: ee_state(link, _env, _m)
    pb.getJointState = lambda body, joint, *a, **k: (float(env2._action[list(joints).index(joint)]), 0.0, (0.0,) * 6, 0.0)
    penv2 = _Proxy(env2)
    env2.step(start)
    saved2 = pb.saveState()
    np.random.seed(1)
    random.seed(1)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        planner2 = Planner()
        cmd2 = planner2.move_arm_conf2conf_multi_world_benchmark(start, goal2, penv2, saved2)
        nxt2 = (float(np.random.rand()), float(random.random()))
    result["benchmark"] = {
        "avoid_all": True, "goal": list(goal2),
        "path": [[float(v) for v in q] for q in cmd2.body_paths[0].path],
        "next_np_rand": nxt2[0], "next_py_random": nxt2[1], "kernel_evaluations": int(env2.kernel_evaluations),
    }
    out["reference_planner.json"] = result


def gen_fixtures(out):
    def plant_numbers(urdf_path, pkl_path):
        tree = parse_urdf(urdf_path)
        with open(pkl_path, "rb") as fp:
            joint_list, mains, base_points, base_pos = pickle.load(fp)
        with open(urdf_path) as fp:
            text = fp.read()
        return {"urdf": text, "joint_list": [int(v) for v in joint_list], "main_stem_indices": [int(v) for v in mains],
                "base_points": {str(k): [float(x) for x in v] for k, v in base_points.items()},
                "base_pos": [float(v) for v in base_pos], "n_joints": len(tree.joints)}
    out["seed1_plant.json"] = plant_numbers(os.path.join(REF, "models/plant.urdf"), os.path.join(REF, "models/plant_params.pkl"))
    fx = {}
    for env in ("env1", "env2", "env3", "env4"):
        d = os.path.join(REF, "environments/multi_branch", env)
        fx[env] = plant_numbers(os.path.join(d, "plant.urdf"), os.path.join(d, "plant_params.pkl"))
    out["plant_fixtures.json"] = fx
    import plant_motion_planning.utils as PU
    envs = {k: [[float(x) for x in p_] for p_ in v] for k, v in PU.envs.items()}
    out["single_stem_envs.json"] = envs


def main():
    check_transformations()
    out = {}
    gen_planner_fns(out)
    gen_sampler(out)
    gen_observers(out)
    gen_gate(out)
    gen_fk(out)
    gen_rrt(out)
    gen_planner(out)
    gen_fixtures(out)
    for name, obj in out.items():
        path = os.path.join(HERE, name)
        if name == "single_stem_envs.json":
            path = os.path.join(ROOT, "plant_motion_planning_b200", "data", name)
        with open(path, "w") as fp:
            json.dump(obj, fp, indent=0 if name in ("plant_fixtures.json", "seed1_plant.json") else None)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
