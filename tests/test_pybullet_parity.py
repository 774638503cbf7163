"""Check against the real engine -- runs only where PyBullet is importable (it is NOT in this image, SURVEY 8c:
"parity unpinned" for FK / contact at the Bullet boundary; the oracle is pinned on everything the reference computes
itself).  Set PMP_IIWA_URDF to the reference's models/drake/iiwa_description/urdf/iiwa14_polytope_collision_edit.urdf.

What it would establish:
  * pmp_fk / the oracle's link_poses == p.getLinkState(...)[4:6] (URDF link frame) and [0:2] (COM frame) after
    p.resetJointState -- the semantics stated in SURVEY 8c ("Bullet semantics the restatement must state explicitly");
  * the restated plant sampler's URDF loads and its stem link frames at zero deflection equal ours.
"""
import os

import numpy as np
import pytest

pybullet = pytest.importorskip("pybullet")


@pytest.fixture(scope="module")
def client():
    cid = pybullet.connect(pybullet.DIRECT)
    yield cid
    pybullet.disconnect(cid)


def test_link_frames_equal_get_link_state(client):
    path = os.environ.get("PMP_IIWA_URDF")
    if not path or not os.path.exists(path):
        pytest.skip("PMP_IIWA_URDF not set")
    from oracle import edgecheck_oracle as O
    from plant_motion_planning_b200.robots import load_iiwa14
    model = load_iiwa14()
    body = pybullet.loadURDF(path, useFixedBase=True, physicsClientId=client)
    joints = [j for j in range(pybullet.getNumJoints(body, physicsClientId=client))
              if pybullet.getJointInfo(body, j, physicsClientId=client)[2] != pybullet.JOINT_FIXED]
    rng = np.random.RandomState(0)
    for _ in range(32):
        q = rng.uniform(model.lower, model.upper)
        for j, v in zip(joints, q):
            pybullet.resetJointState(body, j, float(v), physicsClientId=client)
        pos, quat, com = O.link_poses(model, q[None])
        for link in range(model.num_links):
            st = pybullet.getLinkState(body, link, computeForwardKinematics=True, physicsClientId=client)
            assert np.abs(np.asarray(st[4]) - pos[0, link]).max() <= 1e-6
            assert np.abs(np.asarray(st[0]) - com[0, link]).max() <= 1e-6
            qq = np.asarray(st[5])
            assert min(np.abs(qq - quat[0, link]).max(), np.abs(qq + quat[0, link]).max()) <= 1e-6


def test_sampled_plant_urdf_loads_with_our_rest_frames(client, tmp_path):
    import random
    from oracle import edgecheck_oracle as O
    from plant_motion_planning_b200.model.plants import sample_world, save_plant
    np.random.seed(1)
    random.seed(1)
    w = sample_world(1, 2, 1, ((0, 0.35), (-0.5, 0)), stem_base_spacing=0.05)
    save_plant(w, str(tmp_path / "plant.urdf"), str(tmp_path / "plant_params.pkl"))
    body = pybullet.loadURDF(str(tmp_path / "plant.urdf"), basePosition=w.base_pos, useFixedBase=True,
                             physicsClientId=client)
    Rs, ts = O.frames_from_theta(w, np.zeros((1, w.num_stems, 2)))
    for k in range(w.num_stems):
        st = pybullet.getLinkState(body, 2 * k + 1, computeForwardKinematics=True, physicsClientId=client)
        assert np.abs(np.asarray(st[4]) - ts[0, k]).max() <= 2e-5          # URDF text carries 5 decimals


def test_hull_collision_booleans_equal_get_closest_points(client):
    """The rigid decision itself: p.getClosestPoints(distance=0) per self link pair and per (moving link, obstacle) pair
    (pybullet_tools/utils.py:3390-3441, 3996-4014) against the hull oracle that pins the CUDA pre-pass.  Configurations
    whose oracle clearance is within 1e-4 m of the boundary are skipped (Bullet's GJK tolerance)."""
    path = os.environ.get("PMP_IIWA_URDF")
    if not path or not os.path.exists(path):
        pytest.skip("PMP_IIWA_URDF not set")
    from oracle import edgecheck_oracle as O
    from plant_motion_planning_b200.model.tables import default_obstacles
    from plant_motion_planning_b200.robots import load_iiwa14
    model = load_iiwa14()
    body = pybullet.loadURDF(path, useFixedBase=True, physicsClientId=client)
    joints = [j for j in range(pybullet.getNumJoints(body, physicsClientId=client))
              if pybullet.getJointInfo(body, j, physicsClientId=client)[2] != pybullet.JOINT_FIXED]
    obstacles = [o for o in default_obstacles() if o.kind == 0]          # the floor and the block (boxes)
    fixed = []
    for o in obstacles:
        shape = pybullet.createCollisionShape(pybullet.GEOM_BOX, halfExtents=[float(v) for v in o.half], physicsClientId=client)
        from plant_motion_planning_b200.model.transforms import matrix_to_quat
        fixed.append(pybullet.createMultiBody(0, shape, basePosition=[float(v) for v in o.t],
                                              baseOrientation=[float(v) for v in matrix_to_quat(o.R)], physicsClientId=client))
    n_links = pybullet.getNumJoints(body, physicsClientId=client)
    rng = np.random.RandomState(1)
    Q = rng.uniform(model.lower, model.upper, size=(200, model.dof))
    clear = O.hull_clearance(model, obstacles, Q)
    checked = 0
    for q, c in zip(Q, clear):
        if abs(c) < 1e-4:
            continue
        for j, v in zip(joints, q):
            pybullet.resetJointState(body, j, float(v), physicsClientId=client)
        hit = False
        for a in range(n_links):
            for b in range(a + 2, n_links):                              # non-adjacent links of the serial chain
                hit |= len(pybullet.getClosestPoints(body, body, 0.0, a, b, physicsClientId=client)) > 0
            for f in fixed:
                hit |= len(pybullet.getClosestPoints(body, f, 0.0, a, -1, physicsClientId=client)) > 0
        assert hit == (c <= 0.0), (q, c)
        checked += 1
    assert checked > 150
