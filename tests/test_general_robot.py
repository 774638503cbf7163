"""The general robot path: 9-DOF chain with mixed axes, a continuous joint, a welded mid-chain link, a frozen finger,
20 spheres (A_PAD = 32) against deep plants (21 stems, 3 live parent frames -> NSLOT = 4)."""
import numpy as np
import pytest

from oracle import edgecheck_oracle as O
from oracle.c_oracle import COracle
from plant_motion_planning_b200.model.tables import assign_slots, pack_robot
from plant_motion_planning_b200.workloads import synthetic_worlds
from synthetic_robot import synthetic_robot


@pytest.fixture(scope="module")
def setup():
    m = synthetic_robot()
    worlds = synthetic_worlds(3, seed_base=77, branches=2, vert_stems=3, extensions=2)
    rng = np.random.RandomState(0)
    lo = np.where(m.circular, -np.pi, m.lower)
    hi = np.where(m.circular, np.pi, m.upper)
    Q = rng.uniform(lo, hi, size=(1024, m.dof)).astype(np.float32)
    return m, worlds, Q


def test_compile_invariants(setup):
    m, worlds, _ = setup
    assert m.dof == 9 and m.num_spheres == 20 and m.num_links == 11
    assert list(m.link_joint) == [0, 1, 2, 3, 4, 5, 5, 6, 7, 8, 8]       # welded sensor rides joint 5, tip rides joint 8
    assert m.circular.tolist() == [False] * 4 + [True] + [False] * 4
    assert np.all(np.diff(m.sph_joint) >= 0)
    t = pack_robot(m)
    assert t.n_fixed == 11 and t.circular[4] == 1      # floor, block, 8 block-corner points, the robot's base box
    assert [w.num_stems for w in worlds] == [21, 21, 21]
    assert max(assign_slots(w)[2] for w in worlds) == 3
    # adjacent links are never paired; spheres of one link never pair with each other
    for a in range(20):
        for b in range(a + 1, 20):
            if m.self_pair_mask[a, b]:
                assert abs(int(m.sph_link[a]) - int(m.sph_link[b])) >= 1 and m.sph_link[a] != m.sph_link[b]


def test_numpy_and_c_oracles_agree(setup):
    m, worlds, Q = setup
    c = COracle(m, worlds).check_configs(Q[:300])
    r = O.check_configs(m, worlds, Q[:300].astype(np.float64), O.OracleParams())
    assert ((c["flags"] & 1 > 0) != r.rigid).mean() < 5e-3
    assert np.percentile(np.abs(c["deflection"] - r.deflection), 99.9) < 1e-4
    assert np.abs(c["ee"] - r.ee).max() < 1e-6


def test_frozen_finger_and_welded_link_in_fk(setup):
    m, _, Q = setup
    pos, quat, com = O.link_poses(m, Q[:8].astype(np.float64))
    R, t = O.joint_frames(m, Q[:8].astype(np.float64))
    tip = m.link_names.index("tip")
    # tip frame = joint-8 frame . origin(0, 0.02, 0.1) . Rx(0.4)
    want = np.einsum("bij,j->bi", R[:, 8], np.array([0.0, 0.02, 0.1])) + t[:, 8]
    assert np.abs(pos[:, tip] - want).max() < 1e-12
    from plant_motion_planning_b200.model.transforms import quat_to_matrix, rot_x
    for b in range(8):
        assert np.abs(quat_to_matrix(quat[b, tip]) - R[b, 8] @ rot_x(0.4)).max() < 1e-12


@pytest.mark.gpu
def test_gpu_parity_general_robot(setup):
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    m, worlds, Q = setup
    chk = EdgeChecker(m, worlds, EdgeCheckConfig())
    g = chk.check_configs(torch.from_numpy(Q))
    info = chk.launch_info()
    assert info["n_sphere_pad"] == 20 and info["n_slot_pad"] == 4     # 20 spheres: the A = 20 instantiation
    c = COracle(m, worlds).check_configs(Q)
    fl = g.flags.cpu().numpy()
    Qd = Q.astype(np.float64)
    robust = O.rigid_hit(m, O.default_obstacles(), Qd, inflate=-1e-5) == O.rigid_hit(m, O.default_obstacles(), Qd, inflate=1e-5)
    assert np.array_equal((fl & 1 > 0)[robust], ((c["flags"] & 1) > 0)[robust])
    near = np.any(np.abs(c["deflection"] - 0.30) < 1e-3, axis=2)
    assert np.array_equal((fl & 2 > 0)[~near], ((c["flags"] & 2) > 0)[~near])
    assert np.percentile(np.abs(g.deflection.cpu().numpy() - c["deflection"]), 99.9) <= 1e-4
    assert (c["deflection"] > 0).any()
    pos, quat, com = chk.fk(torch.from_numpy(Q[:256]))
    opos, oquat, ocom = O.link_poses(m, Qd[:256])
    assert np.abs(pos.cpu().numpy() - opos).max() <= 1e-5 and np.abs(com.cpu().numpy() - ocom).max() <= 1e-5
    qg = quat.cpu().numpy()
    assert np.abs(qg * np.sign((qg * oquat).sum(-1, keepdims=True)) - oquat).max() <= 1e-5
    # edges: the continuous joint interpolates along the wrapped difference
    rng = np.random.RandomState(5)
    q0, q1 = Q[:200].copy(), Q[200:400].copy()
    q0[:, 4], q1[:, 4] = 3.0, -3.0                     # short way round is +0.283 rad, not -6 rad
    out = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), want_masks=True)
    ce = COracle(m, worlds).validate_edges(q0, q1, want_masks=True)
    assert np.array_equal(out.n_steps.cpu().numpy(), ce["n_steps"])
    same = np.all(out.rigid_mask.cpu().numpy().view(np.uint32) == ce["rigid_mask"], axis=(1, 2)) & \
        np.all(out.exceed_mask.cpu().numpy().view(np.uint32) == ce["exceed_mask"], axis=(1, 2))
    assert same.mean() > 0.97
    assert np.array_equal(out.first_hit.cpu().numpy()[same], ce["first_hit"][same])
    assert np.abs(out.ee_len.cpu().numpy() - ce["ee_len"]).max() <= 1e-4
    for e in range(5):
        n = O.num_steps(m, q0[e].astype(np.float64), q1[e].astype(np.float64), np.radians(3))
        assert n == ce["n_steps"][e]


@pytest.mark.gpu
def test_padding_spheres_and_far_stems_in_both_kernel_shapes():
    """A 14-sphere arm runs in the A = 16 instantiation with two padding spheres, and two of the four worlds stand
    5 m from the arm: every sphere is more than a metre outside those stems' boxes, where the contact sums saturate their
    distance components (pmp_outside).  Neither may register: production and dense launches agree bit for bit and
    match the oracle, far worlds stay at their rest observers."""
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.model.plants import sample_world
    import random
    m = synthetic_robot(n_joints=6, spheres_per_link=2)
    assert m.num_spheres == 14
    near, far_x, far_y = ((0.0, 0.35), (-0.5, 0.0)), ((5.0, 5.0), (-0.5, 0.0)), ((0.0, 0.35), (5.0, 5.0))
    worlds = [sample_world(1, 2, 1, lim, np_random=np.random.RandomState(5 + k), py_random=random.Random(5 + k))
              for k, lim in enumerate([near, far_x, near, far_y])]
    rng = np.random.RandomState(2)
    lo, hi = np.where(m.circular, -np.pi, m.lower), np.where(m.circular, np.pi, m.upper)
    Q = rng.uniform(lo, hi, size=(2048, m.dof)).astype(np.float32)
    chk = EdgeChecker(m, worlds, EdgeCheckConfig(max_steps=64))
    g = chk.check_configs(torch.from_numpy(Q))
    assert chk.launch_info()["n_sphere_pad"] == 16
    c = COracle(m, worlds).check_configs(Q)
    d = g.deflection.cpu().numpy()
    assert np.percentile(np.abs(d - c["deflection"]), 99.9) <= 1e-4
    rest = chk.check_configs(torch.from_numpy(Q[:1] * 0 + np.float32(0.0))).deflection.cpu().numpy()
    for far in (1, 3):                                   # nothing reaches a plant 5 m away
        assert np.array_equal(d[:, far], np.broadcast_to(rest[0, far], d[:, far].shape))
        assert np.abs(c["deflection"][:, far] - d[:, far]).max() <= 1e-5
    assert (np.abs(d[:, 0] - rest[0, 0]) > 1e-3).any()   # ... and the near ones are pushed
    a, b = torch.from_numpy(Q[:1024]), torch.from_numpy(Q[1024:])
    r1 = chk.validate_edges(a, b, want_masks=True)
    chk.set_config(dense=True)
    r2 = chk.validate_edges(a, b, want_masks=True)
    chk.set_config(dense=False)
    assert torch.equal(r1.first_hit, r2.first_hit) and torch.equal(r1.exceed_mask, r2.exceed_mask)
    assert torch.equal(r1.rigid_mask, r2.rigid_mask)
    assert (r1.max_deflection - r2.max_deflection).abs().max().item() <= 2e-6
    assert torch.equal(r1.max_deflection[:, 1], r2.max_deflection[:, 1])
    chk.close()


@pytest.mark.gpu
def test_radius_and_margin_limits_are_enforced(setup):
    """pmp_set_robot / pmp_set_worlds reject sphere radii above 0.5 m and stem margins above 0.4 m (their sum must stay
    below the metre at which the contact sums saturate)."""
    import dataclasses
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    m, worlds, _ = setup
    big = dataclasses.replace(m, sph_radius=np.where(np.arange(m.num_spheres) == 3, 0.6, m.sph_radius))
    with pytest.raises(RuntimeError, match="sph_radius"):
        EdgeChecker(big, worlds, EdgeCheckConfig())
    with pytest.raises(RuntimeError, match="margin"):
        EdgeChecker(m, worlds, EdgeCheckConfig(stem_margin=0.45))
