"""Maximum sizes: a 16-DOF arm (PMP_MAX_DOF) against plants of 32 stems (PMP_MAX_STEMS); one more of anything is
refused with a clear error.  CPU: packers + the two oracles; -m gpu: the CUDA path against the oracle at the caps."""
import numpy as np
import pytest

from oracle import edgecheck_oracle as O
from oracle.c_oracle import COracle
from plant_motion_planning_b200.model.tables import assign_slots, pack_robot, pack_worlds
from plant_motion_planning_b200.workloads import synthetic_worlds
from synthetic_robot import synthetic_robot


@pytest.fixture(scope="module")
def big():
    m = synthetic_robot(n_joints=16, spheres_per_link=1)
    worlds = synthetic_worlds(2, seed_base=5, branches=3, vert_stems=2, extensions=4)
    rng = np.random.RandomState(0)
    lo = np.where(m.circular, -np.pi, m.lower)
    hi = np.where(m.circular, np.pi, m.upper)
    Q = (0.35 * rng.uniform(lo, hi, size=(512, m.dof))).astype(np.float32)      # a 2.5 m arm: keep it near the plants
    return m, worlds, Q


def test_tables_at_the_caps(big):
    m, worlds, _ = big
    assert m.dof == 16 and m.num_spheres == 18
    assert [w.num_stems for w in worlds] == [32, 32]
    assert pack_worlds(worlds).stems.shape == (2, 32, 48)
    assert max(assign_slots(w)[2] for w in worlds) <= 4
    pack_robot(m)
    with pytest.raises(ValueError, match="dof 17"):
        pack_robot(synthetic_robot(n_joints=17, spheres_per_link=1, prune=False))
    with pytest.raises(ValueError, match="34 spheres"):
        pack_robot(synthetic_robot(n_joints=16, spheres_per_link=2, prune=False))
    with pytest.raises(ValueError, match="48 stems"):
        pack_worlds(synthetic_worlds(1, seed_base=5, branches=3, vert_stems=3, extensions=4))


def test_oracles_agree_at_the_caps(big):
    m, worlds, Q = big
    c = COracle(m, worlds).check_configs(Q[:128])
    r = O.check_configs(m, worlds, Q[:128].astype(np.float64), O.OracleParams())
    assert ((c["flags"] & 1 > 0) != r.rigid).mean() < 5e-3
    assert np.percentile(np.abs(c["deflection"] - r.deflection), 99.5) < 1e-4
    assert np.abs(c["ee"] - r.ee).max() < 1e-6
    assert (r.deflection > 0).any()


@pytest.mark.gpu
def test_gpu_parity_at_the_caps(big):
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    m, worlds, Q = big
    chk = EdgeChecker(m, worlds, EdgeCheckConfig())
    g = chk.check_configs(torch.from_numpy(Q))
    info = chk.launch_info()
    assert info["n_sphere_pad"] == 18 and info["n_slot_pad"] == 4      # 18 spheres: the A = 18 instantiation
    c = COracle(m, worlds).check_configs(Q)
    fl = g.flags.cpu().numpy()
    Qd = Q.astype(np.float64)
    obs = O.default_obstacles()
    robust = O.rigid_hit(m, obs, Qd, inflate=-1e-5) == O.rigid_hit(m, obs, Qd, inflate=1e-5)
    assert np.array_equal((fl & 1 > 0)[robust], ((c["flags"] & 1) > 0)[robust])
    near = np.any(np.abs(c["deflection"] - 0.30) < 1e-3, axis=2)
    assert np.array_equal((fl & 2 > 0)[~near], ((c["flags"] & 2) > 0)[~near])
    assert np.percentile(np.abs(g.deflection.cpu().numpy() - c["deflection"]), 99.5) <= 1e-4
    pos, quat, com = chk.fk(torch.from_numpy(Q[:128]))
    opos, _, ocom = O.link_poses(m, Qd[:128])
    assert np.abs(pos.cpu().numpy() - opos).max() <= 2e-5 and np.abs(com.cpu().numpy() - ocom).max() <= 2e-5
    q0, q1 = Q[:128], Q[128:256]
    out = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), want_masks=True)
    ce = COracle(m, worlds).validate_edges(q0, q1, want_masks=True)
    assert np.array_equal(out.n_steps.cpu().numpy(), ce["n_steps"])
    same = np.all(out.rigid_mask.cpu().numpy().view(np.uint32) == ce["rigid_mask"], axis=(1, 2)) & \
        np.all(out.exceed_mask.cpu().numpy().view(np.uint32) == ce["exceed_mask"], axis=(1, 2))
    assert same.mean() > 0.95
    assert np.array_equal(out.first_hit.cpu().numpy()[same], ce["first_hit"][same])
    fast = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), first_hit_only=True)
    assert torch.equal(fast.first_hit, out.first_hit)
    # the A = 32 instantiation: a 15-joint arm with two spheres per link (32 spheres, the table's capacity)
    m32 = synthetic_robot(n_joints=15, spheres_per_link=2)
    assert m32.num_spheres == 32
    Q32 = np.random.RandomState(9).uniform(np.where(m32.circular, -np.pi, m32.lower), np.where(m32.circular, np.pi, m32.upper),
                                           size=(512, 15)).astype(np.float32)
    chk32 = EdgeChecker(m32, worlds, EdgeCheckConfig())
    g32 = chk32.check_configs(torch.from_numpy(Q32))
    assert chk32.launch_info()["n_sphere_pad"] == 32
    c32 = COracle(m32, worlds).check_configs(Q32)
    rob32 = O.rigid_hit(m32, obs, Q32.astype(np.float64), inflate=-1e-5) == O.rigid_hit(m32, obs, Q32.astype(np.float64), inflate=1e-5)
    f32 = g32.flags.cpu().numpy()
    assert np.array_equal((f32 & 1 > 0)[rob32], ((c32["flags"] & 1) > 0)[rob32])
    assert np.percentile(np.abs(g32.deflection.cpu().numpy() - c32["deflection"]), 99.5) <= 1e-4
    chk32.close()
    # nearest-node search on the DP = 16 path of a 16-DOF arm with a continuous joint
    idx, dist = chk.nearest(torch.from_numpy(Q[:300]), torch.from_numpy(Q[300:364]))
    oi, od = O.nearest(m, Q[:300], Q[300:364])
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
