"""Vectorised plant sampler: its tables equal the scalar packer's on the same draws, its topology
and distributions equal the sequential restatement of create_plant_params, and -m gpu: tables fed straight to the
CUDA library give the results of the PlantWorld route."""
import random

import numpy as np
import pytest

from plant_motion_planning_b200.model import batch_sampler as B
from plant_motion_planning_b200.model.plants import sample_world
from plant_motion_planning_b200.model.tables import pack_worlds

LIMITS = ((0, 0.35), (-0.5, 0))


@pytest.mark.parametrize("args,spacing", [((1, 2, 1), 0.0), ((2, 3, 2), 0.05), ((3, 2, 4), 0.0), ((0, 1, 1), 0.0)])
def test_tables_equal_the_scalar_packer_on_the_same_draws(args, spacing):
    d = B.draw_plants(40, *args, LIMITS, seed=3, stem_base_spacing=spacing)
    wt = B.tables_from_draws(d)
    ref = pack_worlds(B.worlds_from_draws(d, range(40)))
    assert wt.stems.shape == ref.stems.shape and wt.n_slots == ref.n_slots
    assert np.array_equal(wt.stems[..., :24], ref.stems[..., :24]) and np.array_equal(wt.stems[..., 27:], ref.stems[..., 27:])
    if args[0] > 0:
        assert np.abs(wt.stems[..., 28:30]).max() > 0.01, "the rest pose of a branched plant must carry gravity sag"
    assert np.array_equal(wt.stems.view(np.int32)[..., 24:27], ref.stems.view(np.int32)[..., 24:27])   # slots, flags
    assert np.array_equal(wt.boxes, ref.boxes) and np.array_equal(wt.n_stems, ref.n_stems)
    # same topology as the sequential restatement of create_plant_params
    np.random.seed(0)
    random.seed(0)
    w = sample_world(*args, LIMITS, stem_base_spacing=spacing)
    assert [s.parent for s in w.stems] == list(d.parent) and [s.is_main for s in w.stems] == list(d.is_main)


def test_branch_placement_rule_and_ranges():
    d = B.draw_plants(5000, 3, 2, 1, LIMITS, seed=7)
    kinds = B._topology(3, 2, 1)[2]
    sf = 0.25
    for main in {ref for kind, ref in kinds if kind == "branch"}:
        ks = [k for k, (kind, ref) in enumerate(kinds) if kind == "branch" and ref == main]
        for i, a in enumerate(ks):
            x, y = d.xyz[:, a, 0], d.xyz[:, a, 1]
            assert np.all((np.abs(np.abs(x) - sf * 0.2) < 1e-12) | (np.abs(np.abs(y) - sf * 0.2) < 1e-12))
            assert np.all((d.xyz[:, a, 2] >= 0) & (d.xyz[:, a, 2] <= 2 * d.half_height[:, main]))
            for b in ks[:i]:                                   # rand_gen...:196-213: no two branches within 0.05 m in x AND y
                assert not np.any((np.abs(x - d.xyz[:, b, 0]) < sf * 0.2) & (np.abs(y - d.xyz[:, b, 1]) < sf * 0.2))
    assert d.half_height.min() >= sf * 0.7 and d.half_height.max() <= sf * 1.0
    assert d.base_half.min() >= sf * 0.07 and d.base_half.max() <= sf * 0.15
    assert np.all((d.base_xy[:, 0] >= 0) & (d.base_xy[:, 0] <= 0.35) & (d.base_xy[:, 1] >= -0.5) & (d.base_xy[:, 1] <= 0))


def test_distributions_match_the_sequential_sampler():
    worlds = []
    for s in range(1500):
        np.random.seed(10000 + s)
        random.seed(10000 + s)
        worlds.append(sample_world(2, 2, 1, LIMITS, stem_base_spacing=0.05))
    seq = pack_worlds(worlds).stems
    vec = B.sample_world_tables(1500, 2, 2, 1, LIMITS, seed=9, stem_base_spacing=0.05).stems
    for col in (0, 4, 8, 9, 10, 11, 13, 14, 16, 17, 18, 23, 28, 29, 32, 40, 41, 42, 43, 44, 46):   # rest frames, box, observer, sag, origins
        a, b = seq[..., col].astype(np.float64), vec[..., col].astype(np.float64)
        scale = max(a.std(), 1e-3)
        # 15 000 samples per column; the sag-derived columns are correlated within a plant (1500 independent plants)
        k = 4.0 if col in (23, 28, 29) else 1.0
        assert abs(a.mean() - b.mean()) <= k * 0.03 * scale + 1e-6, col
        assert abs(a.std() - b.std()) <= k * 0.05 * scale + 1e-6, col


def test_capacity_and_bad_arguments():
    with pytest.raises(ValueError, match="stems >"):
        B.draw_plants(2, 3, 3, 4, LIMITS)
    with pytest.raises(ValueError):
        B.draw_plants(2, 1, 2, 0, LIMITS)


@pytest.mark.gpu
def test_gpu_accepts_tables_directly():
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14
    robot = load_iiwa14()
    d = B.draw_plants(96, 1, 2, 1, LIMITS, seed=5)
    wt = B.tables_from_draws(d)
    a = EdgeChecker(robot, wt, EdgeCheckConfig())
    b = EdgeChecker(robot, B.worlds_from_draws(d, range(96)), EdgeCheckConfig())
    rng = np.random.RandomState(0)
    Q = torch.from_numpy(rng.uniform(robot.lower, robot.upper, size=(512, 7)).astype(np.float32))
    ra, rb = a.check_configs(Q), b.check_configs(Q)
    assert torch.equal(ra.flags, rb.flags) and torch.equal(ra.deflection, rb.deflection)
    assert (ra.flags & 2).any()
    # a large world count: 4096 sampled worlds read through L1
    big = EdgeChecker(robot, B.sample_world_tables(4096, 1, 2, 1, LIMITS, seed=1), EdgeCheckConfig())
    out = big.validate_edges(Q[:64], Q[64:128], first_hit_only=True)
    sub = EdgeChecker(robot, B.tables_from_draws(B.draw_plants(4096, 1, 2, 1, LIMITS, seed=1)), EdgeCheckConfig())
    assert torch.equal(out.first_hit, sub.validate_edges(Q[:64], Q[64:128]).first_hit)
    with pytest.raises(ValueError):
        a.set_config(stem_margin=0.002)


# ---- device sampler (csrc/pmp_sampler.cuh): same fixed block of uniforms per world -> same tables ---------------------
def test_counter_based_uniforms_and_their_host_mirror():
    nu = B.uniforms_per_world(6)
    U = B.splitmix_uniforms(11, 500, nu)
    assert U.shape == (500, nu) and U.min() >= 0.0 and U.max() < 1.0 and abs(U.mean() - 0.5) < 0.01
    assert np.array_equal(U, B.splitmix_uniforms(11, 500, nu)) and not np.array_equal(U, B.splitmix_uniforms(12, 500, nu))
    # a world's draws do not depend on how many worlds are drawn with it
    assert np.array_equal(U[:100], B.splitmix_uniforms(11, 100, nu))
    d = B.draws_from_uniforms(U, 1, 2, 1, LIMITS)
    wt = B.tables_from_draws(d)
    ref = pack_worlds(B.worlds_from_draws(d, range(50)))
    assert np.array_equal(wt.stems[:50, :, :24], ref.stems[..., :24]) and np.array_equal(wt.stems[:50, :, 27:], ref.stems[..., 27:])
    assert np.array_equal(wt.stems[:50].view(np.int32)[..., 24:27], ref.stems.view(np.int32)[..., 24:27])
    assert np.array_equal(wt.boxes[:50], ref.boxes)
    # ranges of create_plant_params (rand_gen_plants_programmatically_main.py:32, 84-86)
    assert d.half_height.min() >= 0.175 and d.half_height.max() <= 0.25
    assert d.base_xy[:, 0].min() >= 0 and d.base_xy[:, 0].max() <= 0.35 and d.base_xy[:, 1].min() >= -0.5
    # two branches on one main stem never overlap (rand_gen...:196-213), with at most SAMPLER_TRIES attempts
    d2 = B.draws_from_uniforms(B.splitmix_uniforms(5, 2000, B.uniforms_per_world(7)), 2, 1, 2, LIMITS)
    dx = np.abs(d2.xyz[:, 1, 0] - d2.xyz[:, 4, 0]); dy = np.abs(d2.xyz[:, 1, 1] - d2.xyz[:, 4, 1])
    assert not np.any((dx < 0.05) & (dy < 0.05))
    topo = B.sampler_topology(1, 2, 1)
    assert topo["parent"] == [-1, 0, 1, 0, 3, 4] and topo["kind"] == [0, 2, 3, 1, 2, 3] and topo["n_slots"] == 2
    with pytest.raises(ValueError):
        B.sampler_topology(3, 2, 4)          # 32 stems: host sampler only


def _tables_close(dev, host, min_equal=0.99):
    """float fields within one float32 ulp-ish of the float64 host result (and almost all of them bit-identical),
    integer fields identical."""
    a, b = dev.stems, host.stems
    assert a.shape == b.shape
    assert np.array_equal(a.view(np.int32)[..., 24:27], b.view(np.int32)[..., 24:27])
    fa = np.concatenate([a[..., :24], a[..., 27:]], axis=-1).astype(np.float64)
    fb = np.concatenate([b[..., :24], b[..., 27:]], axis=-1).astype(np.float64)
    err = np.abs(fa - fb)
    assert err.max() <= 2.5e-7, err.max()
    assert np.mean(fa == fb) > min_equal, np.mean(fa == fb)
    assert np.abs(dev.boxes.astype(np.float64) - host.boxes).max() <= 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("args,n,same", [((1, 2, 1), 2048, 0.999), ((2, 1, 2), 512, 0.99), ((0, 3, 1), 256, 0.9)])
def test_device_sampler_equals_the_host_packer_on_the_same_draws(args, n, same):
    """`same`: fraction of float fields that must be bit-identical (the rest differ in the last float32 bit, or are
    rounding residues ~1e-17 of quantities that are zero, e.g. the sag of an exactly vertical stem)."""
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14
    from plant_motion_planning_b200.workloads import synthetic_worlds
    m = load_iiwa14()
    chk = EdgeChecker(m, synthetic_worlds(1), EdgeCheckConfig())
    P = B.sampler_topology(*args)["n_stems"]
    U = np.random.default_rng(1).random((n, B.uniforms_per_world(P)))
    dev = chk.sample_worlds_on_device(n, *args, plant_pos_xy_limits=LIMITS, uniforms=U)
    host = B.tables_from_draws(B.draws_from_uniforms(U, *args, LIMITS))
    assert dev.n_slots == host.n_slots and chk.num_worlds == n
    _tables_close(dev, host, same)
    # the on-device generator is the splitmix mirror
    dev2 = chk.sample_worlds_on_device(n, *args, plant_pos_xy_limits=LIMITS, seed=77)
    host2 = B.tables_from_draws(B.draws_from_uniforms(B.splitmix_uniforms(77, n, B.uniforms_per_world(P)), *args, LIMITS))
    _tables_close(dev2, host2, same)
    # without gravity sag the records carry th0 = 0
    chk0 = EdgeChecker(m, synthetic_worlds(1), EdgeCheckConfig(gravity_sag=False))
    dev0 = chk0.sample_worlds_on_device(64, *args, plant_pos_xy_limits=LIMITS, uniforms=U[:64])
    assert np.all(dev0.stems[..., 28:30] == 0)
    # (the rest-pose observer value is then a rounding residue ~1e-17 on both sides, not bit-identical)
    _tables_close(dev0, B.tables_from_draws(B.draws_from_uniforms(U[:64], *args, LIMITS), gravity_sag=False), min_equal=0.9)


@pytest.mark.gpu
def test_large_world_sweep_never_leaves_the_device():
    """W = 65 536 worlds drawn and consumed on the device; a slice of them, downloaded and fed back through the host
    route, gives bit-identical per-world results."""
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14
    from plant_motion_planning_b200.model.tables import WorldTables
    from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds
    m = load_iiwa14()
    W = 1 << 16
    chk = EdgeChecker(m, synthetic_worlds(1), EdgeCheckConfig(max_steps=32))
    wt = chk.sample_worlds_on_device(W, 1, 2, 1, plant_pos_xy_limits=LIMITS, seed=3)
    assert chk.num_worlds == W and wt._stems is None                     # nothing was downloaded
    q0, q1 = synthetic_edges(m, 16, seed=2)
    out = chk.validate_edges(torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda())
    md = out.max_deflection.cpu().numpy()
    assert md.shape == (16, W) and np.isfinite(md).all() and md.max() > 0.3
    sl = slice(1000, 1064)
    sub = WorldTables(64, 6, 1, wt.n_slots, np.full(64, 6, np.int32), wt.stems[sl].copy(), np.ones(64, np.int32), wt.boxes[sl].copy())
    ref = EdgeChecker(m, sub, EdgeCheckConfig(max_steps=32)).validate_edges(torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda())
    assert np.array_equal(ref.max_deflection.cpu().numpy(), md[:, sl])
    assert np.array_equal(ref.over_limit.cpu().numpy(), out.over_limit.cpu().numpy()[:, sl])
