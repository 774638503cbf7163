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
        worlds.append(sample_world(2, 3, 2, LIMITS, stem_base_spacing=0.05))
    seq = pack_worlds(worlds).stems
    vec = B.sample_world_tables(1500, 2, 3, 2, LIMITS, seed=9, stem_base_spacing=0.05).stems
    for col in (0, 4, 8, 9, 10, 11, 13, 14, 16, 17, 18, 23, 28, 29, 32, 40, 41, 42, 43, 44, 46):   # rest frames, box, observer, sag, origins
        a, b = seq[..., col].astype(np.float64), vec[..., col].astype(np.float64)
        scale = max(a.std(), 1e-3)
        # 31 500 samples per column; the sag-derived columns are correlated within a plant (1500 independent plants)
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
