"""Parity of the CUDA path (through the C ABI) with the oracle on the same seeded inputs.   -m gpu

Bars (BASELINE.json north_star):
  FK link poses ............. <= 1e-5 m / 1e-5 (quaternion components), float32 vs float64
  collision booleans ........ identical except where the oracle itself is within EPS_M of a decision boundary
  first-violation step ...... identical on edges whose steps are all outside the bands
  plant deflection .......... |gpu - oracle| <= 1e-4 rad for >= 99.9 % of (config, world, stem), <= 1e-3 for
                              >= 99.99 %; the remainder are deep-penetration configurations where the K-step
                              solve is ill-conditioned (reported)
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPS_M = 1e-5        # metres: band around rigid contact / joint limits
EPS_RAD = 1e-3      # radians: band around the deflection limit

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14
    from plant_motion_planning_b200.workloads import synthetic_worlds
    model = load_iiwa14()
    worlds = synthetic_worlds(8)
    chk = EdgeChecker(model, worlds, EdgeCheckConfig())
    rng = np.random.RandomState(0)
    Q = rng.uniform(model.lower, model.upper, size=(4096, 7)).astype(np.float32)
    return {"torch": torch, "model": model, "worlds": worlds, "chk": chk, "Q": Q}


def test_library_is_the_cuda_one(ctx):
    info_before = ctx["chk"].launch_info()["kernel_launches"]
    ctx["chk"].check_configs(ctx["torch"].from_numpy(ctx["Q"][:8]))
    # two of this library's kernels: the hull pre-pass (exact rigid test) and the configuration kernel
    assert ctx["chk"].launch_info()["kernel_launches"] == info_before + 2
    with open("/proc/self/maps") as fp:
        assert "libpmp_edgecheck.so" in fp.read()


def test_fk_within_1e5(ctx):
    from oracle import edgecheck_oracle as O
    torch = ctx["torch"]
    pos, quat, com = ctx["chk"].fk(torch.from_numpy(ctx["Q"][:1024]))
    opos, oquat, ocom = O.link_poses(ctx["model"], ctx["Q"][:1024].astype(np.float64))
    assert np.abs(pos.cpu().numpy() - opos).max() <= 1e-5
    assert np.abs(com.cpu().numpy() - ocom).max() <= 1e-5
    qg = quat.cpu().numpy()
    sgn = np.sign(np.sum(qg * oquat, axis=-1, keepdims=True))
    assert np.abs(qg * sgn - oquat).max() <= 1e-5
    # end-effector link = PyBullet link 9 (kuka_primitives.py:478)
    ee = ctx["chk"].check_configs(torch.from_numpy(ctx["Q"][:1024])).ee.cpu().numpy()
    assert np.abs(ee - ocom[:, 9]).max() <= 1e-5


def test_fk_against_reference_transformations_golden(ctx):
    """pmp_fk vs link frames composed with the reference's transformations.py (tests/golden/reference_fk.json)."""
    with open(os.path.join(G, "reference_fk.json")) as fp:
        d = json.load(fp)
    Q = np.array([c["q"] for c in d["cases"]], dtype=np.float32)
    pos, quat, com = ctx["chk"].fk(ctx["torch"].from_numpy(Q))
    rp = np.array([c["link_pos"] for c in d["cases"]])
    rq = np.array([c["link_quat"] for c in d["cases"]])
    rc = np.array([c["com_pos"] for c in d["cases"]])
    qg = quat.cpu().numpy()
    sg = np.sign((qg * rq).sum(-1, keepdims=True))
    assert np.abs(pos.cpu().numpy() - rp).max() <= 1e-5
    assert np.abs(com.cpu().numpy() - rc).max() <= 1e-5
    assert np.abs(qg * sg - rq).max() <= 1e-5


@pytest.mark.parametrize("mode", ["our_method", "avoid_all", "ignore_all"])
def test_config_flags_and_deflections(ctx, mode):
    from oracle import edgecheck_oracle as O
    from oracle.c_oracle import COracle
    torch, chk, model, worlds, Q = ctx["torch"], ctx["chk"], ctx["model"], ctx["worlds"], ctx["Q"]
    chk.set_config(mode=mode)
    try:
        g = chk.check_configs(torch.from_numpy(Q))
        torch.cuda.synchronize()
    finally:
        chk.set_config(mode="our_method")
    c = COracle(model, worlds, mode=mode).check_configs(Q)
    fl = g.flags.cpu().numpy()
    gr, gx = (fl & 1) > 0, (fl & 2) > 0
    cr, cx = (c["flags"] & 1) > 0, (c["flags"] & 2) > 0
    Qd = Q.astype(np.float64)
    robust = _rigid_robust(model, Qd)
    if mode == "avoid_all":
        # plant contact adds a boundary: |rest penetration| < eps
        for wi, w in enumerate(worlds):
            cen = O.sphere_centers(model, Qd)
            _, _, _, rp = O.plant_equilibrium(w, cen, model.sph_radius, O.OracleParams(), solve=False)
            ok = robust & (np.abs(rp) > EPS_M)
            assert np.array_equal(gr[ok, wi], cr[ok, wi])
    else:
        assert np.array_equal(gr[robust], cr[robust])
    assert (gr != cr).sum() <= 2, "flag differences are confined to the 1e-5 m band, which holds a handful of 32768"
    near = np.any(np.abs(c["deflection"] - 0.30) < EPS_RAD, axis=2)
    assert np.array_equal(gx[~near], cx[~near])
    err = np.abs(g.deflection.cpu().numpy() - c["deflection"])
    assert np.percentile(err, 99.9) <= 1e-4
    assert (err > 1e-3).mean() <= 1e-4
    terr = np.abs(g.theta.cpu().numpy() - c["theta"])
    assert np.percentile(terr, 99.9) <= 1e-4
    if mode == "our_method":
        assert cx.any() and (c["deflection"] > 0).mean() > 0.01, "test inputs must exercise the plant solve"


def _rigid_robust(model, Qd):
    """Configurations whose rigid decision is at least EPS_M away from flipping: hull clearance (exact geometry,
    float64 GJK) and joint limits."""
    from oracle import edgecheck_oracle as O
    clear = O.hull_clearance(model, O.default_obstacles(), Qd)
    lim = np.minimum(Qd - model.lower[None], model.upper[None] - Qd).min(axis=1)
    return (np.abs(clear) > EPS_M) & (np.abs(lim) > EPS_M)


def test_rigid_flags_equal_the_hull_oracle_outside_the_band(ctx):
    """The exact rigid test (hull pre-pass kernel, float32 GJK) against the INDEPENDENT convex-hull oracle
    (oracle/hull_oracle.py: own URDF kinematics, hulls extracted from the reference's meshes, float64 GJK): identical
    booleans for every configuration whose Bullet clearance is more than 1e-5 m from zero; the band itself holds a
    few configurations in 10^5, counted here."""
    from oracle import hull_oracle as H
    torch, chk = ctx["torch"], ctx["chk"]
    rng = np.random.RandomState(42)
    model = ctx["model"]
    Q = rng.uniform(model.lower, model.upper, size=(100000, 7)).astype(np.float32)
    chk.set_config(mode="ignore_all")
    try:
        fl = chk.check_configs(torch.from_numpy(Q), detail=False).flags.cpu().numpy()[:, 0] & 1
    finally:
        chk.set_config(mode="our_method")
    ref = H.HullScene().rigid(Q.astype(np.float64))
    band = np.abs(ref["clearance"]) <= EPS_M
    assert np.array_equal(fl[~band] > 0, ref["hit"][~band])
    assert band.sum() <= 20 and 0.15 < ref["hit"].mean() < 0.25
    # the script's own goal, unmodified, is free; 2 cm deeper into the block it is not
    from plant_motion_planning_b200.scenarios import REFERENCE_GOAL
    g2 = np.array([REFERENCE_GOAL, REFERENCE_GOAL], dtype=np.float32)
    g2[1, 5] += 0.02
    fl = chk.check_configs(torch.from_numpy(g2), detail=False).flags.cpu().numpy()
    assert not (fl[0] & 1).any() and (fl[1] & 1).all()


def test_numpy_oracle_small(ctx):
    """The normative float64 numpy oracle (quaternion observers, acos) on a small sample."""
    from oracle import edgecheck_oracle as O
    torch, chk, model, worlds, Q = ctx["torch"], ctx["chk"], ctx["model"], ctx["worlds"], ctx["Q"][:512]
    g = chk.check_configs(torch.from_numpy(Q))
    r = O.check_configs(model, worlds, Q.astype(np.float64), O.OracleParams(reg_eps=chk.config.reg_eps))
    fl = g.flags.cpu().numpy()
    robust = _rigid_robust(model, Q.astype(np.float64))
    assert np.array_equal((fl & 1 > 0)[robust], r.rigid[robust])
    near = np.any(np.abs(r.deflection - 0.30) < EPS_RAD, axis=2)
    assert np.array_equal((fl & 2 > 0)[~near], r.exceeded[~near])
    err = np.abs(g.deflection.cpu().numpy() - r.deflection)
    assert np.percentile(err, 99.5) <= 1e-4


def _edge_compare(chk, torch, model, worlds, q0, q1, backward, max_steps=320):
    from oracle.c_oracle import COracle
    out = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), backward=backward, want_masks=True,
                             max_steps=max_steps)
    torch.cuda.synchronize()
    c = COracle(model, worlds, reg_eps=chk.config.reg_eps).validate_edges(q0, q1, backward=backward, want_masks=True,
                                                                          max_steps=max_steps)
    assert np.array_equal(out.n_steps.cpu().numpy(), c["n_steps"])            # integer work: bit exact
    gm_r, gm_x = out.rigid_mask.cpu().numpy().view(np.uint32), out.exceed_mask.cpu().numpy().view(np.uint32)
    diff_bits = sum(int(bin(int(v)).count("1")) for v in np.bitwise_xor(gm_r, c["rigid_mask"]).ravel()) + \
        sum(int(bin(int(v)).count("1")) for v in np.bitwise_xor(gm_x, c["exceed_mask"]).ravel())
    total_bits = int(c["n_steps"].clip(max=max_steps).sum()) * len(worlds) * 2
    assert diff_bits <= max(2, 2e-4 * total_bits), (diff_bits, total_bits)
    same = np.all(gm_r == c["rigid_mask"], axis=(1, 2)) & np.all(gm_x == c["exceed_mask"], axis=(1, 2))
    fh = out.first_hit.cpu().numpy()
    assert np.array_equal(fh[same], c["first_hit"][same])
    # every differing bit must sit inside a decision band of the float64 oracle (1e-5 m of rigid contact or of a joint
    # limit, 1e-3 rad of the deflection limit): checked step by step on the edges that differ
    _assert_differences_in_band(chk, model, worlds, q0, q1, backward, max_steps, np.nonzero(~same)[0],
                                gm_r, c["rigid_mask"], gm_x, c["exceed_mask"])
    assert np.abs(out.ee_len.cpu().numpy() - c["ee_len"]).max() <= 1e-4
    # max over up to max_steps x P per-unit values: the per-unit tail (deep penetrations)
    # is picked up by the max, so the edge-level bar is on the 99th percentile
    md = np.abs(out.max_deflection.cpu().numpy() - c["max_deflection"])
    assert np.percentile(md, 99) <= 1e-4
    ov = np.abs(out.over_limit.cpu().numpy() - c["over_limit"])
    assert np.percentile(ov, 99) <= 1e-3
    cost = np.abs(out.cost.cpu().numpy() - c["cost"])
    assert np.percentile(cost, 99) <= 1e-3
    return out, c


def _assert_differences_in_band(chk, model, worlds, q0, q1, backward, max_steps, edges, gm_r, cm_r, gm_x, cm_x):
    from oracle import edgecheck_oracle as O
    from oracle.c_oracle import COracle
    if len(edges) == 0:
        return
    assert len(edges) <= max(4, len(q0) // 50), "too many edges differ for float32-vs-float64 rounding"
    co = COracle(model, worlds, reg_eps=chk.config.reg_eps, deflection_limit=chk.config.deflection_limit,
                 mode=chk.config.mode)
    tail = []
    for e in edges:
        Qe = O.edge_configs(model, q0[e].astype(np.float64), q1[e].astype(np.float64), chk.config.resolution, backward)[:max_steps]
        robust = _rigid_robust(model, Qe)
        defl = co.check_configs(Qe.astype(np.float32))["deflection"]
        near = np.any(np.abs(defl - chk.config.deflection_limit) < EPS_RAD, axis=2)         # [n, W]
        # the ill-conditioned tail of the K-step solve: where the arm is pressed deep into a plant the K damped steps
        # are not a contraction, and a 1e-6 rad change of the INPUT moves the oracle's own answer by more than the
        # band.  Such steps are identified by exactly that -- the float64 oracle disagreeing
        # with itself under 1e-6 rad perturbations -- excluded, and counted.
        rs = np.random.RandomState(0)
        pert = [co.check_configs((Qe + 1e-6 * rs.choice([-1.0, 1.0], size=Qe.shape)).astype(np.float32))["deflection"]
                for _ in range(4)]
        spread = np.max(np.abs(np.stack(pert) - defl[None]), axis=(0, 3))                     # [n, W]
        wide = spread > EPS_RAD
        for w in range(len(worlds)):
            for ch in range(gm_r.shape[2]):
                dr, dx = int(gm_r[e, w, ch]) ^ int(cm_r[e, w, ch]), int(gm_x[e, w, ch]) ^ int(cm_x[e, w, ch])
                for bit in range(32):
                    i = 32 * ch + bit
                    if (dr >> bit) & 1:
                        assert not robust[i], ("rigid bit differs outside the band", e, w, i)
                    if (dx >> bit) & 1 and not near[i, w]:
                        # float32 tail of the K-step solve: the GPU's own deflections of this step must then be within
                        # 1e-2 rad of the oracle's and the limit must lie between the two (or the step is ill-conditioned)
                        if not wide[i, w]:
                            gd = chk.check_configs_host(Qe[i:i + 1].astype(np.float32))["deflection"][0, w]
                            err = np.abs(gd - defl[i, w])
                            lim = chk.config.deflection_limit
                            assert err.max() < 1e-2 and np.any(np.abs(defl[i, w] - lim) <= err + 1e-4), \
                                ("exceeded bit differs outside the band on a well-conditioned step", e, w, i, err.max())
                        tail.append((int(e), w, i))
    units = int(sum(min(int(O.num_steps(model, q0[e].astype(np.float64), q1[e].astype(np.float64), chk.config.resolution)),
                        max_steps) for e in range(len(q0)))) * len(worlds)
    assert len(tail) <= max(2, units // 200000), ("too many exceeded bits in the float32 / ill-conditioned tail", tail[:5], units)


@pytest.mark.parametrize("backward", [False, True])
def test_ragged_edges(ctx, backward):
    from plant_motion_planning_b200.workloads import random_edges
    q0, q1 = random_edges(ctx["model"], 512, seed=3)
    q1[5] = q0[5]                       # zero-length edge: one step
    out, c = _edge_compare(ctx["chk"], ctx["torch"], ctx["model"], ctx["worlds"], q0, q1, backward)
    assert int(out.n_steps[5]) == 1
    assert out.n_steps.max() > 32, "need multi-chunk edges in this test"


def test_long_edges_and_step_capacity(ctx):
    """Edges longer than max_steps are evaluated up to max_steps (documented contract); multi-chunk masks."""
    model = ctx["model"]
    rng = np.random.RandomState(8)
    q0 = rng.uniform(model.lower, model.upper, size=(96, 7)).astype(np.float32)
    q1 = rng.uniform(model.lower, model.upper, size=(96, 7)).astype(np.float32)
    _edge_compare(ctx["chk"], ctx["torch"], model, ctx["worlds"], q0, q1, False, max_steps=320)
    out, c = _edge_compare(ctx["chk"], ctx["torch"], model, ctx["worlds"], q0, q1, False, max_steps=40)
    assert (c["n_steps"] > 40).any()
    assert int(out.first_hit.max()) <= 40


@pytest.mark.parametrize("mode", ["our_method", "avoid_all", "ignore_all"])
@pytest.mark.parametrize("backward", [False, True])
def test_first_hit_only_calls_stop_early_with_identical_answers(ctx, mode, backward):
    """A call that asks only for first_hit / n_steps retires steps and worlds like the reference's loops
    (primitives.py:225-241, env.py:230-238); the answers are bit-identical to the full evaluation, also with the
    on-device gate and with a step capacity."""
    from plant_motion_planning_b200.workloads import random_edges
    torch, chk = ctx["torch"], ctx["chk"]
    q0, q1 = random_edges(ctx["model"], 1024, seed=12)
    q1[3] = q0[3]
    q1[512:] = q0[512:] + 0.25 * (q1[512:] - q0[512:])          # short edges: some stay free in every mode
    a, b = torch.from_numpy(q0), torch.from_numpy(q1)
    seen_hit = seen_free = False
    try:
        for gate in (None, 7):
            chk.set_config(mode=mode, device_gate_seed=gate, deflection_limit=0.10 if gate else 0.30)
            for ms in (320, 40):
                full = chk.validate_edges(a, b, backward=backward, max_steps=ms)
                fast = chk.validate_edges(a, b, backward=backward, max_steps=ms, first_hit_only=True)
                assert fast.max_deflection is None and fast.cost is None
                fh, n = full.first_hit.cpu().numpy(), np.minimum(full.n_steps.cpu().numpy(), ms)
                bad = np.nonzero(fh != fast.first_hit.cpu().numpy())[0]
                assert len(bad) == 0, (gate, ms, bad[:8], fh[bad[:8]], fast.first_hit.cpu().numpy()[bad[:8]], n[bad[:8]])
                assert torch.equal(full.n_steps, fast.n_steps)
                seen_hit |= bool((fh < n).any())
                seen_free |= bool((fh >= n).any())
        assert seen_hit and (seen_free or mode == "avoid_all")     # 8 rigid plants leave no random edge free
        host = chk.validate_edges_host(q0, q1, backward=backward, max_steps=40, first_hit_only=True)
        assert sorted(host) == ["first_hit", "n_steps"] and np.array_equal(host["first_hit"], fast.first_hit.cpu().numpy())
    finally:
        chk.set_config(mode="our_method", device_gate_seed=None, deflection_limit=0.30)


def test_worlds_without_plants(ctx):
    """A world with no stems and no plant base (an empty plot) next to planted ones, and a checker whose only world
    is empty: only the arm's own rigid tests can fire there, deflections are zero."""
    from oracle.c_oracle import COracle
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.model.plants import PlantWorld
    from plant_motion_planning_b200.workloads import random_edges
    torch, model = ctx["torch"], ctx["model"]
    empty = PlantWorld(stems=[], base_boxes=[])
    Q = ctx["Q"][:512]
    for worlds in ([empty, ctx["worlds"][0], empty], [empty]):
        for mode in ("our_method", "avoid_all"):
            chk = EdgeChecker(model, worlds, EdgeCheckConfig(mode=mode))
            g = chk.check_configs(torch.from_numpy(Q))
            c = COracle(model, worlds, mode=mode).check_configs(Q)
            fl = g.flags.cpu().numpy()
            assert np.array_equal(fl[:, 0], fl[:, -1]) and not (fl[:, 0] & 2).any()
            assert (fl != c["flags"]).sum() <= 1
            assert np.abs(g.deflection.cpu().numpy()[:, 0]).max() == 0.0
            q0, q1 = random_edges(model, 64, seed=2)
            out = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1))
            ce = COracle(model, worlds, mode=mode).validate_edges(q0, q1)
            assert np.array_equal(out.n_steps.cpu().numpy(), ce["n_steps"])
            assert (out.first_hit.cpu().numpy() != ce["first_hit"]).sum() <= 1
            assert float(out.max_deflection[:, 0].abs().max()) == 0.0
            fast = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), first_hit_only=True)
            assert torch.equal(fast.first_hit, out.first_hit)
            chk.close()


def test_empty_batch_and_bad_shapes(ctx):
    torch, chk = ctx["torch"], ctx["chk"]
    out = chk.validate_edges(torch.zeros((0, 7)), torch.zeros((0, 7)))
    assert out.first_hit.numel() == 0
    with pytest.raises(ValueError):
        chk.validate_edges(torch.zeros((4, 6)), torch.zeros((4, 6)))
    with pytest.raises(ValueError):
        chk.validate_edges(torch.zeros((4, 7)), torch.zeros((5, 7)))


def test_device_gate_matches_host_recomputation(ctx):
    """validate_edges with the on-device 5 % pass-through (counter-based hash) vs the same rule applied on the host to
    the raw masks; the raw masks themselves must not change."""
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, planner_fns as F
    from plant_motion_planning_b200.workloads import random_edges
    torch, model, worlds = ctx["torch"], ctx["model"], ctx["worlds"]
    q0, q1 = random_edges(model, 2048, seed=17)
    raw = ctx["chk"].validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), want_masks=True)
    rm = raw.rigid_mask.cpu().numpy().view(np.uint32)
    xm = raw.exceed_mask.cpu().numpy().view(np.uint32)
    n = raw.n_steps.cpu().numpy()
    for seed, prob in ((5, 0.95), (6, 0.5), (7, 1.0)):
        chk = EdgeChecker(model, worlds, EdgeCheckConfig(device_gate_seed=seed, gate_probability=prob))
        g = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), want_masks=True)
        assert torch.equal(g.rigid_mask, raw.rigid_mask) and torch.equal(g.exceed_mask, raw.exceed_mask)
        want = F.device_gate_first_hit(n, rm, xm, seed, prob)
        got = g.first_hit.cpu().numpy()
        assert np.array_equal(got, want)
        if prob < 1.0:
            assert (got != raw.first_hit.cpu().numpy()).any(), "the gate must let some exceeded steps through"
        else:
            assert np.array_equal(got, raw.first_hit.cpu().numpy())
    # the gate passes ~5 % of the exceeded (edge, step, world) triples
    h = F.gate_hash(5, np.arange(200000, dtype=np.uint32), 3, 1)
    assert abs((h < np.uint32(int(0.95 * 2 ** 32))).mean() - 0.95) < 0.005


def test_device_gate_does_not_depend_on_sharding(ctx):
    """The on-device gate hashes the GLOBAL edge index (edge_base): shards of a batch -- what ShardedEdgeChecker hands
    every rank -- decide exactly like the unsplit batch, and the chunked host entry point like the device one."""
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.distributed import shard_bounds
    from plant_motion_planning_b200.workloads import random_edges
    torch, model, worlds = ctx["torch"], ctx["model"], ctx["worlds"]
    q0, q1 = random_edges(model, 1537, seed=23)
    chk = EdgeChecker(model, worlds, EdgeCheckConfig(device_gate_seed=9, gate_probability=0.5))
    a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
    whole = chk.validate_edges(a, b).first_hit.cpu().numpy()
    for ws in (2, 3):
        parts = []
        for r in range(ws):
            lo, hi = shard_bounds(len(q0), r, ws)
            parts.append(chk.validate_edges(a[lo:hi], b[lo:hi], edge_base=lo).first_hit.cpu().numpy())
        assert np.array_equal(np.concatenate(parts), whole)
    lo, hi = shard_bounds(len(q0), 1, 2)
    assert not np.array_equal(chk.validate_edges(a[lo:hi], b[lo:hi]).first_hit.cpu().numpy(), whole[lo:hi]), \
        "without edge_base the shard must see other gate draws (else this test proves nothing)"
    assert np.array_equal(chk.validate_edges_host(q0, q1)["first_hit"], whole)


def test_per_kernel_timing_through_the_c_abi(ctx):
    """pmp_set_kernel_timing / pmp_get_kernel_times: what bench.py's roofline leg reads (CUDA events recorded by the
    library on the launching stream, around the pre-pass and around the edge kernel)."""
    from plant_motion_planning_b200.workloads import random_edges
    torch, chk, model = ctx["torch"], ctx["chk"], ctx["model"]
    q0, q1 = random_edges(model, 4096, seed=3)
    a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
    with pytest.raises(RuntimeError):
        chk.kernel_times()                      # timing is off by default
    chk.set_kernel_timing(True)
    try:
        ref = chk.validate_edges(a, b)
        rigid_ms, edge_ms = chk.kernel_times()
        assert 0.0 < rigid_ms < 100.0 and 0.0 < edge_ms < 1000.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chk.validate_edges(a, b)
        e1.record()
        torch.cuda.synchronize()
        r2, k2 = chk.kernel_times()
        assert r2 + k2 <= e0.elapsed_time(e1) * 1.05 + 0.05      # the two kernels are the call
    finally:
        chk.set_kernel_timing(False)
    again = chk.validate_edges(a, b)
    assert torch.equal(again.first_hit, ref.first_hit)


def test_non_finite_inputs_are_rejected_not_crashed(ctx):
    torch, chk = ctx["torch"], ctx["chk"]
    q0 = np.zeros((4, 7), np.float32)
    q1 = np.full((4, 7), 0.2, np.float32)
    q1[1, 2] = np.nan
    q0[2, 0] = np.inf
    q1[3, 5] = 1e30
    out = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), max_steps=64)
    torch.cuda.synchronize()
    fh, n = out.first_hit.cpu().numpy(), out.n_steps.cpu().numpy()
    solo = chk.validate_edges(torch.from_numpy(q0[:1]), torch.from_numpy(q1[:1]), max_steps=64)
    assert n[0] == int(solo.n_steps[0]) >= 1 and fh[0] == int(solo.first_hit[0])     # the finite edge is untouched by its neighbours
    assert n[1] == 1 and fh[1] == 0                   # NaN: one step, rejected
    assert n[2] > 64 and fh[2] == 0                    # inf: infinitely long edge, every configuration rejected
    assert n[3] > 64 and fh[3] < 64                    # absurdly long edge: truncated at max_steps and rejected (limits)
    assert not bool(out.valid[1]) and not bool(out.valid[2]) and not bool(out.valid[3])


def test_limits_violation_is_a_rigid_hit(ctx):
    torch, chk, model = ctx["torch"], ctx["chk"], ctx["model"]
    q = np.zeros((3, 7), np.float32)
    q[:, 1] = 0.6
    q[1, 3] = model.upper[3] + 1e-3
    q[2, 0] = model.lower[0] - 1e-3
    fl = chk.check_configs(torch.from_numpy(q)).flags.cpu().numpy()
    assert not (fl[0] & 1).any() and (fl[1] & 1).all() and (fl[2] & 1).all()


def test_host_api_equals_device_api(ctx):
    from plant_motion_planning_b200.workloads import random_edges
    torch, chk = ctx["torch"], ctx["chk"]
    q0, q1 = random_edges(ctx["model"], 300, seed=21)
    dev = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), want_masks=True)
    host = chk.validate_edges_host(q0, q1, want_masks=True)
    for k, t in (("first_hit", dev.first_hit), ("n_steps", dev.n_steps), ("max_deflection", dev.max_deflection),
                 ("over_limit", dev.over_limit), ("ee_len", dev.ee_len), ("cost", dev.cost)):
        assert np.array_equal(host[k], t.cpu().numpy()), k
    assert np.array_equal(host["rigid_mask"], dev.rigid_mask.cpu().numpy().view(np.uint32))
    # large batches take the chunked two-stream route of the host entry point: same bits, also with the on-device gate
    # (whose hash sees the index of the edge in the caller's batch, not in the chunk)
    from plant_motion_planning_b200.workloads import synthetic_edges
    b0, b1 = synthetic_edges(ctx["model"], (1 << 16) + 37, seed=3, steps=12)
    try:
        chk.set_config(device_gate_seed=11, deflection_limit=0.05, max_steps=32)
        dev = chk.validate_edges(torch.from_numpy(b0), torch.from_numpy(b1), want_masks=True)
        host = chk.validate_edges_host(b0, b1, want_masks=True)
        for k, t in (("first_hit", dev.first_hit), ("n_steps", dev.n_steps), ("max_deflection", dev.max_deflection),
                     ("over_limit", dev.over_limit), ("ee_len", dev.ee_len), ("cost", dev.cost)):
            assert np.array_equal(host[k], t.cpu().numpy()), k
        assert np.array_equal(host["exceed_mask"], dev.exceed_mask.cpu().numpy().view(np.uint32))
        assert host["exceed_mask"].any()
        fast = chk.validate_edges_host(b0, b1, first_hit_only=True)
        assert np.array_equal(fast["first_hit"], host["first_hit"])
    finally:
        chk.set_config(device_gate_seed=None, deflection_limit=0.30, max_steps=320)
    hc = chk.check_configs_host(ctx["Q"][:100])
    dc = chk.check_configs(torch.from_numpy(ctx["Q"][:100]))
    assert np.array_equal(hc["flags"], dc.flags.cpu().numpy())
    assert np.array_equal(hc["deflection"], dc.deflection.cpu().numpy())


# ---------------------------------------------------------------- size-independent properties at full size
@pytest.fixture(scope="module")
def sweep():
    import torch
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14
    from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds
    model = load_iiwa14()
    worlds = synthetic_worlds(64)
    chk = EdgeChecker(model, worlds, EdgeCheckConfig(max_steps=32))
    q0, q1 = synthetic_edges(model, 1 << 17)
    return torch, model, worlds, chk, torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()


def test_dense_and_early_exit_kernels_agree_bitwise(sweep):
    """Skipping provably no-op iterations must not change a single bit (BASELINE sweep shape, 2^17 edges)."""
    torch, model, worlds, chk, a, b = sweep
    r1 = chk.validate_edges(a, b, want_masks=True)
    chk.set_config(dense=True)
    try:
        r2 = chk.validate_edges(a, b, want_masks=True)
    finally:
        chk.set_config(dense=False)
    torch.cuda.synchronize()
    assert torch.equal(r1.first_hit, r2.first_hit) and torch.equal(r1.n_steps, r2.n_steps)
    assert torch.equal(r1.rigid_mask, r2.rigid_mask) and torch.equal(r1.exceed_mask, r2.exceed_mask)
    assert (r1.max_deflection - r2.max_deflection).abs().max().item() <= 2e-6
    assert int((r1.n_steps != 32).sum()) == 0


def test_reversed_edge_visits_the_same_configurations(sweep):
    """forward(q0->q1) and backward(q1->q0) evaluate the same configuration set {q0 + k/n d, k=1..n}: per-world
    max deflection is identical and the violation masks are mirror images."""
    torch, model, worlds, chk, a, b = sweep
    a, b = a[:1 << 14], b[:1 << 14]
    f = chk.validate_edges(a, b, want_masks=True)
    r = chk.validate_edges(b, a, backward=True, want_masks=True)
    torch.cuda.synchronize()
    assert torch.equal(f.n_steps, r.n_steps)
    # the two launches round the interpolated angles differently (1 ulp); deep-penetration configurations amplify
    # that, so the bar is on the distribution, not the maximum
    dd = (f.max_deflection - r.max_deflection).abs()
    assert (dd > 1e-4).float().mean().item() < 1e-3
    assert dd.median().item() <= 1e-6
    fm = f.rigid_mask.cpu().numpy().view(np.uint32)[..., 0]
    rm = r.rigid_mask.cpu().numpy().view(np.uint32)[..., 0]
    rev = np.zeros_like(rm)
    for s in range(32):
        rev |= ((rm >> np.uint32(s)) & np.uint32(1)) << np.uint32(31 - s)
    assert (fm != rev).mean() < 1e-4


def test_batch_composition_and_world_order_do_not_matter(sweep):
    torch, model, worlds, chk, a, b = sweep
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    n = 1 << 13
    full = chk.validate_edges(a[:n], b[:n])
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).cuda()
    sh = chk.validate_edges(a[:n][perm], b[:n][perm])
    assert torch.equal(full.first_hit[perm], sh.first_hit)
    assert torch.equal(full.max_deflection[perm], sh.max_deflection)
    order = list(reversed(range(64)))
    chk2 = EdgeChecker(model, [worlds[i] for i in order], EdgeCheckConfig(max_steps=32))
    r2 = chk2.validate_edges(a[:n], b[:n])
    assert torch.equal(full.first_hit, r2.first_hit)
    assert torch.equal(full.max_deflection, r2.max_deflection.flip(1))
    assert torch.equal(full.over_limit, r2.over_limit.flip(1))


def test_full_sweep_shape_checksum_against_oracle_sample(sweep):
    """2^17 x 32 x 64 units on the GPU; a 4096-edge slice (8.4 M units) re-evaluated by the C oracle: step masks
    identical except inside the decision bands, first violation identical wherever the masks are."""
    from oracle.c_oracle import COracle
    torch, model, worlds, chk, a, b = sweep
    r = chk.validate_edges(a, b, want_masks=True)
    torch.cuda.synchronize()
    idx = np.arange(0, 1 << 17, 32)
    q0, q1 = a.cpu().numpy()[idx], b.cpu().numpy()[idx]
    c = COracle(model, worlds).validate_edges(q0, q1, max_steps=32, want_masks=True)
    gm_r = r.rigid_mask.cpu().numpy().view(np.uint32)[idx]
    gm_x = r.exceed_mask.cpu().numpy().view(np.uint32)[idx]
    same = np.all(gm_r == c["rigid_mask"], axis=(1, 2)) & np.all(gm_x == c["exceed_mask"], axis=(1, 2))
    fh = r.first_hit.cpu().numpy()[idx]
    assert np.array_equal(fh[same], c["first_hit"][same])
    _assert_differences_in_band(chk, model, worlds, q0, q1, False, 32, np.nonzero(~same)[0], gm_r, c["rigid_mask"],
                                gm_x, c["exceed_mask"])
    # max over 32 steps x 6 stems of per-unit values: the per-unit tail (ill-conditioned deep penetrations, DESIGN.md
    # section 6) is picked up by the max
    md = np.abs(r.max_deflection.cpu().numpy()[idx] - c["max_deflection"])
    assert np.percentile(md, 99) <= 1e-4 and np.percentile(md, 99.9) <= 2e-3


def test_many_worlds_tables_outside_shared_memory(ctx):
    """BASELINE config 5 shape (W = 256): the stem tables (295 KB) no longer fit in shared memory and are read
    through L1; results must not depend on where the tables live."""
    from oracle.c_oracle import COracle
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds
    torch, model = ctx["torch"], ctx["model"]
    worlds = synthetic_worlds(256, seed_base=1)
    chk = EdgeChecker(model, worlds, EdgeCheckConfig(max_steps=32))
    q0, q1 = synthetic_edges(model, 4096, seed=1)
    out = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1))
    torch.cuda.synchronize()
    assert chk.launch_info()["worlds_in_smem"] == 0
    idx = np.arange(0, 4096, 64)
    c = COracle(model, worlds).validate_edges(q0[idx], q1[idx], max_steps=32)
    assert (out.first_hit.cpu().numpy()[idx] != c["first_hit"]).sum() <= 1
    md = np.abs(out.max_deflection.cpu().numpy()[idx] - c["max_deflection"])
    assert np.percentile(md, 99) <= 1e-4
    # same worlds, first 64 only, tables in shared memory: identical columns
    chk64 = EdgeChecker(model, worlds[:64], EdgeCheckConfig(max_steps=32))
    o64 = chk64.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1))
    assert chk64.launch_info()["worlds_in_smem"] == 1
    assert torch.equal(o64.max_deflection, out.max_deflection[:, :64])
    assert torch.equal(o64.over_limit, out.over_limit[:, :64])


def test_small_batches_spread_over_sms(ctx):
    torch, chk = ctx["torch"], ctx["chk"]
    from plant_motion_planning_b200.workloads import random_edges
    q0, q1 = random_edges(ctx["model"], 40, seed=2)
    a = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1))
    info = chk.launch_info()
    assert info["grid"] == 40 and info["block"] == 32
    b = chk.validate_edges(torch.from_numpy(np.tile(q0, (100, 1))), torch.from_numpy(np.tile(q1, (100, 1))))
    assert torch.equal(a.first_hit, b.first_hit[:40]) and torch.equal(a.max_deflection, b.max_deflection[:40])


# ---------------------------------------------------------------- fixtures of the reference and other world kinds
def test_saved_multi_branch_fixtures(ctx):
    """environments/multi_branch/env1..4 (URDF + pickle -> CharacterizePlant / _mod observer)."""
    from oracle.c_oracle import COracle
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.model.plants import plant_from_parsed
    from plant_motion_planning_b200.model.urdf import parse_urdf
    with open(os.path.join(G, "plant_fixtures.json")) as fp:
        fx = json.load(fp)
    worlds = [plant_from_parsed(parse_urdf(f["urdf"], is_text=True), f["joint_list"], f["main_stem_indices"],
                                {int(k): v for k, v in f["base_points"].items()}, f["base_pos"]) for f in fx.values()]
    assert [w.num_stems for w in worlds] == [6, 6, 5, 6]
    chk = EdgeChecker(ctx["model"], worlds, EdgeCheckConfig())
    Q = ctx["Q"][:2048]
    g = chk.check_configs(ctx["torch"].from_numpy(Q))
    c = COracle(ctx["model"], worlds).check_configs(Q)
    fl = g.flags.cpu().numpy()
    near = np.any(np.abs(c["deflection"] - 0.30) < EPS_RAD, axis=2)
    assert np.array_equal((fl & 2 > 0)[~near], ((c["flags"] & 2) > 0)[~near])
    err = np.abs(g.deflection.cpu().numpy() - c["deflection"])
    assert np.percentile(err, 99.9) <= 1e-4
    assert (c["deflection"] > 0.05).any()


def test_single_stem_env_sets(ctx):
    """plant_motion_planning/utils.py:120-230 position sets, TwoAngleRepresentation observer, 5-29 stems per world."""
    from oracle.c_oracle import COracle
    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.model.plants import SINGLE_STEM_ENVS, single_stem_world
    assert len(SINGLE_STEM_ENVS) >= 8
    worlds = [single_stem_world(SINGLE_STEM_ENVS[k]) for k in sorted(SINGLE_STEM_ENVS)]
    chk = EdgeChecker(ctx["model"], worlds, EdgeCheckConfig())
    Q = ctx["Q"][:1024]
    g = chk.check_configs(ctx["torch"].from_numpy(Q))
    c = COracle(ctx["model"], worlds).check_configs(Q)
    fl = g.flags.cpu().numpy()
    near = np.any(np.abs(c["deflection"] - 0.30) < EPS_RAD, axis=2)
    assert np.array_equal((fl & 2 > 0)[~near], ((c["flags"] & 2) > 0)[~near])
    err = np.abs(g.deflection.cpu().numpy() - c["deflection"])
    assert np.percentile(err, 99.9) <= 2e-4


def test_facade_matches_reference_call_pattern(ctx):
    """MultiPlantWorld facade: step(q) then checker(q), OR over worlds, gate draws in the reference's order."""
    from oracle.c_oracle import COracle
    from plant_motion_planning_b200.env import MultiPlantWorld
    import random
    np.random.seed(19)
    random.seed(19)
    env = MultiPlantWorld((0, 5), (0, 5), 0.30, 1, 2, 1, ((0.2, 0.35), (-0.5, 0.0)))   # multi_world_avoid_all.py:73-92
    assert len(env.envs) == 4 and list(env.envs) == [(0, 0), (0, 5), (5, 0), (5, 5)]
    worlds = [e.world for e in env.envs.values()]
    co = COracle(env.sample_robot, worlds)
    Q = ctx["Q"][:200]
    c = co.check_configs(Q)
    np.random.seed(5)
    got = []
    for q in Q:
        env.step(tuple(q))
        got.append(env.net_constraint_violation_checker_forward(tuple(q)))
    after = np.random.rand()
    np.random.seed(5)
    want = []
    for i in range(len(Q)):
        hit = False
        for w in range(4):
            if c["flags"][i, w] & 1:
                hit = True
                break
            if (c["flags"][i, w] & 2) and np.random.rand() < 0.95:
                hit = True
                break
        want.append(hit)
    assert got == want and after == np.random.rand()
    # the backward checker ignores the plants: OR over the worlds of the rigid flag alone (env.py:241-247)
    for i in range(20):
        env.step(tuple(Q[i]))
        assert env.net_constraint_violation_checker_backward(tuple(Q[i])) == bool((c["flags"][i] & 1).any())
    env.step(tuple(Q[3]))
    env.envs[(0, 5)].plant_rep.observe_all()
    assert np.allclose(env.envs[(0, 5)].plant_rep.deflections, c["deflection"][3, 1, :6], atol=1e-4)
