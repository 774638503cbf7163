"""The kernels take the components of (p - nearest box point) as copysign(sat(|p| - h), p) (csrc/pmp_kernels.cuh:
pmp_outside, and sat(|p| - h) in the contact tests) where the oracle writes p - clamp(p, -h, h).  IEEE float32 on the
host shows what the GPU tests rely on: the two are the same float while the component is at most 1 m, a signed zero
inside the box, and a saturated component alone already puts the squared distance at >= 1 m^2 -- beyond any sphere
radius (<= 0.5 m) plus stem margin (<= 0.4 m), which pmp_set_robot / pmp_set_worlds enforce."""
import numpy as np

f32 = np.float32


def _outside(v, h):
    return np.copysign(np.clip(np.abs(v) - h, f32(0), f32(1)), v).astype(f32)


def test_same_float_below_one_metre_and_signed_zero_inside():
    rng = np.random.RandomState(0)
    for h in (f32(0.025), f32(0.1), f32(0.31)):
        v = np.concatenate([rng.uniform(-1.3, 1.3, 200000), rng.normal(0, 0.05, 200000), [h, -h, 0.0, -0.0, 1 + h, -1 - h]]).astype(f32)
        ref = (v - np.clip(v, -h, h)).astype(f32)
        got = _outside(v, h)
        near = np.abs(v) - h <= f32(1)
        assert np.array_equal(ref[near], got[near])                      # == also equates -0.0 with +0.0
        nz = near & (ref != 0)
        assert np.array_equal(ref[nz].view(np.uint32), got[nz].view(np.uint32))
        inside = np.abs(v) <= h
        assert np.all(got[inside] == 0) and np.all(ref[inside] == 0)
        far = ~near
        assert np.all(np.abs(got[far]) == 1) and np.all(np.abs(ref[far]) > 1)


def test_saturated_distance_still_says_apart():
    """(r + margin)^2 - d2 keeps its sign when a component saturates: d2 >= 1 > 0.81 >= (r + margin)^2, and the
    penetration max(r + margin - |d|, 0) is the same exact zero; padding spheres (radius -0.5, parked 1e6 m away)
    never reach a positive margin either."""
    rng = np.random.RandomState(1)
    p = rng.uniform(-4, 4, size=(100000, 3)).astype(f32)
    h = np.array([0.025, 0.025, 0.3], f32)
    exact = (p - np.clip(p, -h, h)).astype(f32)
    sat = _outside(p, h)
    d2e = (exact * exact).sum(1) + f32(1e-24)
    d2s = (sat * sat).sum(1) + f32(1e-24)
    for rr in (f32(0.05), f32(0.5), f32(0.9)):
        assert np.array_equal(rr * rr - d2e > 0, rr * rr - d2s > 0)
        assert np.array_equal(np.maximum(rr - np.sqrt(d2e), 0) > 0, np.maximum(rr - np.sqrt(d2s), 0) > 0)
        same = np.all(np.abs(p) - h <= 1, axis=1)
        assert np.array_equal(d2e[same], d2s[same]) and same.sum() > 1000 and (~same).sum() > 1000
    far = _outside(np.full(3, 1e6, f32), h)
    nrr = f32(0.5) - f32(1e-3)                                           # -(r_pad + margin)
    assert nrr * nrr - (far * far).sum() < 0 and max(-(np.sqrt((far * far).sum()) + nrr), 0) == 0
