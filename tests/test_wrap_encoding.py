"""The wrap-encoded orientation texel of the scan loop (encode_theta_pair in sdm_kernels.cuh, used by scan_columns2),
checked on the CPU in numpy float32 (one IEEE round-to-nearest per elementwise operation, no contraction):

    reference  yangle(a0, a1; w0, w1)   ProbabilityMapping.cc:83-111
    device     e = encode(a0, a1) once per texel;  s = e0*w0 + e1*w1;  if s <= -360: s += 360;  gth = |s|

must give the same bits for every orientation pair in [0, 360] (what cv::phase produces; the loop's precondition) and
every weight pair the loop can form (w1 = v - floor(v) in [0, 1), w0 = 1 - w1).  Random pairs plus a grid of the values
where the branches of yangle switch (0, -0, 180 apart, next to 360)."""
import numpy as np

F = np.float32


def yangle_ref(a0, a1, w0, w1):
    a0, a1 = a0.copy(), a1.copy()
    direct = np.abs(a0 - a1) < F(180)
    lo0 = (a0 < a1) & ~direct
    lo1 = ~(a0 < a1) & ~direct
    a0 = np.where(lo0, a0 + F(360), a0)
    a1 = np.where(lo1, a1 + F(360), a1)
    inter = a0 * w0 + a1 * w1
    wrapped = np.where(inter >= F(360), inter - F(360), inter)
    return np.where(direct, inter, wrapped)


def encode(a0, a1):
    direct = np.abs(a0 - a1) < F(180)
    e0 = np.where(direct, a0, -np.where(a0 < a1, a0 + F(360), a0))
    e1 = np.where(direct, a1, -np.where(a0 < a1, a1, a1 + F(360)))
    return e0.astype(F), e1.astype(F)


def yangle_dev(e0, e1, w0, w1):
    s = e0 * w0 + e1 * w1
    s = np.where(s <= F(-360), s + F(360), s)
    return np.abs(s)


def _weights(rng, n):
    # v = row coordinate as the loop has it (float32, 0 <= v <= 2047): w1 = v - floor(v) exact, w0 = 1 - w1 rounded
    v = (rng.random(n) * rng.choice([1.0, 8.0, 480.0, 2047.0], n)).astype(F)
    v[: n // 50] = np.floor(v[: n // 50])  # integer rows: w1 = 0
    w1 = (v - np.floor(v)).astype(F)
    return (F(1) - w1).astype(F), w1


def _check(a0, a1, w0, w1):
    ref = yangle_ref(a0, a1, w0, w1).astype(F)
    e0, e1 = encode(a0, a1)
    dev = yangle_dev(e0, e1, w0, w1).astype(F)
    # the sign of a zero result is the one thing the loop cannot see (|s|; gates compare differences): compare values
    same = (ref.view(np.uint32) == dev.view(np.uint32)) | ((ref == 0) & (dev == 0))
    assert same.all(), (a0[~same][:5], a1[~same][:5], w0[~same][:5], w1[~same][:5], ref[~same][:5], dev[~same][:5])
    # the flag: a wrapped pair is strictly negative, any other pair is >= +-0
    direct = np.abs(a0 - a1) < F(180)
    assert (e0[~direct] < 0).all() and (e1[~direct] < 0).all()
    assert (e0[direct] >= 0).all() and (e1[direct] >= 0).all()
    assert ((e0 * w0 + e1 * w1)[direct] >= 0).all()


def test_random_pairs():
    rng = np.random.default_rng(7)
    n = 4_000_000
    a0 = (rng.random(n) * 360).astype(F)
    a1 = (rng.random(n) * 360).astype(F)
    # half of the pairs close to each other modulo 360 (what neighbouring rows of a real orientation plane look like)
    near = rng.random(n) < 0.5
    a1 = np.where(near, np.mod(a0 + (rng.standard_normal(n) * 8).astype(F), F(360)).astype(F), a1)
    a1 = np.clip(a1, F(0), F(360))
    _check(a0, a1, *_weights(rng, n))


def test_branch_boundaries():
    rng = np.random.default_rng(8)
    up = lambda x: np.nextafter(F(x), F(1e9))
    dn = lambda x: np.nextafter(F(x), F(-1e9))
    vals = np.array([0.0, -0.0, 1e-30, up(0), 3e-5, 45.0, 90.0, dn(180), 180.0, up(180), 270.0, 315.0, dn(360), 360.0,
                     0.5, 179.5, 180.5, 359.5, 100.0, 280.0, dn(100), up(280)], F)
    a0, a1 = [g.ravel() for g in np.meshgrid(vals, vals)]
    reps = 4000
    a0, a1 = np.tile(a0, reps), np.tile(a1, reps)
    w0, w1 = _weights(rng, a0.size)
    _check(a0, a1, w0, w1)
    # pairs exactly 180 apart take the wrapped branch (the reference's test is a strict `<`)
    b0 = (rng.random(200000) * 180).astype(F)
    b1 = (b0 + F(180)).astype(F)
    keep = np.abs(b0 - b1) == F(180)
    _check(b0[keep], b1[keep], *_weights(rng, int(keep.sum())))
    _check(b1[keep], b0[keep], *_weights(rng, int(keep.sum())))
