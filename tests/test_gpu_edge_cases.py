"""Edge cases of the path, CUDA (through the C-ABI) vs oracle: image sizes that are not multiples of the tile,
keyframes without candidates, duplicate / degenerate poses (NaN epipolar lines), the largest neighbour count the ABI
accepts, planes with out-of-range and non-finite values, ragged batches, slot re-use.  The reference has no tests for
any of these; the oracle defines the behaviour (it never indexes out of bounds where the reference would)."""
import json

import numpy as np
import pytest

from helpers import compare_planes, run_device, run_oracle
from sdmb200 import api, synth

pytestmark = pytest.mark.gpu


def _bit_equal(dev, osc, keys=("depth", "sigma", "checked", "points")):
    ref = {"depth": osc.depth, "sigma": osc.sigma, "checked": osc.checked, "points": osc.points}
    return {k: int((dev[k].view(np.uint32) != ref[k].view(np.uint32)).sum()) for k in keys}


@pytest.mark.parametrize("W,H", [(333, 217), (64, 48), (17, 9)])
def test_odd_image_sizes(W, H):
    sc = synth.make_scene(8, W, H, 6, seed=3, contrast=0.9)
    osc, dev = run_oracle(sc), run_device(sc)
    assert _bit_equal(dev, osc) == {"depth": 0, "sigma": 0, "checked": 0, "points": 0}
    assert dev["stats"]["candidates"] == osc.stats.as_dict()["candidates"]


def test_keyframes_without_candidates_and_flat_neighbours():
    """a flat keyframe has no candidate pixel (ragged batch); a flat neighbour never passes condition 1"""
    sc = synth.make_scene(8, 160, 120, 6, seed=9)
    for i in (2, 5):
        sc.im[i][:] = 77
        sc.grad[i][:] = 0
        sc.theta[i][:] = 0
    osc, dev = run_oracle(sc), run_device(sc)
    assert _bit_equal(dev, osc) == {"depth": 0, "sigma": 0, "checked": 0, "points": 0}
    assert (dev["depth"][2] == 0).all() and (dev["checked"][5] == 0).all()
    assert (osc.depth[0] > 0).sum() > 100  # the other keyframes still reconstruct


def test_duplicate_poses_give_nan_lines_not_faults():
    """two keyframes at the same pose: t12 = 0, F12 = 0, a/b = NaN -> EpipolarSearch yields nothing for that pair"""
    sc = synth.make_scene(8, 160, 120, 6, seed=10)
    sc.Tcw[3] = sc.Tcw[4]
    osc, dev = run_oracle(sc), run_device(sc)
    assert _bit_equal(dev, osc) == {"depth": 0, "sigma": 0, "checked": 0, "points": 0}
    assert np.isfinite(dev["depth"]).all()


def test_sixteen_neighbours():
    sc = synth.make_scene(18, 128, 96, api.SDM_MAX_NBR, seed=12, contrast=0.9)
    osc, dev = run_oracle(sc), run_device(sc)
    rep = compare_planes(dev, osc)
    print(json.dumps(rep))
    assert rep["depth_bit_mismatch"] == 0 and rep["checked_bit_mismatch"] == 0
    with api.Context(width=128, height=96, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        items = api.make_items([0], sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        items[0].n_nbr = api.SDM_MAX_NBR + 1
        with pytest.raises(api.SdmError) as e:
            ctx.pass1(items)
        assert e.value.code == -1


def test_out_of_range_and_non_finite_planes():
    """GradTheta outside [0,360), huge / NaN gradient magnitudes, extreme depth bounds: same decisions, no faults"""
    sc = synth.make_scene(8, 160, 120, 6, seed=14)
    rng = np.random.default_rng(0)
    sc.theta[1] += 360.0                       # every angle one turn too large
    sc.theta[2] -= 400.0
    m = rng.random(sc.grad[3].shape) < 0.02
    sc.grad[3][m] = np.nan
    sc.theta[3][rng.random(sc.grad[3].shape) < 0.02] = np.inf
    sc.grad[6][rng.random(sc.grad[6].shape) < 0.01] = 3e38
    sc.min_depth[4], sc.max_depth[4] = 1e6, 1e-6     # search range = the whole epipolar line
    sc.min_depth[5] = -2.0                           # mean - 2 sigma < 0 (no guard in the reference, Appendix A.9)
    osc, dev = run_oracle(sc), run_device(sc)
    diff = _bit_equal(dev, osc)
    # NaN payloads may differ in their mantissa bits: compare NaN-ness there and bits elsewhere
    for k, ref in (("depth", osc.depth), ("sigma", osc.sigma), ("checked", osc.checked)):
        a, b = dev[k], ref
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        ok = ~np.isnan(b)
        assert np.array_equal(a[ok].view(np.uint32), b[ok].view(np.uint32)), (k, diff)


def test_slot_reuse_and_partial_batches():
    """re-uploading other keyframes into used slots and running the passes keyframe by keyframe in arbitrary order"""
    a = synth.make_scene(8, 160, 120, 6, seed=15)
    b = synth.make_scene(8, 160, 120, 6, seed=16)
    ob = run_oracle(b)
    with api.Context(width=160, height=120, max_keyframes=8) as ctx:
        run_device(a, ctx=ctx)                         # fills every slot and every output plane
        ctx.upload_scene(b)                            # same slots, different keyframes
        order = [5, 0, 7, 2, 1, 6, 3, 4]
        for i in order:
            ctx.pass1(api.make_items([i], b.nbr_idx, b.rot, b.min_depth, b.max_depth))
        for i in reversed(order):
            ctx.pass2(api.make_items([i], b.nbr_idx, b.rot, b.min_depth, b.max_depth))
        dev = {k: np.stack([ctx.download(i)[k] for i in range(8)]) for k in ("depth", "sigma", "checked", "points")}
    assert _bit_equal(dev, ob) == {"depth": 0, "sigma": 0, "checked": 0, "points": 0}


def test_scan_generations_agree(monkeypatch):
    """the three generations of the column loop (SDM_SCAN=lane1 | lane2 | lane3; the first is also what any other
    threshold set runs) and the warp-per-pixel kernel (SDM_SCAN=warp) give the oracle's bits"""
    sc = synth.make_scene(10, 320, 240, 6, seed=21, contrast=0.9)
    osc = run_oracle(sc)
    zero = {"depth": 0, "sigma": 0, "checked": 0, "points": 0}
    monkeypatch.delenv("SDM_SCAN", raising=False)
    with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
        assert ctx.scan_generation() in (2, 3)     # k_verify_div passed for THETA = 0.23f
        assert _bit_equal(run_device(sc, ctx=ctx), osc) == zero
    # warp / warp_tma = the warp-per-pixel A/B kernels (the second stages neighbour tiles in shared memory by TMA)
    for env, gen in (("lane1", 1), ("lane2", 2), ("lane3", 3), ("warp", 0), ("warp_tma", 0)):
        monkeypatch.setenv("SDM_SCAN", env)
        with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
            assert ctx.scan_generation() == gen
            assert _bit_equal(run_device(sc, ctx=ctx), osc) == zero
    monkeypatch.delenv("SDM_SCAN")
    # the two builds of the third-generation kernel (48 registers / 10 blocks per SM; 64 / 8 for long scans): forced, then
    # chosen by sdm_pass1 from the batch's mean search range (this scene: short; a wide depth range: long)
    for force, want in (("0", False), ("1", True)):
        monkeypatch.setenv("SDM_SCAN_LONG", force)
        with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
            assert _bit_equal(run_device(sc, ctx=ctx), osc) == zero
            assert ctx.last_scan_long() is want
    monkeypatch.delenv("SDM_SCAN_LONG")
    with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
        run_device(sc, ctx=ctx)
        assert ctx.last_scan_long() is False
    wide = synth.make_scene(8, 640, 480, 6, seed=21, contrast=0.9, wide_range=True)
    with api.Context(width=640, height=480, max_keyframes=wide.n) as ctx:
        dev = run_device(wide, ctx=ctx)
        picked = ctx.last_scan_long()
    assert _bit_equal(dev, run_oracle(wide)) == zero
    print("wide-range scene: long-scan build picked:", picked)
    with api.Context(width=320, height=240, max_keyframes=sc.n, lambdaG=9) as ctx:
        assert ctx.scan_generation() == 1          # not the reference's constants


def test_gate_boundaries_and_irregular_rot():
    """orientation planes quantised to 5 degrees (with exact 0 and 360 entries) make the raw differences of gates 2 / 3
    land exactly on 45, 315, -45, -315, 80, 100 ...: the short gate forms of scan_columns2 must decide like the
    reference's fold sequences there.  A relative roll outside [-360, 360] sends its keyframe through the
    first-generation loop (same bits)."""
    sc = synth.make_scene(8, 200, 150, 6, seed=22, contrast=0.9)
    q = np.round(sc.theta / 5.0) * 5.0
    q[q > 360.0] = 360.0
    sc.theta[:] = q.astype(np.float32)
    assert (sc.theta == 360.0).any() and (sc.theta == 0.0).any()
    zero = {"depth": 0, "sigma": 0, "checked": 0, "points": 0}
    osc, dev = run_oracle(sc), run_device(sc)
    assert _bit_equal(dev, osc) == zero
    assert (osc.depth > 0).sum() > 1000
    sc.rot[2][:] = 45.0                    # regular: gate 3 against th_pi + 45
    sc.rot[4][:] = -315.0
    sc.rot[5][:] = 725.0                   # irregular: first-generation loop for keyframe 5
    osc, dev = run_oracle(sc), run_device(sc)
    assert _bit_equal(dev, osc) == zero


def test_wrap_encoded_orientation_pairs():
    """scan_columns2 reads texels whose orientation pair is pre-processed for yangle (:83-111): pairs more than 180
    degrees apart carry the `+ 360` of the smaller angle and a sign flag (encode_theta_pair).  Planes built to stress
    exactly that: most vertical pairs straddle the 0/360 seam, with exact 0, -0, 360, the floats next to 360 and to
    180 apart, and pairs exactly 180 apart (the reference's `< 180` test is strict).  Own-pixel orientations (th_pi of
    gate 3) come from the same planes, so candidates whose own texel is wrapped are covered too."""
    sc = synth.make_scene(8, 200, 150, 6, seed=23, contrast=0.9)
    rng = np.random.default_rng(5)
    H, W = sc.shape
    near0 = np.float32([0.0, -0.0, 1e-30, 3e-5, 0.5, 7.25, 44.99999, 90.0, 179.99998, 180.0])
    near360 = np.float32([360.0, 359.99997, 359.5, 352.75, 315.00003, 270.0, 180.00002, 180.0])
    for i in range(sc.n):
        seam = rng.random((H, W)) < 0.6                      # pixels moved next to the seam
        hi_row = (np.arange(H)[:, None] + rng.integers(0, 2)) % 2 == 0
        lo = rng.choice(near0, size=(H, W))
        hi = rng.choice(near360, size=(H, W))
        jitter = (rng.random((H, W)) * 20).astype(np.float32)
        lo = np.where(rng.random((H, W)) < 0.5, lo, jitter)
        hi = np.where(rng.random((H, W)) < 0.5, hi, np.float32(360.0) - jitter)
        sc.theta[i] = np.where(seam, np.where(hi_row, hi, lo), sc.theta[i]).astype(np.float32)
    assert sc.theta.min() >= 0 and sc.theta.max() <= 360
    d = np.abs(sc.theta[:, 1:] - sc.theta[:, :-1])
    assert (d >= 180).mean() > 0.2 and (d == 180).any()
    zero = {"depth": 0, "sigma": 0, "checked": 0, "points": 0}
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        assert ctx.scan_generation() >= 2
        osc, dev = run_oracle(sc), run_device(sc, ctx=ctx)
    assert _bit_equal(dev, osc) == zero
    assert (osc.depth > 0).sum() > 300
    sc.rot[1][:] = 90.0
    sc.rot[3][:] = -180.0
    osc, dev = run_oracle(sc), run_device(sc)
    assert _bit_equal(dev, osc) == zero
