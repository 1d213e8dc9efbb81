"""The single-method / bookkeeping entry points of the C-ABI that the batch tests do not reach: sdm_inter_check on
planes written from outside (dense pass-2 kernel), sdm_update_points after sdm_set_pose (PoseChanged refresh,
ProbabilityMapping.cc:678-697), sdm_set_intrinsics / sdm_mark_pass1_done / sdm_depth_plane_ptr for halo slots,
pinned host memory, timing and statistics calls."""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
from helpers import planes_that_grow as _planes_that_grow, run_oracle
from sdmb200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return synth.make_scene(9, 320, 240, 6, seed=23)


def _bits(a, b):
    return int((np.ascontiguousarray(a).view(np.uint32) != np.ascontiguousarray(b).view(np.uint32)).sum())


def test_inter_check_on_external_planes(scene):
    """InterKeyFrameDepthChecking(kf, neighbours) as a single call: every (rho,sigma) plane arrives through
    sdm_upload_depth (like cv::Mat planes of an earlier loop), so pass 2 runs its every-pixel kernel."""
    sc = scene
    osc = O.OracleScene(sc)
    osc.run(pass_mask=1)                       # pass-1 planes of every keyframe
    O.inter_check(osc, 4)
    with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        for i in range(sc.n):
            ctx.upload_depth(i, osc.depth[i], osc.sigma[i])
        item = api.make_items([4], sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx._chk(ctx.lib.sdm_inter_check(ctx.h, item))
        got = ctx.download(4)
    assert (osc.checked[4] > 0).sum() > 5000
    assert _bits(got["checked"], osc.checked[4]) == 0 and _bits(got["points"], osc.points[4]) == 0
    assert _bits(got["depth"], osc.depth[4]) == 0 and _bits(got["sigma"], osc.sigma[4]) == 0  # de-interleave path


def test_update_points_after_pose_change(scene):
    sc = scene
    osc = run_oracle(sc)
    with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items); ctx.pass2(items)
        T = sc.Tcw[2].copy()
        T[:, 3] += np.array([0.03, -0.02, 0.05], np.float32)   # loop closure moved the keyframe
        ctx.set_pose(2, T)
        ctx.update_points([2, 6])
        O.update_points(osc, 2, T)
        O.update_points(osc, 6)
        for i in (2, 6):
            got = ctx.download(i)
            assert _bits(got["points"], osc.points[i]) == 0 and _bits(got["checked"], osc.checked[i]) == 0
        assert np.abs(osc.points[2]).max() > 0


def test_halo_slot_without_images(scene):
    """a halo slot may carry only calibration, pose and (rho,sigma) received from outside: enough for pass 2"""
    sc = scene
    osc = run_oracle(sc)
    with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
        nb = [int(j) for j in sc.nbr_idx[4]]
        for i in [4] + nb[:3]:
            ctx.upload_keyframe(i, sc.im[i], sc.grad[i], sc.theta[i], None, sc.K, sc.Tcw[i])
        for i in nb[3:]:                                      # no image planes for these
            ctx.set_intrinsics(i, sc.K); ctx.set_pose(i, sc.Tcw[i])
        for i in [4] + nb:
            ctx.upload_depth(i, osc.depth[i], osc.sigma[i])
        ptr, nbytes = ctx.depth_plane_ptr(nb[4])
        assert ptr and nbytes == 320 * 240 * 8
        ctx.mark_pass1_done(nb[4])
        ctx.pass2(api.make_items([4], sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth))
        got = ctx.download(4, depth=False, sigma=False)
    assert _bits(got["checked"], osc.checked[4]) == 0


def test_pinned_memory_timing_and_stats(scene):
    sc = scene
    lib = api.load()
    p = C.c_void_p()
    assert lib.sdm_host_alloc(C.byref(p), 1 << 20) == 0 and p.value
    assert lib.sdm_host_free(p) == 0
    assert b"sm_100a" in lib.sdm_version()
    with api.Context(width=320, height=240, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        l0 = ctx.launch_count()
        ctx.mark(0); ctx.pass1(items); ctx.pass2(items); ctx.mark(1)
        total = ctx.elapsed_ms(0, 1)
        p1, p2 = ctx.last_pass_ms()
        t = ctx.last_timing()
        assert 0 < p1 <= total and 0 < p2 <= total and abs(t["pass1_scan_ms"] + t["pass1_intra_ms"] - p1) < 0.05
        assert ctx.launch_count() - l0 == 4               # k_plan + scan, k_plan + pass 2
        st = ctx.stats()
        assert st["candidates"] == int((sc.grad > 8).sum()) and 0 < st["checked"] <= st["fused"] <= st["candidates"]
        with pytest.raises(api.SdmError):
            ctx.elapsed_ms(2, 3)                           # marks never recorded


def test_planes_produced_on_the_device(scene):
    """sdm_upload_keyframes with grad = theta = NULL: Scharr/32 + sqrt + scalar fastAtan2 on the device reproduce
    the scalar-form planes of the generator bit for bit, so the whole loop gives the same planes as the upload of
    host-produced GradImg / GradTheta (and H2D shrinks from 9 to 1 byte per pixel)."""
    sc = scene
    osc = run_oracle(sc)
    H, W = sc.shape
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_keyframes(ctx.upload_descs(sc, range(sc.n), images_only=True))
        for i in (0, 4, sc.n - 1):
            g, t = ctx.download_planes(i)
            assert _bits(g, sc.grad[i]) == 0 and _bits(t, sc.theta[i]) == 0
            assert ctx.candidate_count(i) == int((sc.grad[i] > 8).sum())
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items); ctx.pass2(items)
        for i in range(sc.n):
            got = ctx.download(i)
            assert _bits(got["depth"], osc.depth[i]) == 0 and _bits(got["checked"], osc.checked[i]) == 0
    # odd size: borders (BORDER_REFLECT_101) and partial tiles
    sc2 = synth.make_scene(2, 45, 19, 1, seed=3, contrast=0.9)
    with api.Context(width=45, height=19, max_keyframes=2) as ctx:
        ctx.upload_keyframe(0, sc2.im[0], None, None, None, sc2.K, sc2.Tcw[0])
        g, t = ctx.download_planes(0)
        assert _bits(g, sc2.grad[0]) == 0 and _bits(t, sc2.theta[0]) == 0


def test_device_planes_against_real_cv2_magnitude_and_phase(scene):
    """SURVEY 8f-1 pinned to OpenCV itself: the planes k_pack_image produces from im_ against the REAL calls of
    KeyFrame.cc:69-74 (cv2.Scharr x2, cv2.magnitude, cv2.phase(angleInDegrees=True); cv2 4.x of this image).  OpenCV's
    SIMD magnitude / phase are not the scalar forms (and differ between calls on a multi-threaded host), so the bar is
    what can be measured: <= 2 ulp on GradImg (SURVEY 8(c) saw 1 on its samples; cv2 4.13's magnitude reaches 2 on these
    images), <= 3.1e-5 degrees on GradTheta - and, downstream, the accepted sets of a whole loop on OpenCV's planes vs
    the device's planes differ by at most BASELINE's 0.1 %."""
    cv2 = pytest.importorskip("cv2")
    sc = scene
    H, W = sc.shape
    mag, ph = np.zeros_like(sc.grad), np.zeros_like(sc.theta)
    for i in range(sc.n):
        gx = cv2.Scharr(sc.im[i], cv2.CV_32F, 1, 0, scale=1 / 32.0)
        gy = cv2.Scharr(sc.im[i], cv2.CV_32F, 0, 1, scale=1 / 32.0)
        mag[i] = cv2.magnitude(gx, gy)
        ph[i] = cv2.phase(gx, gy, angleInDegrees=True)
    import copy
    sc_cv = copy.copy(sc)
    sc_cv.grad, sc_cv.theta = mag, ph
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_keyframes(ctx.upload_descs(sc, range(sc.n), images_only=True))
        worst_ulp, worst_deg = 0, 0.0
        for i in range(sc.n):
            g, t = ctx.download_planes(i)
            ulp = np.abs(g.view(np.int32).astype(np.int64) - mag[i].view(np.int32).astype(np.int64))  # both >= +0
            worst_ulp = max(worst_ulp, int(ulp.max()))
            d = np.abs(t.astype(np.float64) - ph[i].astype(np.float64))
            worst_deg = max(worst_deg, float(np.minimum(d, 360.0 - d).max()))
        assert worst_ulp <= 2, worst_ulp
        assert worst_deg <= 3.1e-5, worst_deg
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items); ctx.pass2(items)
        dev = [ctx.download(i) for i in range(sc.n)]
        ctx.upload_keyframes(ctx.upload_descs(sc_cv, range(sc.n)))   # OpenCV's own planes through the same loop
        ctx.pass1(items); ctx.pass2(items)
        ocv = [ctx.download(i) for i in range(sc.n)]
    for key in ("depth", "checked"):
        a = np.stack([d[key] for d in dev]) > 0
        b = np.stack([d[key] for d in ocv]) > 0
        assert b.sum() > 10000
        assert (a != b).sum() <= 1e-3 * b.sum(), (key, int((a != b).sum()), int(b.sum()))


def test_scatter_download_equals_dense_download(scene):
    """sdm_scatter_keyframes (candidate records over PCIe + host-side scatter by the library's worker threads) into
    zero-initialised planes == sdm_download_keyframes == the oracle, bit for bit; row-pitched destination planes;
    pass-1-only call; refused for slots whose planes came from outside"""
    sc = scene
    osc = run_oracle(sc)
    H, W = sc.shape
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items)
        # pass-1 planes only, into a pitched destination (ROI of a wider array)
        wide = np.zeros((2, H, W + 24), np.float32)
        d = (api.DownloadDesc * 1)()
        d[0].kf = 3
        d[0].depth, d[0].depth_step = wide[0, :, 8:].ctypes.data, wide.strides[1]
        d[0].sigma, d[0].sigma_step = wide[1, :, 8:].ctypes.data, wide.strides[1]
        ctx.scatter_keyframes(d); ctx.synchronize()
        assert _bits(wide[0, :, 8:8 + W], osc.depth[3]) == 0 and _bits(wide[1, :, 8:8 + W], osc.sigma[3]) == 0
        assert not wide[:, :, :8].any() and not wide[:, :, 8 + W:].any()
        ctx.pass2(items)
        for rep in range(2):  # twice: staging ring re-use, same planes overwritten
            got = ctx.scatter_all(list(range(sc.n)))
            for k, ref in (("depth", osc.depth), ("sigma", osc.sigma), ("checked", osc.checked), ("points", osc.points)):
                assert _bits(got[k], ref) == 0, k
        dense = ctx.download(2)
        assert all(_bits(got[k][2], dense[k]) == 0 for k in ("depth", "sigma", "checked", "points"))
        ctx.upload_depth(1, osc.depth[1], osc.sigma[1])
        d[0].kf = 1
        with pytest.raises(api.SdmError) as e:
            ctx.scatter_keyframes(d)
        assert e.value.code == -3  # SDM_ERR_STATE


def test_inter_chi_gate_float_shortcut_is_exact():
    """chi_inter_accept decides in float unless the value is within 2^-16 of 3.84; the decision must equal the
    reference's double evaluation `(float)(dd*dd / (sigma*sigma)) < 3.84` everywhere: random inputs, inputs placed
    within a few ulp of the threshold, zeros, subnormal / huge squares, infinities and NaN."""
    rng = np.random.default_rng(7)
    n = 400000
    sg = np.exp(rng.uniform(np.log(1e-4), np.log(0.5), n)).astype(np.float32)
    chi = np.empty(n)
    chi[:n // 2] = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n // 2))
    # the other half hugs the threshold: 3.84 * (1 + eps), |eps| from 1e-9 to 1e-3, both signs
    eps = np.exp(rng.uniform(np.log(1e-9), np.log(1e-3), n - n // 2)) * rng.choice([-1.0, 1.0], n - n // 2)
    chi[n // 2:] = 3.84 * (1.0 + eps)
    diff = (np.sqrt(chi) * sg.astype(np.float64) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    extra_d = np.array([0, 0, 1e-30, 1e-30, 1e25, 1e25, np.inf, 1.0, np.nan, 1.0, 1e-20, 3e19, -0.0, 1e-44], np.float32)
    extra_s = np.array([0, 1, 1e-30, 1e-25, 1e25, 1e20, 1.0, np.inf, 1.0, np.nan, 1e-20, 3e19, 1e-44, 1e-44], np.float32)
    diff, sg = np.concatenate([diff, extra_d]), np.concatenate([sg, extra_s])
    with np.errstate(all="ignore"):
        dd = diff.astype(np.float64)
        ref = ((dd * dd) / (sg.astype(np.float64) * sg.astype(np.float64))).astype(np.float32) < np.float64(3.84)
    with api.Context(width=64, height=48, max_keyframes=1) as ctx:
        got = ctx.inter_chi_test(diff, sg)
    assert np.array_equal(got, ref), np.flatnonzero(got != ref)[:10]
    near = np.abs(chi[n // 2:] / 3.84 - 1.0) < 2.0 ** -16
    assert near.sum() > 1000 and (~near).sum() > 1000   # both the double fallback and the float shortcut were exercised


def test_growing_stage_where_it_does_grow(scene):
    """IntraKeyFrameDepthGrowing (:929-976) is a no-op on the planes the shipped pipeline produces (an empty pixel is
    (0, 0) and a centre with sigma = 0 fails every ChiTest); the kernel uses exactly that as an early exit.  Planes
    that break the invariant on purpose - holes that keep a non-zero sigma, rho just under 1e-6, sigma = +-0 next to
    them - must still grow like the reference (tests/test_ref_vs_oracle.py pins the oracle on the same construction)."""
    sc = scene
    osc = O.OracleScene(sc)
    osc.run(pass_mask=1)
    H, W = sc.shape
    p = O.default_params()
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        for i in (2, 5):
            d, s = _planes_that_grow(osc.depth[i], osc.sigma[i], seed=60 + i)
            ctx.upload_depth(i, d, s)
            ctx.intra_grow(i)
            got = ctx.download(i, checked=False, points=False)
            d2, s2 = d.copy(), s.copy()
            O.lib().oracle_intra_grow(O.fptr(d2), O.fptr(s2), O.fptr(np.ascontiguousarray(sc.grad[i])), W, H, C.byref(p))
            assert np.array_equal(got["depth"].view(np.uint32), d2.view(np.uint32))
            assert np.array_equal(got["sigma"].view(np.uint32), s2.view(np.uint32))
            assert ((d2 != d) | (s2 != s)).sum() > 50
