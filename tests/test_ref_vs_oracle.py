"""The oracle (oracle/sdm_oracle.c) against the REFERENCE'S OWN SOURCE: /root/reference/src/ProbabilityMapping.cc is
compiled where it lies against the stand-in headers of oracle/refshim/ (cv::Mat with OpenCV's cv2-pinned evaluation
rules, ORB_SLAM2::KeyFrame / Map with the reference's member names) and its SemiDenseLoop() is run on the same
keyframes.  Gating, loop bounds, candidate filter, EpipolarSearch, fusion, the inter-keyframe check and the point set
are then the reference's text, not a transcription.  Runs where /root/reference exists (this container) or where
oracle/_ref/libref_pm.so was prebuilt (it travels with gpurun); the expected planes are also committed as
tests/golden/ref_loop_small.npz so the comparison survives without either."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_py as O
import ref_py
from helpers import planes_that_grow as _planes_that_grow
from sdmb200 import synth


def _scene():
    sc = synth.make_scene(12, 96, 72, 7, seed=51, contrast=0.9)
    kp_angle = (100.0 + 0.75 * np.arange(sc.n)).astype(np.float32)     # differences are exact in float
    sc.rot = (kp_angle[sc.nbr_idx] - kp_angle[:, None]).astype(np.float32)  # = lower median of GetRotInPlane (:406-415)
    lib = O.lib()
    for i in range(sc.n):  # StereoSearchConstraints (:734-747) from the same sorted inverse depths
        a, b = C.c_float(), C.c_float()
        lib.oracle_stereo_search_constraints(O.fptr(sc.inv_depths[i]), len(sc.inv_depths[i]), C.byref(a), C.byref(b))
        sc.min_depth[i], sc.max_depth[i] = a.value, b.value
    return sc, kp_angle


def _oracle(sc):
    osc = O.OracleScene(sc)
    osc.run()
    return osc


def _check(ref, osc):
    rep = {}
    for k in ("depth", "sigma", "checked", "points"):
        rep[k] = int((ref[k].view(np.uint32) != getattr(osc, k).view(np.uint32)).sum())
    return rep


@pytest.mark.skipif(not ref_py.available(), reason="needs /root/reference or a prebuilt oracle/_ref/libref_pm.so")
def test_oracle_matches_the_reference_source_bit_for_bit():
    sc, kp_angle = _scene()
    ref = ref_py.run_reference_loop(sc, kp_angle)
    assert (ref["flags"] == 1).all(), "every keyframe passes the reference's gating (:359, :365-384, :518-542)"
    osc = _oracle(sc)
    n1, n2 = int((ref["depth"] > 0).sum()), int((ref["checked"] > 0).sum())
    assert n1 > 30000 and n2 > 25000, (n1, n2)
    assert _check(ref, osc) == {"depth": 0, "sigma": 0, "checked": 0, "points": 0}


@pytest.mark.skipif(not ref_py.available(), reason="needs /root/reference or a prebuilt oracle/_ref/libref_pm.so")
def test_reference_with_edge_mask():
    sc = synth.make_scene(12, 96, 72, 7, seed=52, contrast=0.9, edge_mask=True)
    kp_angle = np.full(sc.n, 50.0, np.float32)
    lib = O.lib()
    for i in range(sc.n):
        a, b = C.c_float(), C.c_float()
        lib.oracle_stereo_search_constraints(O.fptr(sc.inv_depths[i]), len(sc.inv_depths[i]), C.byref(a), C.byref(b))
        sc.min_depth[i], sc.max_depth[i] = a.value, b.value
    ref = ref_py.run_reference_loop(sc, kp_angle)
    osc = _oracle(sc)
    assert int((ref["depth"] > 0).sum()) > 3000
    assert _check(ref, osc) == {"depth": 0, "sigma": 0, "checked": 0, "points": 0}


@pytest.mark.skipif(not ref_py.available(), reason="needs /root/reference or a prebuilt oracle/_ref/libref_pm.so")
def test_intra_check_and_grow_match_the_reference_source():
    sc, _ = _scene()
    osc = O.OracleScene(sc)
    osc.run(pass_mask=1)
    p = O.default_params()
    for i in (3, 8):
        d, s = osc.depth[i].copy(), osc.sigma[i].copy()
        rd, rs = ref_py.run_reference_intra(0, d, s, sc.grad[i])
        O.lib().oracle_intra_check(O.fptr(d), O.fptr(s), sc.shape[1], sc.shape[0], C.byref(p))
        assert np.array_equal(rd.view(np.uint32), d.view(np.uint32)) and np.array_equal(rs.view(np.uint32), s.view(np.uint32))
        assert 0 < (d > 0).sum() < (osc.depth[i] > 0).sum()          # the check removes isolated pixels
        gd, gs = ref_py.run_reference_intra(1, d, s, sc.grad[i])
        d2, s2 = d.copy(), s.copy()
        O.lib().oracle_intra_grow(O.fptr(d2), O.fptr(s2), O.fptr(np.ascontiguousarray(sc.grad[i])), sc.shape[1], sc.shape[0], C.byref(p))
        assert np.array_equal(gd.view(np.uint32), d2.view(np.uint32)) and np.array_equal(gs.view(np.uint32), s2.view(np.uint32))
        assert np.array_equal(gd, d)                                   # growing is a no-op


@pytest.mark.skipif(not ref_py.available(), reason="needs /root/reference or a prebuilt oracle/_ref/libref_pm.so")
def test_growing_stage_where_it_does_grow_matches_the_reference_source():
    sc, _ = _scene()
    osc = O.OracleScene(sc)
    osc.run(pass_mask=1)
    p = O.default_params()
    for i in (2, 6):
        d, s = _planes_that_grow(osc.depth[i], osc.sigma[i], seed=40 + i)
        gd, gs = ref_py.run_reference_intra(1, d, s, sc.grad[i])
        d2, s2 = d.copy(), s.copy()
        O.lib().oracle_intra_grow(O.fptr(d2), O.fptr(s2), O.fptr(np.ascontiguousarray(sc.grad[i])), sc.shape[1], sc.shape[0], C.byref(p))
        assert np.array_equal(gd.view(np.uint32), d2.view(np.uint32)) and np.array_equal(gs.view(np.uint32), s2.view(np.uint32))
        grown = (gd != d) | (gs != s)
        assert grown.sum() > 50                                            # the stage did something
        assert not grown[(s == 0) & (d < 1e-6)].any()                      # ... but never where the centre's sigma is +-0


def test_oracle_matches_committed_reference_output(golden_dir):
    """the planes the reference source produced here (tests/golden/ref_loop_small.npz, written by
    `python tests/test_ref_vs_oracle.py`) — no reference tree and no _ref library needed"""
    g = np.load(os.path.join(golden_dir, "ref_loop_small.npz"))
    sc, _ = _scene()
    assert np.array_equal(sc.im, g["im"]), "scene generator drifted: regenerate the fixture"
    osc = _oracle(sc)
    for k in ("depth", "sigma", "checked", "points"):
        assert np.array_equal(getattr(osc, k).view(np.uint32), g[k].view(np.uint32)), k


if __name__ == "__main__":  # regenerate the committed fixture from the reference source
    sc, kp_angle = _scene()
    ref = ref_py.run_reference_loop(sc, kp_angle)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_loop_small.npz")
    np.savez_compressed(out, im=sc.im, depth=ref["depth"], sigma=ref["sigma"], checked=ref["checked"], points=ref["points"])
    print("wrote", out, {k: int((ref[k] > 0).sum()) for k in ("depth", "checked")})
