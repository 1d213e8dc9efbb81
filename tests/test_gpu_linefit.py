"""sdm_line_fit (LineDetector::LineFitting, /root/reference/src/LineDetector.cc:578-840, :884-900) against the
oracle (oracle/linefit_oracle.py: the reference's control flow with REAL cv2 calls for cv::SVD::solveZ / cv::solve /
cv::norm), on the edge chains the reference's own Edge Drawing library found on the scene's images
(tests/golden/ed_chains_small.npz, oracle/make_ed_golden.py).

Parity level: the device solves the two least-squares problems in closed form in double precision, OpenCV with a float
Jacobi SVD, so the fitted (a, b, c) / (alpha, beta) agree to float rounding and a decision that sits within ~1e-2 of a
threshold (lineFitError <= 1, depthFitError <= 1, point-depth distance > 1.5) can flip.  Stated bars: >= 97 % of the
chains that produce a line on either side give the same number of lines with end points within 0.02 px and 3-D end
points within 1e-3 relative; counting and ordering (per-keyframe counts, chain order) are exact properties.

Second checker, without the solver noise: tests/golden/linefit_ref_small.npz holds the rows of the REFERENCE'S OWN LineFit
text (LineDetector.cc:578-840 compiled where it lies, oracle/Makefile target ref_linefit, oracle/make_linefit_golden.py) with
exact stand-ins for the two OpenCV solver calls - the kernel's control flow against the reference's source: >= 99.8 % of the
chains identical on the dense planes, all of them on the loop's own planes."""
import ctypes as C
import os
from collections import defaultdict

import numpy as np
import pytest

import linefit_oracle as LO
import oracle_py as O
from sdmb200 import api, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _twc(Tcw):
    out = np.zeros(16, np.float32)
    t = np.ascontiguousarray(Tcw, np.float32)
    O.lib().oracle_pose_inverse(t.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out.reshape(4, 4)[:3]


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ed_chains_small.npz"))
    n, w, h, nn, seed = (int(v) for v in g["scene"])
    sc = synth.make_scene(n, w, h, nn, seed=seed)
    offs = [g[f"off_{i}"] for i in range(n)]
    pix = [g[f"pix_{i}"] for i in range(n)]
    return sc, offs, pix


def _chains(off, pix):
    return [[(int(p >> 16), int(p & 0xffff)) for p in pix[off[i]:off[i + 1]]] for i in range(len(off) - 1)]


def _compare(dev_lines, dev_counts, ref_per_kf, min_chains, min_share=0.97):
    """dev_lines: LINE3D array of the batch, ref_per_kf: list (per keyframe) of oracle outputs"""
    assert [int(c) for c in dev_counts] == [int((dev_lines["kf_index"] == i).sum()) for i in range(len(ref_per_kf))]
    assert np.all(np.diff(dev_lines["kf_index"]) >= 0), "keyframes in list order"
    tot = same = 0
    for i, ref in enumerate(ref_per_kf):
        d = dev_lines[dev_lines["kf_index"] == i]
        assert np.all(np.diff(d["chain"]) >= 0), "chains in order (push_back order of :822-823)"
        da, db = defaultdict(list), defaultdict(list)
        for row in d:
            da[int(row["chain"])].append(row)
        for cid, seg, xyz in ref:
            db[cid].append((np.array(seg, np.float32), np.array(xyz, np.float32)))
        for cid in set(da) | set(db):
            tot += 1
            if len(da[cid]) != len(db[cid]):
                continue
            ok = True
            for a, (seg, xyz) in zip(da[cid], db[cid]):
                scale = max(1.0, float(np.abs(xyz).max()))
                ok &= bool(np.abs(a["seg"] - seg).max() <= 0.02 and np.abs(a["xyz"] - xyz).max() <= 1e-3 * scale)
            same += ok
    assert tot >= min_chains, f"only {tot} chains produced a line: the comparison would be vacuous"
    assert same >= min_share * tot, (same, tot)
    return same, tot


def test_line_fit_on_pass2_planes_with_real_edge_chains(golden):
    """the planes the device's own SemiDenseLoop left in the arena (sparse: the scene's textured walls)"""
    sc, offs, pix = golden
    H, W = sc.shape
    with api.Context(width=W, height=H, max_keyframes=sc.n, intra_check=1, intra_grow=1) as ctx:
        ctx.upload_scene(sc)
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items); ctx.pass2(items)
        lines, counts = ctx.line_fit(list(range(sc.n)), offs, pix)
        planes = [ctx.download(i) for i in range(sc.n)]
        assert ctx.last_line_fit_ms() > 0
    ref = [LO.line_fitting(LO.Planes(planes[i]["checked"], planes[i]["sigma"], sc.K, _twc(sc.Tcw[i])), _chains(offs[i], pix[i]))
           for i in range(sc.n)]
    same, tot = _compare(lines, counts, ref, min_chains=30, min_share=0.93)
    print("pass-2 planes:", len(lines), "lines;", same, "of", tot, "chains identical")


def test_line_fit_on_dense_planes_batched_and_single(golden):
    """piecewise-planar inverse-depth planes with noise, holes and wide-sigma pixels, written from outside: thousands of
    lines; a batch over all keyframes equals the per-keyframe calls row for row"""
    sc, offs, pix = golden
    H, W = sc.shape
    chk, sig = linefit_dense_planes(H, W, sc.n)
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        for i in range(sc.n):
            ctx.upload_depth(i, chk[i], sig[i])
            ctx.upload_checked(i, chk[i])
        lines, counts = ctx.line_fit(list(range(sc.n)), offs, pix)
        singles = [ctx.line_fit([i], [offs[i]], [pix[i]]) for i in (0, sc.n - 1)]
        none, c0 = ctx.line_fit([2], [np.zeros(1, np.int32)], [np.zeros(0, np.uint32)])
    assert len(none) == 0 and int(c0[0]) == 0
    for i, (ls, cs) in zip((0, sc.n - 1), singles):
        batch = lines[lines["kf_index"] == i]
        assert int(cs[0]) == len(batch) == len(ls)
        assert np.array_equal(ls["seg"].view(np.uint32), batch["seg"].view(np.uint32))
        assert np.array_equal(ls["xyz"].view(np.uint32), batch["xyz"].view(np.uint32))
        assert np.array_equal(ls["chain"], batch["chain"])
    ref = [LO.line_fitting(LO.Planes(chk[i], sig[i], sc.K, _twc(sc.Tcw[i])), _chains(offs[i], pix[i])) for i in range(sc.n)]
    n_ref = sum(len(r) for r in ref)
    assert n_ref > 1500 and abs(len(lines) - n_ref) <= 0.01 * n_ref
    same, tot = _compare(lines, counts, ref, min_chains=1000)
    print("dense planes:", len(lines), "lines (oracle", n_ref, ");", same, "of", tot, "chains identical")
    # against the reference's own text with exact solvers: the control flow itself, no solver noise
    gref = _golden_ref("dense", sc.n)
    same_r, tot_r = _compare(lines, counts, gref, min_chains=1000, min_share=0.998)
    assert len(lines) == sum(len(r) for r in gref)
    print("dense planes vs the reference's LineFit text (exact solvers):", same_r, "of", tot_r, "chains identical")


from helpers import edge_index_from_chains, linefit_dense_planes  # noqa: E402


def _golden_ref(tag, n):
    """rows of the REFERENCE'S OWN LineFit text with exact solver stand-ins (oracle/make_linefit_golden.py), in the oracle's format"""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "linefit_ref_small.npz"))
    return [[(int(c), list(s), list(x)) for c, s, x in zip(gold[f"{tag}_chain_{i}"], gold[f"{tag}_seg_{i}"], gold[f"{tag}_xyz_{i}"])]
            for i in range(n)]


def test_loop_with_the_real_edge_drawing_mask_then_line_fit(golden):
    """What the reference actually runs: the candidate filter of ProbabilityMapping.cc:454 on the mEdgeIndex plane of the
    real edge detector (golden chains from EDLib.a), the whole SemiDenseLoop bit for bit against the oracle, then the
    line fitting over the same chains on the planes that loop left on the device - depth now sits exactly on the chains."""
    from helpers import compare_planes, run_oracle
    sc, offs, pix = golden
    H, W = sc.shape
    sc = synth.Scene(im=sc.im, grad=sc.grad, theta=sc.theta,
                     edge=np.stack([edge_index_from_chains(offs[i], pix[i], H, W) for i in range(sc.n)]), K=sc.K, Tcw=sc.Tcw,
                     nbr_idx=sc.nbr_idx, rot=sc.rot, min_depth=sc.min_depth, max_depth=sc.max_depth)
    assert 0.05 < (sc.edge >= 0).mean() < 0.3
    osc = run_oracle(sc)   # the shipped loop: intra-keyframe stages commented out (:491-494)
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        for i in range(sc.n):
            assert ctx.candidate_count(i) == int(((sc.grad[i] > 8) & (sc.edge[i] >= 0)).sum())
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items); ctx.pass2(items)
        lines, counts = ctx.line_fit(list(range(sc.n)), offs, pix)
        planes = [ctx.download(i) for i in range(sc.n)]
    dev = {k: np.stack([p[k] for p in planes]) for k in ("depth", "sigma", "checked", "points")}
    rep = compare_planes(dev, osc)
    assert all(rep[k + "_bit_mismatch"] == 0 for k in ("depth", "sigma", "checked", "points")), rep
    assert not (dev["checked"] > 0)[sc.edge < 0].any(), "depth only on edge-chain pixels"
    ref = [LO.line_fitting(LO.Planes(osc.checked[i], osc.sigma[i], sc.K, _twc(sc.Tcw[i])), _chains(offs[i], pix[i]))
           for i in range(sc.n)]
    same, tot = _compare(lines, counts, ref, min_chains=12, min_share=0.85)
    same_r, tot_r = _compare(lines, counts, _golden_ref("mask", sc.n), min_chains=12, min_share=1.0)
    print("real ED mask:", rep["pass2_accepted_ref"], "checked pixels;", len(lines), "lines;", same, "of", tot, "chains identical to the cv2 "
          "oracle,", same_r, "of", tot_r, "to the reference's LineFit text")


def test_line_fit_rejects_bad_arguments(golden):
    sc, offs, pix = golden
    H, W = sc.shape
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        with pytest.raises(api.SdmError):      # no depth planes yet
            ctx.line_fit([0], [offs[0]], [pix[0]])
        ctx.upload_depth(0, np.zeros((H, W), np.float32), np.zeros((H, W), np.float32))
        with pytest.raises(api.SdmError):      # pixel outside the plane
            ctx.line_fit([0], [np.array([0, 12], np.int32)], [np.full(12, (H << 16) | 3, np.uint32)])
        with pytest.raises(api.SdmError):      # slot out of range
            ctx.line_fit([sc.n + 5], [offs[0]], [pix[0]])
