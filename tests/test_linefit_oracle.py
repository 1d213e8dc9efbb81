"""CPU tests of the line-fitting oracle (oracle/linefit_oracle.py) and of the committed Edge Drawing chains.

The oracle's OpenCV calls are real cv2 calls; what is pinned here is the restated control flow of LineDetector::LineFit
(/root/reference/src/LineDetector.cc:713-840) through known answers that follow from the reference's text:
  * a chain on an exact 2-D line whose depths are exactly linear in the distance along it -> ONE segment with the analytic
    end points (:788-823), nothing else (the recursion at :838 finds no second window);
  * the same chain with a 3-D direction closer than MIN_SEGMENT_ANGLE (30 deg) to the viewing ray -> rejected (:809-814);
  * a chain without depth, a chain of 10 pixels (noPixels > initLength fails, :723) -> nothing;
  * a depth discontinuity half way -> the first segment stops at the step (:757-771), a second one starts after it.
"""
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
cv2 = pytest.importorskip("cv2")
import linefit_oracle as LO  # noqa: E402
from sdmb200 import synth  # noqa: E402

K = (500.0, 500.0, 160.0, 120.0)
TWC = np.hstack([np.eye(3, dtype=np.float32), np.array([[1.0], [2.0], [3.0]], np.float32)])
H, W = 240, 320


def planes_for(chain, depth_of_d, sigma=0.01):
    chk = np.zeros((H, W), np.float32)
    sig = np.zeros((H, W), np.float32)
    r0, c0 = chain[0]
    for r, c in chain:
        d = np.hypot(r - r0, c - c0)
        chk[r, c] = np.float32(1.0 / depth_of_d(d))
        sig[r, c] = sigma
    return LO.Planes(chk, sig, K, TWC)


def test_exact_line_gives_one_segment_with_the_analytic_end_points():
    chain = [(100, 40 + i) for i in range(60)]                      # row 100, columns 40..99
    P = planes_for(chain, lambda d: 2.0 + 0.002 * d)                # depth 2.0 .. 2.118 m along the chain
    out = LO.line_fitting(P, [chain])
    assert len(out) == 1
    cid, seg, xyz = out[0]
    assert cid == 0
    assert np.allclose(seg, [40, 100, 99, 100], atol=2e-3)
    fx, fy, cx, cy = K
    ps = np.array([2.0 * (40 - cx) / fx, 2.0 * (100 - cy) / fy, 2.0]) + TWC[:, 3]
    pe = np.array([2.118 * (99 - cx) / fx, 2.118 * (100 - cy) / fy, 2.118]) + TWC[:, 3]
    assert np.allclose(xyz[:3], ps, atol=2e-3) and np.allclose(xyz[3:], pe, atol=2e-3)


def test_segment_along_the_viewing_ray_is_rejected():
    chain = [(120, 170 + i) for i in range(50)]                     # starts 10 px right of the principal point
    P = planes_for(chain, lambda d: 1.0 + 2.0 * d / 49.0)           # depth 1 -> 3 m: 8.5 deg off the ray
    assert LO.line_fitting(P, [chain]) == []


def test_chains_without_depth_or_too_short_give_nothing():
    chain = [(50, 30 + i) for i in range(80)]
    empty = LO.Planes(np.zeros((H, W), np.float32), np.zeros((H, W), np.float32), K, TWC)
    assert LO.line_fitting(empty, [chain]) == []
    short = chain[:10]
    assert LO.line_fitting(planes_for(short, lambda d: 2.0), [short]) == []
    wide_sigma = planes_for(chain, lambda d: 2.0, sigma=0.02)       # sigma < sigma_limit fails (:593-594)
    assert LO.line_fitting(wide_sigma, [chain]) == []


def test_depth_step_splits_the_chain():
    chain = [(60, 20 + i) for i in range(100)]
    P = planes_for(chain, lambda d: 2.0 + 0.004 * d if d < 50 else 3.0 + 0.004 * d)
    out = LO.line_fitting(P, [chain])
    assert len(out) == 2
    (_, s0, _), (_, s1, _) = out
    assert abs(s0[0] - 20) < 0.01 and 60 <= s0[2] <= 70            # stops in the 10-pixel check window holding the step
    assert s1[0] >= 69 and abs(s1[2] - 119) < 0.01


def test_golden_edge_chains_are_consistent_with_the_scene():
    """tests/golden/ed_chains_small.npz (oracle/make_ed_golden.py, the reference's EDLib.a): the images the chains were
    detected on are the ones synth.make_scene still renders; chains are 8-connected pixel walks inside the image."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ed_chains_small.npz"))
    n, w, h, nn, seed = (int(v) for v in g["scene"])
    sc = synth.make_scene(n, w, h, nn, seed=seed)
    assert zlib.crc32(sc.im.tobytes()) == int(g["im_crc"][0])
    for i in range(n):
        off, pix = g[f"off_{i}"], g[f"pix_{i}"]
        assert off[0] == 0 and off[-1] == pix.size and np.all(np.diff(off) > 0)
        r, c = (pix >> 16).astype(np.int64), (pix & 0xffff).astype(np.int64)
        assert r.max() < h and c.max() < w
        step = np.maximum(np.abs(np.diff(r)), np.abs(np.diff(c)))
        inside = np.ones(pix.size - 1, bool)
        inside[off[1:-1] - 1] = False                                # steps across chain boundaries
        assert step[inside].max() <= 1 + 1, "ED chains are (nearly) 8-connected walks"
        assert len(off) - 1 > 100


def test_oracle_agrees_with_the_reference_source_up_to_solver_noise():
    """tests/golden/linefit_ref_small.npz = the REFERENCE'S OWN LineFit text (LineDetector.cc:578-840 compiled where it lies,
    oracle/Makefile target ref_linefit) with exact stand-ins for the two OpenCV solver calls; the python oracle makes the
    real cv2 calls (float SVD through LAPACK).  Same control flow, so the only differences are threshold decisions inside the
    solver noise: >= 97 % of the line-producing chains give identical line lists (two keyframes checked here)."""
    from collections import defaultdict
    from helpers import linefit_dense_planes
    gold = np.load(os.path.join(ROOT, "tests", "golden", "linefit_ref_small.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "ed_chains_small.npz"))
    n, w, h, nn, seed = (int(v) for v in g["scene"])
    sc = synth.make_scene(n, w, h, nn, seed=seed)
    chk, sig = linefit_dense_planes(h, w, n)
    tot = same = 0
    for i in (0, 3):
        off, pix = g[f"off_{i}"], g[f"pix_{i}"]
        chains = [[(int(p >> 16), int(p & 0xffff)) for p in pix[off[k]:off[k + 1]]] for k in range(len(off) - 1)]
        Twc = np.linalg.inv(np.vstack([sc.Tcw[i].astype(np.float64), [0, 0, 0, 1]]))[:3].astype(np.float32)
        mine = defaultdict(list)
        for cid, seg, xyz in LO.line_fitting(LO.Planes(chk[i], sig[i], sc.K, Twc), chains):
            mine[cid].append(np.array(seg, np.float32))
        ref = defaultdict(list)
        for cid, seg in zip(gold[f"dense_chain_{i}"], gold[f"dense_seg_{i}"]):
            ref[int(cid)].append(seg)
        for cid in set(mine) | set(ref):
            tot += 1
            same += len(mine[cid]) == len(ref[cid]) and all(np.abs(a - b).max() <= 0.02 for a, b in zip(mine[cid], ref[cid]))
    assert tot > 300 and same >= 0.97 * tot, (same, tot)
