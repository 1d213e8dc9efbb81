"""Generate tests/golden/linefit_ref_small.npz: the lines the REFERENCE'S OWN LineDetector::LineFit finds
(oracle/_ref/libref_linefit.so = the text of /root/reference/src/LineDetector.cc:578-840 compiled where it lies, with exact
stand-ins for cv::SVD::solveZ / cv::solve, refshim/cvstub_linefit.h) on
  dense_*  the dense test planes of tests/helpers.linefit_dense_planes and
  mask_*   the planes the (CPU oracle's) SemiDenseLoop leaves on the scene with the real Edge Drawing mask,
both over the committed Edge Drawing chains (tests/golden/ed_chains_small.npz).  The GPU test compares sdm_line_fit with
these rows: the kernel's control flow against the reference's text, solver noise taken out.  TEST INFRASTRUCTURE; runs only
where /root/reference exists.  Run: make -C oracle ref_linefit && python oracle/make_linefit_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(HERE, "..", "eao-slam_b200", "python"), HERE, os.path.join(HERE, "..", "tests")):
    sys.path.insert(0, p)
from helpers import edge_index_from_chains, linefit_dense_planes, run_oracle  # noqa: E402
from sdmb200 import synth  # noqa: E402


def ref_lines(lib, checked, sigma, K, Tcw, off, pix):
    rc = np.stack([(pix >> 16).astype(np.int32), (pix & 0xffff).astype(np.int32)], 1).copy()
    cap = int((np.diff(off) // 10).sum()) + 1
    seg, xyz, ch = np.zeros((cap, 4), np.float32), np.zeros((cap, 6), np.float32), np.zeros(cap, np.int32)
    H, W = checked.shape
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    n = lib.ref_line_fitting(W, H, np.ascontiguousarray(checked, np.float32).ctypes.data_as(fp),
                             np.ascontiguousarray(sigma, np.float32).ctypes.data_as(fp),
                             np.asarray(K, np.float32).ctypes.data_as(fp), np.ascontiguousarray(Tcw, np.float32).ctypes.data_as(fp),
                             len(off) - 1, np.ascontiguousarray(off, np.int32).ctypes.data_as(ip), rc.ctypes.data_as(ip), cap,
                             seg.ctypes.data_as(fp), xyz.ctypes.data_as(fp), ch.ctypes.data_as(ip))
    assert n >= 0
    return seg[:n], xyz[:n], ch[:n]


def main(out):
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_linefit.so"))
    lib.ref_line_fitting.restype = C.c_int
    g = np.load(os.path.join(HERE, "..", "tests", "golden", "ed_chains_small.npz"))
    n, W, H, nn, seed = (int(v) for v in g["scene"])
    sc = synth.make_scene(n, W, H, nn, seed=seed)
    offs, pixs = [g[f"off_{i}"] for i in range(n)], [g[f"pix_{i}"] for i in range(n)]
    d = {}
    chk, sig = linefit_dense_planes(H, W, n)
    masked = synth.Scene(im=sc.im, grad=sc.grad, theta=sc.theta,
                         edge=np.stack([edge_index_from_chains(offs[i], pixs[i], H, W) for i in range(n)]), K=sc.K, Tcw=sc.Tcw,
                         nbr_idx=sc.nbr_idx, rot=sc.rot, min_depth=sc.min_depth, max_depth=sc.max_depth)
    osc = run_oracle(masked)
    for tag, planes in (("dense", (chk, sig)), ("mask", (osc.checked, osc.sigma))):
        tot = 0
        for i in range(n):
            seg, xyz, ch = ref_lines(lib, planes[0][i], planes[1][i], sc.K, sc.Tcw[i], offs[i], pixs[i])
            d[f"{tag}_seg_{i}"], d[f"{tag}_xyz_{i}"], d[f"{tag}_chain_{i}"] = seg, xyz, ch
            tot += len(seg)
        print(tag, tot, "lines")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "tests", "golden", "linefit_ref_small.npz"))
