"""Timing of sdm_line_fit on VGA keyframes with real Edge Drawing chains.

Needs oracle/_ref/ed_chains_vga.npz (python oracle/make_ed_golden.py --bench, here, where the reference's EDLib.a is;
the file travels to the GPU box with gpurun but stays out of git).  Runs the SemiDenseLoop on the keyframes, then fits
the lines of all of them in one call and per keyframe; prints one JSON line.  The oracle (python + cv2) is timed on two
keyframes only to show the order of magnitude of a scalar host implementation - it is not a tuned CPU baseline.
usage: python tools/linefit_bench.py [n_keyframes]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("eao-slam_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
from sdmb200 import api, synth  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "oracle", "_ref", "ed_chains_vga.npz"))
    n, W, H, nn, seed = (int(v) for v in g["scene"])
    n = min(n, int(sys.argv[1])) if len(sys.argv) > 1 else n
    sc = synth.make_scene(int(g["scene"][0]), W, H, nn, seed=seed, workers=8)
    offs = [g[f"off_{i}"] for i in range(n)]
    pix = [g[f"pix_{i}"] for i in range(n)]
    with api.Context(width=W, height=H, max_keyframes=sc.n, intra_check=1, intra_grow=1) as ctx:
        ctx.upload_scene(sc)
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items); ctx.pass2(items); ctx.synchronize()
        ctx.line_fit(list(range(n)), offs, pix)  # warm-up (allocations)
        wall, dev = [], []
        for _ in range(5):
            t = time.perf_counter()
            lines, counts = ctx.line_fit(list(range(n)), offs, pix)
            wall.append((time.perf_counter() - t) * 1e3)
            dev.append(ctx.last_line_fit_ms())
        t = time.perf_counter()
        for i in range(n):
            ctx.line_fit([i], [offs[i]], [pix[i]])
        per_kf_wall = (time.perf_counter() - t) * 1e3 / n
        planes = [ctx.download(i) for i in range(min(n, 8))]
    out = dict(what="sdm_line_fit over real ED chains, one call for all keyframes", keyframes=n, image=f"{W}x{H}",
               chains=int(sum(len(o) - 1 for o in offs)), chain_pixels=int(sum(p.size for p in pix)), lines=int(len(lines)),
               device_ms_per_call=float(np.median(dev)), wall_ms_per_call=float(np.median(wall)),
               device_us_per_keyframe=float(np.median(dev)) * 1e3 / n, wall_ms_per_keyframe_single_calls=per_kf_wall)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_linefit.so")
    if os.path.exists(ref_so):  # the reference's own LineFit text (exact solver stand-ins), one thread like LineFitting (:884-900)
        import ctypes as C
        lib = C.CDLL(ref_so)
        lib.ref_line_fitting.restype = C.c_int
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        m = min(n, 8)
        t, nl = time.perf_counter(), 0
        for i in range(m):
            rc = np.stack([(pix[i] >> 16).astype(np.int32), (pix[i] & 0xffff).astype(np.int32)], 1).copy()
            cap = int((np.diff(offs[i]) // 10).sum()) + 1
            seg, xyz, chn = np.zeros((cap, 4), np.float32), np.zeros((cap, 6), np.float32), np.zeros(cap, np.int32)
            pl = planes[i] if i < len(planes) else None
            if pl is None:
                break
            nl += lib.ref_line_fitting(W, H, np.ascontiguousarray(pl["checked"]).ctypes.data_as(fp), np.ascontiguousarray(pl["sigma"]).ctypes.data_as(fp),
                                       np.asarray(sc.K, np.float32).ctypes.data_as(fp), np.ascontiguousarray(sc.Tcw[i], np.float32).ctypes.data_as(fp),
                                       len(offs[i]) - 1, np.ascontiguousarray(offs[i], np.int32).ctypes.data_as(ip), rc.ctypes.data_as(ip), cap,
                                       seg.ctypes.data_as(fp), xyz.ctypes.data_as(fp), chn.ctypes.data_as(ip))
        out["reference_source_cpu_ms_per_keyframe"] = (time.perf_counter() - t) * 1e3 / m
        out["reference_source_note"] = (f"LineDetector.cc:578-840 compiled where it lies (oracle/_ref/libref_linefit.so, exact solver stand-ins instead "
                                        f"of OpenCV's SVD), one host thread, {m} keyframes, {nl} lines; device lines of the same keyframes: "
                                        f"{int((lines['kf_index'] < m).sum())}")
    try:
        import ctypes as C
        import linefit_oracle as LO
        import oracle_py as O
        t = time.perf_counter()
        for i in range(2):
            twc = np.zeros(16, np.float32)
            tc = np.ascontiguousarray(sc.Tcw[i], np.float32)
            O.lib().oracle_pose_inverse(tc.ctypes.data_as(C.POINTER(C.c_float)), twc.ctypes.data_as(C.POINTER(C.c_float)))
            ch = [[(int(p >> 16), int(p & 0xffff)) for p in pix[i][offs[i][k]:offs[i][k + 1]]] for k in range(len(offs[i]) - 1)]
            LO.line_fitting(LO.Planes(planes[i]["checked"], planes[i]["sigma"], sc.K, twc.reshape(4, 4)[:3]), ch)
        out["oracle_python_cv2_ms_per_keyframe"] = (time.perf_counter() - t) * 1e3 / 2
    except Exception as e:  # noqa: BLE001
        out["oracle_python_cv2_ms_per_keyframe"] = f"unavailable: {e}"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
