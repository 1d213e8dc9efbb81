"""Timing of sdm_edge_drawing (LineDetector::DetectEdgeMap for a batch of keyframes: k_ed_planes on the device + routing on host
threads) on VGA keyframes of the bench trajectory, beside the same implementation run entirely on one host thread.
    python tools/ed_bench.py [--n 200] [--out gpurun_out/ed_bench.json]
Prints one JSON object: per thread count the wall time per keyframe, the device time of k_ed_planes per keyframe with its
share of the HBM roofline (4 algorithmic bytes per pixel: 1 read, 3 written), the routing thread time per keyframe."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eao-slam_b200", "python"))
from sdmb200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--out", default=None)
    ap.add_argument("--n-device", type=int, nargs="*", default=[200, 1000], help="batch sizes of the device-routing runs")
    a = ap.parse_args()
    W, H = 640, 480
    sc = synth.make_scene(32, W, H, 6, seed=2, workers=16)
    ims = np.concatenate([sc.im, sc.im[:, ::-1], sc.im[:, :, ::-1], sc.im[:, ::-1, ::-1]] * ((a.n + 127) // 128))[:a.n]
    ims = np.ascontiguousarray(ims)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6552.3))
    res = dict(workload=f"{a.n} keyframes {W}x{H}", hbm_peak_gbps=hbm, runs=[])
    with api.Context(width=W, height=H, max_keyframes=2) as ctx:
        ctx.edge_drawing(ims[:16])  # warm-up: buffers, threads
        ctx.edge_drawing(ims)
        n_chains = None
        for threads in (1, 4, 8, 16, 32, 0):
            best = None
            for _ in range(3):
                offs, pix, _e = ctx.edge_drawing(ims, n_threads=threads, edge_index=True)
                t = ctx.last_edge_drawing_ms()
                if best is None or t["wall_ms"] < best["wall_ms"]:
                    best = t
            n_chains = int(sum(len(o) - 1 for o in offs))
            k_us = best["kernel_ms"] * 1e3 / a.n
            res["runs"].append(dict(threads=threads, wall_ms_per_kf=best["wall_ms"] / a.n, kernel_us_per_kf=k_us,
                                    kernel_gbps=4.0 * W * H / (k_us * 1e-6) / 1e9, kernel_roofline_frac=4.0 * W * H / (k_us * 1e-6) / 1e9 / hbm,
                                    route_thread_ms_per_kf=best["route_thread_ms"] / a.n))
        res["chains"] = n_chains
    # stage 2 on the device (k_ed_route, one warp per keyframe): wall time per keyframe against the batch size; the edge-index
    # planes land in pinned memory (the host-thread runs above write them from the routing threads)
    res["device_route"] = []
    for nd in a.n_device:
        imd = np.ascontiguousarray(np.concatenate([ims] * ((nd + len(ims) - 1) // len(ims)))[:nd])
        with api.Context(width=W, height=H, max_keyframes=2) as ctx:
            ctx.set_edge_drawing_route(True)
            ptr = api.C.c_void_p()
            ctx._chk(ctx.lib.sdm_host_alloc(api.C.byref(ptr), nd * H * W * 4))
            edge = np.ctypeslib.as_array(api.C.cast(ptr, api.C.POINTER(api.C.c_int32)), shape=(nd, H, W))
            ctx.edge_drawing(imd[:8], edge_index=False)
            best = None
            for _ in range(3):
                offs, pix, _e = ctx.edge_drawing(imd, edge_out=edge)
                t = ctx.last_edge_drawing_ms()
                if best is None or t["wall_ms"] < best["wall_ms"]:
                    best = t
            res["device_route"].append(dict(keyframes=nd, wall_ms_per_kf=best["wall_ms"] / nd, route_kernel_ms=best["route_thread_ms"],
                                            stage1_kernel_us_per_kf=best["kernel_ms"] * 1e3 / nd, fallbacks=ctx.last_edge_drawing_fallbacks(),
                                            chains=int(sum(len(o) - 1 for o in offs))))
            del edge
            ctx.lib.sdm_host_free(ptr)
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "ed")
        subprocess.run(["g++", "-O2", "-std=c++11", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_edge_drawing.cpp")], check=True)
        sub = ims[:64]
        sub.tofile(os.path.join(tmp, "in.raw"))
        r = subprocess.run([exe, str(W), str(H), str(len(sub)), os.path.join(tmp, "in.raw"), os.path.join(tmp, "out.bin")],
                           check=True, capture_output=True, text=True)
        res["host_only_one_thread_ms_per_kf"] = float(re.search(r"host_ms_per_image ([0-9.]+)", r.stderr).group(1))
    s = json.dumps(res)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
