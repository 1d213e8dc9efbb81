"""What a bit-mask-driven column walk ("next") would buy on the bench scene - CPU estimate.

For one keyframe of bench.py's trajectory and its N neighbours this replays the scan's geometry in numpy (float64; the
statistics do not depend on last-bit rounding), groups the candidates into warps the way k_pack / k_pass1_lane do (32x8
tiles, row-major inside a tile, 32 consecutive candidates per warp) and counts per warp
    trips_now   = max over lanes of the scanned columns               (what the loop iterates today)
    trips_P     = max over lanes of columns whose row pair may pass gate 1 (G0 > 8 or G1 > 8: the exact superset)
    trips_D     = the same with the mask dilated by one row up and down   (what a word-level lookup would need)
and the lane occupancy of the sections.  usage: python tools/sim_mask_walk.py [keyframe] [n_keyframes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("eao-slam_b200/python", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle_py as O  # noqa: E402  (analysis tool: not part of the product path)
from sdmb200 import synth  # noqa: E402


def main():
    kf = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    sc = synth.make_scene(n, 640, 480, 6, seed=2, workers=4)
    osc = O.OracleScene(sc)
    H, W = sc.shape
    fx, fy, cx, cy = sc.K
    G = sc.grad.astype(np.float64)
    cand = G[kf] > 8
    # warp order: tiles of 32x8 in raster order of tiles (the device's tile order is arbitrary, the grouping is the same)
    ys, xs = np.nonzero(cand)
    key = ((ys // 8) * (W // 32 + 1) + xs // 32) * 256 + (ys % 8) * 32 + xs % 32
    o = np.argsort(key, kind="stable")
    ys, xs = ys[o], xs[o]
    nc = len(ys)
    pad = (-nc) % 32
    tot = dict(now=0, P=0, D=0, lanes_now=0, lanes_pass=0, cols=0, pass1=0, setP=0, setD=0)
    slopes = []
    for j in sc.nbr_idx[kf]:
        pr = osc.pair(kf, int(j))
        F = np.array(pr.F12, np.float64).reshape(3, 3)
        R = np.array(pr.R21, np.float64).reshape(3, 3)
        t = np.array(pr.t21, np.float64)
        a = xs * F[0, 0] + ys * F[1, 0] + F[2, 0]
        b = xs * F[0, 1] + ys * F[1, 1] + F[2, 1]
        c = xs * F[0, 2] + ys * F[1, 2] + F[2, 2]
        ab, cb = a / b, c / b
        xn, yn = (xs - cx) / fx, (ys - cy) / fy
        s0 = R[0, 0] * xn + R[0, 1] * yn + R[0, 2]
        s2 = R[2, 0] * xn + R[2, 1] * yn + R[2, 2]
        u1 = fx * (s0 * sc.min_depth[kf] + t[0]) / (s2 * sc.min_depth[kf] + t[2]) + cx
        u2 = fx * (s0 * sc.max_depth[kf] + t[0]) / (s2 * sc.max_depth[kf] + t[2]) + cx
        umin, umax = np.minimum(u1, u2), np.maximum(u1, u2)
        lo = np.maximum(np.ceil(np.clip(umin, 0, W - 1)), 1).astype(int)
        hi = np.minimum(np.floor(np.clip(umax, 0, W - 1)), W - 2).astype(int)
        ok = (np.abs(ab) <= 4) & (hi >= lo)
        slopes.append(np.abs(ab[ok]))
        L = int(np.max(np.where(ok, hi - lo + 1, 0)))
        k = np.arange(L)[None, :]
        u = lo[:, None] + k
        inside = ok[:, None] & (u <= hi[:, None])
        v = -(ab[:, None] * u + cb[:, None])
        vm, vp = -(ab[:, None] * (u - 1) + cb[:, None]), -(ab[:, None] * (u + 1) + cb[:, None])
        inside &= (v >= 0) & (v <= H - 1) & (vm >= 0) & (vm <= H - 1) & (vp >= 0) & (vp <= H - 1)
        uu = np.clip(u, 0, W - 1)
        y0 = np.clip(np.floor(v), 0, H - 2).astype(int)
        w1 = v - np.floor(v)
        Gj = G[int(j)]
        g0, g1 = Gj[y0, uu], Gj[y0 + 1, uu]
        pass1 = inside & (g0 * (1 - w1) + g1 * w1 > 8)
        setP = inside & ((g0 > 8) | (g1 > 8))
        gm, gp = Gj[np.clip(y0 - 1, 0, H - 1), uu], Gj[np.clip(y0 + 2, 0, H - 1), uu]
        setD = inside & ((g0 > 8) | (g1 > 8) | (gm > 8) | (gp > 8))

        def warps(x):  # per-lane counts -> [n_warps, 32]
            return np.pad(x.sum(1), (0, pad)).reshape(-1, 32)
        wn, wP, wD = warps(inside), warps(setP), warps(setD)
        tot["now"] += wn.max(1).sum()
        tot["P"] += wP.max(1).sum()
        tot["D"] += wD.max(1).sum()
        tot["cols"] += inside.sum()
        tot["pass1"] += pass1.sum()
        tot["setP"] += setP.sum()
        tot["setD"] += setD.sum()
    print(f"keyframe {kf}: {nc} candidates, {len(sc.nbr_idx[kf])} neighbours")
    print(f"columns per (pixel, neighbour): {tot['cols'] / (nc * len(sc.nbr_idx[kf])):.1f}; "
          f"gate-1 pass rate {tot['pass1'] / tot['cols']:.3f}; mask density P {tot['setP'] / tot['cols']:.3f}, "
          f"dilated D {tot['setD'] / tot['cols']:.3f}")
    sl = np.concatenate(slopes)
    print("epipolar slope |a/b|: median %.3f, 90 %% %.3f; rows crossed per 32 columns <= 1: %.2f, <= 2: %.2f, <= 4: %.2f of the pairs"
          % (np.median(sl), np.quantile(sl, 0.9), (sl * 32 <= 1).mean(), (sl * 32 <= 2).mean(), (sl * 32 <= 4).mean()))
    print(f"warp-level trips: now {tot['now']}  (lane occupancy {tot['cols'] / (32 * tot['now']):.3f})")
    print(f"                  P   {tot['P']}  = {tot['P'] / tot['now']:.3f} of now (lane occupancy {tot['setP'] / (32 * tot['P']):.3f})")
    print(f"                  D   {tot['D']}  = {tot['D'] / tot['now']:.3f} of now (lane occupancy {tot['setD'] / (32 * tot['D']):.3f})")


if __name__ == "__main__":
    main()
