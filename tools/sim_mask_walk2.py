"""Sizing of mask-driven column walks for the scan kernel (round 2) - CPU replay on the bench scene.

Extends tools/sim_mask_walk.py: besides the gate-1 superset it evaluates masks that also fold in a superset of gate 3
(orientation bins: the texel's orientation arc against a 360/K-degree bin of the pixel's th_pi + rot), looked up either
at the exact row of every column or - what a kernel can afford - once per window of `win` columns from a plane that
ORs two vertically adjacent row pairs (valid when the line crosses at most one row boundary inside the window; windows
with more crossings fall back to all ones).  For every design it prints the warp-level trip count relative to today's
loop and a modelled instruction count per warp (front / gate sections executed whenever ONE lane needs them).

usage: python tools/sim_mask_walk2.py [keyframe] [n_keyframes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("eao-slam_b200/python", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle_py as O  # noqa: E402  (analysis tool: not part of the product path)
from sdmb200 import synth  # noqa: E402

# modelled warp-instructions: front, yangle + gate 3, gate 2, residual, per-trip overhead (loop + reconvergence)
F_NOW, G3, G2, RES, OVH = 16, 9, 7, 15, 4.6
F_MASK = 24  # front of a bit-driven walk (find / clear bit, column from bit, refill test)


def circ_dist(a, b):
    d = np.abs(a - b) % 360.0
    return np.minimum(d, 360.0 - d)


def main():
    kf = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    sc = synth.make_scene(n, 640, 480, 6, seed=2, workers=4)
    osc = O.OracleScene(sc)
    H, W = sc.shape
    fx, fy, cx, cy = sc.K
    G = sc.grad.astype(np.float64)
    TH = sc.theta.astype(np.float64)
    cand = G[kf] > 8
    ys, xs = np.nonzero(cand)
    key = ((ys // 8) * (W // 32 + 1) + xs // 32) * 256 + (ys % 8) * 32 + xs % 32
    o = np.argsort(key, kind="stable")
    ys, xs = ys[o], xs[o]
    nc = len(ys)
    pad = (-nc) % 32
    th_pi = TH[kf][ys, xs]

    designs = {}
    stats = {}

    def acc(name, trips, cost):
        d = designs.setdefault(name, [0.0, 0.0])
        d[0] += trips
        d[1] += cost

    cols_total = 0
    for jn, j in enumerate(sc.nbr_idx[kf]):
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
        L = int(np.max(np.where(ok, hi - lo + 1, 0)))
        L = (L + 31) // 32 * 32
        k = np.arange(L)[None, :]
        u = lo[:, None] + k
        inside = ok[:, None] & (u <= hi[:, None])
        v = -(ab[:, None] * u + cb[:, None])
        vm, vp = -(ab[:, None] * (u - 1) + cb[:, None]), -(ab[:, None] * (u + 1) + cb[:, None])
        inside &= (v >= 0) & (v <= H - 1) & (vm >= 0) & (vm <= H - 1) & (vp >= 0) & (vp <= H - 1)
        uu = np.clip(u, 0, W - 1)
        y0 = np.clip(np.floor(v), 0, H - 2).astype(int)
        w1 = v - np.floor(v)
        Gj, Tj = G[int(j)], TH[int(j)]
        g0, g1 = Gj[y0, uu], Gj[y0 + 1, uu]
        a0, a1 = Tj[y0, uu], Tj[y0 + 1, uu]
        # yangle
        wrapd = np.abs(a0 - a1) >= 180
        b0 = np.where(wrapd & (a0 < a1), a0 + 360, a0)
        b1 = np.where(wrapd & (a0 >= a1), a1 + 360, a1)
        gth = (b0 * (1 - w1) + b1 * w1) % 360.0
        th_line = np.degrees(np.arctan2(-ab, 1.0)) % 360.0
        apr = (th_pi + sc.rot[kf][jn]) % 360.0
        p1 = inside & (g0 * (1 - w1) + g1 * w1 > 8)
        p3 = p1 & (circ_dist(gth, apr[:, None]) < 45)
        d2 = circ_dist(gth, th_line[:, None])
        d2 = np.minimum(d2, 180 - d2)
        p2 = p3 & (d2 < 80)
        cols_total += inside.sum()

        # texel-level supersets
        S1 = (Gj[:-1] > 8) | (Gj[1:] > 8)  # row pair (y, y+1) may pass gate 1   [H-1, W]
        S1 = np.vstack([S1, S1[-1:]])
        T0, T1 = Tj, np.vstack([Tj[1:], Tj[-1:]])

        def bin_mask(K):
            """[K, H, W]: texel's orientation arc (short way from T0 to T1, +- margin) meets (bin_lo - 45, bin_hi + 45)"""
            wdt = 360.0 / K
            mid = (T0 + ((T1 - T0 + 540.0) % 360.0 - 180.0) / 2.0) % 360.0
            half = np.abs((T1 - T0 + 540.0) % 360.0 - 180.0) / 2.0 + 0.01
            out = np.zeros((K,) + T0.shape, bool)
            for bq in range(K):
                cen = (bq + 0.5) * wdt
                out[bq] = circ_dist(mid, cen) < half + wdt / 2 + 45.0
            return out

        def per_warp(x):
            return np.pad(x, ((0, pad), (0, 0))).reshape(-1, 32, x.shape[1])

        def model(name, setm, front, sync_win=None):
            """setm [nc, L] bool: columns a lane still has to evaluate (superset of p1 [and p3])."""
            assert not (p1 & ~setm & (name.find("g3") < 0)).any() or True
            cnt = setm.sum(1)
            W_set = per_warp(setm)
            # lane-asynchronous walk: lane's t-th set column
            order = np.argsort(~setm, axis=1, kind="stable")  # set columns first, in column order
            tmax = int(cnt.max()) if cnt.size else 0
            sel = order[:, :tmax]
            valid = np.arange(tmax)[None, :] < cnt[:, None]
            q1 = np.take_along_axis(p1, sel, 1) & valid
            q3 = np.take_along_axis(p3, sel, 1) & valid
            q2 = np.take_along_axis(p2, sel, 1) & valid
            wv, w1_, w3_, w2_ = per_warp(valid), per_warp(q1), per_warp(q3), per_warp(q2)
            trips = wv.any(1).sum()
            cost = (front + OVH) * trips + G3 * w1_.any(1).sum() + G2 * w3_.any(1).sum() + RES * w2_.any(1).sum()
            acc(name + " async", trips, cost)
            st = stats.setdefault(name, [0, 0, 0, 0, 0])
            for i_, x_ in enumerate((trips, w1_.any(1).sum(), w3_.any(1).sum(), w2_.any(1).sum(), wv.sum())):
                st[i_] += x_
            if sync_win:
                nw = L // sync_win
                pc = W_set.reshape(W_set.shape[0], 32, nw, sync_win).sum(3)  # [warps, 32, nw]
                acc(name + f" sync{sync_win} (trips only)", pc.max(1).sum(), 0)

        model("now", inside, F_NOW)
        rowsel = (y0, uu)
        model("exact-row g1", inside & S1[rowsel], F_MASK, 32)
        for K in (8, 16):
            B = bin_mask(K)
            bq = np.minimum((apr / (360.0 / K)).astype(int), K - 1)
            Bsel = B[bq[:, None], y0, uu]
            exact = inside & S1[rowsel] & Bsel
            assert not (p3 & ~exact).any(), "orientation-bin mask is not a superset"
            model(f"exact-row g1&g3 K={K}", exact, F_MASK, 32)
            # window-level lookup from the 2-row OR plane at row min(y_first, y_last) of the window
            M = S1[None] & B  # [K,H,W]
            M2 = M | np.concatenate([M[:, 1:], M[:, -1:]], axis=1)
            for win in (16, 32):
                nw = L // win
                yw = y0.reshape(nc, nw, win)
                iw = inside.reshape(nc, nw, win)
                big = np.where(iw, yw, 10 ** 6).min(2)
                sml = np.where(iw, yw, -1).max(2)
                base = np.minimum(big, H - 1)
                flat = (sml - big) <= 1
                rows = np.repeat(base[:, :, None], win, 2).reshape(nc, L)
                look = M2[bq[:, None], np.clip(rows, 0, H - 1), uu]
                look = np.where(np.repeat(flat[:, :, None], win, 2).reshape(nc, L), look, True)
                sm = inside & look
                assert not (p3 & ~sm).any()
                model(f"window{win} OR2 g1&g3 K={K}", sm, F_MASK, win)
            if K == 8:
                M1 = S1 | np.vstack([S1[1:], S1[-1:]])
                for win in (16, 32):
                    nw = L // win
                    yw = y0.reshape(nc, nw, win)
                    iw = inside.reshape(nc, nw, win)
                    big = np.where(iw, yw, 10 ** 6).min(2)
                    sml = np.where(iw, yw, -1).max(2)
                    base = np.minimum(big, H - 1)
                    flat = (sml - big) <= 1
                    rows = np.repeat(base[:, :, None], win, 2).reshape(nc, L)
                    look = M1[np.clip(rows, 0, H - 1), uu]
                    look = np.where(np.repeat(flat[:, :, None], win, 2).reshape(nc, L), look, True)
                    model(f"window{win} OR2 g1 only", inside & look, F_MASK, win)

    base_trips, base_cost = designs["now async"]
    print(f"keyframe {kf}: {nc} candidates, {len(sc.nbr_idx[kf])} neighbours, {cols_total / (nc * 6):.1f} columns per pair")
    print(f"{'design':44s} {'trips':>8s} {'cost/now':>9s}")
    for name, (tr, co) in designs.items():
        print(f"{name:44s} {tr / base_trips:8.3f} {co / base_cost if co else float('nan'):9.3f}")
    print("section execution per trip (gate-3 section, gate-2 section, residual) and lanes per trip:")
    for name, st in stats.items():
        print(f"{name:44s} {st[1] / st[0]:.3f} {st[2] / st[0]:.3f} {st[3] / st[0]:.3f}   {st[4] / st[0]:.1f}")


if __name__ == "__main__":
    main()
