"""CPU replay of the skip-distance column walk (third-generation scan loop) and of variants of it, on the bench scene.

For every (candidate pixel, neighbour) the replay walks the search range the way scan_columns3 does - evaluate a column,
advance by min(skip distance of the texel, columns left in the row, columns left in the range) - and counts the visits of
every lane; lanes are grouped into warps exactly like k_pack orders the candidates, and a warp's trip count for a neighbour
is the maximum over its lanes (the lanes re-converge after every neighbour).  ncu of the real kernel (config 2): 78 M
warp-trips per 200 keyframes = 390 k per keyframe at 24 of 32 lanes; the replay of "gen3" should land there.

Variants:
  gen3       skip plane of the texel's own row; the first column after every row change is visited unconditionally
  or2        skip plane that ORs the row pair the line is in with the one it enters next: a jump may cross ONE row boundary
  or2+flag   or2, and a visited column whose own texel cannot survive costs no gate evaluation (reported as 'cheap' visits)
  ideal      only columns that can survive (exact row), no forced visits: the bound of any skip scheme with these bins

usage: python tools/sim_skip_walk.py [keyframe] [n_keyframes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("eao-slam_b200/python", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle_py as O  # noqa: E402  (analysis tool: not part of the product path)
from sdmb200 import synth  # noqa: E402

K = int(os.environ.get("SIM_BINS", "8"))  # orientation bins of the skip planes


def circ_dist(a, b):
    d = np.abs(a - b) % 360.0
    return np.minimum(d, 360.0 - d)


def next_true_distance(m):
    """m [.., W] bool -> uint16 distance from x to the next True strictly after x (W if none)"""
    W = m.shape[-1]
    idx = np.where(m, np.arange(W), 10 ** 6)
    nxt = np.minimum.accumulate(idx[..., ::-1], axis=-1)[..., ::-1]       # next True at or after x
    nxt_after = np.concatenate([nxt[..., 1:], np.full(m.shape[:-1] + (1,), 10 ** 6)], axis=-1)
    return np.minimum(nxt_after - np.arange(W), W).astype(np.int32)


def main():
    kf = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    sc = synth.make_scene(n, 640, 480, 6, seed=2, workers=4)
    osc = O.OracleScene(sc)
    H, W = sc.shape
    fx, fy, cx, cy = sc.K
    G = sc.grad.astype(np.float64)
    TH = sc.theta.astype(np.float64)
    ys, xs = np.nonzero(G[kf] > 8)
    key = ((ys // 8) * (W // 32 + 1) + xs // 32) * 256 + (ys % 8) * 32 + xs % 32
    o = np.argsort(key, kind="stable")
    ys, xs = ys[o], xs[o]
    nc = len(ys)
    pad = (-nc) % 32
    th_pi = TH[kf][ys, xs]
    tot = {}
    flat = {}

    def acc(name, visits, cheap=None):
        v = np.pad(visits, (0, pad)).reshape(-1, 32)
        t = tot.setdefault(name, [0, 0, 0])
        t[0] += int(v.max(1).sum())
        t[1] += int(visits.sum())
        if cheap is not None:
            t[2] += int(cheap.sum())

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
        # valid column interval (rows of u-1, u, u+1 inside the image): shrink from both ends
        def vrow(u):
            return -(ab * u + cb)
        for _ in range(3):  # a few fixed-point steps are enough for these slopes
            bad_lo = ok & ~((vrow(lo - 1) >= 0) & (vrow(lo - 1) <= H - 1) & (vrow(lo + 1) >= 0) & (vrow(lo + 1) <= H - 1) & (vrow(lo) >= 0) & (vrow(lo) <= H - 1))
            lo = lo + bad_lo
            bad_hi = ok & ~((vrow(hi - 1) >= 0) & (vrow(hi - 1) <= H - 1) & (vrow(hi + 1) >= 0) & (vrow(hi + 1) <= H - 1) & (vrow(hi) >= 0) & (vrow(hi) <= H - 1))
            hi = hi - bad_hi
            ok &= hi >= lo
        Gj, Tj = G[int(j)], TH[int(j)]
        S1 = (Gj[:-1] > 8) | (Gj[1:] > 8)
        S1 = np.vstack([S1, S1[-1:]])
        T0, T1 = Tj, np.vstack([Tj[1:], Tj[-1:]])
        wdt = 360.0 / K
        mid = (T0 + ((T1 - T0 + 540.0) % 360.0 - 180.0) / 2.0) % 360.0
        half = np.abs((T1 - T0 + 540.0) % 360.0 - 180.0) / 2.0 + 0.05
        surv = np.zeros((K, H, W), bool)
        for bq in range(K):
            surv[bq] = S1 & (circ_dist(mid, (bq + 0.5) * wdt) < half + wdt / 2 + 45.0)
        apr = (th_pi + sc.rot[kf][jn]) % 360.0
        q = np.minimum((apr / wdt).astype(int), K - 1)
        up = ab < 0  # v grows with u
        d_same = next_true_distance(surv)                                   # [K,H,W]
        surv_up = surv | np.concatenate([surv[:, 1:], surv[:, -1:]], axis=1)   # rows y | y+1, indexed at y
        d_or_up = next_true_distance(surv_up)
        # for lines going down (rows y | y-1) the plane is indexed at y-1
        def walk(mode):
            u = lo.copy()
            act = ok.copy()
            visits = np.zeros(nc, np.int64)
            cheap = np.zeros(nc, np.int64)
            while act.any():
                uu = np.where(act, u, 1)
                v = vrow(uu)
                y = np.clip(np.floor(v), 0, H - 2).astype(int)
                visits += act
                if mode == "gen3":
                    S = d_same[q, y, uu]
                    # columns until the row changes
                    frac = v - np.floor(v)
                    rem = np.where(up, 1 - frac, frac)
                    m1 = np.floor(rem / np.maximum(np.abs(ab), 1e-12) + 1e-9).astype(np.int64) + 1
                    m1 = np.where(np.abs(ab) < 1e-12, 10 ** 6, m1)
                    step = np.maximum(1, np.minimum(S, m1))
                elif mode in ("or2", "or2+flag"):
                    yi = np.where(up, y, np.maximum(y - 1, 0))
                    S = d_or_up[q, yi, uu]
                    frac = v - np.floor(v)
                    rem = np.where(up, 1 - frac, frac) + 1.0   # until the SECOND row change
                    m2 = np.floor(rem / np.maximum(np.abs(ab), 1e-12) + 1e-9).astype(np.int64) + 1
                    m2 = np.where(np.abs(ab) < 1e-12, 10 ** 6, m2)
                    step = np.maximum(1, np.minimum(S, m2))
                    if mode == "or2+flag":
                        cheap += act & ~surv[q, y, uu]
                else:  # ideal: jump to the next column whose own-row texel can survive (oracle knowledge of the rows ahead)
                    step = np.ones(nc, np.int64)
                u = u + step
                act &= u <= hi
            return visits, cheap
        for mode in ("gen3", "or2", "or2+flag"):
            vis, ch = walk(mode)
            acc(mode, vis, ch if mode == "or2+flag" else None)
            if mode == "gen3":
                flat["gen3"] = flat.get("gen3", 0) + vis
        # ideal: count surviving columns per lane directly
        L = int(np.max(np.where(ok, hi - lo + 1, 0)))
        k = np.arange(L)[None, :]
        ucol = lo[:, None] + k
        ins = ok[:, None] & (ucol <= hi[:, None])
        uu = np.clip(ucol, 0, W - 1)
        vv = -(ab[:, None] * uu + cb[:, None])
        yy = np.clip(np.floor(vv), 0, H - 2).astype(int)
        sv = ins & surv[q[:, None], yy, uu]
        acc("ideal", sv.sum(1))
        flat["ideal"] = flat.get("ideal", 0) + sv.sum(1)
        acc("all columns (gen2)", ins.sum(1))
    for name, vis in flat.items():  # lanes re-converge only after all neighbours (one flattened walk per lane)
        v = np.pad(vis, (0, pad)).reshape(-1, 32)
        tot[name + ", no sync between neighbours"] = [int(v.max(1).sum()), int(vis.sum()), 0]
    print(f"keyframe {kf}: {nc} candidates")
    base = tot["gen3"][0]
    print(f"{'design':44s} {'warp-trips':>11s} {'vs gen3':>8s} {'lanes/trip':>10s} {'visits/pair':>11s} {'cheap share':>11s}")
    for name, (tr, vis, ch) in tot.items():
        print(f"{name:44s} {tr:11d} {tr / base:8.3f} {vis / max(tr, 1):10.1f} {vis / (nc * 6):11.1f} {ch / max(vis, 1):11.2f}")


if __name__ == "__main__":
    main()
