"""Shared helpers of the parity tests: run the oracle and the device path on one scene, compare."""
from __future__ import annotations

import numpy as np

import oracle_py as O
from sdmb200 import api

# BASELINE.json north_star tolerances
SET_MISMATCH_MAX = 1e-3   # accepted-pixel sets: at most 0.1 % tie / boundary cases
REL_TOL = 1e-4            # fused inverse depth and variance: 1e-4 relative


def run_oracle(scene, intra_check=0, intra_grow=0, kind="canonical", **params):
    osc = O.OracleScene(scene, kind)
    p = O.default_params(kind, intra_check=intra_check, intra_grow=intra_grow, **params)
    osc.run(params=p)
    return osc


def run_device(scene, intra_check=0, intra_grow=0, ctx=None, **cfg):
    """Full SemiDenseLoop through the C-ABI: upload, pass 1 (all), pass 2 (all), download."""
    H, W = scene.shape
    own = ctx is None
    if own:
        ctx = api.Context(width=W, height=H, max_keyframes=scene.n, intra_check=intra_check, intra_grow=intra_grow, **cfg)
    ctx.upload_scene(scene)
    items = api.make_items(range(scene.n), scene.nbr_idx, scene.rot, scene.min_depth, scene.max_depth)
    ctx.pass1(items)
    st1 = ctx.stats()
    ctx.pass2(items)
    st2 = ctx.stats()
    out = {k: np.zeros((scene.n, H, W) + ((3,) if k == "points" else ()), np.float32)
           for k in ("depth", "sigma", "checked", "points")}
    for i in range(scene.n):
        r = ctx.download(i)
        for k in out:
            out[k][i] = r[k]
    out["stats"] = {"candidates": st1["candidates"], "fused": st1["fused"], "checked": st2["checked"]}
    if own:
        ctx.close()
    return out


def rel_err(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


def compare_planes(dev, osc, what=("depth", "sigma", "checked", "points")):
    """Returns a report dict; asserts the BASELINE tolerances."""
    rep = {}
    ref = {"depth": osc.depth, "sigma": osc.sigma, "checked": osc.checked, "points": osc.points}
    for setname, key in (("pass1", "depth"), ("pass2", "checked")):
        if key not in what:
            continue
        a, b = dev[key] > 0, ref[key] > 0
        n_ref = int(b.sum())
        mism = int((a != b).sum())
        rep[setname + "_accepted_ref"] = n_ref
        rep[setname + "_set_mismatch"] = mism
        assert n_ref > 0, "oracle accepted nothing: the parity check would be vacuous"
        assert mism <= SET_MISMATCH_MAX * n_ref, (setname, mism, n_ref)
    for key in what:
        a, b = dev[key], ref[key]
        gate = "depth" if key in ("depth", "sigma") else "checked"
        both = (dev[gate] > 0) & (ref[gate] > 0)
        if key == "points":
            both = both[..., None] & np.ones(3, bool)
        if both.any():
            e = rel_err(a[both], b[both])
            if key == "points":  # world coordinates pass through 0: scale by the point norm instead
                nrm = np.linalg.norm(ref["points"], axis=-1, keepdims=True).astype(np.float64)
                e = (np.abs(a.astype(np.float64) - b) / np.maximum(nrm, 1e-30))[both]
            rep[key + "_max_rel"] = float(e.max())
            assert e.max() <= REL_TOL, (key, float(e.max()))
        rep[key + "_bit_mismatch"] = int((a.view(np.uint32) != b.view(np.uint32)).sum())
    return rep


def golden_scene(golden_dir):
    """Scene + oracle outputs committed in tests/golden/oracle_small.npz (oracle/make_golden.py)."""
    import os
    from sdmb200 import synth
    g = np.load(os.path.join(golden_dir, "oracle_small.npz"))
    sc = synth.Scene(im=g["im"], grad=g["grad"], theta=g["theta"], edge=None, K=tuple(float(v) for v in g["K"]),
                     Tcw=g["Tcw"], nbr_idx=g["nbr_idx"], rot=g["rot"], min_depth=g["min_depth"], max_depth=g["max_depth"])
    return sc, g


class GoldenRef:
    """Quacks like OracleScene for compare_planes()."""

    def __init__(self, g, tag):
        self.depth, self.sigma = g[f"{tag}_depth"], g[f"{tag}_sigma"]
        self.checked, self.points = g[f"{tag}_checked"], g[f"{tag}_points"]


def planes_that_grow(depth, sigma, seed):
    """(rho, sigma) planes on which IntraKeyFrameDepthGrowing (:929-976) is NOT a no-op: holes punched into a pass-1
    result keep a non-zero sigma (the invariant `rho = 0 => sigma = 0` of the shipped pipeline is broken on purpose),
    some with rho just below the 1e-6 threshold, some with sigma = +-0 (those can never grow)."""
    rng = np.random.default_rng(seed)
    d, s = depth.copy(), sigma.copy()
    filled = d > 0
    hole = filled & (rng.random(d.shape) < 0.25)
    kind = rng.integers(0, 4, d.shape)
    s = np.where(hole & (kind == 0), np.float32(0.0), s)
    s = np.where(hole & (kind == 1), np.float32(-0.0), s)
    s = np.where(hole & (kind == 2), s * np.float32(3.0), s)        # wide sigma: neighbours become compatible
    s = np.where(hole & (kind == 3), np.float32(0.5), s)
    d = np.where(hole, np.where(rng.random(d.shape) < 0.5, np.float32(0.0), np.float32(9.9e-7)), d)
    return np.ascontiguousarray(d, np.float32), np.ascontiguousarray(s, np.float32)


def linefit_dense_planes(H, W, n, seed=5):
    """depth_map_checked_ / depth_sigma_ planes on which LineFit finds thousands of lines: piecewise-planar inverse depth
    (80 x 80 blocks) with 0.2 % noise, 30 % holes and 10 % wide-sigma pixels.  Shared by tests/test_gpu_linefit.py and
    oracle/make_linefit_golden.py (same seed = same planes)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    bx, by = (xx // 80).astype(int) % 5, (yy // 80).astype(int) % 4
    chk, sig = [], []
    for _ in range(n):
        a, b = rng.uniform(-1e-3, 1e-3, (2, 4, 5)).astype(np.float32)
        c0 = rng.uniform(0.3, 1.2, (4, 5)).astype(np.float32)
        inv = (c0[by, bx] + a[by, bx] * (xx % 80) + b[by, bx] * (yy % 80)).astype(np.float32)
        inv *= (1 + rng.normal(0, 2e-3, inv.shape)).astype(np.float32)
        chk.append(np.where(rng.random(inv.shape) < 0.3, 0, inv).astype(np.float32))
        sig.append(np.where(rng.random(inv.shape) < 0.1, 0.03, 0.01).astype(np.float32))
    return chk, sig


def edge_index_from_chains(off, pix, H, W):
    """LineDetector::DetectEdgeMap's mask (LineDetector.cc:857-866): mEdgeIndex starts at -1 (KeyFrame.h:174) and
    mEdgeIndex(r, c) = i for every pixel of chain i, in chain order (a later chain overwrites an earlier one)"""
    e = np.full((H, W), -1, np.int32)
    ids = np.repeat(np.arange(len(off) - 1, dtype=np.int32), np.diff(off))
    e[(pix >> 16).astype(np.int64), (pix & 0xffff).astype(np.int64)] = ids  # numpy assigns in order: last write wins
    return e
