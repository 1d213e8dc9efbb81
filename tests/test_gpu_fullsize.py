"""BASELINE config[1] at FULL size (200 keyframes, 640x480, 6 neighbours, intra + inter checks) through the C-ABI.

The oracle needs ~1 s per keyframe on one core, so the whole 200-keyframe loop is checked through
size-independent properties and a sample of keyframes is compared with the oracle bit for bit:
  * sampled keyframes (first, interior, last of the trajectory): pass 1 (+ intra) of the keyframe and of its six
    neighbours, then pass 2 of the keyframe, by the oracle -> all four planes identical;
  * determinism: a second run of the loop gives the same checksum of checksums over all 800 planes;
  * batch independence: issuing the loop in chunks of 10 / 7 keyframes (what bench.py's e2e leg does) or with the
    work items in reversed order gives the same planes as one batch (pass 1 reads only inputs, pass 2 only pass-1
    outputs: ProbabilityMapping.cc:447-489, :1202-1249);
  * structural invariants of the reference's planes: sigma == 0 wherever depth == 0, the 2-pixel border of
    depth_map_checked_ untouched (= 0, KeyFrame.cc:79), points == 0 wherever checked < 1e-6 (:700-731), every
    accepted pixel a candidate (GradImg > 8, :454-456) when the growing stage is off;
  * the device counters (candidates / fused / checked) equal the counts over the downloaded planes.
"""
from __future__ import annotations

import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

N_KF, N_NBR, SEED = 200, 6, 2


@pytest.fixture(scope="module")
def scene200():
    """bench.py's config[1] scene.  Rendered in a fresh interpreter (the generator forks workers, which must not
    happen in a process that already holds a CUDA context) into bench.py's scene cache, then loaded from there."""
    import bench
    from sdmb200 import synth
    nb = synth.neighbours(N_KF, N_NBR)
    code = ("import bench; from sdmb200 import synth; "
            f"bench.load_scene({N_KF}, 0, synth.neighbours({N_KF}, {N_NBR}), {SEED}, 'gpu')")
    subprocess.run([sys.executable, "-c", code], cwd=ROOT, check=True, timeout=600)
    return bench.load_scene(N_KF, 0, nb, SEED, "gpu")


def _loop(ctx, scene, api, order=None, chunk=None):
    """the whole loop; returns the four planes of every keyframe"""
    n = scene.n
    order = list(range(n)) if order is None else list(order)
    chunk = chunk or n
    stats = {"candidates": 0, "fused": 0, "checked": 0}  # the device counters are per call: add the chunks up
    for c0 in range(0, n, chunk):
        ctx.pass1(api.make_items(order[c0:c0 + chunk], scene.nbr_idx, scene.rot, scene.min_depth, scene.max_depth))
        st = ctx.stats()
        stats["candidates"] += st["candidates"]
        stats["fused"] += st["fused"]
    for c0 in range(0, n, chunk):
        ctx.pass2(api.make_items(order[c0:c0 + chunk], scene.nbr_idx, scene.rot, scene.min_depth, scene.max_depth))
        stats["checked"] += ctx.stats()["checked"]
    H, W = scene.shape
    out = {k: np.zeros((n, H, W) + ((3,) if k == "points" else ()), np.float32)
           for k in ("depth", "sigma", "checked", "points")}
    for i in range(n):
        r = ctx.download(i)
        for k in out:
            out[k][i] = r[k]
    return out, stats


def _checksum(out):
    """checksum of the per-plane checksums (bit patterns, so -0 / NaN payloads count)"""
    acc = 0
    for k in ("depth", "sigma", "checked", "points"):
        for i in range(out[k].shape[0]):
            acc = zlib.crc32(np.uint32(zlib.crc32(out[k][i].view(np.uint32))).tobytes(), acc)
    return acc


@pytest.fixture(scope="module")
def full_run(scene200):
    from sdmb200 import api
    H, W = scene200.shape
    ctx = api.Context(width=W, height=H, max_keyframes=N_KF, intra_check=1, intra_grow=1)
    ctx.upload_scene(scene200)
    out, stats = _loop(ctx, scene200, api)
    yield ctx, out, stats
    ctx.close()


def test_sampled_keyframes_equal_the_oracle(scene200, full_run):
    import oracle_py as O
    _, out, _ = full_run
    osc = O.OracleScene(scene200, "canonical")
    prm = O.default_params("canonical", intra_check=1, intra_grow=1)
    done = set()
    for s in (0, 97, N_KF - 1):
        need = [s] + [int(j) for j in scene200.nbr_idx[s]]
        for k in need:  # pass 1 (+ intra) of the keyframe and its neighbours
            if k not in done:
                osc.run(params=prm, first=k, count=1, pass_mask=1)
                done.add(k)
        osc.run(params=prm, first=s, count=1, pass_mask=2)
        for name, ref in (("depth", osc.depth), ("sigma", osc.sigma), ("checked", osc.checked), ("points", osc.points)):
            a, b = out[name][s].view(np.uint32), ref[s].view(np.uint32)
            assert int((a != b).sum()) == 0, f"keyframe {s} plane {name}: {int((a != b).sum())} words differ"
        assert int((osc.checked[s] > 0).sum()) > 10000  # the sample is not trivially empty
    # pass-1 planes of the neighbours as well (they were computed anyway)
    for k in sorted(done):
        assert np.array_equal(out["depth"][k].view(np.uint32), osc.depth[k].view(np.uint32))
        assert np.array_equal(out["sigma"][k].view(np.uint32), osc.sigma[k].view(np.uint32))


def test_structural_invariants_and_counters(scene200, full_run):
    _, out, stats = full_run
    d, s, c, p = out["depth"], out["sigma"], out["checked"], out["points"]
    assert not np.isnan(d).any() and not np.isnan(s).any() and not np.isnan(c).any()
    assert np.all(s[d == 0] == 0)
    border = np.ones(c.shape[1:], bool)
    border[2:-2, 2:-2] = False
    assert np.all(c[:, border] == 0) and np.all(p[:, border] == 0)
    assert np.all(p[c < 1e-6] == 0)
    assert np.all(d[c > 0] > 1e-6)  # pass 2 only keeps pixels that had a pass-1 depth (:1150)
    # the growing stage only fills pixels with GradImg > 8 (:929-976); pass 1 only candidates (:454-456)
    assert np.all(scene200.grad[d > 0] > 8)
    assert stats["candidates"] == int((scene200.grad > 8).sum())
    assert stats["checked"] == int((c > 0).sum())
    assert 0.5 * stats["candidates"] < stats["fused"] <= stats["candidates"]
    assert 0.5 * stats["candidates"] < int((d > 0).sum()) <= stats["candidates"]


def test_rerun_is_deterministic_and_batching_does_not_matter(scene200, full_run):
    from sdmb200 import api
    ctx, out, _ = full_run
    ref = _checksum(out)
    again, _ = _loop(ctx, scene200, api)
    assert _checksum(again) == ref
    del again
    chunked, st = _loop(ctx, scene200, api, chunk=10)
    assert _checksum(chunked) == ref
    assert st == full_run[2]
    del chunked
    rev, _ = _loop(ctx, scene200, api, order=reversed(range(N_KF)), chunk=7)
    assert _checksum(rev) == ref
