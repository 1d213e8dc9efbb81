"""sdm_run_loop (the whole SemiDenseLoop, ProbabilityMapping.cc:348-597, issued as a chunked pipeline inside the library)
must leave exactly the planes of the step-by-step calls: same kernels on the same inputs, only the issue order differs."""
import ctypes as C

import numpy as np
import pytest

from helpers import run_device, run_oracle
from sdmb200 import api, synth

pytestmark = pytest.mark.gpu


def _descs(out, slots, W, first, second):
    n = len(slots)
    d = (api.DownloadDesc * n)()
    for j, s in enumerate(slots):
        d[j].kf = int(s)
        for name in first + second:
            a = out[name][j]
            setattr(d[j], name, a.ctypes.data)
            setattr(d[j], name + "_step", a.strides[0])
    return d


@pytest.mark.parametrize("chunk,intra", [(3, 0), (4, 1), (64, 0)])
def test_run_loop_equals_stepwise(chunk, intra):
    sc = synth.make_scene(14, 320, 240, 6, seed=31, contrast=0.9)
    H, W = sc.shape
    ref = run_device(sc, intra_check=intra, intra_grow=intra)
    osc = run_oracle(sc, intra_check=intra, intra_grow=intra)
    out = {k: np.full((sc.n, H, W) + ((3,) if k == "points" else ()), np.nan, np.float32)
           for k in ("depth", "sigma", "checked", "points")}
    with api.Context(width=W, height=H, max_keyframes=sc.n, intra_check=intra, intra_grow=intra) as ctx:
        items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        up = ctx.upload_descs(sc, range(sc.n))
        d1 = _descs(out, range(sc.n), W, ["depth", "sigma"], [])
        d2 = _descs(out, range(sc.n), W, [], ["checked", "points"])
        for _ in range(2):  # the second run re-uploads over resident slots with downloads in flight
            ctx.run_loop(upload=up, pass1=items, down1=d1, pass2=items, down2=d2, chunk=chunk)
        ctx.synchronize()
    for k in out:
        assert np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)), k
    assert np.array_equal(out["checked"].view(np.uint32), osc.checked.view(np.uint32))


def test_run_loop_with_different_pass_lists_and_resident_planes():
    """pass-2 work orders are a subset in another order (the reference gates the two passes differently, :365-384 vs
    :523-542), nothing is uploaded (planes resident from an earlier call) and only some planes are requested"""
    sc = synth.make_scene(12, 200, 150, 6, seed=32, contrast=0.9)
    H, W = sc.shape
    ref = run_device(sc)
    sel = [9, 2, 5, 6, 0]
    out = {"checked": np.zeros((len(sel), H, W), np.float32)}
    with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
        ctx.upload_scene(sc)
        items1 = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        items2 = api.make_items(sel, sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        d2 = _descs(out, sel, W, [], ["checked"])
        ctx.run_loop(pass1=items1, pass2=items2, down2=d2, chunk=2)
        ctx.synchronize()
        for j, s in enumerate(sel):
            assert np.array_equal(out["checked"][j].view(np.uint32), ref["checked"][s].view(np.uint32)), s
        with pytest.raises(api.SdmError):
            bad = api.make_items([0], sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
            bad[0].nbr[0] = 99
            ctx.run_loop(pass1=bad)
