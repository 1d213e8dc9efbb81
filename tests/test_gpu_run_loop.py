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


def _pinned(lib, shape, keep):
    n = int(np.prod(shape)) * 4
    p = C.c_void_p()
    assert lib.sdm_host_alloc(C.byref(p), n) == 0
    keep.append(p)
    a = np.frombuffer((C.c_char * n).from_address(p.value), dtype=np.float32).reshape(shape)
    a[:] = 0
    return a


@pytest.mark.parametrize("W,H", [(320, 240), (203, 77)])
def test_block_sparse_download_into_pinned_planes(W, H):
    """sdm_loop.sparse_download / sdm_scatter_keyframes with PINNED zero-initialised destination planes: a kernel writes
    only the 16-pixel row blocks that hold a candidate over PCIe.  Same host planes as the dense DMA, bit for bit (odd
    width: partial last block, rows whose mask word is empty); row-pitched destinations; fewer bytes than the dense copy;
    pageable planes with the same flag fall back to the dense DMA."""
    sc = synth.make_scene(10, W, H, 6, seed=33, contrast=0.9)
    ref = run_device(sc)
    keep = []
    lib = api.load()
    pad = 8
    out = {k: _pinned(lib, (sc.n, H, (W + pad) * (3 if k == "points" else 1)), keep) for k in ("depth", "sigma", "checked", "points")}
    view = {k: (out[k][:, :, :3 * W].reshape(sc.n, H, W, 3) if k == "points" else out[k][:, :, :W]) for k in out}
    try:
        with api.Context(width=W, height=H, max_keyframes=sc.n) as ctx:
            items = api.make_items(range(sc.n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
            up = ctx.upload_descs(sc, range(sc.n))
            d1 = (api.DownloadDesc * sc.n)(); d2 = (api.DownloadDesc * sc.n)(); d4 = (api.DownloadDesc * sc.n)()
            for j in range(sc.n):
                d1[j].kf = d2[j].kf = d4[j].kf = j
                for d, names in ((d1[j], ("depth", "sigma")), (d2[j], ("checked", "points")), (d4[j], ("depth", "sigma", "checked", "points"))):
                    for nm in names:
                        setattr(d, nm, out[nm][j].ctypes.data)
                        setattr(d, nm + "_step", out[nm][j].strides[0])
            ctx.run_loop(upload=up, pass1=items, down1=d1, pass2=items, down2=d2, chunk=3, sparse=True)
            ctx.synchronize()
            for k in out:
                assert np.array_equal(view[k].view(np.uint32), ref[k].view(np.uint32)), k
                assert not out[k][:, :, (3 * W if k == "points" else W):].any(), "padding columns untouched"
            nblk = ctx.candidate_blocks(range(sc.n))
            assert 0 < nblk <= sc.n * H * ((W + 15) // 16)
            # sparse_download = 2: the 4-byte planes by DMA, SemiDensePointSets_ block-sparse on a second stream
            for k in out:
                out[k][:] = 0
            ctx.run_loop(upload=up, pass1=items, down1=d1, pass2=items, down2=d2, chunk=4, sparse=2)
            ctx.synchronize()
            for k in out:
                assert np.array_equal(view[k].view(np.uint32), ref[k].view(np.uint32)), ("mode 2", k)
                assert not out[k][:, :, (3 * W if k == "points" else W):].any(), "padding columns untouched"
            for k in out:
                out[k][:] = 0
            ctx.scatter_keyframes(d4)            # the entry point itself: pinned planes take the kernel route
            ctx.synchronize()
            for k in out:
                assert np.array_equal(view[k].view(np.uint32), ref[k].view(np.uint32)), k
            # pageable destination + sparse flag: dense DMA
            pg = {k: np.full((sc.n, H, W) + ((3,) if k == "points" else ()), np.nan, np.float32) for k in ("checked",)}
            d3 = (api.DownloadDesc * sc.n)()
            for j in range(sc.n):
                d3[j].kf = j
                d3[j].checked, d3[j].checked_step = pg["checked"][j].ctypes.data, pg["checked"][j].strides[0]
            ctx.run_loop(pass1=items, pass2=items, down2=d3, sparse=True)
            ctx.synchronize()
            assert np.array_equal(pg["checked"].view(np.uint32), ref["checked"].view(np.uint32))
    finally:
        for p in keep:
            lib.sdm_host_free(p)
