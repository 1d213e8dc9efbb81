"""sdm_edge_drawing on the device (stage 1: k_ed_planes) + host routing threads, through the C-ABI, against
 (1) the chains of the reference's closed-source EDLib.a (LineDetector::DetectEdgeMap, /root/reference/src/LineDetector.cc:843-881)
     committed as tests/golden/ed_chains_small.npz / ed_chains_misc.npz (oracle/make_ed_golden.py) - identity, chain for chain;
 (2) the host form of the same implementation (eao-slam_b200/host/edge_drawing.h via tests/cpp/test_edge_drawing.cpp, compiled
     on the box) - the stage-1 planes G / F bit for bit, the chains of larger and pitched batches.
Nothing here reads /root/reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eao-slam_b200", "python"))
GOLD = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu

from test_edge_drawing import parse_dump, same  # noqa: E402


@pytest.fixture(scope="module")
def ed_bin(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("edg") / "test_edge_drawing")
    subprocess.run(["g++", "-O2", "-std=c++11", "-o", out, os.path.join(ROOT, "tests", "cpp", "test_edge_drawing.cpp")], check=True)
    return out


@pytest.fixture(scope="module")
def scene():
    from sdmb200 import synth
    g = np.load(os.path.join(GOLD, "ed_chains_small.npz"))
    n_kf, W, H, n_nbr, seed = (int(v) for v in g["scene"])
    return synth.make_scene(n_kf, W, H, n_nbr, seed=seed, workers=4), g


def host_run(binary, images, tmp, planes=False):
    images = np.ascontiguousarray(images)
    n, H, W = images.shape
    raw, out, pl = (os.path.join(tmp, f) for f in ("in.raw", "out.bin", "planes.bin"))
    images.tofile(raw)
    subprocess.run([binary, str(W), str(H), str(n), raw, out, "-"] + ([pl] if planes else []), check=True)
    res = parse_dump(out, n)
    if not planes:
        return res
    b = np.fromfile(pl, np.uint8).reshape(n, 3 * H * W)
    G = b[:, :2 * H * W].copy().view(np.int16).reshape(n, H, W)
    F = b[:, 2 * H * W:].reshape(n, H, W)
    return res, G, F


def ctx_for(W, H):
    from sdmb200 import api
    return api.Context(width=W, height=H, max_keyframes=2)


def test_stage1_planes_match_the_host_form(scene, ed_bin, tmp_path):
    sc, _ = scene
    _, G, F = host_run(ed_bin, sc.im, str(tmp_path), planes=True)
    with ctx_for(sc.im.shape[2], sc.im.shape[1]) as ctx:
        for i in range(sc.im.shape[0]):
            g, f = ctx.ed_planes(sc.im[i])
            assert np.array_equal(g, G[i]) and np.array_equal(f, F[i]), i
    assert (F & 0x80).sum() > 10000 and set(np.unique(F & 3)) == {0, 1, 2}
    m = np.load(os.path.join(GOLD, "ed_chains_misc.npz"))
    for name in m["names"]:
        im = m["im_" + name]
        H, W = im.shape
        if W < 8 or H < 8:
            continue
        _, G, F = host_run(ed_bin, im[None], str(tmp_path), planes=True)
        with ctx_for(W, H) as ctx:
            g, f = ctx.ed_planes(im)
        assert np.array_equal(g, G[0]) and np.array_equal(f, F[0]), name


@pytest.mark.parametrize("threads", [1, 4, "device"])
def test_chains_and_mask_match_the_library_on_the_scene(scene, threads):
    """threads = "device": stage 2 in k_ed_route (one warp per image) instead of host threads"""
    sc, g = scene
    n, H, W = sc.im.shape
    with ctx_for(W, H) as ctx:
        if threads == "device":
            ctx.set_edge_drawing_route(True)
            threads = 0
        offs, pix, edge = ctx.edge_drawing(sc.im, n_threads=threads)
        t = ctx.last_edge_drawing_ms()
        assert ctx.last_edge_drawing_fallbacks() == 0
    assert t["kernel_ms"] > 0 and t["wall_ms"] > 0 and t["route_thread_ms"] > 0
    for i in range(n):
        assert same((offs[i], pix[i]), (g[f"off_{i}"], g[f"pix_{i}"])), i
        want = np.full((H, W), -1, np.int32)   # LineDetector.cc:857-866 on KeyFrame.cc:87's plane of -1
        ids = np.repeat(np.arange(len(offs[i]) - 1, dtype=np.int32), np.diff(offs[i]))
        want[pix[i] >> 16, pix[i] & 0xffff] = ids
        assert np.array_equal(edge[i], want)


@pytest.mark.parametrize("device_route", [False, True])
def test_chains_match_the_library_on_the_odd_images(device_route):
    m = np.load(os.path.join(GOLD, "ed_chains_misc.npz"))
    done = 0
    for name in m["names"]:
        im = m["im_" + name]
        H, W = im.shape
        if W < 8 or H < 8:   # below the smallest context (sdm_create)
            continue
        with ctx_for(W, H) as ctx:
            ctx.set_edge_drawing_route(device_route)
            offs, pix, edge = ctx.edge_drawing(im[None], edge_index=device_route)
        assert same((offs[0], pix[0]), (m["off_" + name], m["pix_" + name])), name
        if device_route:
            want = np.full((H, W), -1, np.int32)
            want[pix[0] >> 16, pix[0] & 0xffff] = np.repeat(np.arange(len(offs[0]) - 1, dtype=np.int32), np.diff(offs[0]))
            assert np.array_equal(edge[0], want), name
        done += 1
    assert done >= 12


@pytest.mark.parametrize("device_route", [False, True])
def test_large_pitched_batch_matches_the_host_form(scene, ed_bin, tmp_path, device_route):
    """more keyframes than one device chunk, rows with a pitch, more threads than chunks; pitched edge-index planes of
    every second image in the device mode"""
    sc, _ = scene
    n, H, W = sc.im.shape
    ims = np.concatenate([sc.im, sc.im[:, ::-1], sc.im[:, :, ::-1], 255 - sc.im, sc.im[:, ::-1, ::-1]])[:27]
    want = host_run(ed_bin, ims, str(tmp_path))
    pitched = np.zeros((len(ims), H, W + 24), np.uint8)
    pitched[:, :, :W] = ims
    with ctx_for(W, H) as ctx:
        ctx.set_edge_drawing_route(device_route)
        ctx.edge_drawing(ims[:3])  # (the buffers of a context grow with the batch)
        offs, pix, _ = ctx.edge_drawing([pitched[i, :, :W] for i in range(len(ims))], n_threads=16, edge_index=False)
        assert ctx.last_edge_drawing_fallbacks() == 0
        empty = ctx.edge_drawing(np.zeros((0, H, W), np.uint8))
    assert empty[0] == [] and empty[1] == []
    for i in range(len(ims)):
        assert same((offs[i], pix[i]), want[i]), i


def test_device_route_falls_back_to_the_host_when_a_capacity_runs_out(scene, monkeypatch):
    """SDM_ED_ROUTE_TEST_CAPS shrinks k_ed_route's per-tree arrays (120 walked pixels, 30 chains): most keyframes run out of
    room on the device, are routed on the host instead, and still give the library's chains and mask; the rest stay on the
    device path"""
    sc, g = scene
    n, H, W = sc.im.shape
    monkeypatch.setenv("SDM_ED_ROUTE_TEST_CAPS", "120")
    with ctx_for(W, H) as ctx:
        ctx.set_edge_drawing_route(True)
        offs, pix, edge = ctx.edge_drawing(sc.im)
        fb = ctx.last_edge_drawing_fallbacks()
    assert 0 < fb <= n
    for i in range(n):
        assert same((offs[i], pix[i]), (g[f"off_{i}"], g[f"pix_{i}"])), i
        w = np.full((H, W), -1, np.int32)
        w[pix[i] >> 16, pix[i] & 0xffff] = np.repeat(np.arange(len(offs[i]) - 1, dtype=np.int32), np.diff(offs[i]))
        assert np.array_equal(edge[i], w), i


@pytest.mark.parametrize("mode", [1, 2])
def test_mask_goes_from_the_detector_to_the_loop_on_the_device(mode):
    """SDM_ED_ROUTE_DEVICE (1: walks in k_ed_route) and SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE (2: walks on host threads, k_ed_mask
    scatters the chain numbers) keep kf->mEdgeIndex on the device; sdm_upload_keyframes takes the device planes as `edge`
    (sdm_ed_device_edge_plane).  Both passes over images + device masks give the planes of the same loop fed with the host
    planes of the default mode, bit for bit, and the chains of the two calls are the same."""
    from sdmb200 import api, synth
    n, W, H, N = 12, 320, 240, 6
    sc = synth.make_scene(n, W, H, N, seed=23)
    items = api.make_items(range(n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
    with api.Context(width=W, height=H, max_keyframes=n) as ctx:
        offs0, pix0, edge = ctx.edge_drawing(sc.im)  # default mode: host threads, host planes
        assert (edge >= 0).sum() > 10000
        with pytest.raises(api.SdmError):
            ctx.ed_device_edge_plane(0)  # no masks on the device after a default-mode call
        ctx.set_edge_drawing_route(mode)
        offs, pix, none = ctx.edge_drawing(sc.im, edge_index=False)
        assert none is None and ctx.last_edge_drawing_fallbacks() == 0
        for i in range(n):
            assert same((offs[i], pix[i]), (offs0[i], pix0[i])), i
        dev = [ctx.ed_device_edge_plane(i) for i in range(n)]
        with pytest.raises(api.SdmError):
            ctx.ed_device_edge_plane(n)
        ctx.upload_keyframes(ctx.upload_descs(sc, range(n), images_only=True, edge_dev=dev))
        ctx.pass1(items); ctx.pass2(items)

        def planes():
            per = [ctx.download(i) for i in range(n)]
            return {k: np.stack([q[k] for q in per]) for k in ("depth", "sigma", "checked", "points")}
        got = planes()
        cands = [ctx.candidate_count(i) for i in range(n)]
        sce = synth.Scene(im=sc.im, grad=sc.grad, theta=sc.theta, edge=edge, K=sc.K, Tcw=sc.Tcw, nbr_idx=sc.nbr_idx, rot=sc.rot,
                          min_depth=sc.min_depth, max_depth=sc.max_depth)
        ctx.upload_keyframes(ctx.upload_descs(sce, range(n), images_only=True))
        ctx.pass1(items); ctx.pass2(items)
        want = planes()
        assert cands == [ctx.candidate_count(i) for i in range(n)] and 0 < sum(cands) < (edge >= 0).sum() + 1
    for k in ("depth", "sigma", "checked", "points"):
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k
    assert (want["checked"] > 0).sum() > 1000 and not (want["checked"] > 0)[edge < 0].any()


def test_argument_errors():
    from sdmb200 import api
    with ctx_for(64, 48) as ctx:
        im = np.zeros((48, 64), np.uint8)
        with pytest.raises(api.SdmError):
            ctx.edge_drawing(im[None], grad_thresh=0)
        with pytest.raises(api.SdmError):
            ctx.edge_drawing(im[None], grad_thresh=4000)
        offs, pix, edge = ctx.edge_drawing(im[None])
        assert offs[0].tolist() == [0] and pix[0].size == 0 and (edge == -1).all()
