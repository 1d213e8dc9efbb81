"""The C++ drop-in class (eao-slam_b200/host/ProbabilityMapping.h): builds with the reference's dialect
(-std=c++11) against the C-ABI library, and — on a GPU — its SemiDenseLoop / per-method calls reproduce
the oracle on the same keyframes (pitched cv::Mat-like planes, reference gating, reference argument order)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_py as O
from helpers import compare_planes
from sdmb200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "eao-slam_b200", "lib")


@pytest.fixture(scope="module")
def shim_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "test_shim")
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"), "-o", out, "-L", LIBDIR, "-lsdm_b200",
                    f"-Wl,-rpath,{LIBDIR}", "-lpthread"], check=True)
    return out


def test_shim_builds_as_cxx11(shim_binary):
    assert os.path.exists(shim_binary)


def _write_scene(path, sc, N, pad, probes, first_id=1, extra_ids=11, bad=None):
    n, (H, W) = sc.n, sc.shape
    with open(path, "wb") as f:
        f.write(struct.pack("9i", n, W, H, N, pad, len(probes), sc.nbr_idx.shape[1], first_id, extra_ids))
        f.write(np.asarray(sc.K, np.float32).tobytes())
        for i in range(n):
            f.write(np.ascontiguousarray(sc.Tcw[i], np.float32).tobytes())
            f.write(sc.im[i].tobytes()); f.write(sc.grad[i].tobytes()); f.write(sc.theta[i].tobytes())
            f.write(struct.pack("i", len(sc.inv_depths[i]))); f.write(sc.inv_depths[i].tobytes())
            f.write(np.ascontiguousarray(sc.nbr_idx[i], np.int32).tobytes())
            f.write(struct.pack("i", int(bad[i]) if bad is not None else 0))
        f.write(np.asarray(probes, np.int32).reshape(-1).tobytes())


@pytest.mark.gpu
def test_shim_gating_matches_the_reference_source(shim_binary, tmp_path):
    """The class shim against the REFERENCE'S OWN SemiDenseLoop() (oracle/_ref, tests/test_ref_vs_oracle.py) in a
    scenario where the gating matters: the first keyframe gets mapping id 0 and is never "Mapped" (KeyFrame.cc:796-806),
    only keyframes with more than 10 newer mapped ones are processed (:789-794), one keyframe is bad, and every
    covisibility list is longer than covisN so that the first seven GOOD entries are taken (:365-384, :523-542).
    Same flags, same planes."""
    import ref_py
    if not ref_py.available():
        pytest.skip("needs /root/reference or a prebuilt oracle/_ref/libref_pm.so")
    n, W, H, N, n_cov = 24, 96, 72, 7, 11
    sc = synth.make_scene(n, W, H, n_cov, seed=61, contrast=0.9)
    bad = np.zeros(n, np.int32); bad[5] = 1
    first_id, extra_ids = 0, 3          # ids 0..23, then 3 more: MappingIdDelay holds for ids 1..15
    ref = ref_py.run_reference_loop(sc, np.full(n, 50.0, np.float32), first_id=first_id, extra_ids=extra_ids, bad=bad)
    scene_path, out_path = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    _write_scene(scene_path, sc, N, 0, [], first_id=first_id, extra_ids=extra_ids, bad=bad)
    r = subprocess.run([shim_binary, scene_path, out_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out_path, np.float32)
    per = 2 + 6 * W * H
    flags = np.zeros((n, 2), np.int32)
    for i in range(n):
        blk = raw[i * per:(i + 1) * per]
        flags[i] = blk[:2].view(np.int32)
        p = blk[2:]
        assert np.array_equal(p[:W * H].reshape(H, W).view(np.uint32), ref["depth"][i].view(np.uint32)), i
        assert np.array_equal(p[W * H:2 * W * H].reshape(H, W).view(np.uint32), ref["sigma"][i].view(np.uint32)), i
        assert np.array_equal(p[2 * W * H:3 * W * H].reshape(H, W).view(np.uint32), ref["checked"][i].view(np.uint32)), i
        assert np.array_equal(p[3 * W * H:].reshape(H, W, 3).view(np.uint32), ref["points"][i].view(np.uint32)), i
    assert np.array_equal(flags, ref["flags"])
    done = ref["flags"][:, 0] == 1
    assert 5 <= done.sum() < n and not done[0] and not done[5] and not done[-1], ref["flags"].T


@pytest.mark.gpu
def test_shim_semidense_loop_matches_oracle(shim_binary, tmp_path):
    n, W, H, N, pad = 12, 320, 240, 6, 16
    sc = synth.make_scene(n, W, H, N, seed=17)
    lib = O.lib()
    for i in range(n):  # the shim derives min/max depth from GetAllPointDepths() like :734-747
        a, b = C.c_float(), C.c_float()
        lib.oracle_stereo_search_constraints(O.fptr(sc.inv_depths[i]), len(sc.inv_depths[i]), C.byref(a), C.byref(b))
        sc.min_depth[i], sc.max_depth[i] = a.value, b.value
    osc = O.OracleScene(sc)
    osc.run()
    ys, xs = np.nonzero(sc.grad[5] > 8)
    pick = np.random.default_rng(0).choice(len(ys), 24, replace=False)
    probes = [(5, int(sc.nbr_idx[5][k % N]), int(xs[p]), int(ys[p])) for k, p in enumerate(pick)]
    scene_path, out_path = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    _write_scene(scene_path, sc, N, pad, probes)
    r = subprocess.run([shim_binary, scene_path, out_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out_path, np.float32)
    per = 2 + 6 * W * H
    dev = {k: np.zeros((n, H, W) + ((3,) if k == "points" else ()), np.float32) for k in ("depth", "sigma", "checked", "points")}
    for i in range(n):
        blk = raw[i * per:(i + 1) * per]
        assert tuple(blk[:2].view(np.int32)) == (1, 1)  # semidense_flag_, interKF_depth_flag_ set (:497, :554)
        p = blk[2:]
        dev["depth"][i] = p[:W * H].reshape(H, W); dev["sigma"][i] = p[W * H:2 * W * H].reshape(H, W)
        dev["checked"][i] = p[2 * W * H:3 * W * H].reshape(H, W); dev["points"][i] = p[3 * W * H:].reshape(H, W, 3)
    rep = compare_planes(dev, osc)
    print(rep)
    # SetProducePlanesOnDevice(true): im_ only goes up, the device makes GradImg / GradTheta -> the same planes
    out2 = str(tmp_path / "out2.bin")
    r2 = subprocess.run([shim_binary, scene_path, out2], capture_output=True, text=True,
                        env=dict(os.environ, SDM_SHIM_DEVICE_PLANES="1"))
    assert r2.returncode == 0, r2.stdout + r2.stderr
    raw2 = np.fromfile(out2, np.float32)
    assert np.array_equal(raw2[:n * per].view(np.uint32), raw[:n * per].view(np.uint32))
    tail = raw[n * per:]
    rec = tail[:8 * len(probes)].reshape(-1, 8)
    for (k1, k2, x, y), got in zip(probes, rec):
        pr = osc.pair(k1, k2)
        a, b = C.c_float(), C.c_float()
        lib.oracle_get_search_range(C.byref(osc.kfs[k1]), C.byref(pr), x, y, float(sc.min_depth[k1]), float(sc.max_depth[k1]), C.byref(a), C.byref(b))
        assert (got[0], got[1]) == (sc.min_depth[k1], sc.max_depth[k1])
        assert (got[2], got[3]) == (np.float32(a.value), np.float32(b.value))
        d, s, u, ok = osc.pass1_pair(k1, k2, 0.0)
        assert bool(got[6]) == bool(ok[y, x])
        if ok[y, x]:
            assert got[4] == d[y, x] and got[5] == s[y, x] and got[7] == u[y, x]
    fus = tail[8 * len(probes):8 * len(probes) + 3]
    hd = np.array([0.50, 0.51, 0.49, 0.505, 0.9, 0.495], np.float32); hs = np.array([0.02, 0.02, 0.03, 0.01, 0.02, 0.02], np.float32)
    a, b = C.c_float(), C.c_float()
    p = O.default_params()
    okf = lib.oracle_fusion(O.fptr(hd), O.fptr(hs), 6, C.byref(p), C.byref(a), C.byref(b))
    assert okf == int(fus[2]) == 1 and fus[0] == np.float32(a.value) and fus[1] == np.float32(b.value)
    intra = tail[8 * len(probes) + 3:]
    d2, s2 = osc.depth[2].copy(), osc.sigma[2].copy()
    lib.oracle_intra_check(O.fptr(d2), O.fptr(s2), W, H, C.byref(p))
    assert np.array_equal(intra[:W * H].reshape(H, W), d2) and np.array_equal(intra[W * H:2 * W * H].reshape(H, W), s2)
    # ExportSemiDensePoints == the filter loop of SaveSemiDensePoints (:159-186) over the oracle's planes
    ex = intra[2 * W * H:]
    n_pts, n_kf = int(ex[0]), int(ex[1])
    counts = ex[2:2 + n_kf].astype(np.int64)
    rec = ex[2 + n_kf:2 + n_kf + 4 * n_pts].reshape(n_pts, 4)
    exp_xyz, exp_pix, exp_counts = [], [], []
    for i in range(n):
        keep = ~(osc.sigma[i].astype(np.float64) > 0.02) & (osc.checked[i].astype(np.float64) > 0.000001)
        ys, xs = np.nonzero(keep)
        exp_xyz.append(osc.points[i][ys, xs]); exp_pix.append((ys.astype(np.uint32) << 16) | xs.astype(np.uint32))
        exp_counts.append(len(ys))
    assert n_kf == n and list(counts) == exp_counts and n_pts == sum(exp_counts) > 1000
    assert np.array_equal(rec[:, :3].view(np.uint32), np.concatenate(exp_xyz).view(np.uint32))
    assert np.array_equal(rec[:, 3].copy().view(np.uint32), np.concatenate(exp_pix))
