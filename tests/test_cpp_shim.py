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


def test_shim_builds_in_orbslam2_mode(tmp_path):
    """-DSDM_HOST_WITH_ORBSLAM2: the header compiles and links against ORB_SLAM2::KeyFrame / Map / cv::Mat, Modeler and
    LineDetector declarations - here the stand-ins of oracle/refshim/, which carry the reference's member names (the
    same headers the reference's own ProbabilityMapping.cc is compiled against for oracle/_ref)."""
    src = tmp_path / "orb.cpp"
    src.write_text('#define SDM_HOST_WITH_ORBSLAM2\n'
                   f'#include "{ROOT}/eao-slam_b200/host/ProbabilityMapping.h"\n'
                   'long unsigned int ORB_SLAM2::KeyFrame::nNextMappingId = 1;\n'
                   'int main() { ORB_SLAM2::Map m; ProbabilityMapping pm(&m); pm.RequestFinish(); pm.Run();\n'
                   '             pm.GetModeler(); pm.LineFittingOnDevice(); return pm.isFinished() ? 0 : 1; }\n')
    out = str(tmp_path / "orb")
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "oracle", "refshim"), "-I",
                    os.path.join(ROOT, "include"), str(src), "-o", out, "-L", LIBDIR, "-lsdm_b200", f"-Wl,-rpath,{LIBDIR}",
                    "-lpthread"], check=True)
    assert os.path.exists(out)


def _read_planes(raw, n, W, H):
    per = 2 + 6 * W * H
    dev = {k: np.zeros((n, H, W) + ((3,) if k == "points" else ()), np.float32) for k in ("depth", "sigma", "checked", "points")}
    flags = np.zeros((n, 2), np.int32)
    for i in range(n):
        blk = raw[i * per:(i + 1) * per]
        flags[i] = blk[:2].view(np.int32)
        p = blk[2:]
        dev["depth"][i] = p[:W * H].reshape(H, W); dev["sigma"][i] = p[W * H:2 * W * H].reshape(H, W)
        dev["checked"][i] = p[2 * W * H:3 * W * H].reshape(H, W); dev["points"][i] = p[3 * W * H:].reshape(H, W, 3)
    return dev, flags, raw[n * per:]


@pytest.mark.gpu
@pytest.mark.parametrize("route_on_device", [False, True])
def test_shim_with_open_edge_drawing(shim_binary, tmp_path, route_on_device):
    """SetEdgeDrawing(true): the class detects the edge maps of the pass-1 keyframes itself, where the reference calls
    LineDetector::DetectEdgeMap (ProbabilityMapping.cc:394, LineDetector.cc:843-881) - mEdgeIndex and the chains equal the
    C-ABI's (pinned to the reference's EDLib.a in tests/test_gpu_edge_drawing.py), the loop's planes equal the oracle's
    on that mask (the candidate filter of :454), and FitLines over the kept chains equals sdm_line_fit."""
    from helpers import run_oracle
    from sdmb200 import api
    n, W, H, N = 12, 320, 240, 6
    sc = synth.make_scene(n, W, H, N, seed=17)
    lib = O.lib()
    for i in range(n):
        a, b = C.c_float(), C.c_float()
        lib.oracle_stereo_search_constraints(O.fptr(sc.inv_depths[i]), len(sc.inv_depths[i]), C.byref(a), C.byref(b))
        sc.min_depth[i], sc.max_depth[i] = a.value, b.value
    scene_path, out_path, edge_path = (str(tmp_path / f) for f in ("scene.bin", "out.bin", "edge.bin"))
    _write_scene(scene_path, sc, N, 16, [])
    r = subprocess.run([shim_binary, scene_path, out_path], capture_output=True, text=True,
                       env=dict(os.environ, SDM_SHIM_EDGE_OUT=edge_path, **({"SDM_SHIM_EDGE_DEVICE": "1"} if route_on_device else {})))
    assert r.returncode == 0, r.stdout + r.stderr
    with api.Context(width=W, height=H, max_keyframes=n) as ctx:
        offs, pix, edge = ctx.edge_drawing(sc.im)
        e = np.fromfile(edge_path, np.int32)
        p = 0
        for i in range(n):
            nc, npx = int(e[p]), int(e[p + 1]); p += 2
            assert nc == len(offs[i]) - 1 > 100
            assert np.array_equal(e[p:p + nc + 1], offs[i]); p += nc + 1
            assert np.array_equal(e[p:p + npx].view(np.uint32), pix[i]); p += npx
            assert np.array_equal(e[p:p + W * H].reshape(H, W), edge[i]); p += W * H
        n_lines = int(e[p]); p += 1
        shim_lines = e[p:].view(api.LINE3D)
        assert len(shim_lines) == n_lines > 50
        sce = synth.Scene(im=sc.im, grad=sc.grad, theta=sc.theta, edge=edge, K=sc.K, Tcw=sc.Tcw, nbr_idx=sc.nbr_idx, rot=sc.rot,
                          min_depth=sc.min_depth, max_depth=sc.max_depth)
        osc = run_oracle(sce)
        raw = np.fromfile(out_path, np.float32)
        dev, flags, _ = _read_planes(raw, n, W, H)
        assert (flags == 1).all()
        rep = compare_planes(dev, osc)
        assert all(rep[k + "_bit_mismatch"] == 0 for k in ("depth", "sigma", "checked", "points")), rep
        assert rep["pass2_accepted_ref"] > 1000 and not (dev["checked"] > 0)[edge < 0].any()
        ctx.upload_scene(sce)
        items = api.make_items(range(n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        ctx.pass1(items); ctx.pass2(items)
        lines, _ = ctx.line_fit(list(range(n)), offs, pix)
        assert lines.tobytes() == shim_lines.tobytes()


@pytest.mark.gpu
def test_shim_online_mode_matches_the_reference_online_build(shim_binary, tmp_path):
    """SURVEY 8f-4.  The reference compiled with -DOnlineLoop (oracle/_ref/libref_pm_online.so) and the shim are driven
    through the same sequence: keyframes arrive in three batches, after each one the body of Run()'s online loop
    (:224-226: SemiDenseLoop(); UpdateAllSemiDensePointSet();) runs; before the second round local BA "moves" six
    keyframes (SetPose), some already finished (their point sets must follow, :691-694), some still waiting (their
    later passes must use the new pose - the reference reads GetRotation()/GetTranslation() afresh in every loop);
    then the final loop of :244.  The map outgrows the shim's first arena on the way, so finished keyframes are
    re-seeded from their cv::Mat planes.  Same flags, same planes, bit for bit."""
    import ref_py
    if not ref_py.available():
        pytest.skip("needs /root/reference or a prebuilt oracle/_ref")
    n, W, H, N, n_cov = 44, 96, 72, 7, 10
    n1, n2, extra = 24, 34, 11
    sc = synth.make_scene(n, W, H, n_cov, seed=71, contrast=0.9)
    rng = np.random.default_rng(3)
    moved = np.array([3, 7, 11, 20, 26, 30], np.int32)
    Tm = sc.Tcw[moved].copy()
    Tm[:, :, 3] += rng.normal(0, 0.004, (len(moved), 3)).astype(np.float32)   # a few millimetres, like a BA update
    ref = ref_py.run_reference_online(sc, n1, n2, extra, moved, Tm)
    scene_path, seq_path, out_path = (str(tmp_path / f) for f in ("scene.bin", "seq.bin", "out.bin"))
    _write_scene(scene_path, sc, N, 0, [], first_id=1, extra_ids=0)
    with open(seq_path, "wb") as f:
        f.write(struct.pack("4i", n1, n2, extra, len(moved)))
        f.write(moved.tobytes()); f.write(np.ascontiguousarray(Tm.reshape(len(moved), 12), np.float32).tobytes())
    r = subprocess.run([shim_binary, "--online", scene_path, seq_path, out_path], capture_output=True, text=True,
                       env=dict(os.environ, SDM_SHIM_HEADROOM="3"))  # 24 + 1 + 3 slots: outgrown by the second batch
    assert r.returncode == 0, r.stdout + r.stderr
    dev, flags, tail = _read_planes(np.fromfile(out_path, np.float32), n, W, H)
    assert np.array_equal(flags, ref["flags"]), (flags.T, ref["flags"].T)
    done1, done2 = ref["flags"][:, 0] == 1, ref["flags"][:, 1] == 1
    assert done2.sum() >= 25 and done2[moved[:3]].all(), ref["flags"].T
    for k in ("depth", "sigma", "checked", "points"):
        assert np.array_equal(dev[k].view(np.uint32), ref[k].view(np.uint32)), k
    n_pts, regrown, hooks = int(tail[0]), int(tail[1]), int(tail[2])
    assert regrown == 1, "the scenario is sized so that the map outgrows the first arena"
    assert hooks == int(done1.sum()), "the edge-map hook runs once per keyframe entering pass 1 (:394-397)"
    keep = ~(ref["sigma"].astype(np.float64) > 0.02) & (ref["checked"].astype(np.float64) > 0.000001) & done2[:, None, None]
    assert n_pts == int(keep.sum()) > 1000
    pts = tail[3:3 + 4 * n_pts].reshape(n_pts, 4)
    assert np.array_equal(pts[:, :3].view(np.uint32), ref["points"][keep].view(np.uint32))


@pytest.mark.gpu
def test_shim_save_semidense_points_equals_the_reference_writer(shim_binary, tmp_path, monkeypatch):
    """SaveSemiDensePoints (:136-192): the shim writes the OBJ from the device-compacted point stream; the reference's own
    function (oracle/_ref) writes it from the same planes with its per-pixel scan.  Byte-identical files."""
    import ref_py
    if not ref_py.available():
        pytest.skip("needs /root/reference or a prebuilt oracle/_ref")
    n, W, H, N = 14, 160, 120, 6
    sc = synth.make_scene(n, W, H, 9, seed=18, contrast=0.9)   # covisibility lists of 9: the bad keyframe is skipped over
    scene_path, out_path = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    bad = np.zeros(n, np.int32); bad[4] = 1
    _write_scene(scene_path, sc, N, 0, [], bad=bad)
    res = str(tmp_path / "shim_results")
    r = subprocess.run([shim_binary, scene_path, out_path], capture_output=True, text=True, cwd=str(tmp_path),
                       env=dict(os.environ, SDM_SHIM_RESULTS=res))
    assert r.returncode == 0, r.stdout + r.stderr
    dev, flags, _ = _read_planes(np.fromfile(out_path, np.float32), n, W, H)
    monkeypatch.chdir(tmp_path)
    ref_file = ref_py.reference_save_points(sc.im, dev["sigma"], dev["checked"], dev["points"], flags, bad=bad)
    a, b = open(os.path.join(res, "semi_pointcloud.obj"), "rb").read(), open(ref_file, "rb").read()
    assert flags[:, 1].sum() >= 8 and flags[4].sum() == 0
    assert len(b) > 100000 and a == b


@pytest.mark.gpu
def test_shim_timing_mode_runs_the_pipelined_loop(shim_binary, tmp_path):
    n, W, H, N = 16, 320, 240, 6
    sc = synth.make_scene(n, W, H, N, seed=19)
    scene_path, out_path = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    _write_scene(scene_path, sc, N, 0, [])
    outs = []
    for chunk in ("0", "5"):
        r = subprocess.run([shim_binary, "--time", scene_path, out_path, "2"], capture_output=True, text=True,
                           env=dict(os.environ, SDM_SHIM_CHUNK=chunk))
        assert r.returncode == 0, r.stdout + r.stderr
        import json
        t = json.loads(r.stdout.strip().splitlines()[-1])
        assert t["finished"] == n and t["semidense_loop_ms_best"] > 0
        outs.append(np.fromfile(out_path, np.float32))
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
