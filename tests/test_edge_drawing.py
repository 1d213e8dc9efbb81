"""Open Edge Drawing implementation (eao-slam_b200/host/edge_drawing.h) and its restatement (oracle/ed_oracle.py) against the
chains of the reference's closed-source EDLib.a (LineDetector::DetectEdgeMap, /root/reference/src/LineDetector.cc:843-881;
the call is DetectEdgesByED(srcImg, width, height, SOBEL_OPERATOR, 36, 8, 1.0), :855).

Fixtures (oracle/make_ed_golden.py, which runs the library's binary where it lies):
  tests/golden/ed_chains_small.npz - six 320 x 240 keyframes of the synthetic scene (2742 chains)
  tests/golden/ed_chains_misc.npz  - noise, flat, shapes, quantised blobs, widths that are not multiples of four, 5 x 7, 6 x 300
The bar is identity: the same chains, pixel for pixel, in the same order.  Where /root/reference exists (the build container)
two more tests run the library itself on fresh random images and on its own lena.pgm, and the bundled OpenCV 2.4.5 cvSmooth
on random images against the smoothing stage.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eao-slam_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden")
ED_REF = "/root/reference/Thirdparty/EDTest"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ed_chains")
have_ref = os.path.exists(os.path.join(ED_REF, "EDLib.a")) and os.path.exists(REF_BIN)


@pytest.fixture(scope="module")
def ed_bin(tmp_path_factory):
    """tests/cpp/test_edge_drawing.cpp: the header behind a raw-image -> chain-dump main()"""
    out = str(tmp_path_factory.mktemp("ed") / "test_edge_drawing")
    subprocess.run(["g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-o", out,
                    os.path.join(ROOT, "tests", "cpp", "test_edge_drawing.cpp")], check=True)
    return out


def parse_dump(path, n):
    a = np.fromfile(path, np.int32)
    assert a[0] == n
    p, res = 1, []
    for _ in range(n):
        ns = int(a[p]); p += 1
        off, pix = [0], []
        for _ in range(ns):
            m = int(a[p]); p += 1
            rc = a[p:p + 2 * m].reshape(m, 2); p += 2 * m
            pix.append((rc[:, 0].astype(np.uint32) << 16) | rc[:, 1].astype(np.uint32))
            off.append(off[-1] + m)
        res.append((np.array(off, np.int32), np.concatenate(pix) if pix else np.zeros(0, np.uint32)))
    assert p == a.size
    return res


def run_dump(binary, images, tmp, env=None, edge=False):
    images = np.ascontiguousarray(images)
    n, H, W = images.shape
    raw, out, eo = os.path.join(tmp, "in.raw"), os.path.join(tmp, "out.bin"), os.path.join(tmp, "edge.bin")
    images.tofile(raw)
    subprocess.run([binary, str(W), str(H), str(n), raw, out] + ([eo] if edge else []), check=True, env=env)
    res = parse_dump(out, n)
    return (res, np.fromfile(eo, np.int32).reshape(n, H, W)) if edge else res


def oracle_chains(im):
    import ed_oracle
    ch = ed_oracle.detect(im)
    off = np.concatenate([[0], np.cumsum([len(c) for c in ch])]).astype(np.int32)
    rc = np.array([p for c in ch for p in c], np.int64).reshape(-1, 2)
    return off, ((rc[:, 0].astype(np.uint32) << 16) | rc[:, 1].astype(np.uint32)).astype(np.uint32)


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_implementation_reproduces_the_library_on_the_scene_keyframes(ed_bin, tmp_path):
    from sdmb200 import synth
    g = np.load(os.path.join(GOLD, "ed_chains_small.npz"))
    n_kf, W, H, n_nbr, seed = (int(v) for v in g["scene"])
    sc = synth.make_scene(n_kf, W, H, n_nbr, seed=seed, workers=4)
    res, edge = run_dump(ed_bin, sc.im, str(tmp_path), edge=True)
    total = 0
    for i in range(n_kf):
        assert same(res[i], (g[f"off_{i}"], g[f"pix_{i}"])), f"keyframe {i}"
        total += len(res[i][0]) - 1
        # mEdgeIndex as LineDetector.cc:857-866 leaves it: -1, then chain numbers written in chain order (later chains win)
        want = np.full((H, W), -1, np.int32)
        off, pix = g[f"off_{i}"], g[f"pix_{i}"]
        ids = np.repeat(np.arange(len(off) - 1, dtype=np.int32), np.diff(off))
        want[pix >> 16, pix & 0xffff] = ids   # numpy assigns in order: the last write of a repeated pixel stays
        assert np.array_equal(edge[i], want)
    assert total > 2500


def test_implementation_reproduces_the_library_on_the_odd_images(ed_bin, tmp_path):
    g = np.load(os.path.join(GOLD, "ed_chains_misc.npz"))
    assert len(g["names"]) >= 15
    for name in g["names"]:
        res = run_dump(ed_bin, g["im_" + name][None], str(tmp_path))
        assert same(res[0], (g["off_" + name], g["pix_" + name])), name


def test_restatement_reproduces_the_library():
    g = np.load(os.path.join(GOLD, "ed_chains_misc.npz"))
    for name in ("black", "checker", "tiny", "thin", "shapes0", "shapes3", "blobs_w70", "blobs_w179", "blobs_w257"):
        assert same(oracle_chains(g["im_" + name]), (g["off_" + name], g["pix_" + name])), name
    from sdmb200 import synth
    s = np.load(os.path.join(GOLD, "ed_chains_small.npz"))
    n_kf, W, H, n_nbr, seed = (int(v) for v in s["scene"])
    sc = synth.make_scene(n_kf, W, H, n_nbr, seed=seed, workers=4)
    assert same(oracle_chains(sc.im[0]), (s["off_0"], s["pix_0"]))


def test_the_rounding_tail_and_the_joint_rule_are_exercised():
    """the two details found last (half-up rounding in the last W % 4 columns of the smoothing; the first chain of a main
    segment compared with the previous segment's last pixel) must matter on the fixture, or it would not pin them"""
    import ed_oracle
    g = np.load(os.path.join(GOLD, "ed_chains_misc.npz"))
    hits = 0
    for name in ("blobs_w70", "blobs_w179", "blobs_w598"):
        im = g["im_" + name]
        H, W = im.shape
        a = np.pad(im.astype(np.int64), 2, mode="edge")
        k = [1, 4, 6, 4, 1]
        row = sum(k[i] * a[:, i:i + W] for i in range(5))
        s = sum(k[i] * row[i:i + H, :] for i in range(5))
        hits += int(np.count_nonzero(np.rint(s / 256.0).astype(np.int64) != ed_oracle.smooth(im)))
    assert hits > 0


def _lib_env(tmp):
    for lib in ("libopencv_core.so.2.4", "libopencv_imgproc.so.2.4"):
        dst = os.path.join(tmp, lib)
        if not os.path.exists(dst):
            os.symlink(os.path.join(ED_REF, lib + ".5"), dst)
    return dict(os.environ, LD_LIBRARY_PATH=tmp)


def _read_pgm(path):
    with open(path, "rb") as f:
        d = f.read()
    toks, i = [], 0
    while len(toks) < 4:
        while d[i:i + 1].isspace():
            i += 1
        if d[i:i + 1] == b"#":
            i = d.index(b"\n", i)
            continue
        j = i
        while not d[j:j + 1].isspace():
            j += 1
        toks.append(d[i:j]); i = j
    W, H = int(toks[1]), int(toks[2])
    return np.frombuffer(d[i + 1:i + 1 + W * H], np.uint8).reshape(H, W)


@pytest.mark.skipif(not have_ref, reason="needs the reference's EDLib.a (build container only)")
def test_live_against_the_library(ed_bin, tmp_path):
    tmp = str(tmp_path)
    env = _lib_env(tmp)
    lena = _read_pgm(os.path.join(ED_REF, "lena.pgm"))   # the library's own demo image
    cases = [lena, lena[::-1], lena[:, ::-1], lena.T, lena[37:401, 11:500], lena[1::2, ::3]]
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    from make_ed_golden import blobs, shapes
    for _ in range(40):
        H, W = int(rng.integers(12, 400)), int(rng.integers(12, 500))
        k = int(rng.integers(0, 3))
        cases.append(blobs(rng, H, W, int(rng.integers(1, 6)), int(rng.integers(2, 9)) if k else 0) if k < 2
                     else shapes(rng, H, W, int(rng.integers(2, 15))))
    n_chains = 0
    for im in cases:
        im = np.ascontiguousarray(im)
        ref = run_dump(REF_BIN, im[None], tmp, env=env)[0]
        got = run_dump(ed_bin, im[None], tmp)[0]
        if not same(ref, got):
            np.save(os.path.join(ROOT, "gpurun_out", "ed_live_mismatch.npy"), im)
        assert same(ref, got), f"image {im.shape} (saved to gpurun_out/ed_live_mismatch.npy)"
        n_chains += len(ref[0]) - 1
    assert n_chains > 3000


def test_device_kernel_replayed_on_the_cpu_matches_the_host_planes(ed_bin, tmp_path):
    """k_ed_planes4 (csrc/edge_drawing_kernels.cuh: four pixels per thread, two 16-bit lanes per register) is written as
    host-callable phases; the driver replays them thread by thread for every tile and compares G / F with EdPlanesHost (exit
    code 4 on the first differing pixel).  Scene keyframes, noise (the worst case for the lane arithmetic: sums near the
    16-bit limit), few-level images (exact .5 ties of the rounding), saturated images, tile-edge sizes."""
    from sdmb200 import synth
    rng = np.random.default_rng(11)
    cases = [np.ascontiguousarray(synth.make_scene(8, 320, 240, 6, seed=5, workers=4).im[:3])]
    for H, W in [(8, 8), (16, 64), (17, 68), (48, 64), (150, 200), (33, 8), (9, 132), (100, 60), (31, 128), (64, 252)]:
        im = rng.integers(0, 256, (4, H, W)).astype(np.uint8)
        im[1] = im[1] // 64 * 64
        im[2] = np.where(im[2] > 127, 255, 0)
        im[3] = 255
        cases.append(im)
    for im in cases:
        n, H, W = im.shape
        raw = str(tmp_path / "in.raw")
        im.tofile(raw)
        r = subprocess.run([ed_bin, str(W), str(H), str(n), raw, str(tmp_path / "out.bin"), "-", str(tmp_path / "planes.bin")],
                           capture_output=True, text=True)
        assert r.returncode == 0, (H, W, r.stderr)


@pytest.mark.skipif(not have_ref, reason="needs the reference's bundled OpenCV 2.4.5 (build container only)")
def test_smoothing_stage_against_the_bundled_cvSmooth():
    """cvSmooth(src, dst, CV_GAUSSIAN, 5, 5) through the C API of Thirdparty/EDTest/libopencv_imgproc.so.2.4.5.  In a child
    process: the library's cv:: symbols are loaded RTLD_GLOBAL and would interpose the stand-in cv::Mat of oracle/_ref."""
    subprocess.run([sys.executable, os.path.abspath(__file__), "--smoothing-check"], check=True)


def _smoothing_check():
    import ed_oracle
    core = ctypes.CDLL(os.path.join(ED_REF, "libopencv_core.so.2.4.5"), mode=ctypes.RTLD_GLOBAL)
    imgproc = ctypes.CDLL(os.path.join(ED_REF, "libopencv_imgproc.so.2.4.5"))
    core.cvCreateMatHeader.restype = ctypes.c_void_p
    core.cvCreateMatHeader.argtypes = [ctypes.c_int] * 3
    core.cvSetData.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    imgproc.cvSmooth.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_double] * 2
    rng = np.random.default_rng(3)
    for t in range(200):
        H, W = int(rng.integers(5, 90)), int(rng.integers(5, 140))
        im = rng.integers(0, 256, (H, W)).astype(np.uint8)
        if t % 3 == 0:
            im = (im // 64 * 64).astype(np.uint8)
        dst = np.zeros_like(im)
        a, b = core.cvCreateMatHeader(H, W, 0), core.cvCreateMatHeader(H, W, 0)
        core.cvSetData(a, im.ctypes.data, W)
        core.cvSetData(b, dst.ctypes.data, W)
        imgproc.cvSmooth(a, b, 2, 5, 5, 0.0, 0.0)
        assert np.array_equal(dst, ed_oracle.smooth(im)), (H, W)


if __name__ == "__main__" and sys.argv[1:] == ["--smoothing-check"]:
    _smoothing_check()
