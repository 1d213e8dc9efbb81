"""Generate tests/golden/ed_chains_small.npz: the edge chains the reference's closed-source Edge Drawing library
(Thirdparty/EDTest/EDLib.a, called as LineDetector::DetectEdgeMap does, LineDetector.cc:855) finds on the images of a
small synthetic scene - the input of the 3-D line fitting (LineFitting :884-900) and of the mEdgeIndex
mask of the hot loop (:857-866, ProbabilityMapping.cc:454).  TEST INFRASTRUCTURE; runs only where /root/reference exists
(here), the fixture travels.  Run: make -C oracle ed && python oracle/make_ed_golden.py
`--misc` writes tests/golden/ed_chains_misc.npz: the library's chains on images that are not keyframes of the scene - noise,
flat, ramps, drawn shapes, smooth blobs quantised to a few grey levels (long gradient ties), sizes that are not multiples of
four (the scalar tail of the library's smoothing rounds differently), a 5 x 7 image, a 6-row strip - the fixture of
tests/test_edge_drawing.py for the open implementation (eao-slam_b200/host/edge_drawing.h) and its restatement
(oracle/ed_oracle.py).  The images are stored in the fixture next to the chains.
"""
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "eao-slam_b200", "python"))
from sdmb200 import synth  # noqa: E402

ED = "/root/reference/Thirdparty/EDTest"
SCENE = dict(n_kf=6, W=320, H=240, n_nbr=4, seed=77)


def ed_chains(images):
    """images (N, H, W) uint8 -> list of N lists of (n_i, 2) int32 arrays (r, c)"""
    n, H, W = images.shape
    with tempfile.TemporaryDirectory() as tmp:
        for lib in ("libopencv_core.so.2.4", "libopencv_imgproc.so.2.4"):
            os.symlink(os.path.join(ED, lib + ".5"), os.path.join(tmp, lib))
        raw, out = os.path.join(tmp, "in.raw"), os.path.join(tmp, "out.bin")
        np.ascontiguousarray(images).tofile(raw)
        env = dict(os.environ, LD_LIBRARY_PATH=tmp)
        subprocess.run([os.path.join(HERE, "_ref", "ed_chains"), str(W), str(H), str(n), raw, out], check=True, env=env)
        a = np.fromfile(out, np.int32)
    assert a[0] == n
    p, res = 1, []
    for _ in range(n):
        ns = int(a[p]); p += 1
        chains = []
        for _ in range(ns):
            m = int(a[p]); p += 1
            chains.append(a[p:p + 2 * m].reshape(m, 2).copy()); p += 2 * m
        res.append(chains)
    assert p == a.size
    return res


def main(out, scene=SCENE):
    sc = synth.make_scene(scene["n_kf"], scene["W"], scene["H"], scene["n_nbr"], seed=scene["seed"], workers=8)
    chains = ed_chains(sc.im)
    d = dict(scene=np.array([scene[k] for k in ("n_kf", "W", "H", "n_nbr", "seed")], np.int32),
             im_crc=np.array([zlib.crc32(sc.im.tobytes())], np.uint32))
    for i, ch in enumerate(chains):
        lens = np.array([len(c) for c in ch], np.int32)
        d[f"off_{i}"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        rc = np.concatenate(ch) if ch else np.zeros((0, 2), np.int32)
        d[f"pix_{i}"] = ((rc[:, 0].astype(np.uint32) << 16) | rc[:, 1].astype(np.uint32)).astype(np.uint32)
        print(f"keyframe {i}: {len(ch)} chains, {int(lens.sum())} pixels, longest {int(lens.max()) if len(ch) else 0}")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out))


def blobs(rng, H, W, sigma, levels):
    """smooth random field (separable box blurs, numpy only), optionally quantised to `levels` grey levels"""
    f = rng.normal(size=(H + 4 * sigma, W + 4 * sigma))
    for _ in range(3):
        c = np.cumsum(np.pad(f, ((sigma, 0), (0, 0))), axis=0); f = c[sigma:] - c[:-sigma]
        c = np.cumsum(np.pad(f, ((0, 0), (sigma, 0))), axis=1); f = c[:, sigma:] - c[:, :-sigma]
    f = f[2 * sigma:2 * sigma + H, 2 * sigma:2 * sigma + W]
    f = (f - f.min()) / (np.ptp(f) + 1e-12)
    if levels:
        f = np.round(f * levels) / levels
    return np.clip(f * 255, 0, 255).astype(np.uint8)


def shapes(rng, H, W, n):
    im = np.full((H, W), int(rng.integers(20, 90)), np.float64)
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(n):
        v = float(rng.integers(0, 256))
        k = int(rng.integers(0, 3))
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        if k == 0:
            im[(np.abs(yy - cy) < rng.uniform(2, max(H / 3, 3))) & (np.abs(xx - cx) < rng.uniform(2, max(W / 3, 3)))] = v
        elif k == 1:
            im[(yy - cy) ** 2 + (xx - cx) ** 2 < rng.uniform(2, max(min(H, W) / 3, 3)) ** 2] = v
        else:
            a = rng.uniform(0, np.pi)
            im[np.abs((yy - cy) * np.cos(a) - (xx - cx) * np.sin(a)) < rng.uniform(1, 4)] = v
    return np.clip(im + rng.normal(size=(H, W)) * rng.uniform(0, 3), 0, 255).astype(np.uint8)


def misc_images():
    rng = np.random.default_rng(20261019)
    yy, xx = np.mgrid[0:120, 0:160]
    ims = dict(
        noise=rng.integers(0, 256, (150, 200)).astype(np.uint8),
        black=np.zeros((48, 64), np.uint8),
        ramp=np.clip(np.mgrid[0:77, 0:203][1] * 1.3 + (np.mgrid[0:77, 0:203][0] > 40) * 90, 0, 255).astype(np.uint8),
        checker=(((yy // 8 + xx // 8) % 2) * 200 + 20).astype(np.uint8),
        shapes0=shapes(rng, 97, 131, 7), shapes1=shapes(rng, 240, 320, 14), shapes2=shapes(rng, 333, 257, 12),
        shapes3=shapes(rng, 77, 203, 9),
        tiny=rng.integers(0, 256, (5, 7)).astype(np.uint8),
        thin=shapes(rng, 6, 300, 8),
        blobs_w70=blobs(rng, 260, 70, 3, 5), blobs_w179=blobs(rng, 243, 179, 4, 3), blobs_w257=blobs(rng, 120, 257, 2, 6),
        blobs_w598=blobs(rng, 310, 598, 5, 4), blobs_smooth=blobs(rng, 200, 202, 2, 0),
    )
    return ims


def main_misc(out):
    ims = misc_images()
    d = dict(names=np.array(list(ims.keys())))
    for name, im in ims.items():
        ch = ed_chains(im[None])[0]
        lens = np.array([len(c) for c in ch], np.int32)
        d["im_" + name] = im
        d["off_" + name] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        rc = np.concatenate(ch) if ch else np.zeros((0, 2), np.int32)
        d["pix_" + name] = ((rc[:, 0].astype(np.uint32) << 16) | rc[:, 1].astype(np.uint32)).astype(np.uint32)
        print(f"{name} {im.shape}: {len(ch)} chains, {int(lens.sum())} pixels")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--misc":
        main_misc(os.path.join(HERE, "..", "tests", "golden", "ed_chains_misc.npz"))
    elif len(sys.argv) > 1 and sys.argv[1] == "--bench":
        # chains for tools/linefit_bench.py: 32 VGA keyframes of the bench trajectory (config 2's seed); 5 MB, kept out of
        # git under oracle/_ref/ (travels to the GPU box with gpurun)
        main(os.path.join(HERE, "_ref", "ed_chains_vga.npz"), dict(n_kf=32, W=640, H=480, n_nbr=6, seed=2))
    else:
        main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "tests", "golden", "ed_chains_small.npz"))
