"""Generate tests/golden/ed_chains_small.npz: the edge chains the reference's closed-source Edge Drawing library
(Thirdparty/EDTest/EDLib.a, called as LineDetector::DetectEdgeMap does, LineDetector.cc:855) finds on the images of a
small synthetic scene - the input of the 3-D line fitting (LineFitting :884-900) and of the mEdgeIndex
mask of the hot loop (:857-866, ProbabilityMapping.cc:454).  TEST INFRASTRUCTURE; runs only where /root/reference exists
(here), the fixture travels.  Run: make -C oracle ed && python oracle/make_ed_golden.py
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


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--bench":
        # chains for tools/linefit_bench.py: 32 VGA keyframes of the bench trajectory (config 2's seed); 5 MB, kept out of
        # git under oracle/_ref/ (travels to the GPU box with gpurun)
        main(os.path.join(HERE, "_ref", "ed_chains_vga.npz"), dict(n_kf=32, W=640, H=480, n_nbr=6, seed=2))
    else:
        main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "tests", "golden", "ed_chains_small.npz"))
