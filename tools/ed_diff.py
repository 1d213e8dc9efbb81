"""Differential test of two revisions of the open Edge Drawing implementation (eao-slam_b200/host/edge_drawing.h behind
tests/cpp/test_edge_drawing.cpp): both drivers are built (the old one from `git show REV:...` into a scratch tree), run on
the same random images - noise, few grey levels, binary, blocks, quantised blocks + noise, sinusoid fields, three-level - of
random sizes, and their chain dumps and edge-index planes must be byte-identical.  The driver of each revision also checks its
fixed-capacity form (what k_ed_route runs) against its vector form on every image.  CPU only.
    python tools/ed_diff.py [--old-rev HEAD~1] [--rounds 140] [--per-round 24] [--seed 12345]"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["eao-slam_b200/host/edge_drawing.h", "eao-slam_b200/csrc/edge_drawing_kernels.cuh", "tests/cpp/test_edge_drawing.cpp"]


def build_old(rev, tmp):
    for f in FILES:
        dst = os.path.join(tmp, "old", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "wb") as o:
            o.write(subprocess.check_output(["git", "-C", ROOT, "show", f"{rev}:{f}"]))
    exe = os.path.join(tmp, "ted_old")
    subprocess.run(["g++", "-O2", "-std=c++11", "-o", exe, os.path.join(tmp, "old", FILES[2])], check=True)
    return exe


def images(rng, kind, n, H, W):
    if kind == 0:
        ims = rng.integers(0, 256, (n, H, W))
    elif kind == 1:
        ims = rng.integers(0, 4, (n, H, W)) * 64
    elif kind == 2:
        ims = rng.integers(0, 2, (n, H, W)) * 255
    elif kind == 3:
        ims = rng.integers(0, 256, (n, H // 4 + 2, W // 4 + 2)).repeat(4, 1).repeat(4, 2)[:, :H, :W]
    elif kind == 4:
        ims = rng.integers(0, 256, (n, H // 8 + 2, W // 8 + 2)).repeat(8, 1).repeat(8, 2)[:, :H, :W] // 32 * 32 + rng.integers(0, 8, (n, H, W))
    elif kind == 5:
        yy, xx = np.mgrid[0:H, 0:W]
        ims = np.stack([(np.sin(xx / rng.uniform(2, 9) + rng.uniform(0, 6)) + np.cos(yy / rng.uniform(2, 9))) * 60 + 128 +
                        rng.integers(0, 20, (H, W)) for _ in range(n)])
    else:
        ims = rng.integers(0, 256, (n, H, W))
        ims = np.where(ims > 200, 255, np.where(ims < 60, 0, 128))
    return np.ascontiguousarray(np.clip(ims, 0, 255).astype(np.uint8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--old-rev", default="HEAD~1")
    ap.add_argument("--rounds", type=int, default=140)
    ap.add_argument("--per-round", type=int, default=24)
    ap.add_argument("--seed", type=int, default=12345)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    with tempfile.TemporaryDirectory() as tmp:
        old = build_old(a.old_rev, tmp)
        new = os.path.join(tmp, "ted_new")
        subprocess.run(["g++", "-O2", "-std=c++11", "-o", new, os.path.join(ROOT, FILES[2])], check=True)

        def run(exe, ims, tag):
            n, H, W = ims.shape
            raw = os.path.join(tmp, "in.raw")
            ims.tofile(raw)
            out, edge = os.path.join(tmp, f"chains_{tag}.bin"), os.path.join(tmp, f"edge_{tag}.bin")
            r = subprocess.run([exe, str(W), str(H), str(n), raw, out, edge], capture_output=True, text=True)
            assert r.returncode == 0, (tag, r.returncode, r.stderr)
            return open(out, "rb").read(), open(edge, "rb").read()
        total = 0
        for it in range(a.rounds):
            H, W = int(rng.integers(12, 260)), int(rng.integers(12, 340))
            ims = images(rng, it % 7, a.per_round, H, W)
            assert run(old, ims, "old") == run(new, ims, "new"), f"round {it}: {H}x{W}, kind {it % 7}"
            total += len(ims)
    print(f"{total} images: chains and edge-index planes identical between {a.old_rev} and the working tree")


if __name__ == "__main__":
    main()
