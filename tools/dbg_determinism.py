import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("eao-slam_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
from sdmb200 import api, synth
tag = sys.argv[1]
sc = synth.make_scene(13, 320, 240, 6, seed=41)
ctx = api.Context(width=320, height=240, max_keyframes=13)
ctx.upload_scene(sc)
items = api.make_items(range(13), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
ref = None
t0 = time.time()
it = 0
while time.time() - t0 < float(sys.argv[2]):
    ctx.pass1(items); ctx.pass2(items)
    cur = [ctx.download(i) for i in range(10)]
    if ref is None:
        ref = cur
    else:
        for i in range(10):
            for k in ("depth", "sigma", "checked"):
                n = int((cur[i][k].view(np.uint32) != ref[i][k].view(np.uint32)).sum())
                if n:
                    print(tag, "iter", it, "kf", i, k, "differs in", n, "words", flush=True)
    it += 1
print(tag, "done", it, "iterations", flush=True)
