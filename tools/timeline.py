"""Developer probe: per-stream timeline (SDM_TRACE=1) of a chunked upload / pass-1 / pass-2 loop, printed by the library."""
import os, sys, time, ctypes as C
os.environ["SDM_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eao-slam_b200", "python")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from sdmb200 import api, synth
n = 100
sc = synth.make_scene(n, 640, 480, 6, seed=2, workers=8)
lib = api.load()
ctx = api.Context(width=640, height=480, max_keyframes=n, intra_check=1, intra_grow=1)
keep = []
for k in ("im", "grad", "theta"):
    a = getattr(sc, k); h = bench.pinned(lib, a.shape, a.dtype, keep); h[:] = a; setattr(sc, k, h)
up = ctx.upload_descs(sc, range(n))
items = api.make_items(range(n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
CH = 25
chunks = [list(range(i, min(n, i + CH))) for i in range(0, n, CH)]
citems = [api.make_items(ch, sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth) for ch in chunks]
need = [max(ch[-1], max(int(v) for s in ch for v in sc.nbr_idx[s])) for ch in chunks]
usz = C.sizeof(api.UploadDesc)
uptr = lambda i: C.cast(C.byref(up, i * usz), C.POINTER(api.UploadDesc))
def chunked():
    nxt = 0
    for k in range(len(chunks)):
        if nxt <= need[k]:
            ctx._chk(lib.sdm_upload_keyframes(ctx.h, need[k] + 1 - nxt, uptr(nxt))); nxt = need[k] + 1
        ctx.pass1(citems[k])
    ctx.pass2(items)
chunked(); ctx.synchronize()
print("---- second iteration ----", file=sys.stderr)
chunked(); ctx.synchronize()
