"""Developer timing probe (not the bench): device-time of pass 1 / pass 2 on a small VGA scene."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eao-slam_b200", "python"))
from sdmb200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sc = synth.make_scene(n, 640, 480, 6, seed=2)
ctx = api.Context(width=640, height=480, max_keyframes=n)
t = time.time(); ctx.upload_scene(sc); ctx.synchronize(); up = time.time() - t
items = api.make_items(range(n), sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
for it in range(3):
    ctx.pass1(items); ctx.pass2(items); ctx.synchronize()
    p1, p2 = ctx.last_pass_ms()
    print(json.dumps({"n_kf": n, "upload_s": round(up, 3), "pass1_ms": p1, "pass2_ms": p2,
                      "ms_per_kf": (p1 + p2) / n, "stats": ctx.stats()}))
