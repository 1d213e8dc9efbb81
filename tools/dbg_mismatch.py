import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("eao-slam_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
from helpers import run_device, run_oracle
from sdmb200 import synth
sc = synth.make_scene(20, 320, 240, 6, seed=41)
dev = run_device(sc); osc = run_oracle(sc)
for k, ref in (("depth", osc.depth), ("sigma", osc.sigma), ("checked", osc.checked)):
    bad = np.argwhere(dev[k].view(np.uint32) != ref.view(np.uint32))
    print(k, len(bad))
    for i, y, x in bad[:6]:
        print("  kf", i, "y", y, "x", x, "dev", repr(dev[k][i, y, x]), "ref", repr(ref[i, y, x]), "rho", repr(osc.depth[i, y, x]))
