"""Generate tests/golden/oracle_small.npz: inputs + canonical-oracle outputs of a tiny SemiDenseLoop
(8 keyframes, 96x72, 6 neighbours; default mode and intra-checks mode).  Committed so that
 (a) the oracle is pinned against silent drift (source edits, compiler flags), and
 (b) the GPU tests can compare the CUDA path with fixed vectors that do not depend on the oracle
     being rebuilt on the GPU box.
Run here: python oracle/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "eao-slam_b200", "python"))
sys.path.insert(0, HERE)
import oracle_py as O  # noqa: E402
from sdmb200 import synth  # noqa: E402


def main(out):
    sc = synth.make_scene(8, 96, 72, 6, seed=31, contrast=0.9)
    sc.rot[:] = np.random.default_rng(1).uniform(-3, 3, sc.rot.shape).astype(np.float32)
    d = dict(im=sc.im, grad=sc.grad, theta=sc.theta, K=np.asarray(sc.K, np.float32), Tcw=sc.Tcw,
             nbr_idx=sc.nbr_idx, rot=sc.rot, min_depth=sc.min_depth, max_depth=sc.max_depth)
    for tag, intra in (("plain", 0), ("intra", 1)):
        osc = O.OracleScene(sc, "canonical")
        osc.run(params=O.default_params("canonical", intra_check=intra, intra_grow=intra))
        d[f"{tag}_depth"], d[f"{tag}_sigma"] = osc.depth, osc.sigma
        d[f"{tag}_checked"], d[f"{tag}_points"] = osc.checked, osc.points
        st = osc.stats.as_dict()
        d[f"{tag}_stats"] = np.array([st[k] for k in ("candidates", "scanned", "evaluated", "hypotheses", "fused", "checked")], np.int64)
        print(tag, st)
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "tests", "golden", "oracle_small.npz"))
