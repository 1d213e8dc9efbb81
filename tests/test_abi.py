"""The C-ABI library: loads, exports every symbol include/sdm_b200.h declares, host-side entry points
agree bit-exactly with the oracle, and a missing GPU is a loud error (never a fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_py as O
from sdmb200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sdm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdm_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header():
    lib = api.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sdm_b200.h but not exported"
    assert sorted(api.EXPORTS) == syms, "python binding list out of sync with the header"


def test_struct_layouts_match_header():
    """sizes the C compiler gives the ABI structs == the ctypes mirrors (guards silent ABI drift)"""
    import subprocess, tempfile
    prog = r'''
#include <stdio.h>
#include "sdm_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(sdm_config), sizeof(sdm_item), sizeof(sdm_pair_geometry_t),
 sizeof(sdm_hypothesis), sizeof(sdm_stats), sizeof(sdm_timing), sizeof(sdm_upload_desc), sizeof(sdm_download_desc),
 sizeof(sdm_point)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"),
                        os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    want = [C.sizeof(x) for x in (api.Config, api.Item, api.PairGeometry, api.Hypothesis, api.Stats, api.Timing,
                                  api.UploadDesc, api.DownloadDesc)] + [16]
    assert [int(v) for v in out] == want


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU error path")
def test_no_gpu_is_a_loud_error():
    with pytest.raises(api.SdmError) as e:
        api.Context(width=64, height=48, max_keyframes=2)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_bad_config_rejected():
    lib = api.load()
    cfg = api.default_config(width=4, height=4)
    h = C.c_void_p()
    assert lib.sdm_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.sdm_create(None, C.byref(h)) == -1
    assert b"null" in lib.sdm_last_error()


def test_default_config_is_the_reference_defines():
    cfg = api.default_config()
    # include/ProbabilityMapping.h:45-56 and the literals of ProbabilityMapping.cc
    assert (cfg.lambdaG, cfg.lambdaL, cfg.lambdaTheta, cfg.lambdaN) == (8, 80, 45, 3)
    assert np.float32(cfg.theta) == np.float32(0.23) and cfg.sigmaI == 20.0
    assert (cfg.chi2_fusion, cfg.chi2_inter, cfg.eps, cfg.slope_max) == (5.99, 3.84, 0.000001, 4.0)
    assert cfg.intra_check == 0 and cfg.intra_grow == 0  # :491-494 are commented out in the shipped loop


def test_pair_geometry_bit_exact_vs_oracle():
    sc = synth.make_scene(9, 96, 72, 6, seed=4)
    osc = O.OracleScene(sc)
    for i in range(sc.n):
        for j in sc.nbr_idx[i]:
            g, p = api.pair_geometry(sc.K, sc.Tcw[i], sc.K, sc.Tcw[int(j)]), osc.pair(i, int(j))
            for f in ("R21", "t21", "F12"):
                a = np.array(getattr(g, f)[:], np.float32).view(np.uint32)
                b = np.array(getattr(p, f)[:], np.float32).view(np.uint32)
                assert np.array_equal(a, b), (i, j, f)


def test_pair_geometry_vs_cv2_golden(golden_dir):
    """R21 / t21 / F12 evaluated by REAL cv2 calls (oracle/pin_cv2.py: gemm / solve / invert in the
    order OpenCV's MatExpr evaluates :1136-1137 and :1700-1708) == the library's host arithmetic."""
    kat = np.load(os.path.join(golden_dir, "pair_geometry_cv2.npz"))
    for K1, T1, K2, T2, R21, t21, F12 in zip(kat["K1"], kat["T1"], kat["K2"], kat["T2"], kat["R21"], kat["t21"], kat["F12"]):
        g = api.pair_geometry(K1, T1, K2, T2)
        assert np.array_equal(np.array(g.R21[:], np.float32).view(np.uint32), R21.reshape(-1).view(np.uint32))
        assert np.array_equal(np.array(g.t21[:], np.float32).view(np.uint32), t21.reshape(-1).view(np.uint32))
        assert np.array_equal(np.array(g.F12[:], np.float32).view(np.uint32), F12.reshape(-1).view(np.uint32))


def test_stereo_search_constraints_vs_oracle():
    rng = np.random.default_rng(5)
    lib = O.lib()
    for n in (1, 7, 500, 1000):
        d = rng.uniform(0.2, 2.5, n).astype(np.float32)
        a, b = C.c_float(), C.c_float()
        lib.oracle_stereo_search_constraints(O.fptr(d), n, C.byref(a), C.byref(b))
        lo, hi = api.stereo_search_constraints(d)
        assert np.float32(lo).view(np.uint32) == np.float32(a.value).view(np.uint32)
        assert np.float32(hi).view(np.uint32) == np.float32(b.value).view(np.uint32)
