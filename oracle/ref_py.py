"""ctypes wrapper around oracle/_ref/libref_pm.so — the REFERENCE'S OWN src/ProbabilityMapping.cc compiled against the
stand-in headers of oracle/refshim/ (see oracle/Makefile, target `ref`).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libref_pm.so")
REFERENCE = "/root/reference"


def available(build: bool = True) -> bool:
    """True if the library exists (prebuilt, e.g. on the GPU box) or can be built here (reference tree present)."""
    if os.path.exists(LIB):
        return True
    if build and os.path.isdir(os.path.join(REFERENCE, "src")):
        r = subprocess.run(["make", "-C", _HERE, "ref"], capture_output=True, text=True)
        return r.returncode == 0 and os.path.exists(LIB)
    return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(LIB)
        _lib = C.CDLL(LIB)
        _lib.ref_covisN.restype = C.c_int
    return _lib


def run_reference_loop(scene, kp_angle, first_id=1, extra_ids=11, bad=None):
    """SemiDenseLoop() of the reference over a synth.Scene whose covisibility lists (scene.nbr_idx) hold at least
    covisN (= 7) keyframes.  first_id / extra_ids / bad drive the reference's own gating (mapping ids, isBad()).
    Returns dict(depth, sigma, checked, points, flags)."""
    l = lib()
    n, H, W = scene.im.shape
    N = l.ref_covisN()
    assert scene.nbr_idx.shape[0] == n and scene.nbr_idx.shape[1] >= N, f"the reference is compiled with covisN = {N}"
    inv = np.ascontiguousarray(np.stack(scene.inv_depths).astype(np.float32))
    out = {k: np.zeros((n, H, W) + ((3,) if k == "points" else ()), np.float32) for k in ("depth", "sigma", "checked", "points")}
    flags = np.zeros((n, 2), np.int32)
    c = np.ascontiguousarray
    args = [c(scene.im, np.uint8), c(scene.grad, np.float32), c(scene.theta, np.float32)]
    edge = c(scene.edge, np.int32) if scene.edge is not None else None
    K = c(np.asarray(scene.K, np.float32))
    T = c(scene.Tcw.reshape(n, 12), np.float32)
    nb = c(scene.nbr_idx, np.int32)
    ang = c(kp_angle, np.float32)
    badv = c(bad, np.int32) if bad is not None else np.zeros(n, np.int32)
    rc = l.ref_semidense_loop_ex(n, W, H, args[0].ctypes.data_as(C.c_void_p), args[1].ctypes.data_as(C.c_void_p),
                                 args[2].ctypes.data_as(C.c_void_p), edge.ctypes.data_as(C.c_void_p) if edge is not None else None,
                                 K.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p), int(nb.shape[1]),
                                 nb.ctypes.data_as(C.c_void_p), ang.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p),
                                 inv.shape[1], int(first_id), int(extra_ids), badv.ctypes.data_as(C.c_void_p),
                                 out["depth"].ctypes.data_as(C.c_void_p), out["sigma"].ctypes.data_as(C.c_void_p),
                                 out["checked"].ctypes.data_as(C.c_void_p), out["points"].ctypes.data_as(C.c_void_p),
                                 flags.ctypes.data_as(C.c_void_p))
    assert rc == 0
    out["flags"] = flags
    return out


def run_reference_intra(which: int, depth, sigma, grad):
    """IntraKeyFrameDepthChecking (which=0) / Growing (which=1) of the reference on copies of the planes."""
    d, s = np.ascontiguousarray(depth, np.float32).copy(), np.ascontiguousarray(sigma, np.float32).copy()
    g = np.ascontiguousarray(grad, np.float32)
    H, W = d.shape
    rc = lib().ref_intra(which, W, H, d.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return d, s
