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


LIB_ONLINE = os.path.join(_HERE, "_ref", "libref_pm_online.so")
_lib_online = None


def lib_online():
    """the same sources compiled with -DOnlineLoop (the reference's online mode, ProbabilityMapping.cc:42)"""
    global _lib_online
    if _lib_online is None:
        if not os.path.exists(LIB_ONLINE) and not (available() and os.path.exists(LIB_ONLINE)):
            raise FileNotFoundError(LIB_ONLINE)
        _lib_online = C.CDLL(LIB_ONLINE)
        assert _lib_online.ref_online_build() == 1
    return _lib_online


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run_reference_online(scene, n1, n2, extra_ids, moved, Tcw_moved):
    """ref_online_sequence (oracle/refshim/refdriver.cc) of the OnlineLoop build: keyframes arrive in three batches
    [0,n1), [n1,n2), [n2,n); after each one `SemiDenseLoop(); UpdateAllSemiDensePointSet();` (:224-226); the keyframes
    in `moved` get new poses before the second round; a final SemiDenseLoop() after `extra_ids` more mapped keyframes."""
    l = lib_online()
    n, H, W = scene.im.shape
    c = np.ascontiguousarray
    inv = c(np.stack(scene.inv_depths).astype(np.float32))
    out = {k: np.zeros((n, H, W) + ((3,) if k == "points" else ()), np.float32) for k in ("depth", "sigma", "checked", "points")}
    flags = np.zeros((n, 2), np.int32)
    im, g, t = c(scene.im, np.uint8), c(scene.grad, np.float32), c(scene.theta, np.float32)
    K, T, nb = c(np.asarray(scene.K, np.float32)), c(scene.Tcw.reshape(n, 12), np.float32), c(scene.nbr_idx, np.int32)
    mv, Tm = c(moved, np.int32), c(np.asarray(Tcw_moved, np.float32).reshape(len(moved), 12))
    rc = l.ref_online_sequence(n, int(n1), int(n2), W, H, _p(im), _p(g), _p(t), _p(K), _p(T), int(nb.shape[1]), _p(nb), _p(inv),
                               inv.shape[1], int(extra_ids), len(mv), _p(mv), _p(Tm), _p(out["depth"]), _p(out["sigma"]),
                               _p(out["checked"]), _p(out["points"]), _p(flags))
    assert rc == 0
    out["flags"] = flags
    return out


def reference_save_points(im, sigma, checked, points, flags, bad=None, rgb=None):
    """the reference's own SaveSemiDensePoints (:136-192) on these planes; writes
    ./results_line_segments/refshim/semi_pointcloud.obj (the directory must exist) and returns its path"""
    l = lib()
    c = np.ascontiguousarray
    n, H, W = im.shape
    im, sigma, checked, points = c(im, np.uint8), c(sigma, np.float32), c(checked, np.float32), c(points, np.float32)
    flags = c(flags, np.int32)
    badv = c(bad, np.int32) if bad is not None else np.zeros(n, np.int32)
    os.makedirs(os.path.join("results_line_segments", "refshim"), exist_ok=True)
    rc = l.ref_save_semidense_points(n, W, H, _p(im), _p(c(rgb, np.uint8)) if rgb is not None else None, _p(sigma), _p(checked),
                                     _p(points), _p(flags), _p(badv))
    assert rc == 0
    return os.path.join("results_line_segments", "refshim", "semi_pointcloud.obj")
