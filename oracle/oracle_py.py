"""ctypes wrapper around oracle/liboracle*.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (eao-slam_b200/) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OracleParams(C.Structure):
    _fields_ = [
        ("lambdaG", C.c_int), ("lambdaL", C.c_int), ("lambdaTheta", C.c_int), ("lambdaN", C.c_int),
        ("theta", C.c_float), ("sigmaI", C.c_float),
        ("chi2_fusion", C.c_double), ("chi2_inter", C.c_double), ("eps", C.c_double),
        ("slope_max", C.c_float), ("intra_check", C.c_int), ("intra_grow", C.c_int),
    ]


class OracleKF(C.Structure):
    _fields_ = [
        ("W", C.c_int), ("H", C.c_int),
        ("im", C.c_void_p), ("grad", C.c_void_p), ("theta", C.c_void_p), ("edge", C.c_void_p),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("Tcw", C.c_float * 12),
        ("depth", C.c_void_p), ("sigma", C.c_void_p), ("checked", C.c_void_p), ("points", C.c_void_p),
    ]


class OraclePair(C.Structure):
    _fields_ = [("R21", C.c_float * 9), ("t21", C.c_float * 3), ("F12", C.c_float * 9)]


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in
                ("candidates", "scanned", "evaluated", "hypotheses", "fused", "checked")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(native: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc only)."""
    target = ["native"] if native else []
    subprocess.run(["make", "-C", _HERE] + target, check=True, capture_output=True)


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build(native=(name == "liboracle_native.so"))
    lib = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    lib.oracle_default_params.argtypes = [C.POINTER(OracleParams)]
    lib.ocv_fastAtan2.restype = C.c_float
    lib.ocv_fastAtan2.argtypes = [C.c_float, C.c_float]
    lib.ocv_mul33_ABt.argtypes = [fp, fp, C.c_double, fp]
    lib.ocv_mul33.argtypes = [fp, fp, fp]
    lib.ocv_mul33_vec.argtypes = [fp, fp, C.c_double, fp, C.c_double, fp]
    lib.ocv_dot3_d.restype = C.c_float
    lib.ocv_dot3_d.argtypes = [fp, fp, C.c_double]
    lib.ocv_dotn_d.restype = C.c_float
    lib.ocv_dotn_d.argtypes = [fp, fp, C.c_int, C.c_double]
    lib.ocv_inv33.argtypes = [fp, fp]
    lib.ocv_solve33_lu.restype = C.c_int
    lib.ocv_solve33_lu.argtypes = [fp, fp, fp]
    lib.ocv_mul44_vec.argtypes = [fp, fp, fp]
    lib.oracle_pose_inverse.argtypes = [fp, fp]
    lib.oracle_pair_geometry.argtypes = [C.POINTER(OracleKF), C.POINTER(OracleKF), C.POINTER(OraclePair)]
    lib.oracle_stereo_search_constraints.argtypes = [fp, C.c_int, fp, fp]
    lib.oracle_get_search_range.argtypes = [C.POINTER(OracleKF), C.POINTER(OraclePair), C.c_int, C.c_int,
                                            C.c_float, C.c_float, fp, fp]
    lib.oracle_fusion.restype = C.c_int
    lib.oracle_fusion.argtypes = [fp, fp, C.c_int, C.POINTER(OracleParams), fp, fp]
    lib.oracle_intra_check.argtypes = [fp, fp, C.c_int, C.c_int, C.POINTER(OracleParams)]
    lib.oracle_intra_grow.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.POINTER(OracleParams)]
    lib.oracle_pass1_pair.argtypes = [C.POINTER(OracleKF), C.POINTER(OracleKF), C.c_float, C.c_float, C.c_float,
                                      C.POINTER(OracleParams), fp, fp, fp, C.POINTER(C.c_uint8)]
    lib.oracle_semidense_loop.restype = C.c_double
    lib.oracle_semidense_loop.argtypes = [C.POINTER(OracleKF), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_int32), fp, fp, fp, C.POINTER(OracleParams),
                                          C.c_int, C.POINTER(OracleStats)]
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_set_num_threads.argtypes = [C.c_int]
    lib.oracle_set_num_threads.restype = None
    return lib


_LIBS: dict[str, C.CDLL] = {}


def lib(kind: str = "canonical") -> C.CDLL:
    name = {"canonical": "liboracle.so", "fast": "liboracle_fast.so", "native": "liboracle_native.so"}[kind]
    if name not in _LIBS:
        _LIBS[name] = _load(name)
    return _LIBS[name]


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_params(kind: str = "canonical", **over) -> OracleParams:
    p = OracleParams()
    lib(kind).oracle_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


class OracleScene:
    """Holds numpy planes of a keyframe set in the oracle's struct layout and runs the loop."""

    def __init__(self, scene, kind: str = "canonical"):
        # scene: sdmb200.synth.Scene (im, grad, theta, edge|None, K, Tcw, nbr_idx, rot, min_depth, max_depth)
        self.kind = kind
        self.scene = scene
        n, H, W = scene.im.shape
        self.n, self.H, self.W = n, H, W
        self.depth = np.zeros((n, H, W), np.float32)
        self.sigma = np.zeros((n, H, W), np.float32)
        self.checked = np.zeros((n, H, W), np.float32)
        self.points = np.zeros((n, H, W, 3), np.float32)
        self.kfs = (OracleKF * n)()
        for i in range(n):
            self._fill(self.kfs[i], i)
        self.stats = OracleStats()

    def _fill(self, k: OracleKF, i: int):
        s = self.scene
        k.W, k.H = self.W, self.H
        k.im = s.im[i].ctypes.data
        k.grad = s.grad[i].ctypes.data
        k.theta = s.theta[i].ctypes.data
        k.edge = s.edge[i].ctypes.data if s.edge is not None else None
        k.fx, k.fy, k.cx, k.cy = [float(v) for v in s.K]
        for j, v in enumerate(np.asarray(s.Tcw[i], np.float32).reshape(-1)[:12]):
            k.Tcw[j] = float(v)
        k.depth = self.depth[i].ctypes.data
        k.sigma = self.sigma[i].ctypes.data
        k.checked = self.checked[i].ctypes.data
        k.points = self.points[i].ctypes.data

    def run(self, params: OracleParams | None = None, first: int = 0, count: int | None = None,
            pass_mask: int = 3) -> float:
        s = self.scene
        p = params or default_params(self.kind)
        count = self.n - first if count is None else count
        nbr = np.ascontiguousarray(s.nbr_idx, np.int32)
        rot = np.ascontiguousarray(s.rot, np.float32)
        mind = np.ascontiguousarray(s.min_depth, np.float32)
        maxd = np.ascontiguousarray(s.max_depth, np.float32)
        return lib(self.kind).oracle_semidense_loop(
            self.kfs, self.n, first, count, nbr.shape[1], nbr.ctypes.data_as(C.POINTER(C.c_int32)),
            fptr(rot), fptr(mind), fptr(maxd), C.byref(p), pass_mask, C.byref(self.stats))

    def pair(self, i: int, j: int) -> OraclePair:
        pr = OraclePair()
        lib(self.kind).oracle_pair_geometry(C.byref(self.kfs[i]), C.byref(self.kfs[j]), C.byref(pr))
        return pr

    def pass1_pair(self, i: int, j: int, rot: float, params: OracleParams | None = None):
        p = params or default_params(self.kind)
        s = self.scene
        d = np.zeros((self.H, self.W), np.float32)
        sg = np.zeros_like(d)
        u = np.zeros_like(d)
        ok = np.zeros((self.H, self.W), np.uint8)
        lib(self.kind).oracle_pass1_pair(C.byref(self.kfs[i]), C.byref(self.kfs[j]), float(rot),
                                         float(s.min_depth[i]), float(s.max_depth[i]), C.byref(p),
                                         fptr(d), fptr(sg), fptr(u), ok.ctypes.data_as(C.POINTER(C.c_uint8)))
        return d, sg, u, ok


def _bind_extra(l):
    l.oracle_inter_check.argtypes = [C.POINTER(OracleKF), C.c_int, C.POINTER(C.POINTER(OracleKF)), C.POINTER(OracleParams),
                                     C.POINTER(OracleStats)]
    l.oracle_update_points.argtypes = [C.POINTER(OracleKF), C.POINTER(OracleParams)]
    return l


def inter_check(osc: "OracleScene", i: int, params: OracleParams | None = None):
    """InterKeyFrameDepthChecking (:1121-1296) + UpdateSemiDensePointSet (:700-731) of keyframe i alone."""
    l = _bind_extra(lib(osc.kind))
    p = params or default_params(osc.kind)
    nb = osc.scene.nbr_idx[i]
    arr = (C.POINTER(OracleKF) * len(nb))(*[C.pointer(osc.kfs[int(j)]) for j in nb])
    l.oracle_inter_check(C.byref(osc.kfs[i]), len(nb), arr, C.byref(p), None)
    l.oracle_update_points(C.byref(osc.kfs[i]), C.byref(p))


def update_points(osc: "OracleScene", i: int, Tcw=None, params: OracleParams | None = None):
    """UpdateSemiDensePointSet (:700-731) of keyframe i, optionally after KeyFrame::SetPose(Tcw)."""
    l = _bind_extra(lib(osc.kind))
    p = params or default_params(osc.kind)
    if Tcw is not None:
        for j, v in enumerate(np.asarray(Tcw, np.float32).reshape(-1)[:12]):
            osc.kfs[i].Tcw[j] = float(v)
    l.oracle_update_points(C.byref(osc.kfs[i]), C.byref(p))
