"""ctypes binding of libsdm_b200.so (include/sdm_b200.h) — harness-side only.

The product is the C-ABI library; this module is what tests/ and bench.py use to reach it from
Python.  There is no CPU fallback here either: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SDM_MAX_NBR = 16
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDM_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libsdm_b200.so"))


class SdmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sdm error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("max_keyframes", C.c_int),
        ("lambdaG", C.c_int), ("lambdaL", C.c_int), ("lambdaTheta", C.c_int), ("lambdaN", C.c_int),
        ("theta", C.c_float), ("sigmaI", C.c_float),
        ("chi2_fusion", C.c_double), ("chi2_inter", C.c_double), ("eps", C.c_double),
        ("slope_max", C.c_float), ("intra_check", C.c_int), ("intra_grow", C.c_int), ("device", C.c_int),
    ]


class Item(C.Structure):
    _fields_ = [
        ("kf", C.c_int32), ("n_nbr", C.c_int32), ("nbr", C.c_int32 * SDM_MAX_NBR),
        ("rot_deg", C.c_float * SDM_MAX_NBR), ("min_depth", C.c_float), ("max_depth", C.c_float),
    ]


class PairGeometry(C.Structure):
    _fields_ = [("R21", C.c_float * 9), ("t21", C.c_float * 3), ("F12", C.c_float * 9)]


class Hypothesis(C.Structure):
    _fields_ = [("depth", C.c_float), ("sigma", C.c_float), ("supported", C.c_int32),
                ("best_u", C.c_float), ("best_v", C.c_float)]


class UploadDesc(C.Structure):
    _fields_ = [("kf", C.c_int32), ("im", C.c_void_p), ("im_step", C.c_size_t), ("grad", C.c_void_p), ("grad_step", C.c_size_t),
                ("theta", C.c_void_p), ("theta_step", C.c_size_t), ("edge", C.c_void_p), ("edge_step", C.c_size_t),
                ("K", C.c_float * 4), ("Tcw", C.c_float * 12)]


class DownloadDesc(C.Structure):
    _fields_ = [("kf", C.c_int32), ("depth", C.c_void_p), ("depth_step", C.c_size_t), ("sigma", C.c_void_p),
                ("sigma_step", C.c_size_t), ("checked", C.c_void_p), ("checked_step", C.c_size_t),
                ("points", C.c_void_p), ("points_step", C.c_size_t)]


class Loop(C.Structure):
    _fields_ = [("n_upload", C.c_int32), ("upload", C.POINTER(UploadDesc)), ("n_pass1", C.c_int32), ("pass1", C.POINTER(Item)),
                ("down1", C.POINTER(DownloadDesc)), ("n_pass2", C.c_int32), ("pass2", C.POINTER(Item)),
                ("down2", C.POINTER(DownloadDesc)), ("chunk", C.c_int32), ("exchange", C.c_int32), ("sparse_download", C.c_int32)]


SDM_PEER_HANDLE_BYTES = 192


class EdgeChains(C.Structure):  # sdm_edge_chains
    _fields_ = [("kf", C.c_int32), ("n_chains", C.c_int32), ("offsets", C.POINTER(C.c_int32)), ("pixels", C.POINTER(C.c_uint32))]


LINE3D = np.dtype([("seg", "f4", 4), ("xyz", "f4", 6), ("chain", "i4"), ("kf_index", "i4")])  # sdm_line3d


class EdImage(C.Structure):  # sdm_ed_image
    _fields_ = [("im", C.c_void_p), ("im_step", C.c_size_t), ("edge_index", C.c_void_p), ("edge_step", C.c_size_t)]


class Timing(C.Structure):
    _fields_ = [("pass1_scan_ms", C.c_float), ("pass1_intra_ms", C.c_float), ("pass2_ms", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("candidates", C.c_longlong), ("fused", C.c_longlong), ("checked", C.c_longlong)]


# every symbol include/sdm_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "sdm_default_config", "sdm_create", "sdm_destroy", "sdm_last_error", "sdm_version", "sdm_synchronize",
    "sdm_get_stats", "sdm_scan_generation", "sdm_host_alloc", "sdm_host_free", "sdm_upload_keyframe", "sdm_set_pose",
    "sdm_set_intrinsics", "sdm_candidate_count", "sdm_candidate_blocks", "sdm_pass1", "sdm_pass2", "sdm_update_points", "sdm_download", "sdm_download_async", "sdm_upload_keyframes", "sdm_download_keyframes", "sdm_scatter_keyframes", "sdm_export_points", "sdm_download_planes",
    "sdm_upload_depth", "sdm_upload_checked", "sdm_depth_plane_ptr", "sdm_export_arena", "sdm_import_peer_arena", "sdm_pull_halo",
    "sdm_mark_pass1_done", "sdm_export_peer_handle", "sdm_import_peer", "sdm_set_halo", "sdm_exchange", "sdm_run_loop",
    "sdm_pair_geometry", "sdm_stereo_search_constraints", "sdm_search_range", "sdm_inter_chi_test",
    "sdm_epipolar_search", "sdm_epipolar_search_plane", "sdm_fuse", "sdm_intra_check", "sdm_intra_grow",
    "sdm_inter_check", "sdm_last_pass_ms", "sdm_launch_count", "sdm_last_timing", "sdm_last_pack_ms", "sdm_mark", "sdm_elapsed_ms",
    "sdm_line_fit", "sdm_last_line_fit_ms", "sdm_last_scan_long",
    "sdm_edge_drawing", "sdm_ed_chains", "sdm_ed_free", "sdm_last_edge_drawing_ms", "sdm_ed_planes",
    "sdm_set_edge_drawing_route", "sdm_last_edge_drawing_fallbacks", "sdm_ed_device_edge_plane",
]

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    sz = C.c_size_t
    lib.sdm_default_config.argtypes = [C.POINTER(Config)]
    lib.sdm_default_config.restype = None
    lib.sdm_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.sdm_destroy.argtypes = [vp]
    lib.sdm_destroy.restype = None
    lib.sdm_last_error.restype = C.c_char_p
    lib.sdm_version.restype = C.c_char_p
    lib.sdm_synchronize.argtypes = [vp]
    lib.sdm_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.sdm_scan_generation.argtypes = [vp]
    lib.sdm_edge_drawing.argtypes = [vp, C.c_int, C.POINTER(EdImage), C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.sdm_ed_chains.argtypes = [vp, C.c_int, ip, C.POINTER(ip), C.POINTER(C.POINTER(C.c_uint32))]
    lib.sdm_ed_free.argtypes = [vp]
    lib.sdm_ed_free.restype = None
    lib.sdm_last_edge_drawing_ms.argtypes = [vp, fp, fp, fp]
    lib.sdm_set_edge_drawing_route.argtypes = [vp, C.c_int]
    lib.sdm_last_edge_drawing_fallbacks.argtypes = [vp]
    lib.sdm_ed_device_edge_plane.argtypes = [vp, C.c_int, C.POINTER(vp)]
    lib.sdm_ed_planes.argtypes = [vp, vp, sz, C.c_int, C.c_int, vp, vp]
    lib.sdm_host_alloc.argtypes = [C.POINTER(vp), sz]
    lib.sdm_host_free.argtypes = [vp]
    lib.sdm_upload_keyframe.argtypes = [vp, C.c_int, vp, sz, vp, sz, vp, sz, vp, sz, fp, fp]
    lib.sdm_set_pose.argtypes = [vp, C.c_int, fp]
    lib.sdm_set_intrinsics.argtypes = [vp, C.c_int, fp]
    lib.sdm_candidate_count.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    lib.sdm_candidate_blocks.argtypes = [vp, C.c_int, ip, C.POINTER(C.c_uint64)]
    lib.sdm_pass1.argtypes = [vp, C.c_int, C.POINTER(Item)]
    lib.sdm_pass2.argtypes = [vp, C.c_int, C.POINTER(Item)]
    lib.sdm_update_points.argtypes = [vp, C.c_int, ip]
    lib.sdm_download.argtypes = [vp, C.c_int, vp, sz, vp, sz, vp, sz, vp, sz]
    lib.sdm_download_async.argtypes = [vp, C.c_int, vp, sz, vp, sz, vp, sz, vp, sz]
    lib.sdm_upload_keyframes.argtypes = [vp, C.c_int, C.POINTER(UploadDesc)]
    lib.sdm_download_keyframes.argtypes = [vp, C.c_int, C.POINTER(DownloadDesc)]
    lib.sdm_scatter_keyframes.argtypes = [vp, C.c_int, C.POINTER(DownloadDesc)]
    lib.sdm_inter_chi_test.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.sdm_download_planes.argtypes = [vp, C.c_int, vp, sz, vp, sz]
    lib.sdm_export_points.argtypes = [vp, C.c_int, ip, C.c_double, vp, sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.sdm_line_fit.argtypes = [vp, C.c_int, C.POINTER(EdgeChains), vp, sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.sdm_last_line_fit_ms.argtypes = [vp, fp]
    lib.sdm_last_scan_long.argtypes = [vp]
    lib.sdm_upload_depth.argtypes = [vp, C.c_int, vp, sz, vp, sz]
    lib.sdm_upload_checked.argtypes = [vp, C.c_int, vp, sz]
    lib.sdm_depth_plane_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    lib.sdm_export_arena.argtypes = [vp, vp, C.POINTER(sz)]
    lib.sdm_import_peer_arena.argtypes = [vp, C.c_int, vp]
    lib.sdm_pull_halo.argtypes = [vp, C.c_int, ip, ip, ip]
    lib.sdm_mark_pass1_done.argtypes = [vp, C.c_int]
    lib.sdm_export_peer_handle.argtypes = [vp, C.c_int, vp]
    lib.sdm_import_peer.argtypes = [vp, C.c_int, vp]
    lib.sdm_set_halo.argtypes = [vp, C.c_int, ip, ip, ip]
    lib.sdm_exchange.argtypes = [vp]
    lib.sdm_run_loop.argtypes = [vp, C.POINTER(Loop)]
    lib.sdm_pair_geometry.argtypes = [fp, fp, fp, fp, C.POINTER(PairGeometry)]
    lib.sdm_stereo_search_constraints.argtypes = [fp, C.c_int, fp, fp]
    lib.sdm_search_range.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, fp, fp]
    lib.sdm_epipolar_search.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_float,
                                        C.POINTER(Hypothesis)]
    lib.sdm_epipolar_search_plane.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]
    lib.sdm_fuse.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.sdm_intra_check.argtypes = [vp, C.c_int]
    lib.sdm_intra_grow.argtypes = [vp, C.c_int]
    lib.sdm_inter_check.argtypes = [vp, C.POINTER(Item)]
    lib.sdm_last_pass_ms.argtypes = [vp, fp, fp]
    lib.sdm_launch_count.argtypes = [vp]
    lib.sdm_launch_count.restype = C.c_longlong
    lib.sdm_last_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.sdm_last_pack_ms.argtypes = [vp, fp]
    lib.sdm_mark.argtypes = [vp, C.c_int]
    lib.sdm_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, fp]
    _lib = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_config(**over) -> Config:
    cfg = Config()
    load().sdm_default_config(C.byref(cfg))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def make_items(kfs, nbr_idx, rot, min_depth, max_depth, slot_of=None):
    """Item array for keyframes `kfs` (scene indices); slot_of maps scene index -> device slot."""
    kfs = list(kfs)
    arr = (Item * len(kfs))()
    f = (lambda i: int(i)) if slot_of is None else (lambda i: int(slot_of[int(i)]))
    for a, i in zip(arr, kfs):
        nb = nbr_idx[i]
        a.kf = f(i)
        a.n_nbr = len(nb)
        for j, v in enumerate(nb):
            a.nbr[j] = f(v)
            a.rot_deg[j] = float(rot[i][j])
        a.min_depth = float(min_depth[i])
        a.max_depth = float(max_depth[i])
    return arr


class Context:
    """One device context (sdm_ctx).  Methods map 1:1 onto the C entry points."""

    def __init__(self, cfg: Config | None = None, **over):
        self.lib = load()
        self.cfg = cfg if cfg is not None else default_config(**over)
        self.h = C.c_void_p()
        self._chk(self.lib.sdm_create(C.byref(self.cfg), C.byref(self.h)))
        self.W, self.H = self.cfg.width, self.cfg.height

    def _chk(self, rc: int):
        if rc != 0:
            raise SdmError(rc, (self.lib.sdm_last_error() or b"").decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.sdm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- planes
    def upload_keyframe(self, slot, im, grad, theta, edge, K, Tcw):
        """Arrays may be row-pitched views (positive strides, unit column stride)."""
        def chk(a, dt, esz):
            assert a.dtype == dt and a.shape == (self.H, self.W) and a.strides[1] == esz, (a.dtype, a.shape, a.strides)
        chk(im, np.uint8, 1)
        if grad is not None or theta is not None:
            chk(grad, np.float32, 4); chk(theta, np.float32, 4)
        if edge is not None:
            chk(edge, np.int32, 4)
        Kf, Tf = _f32(np.asarray(K).reshape(4)), _f32(np.asarray(Tcw).reshape(-1)[:12])
        self._chk(self.lib.sdm_upload_keyframe(
            self.h, slot, im.ctypes.data, im.strides[0],
            grad.ctypes.data if grad is not None else None, grad.strides[0] if grad is not None else 0,
            theta.ctypes.data if theta is not None else None, theta.strides[0] if theta is not None else 0,
            edge.ctypes.data if edge is not None else None, edge.strides[0] if edge is not None else 0,
            _fp(Kf), _fp(Tf)))

    def upload_descs(self, scene, indices, slot_of=None, images_only=False, edge_dev=None):
        """sdm_upload_desc array for keyframes `indices` of a scene (arrays must stay alive until synchronize).
        images_only: leave grad / theta NULL so that the planes are produced on the device.
        edge_dev: per entry the device address of its edge-index plane (ed_device_edge_plane) instead of scene.edge."""
        idx = list(indices)
        arr = (UploadDesc * len(idx))()
        for a, i in zip(arr, idx):
            a.kf = int(i if slot_of is None else slot_of[i])
            a.im, a.im_step = scene.im[i].ctypes.data, scene.im[i].strides[0]
            if not images_only:
                a.grad, a.grad_step = scene.grad[i].ctypes.data, scene.grad[i].strides[0]
                a.theta, a.theta_step = scene.theta[i].ctypes.data, scene.theta[i].strides[0]
            if edge_dev is not None:
                a.edge, a.edge_step = int(edge_dev[idx.index(i)]), 4 * self.W
            elif scene.edge is not None:
                a.edge, a.edge_step = scene.edge[i].ctypes.data, scene.edge[i].strides[0]
            for j, v in enumerate(scene.K):
                a.K[j] = float(v)
            for j, v in enumerate(np.asarray(scene.Tcw[i], np.float32).reshape(-1)[:12]):
                a.Tcw[j] = float(v)
        return arr

    def upload_keyframes(self, descs):
        self._chk(self.lib.sdm_upload_keyframes(self.h, len(descs), descs))

    def download_keyframes(self, descs):
        """enqueue only: the host arrays are valid after synchronize()"""
        self._chk(self.lib.sdm_download_keyframes(self.h, len(descs), descs))

    def scatter_keyframes(self, descs):
        """sparse form of download_keyframes for zero-initialised planes: enqueue only, valid after synchronize()"""
        self._chk(self.lib.sdm_scatter_keyframes(self.h, len(descs), descs))

    def scatter_all(self, slots, dtype_zero=True):
        """zero-initialised planes (as KeyFrame.cc:78-81) filled through sdm_scatter_keyframes; returns dict of stacks"""
        n = len(slots)
        out = {k: np.zeros((n, self.H, self.W) + ((3,) if k == "points" else ()), np.float32)
               for k in ("depth", "sigma", "checked", "points")}
        descs = (DownloadDesc * n)()
        for j, s in enumerate(slots):
            descs[j].kf = int(s)
            descs[j].depth, descs[j].depth_step = out["depth"][j].ctypes.data, 4 * self.W
            descs[j].sigma, descs[j].sigma_step = out["sigma"][j].ctypes.data, 4 * self.W
            descs[j].checked, descs[j].checked_step = out["checked"][j].ctypes.data, 4 * self.W
            descs[j].points, descs[j].points_step = out["points"][j].ctypes.data, 12 * self.W
        self.scatter_keyframes(descs)
        self.synchronize()
        return out

    def upload_scene(self, scene, indices=None, slot_of=None):
        idx = range(scene.n) if indices is None else indices
        self.upload_keyframes(self.upload_descs(scene, idx, slot_of))

    def set_pose(self, slot, Tcw):
        self._chk(self.lib.sdm_set_pose(self.h, slot, _fp(_f32(np.asarray(Tcw).reshape(-1)[:12]))))

    def set_intrinsics(self, slot, K):
        self._chk(self.lib.sdm_set_intrinsics(self.h, slot, _fp(_f32(np.asarray(K).reshape(4)))))

    def candidate_count(self, slot) -> int:
        n = C.c_int()
        self._chk(self.lib.sdm_candidate_count(self.h, slot, C.byref(n)))
        return n.value

    def candidate_blocks(self, slots) -> int:
        a = np.ascontiguousarray(slots, np.int32)
        n = C.c_uint64()
        self._chk(self.lib.sdm_candidate_blocks(self.h, a.size, a.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)))
        return int(n.value)

    # ---- passes
    def pass1(self, items):
        self._chk(self.lib.sdm_pass1(self.h, len(items), items))

    def pass2(self, items):
        self._chk(self.lib.sdm_pass2(self.h, len(items), items))

    def update_points(self, slots):
        a = np.ascontiguousarray(slots, np.int32)
        self._chk(self.lib.sdm_update_points(self.h, a.size, a.ctypes.data_as(C.POINTER(C.c_int32))))

    def synchronize(self):
        self._chk(self.lib.sdm_synchronize(self.h))

    def download(self, slot, depth=True, sigma=True, checked=True, points=True, out=None, async_=False):
        """async_=True only enqueues the copies: the arrays are valid after synchronize()."""
        H, W = self.H, self.W
        out = out or {}
        res = {}

        def buf(name, want, shape):
            if not want:
                return None
            a = out.get(name)
            if a is None:
                a = np.empty(shape, np.float32)
            res[name] = a
            return a
        d, s = buf("depth", depth, (H, W)), buf("sigma", sigma, (H, W))
        c, p = buf("checked", checked, (H, W)), buf("points", points, (H, W, 3))
        fn = self.lib.sdm_download_async if async_ else self.lib.sdm_download
        self._chk(fn(
            self.h, slot,
            d.ctypes.data if d is not None else None, d.strides[0] if d is not None else 0,
            s.ctypes.data if s is not None else None, s.strides[0] if s is not None else 0,
            c.ctypes.data if c is not None else None, c.strides[0] if c is not None else 0,
            p.ctypes.data if p is not None else None, p.strides[0] if p is not None else 0))
        return res

    def export_points(self, slots, sigma_max=0.02, capacity=None):
        """compacted point cloud of keyframes `slots`: structured array (x, y, z, pixel), per-keyframe counts"""
        a = np.ascontiguousarray(slots, np.int32)
        counts = np.zeros(a.size, np.uint64)
        total = C.c_uint64()
        ip, up = C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        if capacity is None:  # count first
            self._chk(self.lib.sdm_export_points(self.h, a.size, a.ctypes.data_as(ip), sigma_max, None, 0,
                                                 counts.ctypes.data_as(up), C.byref(total)))
            capacity = int(total.value)
        pts = np.zeros(max(capacity, 1), np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("pixel", "u4")]))
        self._chk(self.lib.sdm_export_points(self.h, a.size, a.ctypes.data_as(ip), sigma_max, pts.ctypes.data, capacity,
                                             counts.ctypes.data_as(up), C.byref(total)))
        return pts[:min(capacity, int(total.value))], counts, int(total.value)

    def line_fit(self, slots, offsets, pixels):
        """LineDetector::LineFitting for keyframes `slots`; offsets[i] (n_chains + 1 int32) / pixels[i] (uint32, row << 16 | col)
        are the edge chains of slots[i].  Returns (lines: structured array seg[4], xyz[6], chain, kf_index; per-set counts)"""
        n = len(slots)
        keep = [(np.ascontiguousarray(o, np.int32), np.ascontiguousarray(p, np.uint32)) for o, p in zip(offsets, pixels)]
        sets = (EdgeChains * max(n, 1))()
        cap = 0
        for i, (o, p) in enumerate(keep):
            sets[i].kf, sets[i].n_chains = int(slots[i]), o.size - 1
            sets[i].offsets = o.ctypes.data_as(C.POINTER(C.c_int32))
            sets[i].pixels = p.ctypes.data_as(C.POINTER(C.c_uint32))
            cap += int((np.diff(o) // 10).sum())
        out = np.zeros(max(cap, 1), LINE3D)
        counts = np.zeros(max(n, 1), np.uint64)
        total = C.c_uint64()
        self._chk(self.lib.sdm_line_fit(self.h, n, sets, out.ctypes.data, cap, counts.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        C.byref(total)))
        return out[:int(total.value)], counts[:n]

    def edge_drawing(self, images, grad_thresh=36, anchor_thresh=8, n_threads=0, edge_index=True, edge_out=None):
        """LineDetector::DetectEdgeMap for a batch of 8-bit images [n, H, W] (or a list of [H, W], rows may be pitched).
        Returns (offsets list, pixels list, edge_index [n, H, W] int32 or None): chains of image i = pixels[i][offsets[i][k] :
        offsets[i][k + 1]], packed (row << 16) | col - the layout line_fit takes; edge_index = kf->mEdgeIndex."""
        ims = [images[i] for i in range(len(images))]
        n = len(ims)
        for a in ims:
            assert a.dtype == np.uint8 and a.shape == (self.H, self.W) and a.strides[1] == 1
        edge = np.empty((n, self.H, self.W), np.int32) if edge_index else None
        if edge_out is not None:  # caller's planes (e.g. pinned memory)
            assert edge_out.dtype == np.int32 and edge_out.shape == (n, self.H, self.W) and edge_out.strides[2] == 4
            edge = edge_out
        descs = (EdImage * max(n, 1))()
        for i, a in enumerate(ims):
            descs[i].im, descs[i].im_step = a.ctypes.data, a.strides[0]
            if edge is not None:
                descs[i].edge_index, descs[i].edge_step = edge[i].ctypes.data, edge[i].strides[0]
        res = C.c_void_p()
        self._chk(self.lib.sdm_edge_drawing(self.h, n, descs, grad_thresh, anchor_thresh, n_threads, C.byref(res)))
        offs, pix = [], []
        try:
            for i in range(n):
                nc, po, pp = C.c_int32(), C.POINTER(C.c_int32)(), C.POINTER(C.c_uint32)()
                self._chk(self.lib.sdm_ed_chains(res, i, C.byref(nc), C.byref(po), C.byref(pp)))
                o = np.ctypeslib.as_array(po, shape=(nc.value + 1,)).copy()
                offs.append(o)
                pix.append(np.ctypeslib.as_array(pp, shape=(int(o[-1]),)).copy() if o[-1] > 0 else np.zeros(0, np.uint32))
        finally:
            self.lib.sdm_ed_free(res)
        return offs, pix, edge

    def set_edge_drawing_route(self, device):
        """stage 2 of edge_drawing on host threads (False / 0, default), on the device, one warp per image (True / 1), or on
        host threads with the masks built on the device from the chains (2)"""
        self._chk(self.lib.sdm_set_edge_drawing_route(self.h, int(device)))

    def ed_device_edge_plane(self, i: int) -> int:
        """device address of kf->mEdgeIndex of image i of the last device-routed edge_drawing batch (an `edge` plane for uploads)"""
        p = C.c_void_p()
        self._chk(self.lib.sdm_ed_device_edge_plane(self.h, int(i), C.byref(p)))
        return int(p.value)

    def last_edge_drawing_fallbacks(self) -> int:
        return int(self.lib.sdm_last_edge_drawing_fallbacks(self.h))

    def last_edge_drawing_ms(self) -> dict:
        k, w, r = C.c_float(), C.c_float(), C.c_float()
        self._chk(self.lib.sdm_last_edge_drawing_ms(self.h, C.byref(k), C.byref(w), C.byref(r)))
        return dict(kernel_ms=float(k.value), wall_ms=float(w.value), route_thread_ms=float(r.value))

    def ed_planes(self, im, grad_thresh=36, anchor_thresh=8):
        """stage-1 planes of one image as k_ed_planes computes them: (G int16, F uint8)"""
        assert im.dtype == np.uint8 and im.shape == (self.H, self.W) and im.strides[1] == 1
        G, F = np.empty((self.H, self.W), np.int16), np.empty((self.H, self.W), np.uint8)
        self._chk(self.lib.sdm_ed_planes(self.h, im.ctypes.data, im.strides[0], grad_thresh, anchor_thresh, G.ctypes.data, F.ctypes.data))
        return G, F

    def last_line_fit_ms(self) -> float:
        ms = C.c_float()
        self._chk(self.lib.sdm_last_line_fit_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def download_planes(self, slot):
        g, t = np.empty((self.H, self.W), np.float32), np.empty((self.H, self.W), np.float32)
        self._chk(self.lib.sdm_download_planes(self.h, slot, g.ctypes.data, g.strides[0], t.ctypes.data, t.strides[0]))
        return g, t

    def upload_depth(self, slot, depth, sigma):
        d, s = _f32(depth), _f32(sigma)
        self._chk(self.lib.sdm_upload_depth(self.h, slot, d.ctypes.data, d.strides[0], s.ctypes.data, s.strides[0]))

    def upload_checked(self, slot, checked):
        c = _f32(checked)
        self._chk(self.lib.sdm_upload_checked(self.h, slot, c.ctypes.data, c.strides[0]))

    def inter_chi_test(self, diff, sigma):
        d, s = _f32(np.ravel(diff)), _f32(np.ravel(sigma))
        out = np.empty(d.size, np.uint8)
        self._chk(self.lib.sdm_inter_chi_test(self.h, d.size, d.ctypes.data, s.ctypes.data, out.ctypes.data))
        return out.astype(bool)

    def scan_generation(self) -> int:
        return int(self.lib.sdm_scan_generation(self.h))

    def last_scan_long(self) -> bool:
        return bool(self.lib.sdm_last_scan_long(self.h))

    def stats(self) -> dict:
        st = Stats()
        self._chk(self.lib.sdm_get_stats(self.h, C.byref(st)))
        return {"candidates": st.candidates, "fused": st.fused, "checked": st.checked}

    def last_pass_ms(self):
        a, b = C.c_float(), C.c_float()
        self._chk(self.lib.sdm_last_pass_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_timing(self) -> dict:
        t = Timing()
        self._chk(self.lib.sdm_last_timing(self.h, C.byref(t)))
        return {"pass1_scan_ms": t.pass1_scan_ms, "pass1_intra_ms": t.pass1_intra_ms, "pass2_ms": t.pass2_ms}

    def last_pack_ms(self) -> float:
        ms = C.c_float()
        self._chk(self.lib.sdm_last_pack_ms(self.h, C.byref(ms)))
        return ms.value

    def mark(self, idx: int):
        self._chk(self.lib.sdm_mark(self.h, idx))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._chk(self.lib.sdm_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(self.lib.sdm_launch_count(self.h))

    # ---- per-method entry points
    def search_range(self, kf1, kf2, px, py, mind, maxd):
        a, b = C.c_float(), C.c_float()
        self._chk(self.lib.sdm_search_range(self.h, kf1, kf2, px, py, mind, maxd, C.byref(a), C.byref(b)))
        return a.value, b.value

    def epipolar_search(self, kf1, kf2, x, y, pixel, mind, maxd, th_pi, rot=0.0) -> Hypothesis:
        h = Hypothesis()
        self._chk(self.lib.sdm_epipolar_search(self.h, kf1, kf2, x, y, pixel, mind, maxd, th_pi, rot, C.byref(h)))
        return h

    def epipolar_search_plane(self, kf1, kf2, mind, maxd, rot=0.0):
        H, W = self.H, self.W
        d, s, u = (np.empty((H, W), np.float32) for _ in range(3))
        ok = np.empty((H, W), np.uint8)
        self._chk(self.lib.sdm_epipolar_search_plane(self.h, kf1, kf2, mind, maxd, rot, d.ctypes.data, s.ctypes.data,
                                                     u.ctypes.data, ok.ctypes.data))
        return d, s, u, ok

    def fuse(self, depth, sigma, count):
        depth, sigma = _f32(depth), _f32(sigma)
        m, n = depth.shape
        count = np.ascontiguousarray(count, np.int32)
        od, os_ = np.empty(m, np.float32), np.empty(m, np.float32)
        ok = np.empty(m, np.int32)
        self._chk(self.lib.sdm_fuse(self.h, m, n, depth.ctypes.data, sigma.ctypes.data, count.ctypes.data,
                                    od.ctypes.data, os_.ctypes.data, ok.ctypes.data))
        return od, os_, ok

    def intra_check(self, slot):
        self._chk(self.lib.sdm_intra_check(self.h, slot))

    def intra_grow(self, slot):
        self._chk(self.lib.sdm_intra_grow(self.h, slot))

    # ---- multi-GPU plumbing
    def depth_plane_ptr(self, slot):
        p, n = C.c_void_p(), C.c_size_t()
        self._chk(self.lib.sdm_depth_plane_ptr(self.h, slot, C.byref(p), C.byref(n)))
        return p.value, n.value

    def export_arena(self) -> bytes:
        buf = C.create_string_buffer(64)
        n = C.c_size_t()
        self._chk(self.lib.sdm_export_arena(self.h, buf, C.byref(n)))
        return bytes(buf.raw)

    def import_peer_arena(self, rank: int, handle: bytes):
        self._chk(self.lib.sdm_import_peer_arena(self.h, rank, C.create_string_buffer(handle, 64)))

    def pull_halo(self, local_slot, peer_rank, peer_slot):
        a, b, c = (np.ascontiguousarray(v, np.int32) for v in (local_slot, peer_rank, peer_slot))
        ip = C.POINTER(C.c_int32)
        self._chk(self.lib.sdm_pull_halo(self.h, a.size, a.ctypes.data_as(ip), b.ctypes.data_as(ip), c.ctypes.data_as(ip)))

    def mark_pass1_done(self, slot):
        self._chk(self.lib.sdm_mark_pass1_done(self.h, slot))

    # exchange ordered on the devices (no host barrier per step)
    def export_peer_handle(self, my_rank: int) -> bytes:
        buf = C.create_string_buffer(SDM_PEER_HANDLE_BYTES)
        self._chk(self.lib.sdm_export_peer_handle(self.h, my_rank, buf))
        return bytes(buf.raw)

    def import_peer(self, rank: int, handle: bytes):
        self._chk(self.lib.sdm_import_peer(self.h, rank, C.create_string_buffer(handle, SDM_PEER_HANDLE_BYTES)))

    def set_halo(self, local_slot, peer_rank, peer_slot):
        a, b, c = (np.ascontiguousarray(v, np.int32) for v in (local_slot, peer_rank, peer_slot))
        ip = C.POINTER(C.c_int32)
        self._chk(self.lib.sdm_set_halo(self.h, a.size, a.ctypes.data_as(ip), b.ctypes.data_as(ip), c.ctypes.data_as(ip)))

    def exchange(self):
        self._chk(self.lib.sdm_exchange(self.h))

    def run_loop(self, upload=None, pass1=None, down1=None, pass2=None, down2=None, chunk=0, exchange=False, sparse=False):
        """sdm_run_loop: ctypes arrays (UploadDesc / Item / DownloadDesc) or None; enqueue only"""
        L = Loop()
        L.n_upload = len(upload) if upload is not None else 0
        L.upload = upload if upload is not None else None
        L.n_pass1 = len(pass1) if pass1 is not None else 0
        L.pass1 = pass1 if pass1 is not None else None
        L.down1 = down1 if down1 is not None else None
        L.n_pass2 = len(pass2) if pass2 is not None else 0
        L.pass2 = pass2 if pass2 is not None else None
        L.down2 = down2 if down2 is not None else None
        L.chunk, L.exchange, L.sparse_download = int(chunk), int(bool(exchange)), int(sparse)  # sparse: 0 | 1 (True) | 2
        self._keep_loop = (upload, pass1, down1, pass2, down2)
        self._chk(self.lib.sdm_run_loop(self.h, C.byref(L)))


def pair_geometry(K1, Tcw1, K2, Tcw2) -> PairGeometry:
    g = PairGeometry()
    a, b, c, d = (_f32(np.asarray(v).reshape(-1)) for v in (K1, Tcw1, K2, Tcw2))
    rc = load().sdm_pair_geometry(_fp(a), _fp(b[:12].copy()), _fp(c), _fp(d[:12].copy()), C.byref(g))
    if rc:
        raise SdmError(rc, load().sdm_last_error().decode())
    return g


def stereo_search_constraints(inv_depths):
    a = _f32(inv_depths)
    lo, hi = C.c_float(), C.c_float()
    rc = load().sdm_stereo_search_constraints(_fp(a), a.size, C.byref(lo), C.byref(hi))
    if rc:
        raise SdmError(rc, load().sdm_last_error().decode())
    return lo.value, hi.value
