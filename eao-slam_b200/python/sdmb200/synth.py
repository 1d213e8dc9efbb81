"""Synthetic keyframe sets for the semi-dense mapping path.

An analytic ray-cast of a textured room (inside-out box) with a desk (box) in the middle, seen from
a camera circling the desk — "fr3_long_office-like".  Produces exactly the inputs the reference's
hot path consumes (KeyFrame.h:155-175): im_ (u8), GradImg / GradTheta (f32; the same cv2 calls as
KeyFrame.cc:71-74 when cv2 is importable, a numpy restatement otherwise), poses Tcw (f32),
covisible-neighbour lists, the in-plane rotation per pair (0: the trajectory is roll-free) and the
per-keyframe depth search bounds of StereoSearchConstraints (ProbabilityMapping.cc:734-747) from
500 sampled true depths.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

try:  # cv2 is in the image; keep a fallback so nothing here hard-depends on it
    import cv2  # type: ignore
except Exception:  # pragma: no cover
    cv2 = None

TUM3_K = (535.4, 539.2, 320.1, 247.6)  # Examples/Monocular/TUM3.yaml:8-16 (fx, fy, cx, cy)


@dataclass
class Scene:
    im: np.ndarray          # [n,H,W] uint8
    grad: np.ndarray        # [n,H,W] float32
    theta: np.ndarray       # [n,H,W] float32 degrees
    edge: np.ndarray | None  # [n,H,W] int32 or None (= all pass)
    K: tuple                # fx, fy, cx, cy (float)
    Tcw: np.ndarray         # [n,3,4] float32
    nbr_idx: np.ndarray     # [n,N] int32
    rot: np.ndarray         # [n,N] float32
    min_depth: np.ndarray   # [n] float32
    max_depth: np.ndarray   # [n] float32
    depth_gt: np.ndarray | None = None  # [n,H,W] float32 camera-z depth
    inv_depths: list | None = None      # per keyframe: sorted float32 inverse depths (GetAllPointDepths stand-in)
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return self.im.shape[0]

    @property
    def shape(self):
        return self.im.shape[1:]


# ----------------------------------------------------------------------------------------------
# plane producers (KeyFrame.cc:69-74)
# ----------------------------------------------------------------------------------------------

def _fast_atan2_deg(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    """numpy float32 restatement of cv::fastAtan2 (scalar form); used only without cv2."""
    f = np.float32
    scale = f(180.0 / np.pi)
    p1, p3 = f(0.9997878412794807) * scale, f(-0.3258083974640975) * scale
    p5, p7 = f(0.1555786518463281) * scale, f(-0.04432655554792128) * scale
    ax, ay = np.abs(x), np.abs(y)
    eps = f(2.220446049250313e-16)
    swap = ax < ay
    num = np.where(swap, ax, ay)
    den = np.where(swap, ay, ax) + eps
    c = (num / den).astype(f)
    c2 = c * c
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(swap, f(90.0) - a, a)
    a = np.where(x < 0, f(180.0) - a, a)
    a = np.where(y < 0, f(360.0) - a, a)
    return a.astype(f)


def gradient_planes(im: np.ndarray):
    """GradImg, GradTheta of one u8 image: Scharr(scale 1/32) -> magnitude / phase(degrees) (KeyFrame.cc:69-74).

    Scharr/32 is exact (integer sums times 2^-5), so cv2 and the numpy form agree bit for bit.  magnitude and
    phase are evaluated with numpy (float32 sqrt; the scalar form of cv::fastAtan2): cv2.magnitude was observed
    to return results that differ in the last bit BETWEEN CALLS on a 24-thread host (SIMD body vs scalar tail of
    OpenCV's parallel stripes), which would make two ranks disagree about the planes of a shared keyframe.  The
    planes are inputs of the path, so either producer is equally valid."""
    if cv2 is not None:
        gx = cv2.Scharr(im, cv2.CV_32F, 1, 0, scale=1 / 32.0)
        gy = cv2.Scharr(im, cv2.CV_32F, 0, 1, scale=1 / 32.0)
    else:
        p = np.pad(im.astype(np.int32), 1, mode="reflect")  # BORDER_REFLECT_101
        gx = (3 * (p[:-2, 2:] - p[:-2, :-2]) + 10 * (p[1:-1, 2:] - p[1:-1, :-2]) + 3 * (p[2:, 2:] - p[2:, :-2]))
        gy = (3 * (p[2:, :-2] - p[:-2, :-2]) + 10 * (p[2:, 1:-1] - p[:-2, 1:-1]) + 3 * (p[2:, 2:] - p[:-2, 2:]))
        gx = (gx / 32.0).astype(np.float32)
        gy = (gy / 32.0).astype(np.float32)
    return np.sqrt(gx * gx + gy * gy, dtype=np.float32), _fast_atan2_deg(gy, gx)


# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------

ROOM_LO = np.array([-3.0, -1.3, -3.0])
ROOM_HI = np.array([3.0, 1.3, 3.0])
DESK_LO = np.array([-0.8, 0.25, -0.5])
DESK_HI = np.array([0.8, 1.3, 0.5])


def trajectory(n: int, step_m: float = 0.05, radius: float = 1.6, phase0: float = 0.3, first: int = 0) -> np.ndarray:
    """Tcw [n,3,4] float32: camera circling the desk, looking at it, roll-free, y down.
    `first` is the global index of the first pose (a slice of one long trajectory)."""
    T = np.zeros((n, 3, 4), np.float64)
    dphi = step_m / radius
    for i in range(n):
        phi = phase0 + (first + i) * dphi
        r = radius + 0.2 * np.cos(2 * phi)
        Ow = np.array([r * np.cos(phi), -0.35 + 0.15 * np.sin(3 * phi), r * np.sin(phi)])
        target = np.array([0.1 * np.sin(phi), 0.35, 0.1 * np.cos(1.7 * phi)])
        z = target - Ow
        z /= np.linalg.norm(z)
        x = np.cross(np.array([0.0, 1.0, 0.0]), z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        Rwc = np.stack([x, y, z], axis=1)
        Rcw = Rwc.T
        T[i, :, :3] = Rcw
        T[i, :, 3] = -Rcw @ Ow
    return T.astype(np.float32)


def neighbours(n: int, N: int) -> np.ndarray:
    """[n,N] nearest-by-index neighbour lists, nearest first (covisibility order)."""
    assert n > N, "need more keyframes than neighbours"
    out = np.zeros((n, N), np.int32)
    for i in range(n):
        lo = i - N // 2
        hi = lo + N  # window of N+1 indices containing i
        if lo < 0:
            lo, hi = 0, N
        if hi > n - 1:
            lo, hi = n - 1 - N, n - 1
        idx = [j for j in range(lo, hi + 1) if j != i]
        idx.sort(key=lambda j: (abs(j - i), j))
        out[i] = idx[:N]
    return out


class _Texture:
    """Solid (3-D) texture: a sum of random plane waves, evaluated at surface points."""

    def __init__(self, rng: np.random.Generator, n_waves: int = 20, contrast: float = 1.0):
        lam = np.exp(rng.uniform(np.log(0.04), np.log(0.6), n_waves))  # wavelengths in metres
        d = rng.normal(size=(n_waves, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        self.k = (2 * np.pi / lam)[:, None] * d  # [n_waves,3]
        self.amp = lam ** 0.7
        self.phi = rng.uniform(0, 2 * np.pi, n_waves)
        self.norm = 1.0 / (np.sqrt(0.5 * np.sum(self.amp ** 2)) * 2.2)
        self.contrast = contrast

    def __call__(self, P: np.ndarray) -> np.ndarray:
        # P [...,3] -> intensity in [0,255] (float)
        ph = P.reshape(-1, 3).astype(np.float32) @ self.k.T.astype(np.float32) + self.phi.astype(np.float32)
        v = (np.cos(ph) * self.amp.astype(np.float32)).sum(axis=1) * np.float32(self.norm * self.contrast)
        v = 127.5 + 127.5 * np.clip(v, -1.0, 1.0)
        return v.reshape(P.shape[:-1])


def _render(Tcw: np.ndarray, K, W: int, H: int, tex: _Texture, box_tex: _Texture):
    fx, fy, cx, cy = K
    R = Tcw[:, :3].astype(np.float64)
    t = Tcw[:, 3].astype(np.float64)
    Ow = -R.T @ t
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dc = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)  # camera rays, z = 1
    dw = dc @ R  # = (R^T dc^T)^T -> world direction per pixel
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dw
        # room: inside-out box, exit distance
        t1 = (ROOM_LO - Ow) * inv
        t2 = (ROOM_HI - Ow) * inv
        t_room = np.min(np.maximum(t1, t2), axis=-1)
        # desk: outside-in box, entry distance
        b1 = (DESK_LO - Ow) * inv
        b2 = (DESK_HI - Ow) * inv
        t_near = np.max(np.minimum(b1, b2), axis=-1)
        t_far = np.min(np.maximum(b1, b2), axis=-1)
    hit = (t_near < t_far) & (t_near > 1e-3)
    depth = np.where(hit, t_near, t_room)
    P = Ow + depth[..., None] * dw
    val = np.where(hit, box_tex(P), tex(P))
    im = np.clip(np.rint(val), 0, 255).astype(np.uint8)
    return im, depth.astype(np.float32)


def stereo_search_constraints(inv_depths: np.ndarray):
    """numpy form of ProbabilityMapping::StereoSearchConstraints (:734-747); returns (min_depth, max_depth)."""
    d = np.sort(inv_depths.astype(np.float32))
    mean = np.float32(np.sum(d, dtype=np.float64)) / np.float32(d.size)
    diff = d - mean
    var = np.float32(np.sum((diff * diff).astype(np.float64)) / d.size)
    sd = np.sqrt(var, dtype=np.float32)
    max_depth = np.float32(1) / (mean + np.float32(2) * sd)
    min_depth = np.float32(1) / (mean - np.float32(2) * sd)
    return np.float32(min_depth), np.float32(max_depth)


_GEN = {}  # per-process generator state (textures, poses) for the keyframe worker


def _make_kf(i: int):
    g = _GEN
    W, H, K, seed, first = g["W"], g["H"], g["K"], g["seed"], g["first"]
    im, depth = _render(g["Tcw"][i], K, W, H, g["tex"], g["box_tex"])
    grad, theta = gradient_planes(im)
    pick = np.random.default_rng([seed, first + i]).integers(0, W * H, size=500)
    rho = 1.0 / depth.reshape(-1)[pick]
    if g["wide_range"]:  # config C4: force a wide search interval
        rho = np.concatenate([rho, [1 / 0.4, 1 / 8.0]]).astype(np.float32)
    mind, maxd = stereo_search_constraints(rho)
    if mind <= 0:  # mean - 2 sigma <= 0: the reference has no guard; keep the scan bounded
        mind = np.float32(1.0 / max(float(np.min(rho)) * 0.5, 1e-3))
    edge = None
    if g["edge_mask"]:  # a synthetic mEdgeIndex: segment id >= 0 on strong-gradient ridges, -1 elsewhere
        ridge = (grad > 10) & ((grad >= np.roll(grad, 1, 0)) & (grad >= np.roll(grad, -1, 0)) |
                               (grad >= np.roll(grad, 1, 1)) & (grad >= np.roll(grad, -1, 1)))
        edge = np.where(ridge, (np.arange(W * H, dtype=np.int32).reshape(H, W) // 97), -1).astype(np.int32)
    return i, im, grad, theta, (depth if g["keep_depth"] else None), mind, maxd, edge, np.sort(rho.astype(np.float32))


def make_scene(n_kf: int, W: int = 640, H: int = 480, n_nbr: int = 6, seed: int = 1,
               step_m: float = 0.05, K=None, keep_depth: bool = False, contrast: float = 0.6,
               wide_range: bool = False, edge_mask: bool = False, phase0: float = 0.3, first: int = 0,
               workers: int = 1, nbr_idx: np.ndarray | None = None) -> Scene:
    """Build a keyframe set.  K defaults to TUM fr3 intrinsics scaled by W/640.  `first` = global
    index of keyframe 0 on the trajectory: make_scene(n, first=f) equals keyframes [f, f+n) of one
    long trajectory through the same world (textures depend on `seed` only, the per-keyframe depth
    samples on (seed, global index)), which is how bench.py shards a trajectory over ranks.
    workers > 1 renders keyframes in forked worker processes (call before CUDA is initialised)."""
    rng = np.random.default_rng(seed)
    if K is None:
        s = W / 640.0
        K = tuple(float(np.float32(v * s)) for v in TUM3_K)
    tex, box_tex = _Texture(rng, contrast=contrast), _Texture(rng, contrast=contrast)
    Tcw = trajectory(n_kf, step_m=step_m, phase0=phase0, first=first)
    im = np.zeros((n_kf, H, W), np.uint8)
    grad = np.zeros((n_kf, H, W), np.float32)
    theta = np.zeros((n_kf, H, W), np.float32)
    dgt = np.zeros((n_kf, H, W), np.float32) if keep_depth else None
    mind = np.zeros(n_kf, np.float32)
    maxd = np.zeros(n_kf, np.float32)
    edge = np.zeros((n_kf, H, W), np.int32) if edge_mask else None
    _GEN.update(W=W, H=H, K=K, seed=seed, first=first, Tcw=Tcw, tex=tex, box_tex=box_tex,
                wide_range=wide_range, edge_mask=edge_mask, keep_depth=keep_depth)
    if workers > 1 and n_kf >= 2 * workers:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            results = pool.imap_unordered(_make_kf, range(n_kf), chunksize=2)
            results = list(results)
    else:
        results = (_make_kf(i) for i in range(n_kf))
    inv_depths = [None] * n_kf
    for i, a, g, t, d, lo, hi, e, rho in results:
        inv_depths[i] = rho
        im[i], grad[i], theta[i], mind[i], maxd[i] = a, g, t, lo, hi
        if keep_depth:
            dgt[i] = d
        if edge_mask:
            edge[i] = e
    _GEN.clear()
    if nbr_idx is None:
        nbr_idx = neighbours(n_kf, n_nbr)
    return Scene(im=im, grad=grad, theta=theta, edge=edge, K=K, Tcw=Tcw,
                 nbr_idx=nbr_idx, rot=np.zeros(nbr_idx.shape, np.float32),
                 min_depth=mind, max_depth=maxd, depth_gt=dgt, inv_depths=inv_depths,
                 meta={"seed": seed, "step_m": step_m, "first": first, "planes": "cv2" if cv2 is not None else "numpy"})
