"""Generate tests/golden/cv2_kats.npz: known answers of the OpenCV primitives the oracle restates.

Run here (cv2 4.13 present): python oracle/pin_cv2.py
Inputs are random but shaped like the path's operands (rotations, TUM intrinsics, skew matrices);
outputs come from the real cv2 functions, so the oracle's restatement (oracle/sdm_oracle.c, ocv_*)
and the product's host geometry (eao-slam_b200/csrc/pair_geometry.h) can be checked bit-exactly
against OpenCV without OpenCV being present at test time.
"""
import os
import sys

import cv2
import numpy as np

rng = np.random.default_rng(20261018)
f32 = np.float32


def rand_rot(n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1)
    return R.reshape(n, 3, 3).astype(f32)


def main(out):
    n = 2000
    kat = {}
    # fastAtan2
    y = np.concatenate([rng.normal(size=20000) * 10, rng.normal(size=5000) * 1e-3, [0, 0, 1, -1, 0.0]]).astype(f32)
    x = np.concatenate([rng.normal(size=20000) * 10, np.ones(5000), [1, -1, 0, 0, 0.0]]).astype(f32)
    kat["atan_y"], kat["atan_x"] = y, x
    kat["atan_out"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], f32)
    # A * B^T (GEMM_2_T), alpha = +-1
    A, B = rand_rot(n), rand_rot(n)
    kat["abt_A"], kat["abt_B"] = A, B
    kat["abt_pos"] = np.stack([cv2.gemm(a, b, 1.0, None, 0.0, flags=cv2.GEMM_2_T) for a, b in zip(A, B)])
    kat["abt_neg"] = np.stack([cv2.gemm(a, b, -1.0, None, 0.0, flags=cv2.GEMM_2_T) for a, b in zip(A, B)])
    # 3x3 * 3x3
    A2 = (rng.normal(size=(n, 3, 3)) * rng.choice([1e-3, 1.0, 500.0], size=(n, 1, 1))).astype(f32)
    B2 = (rng.normal(size=(n, 3, 3))).astype(f32)
    kat["mm_A"], kat["mm_B"] = A2, B2
    kat["mm_out"] = np.stack([cv2.gemm(a, b, 1.0, None, 0.0) for a, b in zip(A2, B2)])
    # 3x3 * 3x1 * alpha + c
    xv = rng.normal(size=(n, 3, 1)).astype(f32)
    cv = rng.normal(size=(n, 3, 1)).astype(f32)
    al = np.abs(rng.normal(size=n)).astype(f32).astype(np.float64) * 3
    kat["mv_A"], kat["mv_x"], kat["mv_c"], kat["mv_alpha"] = A, xv, cv, al
    kat["mv_out"] = np.stack([cv2.gemm(a, x_, float(s), c_, 1.0) for a, x_, s, c_ in zip(A, xv, al, cv)])
    kat["mv_out_noc"] = np.stack([cv2.gemm(a, x_, float(s), None, 0.0) for a, x_, s in zip(A, xv, al)])
    # 1x3 * 3x1 * alpha
    rv = rng.normal(size=(n, 1, 3)).astype(f32)
    kat["dot_a"] = rv
    kat["dot_out"] = np.array([cv2.gemm(r, x_, float(s), None, 0.0)[0, 0] for r, x_, s in zip(rv, xv, al)], f32)
    # 1xn * nx1 via GEMM_1_T (J^T r), n in 1..28
    lens = rng.integers(1, 29, size=n)
    Jp = np.zeros((n, 28), f32)
    rp = np.zeros((n, 28), f32)
    outn = np.zeros(n, f32)
    outp = np.zeros(n, f32)
    for i, L in enumerate(lens):
        J = (rng.normal(size=(L, 1)) * 50).astype(f32)
        r = (rng.normal(size=(L, 1)) * 5).astype(f32)
        Jp[i, :L], rp[i, :L] = J[:, 0], r[:, 0]
        outn[i] = cv2.gemm(J, r, -1.0, None, 0.0, flags=cv2.GEMM_1_T)[0, 0]
        outp[i] = cv2.gemm(J, J, 1.0, None, 0.0, flags=cv2.GEMM_1_T)[0, 0]
    kat["jn_len"], kat["jn_J"], kat["jn_r"], kat["jn_neg"], kat["jn_jtj"] = lens.astype(np.int32), Jp, rp, outn, outp
    # invert 3x3 (K-like and random)
    Ks = np.zeros((n, 3, 3), f32)
    Ks[:, 0, 0] = 535.4 * rng.uniform(0.5, 2.5, n)
    Ks[:, 1, 1] = 539.2 * rng.uniform(0.5, 2.5, n)
    Ks[:, 0, 2] = 320.1 * rng.uniform(0.5, 2.5, n)
    Ks[:, 1, 2] = 247.6 * rng.uniform(0.5, 2.5, n)
    Ks[:, 2, 2] = 1
    Ks[n // 2:] = A2[n // 2:]
    kat["inv_A"] = Ks
    kat["inv_out"] = np.stack([cv2.invert(k, flags=cv2.DECOMP_LU)[1] for k in Ks])
    # solve 3x3 with 3x3 rhs (K^T, skew) and random systems
    Kt = np.transpose(Ks, (0, 2, 1)).copy()
    tv = rng.normal(size=(n, 3)).astype(f32)
    Sk = np.zeros((n, 3, 3), f32)
    Sk[:, 0, 1], Sk[:, 0, 2] = -tv[:, 2], tv[:, 1]
    Sk[:, 1, 0], Sk[:, 1, 2] = tv[:, 2], -tv[:, 0]
    Sk[:, 2, 0], Sk[:, 2, 1] = -tv[:, 1], tv[:, 0]
    Sk[n // 2:] = B2[n // 2:]
    kat["solve_A"], kat["solve_B"] = Kt, Sk
    kat["solve_out"] = np.stack([cv2.solve(a, b, flags=cv2.DECOMP_LU)[1] for a, b in zip(Kt, Sk)])
    # 4x4 * 4x1
    T4 = rng.normal(size=(n, 4, 4)).astype(f32)
    p4 = rng.normal(size=(n, 4, 1)).astype(f32)
    kat["m4_A"], kat["m4_x"] = T4, p4
    kat["m4_out"] = np.stack([cv2.gemm(a, b, 1.0, None, 0.0) for a, b in zip(T4, p4)])
    # Scharr / 32 on a random u8 image (exactness of the plane producer)
    img = rng.integers(0, 256, size=(24, 32), dtype=np.uint8)
    kat["sch_img"] = img
    kat["sch_gx"] = cv2.Scharr(img, cv2.CV_32F, 1, 0, scale=1 / 32.0)
    kat["sch_gy"] = cv2.Scharr(img, cv2.CV_32F, 0, 1, scale=1 / 32.0)
    kat["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(out, **kat)
    print("wrote", out, {k: v.shape for k, v in kat.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "tests", "golden", "cv2_kats.npz"))
