"""Generate tests/golden/pair_geometry_cv2.npz: R21 / t21 / F12 of keyframe pairs computed with REAL
cv2 calls in the order OpenCV's MatExpr machinery evaluates the reference's expressions
(ProbabilityMapping.cc:1136-1137 `Rcw2*Rcw1.t()`, `-Rcw2*Rcw1.t()*tcw1+tcw2`; :1700-1708
`K1.t().inv()*t12x*R12*K2.inv()`): gemm(GEMM_2_T) / gemm / solve(DECOMP_LU) / invert(DECOMP_LU).
Run here (cv2 4.13 present): python oracle/pin_pair_geometry.py
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "eao-slam_b200", "python"))
from sdmb200 import synth  # noqa: E402

f32 = np.float32


def pair(K1, T1, K2, T2):
    R1, t1 = np.ascontiguousarray(T1[:, :3]), np.ascontiguousarray(T1[:, 3:4])
    R2, t2 = np.ascontiguousarray(T2[:, :3]), np.ascontiguousarray(T2[:, 3:4])
    R21 = cv2.gemm(R2, R1, 1.0, None, 0.0, flags=cv2.GEMM_2_T)
    t21 = cv2.gemm(cv2.gemm(R2, R1, -1.0, None, 0.0, flags=cv2.GEMM_2_T), t1, 1.0, t2, 1.0)
    R12 = cv2.gemm(R1, R2, 1.0, None, 0.0, flags=cv2.GEMM_2_T)
    t12 = cv2.gemm(cv2.gemm(R1, R2, -1.0, None, 0.0, flags=cv2.GEMM_2_T), t2, 1.0, t1, 1.0).reshape(3)
    t12x = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]], f32)
    Km1 = np.array([[K1[0], 0, K1[2]], [0, K1[1], K1[3]], [0, 0, 1]], f32)
    Km2 = np.array([[K2[0], 0, K2[2]], [0, K2[1], K2[3]], [0, 0, 1]], f32)
    ok, S = cv2.solve(np.ascontiguousarray(Km1.T), t12x, flags=cv2.DECOMP_LU)   # inv(A)*B -> solve
    SR = cv2.gemm(S, R12, 1.0, None, 0.0)
    K2i = cv2.invert(Km2, flags=cv2.DECOMP_LU)[1]
    F12 = cv2.gemm(SR, K2i, 1.0, None, 0.0)
    return R21, t21.reshape(3), F12


def main(out):
    rng = np.random.default_rng(7)
    K1s, T1s, K2s, T2s, Rs, ts, Fs = [], [], [], [], [], [], []
    traj = synth.trajectory(400, step_m=0.05)
    for _ in range(1500):
        i = int(rng.integers(0, 400)); j = int(np.clip(i + rng.integers(-5, 6), 0, 399))
        if i == j:
            j = (i + 1) % 400
        s = float(rng.choice([1.0, 2.0]))
        K = np.array([v * s for v in synth.TUM3_K], f32)
        K1s.append(K); K2s.append(K); T1s.append(traj[i]); T2s.append(traj[j])
    # random rigid poses and unequal intrinsics
    import importlib.util
    spec = importlib.util.spec_from_file_location("pin_cv2", os.path.join(os.path.dirname(os.path.abspath(__file__)), "pin_cv2.py"))
    pin = importlib.util.module_from_spec(spec); spec.loader.exec_module(pin)
    RA, RB = pin.rand_rot(500), pin.rand_rot(500)
    for a, b in zip(RA, RB):
        T1 = np.concatenate([a, rng.normal(size=(3, 1)).astype(f32)], axis=1)
        T2 = np.concatenate([b, rng.normal(size=(3, 1)).astype(f32)], axis=1)
        K1 = (np.array(synth.TUM3_K) * rng.uniform(0.5, 2.5, 4)).astype(f32)
        K2 = (np.array(synth.TUM3_K) * rng.uniform(0.5, 2.5, 4)).astype(f32)
        K1s.append(K1); K2s.append(K2); T1s.append(T1.astype(f32)); T2s.append(T2.astype(f32))
    for K1, T1, K2, T2 in zip(K1s, T1s, K2s, T2s):
        R, t, F = pair(K1, T1, K2, T2)
        Rs.append(R); ts.append(t); Fs.append(F)
    np.savez_compressed(out, K1=np.stack(K1s), T1=np.stack(T1s), K2=np.stack(K2s), T2=np.stack(T2s),
                        R21=np.stack(Rs), t21=np.stack(ts), F12=np.stack(Fs), cv2_version=np.array(cv2.__version__))
    print("wrote", out, len(Rs))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "tests", "golden", "pair_geometry_cv2.npz"))
