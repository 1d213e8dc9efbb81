"""CPU oracle of the edge-aided 3-D line fitting - TEST INFRASTRUCTURE ONLY.

A restatement of LineDetector::LineFit and its helpers (/root/reference/src/LineDetector.cc:578-840; constants :20-29)
on caller-supplied pixel chains, with every OpenCV call the reference makes made through the REAL cv2 of this image:
    cv::SVD::solveZ(A, u)              -> cv2.SVDecomp(A) and the last row of vt     (:617; modules/core/src/lapack.cpp)
    cv::solve(A, b, u, DECOMP_SVD)     -> cv2.solve(A, b, flags=cv2.DECOMP_SVD)      (:671)
    cv::norm(A*u), cv::norm(A*u, b)    -> cv2.norm(cv2.gemm(...)), cv2.norm(.., b)  (:623, :676)
    Twc * Pc                           -> float 4-term dot, left to right (the rule UpdateSemiDensePointSet uses)
so the decisions of the oracle (which windows start a line, where a line stops) are OpenCV's own.  Scalar arithmetic is
float32 where the reference's is float, double for the angle test (:809-814).  The edge chains themselves come from the
closed-source EDLib in the reference (LineDetector.cc:855: DetectEdgesByED); here they are an input.

Only tests/ may import this module."""
from __future__ import annotations

import cv2
import numpy as np

MIN_LINE_LENGTH, MAX_LINE_LENGTH, INIT_DEPTH_COUNT = 10, 1000, 3   # LineDetector.cc:20-22
MIN_SEGMENT_ANGLE, E1, E2, SIGMA_LIMIT = np.float32(30), np.float32(1.0), np.float32(1.5), np.float32(0.02)  # :23-25, :29
f32 = np.float32


def closest_point_on_line(a, b, c, x, y):
    """:578-582, float arithmetic (x, y are ints converted in the products)"""
    a, b, c = f32(a), f32(b), f32(c)
    den = a * a + b * b
    cx = (b * (b * f32(x) - a * f32(y)) - a * c) / den
    cy = (a * (-b * f32(x) + a * f32(y)) - b * c) / den
    return f32(cx), f32(cy)


def norm2(px, py, qx, qy):
    """cv::norm(Point2f): sqrt((double)x*x + (double)y*y) -> cast to float by the caller"""
    dx, dy = f32(px - qx), f32(py - qy)
    return f32(np.sqrt(float(dx) * float(dx) + float(dy) * float(dy)))


class Planes:
    def __init__(self, checked, sigma, K, Twc):
        self.checked, self.sigma = checked, sigma
        self.fx, self.fy, self.cx, self.cy = (f32(v) for v in K)
        self.Twc = np.asarray(Twc, np.float32).reshape(3, 4)


def has_depth(P, r, c):
    return P.checked[r, c] > 0.000001 and P.sigma[r, c] < SIGMA_LIMIT   # float vs double / float literals as in :593-594


def count_depth(P, chain, i0, length):
    return sum(1 for i in range(i0, i0 + length) if has_depth(P, chain[i][0], chain[i][1]))


def ls_line_fit(chain, i0, length):
    A = np.zeros((length, 3), np.float32)
    for i in range(length):
        A[i] = (chain[i0 + i][1], chain[i0 + i][0], 1)
    _, _, vt = cv2.SVDecomp(A)
    u = vt[-1].astype(np.float32)
    err = f32(cv2.norm(cv2.gemm(A, u.reshape(3, 1), 1, None, 0)))
    return u[0], u[1], u[2], err


def ls_depth_fit(P, chain, i0, length, la, lb, lc):
    A = np.zeros((length, 2), np.float32)
    b = np.zeros((length, 1), np.float32)
    sx, sy = closest_point_on_line(la, lb, lc, chain[i0][1], chain[i0][0])
    for i in range(length):
        r, c = chain[i0 + i]
        invz, sg = P.checked[r, c], P.sigma[r, c]
        if invz > 0.000001 and sg < SIGMA_LIMIT:
            cx, cy = closest_point_on_line(la, lb, lc, c, r)
            d = norm2(cx, cy, sx, sy)
            A[i] = (d, 1)
            b[i, 0] = f32(f32(1) * f32(1) / invz) * (f32(P.fx + P.fy) / f32(2))
    _, u = cv2.solve(A, b, flags=cv2.DECOMP_SVD)
    err = f32(cv2.norm(cv2.gemm(A, u, 1, None, 0), b))
    return f32(u[0, 0]), f32(u[1, 0]), err


def point_distance_to_line(a, b, c, r, col):
    return f32(abs(f32(a * f32(col) + b * f32(r)) + c) / np.sqrt(f32(a * a + b * b)))


def point_depth_to_line(P, chain, i0, a, b, c, alpha, beta, r, col):
    with np.errstate(divide="ignore"):
        z = f32(1) / P.checked[r, col]
    if z < 0.000001:
        return f32(-1)
    if P.sigma[r, col] > SIGMA_LIMIT:
        return f32(-1)
    sx, sy = closest_point_on_line(a, b, c, chain[i0][1], chain[i0][0])
    cx, cy = closest_point_on_line(a, b, c, col, r)
    t = norm2(cx, cy, sx, sy)
    z = z * (f32(P.fx + P.fy) / f32(2))
    return f32(abs(f32(alpha * t - z) + beta) / np.sqrt(f32(alpha * alpha + f32(1))))


def line_fit(P, chain, out, chain_id=0):
    """LineFit (:713-840) with its tail recursion as a loop.  chain: list of (r, c).  Appends (chain_id, line2D[4], line3D[6])."""
    inf = f32(np.inf)
    i0, n = 0, len(chain)
    while True:
        err_l, err_d = inf, inf
        a = b = c = alpha = beta = f32(0)
        init = MIN_LINE_LENGTH
        while n > init and init < MAX_LINE_LENGTH:
            if count_depth(P, chain, i0, 1) < 1 or count_depth(P, chain, i0, init) < INIT_DEPTH_COUNT:
                i0 += 1; n -= 1
                continue
            a, b, c, err_l = ls_line_fit(chain, i0, init)
            alpha, beta, err_d = ls_depth_fit(P, chain, i0, init, a, b, c)
            if err_l <= 1.0 and err_d <= 1.0:
                break
            i0 += 1; n -= 1
        if err_l > E1 or err_d > E2:
            return
        interval, length = 0, init
        while length < MAX_LINE_LENGTH and length < n:
            r, col = chain[i0 + length]
            if point_distance_to_line(a, b, c, r, col) > E1:
                break
            length += 1
            interval += 1
            if interval >= MIN_LINE_LENGTH:
                interval = 0
                if count_depth(P, chain, i0 + length - MIN_LINE_LENGTH, MIN_LINE_LENGTH) < 1:
                    length -= MIN_LINE_LENGTH
                    break
                prev, stop = length - MIN_LINE_LENGTH, False
                for i in range(prev, length):
                    r2, c2 = chain[i0 + i]
                    dept = point_depth_to_line(P, chain, i0, a, b, c, alpha, beta, r2, c2)
                    if dept > E2:
                        length, stop = prev, True
                        break
                    if dept >= 0.0:
                        prev = i
                if stop:
                    break
        a, b, c, err_l = ls_line_fit(chain, i0, length)
        if f32(count_depth(P, chain, i0, length)) / f32(length) > f32(INIT_DEPTH_COUNT) / f32(MIN_LINE_LENGTH):
            alpha, beta, err_d = ls_depth_fit(P, chain, i0, length, a, b, c)
            sx, sy = closest_point_on_line(a, b, c, chain[i0][1], chain[i0][0])
            ex, ey = closest_point_on_line(a, b, c, chain[i0 + length - 1][1], chain[i0 + length - 1][0])
            half = f32(P.fx + P.fy) / f32(2)
            Zs = f32(beta) / half
            Xs, Ys = f32(Zs * (sx - P.cx)) / P.fx, f32(Zs * (sy - P.cy)) / P.fy
            Ze = f32(f32(alpha * norm2(ex, ey, sx, sy)) + beta) / half
            Xe, Ye = f32(Ze * (ex - P.cx)) / P.fx, f32(Ze * (ey - P.cy)) / P.fy
            ps, pe = np.array([Xs, Ys, Zs], np.float32), np.array([Xe, Ye, Ze], np.float32)
            diff = (pe - ps).astype(np.float32)
            with np.errstate(invalid="ignore", divide="ignore"):
                nd = np.sqrt(float(np.sum(diff.astype(np.float64) ** 2)))
                cos_s = float(np.dot(diff.astype(np.float64), ps.astype(np.float64))) / nd / np.sqrt(float(np.sum(ps.astype(np.float64) ** 2)))
                cos_e = float(np.dot(diff.astype(np.float64), pe.astype(np.float64))) / nd / np.sqrt(float(np.sum(pe.astype(np.float64) ** 2)))
                ang_s = np.degrees(np.arccos(abs(cos_s)))
                ang_e = np.degrees(np.arccos(abs(cos_e)))
            if ang_s > MIN_SEGMENT_ANGLE and ang_e > MIN_SEGMENT_ANGLE:
                def world(p):
                    q = np.array([p[0], p[1], p[2], 1], np.float32)
                    return [f32(f32(f32(f32(P.Twc[k, 0] * q[0]) + f32(P.Twc[k, 1] * q[1])) + f32(P.Twc[k, 2] * q[2])) + f32(P.Twc[k, 3] * q[3]))
                            for k in range(3)]
                out.append((chain_id, [sx, sy, ex, ey], world(ps) + world(pe)))
        i0 += length
        n -= length


def line_fitting(P, chains):
    """LineFitting (:884-900): LineFit over every edge chain of the keyframe, in order"""
    out = []
    for k, ch in enumerate(chains):
        line_fit(P, [(int(r), int(c)) for r, c in ch], out, k)
    return out
