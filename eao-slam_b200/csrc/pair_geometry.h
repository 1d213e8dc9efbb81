// pair_geometry.h — host-side per-(keyframe, neighbour) geometry with OpenCV's evaluation rules.
//
// Replaces (reference: yanmin-wu/EAO-SLAM, src/ProbabilityMapping.cc):
//   R21 / t21           :1136-1137 (also recomputed per pixel at :1582-1583, :1611-1612)
//   ComputeFundamental  :1694-1709, GetSkewSymmetricMatrix :1711-1715
//   KeyFrame::SetPose   src/KeyFrame.cc:108-124 (Twc for UpdateSemiDensePointSet)
//   StereoSearchConstraints :734-747
//
// The reference evaluates these with cv::Mat expressions; results depend on where OpenCV uses
// float or double.  The rules (checked bit-exactly against real cv2 calls: tests/test_abi.py on tests/golden/pair_geometry_cv2.npz,
// primitives in tests/test_oracle_cv2_kats.py on tests/golden/cv2_kats.npz):
//   * A * B^T          : general gemm, double accumulation, (float)alpha, result cast to float
//   * 3x3 * 3x3        : float 3-term dot product, left to right
//   * 3x3 * 3x1 *a + c : float dot, then float(double(t) * a + double(c)), a rounded to float first
//   * inv(3x3)         : adjugate with double products
//   * inv(A) * B       : cv::solve(A, B, DECOMP_LU) = float LU with partial pivoting (no FMA)
// This file must be compiled without floating-point contraction (-ffp-contract=off / -fmad=false).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstring>

namespace sdm {

struct Mat3 {
    float m[9];
    float& operator()(int r, int c) { return m[r * 3 + c]; }
    float operator()(int r, int c) const { return m[r * 3 + c]; }
};
struct Vec3 {
    float v[3];
};

struct Pose {
    Mat3 R;
    Vec3 t;
    static Pose from_Tcw(const float* Tcw)  // rows 0..2 of [R|t], row-major 3x4
    {
        Pose p;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) p.R(r, c) = Tcw[r * 4 + c];
            p.t.v[r] = Tcw[r * 4 + 3];
        }
        return p;
    }
};

// A * B^T scaled by s (s is +1 or -1 on this path): double accumulation over k, cast at the end.
inline Mat3 mul_ABt(const Mat3& A, const Mat3& B, float s)
{
    Mat3 D;
    const double sd = (double)s;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += (double)A(r, k) * (double)B(c, k);
            D(r, c) = (float)(acc * sd);
        }
    return D;
}

inline float dot3f(float a0, float a1, float a2, float b0, float b1, float b2)
{
    float s = a0 * b0;
    s = s + a1 * b1;
    s = s + a2 * b2;
    return s;
}

inline Mat3 mul_f32(const Mat3& A, const Mat3& B)
{
    Mat3 D;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) D(r, c) = dot3f(A(r, 0), A(r, 1), A(r, 2), B(0, c), B(1, c), B(2, c));
    return D;
}

// M * x + c  (alpha = beta = 1)
inline Vec3 mul_vec_add(const Mat3& M, const Vec3& x, const Vec3& c)
{
    Vec3 d;
    for (int r = 0; r < 3; ++r) {
        float t = dot3f(M(r, 0), M(r, 1), M(r, 2), x.v[0], x.v[1], x.v[2]);
        d.v[r] = (float)((double)t + (double)c.v[r]);
    }
    return d;
}

inline Mat3 inverse_adjugate(const Mat3& S)
{
    Mat3 D;
    double det = S(0, 0) * ((double)S(1, 1) * S(2, 2) - (double)S(1, 2) * S(2, 1)) -
                 S(0, 1) * ((double)S(1, 0) * S(2, 2) - (double)S(1, 2) * S(2, 0)) +
                 S(0, 2) * ((double)S(1, 0) * S(2, 1) - (double)S(1, 1) * S(2, 0));
    if (det == 0.0) {
        std::memset(D.m, 0, sizeof(D.m));
        return D;
    }
    const double id = 1.0 / det;
    const int nx[3] = {1, 2, 0}, pv[3] = {2, 0, 1};
    // cofactor(c, r) / det ; the explicit (a*b - c*d) ordering matters for bit parity
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            const int r1 = nx[c], r2 = pv[c], c1 = nx[r], c2 = pv[r];
            double cof = (double)S(r1, c1) * S(r2, c2) - (double)S(r1, c2) * S(r2, c1);
            D(r, c) = (float)(cof * id);
        }
    return D;
}

// X = A^-1 B by float LU with partial pivoting, forward elimination on B, back substitution.
inline bool solve_lu(Mat3 A, Mat3 B, Mat3& X)
{
    const float tiny = FLT_EPSILON * 10;
    for (int i = 0; i < 3; ++i) {
        int p = i;
        for (int r = i + 1; r < 3; ++r)
            if (std::fabs(A(r, i)) > std::fabs(A(p, i))) p = r;
        if (std::fabs(A(p, i)) < tiny) {
            std::memset(X.m, 0, sizeof(X.m));
            return false;
        }
        if (p != i) {
            for (int c = i; c < 3; ++c) { float t = A(i, c); A(i, c) = A(p, c); A(p, c) = t; }
            for (int c = 0; c < 3; ++c) { float t = B(i, c); B(i, c) = B(p, c); B(p, c) = t; }
        }
        const float ninv = -1 / A(i, i);
        for (int r = i + 1; r < 3; ++r) {
            const float f = A(r, i) * ninv;
            for (int c = i + 1; c < 3; ++c) { float q = f * A(i, c); A(r, c) = A(r, c) + q; }
            for (int c = 0; c < 3; ++c) { float q = f * B(i, c); B(r, c) = B(r, c) + q; }
        }
    }
    for (int i = 2; i >= 0; --i)
        for (int c = 0; c < 3; ++c) {
            float s = B(i, c);
            for (int k = i + 1; k < 3; ++k) { float q = A(i, k) * B(k, c); s = s - q; }
            B(i, c) = s / A(i, i);
        }
    X = B;
    return true;
}

struct PairGeometry {
    Mat3 R21;
    Vec3 t21;
    Mat3 F12;
};

// kf1 = the keyframe being mapped, kf2 = one covisible neighbour.  K = {fx, fy, cx, cy}.
inline PairGeometry pair_geometry(const float* K1, const float* Tcw1, const float* K2, const float* Tcw2)
{
    const Pose p1 = Pose::from_Tcw(Tcw1), p2 = Pose::from_Tcw(Tcw2);
    PairGeometry g;
    g.R21 = mul_ABt(p2.R, p1.R, 1.0f);
    g.t21 = mul_vec_add(mul_ABt(p2.R, p1.R, -1.0f), p1.t, p2.t);

    const Mat3 R12 = mul_ABt(p1.R, p2.R, 1.0f);
    const Vec3 t12 = mul_vec_add(mul_ABt(p1.R, p2.R, -1.0f), p2.t, p1.t);
    Mat3 skew = {{0.f, -t12.v[2], t12.v[1], t12.v[2], 0.f, -t12.v[0], -t12.v[1], t12.v[0], 0.f}};
    Mat3 K1t = {{K1[0], 0.f, 0.f, 0.f, K1[1], 0.f, K1[2], K1[3], 1.f}};
    Mat3 K2m = {{K2[0], 0.f, K2[2], 0.f, K2[1], K2[3], 0.f, 0.f, 1.f}};
    Mat3 S;
    solve_lu(K1t, skew, S);
    g.F12 = mul_f32(mul_f32(S, R12), inverse_adjugate(K2m));
    return g;
}

// Twc rows 0..2 (3x4, row-major): Rwc = Rcw^T, Ow = -(Rwc * tcw)
inline void pose_inverse(const float* Tcw, float* Twc12)
{
    const Pose p = Pose::from_Tcw(Tcw);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Twc12[r * 4 + c] = p.R(c, r);
        float t = dot3f(p.R(0, r), p.R(1, r), p.R(2, r), p.t.v[0], p.t.v[1], p.t.v[2]);
        Twc12[r * 4 + 3] = -t;
    }
}

inline void stereo_search_constraints(const float* inv_depths, int n, float* min_depth, float* max_depth)
{
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += (double)inv_depths[i];
    const float mean = (float)acc / (float)n;
    double sq = 0.0;
    for (int i = 0; i < n; ++i) {
        const float d = inv_depths[i] - mean;
        const float dd = d * d;
        sq += (double)dd;
    }
    const float stdev = std::sqrt((float)(sq / (double)n));
    *max_depth = 1.0f / (mean + 2.0f * stdev);
    *min_depth = 1.0f / (mean - 2.0f * stdev);
}

// float thresholds equivalent to the reference's float-vs-double-literal compares:
//   (double)x > c  <=>  x > thr_gt(c)      (double)x < c  <=>  x < thr_lt(c)
inline float thr_gt(double c)
{
    float f = (float)c;
    if ((double)f > c) f = std::nextafterf(f, -INFINITY);
    return f;  // largest float <= c
}
inline float thr_lt(double c)
{
    float f = (float)c;
    if ((double)f < c) f = std::nextafterf(f, INFINITY);
    return f;  // smallest float >= c
}

}  // namespace sdm
