// cvstub.h — a minimal stand-in for the part of OpenCV core that EAO-SLAM's ProbabilityMapping.cc
// touches, so that the REFERENCE'S OWN SOURCE FILE can be compiled in this image (OpenCV C++ headers are
// absent) and executed as oracle/_ref/libref_pm.so.  TEST INFRASTRUCTURE ONLY.
//
// It is NOT OpenCV: only cv::Mat for CV_8U / CV_32S / CV_32F, the lazy MatExpr folding rules of
// modules/core/src/matop.cpp for the expression forms the file uses, and the CV_32F evaluation rules of
// gemm / invert / solve / convertTo / fastAtan2 that were pinned bit-exactly against real cv2 4.13
// (tests/golden/cv2_kats.npz, tests/golden/pair_geometry_cv2.npz).  What this buys: the control flow,
// loop bounds, gating, ordering and float-vs-double expression structure of the path come from the
// reference's text, not from a transcription.  Functions only reached from dead code of the reference
// (SVD, Scharr, magnitude, phase, DECOMP_SVD inversion) abort if they are ever executed.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6

namespace cv {

enum { DECOMP_LU = 0, DECOMP_SVD = 1 };
enum { GEMM_1_T = 1, GEMM_2_T = 2 };

[[noreturn]] inline void stub_dead(const char* what)
{
    fprintf(stderr, "cvstub: %s is only reachable from dead code of the reference and is not implemented\n", what);
    abort();
}

struct Scalar {
    double v[4];
    Scalar(double a = 0) { v[0] = a; v[1] = v[2] = v[3] = 0; }
};
struct KeyPoint {
    float angle;
    KeyPoint() : angle(-1) {}
};

class MatExpr;

class Mat {
public:
    int rows, cols;
    size_t step;  // bytes per row
    uchar* data;
    Mat() : rows(0), cols(0), step(0), data(NULL), type_(CV_32F) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); *this = s; }
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Scalar& s)  // Mat::operator=(const Scalar&): every element
    {
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) {
                if (type_ == CV_32F) at<float>(i, j) = (float)s.v[0];
                else if (type_ == CV_32S) at<int>(i, j) = (int)s.v[0];
                else at<uchar>(i, j) = (uchar)s.v[0];
            }
        return *this;
    }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        buf_.reset(new uchar[(size_t)r * step + 16](), std::default_delete<uchar[]>());
        data = buf_.get();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }  // create() value-initialises
    size_t elemSize() const { return type_ == CV_8U ? 1 : 4; }
    int type() const { return type_; }
    bool empty() const { return data == NULL || rows == 0 || cols == 0; }
    template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }  // cv::Mat::ptr<T>(row)
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    // single index: element i of a row or column vector (cv::Mat::at(int i0))
    template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <class T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    Mat view(int r0, int r1, int c0, int c1) const
    {
        Mat m;
        m.rows = r1 - r0; m.cols = c1 - c0; m.step = step; m.type_ = type_; m.buf_ = buf_;
        m.data = data + (size_t)r0 * step + (size_t)c0 * elemSize();
        return m;
    }
    Mat row(int r) const { return view(r, r + 1, 0, cols); }
    Mat col(int c) const { return view(0, rows, c, c + 1); }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int i = 0; i < rows; i++) memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat dst) const
    {
        for (int i = 0; i < rows; i++) memcpy(dst.data + (size_t)i * dst.step, data + (size_t)i * step, (size_t)cols * elemSize());
    }
    void push_back(const Mat& m)  // cv::Mat::push_back(const Mat&): append the rows of m
    {
        Mat n;
        n.create(rows + m.rows, m.cols, m.type_);
        for (int i = 0; i < rows; i++) memcpy(n.data + (size_t)i * n.step, data + (size_t)i * step, (size_t)cols * elemSize());
        for (int i = 0; i < m.rows; i++) memcpy(n.data + (size_t)(rows + i) * n.step, m.data + (size_t)i * m.step, (size_t)m.cols * m.elemSize());
        *this = n;
    }
    double dot(const Mat& m) const  // cv::Mat::dot: products accumulated in double (the semi-dense path never calls it; LineFit does, :809-810)
    {
        double s = 0;
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) s += (double)at<float>(i, j) * (double)m.at<float>(i, j);
        return s;
    }
    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const;

private:
    int type_;
    std::shared_ptr<uchar> buf_;
};

template <class T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, CV_32F) {}
};
// (cv::Mat_<float>(3,1) << a, b, c): MatCommaInitializer_, values converted to the element type
template <class T> struct CommaInit {
    Mat_<T> m;
    int i;
    CommaInit(const Mat_<T>& m_, T v) : m(m_), i(0) { put(v); }
    void put(T v) { m.template at<T>(i / m.cols, i % m.cols) = v; i++; }
    template <class V> CommaInit& operator,(V v) { put((T)v); return *this; }
    operator Mat() const { return m; }
};
template <class T, class V> CommaInit<T> operator<<(const Mat_<T>& m, V v) { return CommaInit<T>(m, (T)v); }

// ---------------------------------------------------------------------------------------------
// evaluation kernels (CV_32F), cv2-pinned rules
// ---------------------------------------------------------------------------------------------
inline float f_at(const Mat& m, int r, int c) { return m.at<float>(r, c); }

// cv::gemm(A, B, alpha, C, beta, D, flags) for CV_32F.  alpha/beta reach the kernels as FLOAT (hal::gemm32f).
inline Mat gemm_eval(const Mat& A, const Mat& B, double alpha, const Mat& C, double beta, int flags)
{
#ifndef CVSTUB_GEMM_ALPHA_DOUBLE  // OpenCV >= 3.3 (hal::gemm32f takes float alpha/beta; measured on cv2 4.13).  3.2.0 kept doubles.
    alpha = (double)(float)alpha;
    beta = (double)(float)beta;
#endif
    const bool at = flags & GEMM_1_T, bt = flags & GEMM_2_T;
    const int M = at ? A.cols : A.rows, len = at ? A.rows : A.cols, N = bt ? B.rows : B.cols;
    Mat D(M, N, CV_32F);
    const bool hasC = !C.empty();
    if (flags == 0 && 2 <= len && len <= 4 && (len == N || len == M)) {
        // small-matrix path: float products summed left to right, then (float)(t*alpha + c*beta) in double
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                float t = f_at(A, i, 0) * f_at(B, 0, j);
                for (int k = 1; k < len; k++) t = t + f_at(A, i, k) * f_at(B, k, j);
                D.at<float>(i, j) = hasC ? (float)((double)t * alpha + (double)f_at(C, i, j) * beta) : (float)((double)t * alpha);
            }
        return D;
    }
    // general path: double accumulation, k sequential
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int k = 0; k < len; k++) {
                const float a = at ? f_at(A, k, i) : f_at(A, i, k);
                const float b = bt ? f_at(B, j, k) : f_at(B, k, j);
                s += (double)a * (double)b;
            }
            s *= alpha;
            if (hasC) s += (double)f_at(C, i, j) * beta;
            D.at<float>(i, j) = (float)s;
        }
    return D;
}

// cv::invert, 3x3 CV_32F, DECOMP_LU: closed form with double products
inline Mat invert_eval(const Mat& S, int method)
{
    if (method != DECOMP_LU || S.rows != 3 || S.cols != 3) stub_dead("invert (other than 3x3 DECOMP_LU)");
#define Sf(y, x) ((double)S.at<float>(y, x))
    Mat D(3, 3, CV_32F);
    double d = S.at<float>(0, 0) * (Sf(1, 1) * S.at<float>(2, 2) - Sf(1, 2) * S.at<float>(2, 1)) -
               S.at<float>(0, 1) * (Sf(1, 0) * S.at<float>(2, 2) - Sf(1, 2) * S.at<float>(2, 0)) +
               S.at<float>(0, 2) * (Sf(1, 0) * S.at<float>(2, 1) - Sf(1, 1) * S.at<float>(2, 0));
    if (d != 0.) {
        d = 1. / d;
        double t[9];
        t[0] = (Sf(1, 1) * S.at<float>(2, 2) - Sf(1, 2) * S.at<float>(2, 1)) * d;
        t[1] = (Sf(0, 2) * S.at<float>(2, 1) - Sf(0, 1) * S.at<float>(2, 2)) * d;
        t[2] = (Sf(0, 1) * S.at<float>(1, 2) - Sf(0, 2) * S.at<float>(1, 1)) * d;
        t[3] = (Sf(1, 2) * S.at<float>(2, 0) - Sf(1, 0) * S.at<float>(2, 2)) * d;
        t[4] = (Sf(0, 0) * S.at<float>(2, 2) - Sf(0, 2) * S.at<float>(2, 0)) * d;
        t[5] = (Sf(0, 2) * S.at<float>(1, 0) - Sf(0, 0) * S.at<float>(1, 2)) * d;
        t[6] = (Sf(1, 0) * S.at<float>(2, 1) - Sf(1, 1) * S.at<float>(2, 0)) * d;
        t[7] = (Sf(0, 1) * S.at<float>(2, 0) - Sf(0, 0) * S.at<float>(2, 1)) * d;
        t[8] = (Sf(0, 0) * S.at<float>(1, 1) - Sf(0, 1) * S.at<float>(1, 0)) * d;
        for (int i = 0; i < 9; i++) D.at<float>(i / 3, i % 3) = (float)t[i];
    }
#undef Sf
    return D;
}

// cv::solve(A, B, X, DECOMP_LU) with a multi-column right-hand side: hal::LU32f (float, partial pivoting)
inline Mat solve_eval(const Mat& A_, const Mat& B_, int method)
{
    if (method != DECOMP_LU || A_.rows != A_.cols || B_.cols < 2) stub_dead("solve (other than LU with a matrix right-hand side)");
    Mat A = A_.clone(), b = B_.clone();
    const int m = A.rows, n = b.cols;
    const float eps = FLT_EPSILON * 10;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::abs(A.at<float>(j, i)) > std::abs(A.at<float>(k, i))) k = j;
        if (std::abs(A.at<float>(k, i)) < eps) return Mat(m, n, CV_32F);
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A.at<float>(i, j), A.at<float>(k, j));
            for (int j = 0; j < n; j++) std::swap(b.at<float>(i, j), b.at<float>(k, j));
        }
        const float d = -1 / A.at<float>(i, i);
        for (int j = i + 1; j < m; j++) {
            const float alpha = A.at<float>(j, i) * d;
            for (int q = i + 1; q < m; q++) A.at<float>(j, q) += alpha * A.at<float>(i, q);
            for (int q = 0; q < n; q++) b.at<float>(j, q) += alpha * b.at<float>(i, q);
        }
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
            float s = b.at<float>(i, j);
            for (int k = i + 1; k < m; k++) s -= A.at<float>(i, k) * b.at<float>(k, j);
            b.at<float>(i, j) = s / A.at<float>(i, i);
        }
    return b;
}

// ---------------------------------------------------------------------------------------------
// MatExpr: the lazy expression forms of matop.cpp that the file produces
// ---------------------------------------------------------------------------------------------
class MatExpr {
public:
    enum Op { IDENT, SCALED, T, INV, SOLVE, GEMM };  // SCALED = MatOp_AddEx with one operand: a*alpha (+ nothing)
    Op op;
    Mat a, b, c;
    double alpha, beta;
    int flags;
    MatExpr() : op(IDENT), alpha(1), beta(0), flags(0) {}
    MatExpr(const Mat& m) : op(IDENT), a(m), alpha(1), beta(0), flags(0) {}
    static MatExpr make(Op op, const Mat& a, const Mat& b, const Mat& c, double alpha, double beta, int flags)
    {
        MatExpr e; e.op = op; e.a = a; e.b = b; e.c = c; e.alpha = alpha; e.beta = beta; e.flags = flags;
        return e;
    }
    Mat eval() const
    {
        switch (op) {
        case IDENT: return a;
        case SCALED: {  // convertTo(type, alpha): cvtScale 32f->32f works in float
            Mat m(a.rows, a.cols, CV_32F);
            const float s = (float)alpha;
            for (int i = 0; i < a.rows; i++)
                for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = a.at<float>(i, j) * s;
            return m;
        }
        case T: {  // transpose, then scale if alpha != 1 (MatOp_T::assign)
            Mat m(a.cols, a.rows, CV_32F);
            for (int i = 0; i < a.rows; i++)
                for (int j = 0; j < a.cols; j++) m.at<float>(j, i) = a.at<float>(i, j);
            if (alpha != 1) return make(SCALED, m, Mat(), Mat(), alpha, 0, 0).eval();
            return m;
        }
        case INV: return invert_eval(a, flags);
        case SOLVE: return solve_eval(a, b, flags);
        case GEMM: return gemm_eval(a, b, alpha, c, beta, flags);
        }
        return Mat();
    }
    operator Mat() const { return eval(); }
    template <class Tp> Tp at(int r, int c) const { return eval().at<Tp>(r, c); }
    MatExpr t() const  // only on identity in this file
    {
        if (op == IDENT) return make(T, a, Mat(), Mat(), 1, 0, 0);
        return make(T, eval(), Mat(), Mat(), 1, 0, 0);
    }
    MatExpr inv(int method = DECOMP_LU) const { return make(INV, eval(), Mat(), Mat(), 1, 0, method); }  // MatOp::invert: materialise
};

inline Mat::Mat(const MatExpr& e) { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e)
{
    // MatOp::assign -> Mat::create(): a destination of the same size and type keeps its buffer and is written in
    // place (this is what makes `A.row(0) = expr` work); otherwise it is re-allocated
    Mat m = e.eval();
    if (data && rows == m.rows && cols == m.cols && type_ == m.type_) {
        if (m.data != data) m.copyTo(*this);
        return *this;
    }
    rows = m.rows; cols = m.cols; step = m.step; data = m.data; type_ = m.type_; buf_ = m.buf_;
    return *this;
}
inline MatExpr Mat::t() const { return MatExpr::make(MatExpr::T, *this, Mat(), Mat(), 1, 0, 0); }
inline MatExpr Mat::inv(int method) const { return MatExpr::make(MatExpr::INV, *this, Mat(), Mat(), 1, 0, method); }

// MatOp::matmul (+ MatOp_Invert::matmul): fold transposes / scales into one gemm, inv(A)*B into solve
inline MatExpr matmul(const MatExpr& e1, const MatExpr& e2)
{
    if (e1.op == MatExpr::INV && e2.op == MatExpr::IDENT)
        return MatExpr::make(MatExpr::SOLVE, e1.a, e2.a, Mat(), 1, 0, e1.flags);
    double alpha = 1;
    int flags = 0;
    Mat m1, m2;
    if (e1.op == MatExpr::T) { flags = GEMM_1_T; alpha = e1.alpha; m1 = e1.a; }
    else if (e1.op == MatExpr::SCALED) { alpha = e1.alpha; m1 = e1.a; }
    else m1 = e1.eval();
    if (e2.op == MatExpr::T) { flags |= GEMM_2_T; alpha *= e2.alpha; m2 = e2.a; }
    else if (e2.op == MatExpr::SCALED) { alpha *= e2.alpha; m2 = e2.a; }
    else m2 = e2.eval();
    return MatExpr::make(MatExpr::GEMM, m1, m2, Mat(), alpha, 0, flags);
}
inline MatExpr operator*(const Mat& a, const Mat& b) { return matmul(MatExpr(a), MatExpr(b)); }
inline MatExpr operator*(const MatExpr& a, const Mat& b) { return matmul(a, MatExpr(b)); }
inline MatExpr operator*(const Mat& a, const MatExpr& b) { return matmul(MatExpr(a), b); }
inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return matmul(a, b); }

// op->multiply(e, s): GEMM scales alpha (and beta), T scales alpha, a plain Mat becomes SCALED
inline MatExpr scale(const MatExpr& e, double s)
{
    if (e.op == MatExpr::GEMM) { MatExpr r = e; r.alpha *= s; r.beta *= s; return r; }
    if (e.op == MatExpr::T || e.op == MatExpr::SCALED) { MatExpr r = e; r.alpha *= s; return r; }
    return MatExpr::make(MatExpr::SCALED, e.eval(), Mat(), Mat(), s, 0, 0);
}
inline MatExpr operator*(const MatExpr& e, double s) { return scale(e, s); }
inline MatExpr operator*(double s, const MatExpr& e) { return scale(e, s); }
inline MatExpr operator*(const Mat& m, double s) { return scale(MatExpr(m), s); }
inline MatExpr operator*(double s, const Mat& m) { return scale(MatExpr(m), s); }
inline MatExpr operator/(const MatExpr& e, double s) { return scale(e, 1. / s); }
inline MatExpr operator/(const Mat& m, double s) { return scale(MatExpr(m), 1. / s); }
// unary minus: on a Mat a SCALED(-1); on an expression MatOp::subtract(Scalar(0), e): materialise, then SCALED(-1)
inline MatExpr operator-(const Mat& m) { return MatExpr::make(MatExpr::SCALED, m, Mat(), Mat(), -1, 0, 0); }
inline MatExpr operator-(const MatExpr& e) { return MatExpr::make(MatExpr::SCALED, e.eval(), Mat(), Mat(), -1, 0, 0); }
// MatOp_GEMM::add: a product without C absorbs a plain Mat as C with beta = 1
inline MatExpr add_expr(const MatExpr& e1, const MatExpr& e2)
{
    const bool prod1 = e1.op == MatExpr::GEMM && (e1.c.empty() || e1.beta == 0);
    const bool prod2 = e2.op == MatExpr::GEMM && (e2.c.empty() || e2.beta == 0);
    if (prod1 && e2.op == MatExpr::IDENT) return MatExpr::make(MatExpr::GEMM, e1.a, e1.b, e2.a, e1.alpha, 1, e1.flags);
    if (prod2 && e1.op == MatExpr::IDENT) return MatExpr::make(MatExpr::GEMM, e2.a, e2.b, e1.a, e2.alpha, 1, e2.flags);
    stub_dead("Mat addition other than product + matrix");
}
inline MatExpr operator+(const MatExpr& a, const Mat& b) { return add_expr(a, MatExpr(b)); }
inline MatExpr operator+(const Mat& a, const MatExpr& b) { return add_expr(MatExpr(a), b); }
inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return add_expr(a, b); }
inline MatExpr operator-(const MatExpr&, const Mat&) { stub_dead("Mat subtraction"); }
inline MatExpr operator-(const MatExpr&, const MatExpr&) { stub_dead("Mat subtraction"); }

// ---------------------------------------------------------------------------------------------
// free functions
// ---------------------------------------------------------------------------------------------
inline float fastAtan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = std::abs(x), ay = std::abs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}
template <class Tp> inline const Tp& max(const Tp& a, const Tp& b) { return std::max(a, b); }
inline void Scharr(const Mat&, Mat&, int, int, int, double) { stub_dead("cv::Scharr"); }
inline void magnitude(const Mat&, const Mat&, Mat&) { stub_dead("cv::magnitude"); }
inline void phase(const Mat&, const Mat&, Mat&, bool) { stub_dead("cv::phase"); }
struct SVD {
    enum { MODIFY_A = 1, FULL_UV = 4 };
    static void compute(const Mat&, Mat&, Mat&, Mat&, int) { stub_dead("cv::SVD::compute"); }
    static void solveZ(const Mat& A, Mat& u);  // defined in cvstub_linefit.h (exact stand-in), used by LineDetector::LeastSquaresLineFit only
};

}  // namespace cv
