// cvstub_linefit.h - the OpenCV calls of LineDetector::LineFit and its helpers (LineDetector.cc:578-840) on top of cvstub.h.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_linefit.so: the reference's own text of those functions, extracted at build time
// where it lies, compiled against these stand-ins; see oracle/Makefile target `ref_linefit`).
//
// The two solver calls are stand-ins with EXACT arithmetic, not OpenCV's float SVD:
//   cv::SVD::solveZ(A, u)            unit vector minimising |A u|: smallest eigenvector of A^T A, cyclic Jacobi in double
//   cv::solve(A, b, u, DECOMP_SVD)   minimum-norm least squares of an n x 2 system, normal equations in double
// so this library pins the CONTROL FLOW of the device kernel (which windows start a line, where a line stops, what is
// emitted, in which order) against the reference's own source with the solver noise taken out; how far OpenCV's float SVD
// moves the threshold decisions is what the cv2-based oracle (oracle/linefit_oracle.py) measures.
// Everything else follows OpenCV: cv::norm(Point2f) = sqrt((double)x*x + (double)y*y); cv::norm(Mat) / cv::norm(Mat, Mat)
// accumulate squares of the float elements in double; Mat::dot accumulates in double; A*u (n x 3 times 3 x 1) is the
// general gemm (double accumulation, rounded to float) of cvstub.h; Twc*Pc (4 x 4 times 4 x 1) its small-matrix path.
#pragma once
#include <limits>

#include "cvstub.h"

#ifndef CV_PI
#define CV_PI 3.1415926535897932384626433832795
#endif

namespace cv {

struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
inline Point2f operator-(const Point2f& a, const Point2f& b) { return Point2f(a.x - b.x, a.y - b.y); }
inline double norm(const Point2f& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

inline double norm(const Mat& m)
{
    double s = 0;
    for (int i = 0; i < m.rows; i++)
        for (int j = 0; j < m.cols; j++) s += (double)m.at<float>(i, j) * (double)m.at<float>(i, j);
    return std::sqrt(s);
}
inline double norm(const Mat& a, const Mat& b)
{
    double s = 0;
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) {
            const double d = (double)a.at<float>(i, j) - (double)b.at<float>(i, j);  // (float difference is exact enough: see header)
            s += d * d;
        }
    return std::sqrt(s);
}
inline Mat operator-(const Mat& a, const Mat& b)
{
    Mat d(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) d.at<float>(i, j) = a.at<float>(i, j) - b.at<float>(i, j);
    return d;
}

// smallest eigenvector of the symmetric 3x3 M, cyclic Jacobi in double
inline void jacobi_smallest3(double M[3][3], double v[3])
{
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = std::fabs(M[0][1]) + std::fabs(M[0][2]) + std::fabs(M[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (M[p][q] == 0.0) continue;
                const double theta = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) { const double a = M[k][p], b = M[k][q]; M[k][p] = cs * a - sn * b; M[k][q] = sn * a + cs * b; }
                for (int k = 0; k < 3; ++k) { const double a = M[p][k], b = M[q][k]; M[p][k] = cs * a - sn * b; M[q][k] = sn * a + cs * b; }
                for (int k = 0; k < 3; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = cs * a - sn * b; V[k][q] = sn * a + cs * b; }
            }
    }
    int m = 0;
    if (M[1][1] < M[m][m]) m = 1;
    if (M[2][2] < M[m][m]) m = 2;
    for (int r = 0; r < 3; ++r) v[r] = V[r][m];
}

inline void SVD::solveZ(const Mat& A, Mat& u)
{
    {
        if (A.cols != 3) stub_dead("SVD::solveZ (other than n x 3)");
        double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < A.rows; i++)
            for (int p = 0; p < 3; p++)
                for (int q = 0; q < 3; q++) M[p][q] += (double)A.at<float>(i, p) * (double)A.at<float>(i, q);
        double v[3];
        jacobi_smallest3(M, v);
        u = Mat(3, 1, CV_32F);
        for (int r = 0; r < 3; r++) u.at<float>(r, 0) = (float)v[r];
    }
}

inline bool solve(const Mat& A, const Mat& b, Mat& x, int method)
{
    if (method != DECOMP_SVD || A.cols != 2 || b.cols != 1) stub_dead("solve (other than n x 2 DECOMP_SVD)");
    double Saa = 0, Sab = 0, Sbb = 0, Say = 0, Sby = 0;
    for (int i = 0; i < A.rows; i++) {
        const double a0 = A.at<float>(i, 0), a1 = A.at<float>(i, 1), y = b.at<float>(i, 0);
        Saa += a0 * a0; Sab += a0 * a1; Sbb += a1 * a1; Say += a0 * y; Sby += a1 * y;
    }
    const double det = Saa * Sbb - Sab * Sab;
    double x0 = 0, x1 = 0;
    if (Sbb > 0 && det > 1e-9 * (Saa * Sbb + 1e-30)) {
        x0 = (Say * Sbb - Sab * Sby) / det;
        x1 = (Saa * Sby - Sab * Say) / det;
    } else if (Sbb > 0) {  // every non-zero row is a multiple of (d, 1): minimum-norm solution along that row
        const double d = Sab / Sbb, s = (Sby / Sbb) / (d * d + 1.0);
        x0 = d * s; x1 = s;
    }
    x = Mat(2, 1, CV_32F);
    x.at<float>(0, 0) = (float)x0;
    x.at<float>(1, 0) = (float)x1;
    return true;
}

}  // namespace cv
