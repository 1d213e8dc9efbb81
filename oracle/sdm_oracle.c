/*
 * sdm_oracle.c — CPU restatement of EAO-SLAM's semi-dense ProbabilityMapping hot path.
 * TEST INFRASTRUCTURE ONLY (see sdm_oracle.h).  "parity unpinned" by reference tests; OpenCV
 * primitives pinned against cv2 4.13 (tests/golden/cv2_kats.npz).
 *
 * Canonical build: gcc -O2 -ffp-contract=off -fopenmp  (no FMA contraction; IEEE float/double).
 * All citations are /root/reference/src/ProbabilityMapping.cc unless another file is named.
 *
 * Arithmetic conventions:  f() = round to float, d() = to double.
 */
#include "sdm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_default_params(oracle_params* p)
{
    p->lambdaG = 8;
    p->lambdaL = 80;
    p->lambdaTheta = 45;
    p->lambdaN = 3;
    p->theta = (float)0.23;
    p->sigmaI = 20.0f;
    p->chi2_fusion = 5.99;
    p->chi2_inter = 3.84;
    p->eps = 0.000001;
    p->slope_max = 4.0f;
    p->intra_check = 0;
    p->intra_grow = 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1 to their workers: the CPU baseline sets its thread count
 * explicitly (the reference runs its omp loops on every core, CMakeLists.txt:47-52) */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * OpenCV core primitives, restated (OpenCV 3.x modules/core/src/{matmul,lapack,mathfuncs}.cpp;
 * OpenCV is an un-vendored dependency of the reference: CMakeLists.txt:35-41).
 * ------------------------------------------------------------------------------------------ */

/* cv::fastAtan2 (used at :791) — degree polynomial, float arithmetic */
float ocv_fastAtan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
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

/* NOTE on alpha/beta: for CV_32F operands cv::gemm hands alpha and beta to hal::gemm32f as FLOAT
 * (static_cast<float>) before any arithmetic; inside, they are used as doubles again.  Verified
 * against cv2 4.13 (tests/golden/cv2_kats.npz: a double alpha that is not float-representable
 * behaves exactly like (float)alpha on every path: small-matrix, general, transposed). */
#define F32(x) ((double)(float)(x))

/* gemm(A, B, alpha, GEMM_2_T) for 3x3: general path, double accumulation, k sequential */
void ocv_mul33_ABt(const float* A, const float* B, double alpha, float* D)
{
    alpha = F32(alpha);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += (double)A[i * 3 + k] * (double)B[j * 3 + k];
            D[i * 3 + j] = (float)(s * alpha);
        }
}

/* gemm 3x3 * 3x3, flags 0: small-matrix case, float 3-term dot left to right */
void ocv_mul33(const float* A, const float* B, float* D)
{
    float T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float t = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
            T[i * 3 + j] = t;
        }
    memcpy(D, T, sizeof(T));
}

/* gemm 3x3 * 3x1 (flags 0): float dot, then (float)(t*alpha + c*beta) in double */
void ocv_mul33_vec(const float* A, const float* x, double alpha, const float* c, double beta, float* d)
{
    float T[3];
    alpha = F32(alpha);
    beta = F32(beta);
    for (int i = 0; i < 3; i++) {
        float t = A[i * 3 + 0] * x[0] + A[i * 3 + 1] * x[1] + A[i * 3 + 2] * x[2];
        if (c)
            T[i] = (float)((double)t * alpha + (double)c[i] * beta);
        else
            T[i] = (float)((double)t * alpha);
    }
    d[0] = T[0]; d[1] = T[1]; d[2] = T[2];
}

/* gemm 1x3 * 3x1: general path (double accumulation), * alpha in double */
float ocv_dot3_d(const float* a, const float* x, double alpha)
{
    double s = 0;
    alpha = F32(alpha);
    for (int k = 0; k < 3; k++) s += (double)a[k] * (double)x[k];
    return (float)(s * alpha);
}

/* gemm 1xn * nx1 (J^T r, J^T J at :1285-1286): general path, double accumulation */
float ocv_dotn_d(const float* a, const float* b, int n, double alpha)
{
    double s = 0;
    alpha = F32(alpha);
    for (int k = 0; k < n; k++) s += (double)a[k] * (double)b[k];
    return (float)(s * alpha);
}

/* gemm 4x4 * 4x1 (Twc*Pc at :723): small-matrix case, float 4-term dot */
void ocv_mul44_vec(const float* A, const float* x, float* d)
{
    for (int i = 0; i < 4; i++)
        d[i] = A[i * 4 + 0] * x[0] + A[i * 4 + 1] * x[1] + A[i * 4 + 2] * x[2] + A[i * 4 + 3] * x[3];
}

/* cv::invert, 3x3 CV_32F, DECOMP_LU: closed-form adjugate with double products */
void ocv_inv33(const float* S, float* D)
{
#define Sf(y, x) S[(y) * 3 + (x)]
    double d = Sf(0, 0) * ((double)Sf(1, 1) * Sf(2, 2) - (double)Sf(1, 2) * Sf(2, 1)) -
               Sf(0, 1) * ((double)Sf(1, 0) * Sf(2, 2) - (double)Sf(1, 2) * Sf(2, 0)) +
               Sf(0, 2) * ((double)Sf(1, 0) * Sf(2, 1) - (double)Sf(1, 1) * Sf(2, 0));
    if (d != 0.) {
        double t[9];
        d = 1. / d;
        t[0] = (((double)Sf(1, 1) * Sf(2, 2) - (double)Sf(1, 2) * Sf(2, 1)) * d);
        t[1] = (((double)Sf(0, 2) * Sf(2, 1) - (double)Sf(0, 1) * Sf(2, 2)) * d);
        t[2] = (((double)Sf(0, 1) * Sf(1, 2) - (double)Sf(0, 2) * Sf(1, 1)) * d);
        t[3] = (((double)Sf(1, 2) * Sf(2, 0) - (double)Sf(1, 0) * Sf(2, 2)) * d);
        t[4] = (((double)Sf(0, 0) * Sf(2, 2) - (double)Sf(0, 2) * Sf(2, 0)) * d);
        t[5] = (((double)Sf(0, 2) * Sf(1, 0) - (double)Sf(0, 0) * Sf(1, 2)) * d);
        t[6] = (((double)Sf(1, 0) * Sf(2, 1) - (double)Sf(1, 1) * Sf(2, 0)) * d);
        t[7] = (((double)Sf(0, 1) * Sf(2, 0) - (double)Sf(0, 0) * Sf(2, 1)) * d);
        t[8] = (((double)Sf(0, 0) * Sf(1, 1) - (double)Sf(0, 1) * Sf(1, 0)) * d);
        for (int i = 0; i < 9; i++) D[i] = (float)t[i];
    } else {
        for (int i = 0; i < 9; i++) D[i] = 0.f;
    }
#undef Sf
}

/* cv::solve(A, B, X, DECOMP_LU), 3x3 with a 3x3 right-hand side: hal::LU32f, float, partial pivoting.
 * This is what `K1.t().inv() * t12x` evaluates to (:1708; MatOp_Invert::matmul -> MatOp_Solve). */
int ocv_solve33_lu(const float* A_, const float* B_, float* X)
{
    float A[9], b[9];
    memcpy(A, A_, sizeof(A));
    memcpy(b, B_, sizeof(b));
    const int m = 3, n = 3;
    const float eps = FLT_EPSILON * 10;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (fabsf(A[j * 3 + i]) > fabsf(A[k * 3 + i])) k = j;
        if (fabsf(A[k * 3 + i]) < eps) {
            for (int q = 0; q < 9; q++) X[q] = 0.f;
            return 0;
        }
        if (k != i) {
            for (int j = i; j < m; j++) { float t = A[i * 3 + j]; A[i * 3 + j] = A[k * 3 + j]; A[k * 3 + j] = t; }
            for (int j = 0; j < n; j++) { float t = b[i * 3 + j]; b[i * 3 + j] = b[k * 3 + j]; b[k * 3 + j] = t; }
        }
        float d = -1 / A[i * 3 + i];
        for (int j = i + 1; j < m; j++) {
            float alpha = A[j * 3 + i] * d;
            for (int q = i + 1; q < m; q++) A[j * 3 + q] += alpha * A[i * 3 + q];
            for (int q = 0; q < n; q++) b[j * 3 + q] += alpha * b[i * 3 + q];
        }
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
            float s = b[i * 3 + j];
            for (int k = i + 1; k < m; k++) s -= A[i * 3 + k] * b[k * 3 + j];
            b[i * 3 + j] = s / A[i * 3 + i];
        }
    memcpy(X, b, sizeof(b));
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Pose helpers
 * ------------------------------------------------------------------------------------------ */

static void kf_Rt(const oracle_kf* kf, float* R, float* t)
{
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) R[i * 3 + j] = kf->Tcw[i * 4 + j];
        t[i] = kf->Tcw[i * 4 + 3];
    }
}

/* KeyFrame::SetPose, src/KeyFrame.cc:108-124: Rwc = Rcw^T, Ow = -Rwc*tcw (gemm alpha=-1) */
void oracle_pose_inverse(const float* Tcw, float* Twc)
{
    float Rwc[9], tcw[3], Ow[3];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) Rwc[i * 3 + j] = Tcw[j * 4 + i];
        tcw[i] = Tcw[i * 4 + 3];
    }
    ocv_mul33_vec(Rwc, tcw, -1.0, NULL, 0.0, Ow);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) Twc[i * 4 + j] = Rwc[i * 3 + j];
        Twc[i * 4 + 3] = Ow[i];
    }
    Twc[12] = 0; Twc[13] = 0; Twc[14] = 0; Twc[15] = 1;
}

/* R21, t21 (:1136-1137, :1582-1583, :1611-1612) and F12 = ComputeFundamental (:1694-1709),
 * GetSkewSymmetricMatrix (:1711-1715).  Hoisted per pair; the reference recomputes R21/t21 per
 * pixel from the same poses, with identical results. */
void oracle_pair_geometry(const oracle_kf* kf1, const oracle_kf* kf2, oracle_pair* out)
{
    float R1[9], t1[3], R2[9], t2[3], M[9];
    kf_Rt(kf1, R1, t1);
    kf_Rt(kf2, R2, t2);

    /* R21 = Rcw2*Rcw1.t() ; t21 = -Rcw2*Rcw1.t()*tcw1 + tcw2 */
    ocv_mul33_ABt(R2, R1, 1.0, out->R21);
    ocv_mul33_ABt(R2, R1, -1.0, M);
    ocv_mul33_vec(M, t1, 1.0, t2, 1.0, out->t21);

    /* R12 = R1w*R2w.t() ; t12 = -R1w*R2w.t()*t2w + t1w */
    float R12[9], t12[3];
    ocv_mul33_ABt(R1, R2, 1.0, R12);
    ocv_mul33_ABt(R1, R2, -1.0, M);
    ocv_mul33_vec(M, t2, 1.0, t1, 1.0, t12);

    float t12x[9] = {0, -t12[2], t12[1], t12[2], 0, -t12[0], -t12[1], t12[0], 0};
    float K1t[9] = {kf1->fx, 0, 0, 0, kf1->fy, 0, kf1->cx, kf1->cy, 1};
    float K2[9] = {kf2->fx, 0, kf2->cx, 0, kf2->fy, kf2->cy, 0, 0, 1};
    float S[9], SR[9], K2i[9];
    /* K1.t().inv()*t12x*R12*K2.inv() */
    ocv_solve33_lu(K1t, t12x, S);
    ocv_mul33(S, R12, SR);
    ocv_inv33(K2, K2i);
    ocv_mul33(SR, K2i, out->F12);
}

/* StereoSearchConstraints (:734-747): inv_depths = KeyFrame::GetAllPointDepths (KeyFrame.cc:756-787) */
void oracle_stereo_search_constraints(const float* inv_depths, int n, float* min_depth, float* max_depth)
{
    double acc = 0.0;
    for (int i = 0; i < n; i++) acc = acc + inv_depths[i];
    float sum = (float)acc;
    float mean = sum / (float)n;
    double var = 0.0;
    for (int i = 0; i < n; i++) {
        float df = inv_depths[i] - mean;
        var = var + df * df; /* float product promoted, std::inner_product with init 0.0 */
    }
    float variance = (float)(var / n);
    float stdev = sqrtf(variance);
    *max_depth = 1 / (mean + 2 * stdev);
    *min_depth = 1 / (mean - 2 * stdev);
}

/* ------------------------------------------------------------------------------------------
 * Interpolators (:66-111). x is an integer column; row y0+1 is clamped to H-1 (Appendix A.11:
 * the reference reads one row past the end with weight 0 when vj == rows-1 exactly).
 * ------------------------------------------------------------------------------------------ */
static inline float ylinear_f(const float* img, int W, int H, float y, int x0)
{
    int y0 = (int)floorf(y);
    int y1 = y0 + 1;
    float w0 = y1 - y;
    float w1 = y - y0;
    int y1c = y1 < H ? y1 : H - 1;
    return img[y0 * W + x0] * w0 + img[y1c * W + x0] * w1;
}

static inline float ylinear_u8(const uint8_t* img, int W, int H, float y, int x0)
{
    int y0 = (int)floorf(y);
    int y1 = y0 + 1;
    float w0 = y1 - y;
    float w1 = y - y0;
    int y1c = y1 < H ? y1 : H - 1;
    return img[y0 * W + x0] * w0 + img[y1c * W + x0] * w1;
}

static inline float yangle_f(const float* img, int W, int H, float y, int x0)
{
    int y0 = (int)floorf(y);
    int y1 = y0 + 1;
    float w0 = y1 - y;
    float w1 = y - y0;
    int y1c = y1 < H ? y1 : H - 1;
    float a0 = img[y0 * W + x0];
    float a1 = img[y1c * W + x0];
    if (fabsf(a0 - a1) < 180) {
        return a0 * w0 + a1 * w1;
    } else {
        if (a0 < a1) a0 += 360; else a1 += 360;
        float inter = a0 * w0 + a1 * w1;
        if (inter >= 360) inter -= 360;
        return inter;
    }
}

/* GetSearchRange (:1598-1631) */
void oracle_get_search_range(const oracle_kf* kf, const oracle_pair* pr, int px, int py,
                             float mind, float maxd, float* umin_, float* umax_)
{
    float fx = kf->fx, cx = kf->cx, fy = kf->fy, cy = kf->cy;
    float xp1[3] = {(px - cx) / fx, (py - cy) / fy, 1.0f};
    float xmin[3], xmax[3];
    ocv_mul33_vec(pr->R21, xp1, (double)mind, pr->t21, 1.0, xmin);
    ocv_mul33_vec(pr->R21, xp1, (double)maxd, pr->t21, 1.0, xmax);
    float umin = fx * xmin[0] / xmin[2] + cx;
    float umax = fx * xmax[0] / xmax[2] + cx;
    if (umin > umax) { float t = umax; umax = umin; umin = t; }
    if (umin < 0) umin = 0;
    if (umax < 0) umax = 0;
    if (umin > kf->W) umin = kf->W - 1;
    if (umax > kf->W) umax = kf->W - 1;
    *umin_ = umin;
    *umax_ = umax;
}

/* GetPixelDepth, equation 8 (:1568-1596) */
static inline float pixel_depth(float uj, int px, int py, const oracle_kf* kf, const oracle_pair* pr)
{
    float fx = kf->fx, cx = kf->cx, fy = kf->fy, cy = kf->cy;
    float ucx = uj - cx;
    float xp[3] = {(px - cx) / fx, (py - cy) / fy, 1.0f};
    float num1 = ocv_dot3_d(&pr->R21[6], xp, (double)ucx);
    float num2 = ocv_dot3_d(&pr->R21[0], xp, (double)fx);
    float denom1 = -pr->t21[2] * ucx;
    float denom2 = fx * pr->t21[0];
    return (num1 - num2) / (denom1 + denom2);
}

/* EpipolarSearch (:749-845) + ComputeInvDepthHypothesis (:1310-1335) */
int oracle_epipolar_search(const oracle_kf* kf1, const oracle_kf* kf2, const oracle_pair* pr,
                           int x, int y, float pixel, float min_depth, float max_depth,
                           float th_pi, float rot, const oracle_params* prm,
                           float* depth, float* sigma, float* best_u_, float* best_v_,
                           oracle_stats* st)
{
    const float* F12 = pr->F12;
    const int W = kf2->W, H = kf2->H;
    float a = x * F12[0] + y * F12[3] + F12[6];
    float b = x * F12[1] + y * F12[4] + F12[7];
    float c = x * F12[2] + y * F12[5] + F12[8];

    if ((a / b) < -prm->slope_max || a / b > prm->slope_max) return 0;
    if (a / b != a / b) return 0; /* NaN line (a=b=0): the reference would index out of bounds */

    float old_err = 100000.0f;
    float best_photometric_err = 0.0f;
    float best_gradient_modulo_err = 0.0f;
    int best_pixel = 0;

    float umin = 0.0f, umax = 0.0f;
    oracle_get_search_range(kf1, pr, x, y, min_depth, max_depth, &umin, &umax);
    if (umin != umin || umax != umax) return 0; /* NaN range: no iterations of interest */

    const float gradc = kf1->grad[y * kf1->W + x];
    const float theta_ = prm->theta;
    long long scanned = 0, evaluated = 0;

    for (int uj = (int)ceilf(umin); uj <= floorf(umax); uj++) {
        scanned++;
        float vj = -((a / b) * uj + (c / b));
        if (floorf(vj) < 0 || ceilf(vj) >= H) continue;

        int uj_plus = uj + 1;
        int uj_minus = uj - 1;
        if (uj_plus >= W) continue;
        if (uj_minus < 0) continue;

        float vj_plus = -((a / b) * uj_plus + (c / b));
        float vj_minus = -((a / b) * uj_minus + (c / b));
        if (floorf(vj_plus) < 0 || ceilf(vj_plus) >= H) continue;
        if (floorf(vj_minus) < 0 || ceilf(vj_minus) >= H) continue;

        /* condition 1 */
        if (ylinear_f(kf2->grad, W, H, vj, uj) <= prm->lambdaG) continue;

        /* condition 2 */
        float th_epipolar_line = ocv_fastAtan2(-a / b, 1);
        float temp_gradth = yangle_f(kf2->theta, W, H, vj, uj);
        float ang_diff = temp_gradth - th_epipolar_line;
        if (ang_diff >= 360) ang_diff -= 360;
        if (ang_diff < 0) ang_diff += 360;
        if (ang_diff > 180) ang_diff = 360 - ang_diff;
        if (ang_diff > 90) ang_diff = 180 - ang_diff;
        if (ang_diff >= prm->lambdaL) continue;

        /* condition 3 */
        float ang_pi_rot = th_pi + rot;
        if (ang_pi_rot >= 360) ang_pi_rot -= 360;
        if (ang_pi_rot < 0) ang_pi_rot += 360;
        float th_diff = temp_gradth - ang_pi_rot;
        if (th_diff >= 360) th_diff -= 360;
        if (th_diff < 0) th_diff += 360;
        if (th_diff > 180) th_diff = 360 - th_diff;
        if (th_diff >= prm->lambdaTheta) continue;

        evaluated++;
        float photometric_err = pixel - ylinear_u8(kf2->im, W, H, vj, uj);
        float gradient_modulo_err = gradc - ylinear_f(kf2->grad, W, H, vj, uj);
        float err = (photometric_err * photometric_err + (gradient_modulo_err * gradient_modulo_err) / theta_);
        if (err < old_err) {
            best_pixel = uj;
            old_err = err;
            best_photometric_err = photometric_err;
            best_gradient_modulo_err = gradient_modulo_err;
        }
    }
    if (st) { st->scanned += scanned; st->evaluated += evaluated; }

    if (old_err < 100000.0) {
        int uj_plus = best_pixel + 1;
        int uj_minus = best_pixel - 1;
        float vj_plus = -((a / b) * uj_plus + (c / b));
        float vj_minus = -((a / b) * uj_minus + (c / b));

        float g = (ylinear_u8(kf2->im, W, H, vj_plus, uj_plus) - ylinear_u8(kf2->im, W, H, vj_minus, uj_minus)) / 2;
        float q = (ylinear_f(kf2->grad, W, H, vj_plus, uj_plus) - ylinear_f(kf2->grad, W, H, vj_minus, uj_minus)) / 2;

        float denomiator = (g * g + (1 / theta_) * q * q);
        float ustar = best_pixel + (g * best_photometric_err + (1 / theta_) * q * best_gradient_modulo_err) / denomiator;
        float ustar_var = (2 * prm->sigmaI * prm->sigmaI / denomiator);

        *best_u_ = ustar;
        *best_v_ = -((a / b) * ustar + (c / b));

        /* ComputeInvDepthHypothesis (:1310-1335) */
        float inv_pixel_depth = pixel_depth(ustar, x, y, kf1, pr);
        float ustar_min = ustar - sqrtf(ustar_var);
        float inv_depth_min = pixel_depth(ustar_min, x, y, kf1, pr);
        float ustar_max = ustar + sqrtf(ustar_var);
        float inv_depth_max = pixel_depth(ustar_max, x, y, kf1, pr);

        float s1 = fabsf(inv_depth_max - inv_pixel_depth);
        float s2 = fabsf(inv_depth_min - inv_pixel_depth);
        float sigma_depth = (s1 < s2) ? s2 : s1; /* cv::max == std::max */

        *depth = inv_pixel_depth;
        *sigma = sigma_depth;
        return 1;
    }
    return 0;
}

/* ChiTest(depthHo) (:1633-1639) and ChiTest(float...) (:1641-1645) */
static inline int chi_test(float a, float b, float sa, float sb, const oracle_params* prm)
{
    float num = (a - b) * (a - b);
    float chi = num / (sa * sa) + num / (sb * sb);
    return (double)chi < prm->chi2_fusion;
}

/* GetFusion(vector<depthHo>) (:1669-1692): float accumulators updated through double */
static void get_fusion(const float* depth, const float* sigma, int n, float* out_depth,
                       float* out_sigma, float* min_sigma)
{
    float temp_min_sigma = sigma[0];
    float pjsj = 0, rsj = 0;
    for (int j = 0; j < n; j++) {
        double s2 = (double)sigma[j] * (double)sigma[j]; /* pow(sigma,2) */
        pjsj = (float)((double)pjsj + (double)depth[j] / s2);
        rsj = (float)((double)rsj + 1 / s2);
        if (s2 < (double)temp_min_sigma * (double)temp_min_sigma) temp_min_sigma = sigma[j];
    }
    *out_depth = pjsj / rsj;
    *out_sigma = sqrtf(1 / rsj);
    if (min_sigma) *min_sigma = temp_min_sigma;
}

/* InverseDepthHypothesisFusion (:978-1009); n <= 32 */
int oracle_fusion(const float* depth, const float* sigma, int n, const oracle_params* prm,
                  float* out_depth, float* out_sigma)
{
    *out_depth = 0;
    *out_sigma = 0;
    int best_idx[32], best_n = 0;
    for (int a = 0; a < n; a++) {
        int idx[32], cnt = 0;
        for (int b = 0; b < n; b++) {
            if (a == b) { idx[cnt++] = b; continue; }
            if (chi_test(depth[a], depth[b], sigma[a], sigma[b], prm)) idx[cnt++] = b;
        }
        if (best_n < cnt) { best_n = cnt; memcpy(best_idx, idx, sizeof(int) * cnt); }
    }
    if (best_n > prm->lambdaN) {
        float dd[32], ss[32];
        for (int i = 0; i < best_n; i++) { dd[i] = depth[best_idx[i]]; ss[i] = sigma[best_idx[i]]; }
        get_fusion(dd, ss, best_n, out_depth, out_sigma, NULL);
        return 1;
    }
    return 0;
}

/* IntraKeyFrameDepthChecking(cv::Mat&, cv::Mat&, cv::Mat) (:866-927) */
void oracle_intra_check(float* depth_map, float* depth_sigma, int W, int H, const oracle_params* prm)
{
    size_t P = (size_t)W * H;
    float* dnew = (float*)malloc(P * sizeof(float));
    float* snew = (float*)malloc(P * sizeof(float));
    memcpy(dnew, depth_map, P * sizeof(float));
    memcpy(snew, depth_sigma, P * sizeof(float));
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int py = 2; py < (H - 2); py++) {
        for (int px = 2; px < (W - 2); px++) {
            if ((double)depth_map[py * W + px] > prm->eps) {
                float dd[9], ss[9];
                int n = 0;
                float da = depth_map[py * W + px], sa = depth_sigma[py * W + px];
                for (int y = py - 1; y <= py + 1; y++)
                    for (int x = px - 1; x <= px + 1; x++) {
                        if (x == px && y == py) continue;
                        if ((double)depth_map[y * W + x] > prm->eps) {
                            if (chi_test(depth_map[y * W + x], da, depth_sigma[y * W + x], sa, prm)) {
                                dd[n] = depth_map[y * W + x];
                                ss[n] = depth_sigma[y * W + x];
                                n++;
                            }
                        }
                    }
                dd[n] = da; ss[n] = sa; n++; /* "dont forget itself" :902 */
                if (n >= 3) {
                    float fd, fs, ms = 0;
                    get_fusion(dd, ss, n, &fd, &fs, &ms);
                    dnew[py * W + px] = fd;
                    snew[py * W + px] = ms;
                } else {
                    dnew[py * W + px] = 0.0f;
                    snew[py * W + px] = 0.0f;
                }
            }
        }
    }
    memcpy(depth_map, dnew, P * sizeof(float));
    memcpy(depth_sigma, snew, P * sizeof(float));
    free(dnew);
    free(snew);
}

/* IntraKeyFrameDepthGrowing (:929-976) with GetFusion(vector<pair>) (:1647-1667) */
void oracle_intra_grow(float* depth_map, float* depth_sigma, const float* grad, int W, int H,
                       const oracle_params* prm)
{
    size_t P = (size_t)W * H;
    float* dnew = (float*)malloc(P * sizeof(float));
    float* snew = (float*)malloc(P * sizeof(float));
    memcpy(dnew, depth_map, P * sizeof(float));
    memcpy(snew, depth_sigma, P * sizeof(float));
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int py = 2; py < (H - 2); py++) {
        for (int px = 2; px < (W - 2); px++) {
            if ((double)depth_map[py * W + px] < prm->eps) {
                if (grad[py * W + px] <= prm->lambdaG) continue;
                float dd[8], ss[8];
                int n = 0;
                for (int y = py - 1; y <= py + 1; y++)
                    for (int x = px - 1; x <= px + 1; x++) {
                        if (x == px && y == py) continue;
                        if (chi_test(depth_map[y * W + x], depth_map[py * W + px], depth_sigma[y * W + x],
                                     depth_sigma[py * W + px], prm)) {
                            dd[n] = depth_map[y * W + x];
                            ss[n] = depth_sigma[y * W + x];
                            n++;
                        }
                    }
                if (n >= 2) {
                    float pjsj = 0, rsj = 0, min_sigma = ss[0];
                    for (int i = 0; i < n; i++) {
                        double s2 = (double)ss[i] * (double)ss[i];
                        pjsj = (float)((double)pjsj + (double)dd[i] / s2);
                        rsj = (float)((double)rsj + 1 / s2);
                        if (ss[i] < min_sigma) min_sigma = ss[i];
                    }
                    dnew[py * W + px] = pjsj / rsj;
                    snew[py * W + px] = min_sigma;
                }
            }
        }
    }
    memcpy(depth_map, dnew, P * sizeof(float));
    memcpy(depth_sigma, snew, P * sizeof(float));
    free(dnew);
    free(snew);
}

/* per-pair raw hypotheses for every candidate pixel (granularity of one EpipolarSearch call) */
void oracle_pass1_pair(const oracle_kf* kf1, const oracle_kf* kf2, float rot, float min_depth,
                       float max_depth, const oracle_params* prm,
                       float* hyp_depth, float* hyp_sigma, float* hyp_u, uint8_t* hyp_ok)
{
    oracle_pair pr;
    oracle_pair_geometry(kf1, kf2, &pr);
    const int W = kf1->W, H = kf1->H;
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            size_t i = (size_t)y * W + x;
            hyp_depth[i] = 0; hyp_sigma[i] = 0; hyp_u[i] = 0; hyp_ok[i] = 0;
            if (kf1->edge && kf1->edge[i] < 0) continue;
            if (kf1->grad[i] <= prm->lambdaG) continue;
            float pixel = (float)kf1->im[i];
            float d = 0, s = 0, bu = 0, bv = 0;
            int ok = oracle_epipolar_search(kf1, kf2, &pr, x, y, pixel, min_depth, max_depth,
                                            kf1->theta[i], rot, prm, &d, &s, &bu, &bv, NULL);
            if (ok) {
                hyp_depth[i] = d; hyp_sigma[i] = s; hyp_u[i] = bu;
                hyp_ok[i] = (1 / d > 0.0) ? 2 : 1; /* 2 = also passes the keep test of :472 */
            }
        }
    }
}

static void stats_add(oracle_stats* dst, const oracle_stats* src)
{
    dst->candidates += src->candidates;
    dst->scanned += src->scanned;
    dst->evaluated += src->evaluated;
    dst->hypotheses += src->hypotheses;
    dst->fused += src->fused;
    dst->checked += src->checked;
}

/* hot loop 1 (:447-489) for one keyframe */
void oracle_pass1_kf(oracle_kf* kf, int n_nbr, const oracle_kf* const* nbrs, const float* rot,
                     float min_depth, float max_depth, const oracle_params* prm, oracle_stats* st)
{
    oracle_pair* pairs = (oracle_pair*)malloc(sizeof(oracle_pair) * (size_t)n_nbr);
    for (int j = 0; j < n_nbr; j++) oracle_pair_geometry(kf, nbrs[j], &pairs[j]);
    const int W = kf->W, H = kf->H;

#pragma omp parallel
    {
        oracle_stats loc;
        memset(&loc, 0, sizeof(loc));
#pragma omp for schedule(dynamic) collapse(2)
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < W; x++) {
                size_t i = (size_t)y * W + x;
                if (kf->edge && kf->edge[i] < 0) continue;
                if (kf->grad[i] <= prm->lambdaG) continue;
                loc.candidates++;
                float pixel = (float)kf->im[i];
                float hd[32], hs[32];
                int nh = 0;
                for (int j = 0; j < n_nbr; j++) {
                    float d = 0, s = 0, bu = 0, bv = 0;
                    int ok = oracle_epipolar_search(kf, nbrs[j], &pairs[j], x, y, pixel, min_depth, max_depth,
                                                    kf->theta[i], rot[j], prm, &d, &s, &bu, &bv, &loc);
                    if (ok && 1 / d > 0.0) {
                        hd[nh] = d; hs[nh] = s; nh++;
                    }
                }
                loc.hypotheses += nh;
                if (nh > prm->lambdaN) {
                    float fd, fs;
                    if (oracle_fusion(hd, hs, nh, prm, &fd, &fs)) {
                        kf->depth[i] = fd;
                        kf->sigma[i] = fs;
                        loc.fused++;
                    }
                }
            }
        }
        if (st) {
#pragma omp critical
            stats_add(st, &loc);
        }
    }
    free(pairs);

    /* optional stages, placed where the commented calls sit (:491-494) */
    if (prm->intra_check) oracle_intra_check(kf->depth, kf->sigma, W, H, prm);
    if (prm->intra_grow) oracle_intra_grow(kf->depth, kf->sigma, kf->grad, W, H, prm);
}

/* InterKeyFrameDepthChecking(KeyFrame*, vector<KeyFrame*>) (:1121-1296) */
void oracle_inter_check(oracle_kf* kf, int n_nbr, const oracle_kf* const* nbrs,
                        const oracle_params* prm, oracle_stats* st)
{
    oracle_pair* pairs = (oracle_pair*)malloc(sizeof(oracle_pair) * (size_t)n_nbr);
    for (int j = 0; j < n_nbr; j++) oracle_pair_geometry(kf, nbrs[j], &pairs[j]);

    const int cols = kf->W, rows = kf->H;
    const float fx = kf->fx, fy = kf->fy, cx = kf->cx, cy = kf->cy;
    long long nchecked = 0;

#pragma omp parallel for schedule(dynamic) collapse(2) reduction(+ : nchecked)
    for (int py = 2; py < rows - 2; py++) {
        for (int px = 2; px < cols - 2; px++) {
            size_t pi = (size_t)py * cols + px;
            if ((double)kf->depth[pi] < prm->eps) {
                kf->checked[pi] = 0.0f;
                continue;
            }
            float depthp = kf->depth[pi];
            int compatible_neighbor_keyframes_count = 0;
            float cd[32 * 4], cs[32 * 4];
            int cj[32 * 4];
            int num_compatible_pixels = 0;
            float xp[3] = {(px - cx) / fx, (py - cy) / fy, 1.0f};

            for (int j = 0; j < n_nbr; j++) {
                const oracle_kf* pKFj = nbrs[j];
                const oracle_pair* pr = &pairs[j];
                float K[9] = {pKFj->fx, 0, pKFj->cx, 0, pKFj->fy, pKFj->cy, 0, 0, 1};
                float temp[3], Xj[3];
                ocv_mul33_vec(pr->R21, xp, 1.0 * (1. / (double)depthp), pr->t21, 1.0, temp);
                ocv_mul33_vec(K, temp, 1.0, NULL, 0.0, Xj);
                float iz = (float)(1. / (double)Xj[2]); /* Mat / s -> convertTo with (float)scale */
                float xj = Xj[0] * iz;
                float yj = Xj[1] * iz;

                float denom1 = ocv_dot3_d(&pr->R21[6], xp, 1.0);
                float denom2 = depthp * pr->t21[2];
                float depthj = depthp / (denom1 + denom2);

                if (xj < 0 || xj >= cols - 1 || yj < 0 || yj >= rows - 1) continue;
                if (xj != xj || yj != yj) continue; /* NaN: the reference would index out of bounds */
                int x0 = (int)floorf(xj);
                int y0 = (int)floorf(yj);
                int x1 = x0 + 1;
                int y1 = y0 + 1;
                const int ys[4] = {y0, y1, y0, y1};
                const int xs[4] = {x0, x0, x1, x1};
                int nj = 0;
                for (int q = 0; q < 4; q++) {
                    float d = pKFj->depth[(size_t)ys[q] * cols + xs[q]];
                    float sg = pKFj->sigma[(size_t)ys[q] * cols + xs[q]];
                    if ((double)d > prm->eps) {
                        double dd = (double)(depthj - d);
                        double s2 = (double)sg * (double)sg;
                        float test = (float)((dd * dd) / s2);
                        if ((double)test < prm->chi2_inter) {
                            cd[num_compatible_pixels + nj] = d;
                            cs[num_compatible_pixels + nj] = sg;
                            cj[num_compatible_pixels + nj] = j;
                            nj++;
                        }
                    }
                }
                if (nj >= 1) compatible_neighbor_keyframes_count++;
                num_compatible_pixels += nj;
            }

            if (compatible_neighbor_keyframes_count < prm->lambdaN) {
                kf->checked[pi] = 0.0f;
            } else {
                float dp = 1 / depthp;
                float J[32 * 4], r0[32 * 4];
                for (int n = 0; n < num_compatible_pixels; n++) {
                    const oracle_pair* pr = &pairs[cj[n]];
                    float rz = ocv_dot3_d(&pr->R21[6], xp, 1.0);
                    float djn = 1 / cd[n];
                    float sigmajn = cs[n];
                    float d2sigma = djn * djn * sigmajn;
                    J[n] = -rz / d2sigma;
                    r0[n] = (djn - dp * rz - pr->t21[2]) / d2sigma;
                }
                float Jtr0 = ocv_dotn_d(J, r0, num_compatible_pixels, -1.0);
                float JtJ = ocv_dotn_d(J, J, num_compatible_pixels, 1.0);
                float dpDelta = Jtr0 / JtJ;
                float v = 1 / (dp + dpDelta);
                kf->checked[pi] = v;
                if (v > 0) nchecked++;
            }
        }
    }
    if (st) st->checked += nchecked;
    free(pairs);
}

/* UpdateSemiDensePointSet (:700-731) */
void oracle_update_points(oracle_kf* kf, const oracle_params* prm)
{
    float Twc[16];
    oracle_pose_inverse(kf->Tcw, Twc);
    const int cols = kf->W, rows = kf->H;
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int y = 2; y < rows - 2; y++) {
        for (int x = 2; x < cols - 2; x++) {
            size_t i = (size_t)y * cols + x;
            if ((double)kf->checked[i] < prm->eps) {
                kf->points[3 * i + 0] = 0.0f;
                kf->points[3 * i + 1] = 0.0f;
                kf->points[3 * i + 2] = 0.0f;
                continue;
            }
            float inv_d = kf->checked[i];
            float Z = 1 / inv_d;
            float X = Z * (x - kf->cx) / kf->fx;
            float Y = Z * (y - kf->cy) / kf->fy;
            float Pc[4] = {X, Y, Z, 1}, pos[4];
            ocv_mul44_vec(Twc, Pc, pos);
            kf->points[3 * i + 0] = pos[0];
            kf->points[3 * i + 1] = pos[1];
            kf->points[3 * i + 2] = pos[2];
        }
    }
}

/* SemiDenseLoop (:348-597) over keyframes [first, first+count) that already passed the gating of
 * :359 and :365-384 (gating is host logic; see eao-slam_b200/host/).  Keyframes sequential, OpenMP
 * inside, as the reference. */
double oracle_semidense_loop(oracle_kf* kfs, int nkf, int first, int count, int n_nbr,
                             const int32_t* nbr_idx, const float* rot, const float* min_depth,
                             const float* max_depth, const oracle_params* prm, int pass_mask,
                             oracle_stats* st)
{
    (void)nkf;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const oracle_kf** nb = (const oracle_kf**)malloc(sizeof(oracle_kf*) * (size_t)n_nbr);
    if (pass_mask & 1)
        for (int i = first; i < first + count; i++) {
            for (int j = 0; j < n_nbr; j++) nb[j] = &kfs[nbr_idx[(size_t)i * n_nbr + j]];
            oracle_pass1_kf(&kfs[i], n_nbr, nb, rot + (size_t)i * n_nbr, min_depth[i], max_depth[i], prm, st);
        }
    if (pass_mask & 2)
        for (int i = first; i < first + count; i++) {
            for (int j = 0; j < n_nbr; j++) nb[j] = &kfs[nbr_idx[(size_t)i * n_nbr + j]];
            oracle_inter_check(&kfs[i], n_nbr, nb, prm, st);
            oracle_update_points(&kfs[i], prm);
        }
    free(nb);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) / 1e9;
}
