/*
 * sdm_oracle.h — CPU restatement of EAO-SLAM's semi-dense ProbabilityMapping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle and the timed CPU baseline.  Nothing in
 * the product path (eao-slam_b200/, include/) may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS: pinned against the reference's own source, with one stated caveat.
 *   - The reference ships no test, golden vector or fixture for this path and its build needs
 *     OpenCV C++, Eigen, Boost and CGAL, none present here.  Instead, oracle/Makefile (target `ref`) compiles
 *     /root/reference/src/ProbabilityMapping.cc WHERE IT LIES against stand-in headers (oracle/refshim/: a minimal
 *     cv::Mat / MatExpr, ORB_SLAM2::KeyFrame / Map with the reference's member names, no-op LineDetector / Modeler)
 *     into oracle/_ref/libref_pm.so, and tests/test_ref_vs_oracle.py runs the reference's SemiDenseLoop(),
 *     IntraKeyFrameDepthChecking and IntraKeyFrameDepthGrowing on the same keyframes: this file reproduces their
 *     planes bit for bit (0 differing words; also committed as tests/golden/ref_loop_small.npz).
 *   - Caveat: the arithmetic INSIDE cv::Mat expressions is the stand-in's, not OpenCV's.  Those evaluation rules
 *     (small gemm forms, A*B^T, 1xn*nx1, invert, LU solve, convertTo scale, fastAtan2) are pinned bit-exactly
 *     against real cv2 4.13 (tests/golden/cv2_kats.npz, generator oracle/pin_cv2.py), and the composed pair geometry
 *     (R21, t21, F12) against real cv2 calls in MatExpr evaluation order (tests/golden/pair_geometry_cv2.npz).
 *   - The end-to-end outputs of this file on a small scene are committed (tests/golden/oracle_small.npz) so the
 *     oracle cannot drift silently.
 * cpu_baseline.kind stays "port": timing the reference text through a stand-in cv::Mat would not be the reference's
 * real cost (OpenCV's allocator and locks dominate it).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef SDM_ORACLE_H
#define SDM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* #defines of include/ProbabilityMapping.h:45-56 and the literals in src/ProbabilityMapping.cc */
typedef struct {
    int lambdaG;        /* 8   :48  gradient-magnitude threshold                     */
    int lambdaL;        /* 80  :49  epipolar-line / gradient direction threshold     */
    int lambdaTheta;    /* 45  :50  gradient orientation consistency threshold       */
    int lambdaN;        /* 3   :51  hypothesis / neighbour-support count             */
    float theta;        /* (float)0.23 :55 THETA                                     */
    float sigmaI;       /* 20  :47  I_stddev (KeyFrame.cc:65)                        */
    double chi2_fusion; /* 5.99  ProbabilityMapping.cc:1638,1644                     */
    double chi2_inter;  /* 3.84  :1207                                               */
    double eps;         /* 0.000001 :877,1156,1204                                   */
    float slope_max;    /* 4   :757                                                  */
    int intra_check;    /* 0 = shipped loop (:491-494 commented out), 1 = enabled    */
    int intra_grow;     /* same, for IntraKeyFrameDepthGrowing                       */
} oracle_params;

void oracle_default_params(oracle_params* p);

/* One keyframe: inputs (KeyFrame.h:155-175) and the planes the path writes.
 * All planes are dense row-major W*H (points: W*H*3, (y, 3x+c)). */
typedef struct {
    int W, H;
    const uint8_t* im;    /* im_       CV_8U   */
    const float* grad;    /* GradImg   CV_32F  */
    const float* theta;   /* GradTheta CV_32F, degrees [0,360) */
    const int32_t* edge;  /* mEdgeIndex CV_32S or NULL (= every pixel passes :454) */
    float fx, fy, cx, cy;
    float Tcw[12];        /* rows 0..2 of the 4x4 pose, row-major [R|t] */
    float* depth;         /* depth_map_         (inverse depth) */
    float* sigma;         /* depth_sigma_                        */
    float* checked;       /* depth_map_checked_                  */
    float* points;        /* SemiDensePointSets_ (CV_32FC3)      */
} oracle_kf;

typedef struct {
    float R21[9];
    float t21[3];
    float F12[9];
} oracle_pair;

typedef struct {
    long long candidates;      /* pixels passing :454-456                      */
    long long scanned;         /* uj iterations of :770                        */
    long long evaluated;       /* uj iterations reaching the err computation   */
    long long hypotheses;      /* accepted per-pair hypotheses (:472)          */
    long long fused;           /* pixels written at :483                       */
    long long checked;         /* pixels with depth_map_checked_ > 0           */
} oracle_stats;

/* ---- OpenCV primitives (exposed for the cv2 known-answer tests) ---- */
float ocv_fastAtan2(float y, float x);
void ocv_mul33_ABt(const float* A, const float* B, double alpha, float* D);   /* gemm GEMM_2_T, double acc */
void ocv_mul33(const float* A, const float* B, float* D);                     /* 3x3*3x3 float dots */
void ocv_mul33_vec(const float* A, const float* x, double alpha, const float* c, double beta, float* d);
float ocv_dot3_d(const float* a, const float* x, double alpha);               /* 1x3 * 3x1, double acc */
void ocv_inv33(const float* A, float* D);                                     /* cv::invert 3x3 DECOMP_LU */
int ocv_solve33_lu(const float* A, const float* B, float* X);                 /* cv::solve 3x3, 3x3 RHS */
float ocv_dotn_d(const float* a, const float* b, int n, double alpha);        /* 1xn * nx1, double acc */
void ocv_mul44_vec(const float* A, const float* x, float* d);                 /* 4x4*4x1 float dots */

/* ---- path functions ---- */
void oracle_pose_inverse(const float* Tcw, float* Twc16);                     /* KeyFrame.cc:108-124 */
void oracle_pair_geometry(const oracle_kf* kf1, const oracle_kf* kf2, oracle_pair* out);
void oracle_stereo_search_constraints(const float* inv_depths, int n, float* min_depth, float* max_depth);
void oracle_get_search_range(const oracle_kf* kf1, const oracle_pair* pr, int px, int py,
                             float mind, float maxd, float* umin, float* umax);
/* returns 1 if dh was produced (supported), fills depth/sigma/best_u/best_v */
int oracle_epipolar_search(const oracle_kf* kf1, const oracle_kf* kf2, const oracle_pair* pr,
                           int x, int y, float pixel, float min_depth, float max_depth,
                           float th_pi, float rot, const oracle_params* prm,
                           float* depth, float* sigma, float* best_u, float* best_v,
                           oracle_stats* st);
int oracle_fusion(const float* depth, const float* sigma, int n, const oracle_params* prm,
                  float* out_depth, float* out_sigma);
void oracle_intra_check(float* depth, float* sigma, int W, int H, const oracle_params* prm);
void oracle_intra_grow(float* depth, float* sigma, const float* grad, int W, int H, const oracle_params* prm);

/* per-pair raw hypotheses over the whole image (debug / parity granularity of EpipolarSearch) */
void oracle_pass1_pair(const oracle_kf* kf1, const oracle_kf* kf2, float rot, float min_depth,
                       float max_depth, const oracle_params* prm,
                       float* hyp_depth, float* hyp_sigma, float* hyp_u, uint8_t* hyp_ok);

/* hot loop 1 (:447-489) for one keyframe; nbrs in neighbour order */
void oracle_pass1_kf(oracle_kf* kf, int n_nbr, const oracle_kf* const* nbrs, const float* rot,
                     float min_depth, float max_depth, const oracle_params* prm, oracle_stats* st);
/* hot loop 2 (:1121-1296) + UpdateSemiDensePointSet (:700-731) */
void oracle_inter_check(oracle_kf* kf, int n_nbr, const oracle_kf* const* nbrs,
                        const oracle_params* prm, oracle_stats* st);
void oracle_update_points(oracle_kf* kf, const oracle_params* prm);

/* SemiDenseLoop (:348-597) over pre-gated keyframes: pass 1 for all, then pass 2 for all.
 * nbr_idx: [nkf][n_nbr] indices into kfs; rot: [nkf][n_nbr]; min/max_depth: [nkf].
 * pass_mask bit0 = pass 1, bit1 = pass 2. Returns elapsed seconds (CLOCK_MONOTONIC, :246-254). */
double oracle_semidense_loop(oracle_kf* kfs, int nkf, int first, int count, int n_nbr,
                             const int32_t* nbr_idx, const float* rot, const float* min_depth,
                             const float* max_depth, const oracle_params* prm, int pass_mask,
                             oracle_stats* st);

int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
