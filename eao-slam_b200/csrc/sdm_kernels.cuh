// sdm_kernels.cuh — sm_100a kernels of the semi-dense mapping path.
//
// Reference being replaced: yanmin-wu/EAO-SLAM src/ProbabilityMapping.cc (line numbers cited per
// function).  The arithmetic follows the reference operation by operation (float where the
// reference is float, double where OpenCV accumulates in double); this translation unit MUST be
// compiled with -fmad=false and default (IEEE) division / sqrt so that results are bit-identical
// to a non-contracting CPU evaluation.  No tensor cores: the path is a gather + reduction.
//
// Data layout in HBM (per keyframe slot, P = W*H):
//   tex    float4[P]  row-pair texel {G(y,x), G(y+1,x), Th(y,x), Th(y+1,x)}  (row H-1 pairs with itself)
//   ipair  uchar2[P]  {I(y,x), I(y+1,x)}
//   texw   float4[P]  the texel the second-generation scan loop reads: {G(y,x), G(y+1,x), Th'(y,x), Th'(y+1,x)} with the
//                     orientation pair pre-processed for yangle (:83-111): where |Th0 - Th1| >= 180 the smaller one
//                     already carries its + 360 and BOTH are stored negated (the sign is the "wrapped" flag)
//   cand   u32[P]     compacted candidate pixels (y<<16 | x), count in cand_count[slot]
//   rs     float2[P]  {rho, sigma} = depth_map_, depth_sigma_   (pass-1 output; the plane neighbours and peers read)
//   rs2    float2[P]  second plane of the intra check / grow ping-pong (only with the intra stage)
//   dpl    float[P]   } dense copies of rho / sigma for the D2H path
//   spl    float[P]   }
//   chk    float[P]   depth_map_checked_                          (pass-2 output)
//   pts    float[3P]  SemiDensePointSets_                         (pass-2 output)
// The row-pair texel makes every ylinear/yangle evaluation of the epipolar scan (:66-111) one
// 16-byte load + one 2-byte load, coalesced across the warp's consecutive columns.
// Kernels: k_pack, k_plan, k_pass1_lane (default scan) / k_pass1 (warp-per-pixel A/B),
// k_intra_cand / k_intra_check / k_intra_grow, k_pass2_cand / k_pass2, k_export_*, single-method helpers.
#pragma once

#include <cuda.h>  // CUtensorMap (type only: the encode entry point is fetched through the runtime, no -lcuda)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdm_b200.h"

namespace sdm {

struct DevParams {
    int W, H;
    float lambdaG, lambdaL, lambdaTheta;
    int lambdaN;
    float theta;      // (float)0.23
    float inv_theta;  // 1 / (float)0.23 (float division)
    float var_num;    // (2 * sigmaI) * sigmaI
    float chi_fusion_lt;  // x < this  <=>  (double)x < 5.99
    float chi_inter_lt;   // x < this  <=>  (double)x < 3.84
    float chi_inter_lo, chi_inter_hi;  // chi_inter_lt * (1 -+ 2^-16): the float shortcut of chi_inter_accept
    float eps_gt;         // x > this  <=>  (double)x > 0.000001
    float eps_lt;         // x < this  <=>  (double)x < 0.000001
    float slope_max;
    int scan2;        // scan loop generation: 0 = first (any thresholds), 2 = second (reference thresholds + verified
                      // reciprocal division), 3 = third (second + skip-distance planes, scan_columns3)
};

constexpr int kSkipBins = 8;  // orientation bins of the skip-distance planes (k_skip, scan_columns3)

struct DevArena {
    float4* tex;
    uchar2* ipair;
    float4* texw;  // scan_columns2's texel: as tex, orientation pair wrap-encoded (encode_theta_pair)
    uint32_t* cand;
    int* cand_count;
    int* plane_irregular;  // per slot: != 0 if an uploaded plane breaks scan_columns2's preconditions (k_pack)
    float2* rs;
    float* chk;
    float* pts;
    float* dpl;  // dense depth_map_ plane   } copies of rs.x / rs.y kept for the D2H path, so a download is
    float* spl;  // dense depth_sigma_ plane } pure DMA (no de-interleave kernel competing for SMs)
    float2* rs2;  // second (rho,sigma) plane of the intra ping-pong: zero outside the candidate pixels
    uint32_t* blk;  // [slot][H][blk_words] bit b of a row: the 16-pixel block b holds a candidate pixel (k_pack; k_sparse_rows)
    int blk_words;  // words per row = ceil(ceil(W / 16) / 32)
    uint8_t* skip;  // [slot][kSkipBins][P] skip distances of the third-generation scan loop (k_skip); nullptr = not built
    size_t P;  // pixels per plane
};

struct DevPair {
    float F[9];
    float R[9];
    float t[3];
    float rot;
    int slot;
    float K2[4];  // the neighbour's own fx fy cx cy: InterKeyFrameDepthChecking projects with pKFj->GetCalibrationMatrix() (:1172)
};  // 27 words

struct DevItem {
    int kf;
    int n_nbr;
    float min_depth, max_depth;
    float K[4];     // fx fy cx cy
    float Twc[12];  // rows 0..2 of the inverse pose
    DevPair pair[SDM_MAX_NBR];
};
static_assert(sizeof(DevItem) % 4 == 0, "word copy");

struct DevStats {
    unsigned long long candidates, fused, checked;
};

#define SDM_FULL 0xffffffffu

// constants of the second-generation scan loop (scan_columns2)
constexpr float kTheta2 = 0.23f, kRTheta2 = 1.0f / 0.23f, kLambdaG2 = 8.0f, kGradMax2 = 0x1p40f;

// ---------------------------------------------------------------------------------------------
// cv::fastAtan2 (used at :791), float polynomial in degrees
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = 2.220446049250313e-16f;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------
// interpolators on the packed planes (ylinear<float>, ylinear<uchar>, yangle<float>; :66-111)
// ---------------------------------------------------------------------------------------------
struct RowW {
    int y0;
    float w0, w1;
};
__device__ __forceinline__ RowW row_weights(float v)
{
    RowW r;
    float fl = floorf(v);
    r.y0 = (int)fl;
    r.w0 = (fl + 1.0f) - v;  // y1 - y  (y1 = y0 + 1 converted to float: exact)
    r.w1 = v - fl;           // y - y0
    return r;
}
__device__ __forceinline__ float ylin_grad(const float4* __restrict__ tex, int W, float v, int u)
{
    RowW r = row_weights(v);
    float4 t = __ldg(&tex[(size_t)r.y0 * W + u]);
    return t.x * r.w0 + t.y * r.w1;
}
__device__ __forceinline__ float ylin_im(const uchar2* __restrict__ ip, int W, float v, int u)
{
    RowW r = row_weights(v);
    uchar2 i2 = __ldg(&ip[(size_t)r.y0 * W + u]);
    return (float)i2.x * r.w0 + (float)i2.y * r.w1;
}
__device__ __forceinline__ float yangle_interp(float a0, float a1, float w0, float w1)
{
    if (fabsf(a0 - a1) < 180.f) {
        return a0 * w0 + a1 * w1;
    } else {
        if (a0 < a1) a0 += 360.f; else a1 += 360.f;
        float inter = a0 * w0 + a1 * w1;
        if (inter >= 360.f) inter -= 360.f;
        return inter;
    }
}

// ---------------------------------------------------------------------------------------------
// per-(pixel, neighbour) set-up: epipolar line (:753-757), GetSearchRange (:1598-1631), the
// loop-invariant fastAtan2 of :791 and the wrapped th_pi + rot of :801-803
// ---------------------------------------------------------------------------------------------
struct PairSetup {
    float ab, cb;      // a/b, c/b
    int u_lo, u_hi;    // scan bounds, already restricted to 1 <= u <= W-2 (:778-779)
    float th_line;
    float ang_pi_rot;
    bool valid;
};

__device__ __forceinline__ float wrap360(float a)
{
    if (a >= 360.f) a -= 360.f;
    if (a < 0.f) a += 360.f;
    return a;
}

__device__ __forceinline__ void search_range(const DevPair& g, const float* K, int W, float xn, float yn,
                                             float mind, float maxd, float& umin, float& umax)
{
    const float fx = K[0], cx = K[2];
    // R21*xp1*d + t21 : one gemm, float dot then double scale-add (only rows 0 and 2 are used)
    float s0 = g.R[0] * xn + g.R[1] * yn + g.R[2] * 1.0f;
    float s2 = g.R[6] * xn + g.R[7] * yn + g.R[8] * 1.0f;
    float x0min = (float)((double)s0 * (double)mind + (double)g.t[0]);
    float x2min = (float)((double)s2 * (double)mind + (double)g.t[2]);
    float x0max = (float)((double)s0 * (double)maxd + (double)g.t[0]);
    float x2max = (float)((double)s2 * (double)maxd + (double)g.t[2]);
    umin = fx * x0min / x2min + cx;
    umax = fx * x0max / x2max + cx;
    if (umin > umax) { float t = umax; umax = umin; umin = t; }
    if (umin < 0) umin = 0;
    if (umax < 0) umax = 0;
    if (umin > W) umin = W - 1;
    if (umax > W) umax = W - 1;
}

__device__ __forceinline__ PairSetup pair_setup(const DevPair& g, const float* K, const DevParams& P, int x, int y,
                                                float xn, float yn, float mind, float maxd, float th_pi)
{
    PairSetup s;
    float a = x * g.F[0] + y * g.F[3] + g.F[6];
    float b = x * g.F[1] + y * g.F[4] + g.F[7];
    float c = x * g.F[2] + y * g.F[5] + g.F[8];
    s.ab = a / b;
    s.cb = c / b;
    s.valid = (s.ab >= -P.slope_max) && (s.ab <= P.slope_max);  // NaN -> invalid
    float umin, umax;
    search_range(g, K, P.W, xn, yn, mind, maxd, umin, umax);
    s.valid = s.valid && (umin == umin) && (umax == umax);
    int lo = (int)ceilf(umin), hi = (int)floorf(umax);
    s.u_lo = lo < 1 ? 1 : lo;                  // uj - 1 < mnMinX -> continue
    s.u_hi = hi > P.W - 2 ? P.W - 2 : hi;      // uj + 1 >= mnMaxX -> continue
    if (!s.valid) { s.u_lo = 1; s.u_hi = 0; }
    s.th_line = fast_atan2_deg(-s.ab, 1.0f);
    s.ang_pi_rot = wrap360(th_pi + g.rot);
    return s;
}

// one candidate column of the scan body (:772-821).  Returns false if the candidate is skipped.
// cand_rows: the row tests of :773-785 and the interpolation weights; cand_gates: conditions 1-3 and the residual on the
// column's texel (fetched by the caller: global memory, or a shared-memory tile staged by TMA in k_pass1_tma).
__device__ __forceinline__ bool cand_rows(const DevParams& P, float Hm1, int u, float ab, float cb, RowW& r, size_t& idx)
{
    const float uf = (float)u;
    const float v = -(ab * uf + cb);
    const float vp = -(ab * (uf + 1.0f) + cb);
    const float vm = -(ab * (uf - 1.0f) + cb);
    // floor(v) < 0 || ceil(v) >= rows  <=>  !(0 <= v <= rows-1), same for vj_plus / vj_minus
    if (!(v >= 0.f && v <= Hm1 && vp >= 0.f && vp <= Hm1 && vm >= 0.f && vm <= Hm1)) return false;
    r = row_weights(v);
    idx = (size_t)r.y0 * P.W + u;
    return true;
}
__device__ __forceinline__ bool cand_gates(const float4 t, const RowW& r, const uchar2* __restrict__ ip2, size_t idx,
                                           const DevParams& P, float th_line, float ang_pi_rot, float pixel, float gradc,
                                           float& err, float& pe, float& ge)
{
    const float g2 = t.x * r.w0 + t.y * r.w1;
    if (g2 <= P.lambdaG) return false;  // condition 1
    const float gth = yangle_interp(t.z, t.w, r.w0, r.w1);
    float ang = gth - th_line;  // condition 2
    if (ang >= 360.f) ang -= 360.f;
    if (ang < 0.f) ang += 360.f;
    if (ang > 180.f) ang = 360.f - ang;
    if (ang > 90.f) ang = 180.f - ang;
    if (ang >= P.lambdaL) return false;
    float thd = gth - ang_pi_rot;  // condition 3
    if (thd >= 360.f) thd -= 360.f;
    if (thd < 0.f) thd += 360.f;
    if (thd > 180.f) thd = 360.f - thd;
    if (thd >= P.lambdaTheta) return false;
    const uchar2 i2 = __ldg(&ip2[idx]);
    pe = pixel - ((float)i2.x * r.w0 + (float)i2.y * r.w1);
    ge = gradc - g2;
    err = pe * pe + (ge * ge) / P.theta;
    return true;
}
__device__ __forceinline__ bool eval_candidate(const float4* __restrict__ tex2, const uchar2* __restrict__ ip2,
                                               const DevParams& P, float Hm1, int u, float ab, float cb,
                                               float th_line, float ang_pi_rot, float pixel, float gradc,
                                               float& err, float& pe, float& ge)
{
    RowW r;
    size_t idx;
    if (!cand_rows(P, Hm1, u, ab, cb, r, idx)) return false;
    return cand_gates(__ldg(&tex2[idx]), r, ip2, idx, P, th_line, ang_pi_rot, pixel, gradc, err, pe, ge);
}

// GetPixelDepth, equation 8 (:1568-1596).  s2d/s0d are the double-accumulated row products.
__device__ __forceinline__ float pixel_inv_depth(float u, double s2d, double s0d, const DevPair& g, const float* K)
{
    const float fx = K[0], cx = K[2];
    float ucx = u - cx;
    float num1 = (float)(s2d * (double)ucx);
    float num2 = (float)(s0d * (double)fx);
    float denom1 = -g.t[2] * ucx;
    float denom2 = fx * g.t[0];
    return (num1 - num2) / (denom1 + denom2);
}

// sub-pixel refinement (:823-840) + ComputeInvDepthHypothesis (:1310-1335) at the best column
struct Hypo {
    float depth, sigma, best_u, best_v;
};
template <class TG, class TI>
__device__ __forceinline__ Hypo refine_hypothesis(const TG* __restrict__ tex2, const TI* __restrict__ ip2,
                                                  const DevPair& g, const float* K, const DevParams& P, int best,
                                                  float ab, float cb, float pe, float ge, float xn, float yn)
{
    const int up = best + 1, um = best - 1;
    const float vp = -(ab * (float)up + cb);
    const float vm = -(ab * (float)um + cb);
    const float gI = (ylin_im(ip2, P.W, vp, up) - ylin_im(ip2, P.W, vm, um)) / 2;
    const float q = (ylin_grad(tex2, P.W, vp, up) - ylin_grad(tex2, P.W, vm, um)) / 2;
    const float den = gI * gI + P.inv_theta * q * q;
    const float ustar = (float)best + (gI * pe + P.inv_theta * q * ge) / den;
    const float var = P.var_num / den;
    Hypo h;
    h.best_u = ustar;
    h.best_v = -(ab * ustar + cb);
    // R21.row(k) * xp : 1x3 * 3x1 gemm -> double accumulation
    const double s2d = (double)g.R[6] * (double)xn + (double)g.R[7] * (double)yn + (double)g.R[8] * 1.0;
    const double s0d = (double)g.R[0] * (double)xn + (double)g.R[1] * (double)yn + (double)g.R[2] * 1.0;
    const float sd = sqrtf(var);
    const float rho = pixel_inv_depth(ustar, s2d, s0d, g, K);
    const float rho_min = pixel_inv_depth(ustar - sd, s2d, s0d, g, K);
    const float rho_max = pixel_inv_depth(ustar + sd, s2d, s0d, g, K);
    const float s1 = fabsf(rho_max - rho), s2 = fabsf(rho_min - rho);
    h.depth = rho;
    h.sigma = (s1 < s2) ? s2 : s1;  // cv::max == std::max
    return h;
}

// ChiTest (:1633-1645)
__device__ __forceinline__ bool chi_compatible(float a, float b, float sa, float sb, float thr_lt)
{
    float num = (a - b) * (a - b);
    float chi = num / (sa * sa) + num / (sb * sb);
    return chi < thr_lt;
}

// one term of GetFusion (:1669-1692): float accumulators updated through double
__device__ __forceinline__ void fusion_accumulate(float d, float s, float& pjsj, float& rsj)
{
    const double s2 = (double)s * (double)s;
    pjsj = (float)((double)pjsj + (double)d / s2);
    rsj = (float)((double)rsj + 1.0 / s2);
}

// Orientation pair of a texel as scan_columns2 wants it.  yangle (:83-111) interpolates a0*w0 + a1*w1 directly when
// |a0 - a1| < 180; otherwise it first adds 360 to the smaller angle and afterwards takes the result modulo 360.  The
// "+ 360" is applied here, once per texel instead of once per scanned column, and both angles are negated so that the
// sign of the interpolated value tells the loop which case it is in (RN products and sums are odd functions: the
// negated pair interpolates to exactly minus the reference's value).  Valid for orientations in [0, 360] (regular
// planes): then every wrapped pair is strictly negative and every other pair is >= +0 or -0.
__device__ __forceinline__ float2 encode_theta_pair(float a0, float a1)
{
    if (fabsf(a0 - a1) < 180.f) return make_float2(a0, a1);
    if (a0 < a1) a0 += 360.f; else a1 += 360.f;
    return make_float2(-a0, -a1);
}

// ---------------------------------------------------------------------------------------------
// K1 + K2: plane packing + candidate compaction (KeyFrame.cc:63-88 planes as device buffers;
// the candidate test of :454-456).  One thread per pixel, 32x8 tiles; one atomic per tile so a
// tile's candidates stay contiguous (2-D locality for the scan kernel).
// ---------------------------------------------------------------------------------------------
#ifndef SDM_TILE_W
#define SDM_TILE_W 32
#endif
constexpr int kTileW = SDM_TILE_W, kTileH = 256 / SDM_TILE_W;

// A batch of keyframes packed by ONE launch of each packing kernel (grid.z = keyframe): per keyframe the slot and the
// staging planes its H2D landed in.  (One launch per keyframe left the stream idle between 5 launches per keyframe:
// 6.9 ms of stream time per 200 keyframes for 1.9 ms of kernel time.)
constexpr int kPackBatch = 24;
struct PackBatch {
    int slot[kPackBatch];
    const uint8_t* im[kPackBatch];
    const float* grad[kPackBatch];   // nullptr: GradImg / GradTheta are produced from im (k_pack_image)
    const float* theta[kPackBatch];
    const int32_t* edge[kPackBatch];  // nullptr: every pixel passes :454
};

// per-slot counters and block masks back to zero before a batch is packed
__global__ void __launch_bounds__(256) k_pack_reset(DevArena A, DevParams P, PackBatch B)
{
    const int slot = B.slot[blockIdx.x];
    if (threadIdx.x == 0) { A.cand_count[slot] = 0; A.plane_irregular[slot] = 0; }
    uint32_t* blk = A.blk + (size_t)slot * P.H * A.blk_words;
    for (int i = threadIdx.x; i < P.H * A.blk_words; i += blockDim.x) blk[i] = 0u;
}

// Which 16-pixel blocks of each row hold a candidate: every output plane of the path is zero outside the candidate
// pixels, so a download into zero-initialised host planes only has to move these blocks (k_sparse_rows).
constexpr int kBlkPx = 16;
__device__ __forceinline__ void mark_blocks(const DevArena& A, const DevParams& P, int slot, int x, int y, bool in, bool is_cand,
                                            unsigned bal, int lane)
{
    if (!A.blk) return;
    uint32_t* row = A.blk + ((size_t)slot * P.H + (in ? y : 0)) * A.blk_words;
    if (kTileW == 32) {  // a warp = 32 consecutive pixels of one row = two blocks: one atomic per warp
        const unsigned two = ((bal & 0xffffu) ? 1u : 0u) | ((bal >> 16) ? 2u : 0u);
        const int b0 = (x - lane) / kBlkPx;  // lane 0's block (x - lane is a multiple of 32)
        if (lane == 0 && two && y < P.H) atomicOr(&row[b0 >> 5], two << (b0 & 31));
    } else if (is_cand) {
        const int b = x / kBlkPx;
        atomicOr(&row[b >> 5], 1u << (b & 31));
    }
}

__global__ void __launch_bounds__(256) k_pack(DevArena A, DevParams P, PackBatch B)
{
    const int slot = B.slot[blockIdx.z];
    const uint8_t* __restrict__ im = B.im[blockIdx.z];
    const float* __restrict__ grad = B.grad[blockIdx.z];
    const float* __restrict__ theta = B.theta[blockIdx.z];
    const int32_t* __restrict__ edge = B.edge[blockIdx.z];
    if (grad == nullptr) return;  // this keyframe belongs to k_pack_image
    __shared__ int s_wcount[8];
    __shared__ int s_base;
    // a block packs one tile of kTileW x kTileH pixels; a warp covers kTileW x (32 / kTileW) of them
    const int warp = threadIdx.y, lane = threadIdx.x;
    const int x = blockIdx.x * kTileW + (lane % kTileW);
    const int y = blockIdx.y * kTileH + warp * (32 / kTileW) + (lane / kTileW);
    const bool in = (x < P.W) && (y < P.H);
    bool is_cand = false, irregular = false;
    const size_t base = (size_t)slot * A.P;
    if (in) {
        const int y1 = (y + 1 < P.H) ? y + 1 : P.H - 1;
        const size_t i0 = (size_t)y * P.W + x, i1 = (size_t)y1 * P.W + x;
        const float g0 = grad[i0], g1 = grad[i1];
        const float th0 = theta[i0];
        irregular = !(th0 >= 0.f && th0 <= 360.f) || !(g0 <= kGradMax2);
        const float th1 = theta[i1];
        const uint8_t p0 = im[i0], p1 = im[i1];
        A.tex[base + i0] = make_float4(g0, g1, th0, th1);
        A.ipair[base + i0] = make_uchar2(p0, p1);
        const float2 tw = encode_theta_pair(th0, th1);
        if (A.texw) A.texw[base + i0] = make_float4(g0, g1, tw.x, tw.y);
        A.rs[base + i0] = make_float2(0.f, 0.f);
        A.dpl[base + i0] = 0.f;
        A.spl[base + i0] = 0.f;
        if (A.rs2) A.rs2[base + i0] = make_float2(0.f, 0.f);
        A.chk[base + i0] = 0.f;  // pass 2 only visits candidate pixels: the rest of the output planes stays 0
        A.pts[3 * (base + i0) + 0] = 0.f;
        A.pts[3 * (base + i0) + 1] = 0.f;
        A.pts[3 * (base + i0) + 2] = 0.f;
        is_cand = (edge == nullptr || edge[i0] >= 0) && !(g0 <= P.lambdaG);
    }
    const unsigned bal = __ballot_sync(SDM_FULL, is_cand);
    mark_blocks(A, P, slot, x, y, in, is_cand, bal, lane);
    // scan_pixel_lane<2> relies on orientations in [0, 360] (what cv::phase produces) and finite gradients; anything
    // else selects the first-generation loop for the keyframes that touch this slot
    if (__any_sync(SDM_FULL, irregular) && lane == 0) atomicOr(&A.plane_irregular[slot], 1);
    if (lane == 0) s_wcount[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0 && lane == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) { int c = s_wcount[w]; s_wcount[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(&A.cand_count[slot], tot) : 0;
    }
    __syncthreads();
    if (is_cand) {
        const int pos = s_base + s_wcount[warp] + __popc(bal & ((1u << lane) - 1u));
        A.cand[base + pos] = ((uint32_t)y << 16) | (uint32_t)x;
    }
}

// ---------------------------------------------------------------------------------------------
// K1', plane producers on the device: GradImg / GradTheta of KeyFrame.cc:69-74 from im_
// alone, then the same packing + candidate compaction as k_pack.  cv::Scharr(.., CV_32F, 1|0, 0|1, 1/32.0)
// with BORDER_REFLECT_101 is exact (integer sums times 2^-5).  magnitude = sqrtf(gx*gx + gy*gy) and
// phase = the scalar cv::fastAtan2(gy, gx): OpenCV's SIMD magnitude / phase differ from these scalar forms by
// <= 1 ulp / 3e-5 deg (and are not reproducible between calls on a multi-threaded host), so the planes
// are as valid as a host run's, not bit-identical to a particular one.  One block = one 32x8 tile; the
// (G, theta) of rows y0 .. y0+8 are staged in shared memory because a texel pairs row y with row y+1.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__device__ __forceinline__ float2 scharr_mag_phase(const uint8_t* __restrict__ im, int W, int H, int x, int y)
{
    const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
    const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
    const uint8_t* r0 = im + (size_t)ym * W;
    const uint8_t* r1 = im + (size_t)y * W;
    const uint8_t* r2 = im + (size_t)yp * W;
    const int a00 = r0[xm], a01 = r0[x], a02 = r0[xp];
    const int a10 = r1[xm], a12 = r1[xp];
    const int a20 = r2[xm], a21 = r2[x], a22 = r2[xp];
    const int igx = 3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20);
    const int igy = 3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02);
    const float gx = (float)igx * 0.03125f, gy = (float)igy * 0.03125f;  // exact
    return make_float2(sqrtf(gx * gx + gy * gy), fast_atan2_deg(gy, gx));
}

__global__ void __launch_bounds__(256) k_pack_image(DevArena A, DevParams P, PackBatch B)
{
    const int slot = B.slot[blockIdx.z];
    const uint8_t* __restrict__ im = B.im[blockIdx.z];
    const int32_t* __restrict__ edge = B.edge[blockIdx.z];
    if (B.grad[blockIdx.z] != nullptr) return;  // this keyframe's planes were uploaded: k_pack
    __shared__ float2 s_gt[9][32];
    __shared__ int s_wcount[8];
    __shared__ int s_base;
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = blockIdx.y * 8;
    const int y = y0 + warp;
    const bool in = (x < P.W) && (y < P.H);
    if (x < P.W) {
        if (y < P.H) s_gt[warp][lane] = scharr_mag_phase(im, P.W, P.H, x, y);
        if (warp == 0) {  // row y0 + 8: the lower half of the last row's texels (clamped to the last image row)
            const int yl = min(y0 + 8, P.H - 1);
            s_gt[8][lane] = scharr_mag_phase(im, P.W, P.H, x, yl);
        }
    }
    __syncthreads();
    bool is_cand = false;
    const size_t base = (size_t)slot * A.P;
    if (in) {
        const int y1 = (y + 1 < P.H) ? y + 1 : P.H - 1;
        const size_t i0 = (size_t)y * P.W + x;
        const float2 g0 = s_gt[warp][lane];
        const float2 g1 = s_gt[y1 - y0][lane];
        const uint8_t p0 = im[i0], p1 = im[(size_t)y1 * P.W + x];
        A.tex[base + i0] = make_float4(g0.x, g1.x, g0.y, g1.y);
        A.ipair[base + i0] = make_uchar2(p0, p1);
        const float2 tw = encode_theta_pair(g0.y, g1.y);
        if (A.texw) A.texw[base + i0] = make_float4(g0.x, g1.x, tw.x, tw.y);
        A.rs[base + i0] = make_float2(0.f, 0.f);
        A.dpl[base + i0] = 0.f;
        A.spl[base + i0] = 0.f;
        if (A.rs2) A.rs2[base + i0] = make_float2(0.f, 0.f);
        A.chk[base + i0] = 0.f;
        A.pts[3 * (base + i0) + 0] = 0.f;
        A.pts[3 * (base + i0) + 1] = 0.f;
        A.pts[3 * (base + i0) + 2] = 0.f;
        is_cand = (edge == nullptr || edge[i0] >= 0) && !(g0.x <= P.lambdaG);
    }
    const unsigned bal = __ballot_sync(SDM_FULL, is_cand);
    mark_blocks(A, P, slot, x, y, in, is_cand, bal, lane);
    if (lane == 0) s_wcount[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0 && lane == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) { int c = s_wcount[w]; s_wcount[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(&A.cand_count[slot], tot) : 0;
    }
    __syncthreads();
    if (is_cand) {
        const int pos = s_base + s_wcount[warp] + __popc(bal & ((1u << lane) - 1u));
        A.cand[base + pos] = ((uint32_t)y << 16) | (uint32_t)x;
    }
}

// GradImg / GradTheta of a slot back as dense planes (what KeyFrame::GradImg / GradTheta would hold)
__global__ void k_split_tex(const float4* __restrict__ tex, float* __restrict__ g, float* __restrict__ t, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 v = tex[i]; g[i] = v.x; t[i] = v.z; }
}

// ---------------------------------------------------------------------------------------------
// Skip-distance planes of the third-generation scan loop (scan_columns3).  For every texel (row pair y / y+1, column x)
// and each of kSkipBins orientation bins q, M_q(y,x) = 1 if a candidate column that lands on this texel MAY survive
// conditions 1 and 3 of the scan body (:788, :801-809) for a pixel whose th_pi + rot lies in bin q:
//   condition 1: fl(G0*w0) + fl(G1*w1) <= 8 whenever G0 <= 8 and G1 <= 8 (w0 + w1 rounds to <= 1: both products are
//                <= 8*w, the sum is <= 8*fl(w0 + w1) = 8), so "G0 > 8 or G1 > 8" is necessary;
//   condition 3: the interpolated orientation lies on the short arc between the texel's two orientations (through the
//                0/360 seam when they are >= 180 apart, :83-111) up to float rounding (< 1e-3 deg), and the gate passes
//                only within 45 deg (+ 2^-16) of th_pi + rot, which lies in [45q, 45q + 45]: the arc, widened by a
//                0.05 deg margin, must meet (45q - 45, 45q + 90).
// The plane stores S_q(y,x) = distance to the next column x' > x of the SAME row with M_q(y,x') = 1 (capped at 255):
// a lane that has evaluated column x of row y may jump S_q columns ahead as long as its line stays in row y.
// One block per image row, one warp per bin; right-to-left over 32-column chunks.
// ---------------------------------------------------------------------------------------------
constexpr float kSkipBinDeg = 360.0f / kSkipBins;

// One block = one image row.  Phase 1: every texel of the row is evaluated once, one ballot per orientation bin into
// shared mask words.  Phase 2: one thread per bin walks the words right to left and leaves, per word, the distance from
// its first column to the next set bit at or after it.  Phase 3: every thread reads its eight distances off its own mask
// word or, when no higher bit is set there, off the next word's entry - no loop.  (The first version looked ahead word by
// word from every thread: in texture-free regions, where nothing is set, that was 15 iterations per bin and thread, 665
// warp-instructions per 32 pixels and 60 % of the packing time; this form needs about 150.)
constexpr int kSkipSpan = 256;                 // threads per block
constexpr int kSkipMaxWords = 8192 / 32;       // scan_columns3 is used for images up to 8192 columns
__global__ void __launch_bounds__(kSkipSpan) k_skip(DevArena A, DevParams P, PackBatch B)
{
    const int slot = B.slot[blockIdx.z];
    __shared__ unsigned s_m[kSkipBins][kSkipMaxWords];
    __shared__ unsigned short s_nd[kSkipBins][kSkipMaxWords + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y = blockIdx.y;
    const int n_words = (P.W + 31) >> 5;
    const float4* __restrict__ row = A.tex + (size_t)slot * A.P + (size_t)y * P.W;
    for (int w = warp; w < n_words; w += kSkipSpan / 32) {
        const int x = w * 32 + lane;
        bool g1 = false;
        float c = 0.f, h = 0.f;
        if (x < P.W) {
            const float4 t = row[x];
            const float d = fabsf(t.z - t.w);
            c = (t.z + t.w) * 0.5f;
            h = d * 0.5f;
            if (!(d < 180.f)) { c += 180.f; h = (360.f - d) * 0.5f; }
            g1 = t.x > kLambdaG2 || t.y > kLambdaG2;
        }
        const float reach = h + (0.5f * kSkipBinDeg + 45.f + 0.05f);
        unsigned mine = 0u;  // lane q keeps the ballot of bin q
#pragma unroll
        for (int q = 0; q < kSkipBins; ++q) {
            float dist = fabsf(c - ((float)q + 0.5f) * kSkipBinDeg);  // c in [0, 540), bin centre in (0, 360)
            if (dist > 360.f) dist -= 360.f;
            if (dist > 180.f) dist = 360.f - dist;
            const unsigned bal = __ballot_sync(SDM_FULL, g1 && (dist < reach));
            if (lane == q) mine = bal;
        }
        if (lane < kSkipBins) s_m[lane][w] = mine;
    }
    __syncthreads();
    if (threadIdx.x < kSkipBins) {  // distance from the first column of word w to the next set bit at or after it
        const int q = threadIdx.x;
        unsigned nd = 1024u;
        s_nd[q][n_words] = (unsigned short)nd;
        for (int w = n_words - 1; w >= 0; --w) {
            const unsigned v = s_m[q][w];
            nd = v ? (unsigned)(__ffs(v) - 1) : min(nd + 32u, 1024u);
            s_nd[q][w] = (unsigned short)nd;
        }
    }
    __syncthreads();
    uint8_t* __restrict__ out_row = A.skip + (size_t)slot * kSkipBins * A.P + (size_t)y * P.W;  // + q * A.P: plane of bin q
    for (int w = warp; w < n_words; w += kSkipSpan / 32) {
        const int x = w * 32 + lane;
        if (x >= P.W) continue;
#pragma unroll
        for (int q = 0; q < kSkipBins; ++q) {
            const unsigned higher = (s_m[q][w] >> lane) >> 1;
            const int sd = higher ? __ffs(higher) : 32 - lane + (int)s_nd[q][w + 1];
            out_row[(size_t)q * A.P + x] = (uint8_t)min(sd, 255);
        }
    }
}

__device__ __forceinline__ void load_item(DevItem& s_item, const DevItem* __restrict__ src_item)
{
    const int* src = reinterpret_cast<const int*>(src_item);
    int* dst = reinterpret_cast<int*>(&s_item);
    const int header = (int)(offsetof(DevItem, pair) / 4);
    const int words = header + src_item->n_nbr * (int)(sizeof(DevPair) / 4);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// Work plan of a pass: the candidates of the batch's keyframes are cut into chunks of kChunk
// consecutive candidates of ONE keyframe.  k_plan builds the chunk prefix over the batch on the
// device (candidate counts never travel to the host, so a pass can be enqueued while the uploads
// it depends on are still in flight); the pass kernels are persistent: a fixed grid of blocks
// pulls chunk ids from an atomic counter until the plan is exhausted.
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 128;

struct DevPlan {
    int* chunk_off;  // [n_items + 1] exclusive prefix of chunks per batch entry
    int* counter;    // next chunk to hand out
    const int* order;  // batch entry -> index into items[] (nullptr = identity)
    int n_items;
};

__global__ void __launch_bounds__(1024)
k_plan(DevPlan plan, const DevItem* __restrict__ items, const int* __restrict__ cand_count, DevStats* stats)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    __shared__ unsigned long long s_cands;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_base = 0; s_cands = 0ULL; *plan.counter = 0; }
    __syncthreads();
    for (int i0 = 0; i0 < plan.n_items; i0 += 1024) {
        const int i = i0 + tid;
        int cnt = 0;
        if (i < plan.n_items) cnt = cand_count[items[plan.order ? plan.order[i] : i].kf];
        const int chunks = (cnt + kChunk - 1) / kChunk;
        int incl = chunks;  // inclusive warp scan
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(SDM_FULL, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        unsigned long long c64 = (unsigned long long)cnt;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) c64 += __shfl_xor_sync(SDM_FULL, c64, off);
        if (lane == 0 && c64) atomicAdd(&s_cands, c64);
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(SDM_FULL, w, off);
                if (lane >= off) w += v;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int before = s_base + (warp ? s_warp[warp - 1] : 0) + incl - chunks;
        if (i < plan.n_items) plan.chunk_off[i] = before;
        __syncthreads();
        if (tid == 0) s_base += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) {
        plan.chunk_off[plan.n_items] = s_base;
        if (stats) stats->candidates = s_cands;
    }
}

// fetch the next chunk for this block: returns false when the plan is exhausted.  On return s_item holds the
// chunk's keyframe and first = index of the chunk's first candidate in that keyframe's list.
__device__ __forceinline__ bool next_chunk(const DevPlan& plan, const DevItem* __restrict__ items, DevItem& s_item,
                                           int& s_chunk, int& cur_entry, int& first)
{
    __syncthreads();  // everyone is done with the previous chunk (s_item, s_chunk)
    if (threadIdx.x == 0) s_chunk = atomicAdd(plan.counter, 1);
    __syncthreads();
    const int chunk = s_chunk;
    const int total = plan.chunk_off[plan.n_items];
    if (chunk >= total) return false;
    int lo = 0, hi = plan.n_items;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (plan.chunk_off[mid] <= chunk) lo = mid; else hi = mid;
    }
    if (lo != cur_entry) {
        load_item(s_item, &items[plan.order ? plan.order[lo] : lo]);
        cur_entry = lo;
    }
    __syncthreads();
    first = (chunk - plan.chunk_off[lo]) * kChunk;
    return true;
}

// ---------------------------------------------------------------------------------------------
// K3 + K4, warp-per-pixel variant (A/B reference, env SDM_SCAN=warp): lanes stride the columns of a
// search range, argmin by shuffles, lane j refines neighbour j, chi-square sets across lanes.
// ---------------------------------------------------------------------------------------------
constexpr int kPass1Warps = 8;

// --- the north star's wording of this kernel: "warp-per-pixel ... stages neighbour-keyframe tiles in shared memory via
// TMA" (k_pass1_tma, env SDM_SCAN=warp_tma; an A/B variant like k_pass1).  The texel arena is described to the TMA unit as
// a 3-D tensor {4 floats, W, slots * H}; for every 64-column segment of a search range whose line positions span at most
// kTmaBH row pairs, one elected lane requests the box {4, 64, kTmaBH} into the warp's shared-memory tile
// (cp.async.bulk.tensor.3d ... mbarrier::complete_tx::bytes) and the lanes read their texels from it; segments that do not
// fit the box (steep lines) and the intensity pairs come from global memory as before.  Same arithmetic, same bits.
// What the A/B measures (profiles/README.md): a box is a rectangle, an epipolar line is not - the box holds
// 64 x kTmaBH texels of which the 64 on the line are used.
constexpr int kTmaBW = 64, kTmaBH = 16, kTmaWarps = 4;
constexpr int kTmaTileBytes = kTmaBW * kTmaBH * 16;
struct TmaTile {
    float4* tile;        // this warp's [kTmaBH][kTmaBW] texels
    uint64_t* mbar;      // this warp's mbarrier
    const CUtensorMap* map;
    unsigned phase;
    int rows_per_slot;   // H
};
__device__ __forceinline__ void tma_load_box(const TmaTile& T, int x0, int y0_global)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(T.tile), bar = (unsigned)__cvta_generic_to_shared(T.mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTmaTileBytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(T.map), "r"(0), "r"(x0), "r"(y0_global), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_wait(const TmaTile& T)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(T.mbar);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(T.phase) : "memory");
}

template <bool kTma>
__device__ __forceinline__ bool scan_pixel_warp(const DevArena& A, const DevParams& P, const DevItem& s_item, int ci, int lane,
                                                TmaTile* T = nullptr)
{
    const int kf = s_item.kf;
    const int N = s_item.n_nbr;
    const uint32_t packed = A.cand[(size_t)kf * A.P + ci];
    const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);

    const float4 t1 = __ldg(&A.tex[(size_t)kf * A.P + (size_t)y * P.W + x]);
    const float gradc = t1.x, th_pi = t1.z;
    const float pixel = (float)__ldg(&A.ipair[(size_t)kf * A.P + (size_t)y * P.W + x]).x;
    const float* K = s_item.K;
    const float xn = (x - K[2]) / K[0], yn = (y - K[3]) / K[1];
    const float Hm1 = (float)(P.H - 1);

    // phase A: lane j sets up neighbour j
    PairSetup my;
    my.valid = false; my.u_lo = 1; my.u_hi = 0; my.ab = 0.f; my.cb = 0.f; my.th_line = 0.f; my.ang_pi_rot = 0.f;
    if (lane < N) my = pair_setup(s_item.pair[lane], K, P, x, y, xn, yn, s_item.min_depth, s_item.max_depth, th_pi);

    // phase B
    int my_best = -1;  // lane j: best column of neighbour j (or -1)
    for (int j = 0; j < N; ++j) {
        const int u_lo = __shfl_sync(SDM_FULL, my.u_lo, j);
        const int u_hi = __shfl_sync(SDM_FULL, my.u_hi, j);
        if (u_lo > u_hi) continue;
        const float ab = __shfl_sync(SDM_FULL, my.ab, j);
        const float cb = __shfl_sync(SDM_FULL, my.cb, j);
        const float th_line = __shfl_sync(SDM_FULL, my.th_line, j);
        const float apr = __shfl_sync(SDM_FULL, my.ang_pi_rot, j);
        const size_t nb = (size_t)s_item.pair[j].slot * A.P;
        const float4* tex2 = A.tex + nb;
        const uchar2* ip2 = A.ipair + nb;
        float best_err = 100000.0f;
        int best_u = 0x7fffffff;
        if (kTma) {
            for (int us = u_lo; us <= u_hi; us += kTmaBW) {
                const int ue = min(us + kTmaBW - 1, u_hi);
                // row pairs the segment touches (v is monotone in u): only if every line position is inside the image
                const float v0 = -(ab * (float)us + cb), v1 = -(ab * (float)ue + cb);
                const float vlo = fminf(v0, v1), vhi = fmaxf(v0, v1);
                const int ya = (int)floorf(vlo), yb = (int)floorf(vhi);
                const bool boxed = vlo >= 0.f && vhi <= Hm1 && yb - ya < kTmaBH;  // (NaN: false)
                if (boxed) {
                    __syncwarp();  // everybody is done with the previous tile
                    if (lane == 0) tma_load_box(*T, us, s_item.pair[j].slot * T->rows_per_slot + ya);
                    tma_wait(*T);
                    T->phase ^= 1u;
                }
                for (int u = us + lane; u <= ue; u += 32) {
                    float err, pe, ge;
                    RowW r;
                    size_t idx;
                    if (!cand_rows(P, Hm1, u, ab, cb, r, idx)) continue;
                    const float4 t = boxed ? T->tile[(r.y0 - ya) * kTmaBW + (u - us)] : __ldg(&tex2[idx]);
                    if (cand_gates(t, r, ip2, idx, P, th_line, apr, pixel, gradc, err, pe, ge))
                        if (err < best_err) { best_err = err; best_u = u; }
                }
            }
        } else
        for (int u = u_lo + lane; u <= u_hi; u += 32) {
            float err, pe, ge;
            if (eval_candidate(tex2, ip2, P, Hm1, u, ab, cb, th_line, apr, pixel, gradc, err, pe, ge)) {
                if (err < best_err) { best_err = err; best_u = u; }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float oe = __shfl_xor_sync(SDM_FULL, best_err, off);
            const int ou = __shfl_xor_sync(SDM_FULL, best_u, off);
            if (oe < best_err || (oe == best_err && ou < best_u)) { best_err = oe; best_u = ou; }
        }
        if (lane == j && best_err < 100000.0f) my_best = best_u;
    }

    // phase C
    float hd = 0.f, hs = 0.f;
    bool keep = false;
    if (my_best >= 0) {
        const DevPair& g = s_item.pair[lane];
        const size_t nb = (size_t)g.slot * A.P;
        float err, pe, ge;
        eval_candidate(A.tex + nb, A.ipair + nb, P, Hm1, my_best, my.ab, my.cb, my.th_line, my.ang_pi_rot, pixel,
                       gradc, err, pe, ge);
        Hypo h = refine_hypothesis(A.tex + nb, A.ipair + nb, g, K, P, my_best, my.ab, my.cb, pe, ge, xn, yn);
        hd = h.depth;
        hs = h.sigma;
        keep = (1.0f / hd > 0.0f);  // :472
    }

    // phase D: InverseDepthHypothesisFusion (:978-1009)
    const unsigned vmask = __ballot_sync(SDM_FULL, keep);
    float out_d = 0.f, out_s = 0.f;
    bool fused = false;
    if (__popc(vmask) > P.lambdaN) {
        unsigned m = 0;
        for (unsigned rem = vmask; rem; rem &= rem - 1) {
            const int b = __ffs(rem) - 1;
            const float db = __shfl_sync(SDM_FULL, hd, b);
            const float sb = __shfl_sync(SDM_FULL, hs, b);
            if (keep && (b == lane || chi_compatible(hd, db, hs, sb, P.chi_fusion_lt))) m |= 1u << b;
        }
        int key = keep ? ((__popc(m) << 8) | (31 - lane)) : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) key = max(key, __shfl_xor_sync(SDM_FULL, key, off));
        const int best_n = key >> 8, winner = 31 - (key & 0xff);
        if (best_n > P.lambdaN) {
            const unsigned wm = __shfl_sync(SDM_FULL, m, winner);
            // GetFusion (:1669-1692): per-lane double terms, then the reference's sequential float sums
            double q1 = 0.0, q2 = 0.0;
            if (keep) {
                const double s2 = (double)hs * (double)hs;
                q1 = (double)hd / s2;
                q2 = 1.0 / s2;
            }
            float pjsj = 0.f, rsj = 0.f;
            for (unsigned rem = wm; rem; rem &= rem - 1) {
                const int b = __ffs(rem) - 1;
                const double t1d = __shfl_sync(SDM_FULL, q1, b);
                const double t2d = __shfl_sync(SDM_FULL, q2, b);
                pjsj = (float)((double)pjsj + t1d);
                rsj = (float)((double)rsj + t2d);
            }
            out_d = pjsj / rsj;
            out_s = sqrtf(1.0f / rsj);
            fused = true;
        }
    }
    if (lane == 0) {
        const size_t own = (size_t)kf * A.P + (size_t)y * P.W + x;
        A.rs[own] = make_float2(out_d, out_s);
        A.dpl[own] = out_d;
        A.spl[own] = out_s;
    }
    return fused;
}

__global__ void __launch_bounds__(kPass1Warps * 32)
k_pass1(DevArena A, DevParams P, const DevItem* __restrict__ items, DevPlan plan, DevStats* stats)
{
    __shared__ DevItem s_item;
    __shared__ int s_chunk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int cur_entry = -1, first = 0;
    unsigned long long n_fused = 0;
    while (next_chunk(plan, items, s_item, s_chunk, cur_entry, first)) {
        const int cnt = A.cand_count[s_item.kf];
        for (int k = warp; k < kChunk; k += kPass1Warps) {
            const int ci = first + k;
            if (ci >= cnt) break;
            if (scan_pixel_warp<false>(A, P, s_item, ci, lane)) ++n_fused;
        }
    }
    if (stats && lane == 0 && n_fused) atomicAdd(&stats->fused, n_fused);
}

__global__ void __launch_bounds__(kTmaWarps * 32)
k_pass1_tma(DevArena A, DevParams P, const DevItem* __restrict__ items, DevPlan plan, DevStats* stats,
            const __grid_constant__ CUtensorMap tex_map)
{
    __shared__ DevItem s_item;
    __shared__ int s_chunk;
    __shared__ __align__(8) uint64_t s_bar[kTmaWarps];
    extern __shared__ __align__(128) unsigned char s_tiles[];  // kTmaWarps tiles of kTmaTileBytes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar[warp]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    TmaTile T;
    T.tile = reinterpret_cast<float4*>(s_tiles + (size_t)warp * kTmaTileBytes);
    T.mbar = &s_bar[warp];
    T.map = &tex_map;
    T.phase = 0;
    T.rows_per_slot = P.H;
    int cur_entry = -1, first = 0;
    unsigned long long n_fused = 0;
    while (next_chunk(plan, items, s_item, s_chunk, cur_entry, first)) {
        const int cnt = A.cand_count[s_item.kf];
        for (int k = warp; k < kChunk; k += kTmaWarps) {
            const int ci = first + k;
            if (ci >= cnt) break;
            if (scan_pixel_warp<true>(A, P, s_item, ci, lane, &T)) ++n_fused;
        }
    }
    if (stats && lane == 0 && n_fused) atomicAdd(&stats->fused, n_fused);
}

// ---------------------------------------------------------------------------------------------
// K3 + K4, lane-per-pixel variant (the default).  One thread owns one candidate pixel and walks the
// search ranges of its N neighbours column by column, exactly like the reference's inner loop
// (:770-821); the 32 lanes of a warp hold 32 neighbouring candidates of one 32x8 tile, whose
// epipolar lines in a neighbour keyframe are nearly parallel and one column apart, so at every
// step of the walk the warp's 16-byte texel loads fall on a few contiguous row segments.
// Compared with the warp-per-pixel kernel: no shuffles, no idle lanes during set-up / refinement /
// fusion, v(u+1) is carried to the next column instead of recomputed, the best column's residuals
// are kept instead of re-evaluated.  Hypotheses of a pixel stay in shared memory ([nbr][thread]).
// ---------------------------------------------------------------------------------------------
constexpr int kLaneBlock = kChunk;

__device__ __forceinline__ bool in_rows(float v, float Hm1) { return v >= 0.f && v <= Hm1; }

// The scan body skips column u unless rows v(u-1), v(u), v(u+1) all lie in [0, H-1] (:773-785), with
// v(k) = -((a/b)*k + c/b) evaluated in float.  Float rounding is monotone, so v(k) is monotone in k and
// the set {k : 0 <= v(k) <= H-1} is ONE integer interval [p, q]; the admissible columns are therefore
// [max(u_lo, p+1), min(u_hi, q-1)].  p and q are located with the exact float predicate (endpoints first:
// in the common case the whole range is inside; otherwise a binary search on the monotone half-tests).
__device__ __forceinline__ float line_row(float ab, float cb, int k) { return -(ab * (float)k + cb); }

__device__ __forceinline__ void valid_columns(float ab, float cb, float Hm1, int u_lo, int u_hi, int& ua, int& ub)
{
    const int lo = u_lo - 1, hi = u_hi + 1;
    const float v_lo = line_row(ab, cb, lo), v_hi = line_row(ab, cb, hi);
    if (in_rows(v_lo, Hm1) && in_rows(v_hi, Hm1)) { ua = u_lo; ub = u_hi; return; }
    ua = 1; ub = 0;
    if (!(v_lo == v_lo) || !(v_hi == v_hi)) return;  // NaN line: nothing is inside
    // orient so that f(k) = s * v(k) is non-decreasing in k: rows too small on the left, too large on the right
    const bool inc = v_hi >= v_lo;
    // p = first k with v(k) inside the lower bound of its direction, q = last k inside the upper bound
    int a = lo, b = hi + 1;  // first k in [lo, hi+1) with "left test" true
    while (a < b) {
        const int m = (a + b) >> 1;
        const float v = line_row(ab, cb, m);
        const bool ok = inc ? (v >= 0.f) : (v <= Hm1);
        if (ok) b = m; else a = m + 1;
    }
    const int p = a;
    a = lo - 1; b = hi;  // last k in (lo-1, hi] with "right test" true
    while (a < b) {
        const int m = (a + b + 1) >> 1;
        const float v = line_row(ab, cb, m);
        const bool ok = inc ? (v <= Hm1) : (v >= 0.f);
        if (ok) a = m; else b = m - 1;
    }
    const int q = a;
    ua = max(u_lo, p + 1);
    ub = min(u_hi, q - 1);
}

// ---------------------------------------------------------------------------------------------
// Second-generation column loop of the lane-per-pixel scan (kMode 2), for the reference's constants
// (lambdaG = 8, lambdaL = 80, lambdaTheta = 45, THETA = 0.23f; ProbabilityMapping.h:45-56).  Same arithmetic results
// as the loop in scan_pixel_lane, fewer issued instructions per column (the scan is issue-bound: profiles/README.md):
//  * no int<->float conversions: the column is carried as a float (exact below 2^24); floor(v) comes from ONE
//    round-down add, r = v (+)RD 2^23 = 2^23 + floor(v) for 0 <= v < 2^23 (valid_columns guarantees 0 <= v <= H-1
//    for every column touched, the prefetched one included), so floor(v) = r - 2^23 as a float (exact) and the row
//    index is the low mantissa of r: texel index = bits(r) * W + (u - 0x4B000000 * W) modulo 2^32.
//    w1 = v - floor(v) is exact (Sterbenz), hence 1 - w1 and the reference's (floor(v) + 1) - v (:66-84) are roundings
//    of the same real number: w0 = 1 - w1.  [v = -0 gives w1 = -0 instead of +0: no product, sum or comparison
//    downstream can tell the two apart.]
//  * gates 2 and 3 on the RAW difference d = gth - c (c = th_line, ang_pi_rot): the `>= 360` wrap step is dropped and
//    gate 3 needs no wrap at all; exact for d < 400, every negative d and NaN (tools/verify_gate_algebra2.py, all
//    2^32 floats).  gth <= 360.0001 when the orientation planes hold [0, 360], and 0 <= c <= 360 when |rot| <= 360.
//  * x / THETA for x = ge*ge by the correctly rounded reciprocal and two fused steps (Markstein): q0 = x*r,
//    q = fma(fma(-q0, THETA, x), r, q0), == the IEEE quotient for x = 0 and every x in [2^-80, 2^80] (k_verify_div
//    runs over that whole range when a context is created and disables this loop on any mismatch).  ge = G1 - G2
//    with both gradients in (8, 2^40] is 0 or a multiple of 2^-20 below 2^40, so x is always inside.
//  * only (err, column) of the best column are tracked; its residuals are re-evaluated once after the loop.
// The preconditions on the planes (orientation in [0, 360], gradient magnitude <= 2^40, no NaN) are checked per
// keyframe slot by k_pack (plane_irregular); k_pass1_lane falls back to the first-generation loop otherwise.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float div_by_theta2(float x)
{
    const float q0 = __fmul_rn(x, kRTheta2);
    return __fmaf_rn(__fmaf_rn(-q0, kTheta2, x), kRTheta2, q0);
}

__global__ void k_verify_div(unsigned long long* mismatches)
{
    // x = 0 and every float in [2^-80, 2^80]: bit patterns 0x17800000 .. 0x67800000
    const unsigned lo = 0x17800000u, hi = 0x67800000u;
    unsigned long long bad = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) bad += (__float_as_uint(div_by_theta2(0.f)) != 0u);
    for (unsigned long long b = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= hi;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)b);
        bad += (__float_as_uint(div_by_theta2(x)) != __float_as_uint(__fdiv_rn(x, kTheta2)));
    }
    if (bad) atomicAdd(mismatches, bad);
}

#ifndef SDM_SCAN2_UNROLL
#define SDM_SCAN2_UNROLL 2
#endif
constexpr int kScan2Unroll = SDM_SCAN2_UNROLL;
// Scans columns ua .. ub of one neighbour; returns the best column (strict minimum of err, lowest column wins) or -1.
// texw2 = the neighbour's wrap-encoded texels (encode_theta_pair, k_pack), ip2 = its intensity pairs.
// Measured on B200 against the raw texel + yangle branch: 10.09 -> 9.27 ms per 200 keyframes (the branch ran for 31 %
// of the warp-level gate evaluations); float intensity pairs instead of uchar2 made no difference and were dropped.
__device__ __forceinline__ int scan_columns2(const float4* __restrict__ texw2, const uchar2* __restrict__ ip2, int W,
                                             int ua, int ub, float ab, float cb, float th_line, float ang_pi_rot,
                                             float pixel, float gradc, float& best_err)
{
    constexpr float kMagic = 8388608.0f;
    constexpr float kT2 = -0x1.67fff8p+5f;  // -(45 - 2^-16): fl(d + 360) > 315  <=>  d > kT2
    const char* tb = reinterpret_cast<const char*>(texw2);
    const char* ib = reinterpret_cast<const char*>(ip2);
    unsigned Wm = (unsigned)W;
    asm volatile("" : "+l"(tb), "+l"(ib), "+r"(Wm));
    unsigned k = (unsigned)ua - 0x4B000000u * Wm;  // texel index = bits(r) * W + k (mod 2^32)
    int best_n = 0;                                // columns left when the best one was found: 0 = none
    float uf = (float)ua;
    float vn = -(ab * uf + cb);
    float r = __fadd_rd(vn, kMagic);
    float w1n = vn - (r - kMagic);
    unsigned idxn = __float_as_uint(r) * Wm + k;
    float4 tn = __ldg(reinterpret_cast<const float4*>(tb + (size_t)idxn * 16));
#pragma unroll kScan2Unroll
    for (int n = ub - ua + 1; n > 0; --n) {
        const float4 t = tn;
        const float w1 = w1n, w0 = 1.0f - w1n;
        const unsigned idx = idxn;
        uf += 1.0f;
        k += 1u;
        vn = -(ab * uf + cb);
        r = __fadd_rd(vn, kMagic);
        w1n = vn - (r - kMagic);
        idxn = __float_as_uint(r) * Wm + k;
        tn = __ldg(reinterpret_cast<const float4*>(tb + (size_t)idxn * 16));
        // Blackwell packed fp32x2 multiplies (FMUL2: two independent RN products per issue slot)
        const float2 w01 = make_float2(w0, w1);
        const float2 gp = __fmul2_rn(make_float2(t.x, t.y), w01);
        const float g2 = gp.x + gp.y;
        if (g2 <= kLambdaG2) continue;  // condition 1
        // yangle on the wrap-encoded pair: a negative sum is minus the reference's value before its `>= 360` step
        const float2 tp = __fmul2_rn(make_float2(t.z, t.w), w01);
        float gs = tp.x + tp.y;
        if (gs <= -360.f) gs += 360.f;
        const float gth = fabsf(gs);
        // the two direction gates only skip, so their order is free: condition 3 (5 instructions, rejects ~3 of 4
        // random orientations) goes first and fewer warps still need condition 2 (7 instructions);
        // measured 9.29 -> 9.15 ms per 200 keyframes
        const float d3 = gth - ang_pi_rot;
        if ((d3 >= 45.f || d3 <= kT2) && fabsf(d3) <= 315.f) continue;  // condition 3
        const float d2 = gth - th_line;
        const float ang = d2 < 0.f ? d2 + 360.f : d2;  // condition 2
        if (fabsf(fabsf(ang - 180.f) - 90.f) <= 10.f) continue;
        const uchar2 i2 = __ldg(reinterpret_cast<const uchar2*>(ib + (size_t)idx * 2));
        const float2 ip = __fmul2_rn(make_float2((float)i2.x, (float)i2.y), w01);
        // {pe, ge} = {pixel, gradc} - {I2, G2} and their squares as packed operations
        const float2 res = __fadd2_rn(make_float2(pixel, gradc), make_float2(-(ip.x + ip.y), -g2));
        const float2 sq = __fmul2_rn(res, res);
        const float err = sq.x + div_by_theta2(sq.y);
        if (err < best_err) { best_err = err; best_n = n; }
    }
    return best_n > 0 ? ub + 1 - best_n : -1;
}

// Cache policies of the walk's three loads, A/B-timed on one box (scan ms, config 2 / config 4; plain ld.global.nc everywhere:
// 8.012 / 28.08): skip byte L1::evict_last 7.977 / 27.93 (kept: the byte's sector serves the next ~12 visits of the lane),
// intensity pair L1::no_allocate 8.436 / 30.87, both 8.415 / 30.71, texel L1::evict_first 8.721 / 33.42.
__device__ __forceinline__ unsigned ld_skip(const char* p)
{
#if !defined(SDM_SKIP_PLAIN_LOAD)
    unsigned v;
    asm volatile("ld.global.nc.L1::evict_last.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(reinterpret_cast<const uint8_t*>(p));
#endif
}
__device__ __forceinline__ uchar2 ld_ipair(const char* p)
{
#if defined(SDM_IPAIR_NOALLOC)
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return make_uchar2((unsigned char)(v & 0xffu), (unsigned char)(v >> 8));
#else
    return __ldg(reinterpret_cast<const uchar2*>(p));
#endif
}
__device__ __forceinline__ float4 ld_texel(const char* p)
{
#if defined(SDM_TEXEL_EVICT_FIRST)
    float4 v;
    asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(reinterpret_cast<const float4*>(p));
#endif
}
// ---------------------------------------------------------------------------------------------
// Third-generation column loop (kMode 3): scan_columns2's arithmetic on the columns that can matter.  The loop of
// scan_columns2 is bound by instruction issue and evaluates every column of the search range although 53 % of them
// fail condition 1 and 57 % of the rest fail condition 3; a warp pays for a column as soon as one lane needs it, so
// what helps is fewer columns per LANE.  After evaluating column u in row pair y, the lane reads S = skip[q][y][u]
// (k_skip: no texel of row y in columns u+1 .. u+S-1 can survive conditions 1 and 3 for this pixel's orientation bin q)
// and m = the number of following columns whose line position certainly stays in row pair y, and advances by
// max(1, min(S, m + 1)).  Skipped columns would have been skipped by the reference's `continue`s, so best_err /
// best column are untouched: same bits as scan_columns2.  The first column of a search range and the first column
// after every row change are evaluated unconditionally (a superset is all that exactness needs).
//   m: with rem = distance of v(u) to the row boundary the line is heading for (1 - w1 if v grows with u, else w1),
//   the columns u + i with i * |a/b| <= rem - 2^-8 stay in the row: v is evaluated as -fl(fl(ab*u) + cb), each
//   rounding <= 2^-13 for |values| < 4096 (2^-11 below 16384: the loop is used for images up to 8192 columns / rows),
//   so two evaluations differ from the real line by < 2^-10 rows, a quarter of the 2^-8 margin; 1/|a/b| is scaled by
//   1 - 2^-10 and capped at 2^20 so that the float product cannot exceed the real quotient.
// CPU sizing of the design: tools/sim_mask_walk2.py / DESIGN.md section 5.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int scan_columns3(const float4* __restrict__ texw2, const uchar2* __restrict__ ip2,
                                             const uint8_t* __restrict__ skip2, int W, int ua, int ub, float ab, float cb,
                                             float th_line, float ang_pi_rot, float pixel, float gradc, float& best_err)
{
    constexpr float kMagic = 8388608.0f;
    constexpr float kT2 = -0x1.67fff8p+5f;  // -(45 - 2^-16)
    const char* tb = reinterpret_cast<const char*>(texw2);
    const char* ib = reinterpret_cast<const char*>(ip2);
    const char* sb = reinterpret_cast<const char*>(skip2);  // the plane of this pixel's orientation bin
    unsigned Wm = (unsigned)W;
    asm volatile("" : "+l"(tb), "+l"(ib), "+l"(sb), "+r"(Wm));
    const float inv = fminf((1.0f - 0x1p-10f) / fabsf(ab), 0x1p20f);
    const float dinv1 = 0x1p-8f * inv - 1.0f;  // floor(rem * inv - dinv1) = m + 1
    const bool up = ab < 0.f;                  // v grows with u: the line leaves the row pair through its upper boundary
    unsigned k = (unsigned)ua - 0x4B000000u * Wm;
    int best_n = 0;
    int n = ub - ua + 1;
    float uf = (float)ua;
    float vn = -(ab * uf + cb);
    float r = __fadd_rd(vn, kMagic);
    float w1n = vn - (r - kMagic);
    unsigned idxn = __float_as_uint(r) * Wm + k;
    float4 tn = ld_texel(tb + (size_t)idxn * 16);
    unsigned sn = ld_skip(sb + (size_t)idxn);
#pragma unroll kScan2Unroll
    while (n > 0) {
        const float4 t = tn;
        const float w1 = w1n, w0 = 1.0f - w1n;
        const unsigned idx = idxn;
        const int ncur = n;
        // advance: as far as the skip byte and the row allow, never past column ub + 1 (the last valid address)
        const float rem = up ? w0 : w1;
        const int m1 = __float2int_rd(rem * inv - dinv1);
        const int step = max(1, min(min((int)sn, m1), ncur));
        n -= step;
        k += (unsigned)step;
        uf += (float)step;
        vn = -(ab * uf + cb);
        r = __fadd_rd(vn, kMagic);
        w1n = vn - (r - kMagic);
        idxn = __float_as_uint(r) * Wm + k;
        tn = ld_texel(tb + (size_t)idxn * 16);
        sn = ld_skip(sb + (size_t)idxn);
        const float2 w01 = make_float2(w0, w1);
        const float2 gp = __fmul2_rn(make_float2(t.x, t.y), w01);
        const float g2 = gp.x + gp.y;
        if (g2 <= kLambdaG2) continue;  // condition 1
        const float2 tp = __fmul2_rn(make_float2(t.z, t.w), w01);
        float gs = tp.x + tp.y;
        if (gs <= -360.f) gs += 360.f;
        const float gth = fabsf(gs);
        const float d3 = gth - ang_pi_rot;
        if ((d3 >= 45.f || d3 <= kT2) && fabsf(d3) <= 315.f) continue;  // condition 3
        const float d2 = gth - th_line;
        const float ang = d2 < 0.f ? d2 + 360.f : d2;  // condition 2
        if (fabsf(fabsf(ang - 180.f) - 90.f) <= 10.f) continue;
        const uchar2 i2 = ld_ipair(ib + (size_t)idx * 2);
        const float2 ip = __fmul2_rn(make_float2((float)i2.x, (float)i2.y), w01);
        const float2 res = __fadd2_rn(make_float2(pixel, gradc), make_float2(-(ip.x + ip.y), -g2));
        const float2 sq = __fmul2_rn(res, res);
        const float err = sq.x + div_by_theta2(sq.y);
        if (err < best_err) { best_err = err; best_n = ncur; }
    }
    return best_n > 0 ? ub + 1 - best_n : -1;
}

// kMode 0: gates for any thresholds; 1: exact short gate forms for lambdaL = 80, lambdaTheta = 45;
// 2: second-generation column loop (below), same decisions and values as 1 under its two preconditions:
//    (i) regular planes and |rot| <= 360 (checked per keyframe: k_pack's plane_irregular flags / k_pass1_lane),
//    (ii) the reference's constants and div_by_theta2 exact (checked when the context is created: k_verify_div).
template <int kMode>
__device__ __forceinline__ bool scan_pixel_lane(const DevArena& A, const DevParams& P, const DevItem& s_item,
                                                float2 (*s_h)[kLaneBlock], int ci, int tid)
{
    constexpr bool kFastGates = kMode >= 1;
    const int kf = s_item.kf;
    const int N = s_item.n_nbr;
    const uint32_t packed = A.cand[(size_t)kf * A.P + ci];
    const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
    const size_t own = (size_t)kf * A.P + (size_t)y * P.W + x;
    float gradc, th_pi, pixel;
    if (kMode >= 2) {  // the planes the column loop reads anyway (the raw orientation only where the pair is wrapped)
        const float4 t1 = __ldg(&A.texw[own]);
        gradc = t1.x;
        th_pi = t1.z;
        if (t1.z < 0.f) th_pi = __ldg(&A.tex[own]).z;
        pixel = (float)__ldg(&A.ipair[own]).x;
    } else {
        const float4 t1 = __ldg(&A.tex[own]);
        gradc = t1.x;
        th_pi = t1.z;
        pixel = (float)__ldg(&A.ipair[own]).x;
    }
    const float* K = s_item.K;
    const float xn = (x - K[2]) / K[0], yn = (y - K[3]) / K[1];
    const float Hm1 = (float)(P.H - 1);
    int W = P.W;
    float lamG = P.lambdaG, lamL = P.lambdaL, lamT = P.lambdaTheta, theta = P.theta;
    asm volatile("" : "+r"(W), "+f"(lamG), "+f"(lamL), "+f"(lamT), "+f"(theta));
    int nh = 0;
    for (int j = 0; j < N; ++j) {
        const DevPair& g = s_item.pair[j];
        const PairSetup s = pair_setup(g, K, P, x, y, xn, yn, s_item.min_depth, s_item.max_depth, th_pi);
        if (s.u_lo > s.u_hi) continue;
        const size_t nb = (size_t)g.slot * A.P;
        const float4* __restrict__ tex2 = A.tex + nb;
        const uchar2* __restrict__ ip2 = A.ipair + nb;
        const float ab = s.ab, cb = s.cb;
        float best_err = 100000.0f, best_pe = 0.f, best_ge = 0.f;
        int best_u = -1;
        // columns whose three rows v(u-1), v(u), v(u+1) are inside the image (:773-785), as one interval
        int ua, ub;
        valid_columns(ab, cb, Hm1, s.u_lo, s.u_hi, ua, ub);
        if (kMode >= 2) {
            if (ua <= ub) {
                const float4* __restrict__ texw2 = A.texw + nb;
                if (kMode == 3) {
                    const int q = min(kSkipBins - 1, (int)(s.ang_pi_rot * (1.0f / kSkipBinDeg)));
                    const uint8_t* __restrict__ skip2 = A.skip + ((size_t)g.slot * kSkipBins + q) * A.P;
                    best_u = scan_columns3(texw2, ip2, skip2, W, ua, ub, ab, cb, s.th_line, s.ang_pi_rot, pixel, gradc, best_err);
                } else
                best_u = scan_columns2(texw2, ip2, W, ua, ub, ab, cb, s.th_line, s.ang_pi_rot, pixel, gradc, best_err);
                if (best_u >= 0) {  // residuals of the best column, same expressions as in the loop
                    const float vb = -(ab * (float)best_u + cb);
                    best_pe = pixel - ylin_im(ip2, W, vb, best_u);
                    best_ge = gradc - ylin_grad(texw2, W, vb, best_u);
                }
            }
        } else if (ua <= ub) {
            // keep the loop's operands in registers (ptxas otherwise re-derives them from constant memory)
            const char* tb = reinterpret_cast<const char*>(tex2);
            const char* ib = reinterpret_cast<const char*>(ip2);
            asm volatile("" : "+l"(tb), "+l"(ib));
            // software pipeline: the texel of column u+1 is requested while column u is evaluated.  Column
            // ub+1 is a valid address (ub <= W-2 and its row is inside the image by construction).
            float vn = -(ab * (float)ua + cb);
            float fl = floorf(vn);
            float w0n = (fl + 1.0f) - vn, w1n = vn - fl;
            unsigned idxn = (unsigned)((int)fl * W + ua);
            float4 tn = __ldg(reinterpret_cast<const float4*>(tb + (size_t)idxn * 16));
            // unrolled by two so that the pipeline registers (current / next texel) swap roles instead of being copied
#pragma unroll 2
            for (int u = ua; u <= ub; ++u) {
                const float4 t = tn;
                const float w0 = w0n, w1 = w1n;
                vn = -(ab * (float)(u + 1) + cb);
                fl = floorf(vn);
                w0n = (fl + 1.0f) - vn;
                w1n = vn - fl;
                idxn = (unsigned)((int)fl * W + (u + 1));
                tn = __ldg(reinterpret_cast<const float4*>(tb + (size_t)idxn * 16));
                const float g2 = t.x * w0 + t.y * w1;
                if (g2 <= lamG) continue;  // condition 1
                const float gth = yangle_interp(t.z, t.w, w0, w1);
                float ang = gth - s.th_line;  // condition 2
                if (ang >= 360.f) ang -= 360.f;
                if (ang < 0.f) ang += 360.f;
                float thd = gth - s.ang_pi_rot;  // condition 3
                if (thd >= 360.f) thd -= 360.f;
                if (thd < 0.f) thd += 360.f;
                if (kFastGates) {
                    // exact short forms of the fold sequences for lambdaL = 80, lambdaTheta = 45, checked over all
                    // 2^32 float bit patterns (tools/verify_gate_algebra.py): same decision, NaN included
                    if (fabsf(fabsf(ang - 180.f) - 90.f) <= 10.f) continue;
                    if (thd >= 45.f && thd <= 315.f) continue;
                } else {
                    if (ang > 180.f) ang = 360.f - ang;
                    if (ang > 90.f) ang = 180.f - ang;
                    if (ang >= lamL) continue;
                    if (thd > 180.f) thd = 360.f - thd;
                    if (thd >= lamT) continue;
                }
                const unsigned idx = (unsigned)((int)floorf(-(ab * (float)u + cb)) * W + u);
                const uchar2 i2 = __ldg(reinterpret_cast<const uchar2*>(ib + (size_t)idx * 2));
                const float pe = pixel - ((float)i2.x * w0 + (float)i2.y * w1);
                const float ge = gradc - g2;
                const float err = pe * pe + (ge * ge) / theta;
                if (err < best_err) { best_err = err; best_u = u; best_pe = pe; best_ge = ge; }
            }
        }
        if (best_err < 100000.0f) {
            // the gradient halves of tex and texw are the same: mode 2 stays on the plane its column loop has in cache
            const Hypo h = refine_hypothesis(kMode >= 2 ? A.texw + nb : tex2, ip2, g, K, P, best_u, ab, cb, best_pe, best_ge, xn, yn);
            if (1.0f / h.depth > 0.0f) {  // :472
                s_h[nh][tid] = make_float2(h.depth, h.sigma);
                ++nh;
            }
        }
    }
    // InverseDepthHypothesisFusion (:978-1009) over this pixel's nh hypotheses (neighbour order)
    float out_d = 0.f, out_s = 0.f;
    bool fused = false;
    if (nh > P.lambdaN) {
        unsigned best_mask = 0;
        int best_n = 0;
        if (nh <= 8) {
            // ChiTest is symmetric bit for bit ((a-b)^2 == (b-a)^2, float + commutes): each unordered pair once,
            // compatibility matrix in 64 bits (row a = bits 8a .. 8a+7)
            unsigned long long M = 0ULL;
            for (int a = 0; a + 1 < nh; ++a) {
                const float2 ha = s_h[a][tid];
                for (int b = a + 1; b < nh; ++b) {
                    const float2 hb = s_h[b][tid];
                    if (chi_compatible(ha.x, hb.x, ha.y, hb.y, P.chi_fusion_lt))
                        M |= (1ULL << (8 * a + b)) | (1ULL << (8 * b + a));
                }
            }
            for (int a = 0; a < nh; ++a) {
                const unsigned m = ((unsigned)(M >> (8 * a)) & 0xffu) | (1u << a);
                const int cnt = __popc(m);
                if (best_n < cnt) { best_n = cnt; best_mask = m; }
            }
        } else
        for (int a = 0; a < nh; ++a) {
            const float2 ha = s_h[a][tid];
            unsigned m = 1u << a;
            for (int b = 0; b < nh; ++b) {
                if (b == a) continue;
                const float2 hb = s_h[b][tid];
                if (chi_compatible(ha.x, hb.x, ha.y, hb.y, P.chi_fusion_lt)) m |= 1u << b;
            }
            const int cnt = __popc(m);
            if (best_n < cnt) { best_n = cnt; best_mask = m; }
        }
        if (best_n > P.lambdaN) {
            float pjsj = 0.f, rsj = 0.f;
            for (unsigned rem = best_mask; rem; rem &= rem - 1) {
                const float2 hb = s_h[__ffs(rem) - 1][tid];
                fusion_accumulate(hb.x, hb.y, pjsj, rsj);
            }
            out_d = pjsj / rsj;
            out_s = sqrtf(1.0f / rsj);
            fused = true;
        }
    }
    A.rs[own] = make_float2(out_d, out_s);
    A.dpl[own] = out_d;
    A.spl[own] = out_s;
    return fused;
}

#ifndef SDM_LANE_MINB
#define SDM_LANE_MINB 12  // 40 registers, 48 warps / SM: measured optimum on B200 with the second-generation loop
                          // (blocks/SM 9: 11.09 ms, 10: 10.71, 12: 10.32, 16: 10.50 per 200 keyframes; the spills are outside the loop)
#endif
// scan_pixel_lane<2>'s precondition (i) for the keyframe of s_item (block-uniform)
__device__ __forceinline__ bool item_regular(const DevArena& A, const DevItem& s_item)
{
    int irr = A.plane_irregular[s_item.kf];
    for (int j = 0; j < s_item.n_nbr; ++j) {
        const float rot = s_item.pair[j].rot;
        irr |= A.plane_irregular[s_item.pair[j].slot] | !(rot >= -360.f && rot <= 360.f);
    }
    return irr == 0;
}

#ifndef SDM_LANE3_MINB
#define SDM_LANE3_MINB 10  // 48 registers: the skip walk carries three plane pointers; 12 blocks (40 registers) spill inside the loop (8.68 vs 8.04 ms)
#endif
// kGen = 3: keyframes that qualify for the second-generation loop run the third-generation one (skip planes built)
// Work distribution of the scan kernel: every WARP pulls pieces of 32 consecutive candidates (a quarter of a plan
// chunk) from the plan's counter and keeps its own copy of the keyframe's work order in shared memory, so the kernel
// has no block-wide barrier: with one chunk per block and a __syncthreads per chunk, 8 % of the resident warp time
// was spent waiting for the block's slowest warp (ncu: stall_barrier 1.05 of 13.4 warp-cycles per issue).
constexpr int kPiecesPerChunk = kChunk / 32;
__device__ __forceinline__ int item_words(int n_nbr) { return (int)(offsetof(DevItem, pair) / 4) + n_nbr * (int)(sizeof(DevPair) / 4); }

__device__ __forceinline__ bool next_piece(const DevPlan& plan, const DevItem* __restrict__ items, DevItem& w_item, int lane,
                                           int& cur_entry, int& first)
{
    int id = 0;
    if (lane == 0) id = atomicAdd(plan.counter, 1);
    id = __shfl_sync(SDM_FULL, id, 0);
    const int chunk = id / kPiecesPerChunk;
    if (chunk >= plan.chunk_off[plan.n_items]) return false;
    int lo = 0, hi = plan.n_items;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (plan.chunk_off[mid] <= chunk) lo = mid; else hi = mid;
    }
    if (lo != cur_entry) {
        __syncwarp();  // every lane is done with the previous work order
        const DevItem* src_item = &items[plan.order ? plan.order[lo] : lo];
        const int* src = reinterpret_cast<const int*>(src_item);
        int* dst = reinterpret_cast<int*>(&w_item);
        const int words = item_words(src_item->n_nbr);
        for (int i = lane; i < words; i += 32) dst[i] = src[i];
        cur_entry = lo;
        __syncwarp();
    }
    first = (chunk - plan.chunk_off[lo]) * kChunk + (id % kPiecesPerChunk) * 32;
    return true;
}

// kLong: the build for long scans (hundreds of columns per (pixel, neighbour): large images, wide depth ranges).  There the
// skip walk is bound by the latency of its dependent loads, not by issue slots, and 64 registers / 8 blocks per SM (no
// spill, a larger share of L1 per warp) beat 48 registers / 10 blocks: 27.9 vs 30.4 ms per 48 keyframes of BASELINE config 4
// (7 / 6 / 5 / 4 blocks: 29.1 / 29.1 / 31.6 / 31.7), while config 2 (55 columns per scan) measures 8.46 vs 8.02 ms the other
// way round.  sdm_pass1 picks the instantiation per launch from the mean search range of the batch's pairs.
#ifndef SDM_LANE3_LONG_MINB
#define SDM_LANE3_LONG_MINB 8
#endif
template <bool kFastGates, int kGen = 2, bool kLong = false>
__global__ void __launch_bounds__(kLaneBlock, kGen == 3 ? (kLong ? SDM_LANE3_LONG_MINB : SDM_LANE3_MINB) : SDM_LANE_MINB)
k_pass1_lane(DevArena A, DevParams P, const DevItem* __restrict__ items, DevPlan plan, DevStats* stats, int item_bytes)
{
    // dynamic shared memory, sized at launch for the largest neighbour count of the batch (the rest stays L1):
    // [kLaneBlock / 32 warps][item_bytes] work orders, then [max n_nbr][kLaneBlock] hypotheses
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevItem& w_item = *reinterpret_cast<DevItem*>(s_dyn + (size_t)warp * item_bytes);
    float2 (*s_h)[kLaneBlock] = reinterpret_cast<float2 (*)[kLaneBlock]>(s_dyn + (size_t)(kLaneBlock / 32) * item_bytes);
    int cur_entry = -1, first = 0;
    unsigned n_fused = 0;
    while (next_piece(plan, items, w_item, lane, cur_entry, first)) {
        const int ci = first + lane;
        bool fused = false;
        if (ci < A.cand_count[w_item.kf]) {
            if (kFastGates && P.scan2 && item_regular(A, w_item)) fused = scan_pixel_lane<kGen>(A, P, w_item, s_h, ci, tid);
            else fused = scan_pixel_lane<kFastGates ? 1 : 0>(A, P, w_item, s_h, ci, tid);
        }
        n_fused += __popc(__ballot_sync(SDM_FULL, fused));
    }
    if (stats && lane == 0 && n_fused) atomicAdd(&stats->fused, (unsigned long long)n_fused);
}

// per-pair raw hypotheses for every candidate pixel of kf1 (granularity of one EpipolarSearch call)
__global__ void __launch_bounds__(kPass1Warps * 32)
k_pair_hypotheses(DevArena A, DevParams P, const DevItem* __restrict__ item, float* __restrict__ hyp_d,
                  float* __restrict__ hyp_s, float* __restrict__ hyp_u, float* __restrict__ hyp_v,
                  uint8_t* __restrict__ hyp_ok, int single_xy, float pixel_arg, float th_pi_arg)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kf = item->kf;
    int x, y;
    size_t out;
    if (single_xy >= 0) {
        if (blockIdx.x != 0 || warp != 0) return;
        x = single_xy & 0xffff; y = single_xy >> 16; out = 0;
    } else {
        const int ci = blockIdx.x * kPass1Warps + warp;
        if (ci >= A.cand_count[kf]) return;
        const uint32_t packed = A.cand[(size_t)kf * A.P + ci];
        x = (int)(packed & 0xffffu); y = (int)(packed >> 16);
        out = (size_t)y * P.W + x;
    }
    const DevPair g = item->pair[0];
    const float K[4] = {item->K[0], item->K[1], item->K[2], item->K[3]};
    const float4 t1 = __ldg(&A.tex[(size_t)kf * A.P + (size_t)y * P.W + x]);
    const float gradc = t1.x;
    // EpipolarSearch takes `pixel` and `th_pi` from its caller (:466-470); the plane mode reads them like :457,:466
    const float th_pi = (single_xy >= 0) ? th_pi_arg : t1.z;
    const float pixel = (single_xy >= 0) ? pixel_arg : (float)__ldg(&A.ipair[(size_t)kf * A.P + (size_t)y * P.W + x]).x;
    const float xn = (x - K[2]) / K[0], yn = (y - K[3]) / K[1];
    const float Hm1 = (float)(P.H - 1);
    const PairSetup s = pair_setup(g, K, P, x, y, xn, yn, item->min_depth, item->max_depth, th_pi);
    const size_t nb = (size_t)g.slot * A.P;
    float best_err = 100000.0f;
    int best_u = 0x7fffffff;
    for (int u = s.u_lo + lane; u <= s.u_hi; u += 32) {
        float err, pe, ge;
        if (eval_candidate(A.tex + nb, A.ipair + nb, P, Hm1, u, s.ab, s.cb, s.th_line, s.ang_pi_rot, pixel, gradc,
                           err, pe, ge))
            if (err < best_err) { best_err = err; best_u = u; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float oe = __shfl_xor_sync(SDM_FULL, best_err, off);
        const int ou = __shfl_xor_sync(SDM_FULL, best_u, off);
        if (oe < best_err || (oe == best_err && ou < best_u)) { best_err = oe; best_u = ou; }
    }
    if (lane != 0) return;
    float d = 0.f, sg = 0.f, bu = 0.f, bv = 0.f;
    uint8_t ok = 0;
    if (best_err < 100000.0f) {
        float err, pe, ge;
        eval_candidate(A.tex + nb, A.ipair + nb, P, Hm1, best_u, s.ab, s.cb, s.th_line, s.ang_pi_rot, pixel, gradc,
                       err, pe, ge);
        Hypo h = refine_hypothesis(A.tex + nb, A.ipair + nb, g, K, P, best_u, s.ab, s.cb, pe, ge, xn, yn);
        d = h.depth; sg = h.sigma; bu = h.best_u; bv = h.best_v;
        ok = (1.0f / d > 0.0f) ? 2 : 1;
    }
    hyp_d[out] = d; hyp_s[out] = sg; hyp_u[out] = bu; hyp_ok[out] = ok;
    if (hyp_v) hyp_v[out] = bv;
}

__global__ void k_search_range(DevParams P, const DevItem* __restrict__ item, int px, int py, float* out)
{
    const float* K = item->K;
    const float xn = (px - K[2]) / K[0], yn = (py - K[3]) / K[1];
    float umin, umax;
    search_range(item->pair[0], K, P.W, xn, yn, item->min_depth, item->max_depth, umin, umax);
    out[0] = umin;
    out[1] = umax;
}

// InverseDepthHypothesisFusion for m independent sets (one warp per set)
__global__ void __launch_bounds__(256) k_fuse_sets(DevParams P, int m, int n, const float* __restrict__ depth,
                                                   const float* __restrict__ sigma, const int* __restrict__ count,
                                                   float* __restrict__ out_d, float* __restrict__ out_s,
                                                   int* __restrict__ out_ok)
{
    const int set = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (set >= m) return;
    const int cnt = count[set];
    const bool keep = lane < cnt;
    const float hd = keep ? depth[(size_t)set * n + lane] : 0.f;
    const float hs = keep ? sigma[(size_t)set * n + lane] : 0.f;
    const unsigned vmask = __ballot_sync(SDM_FULL, keep);
    float od = 0.f, os = 0.f;
    int ok = 0;
    unsigned mm = 0;
    for (unsigned rem = vmask; rem; rem &= rem - 1) {
        const int b = __ffs(rem) - 1;
        const float db = __shfl_sync(SDM_FULL, hd, b);
        const float sb = __shfl_sync(SDM_FULL, hs, b);
        if (keep && (b == lane || chi_compatible(hd, db, hs, sb, P.chi_fusion_lt))) mm |= 1u << b;
    }
    int key = keep ? ((__popc(mm) << 8) | (31 - lane)) : 0;
    for (int off = 16; off > 0; off >>= 1) key = max(key, __shfl_xor_sync(SDM_FULL, key, off));
    const int best_n = key >> 8, winner = 31 - (key & 0xff);
    if (best_n > P.lambdaN) {
        const unsigned wm = __shfl_sync(SDM_FULL, mm, winner);
        double q1 = 0.0, q2 = 0.0;
        if (keep) { const double s2 = (double)hs * (double)hs; q1 = (double)hd / s2; q2 = 1.0 / s2; }
        float pjsj = 0.f, rsj = 0.f;
        for (unsigned rem = wm; rem; rem &= rem - 1) {
            const int b = __ffs(rem) - 1;
            pjsj = (float)((double)pjsj + __shfl_sync(SDM_FULL, q1, b));
            rsj = (float)((double)rsj + __shfl_sync(SDM_FULL, q2, b));
        }
        od = pjsj / rsj; os = sqrtf(1.0f / rsj); ok = 1;
    }
    if (lane == 0) { out_d[set] = od; out_s[set] = os; out_ok[set] = ok; }
}

// ---------------------------------------------------------------------------------------------
// K5: IntraKeyFrameDepthChecking (:866-927) and IntraKeyFrameDepthGrowing (:929-976): 3x3 Jacobi
// stencils on the (rho, sigma) planes.  Two stages that ping-pong between two planes (no snapshot
// copies): "check" reads plane A and writes plane B, "grow" reads B and writes A (+ the dense
// depth / sigma copies).  copy_only = 1 turns a stage into a plain copy (only one of the two enabled).
//   * candidate-list kernels (default, planes produced by pass 1): (rho,sigma) is non-zero only at
//     candidate pixels, so both stages visit the compacted candidates of the batch (dense warps, same
//     work plan as pass 1); plane B is the slot's own second plane rs2, zero outside the candidates.
//     Growing is evaluated on the candidates only: a non-candidate pixel holds (0,0), and a centre with
//     sigma = 0 can never pass ChiTest (num/0 is inf or NaN), so it cannot grow.
//   * dense kernels (planes written from outside: sdm_upload_depth, single-method entry points):
//     every pixel of the plane, grid.z = keyframe, plane B = tmp + z * npix.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 intra_check_pixel(const DevParams& P, const float2* __restrict__ src, int px, int py, float2 c)
{
    float pjsj = 0.f, rsj = 0.f, min_sigma = 0.f;
    int n = 0;
    float2 nb[8];  // the eight neighbours requested up front: the stencil is latency-bound
    {
        int k = 0;
#pragma unroll
        for (int y = py - 1; y <= py + 1; ++y)
#pragma unroll
            for (int x = px - 1; x <= px + 1; ++x)
                if (!(x == px && y == py)) nb[k++] = src[(size_t)y * P.W + x];
    }
    int k = 0;
#pragma unroll
    for (int y = py - 1; y <= py + 1; ++y)
#pragma unroll
        for (int x = px - 1; x <= px + 1; ++x) {
            if (x == px && y == py) continue;
            const float2 q = nb[k++];
            if (q.x > P.eps_gt && chi_compatible(q.x, c.x, q.y, c.y, P.chi_fusion_lt)) {
                if (n == 0) min_sigma = q.y;
                fusion_accumulate(q.x, q.y, pjsj, rsj);
                // pow(sigma,2) < pow(min,2) in double (squares of floats are exact there)  <=>  |sigma| < |min|
                if (fabsf(q.y) < fabsf(min_sigma)) min_sigma = q.y;
                ++n;
            }
        }
    if (n == 0) min_sigma = c.y;
    fusion_accumulate(c.x, c.y, pjsj, rsj);  // "dont forget itself" :902 (pushed last)
    if (fabsf(c.y) < fabsf(min_sigma)) min_sigma = c.y;
    ++n;
    return (n >= 3) ? make_float2(pjsj / rsj, min_sigma) : make_float2(0.f, 0.f);
}

__device__ __forceinline__ float2 intra_grow_pixel(const DevParams& P, const float2* __restrict__ src, int px, int py, float2 c)
{
    // A centre with sigma = +-0 cannot grow, whatever its neighbours hold: ChiTest (:1641-1645) adds num / (0 * 0),
    // which is +inf (num > 0) or NaN (num = 0 or NaN), to a term that is >= 0, +inf or NaN, and neither inf nor NaN
    // is < 5.99.  Every empty pixel that pass 1 / the check stage leave behind is (0, 0), so the
    // stage keeps its stencil for centres with a non-zero sigma only; exact for any plane contents.
    if (c.y == 0.f) return c;
    float pjsj = 0.f, rsj = 0.f, min_sigma = 0.f;
    int n = 0;
    float2 nb[8];
    {
        int k = 0;
#pragma unroll
        for (int y = py - 1; y <= py + 1; ++y)
#pragma unroll
            for (int x = px - 1; x <= px + 1; ++x)
                if (!(x == px && y == py)) nb[k++] = src[(size_t)y * P.W + x];
    }
    int k = 0;
#pragma unroll
    for (int y = py - 1; y <= py + 1; ++y)
#pragma unroll
        for (int x = px - 1; x <= px + 1; ++x) {
            if (x == px && y == py) continue;
            const float2 q = nb[k++];
            if (chi_compatible(q.x, c.x, q.y, c.y, P.chi_fusion_lt)) {
                if (n == 0) min_sigma = q.y;
                fusion_accumulate(q.x, q.y, pjsj, rsj);
                if (q.y < min_sigma) min_sigma = q.y;
                ++n;
            }
        }
    return (n >= 2) ? make_float2(pjsj / rsj, min_sigma) : c;
}

__device__ __forceinline__ bool interior2(const DevParams& P, int px, int py)
{
    return px >= 2 && py >= 2 && px < P.W - 2 && py < P.H - 2;
}

__global__ void __launch_bounds__(256)
k_intra_check(DevParams P, const float2* __restrict__ arena, float2* __restrict__ tmp, const int* __restrict__ slots,
              size_t npix, int copy_only)
{
    const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
    if (px >= P.W || py >= P.H) return;
    const float2* __restrict__ src = arena + (size_t)slots[blockIdx.z] * npix;
    float2* __restrict__ dst = tmp + (size_t)blockIdx.z * npix;
    const size_t pi = (size_t)py * P.W + px;
    const float2 c = src[pi];
    float2 out = c;
    if (!copy_only && interior2(P, px, py) && c.x > P.eps_gt) out = intra_check_pixel(P, src, px, py, c);
    dst[pi] = out;
}

__global__ void __launch_bounds__(256)
k_intra_grow(DevParams P, const float2* __restrict__ tmp, float2* __restrict__ arena, float* __restrict__ dpl_arena,
             float* __restrict__ spl_arena, const float4* __restrict__ tex_arena, const int* __restrict__ slots, size_t npix,
             int copy_only)
{
    const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
    if (px >= P.W || py >= P.H) return;
    const size_t slot = (size_t)slots[blockIdx.z];
    const float2* __restrict__ src = tmp + (size_t)blockIdx.z * npix;
    float2* __restrict__ dst = arena + slot * npix;
    const size_t pi = (size_t)py * P.W + px;
    const float2 c = src[pi];
    float2 out = c;
    if (!copy_only && interior2(P, px, py) && c.x < P.eps_lt && !(tex_arena[slot * npix + pi].x <= P.lambdaG))
        out = intra_grow_pixel(P, src, px, py, c);
    dst[pi] = out;
    dpl_arena[slot * npix + pi] = out.x;
    spl_arena[slot * npix + pi] = out.y;
}

// candidate-list stages; stage = 0: check (rs -> rs2), stage = 1: grow (rs2 -> rs, dpl, spl)
#ifndef SDM_INTRA_MINB
#define SDM_INTRA_MINB 16  // 32 registers, full occupancy: the stencil gathers are latency-bound (1.22 -> 0.94 ms per 200 keyframes)
#endif
__global__ void __launch_bounds__(kChunk, SDM_INTRA_MINB)
k_intra_cand(DevArena A, DevParams P, const DevItem* __restrict__ items, DevPlan plan, int stage, int copy_only)
{
    __shared__ DevItem s_item;
    __shared__ int s_chunk;
    int cur_entry = -1, first = 0;
    while (next_chunk(plan, items, s_item, s_chunk, cur_entry, first)) {
        const int ci = first + threadIdx.x;
        const int kf = s_item.kf;
        if (ci >= A.cand_count[kf]) continue;
        const size_t base = (size_t)kf * A.P;
        const uint32_t packed = A.cand[base + ci];
        const int px = (int)(packed & 0xffffu), py = (int)(packed >> 16);
        const size_t pi = (size_t)py * P.W + px;
        if (stage == 0) {
            const float2* __restrict__ src = A.rs + base;
            const float2 c = src[pi];
            float2 out = c;
            if (!copy_only && interior2(P, px, py) && c.x > P.eps_gt) out = intra_check_pixel(P, src, px, py, c);
            A.rs2[base + pi] = out;
        } else {
            const float2* __restrict__ src = A.rs2 + base;
            const float2 c = src[pi];
            float2 out = c;
            // (GradImg > lambdaG holds for every candidate, :942)
            if (!copy_only && interior2(P, px, py) && c.x < P.eps_lt) out = intra_grow_pixel(P, src, px, py, c);
            A.rs[base + pi] = out;
            A.dpl[base + pi] = out.x;
            A.spl[base + pi] = out.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K6: InterKeyFrameDepthChecking (:1121-1296) fused with UpdateSemiDensePointSet (:700-731).
// ---------------------------------------------------------------------------------------------
// The chi-square gate of InterKeyFrameDepthChecking (:1204-1211 and its three copies):
//     test = (float)(dd*dd / (sigma*sigma))  evaluated in double with dd = (double)(float)(depthj - d);  accept iff test < 3.84.
// Decided in float whenever the float evaluation is at least 2^-16 (relative) away from the threshold: with a2 = a*a and
// s2 = sigma*sigma both normal (in [2^-100, 2^100]) the float products are within 3 * 2^-24 of the exact quotient test
// and the reference's own value is within 2^-24 + 2^-53 of it, so `a2 < T(1 - 2^-16) * s2` implies acceptance and
// `a2 > T(1 + 2^-16) * s2` implies rejection.  Anything closer, out of range or NaN takes the reference's double form.
__device__ __forceinline__ bool chi_inter_accept(float a, float sg, const DevParams& P)
{
    const float a2 = a * a, s2 = sg * sg;
    if (fminf(a2, s2) >= 0x1p-100f && fmaxf(a2, s2) <= 0x1p100f) {
        if (a2 < P.chi_inter_lo * s2) return true;
        if (a2 > P.chi_inter_hi * s2) return false;
    }
    const double dd = (double)a;
    const float test = (float)((dd * dd) / ((double)sg * (double)sg));
    return test < P.chi_inter_lt;
}

__global__ void k_chi_inter(DevParams P, int n, const float* __restrict__ diff, const float* __restrict__ sigma,
                            uint8_t* __restrict__ accept)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) accept[i] = chi_inter_accept(diff[i], sigma[i], P) ? 1 : 0;
}

__device__ __forceinline__ void write_point(const DevItem& it, const DevParams& P, float* __restrict__ pts, size_t pi,
                                            int x, int y, float checked)
{
    float X = 0.f, Y = 0.f, Z = 0.f;
    if (!(checked < P.eps_lt)) {
        const float z = 1.0f / checked;
        const float xc = z * (x - it.K[2]) / it.K[0];
        const float yc = z * (y - it.K[3]) / it.K[1];
        const float* T = it.Twc;
        X = T[0] * xc + T[1] * yc + T[2] * z + T[3] * 1.0f;
        Y = T[4] * xc + T[5] * yc + T[6] * z + T[7] * 1.0f;
        Z = T[8] * xc + T[9] * yc + T[10] * z + T[11] * 1.0f;
    }
    pts[3 * pi + 0] = X;
    pts[3 * pi + 1] = Y;
    pts[3 * pi + 2] = Z;
}

// one interior pixel with rho = depthp (not < eps): :1159-1293.  Returns depth_map_checked_.
__device__ __forceinline__ float inter_check_pixel(const DevArena& A, const DevParams& P, const DevItem& it, int px, int py,
                                                   float depthp)
{
    const int cols = P.W, rows = P.H;
    const float fx = it.K[0], fy = it.K[1], cx = it.K[2], cy = it.K[3];
    const float xn = (px - cx) / fx, yn = (py - cy) / fy;
    const float dp = 1.0f / depthp;
    const double alpha = (double)dp;  // gemm alpha = (float)(1.0 / (double)depthp) == 1.0f / depthp (same argument as iz below)
    int support = 0;
    double JtR = 0.0, JtJ = 0.0;
    const int N = it.n_nbr;
    for (int j = 0; j < N; ++j) {
        const DevPair& g = it.pair[j];
        // temp = Rji*xp/depthp + tji ; Xj = K*temp ; Xj/Xj(2)
        const float s0 = g.R[0] * xn + g.R[1] * yn + g.R[2] * 1.0f;
        const float s1 = g.R[3] * xn + g.R[4] * yn + g.R[5] * 1.0f;
        const float s2 = g.R[6] * xn + g.R[7] * yn + g.R[8] * 1.0f;
        const float X0 = (float)((double)s0 * alpha + (double)g.t[0]);
        const float X1 = (float)((double)s1 * alpha + (double)g.t[1]);
        const float X2 = (float)((double)s2 * alpha + (double)g.t[2]);
        const float U = g.K2[0] * X0 + 0.f * X1 + g.K2[2] * X2;
        const float V = 0.f * X0 + g.K2[1] * X1 + g.K2[3] * X2;
        const float Wz = 0.f * X0 + 0.f * X1 + 1.0f * X2;
        // (float)(1.0 / (double)Wz) == 1.0f / Wz: rounding a double quotient of two floats to float cannot differ
        // from the float quotient (53 >= 2*24 + 2: double rounding is innocuous for division, Figueroa 1995)
        const float iz = 1.0f / Wz;
        const float xj = U * iz, yj = V * iz;
        // Eq (12)
        const float rz = (float)((double)g.R[6] * (double)xn + (double)g.R[7] * (double)yn + (double)g.R[8] * 1.0);
        const float depthj = depthp / (rz + depthp * g.t[2]);
        if (!(xj >= 0.f && xj < (float)(cols - 1) && yj >= 0.f && yj < (float)(rows - 1))) continue;
        const int x0 = (int)floorf(xj), y0 = (int)floorf(yj);
        const float2* nrs = A.rs + (size_t)g.slot * A.P;
        const float2 q00 = __ldg(&nrs[(size_t)y0 * cols + x0]);
        const float2 q10 = __ldg(&nrs[(size_t)(y0 + 1) * cols + x0]);
        const float2 q01 = __ldg(&nrs[(size_t)y0 * cols + x0 + 1]);
        const float2 q11 = __ldg(&nrs[(size_t)(y0 + 1) * cols + x0 + 1]);
        const float2 q[4] = {q00, q10, q01, q11};  // (y0,x0) (y1,x0) (y0,x1) (y1,x1)
        int nj = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d = q[k].x, sg = q[k].y;
            if (d > P.eps_gt) {
                if (chi_inter_accept(depthj - d, sg, P)) {
                    ++nj;
                    // Gauss-Newton terms (:1274-1280), accumulated in the reference's order
                    const float djn = 1.0f / d;
                    const float d2sigma = djn * djn * sg;
                    const float Ji = -rz / d2sigma;
                    const float ri = (djn - dp * rz - g.t[2]) / d2sigma;
                    JtR += (double)Ji * (double)ri;
                    JtJ += (double)Ji * (double)Ji;
                }
            }
        }
        if (nj >= 1) ++support;
    }
    if (support < P.lambdaN) return 0.f;
    const float Jtr0 = (float)(JtR * -1.0);
    const float JtJf = (float)(JtJ * 1.0);
    const float dpDelta = Jtr0 / JtJf;
    return 1.0f / (dp + dpDelta);
}

// dense variant: one thread per pixel of the plane (grid.y = item).  Used for keyframes whose (rho,sigma)
// plane did not come from pass 1 (sdm_upload_depth, external receives) and for points_only.
__global__ void __launch_bounds__(256)
k_pass2(DevArena A, DevParams P, const DevItem* __restrict__ items, const int* __restrict__ order, DevStats* stats,
        int points_only)
{
    __shared__ DevItem s_item;
    load_item(s_item, &items[order ? order[blockIdx.y] : blockIdx.y]);
    __syncthreads();
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= A.P) return;
    const int py = (int)(p / P.W), px = (int)(p % P.W);
    const int kf = s_item.kf;
    const size_t base = (size_t)kf * A.P;
    const bool interior = (px >= 2 && py >= 2 && px < P.W - 2 && py < P.H - 2);
    if (points_only) {
        if (interior) write_point(s_item, P, A.pts + 3 * base, p, px, py, A.chk[base + p]);
        return;
    }
    float checked = 0.f;
    if (interior) {
        const float depthp = A.rs[base + p].x;
        if (!(depthp < P.eps_lt)) checked = inter_check_pixel(A, P, s_item, px, py, depthp);
    }
    A.chk[base + p] = checked;
    write_point(s_item, P, A.pts + 3 * base, p, px, py, interior ? checked : 0.f);
    if (stats) {
        const unsigned bal = __ballot_sync(__activemask(), checked > 0.f);
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&stats->checked, (unsigned long long)__popc(bal));
    }
}

// candidate-list variant (the default): pass 1 writes (rho,sigma) only at candidate pixels and k_pack zeroes
// chk / pts of the whole plane, so pass 2 only has to visit the compacted candidates: one thread per
// candidate, kLaneBlock candidates of one keyframe per block (same block -> item map as pass 1).
#ifndef SDM_P2_MINB
#define SDM_P2_MINB 16  // 32 registers, full occupancy (occupancy-bound gathers; 72 / 48 / 40 / 32 registers: 1.49 / 1.37 / 1.33 / 1.33 ms)
#endif
__global__ void __launch_bounds__(kLaneBlock, SDM_P2_MINB)
k_pass2_cand(DevArena A, DevParams P, const DevItem* __restrict__ items, DevPlan plan, DevStats* stats)
{
    __shared__ DevItem s_item;
    __shared__ int s_chunk;
    int cur_entry = -1, first = 0;
    unsigned n_checked = 0;
    while (next_chunk(plan, items, s_item, s_chunk, cur_entry, first)) {
        const int ci = first + threadIdx.x;
        const int kf = s_item.kf;
        float checked = 0.f;
        if (ci < A.cand_count[kf]) {
            const size_t base = (size_t)kf * A.P;
            const uint32_t packed = A.cand[base + ci];
            const int px = (int)(packed & 0xffffu), py = (int)(packed >> 16);
            if (px >= 2 && py >= 2 && px < P.W - 2 && py < P.H - 2) {
                const size_t p = (size_t)py * P.W + px;
                const float depthp = A.rs[base + p].x;
                if (!(depthp < P.eps_lt)) checked = inter_check_pixel(A, P, s_item, px, py, depthp);
                A.chk[base + p] = checked;
                write_point(s_item, P, A.pts + 3 * base, p, px, py, checked);
            }
        }
        n_checked += __popc(__ballot_sync(SDM_FULL, checked > 0.f));
    }
    if (stats && (threadIdx.x & 31) == 0 && n_checked) atomicAdd(&stats->checked, (unsigned long long)n_checked);
}

// ---------------------------------------------------------------------------------------------
// Point-cloud export: the filter every consumer of the semi-dense map applies —
// SaveSemiDensePoints (:167-168), MapDrawer::DrawSemiDense (MapDrawer.cc:104-106), the CARV entry
// (SFMTranscriptInterface_ORBSLAM.cpp:272) — `skip if depth_sigma_ > sigma_max; keep if
// depth_map_checked_ > 0.000001`, as a stream compaction on the device: keyframes in list order,
// pixels in raster order (the order the reference's loops emit), so only the surviving points
// cross PCIe instead of four dense planes.  Three steps: per-block counts, one-block exclusive scan,
// ordered scatter.  A block handles kExportBlock consecutive pixels of one keyframe.
// ---------------------------------------------------------------------------------------------
constexpr int kExportBlock = 1024;

__device__ __forceinline__ bool export_keep(const DevArena& A, const DevParams& P, size_t base, size_t p, float sigma_gt)
{
    const float sg = A.rs[base + p].y;
    if (sg > sigma_gt) return false;  // (double)sigma > sigma_max, as a float threshold
    return A.chk[base + p] > P.eps_gt;
}

__global__ void __launch_bounds__(kExportBlock)
k_export_count(DevArena A, DevParams P, const int* __restrict__ slots, float sigma_gt, int blocks_per_kf, int* __restrict__ counts)
{
    __shared__ int s_warp[32];
    const int kfi = blockIdx.x / blocks_per_kf, blk = blockIdx.x % blocks_per_kf;
    const size_t base = (size_t)slots[kfi] * A.P;
    const size_t p = (size_t)blk * kExportBlock + threadIdx.x;
    const bool keep = p < A.P && export_keep(A, P, base, p, sigma_gt);
    const unsigned bal = __ballot_sync(SDM_FULL, keep);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = s_warp[threadIdx.x];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(SDM_FULL, v, off);
        if (threadIdx.x == 0) counts[blockIdx.x] = v;
    }
}

// exclusive scan of n block counts (one block); offsets[n] = total; per-keyframe totals
__global__ void __launch_bounds__(1024)
k_export_scan(const int* __restrict__ counts, int n, int blocks_per_kf, unsigned long long* __restrict__ offsets,
              unsigned long long* __restrict__ kf_totals)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0ULL;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        const unsigned long long c = i < n ? (unsigned long long)counts[i] : 0ULL;
        unsigned long long incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long v = __shfl_up_sync(SDM_FULL, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long v = __shfl_up_sync(SDM_FULL, w, off);
                if (lane >= off) w += v;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        if (i < n) offsets[i] = s_base + (warp ? s_warp[warp - 1] : 0ULL) + incl - c;
        __syncthreads();
        if (tid == 0) s_base += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) offsets[n] = s_base;
    __syncthreads();
    // per-keyframe totals from the offsets of each keyframe's first block
    const int n_kf = n / blocks_per_kf;
    for (int k = tid; k < n_kf; k += 1024)
        kf_totals[k] = offsets[(k + 1) * blocks_per_kf] - offsets[k * blocks_per_kf];
}

__global__ void __launch_bounds__(kExportBlock)
k_export_scatter(DevArena A, DevParams P, const int* __restrict__ slots, float sigma_gt, int blocks_per_kf,
                 const unsigned long long* __restrict__ offsets, sdm_point* __restrict__ out, unsigned long long capacity)
{
    __shared__ int s_warp[32];
    const int kfi = blockIdx.x / blocks_per_kf, blk = blockIdx.x % blocks_per_kf;
    const size_t base = (size_t)slots[kfi] * A.P;
    const size_t p = (size_t)blk * kExportBlock + threadIdx.x;
    const bool keep = p < A.P && export_keep(A, P, base, p, sigma_gt);
    const unsigned bal = __ballot_sync(SDM_FULL, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = s_warp[threadIdx.x];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(SDM_FULL, v, off);
            if (lane >= off) v += o;
        }
        s_warp[threadIdx.x] = v;  // inclusive over warps
    }
    __syncthreads();
    if (keep) {
        const unsigned long long pos = offsets[blockIdx.x] + (unsigned long long)((warp ? s_warp[warp - 1] : 0) +
                                                                                   __popc(bal & ((1u << lane) - 1u)));
        if (pos < capacity) {
            sdm_point q;
            q.x = A.pts[3 * (base + p) + 0];
            q.y = A.pts[3 * (base + p) + 1];
            q.z = A.pts[3 * (base + p) + 2];
            q.pixel = ((uint32_t)(p / P.W) << 16) | (uint32_t)(p % P.W);
            out[pos] = q;
        }
    }
}

// (rho, sigma) float2 plane -> two dense float planes (download staging), and back
// ---------------------------------------------------------------------------------------------
// Sparse result of a slot for sdm_scatter_keyframes: SemiDenseLoop only ever writes the candidate pixels of a
// keyframe's four output planes non-zero (:483-484, :1290, :725-727 all sit behind the candidate test of
// :454-456 / depth_map_ > 0), so the n candidates' values are the whole result.  SoA record block:
// pix[n] (y << 16 | x) | rho[n] | sigma[n] | checked[n] | points[3n].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_sparse(DevArena A, DevParams P, int slot, int n, uint32_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t base = (size_t)slot * A.P;
    const uint32_t p = A.cand[base + i];
    const size_t pi = base + (size_t)(p >> 16) * P.W + (p & 0xffffu);
    const float2 rs = A.rs[pi];
    float* f = reinterpret_cast<float*>(out);
    out[i] = p;
    f[(size_t)n + i] = rs.x;
    f[2 * (size_t)n + i] = rs.y;
    f[3 * (size_t)n + i] = A.chk[pi];
    f[4 * (size_t)n + 3 * (size_t)i + 0] = A.pts[3 * pi + 0];
    f[4 * (size_t)n + 3 * (size_t)i + 1] = A.pts[3 * pi + 1];
    f[4 * (size_t)n + 3 * (size_t)i + 2] = A.pts[3 * pi + 2];
}

// ---------------------------------------------------------------------------------------------
// Block-sparse download straight into the caller's PINNED host planes (sdm_scatter_keyframes, sdm_run_loop with
// sparse_download): for every row only the 16-pixel blocks that hold a candidate (A.blk) are written, with coalesced
// stores over PCIe; everything else of the destination keeps its zeros (KeyFrame.cc:78-81).  On the bench scene 53 % of
// the blocks hold a candidate, i.e. 47 % of the D2H bytes of a dense DMA never move.  One warp per (keyframe, row):
// two blocks per step for the 4-byte planes (lanes 0-15 / 16-31), one block of 48 floats per 1.5 steps for the points.
// ---------------------------------------------------------------------------------------------
constexpr int kSparseBatch = 16;
struct SparseDst {
    int slot;
    float *depth, *sigma, *checked, *points;  // device-visible addresses of the pinned host planes (or nullptr)
    size_t depth_step, sigma_step, checked_step, points_step;
};
struct SparseBatch {
    SparseDst d[kSparseBatch];
};
// One warp (= one block: small enough to slip in next to the resident blocks of a persistent pass kernel) per
// (keyframe, row); it walks the row in 32-pixel chunks (two blocks), skips chunks whose two mask bits are clear, and
// every lane moves its own pixel of the 4-byte planes and three floats of the chunk's 96 point floats.
__global__ void __launch_bounds__(32) k_sparse_rows(DevArena A, DevParams P, SparseBatch B)
{
    const int lane = threadIdx.x;
    const int y = blockIdx.x;
    const SparseDst& D = B.d[blockIdx.y];
    const size_t src = (size_t)D.slot * A.P + (size_t)y * P.W;
    const uint32_t* mrow = A.blk + ((size_t)D.slot * P.H + y) * A.blk_words;
    float* dd = D.depth ? reinterpret_cast<float*>(reinterpret_cast<char*>(D.depth) + (size_t)y * D.depth_step) : nullptr;
    float* ds = D.sigma ? reinterpret_cast<float*>(reinterpret_cast<char*>(D.sigma) + (size_t)y * D.sigma_step) : nullptr;
    float* dc = D.checked ? reinterpret_cast<float*>(reinterpret_cast<char*>(D.checked) + (size_t)y * D.checked_step) : nullptr;
    float* dp = D.points ? reinterpret_cast<float*>(reinterpret_cast<char*>(D.points) + (size_t)y * D.points_step) : nullptr;
    const int nchunk = (P.W + 31) / 32;
    unsigned word = 0;
    for (int c = 0; c < nchunk; ++c) {
        if ((c & 15) == 0) word = mrow[c >> 4];  // 32 blocks = 16 chunks per mask word
        const unsigned two = (word >> (2 * (c & 15))) & 3u;
        if (!two) continue;
        const int x = c * 32 + lane;
        if (((two >> (lane >> 4)) & 1u) && x < P.W) {
            if (dd) dd[x] = A.dpl[src + x];
            if (ds) ds[x] = A.spl[src + x];
            if (dc) dc[x] = A.chk[src + x];
        }
        if (dp) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int f = k * 32 + lane;            // float f of the chunk's 96 belongs to pixel f / 3
                const int px = c * 32 + f / 3;
                if (((two >> (f / 48)) & 1u) && px < P.W) dp[3 * c * 32 + f] = A.pts[3 * (src + c * 32) + f];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Cross-GPU ordering of the inter-pass exchange without the host (sdm_exchange).  Every rank owns one XFlags block
// in device memory that its peers map through CUDA IPC.  Owner: `done` = number of the last step whose pass-1 planes
// are complete (written by a kernel that follows pass 1 on the compute stream).  A peer that pulls halo planes
// spins on the owner's `done`, copies, then stores the step number into ack[its rank] of the owner; the owner's NEXT
// pass 1 spins on the acks of every registered puller before it overwrites the planes (sdm_pass1).
// Spins sleep between polls and give up after kXWaitNs (the context reports SDM_ERR_STATE at the next synchronise).
// ---------------------------------------------------------------------------------------------
constexpr int kXPeers = 16;
struct XFlags {
    unsigned done;
    unsigned puller_mask;
    unsigned ack[kXPeers];
    unsigned err;  // != 0: a wait of THIS rank timed out
    unsigned pad[13];
};
constexpr unsigned long long kXWaitNs = 20ULL * 1000 * 1000 * 1000;

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool spin_until_geq(const volatile unsigned* w, unsigned want)
{
    const unsigned long long t0 = global_ns();
    // (int) difference: robust against wrap-around of the step counter
    while ((int)(*w - want) < 0) {
        __nanosleep(500);
        if (global_ns() - t0 > kXWaitNs) return false;
    }
    __threadfence_system();
    return true;
}
__global__ void k_xpublish(XFlags* mine, unsigned step)
{
    __threadfence_system();
    *(volatile unsigned*)&mine->done = step;
    __threadfence_system();
}
__global__ void k_xwait_done(const XFlags* peer, unsigned step, XFlags* mine)
{
    if (!spin_until_geq(&peer->done, step)) mine->err = 1;
}
__global__ void k_xack(XFlags* peer, int my_rank, unsigned step)
{
    __threadfence_system();
    *(volatile unsigned*)&peer->ack[my_rank] = step;
    __threadfence_system();
}
__global__ void k_xwait_acks(XFlags* mine, unsigned step)
{
    const int r = threadIdx.x;
    if (r < kXPeers && ((*(volatile unsigned*)&mine->puller_mask >> r) & 1u))
        if (!spin_until_geq(&mine->ack[r], step)) mine->err = 2;
}
__global__ void k_xregister(XFlags* peer, int my_rank) { atomicOr_system(&peer->puller_mask, 1u << my_rank); }

// candidate counts of freshly packed slots into host-mapped pinned memory (no DMA engine involved)
constexpr int kSlotList = 24;
struct SlotList {
    int s[kSlotList];
};
__global__ void k_publish_counts(const int* __restrict__ cand_count, SlotList slots, int m, int* host_counts)
{
    const int i = threadIdx.x;
    if (i < m) host_counts[slots.s[i]] = cand_count[slots.s[i]];
    __threadfence_system();
}

__global__ void k_split_rs(const float2* __restrict__ rs, float* __restrict__ d, float* __restrict__ s, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float2 v = rs[i]; d[i] = v.x; s[i] = v.y; }
}
__global__ void k_merge_rs(float2* __restrict__ rs, const float* __restrict__ d, const float* __restrict__ s, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rs[i] = make_float2(d[i], s[i]);
}

// ---------------------------------------------------------------------------------------------
// edge-aided 3-D line fitting, the immediate consumer of depth_map_checked_ / depth_sigma_ -
// LineDetector::LineFit and its helpers (LineDetector.cc:578-840; constants :20-29) over the edge chains of a
// keyframe (LineFitting, :884-900), reading the planes where pass 2 left them.  One thread per chain (the algorithm
// is a greedy walk along the chain; its tail recursion at :838 is a loop here).  The two least-squares problems are
// solved in closed form in double precision:
//   LeastSquaresLineFit (:601-624, cv::SVD::solveZ): the unit vector (a, b, c) minimising sum (a x + b y + c)^2 = the
//     eigenvector of the smallest eigenvalue of the 3x3 moment matrix (cyclic Jacobi on exact integer sums);
//   LeastSquaresDepthFit (:627-678, cv::solve DECOMP_SVD): z = u1 d + u2 by the 2x2 normal equations (minimum-norm
//     solution when all d coincide).
// OpenCV solves both with a float Jacobi SVD, so decisions that sit on a threshold can differ: parity with the oracle
// (oracle/linefit_oracle.py, real cv2 calls) is tolerance-level - tests/test_gpu_linefit.py bounds the share of
// differing segments and the end-point error.  Everything else is float arithmetic in the reference's order.
// ---------------------------------------------------------------------------------------------
struct LineFitParams {
    int min_len, max_len, init_depth_count;  // MIN_LINE_LENGTH 10, MAX_LINE_LENGTH 1000, INIT_DEPTH_COUNT 3
    float min_angle, e1, e2;                 // MIN_SEGMENT_ANGLE 30, E1 1.0, E2 1.5
    float sigma_lt;                          // x < this <=> x < sigma_limit (0.02f)
    float fx, fy, cx, cy;                    // filled per keyframe from LineFitKf
    float Twc[12];
};
struct LineFitKf {  // one keyframe of a sdm_line_fit batch
    int slot;
    int chain0;     // first chain of the keyframe in the batch's chain list
    float K[4];
    float Twc[12];
};
struct DevLine {    // = sdm_line3d
    float seg[4];   // s.x s.y e.x e.y   (mLinesSeg row)
    float xyz[6];   // Pws, Pwe          (mLines3D row)
    int chain;      // chain index inside its keyframe
    int kf_index;   // index of the keyframe in the batch
};
static_assert(sizeof(DevLine) == sizeof(sdm_line3d), "DevLine is the ABI's sdm_line3d");

__device__ __forceinline__ void lf_closest(float a, float b, float c, int x, int y, float& cx, float& cy)
{
    const float den = a * a + b * b;
    cx = (b * (b * (float)x - a * (float)y) - a * c) / den;
    cy = (a * (-b * (float)x + a * (float)y) - b * c) / den;
}
__device__ __forceinline__ float lf_norm2(float px, float py, float qx, float qy)
{
    const float dx = px - qx, dy = py - qy;
    return (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
}
struct LfPlanes {
    const float* chk;
    const float2* rs;   // depth_sigma_ = rs[].y (the authoritative pass-1 plane)
    int W;
    float eps_gt, sigma_lt;
};
__device__ __forceinline__ bool lf_has_depth(const LfPlanes& Q, uint32_t rc)
{
    const size_t i = (size_t)(rc >> 16) * Q.W + (rc & 0xffffu);
    return Q.chk[i] > Q.eps_gt && Q.rs[i].y < Q.sigma_lt;
}
__device__ __forceinline__ int lf_count(const LfPlanes& Q, const uint32_t* ch, int n)
{
    int k = 0;
    for (int i = 0; i < n; ++i) k += lf_has_depth(Q, ch[i]);
    return k;
}
// smallest eigenpair of the symmetric 3x3 M (cyclic Jacobi, double): the robust form, used where the direct one below declines
__device__ __noinline__ void lf_smallest_eigvec_jacobi(double M[3][3], double v[3], double& lam)
{
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(M[0][1]) + fabs(M[0][2]) + fabs(M[1][2]);
        if (off < 1e-300) break;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {  // (compile-time indices: the matrices stay in registers)
                if (M[p][q] == 0.0) continue;
                const double theta = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double mkp = M[k][p], mkq = M[k][q];
                    M[k][p] = cs * mkp - sn * mkq;
                    M[k][q] = sn * mkp + cs * mkq;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double mpk = M[p][k], mqk = M[q][k];
                    M[p][k] = cs * mpk - sn * mqk;
                    M[q][k] = sn * mpk + cs * mqk;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    // the smallest diagonal entry (first one on ties), its eigenvector = that column of V
    const bool m1 = M[1][1] < M[0][0];
    const double d01 = m1 ? M[1][1] : M[0][0];
    const bool m2 = M[2][2] < d01;
    lam = m2 ? M[2][2] : d01;
#pragma unroll
    for (int r = 0; r < 3; ++r) v[r] = m2 ? V[r][2] : (m1 ? V[r][1] : V[r][0]);
}
// The same eigenvector without iterating: the smallest root of the characteristic cubic in its trigonometric form, one
// Rayleigh-quotient correction of it, and the eigenvector as the largest cross product of two rows of M - lambda I.  For
// windows of pixel chains (eigenvalues ~ 1e-4 / 25 / 5e5, relative gap between the two small ones >= 7e-7) the unit vector
// agrees with Jacobi's / LAPACK's to 5e-12, i.e. (a, b, c) are the same floats except for an occasional last bit; ten times
// fewer double-precision operations than the sweeps (which dominated k_line_fit: 2.59 -> 0.54 ms per 32 VGA keyframes).  When
// the two small eigenvalues (nearly) coincide - repeated pixels, fewer than two distinct points - the cross products
// vanish and Jacobi takes over.  The sign of the vector is arbitrary here as in any SVD; every use is sign-invariant.
__device__ __forceinline__ void lf_smallest_eigvec(double M[3][3], double v[3], double& lam)
{
    const double m00 = M[0][0], m01 = M[0][1], m02 = M[0][2], m11 = M[1][1], m12 = M[1][2], m22 = M[2][2];
    const double p1 = m01 * m01 + m02 * m02 + m12 * m12;
    const double q = (m00 + m11 + m22) / 3.0;
    const double b00 = m00 - q, b11 = m11 - q, b22 = m22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    bool ok = p2 > 0.0;
    if (ok) {
        const double pp = sqrt(p2 / 6.0), ip = 1.0 / pp;
        const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = m01 * ip, c02 = m02 * ip, c12 = m12 * ip;
        double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
        r = fmin(1.0, fmax(-1.0, r));
        double l = q + 2.0 * pp * cos(acos(r) / 3.0 + 2.0943951023931954923084289221863);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const double r0x = m00 - l, r0y = m01, r0z = m02, r1x = m01, r1y = m11 - l, r1z = m12, r2x = m02, r2y = m12, r2z = m22 - l;
            double ax = r0y * r1z - r0z * r1y, ay = r0z * r1x - r0x * r1z, az = r0x * r1y - r0y * r1x;  // row0 x row1
            const double bx = r0y * r2z - r0z * r2y, by = r0z * r2x - r0x * r2z, bz = r0x * r2y - r0y * r2x;  // row0 x row2
            const double cx = r1y * r2z - r1z * r2y, cy = r1z * r2x - r1x * r2z, cz = r1x * r2y - r1y * r2x;  // row1 x row2
            double na = ax * ax + ay * ay + az * az;
            const double nb = bx * bx + by * by + bz * bz, nc = cx * cx + cy * cy + cz * cz;
            if (nb > na) { ax = bx; ay = by; az = bz; na = nb; }
            if (nc > na) { ax = cx; ay = cy; az = cz; na = nc; }
            ok = na > 1e-18 * p2 * p2;  // |cross| ~ (largest - l) * (middle - l): the two small eigenvalues are apart
            if (!ok) break;
            const double in = 1.0 / sqrt(na);
            v[0] = ax * in; v[1] = ay * in; v[2] = az * in;
            if (it == 0)
                l = v[0] * (m00 * v[0] + m01 * v[1] + m02 * v[2]) + v[1] * (m01 * v[0] + m11 * v[1] + m12 * v[2]) +
                    v[2] * (m02 * v[0] + m12 * v[1] + m22 * v[2]);
        }
        lam = l;
    }
    if (!ok) lf_smallest_eigvec_jacobi(M, v, lam);
}
__device__ __forceinline__ void lf_line(const uint32_t* ch, int n, float& a, float& b, float& c, float& err)
{
    long long sxx = 0, sxy = 0, syy = 0, sx = 0, sy = 0;
    for (int i = 0; i < n; ++i) {
        const long long x = ch[i] & 0xffffu, y = ch[i] >> 16;
        sxx += x * x; sxy += x * y; syy += y * y; sx += x; sy += y;
    }
    double M[3][3] = {{(double)sxx, (double)sxy, (double)sx}, {(double)sxy, (double)syy, (double)sy}, {(double)sx, (double)sy, (double)n}};
    double v[3], lam;
    lf_smallest_eigvec(M, v, lam);
    a = (float)v[0]; b = (float)v[1]; c = (float)v[2];
    double r2 = 0.0;  // cv::norm(A*u): residuals of the float solution
    for (int i = 0; i < n; ++i) {
        const double r = (double)a * (double)(ch[i] & 0xffffu) + (double)b * (double)(ch[i] >> 16) + (double)c;
        r2 += r * r;
    }
    err = (float)sqrt(r2);
}
__device__ __forceinline__ void lf_depth(const LfPlanes& Q, const LineFitParams& L, const uint32_t* ch, int n, float la, float lb,
                                         float lc, float& u1, float& u2, float& err)
{
    float sx, sy;
    lf_closest(la, lb, lc, (int)(ch[0] & 0xffffu), (int)(ch[0] >> 16), sx, sy);
    const float half = (L.fx + L.fy) / 2;
    double Sdd = 0, Sd = 0, Sn = 0, Sdz = 0, Sz = 0;
    for (int i = 0; i < n; ++i) {
        if (!lf_has_depth(Q, ch[i])) continue;
        float cx, cy;
        lf_closest(la, lb, lc, (int)(ch[i] & 0xffffu), (int)(ch[i] >> 16), cx, cy);
        const double d = lf_norm2(cx, cy, sx, sy);
        const size_t pi = (size_t)(ch[i] >> 16) * Q.W + (ch[i] & 0xffffu);
        const double z = (double)((1.0f / Q.chk[pi]) * half);
        Sdd += d * d; Sd += d; Sn += 1.0; Sdz += d * z; Sz += z;
    }
    const double det = Sdd * Sn - Sd * Sd;
    double x1, x2;
    if (Sn > 0 && det > 1e-9 * (Sdd * Sn + 1e-30)) {
        x1 = (Sdz * Sn - Sd * Sz) / det;
        x2 = (Sdd * Sz - Sd * Sdz) / det;
    } else if (Sn > 0) {  // every row equals (d, 1): minimum-norm solution along that row
        const double d = Sd / Sn, s = Sz / (Sn * (d * d + 1.0));
        x1 = d * s; x2 = s;
    } else {
        x1 = x2 = 0.0;
    }
    u1 = (float)x1; u2 = (float)x2;
    double r2 = 0.0;
    for (int i = 0; i < n; ++i) {
        if (!lf_has_depth(Q, ch[i])) continue;
        float cx, cy;
        lf_closest(la, lb, lc, (int)(ch[i] & 0xffffu), (int)(ch[i] >> 16), cx, cy);
        const float d = lf_norm2(cx, cy, sx, sy);
        const size_t pi = (size_t)(ch[i] >> 16) * Q.W + (ch[i] & 0xffffu);
        const float z = (1.0f / Q.chk[pi]) * half;
        const double r = (double)u1 * (double)d + (double)u2 - (double)z;
        r2 += r * r;
    }
    err = (float)sqrt(r2);
}
__device__ __forceinline__ float lf_point_depth(const LfPlanes& Q, const LineFitParams& L, const uint32_t* ch, float a, float b,
                                                float c, float alpha, float beta, uint32_t rc)
{
    const size_t pi = (size_t)(rc >> 16) * Q.W + (rc & 0xffffu);
    float z = 1.0f / Q.chk[pi];
    if ((double)z < 0.000001) return -1.f;
    if (Q.rs[pi].y > 0.02f) return -1.f;
    float sx, sy, cx, cy;
    lf_closest(a, b, c, (int)(ch[0] & 0xffffu), (int)(ch[0] >> 16), sx, sy);
    lf_closest(a, b, c, (int)(rc & 0xffffu), (int)(rc >> 16), cx, cy);
    const float t = lf_norm2(cx, cy, sx, sy);
    z *= (L.fx + L.fy) / 2;
    return fabsf(alpha * t - z + beta) / sqrtf(alpha * alpha + 1.0f);
}

// chains of a batch of keyframes: pixels packed (r << 16 | c); chain k = pix[off[k] .. off[k+1]) belongs to keyframe kfi[k];
// its lines go to out[slot0[k] ..) (slot0 = prefix sum of len / min_len, the most lines a chain can hold), count in n_out[k].
// One WARP per chain.  What costs time in LineFit is the search for the first window of min_len pixels whose two fits both
// stay below 1.0 (:723-744): on most chains hundreds of start positions are fitted and rejected (two least-squares problems
// each), one after the other in the reference.  Here the 32 lanes fit 32 consecutive start positions at once and the
// ballots below restore the sequential semantics (the first accepted position wins; without one the state of the LAST
// fitted position is what :739-741 test).  The greedy growth and the final fits are short and stay on lane 0.
// One thread per chain took 17.6 ms for the 28 527 chains of 32 VGA keyframes (its longest chain), this form 2.6 ms,
// 0.54 ms with the direct eigenvector (lf_smallest_eigvec).
constexpr int kLineFitBlock = 128;
__global__ void __launch_bounds__(kLineFitBlock) k_line_fit(DevArena A, DevParams P, LineFitParams L, const LineFitKf* __restrict__ kfs,
                                                            int n_chains, const int* __restrict__ off, const int* __restrict__ kfi,
                                                            const uint32_t* __restrict__ pix, const int* __restrict__ slot0,
                                                            DevLine* __restrict__ out, int* __restrict__ n_out)
{
    const int k = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (k >= n_chains) return;  // (a whole warp)
    const int ki = kfi[k];
    const LineFitKf F = kfs[ki];
    L.fx = F.K[0]; L.fy = F.K[1]; L.cx = F.K[2]; L.cy = F.K[3];
#pragma unroll
    for (int i = 0; i < 12; ++i) L.Twc[i] = F.Twc[i];
    LfPlanes Q;
    Q.chk = A.chk + (size_t)F.slot * A.P; Q.rs = A.rs + (size_t)F.slot * A.P; Q.W = P.W; Q.eps_gt = P.eps_gt; Q.sigma_lt = L.sigma_lt;
    const uint32_t* ch = pix + off[k];
    int n = off[k + 1] - off[k];
    DevLine* o = out + slot0[k];
    int no = 0;
    const float inf = __int_as_float(0x7f800000);
    for (;;) {
        float err_l = inf, err_d = inf, a = 0, b = 0, c = 0, alpha = 0, beta = 0;  // warp-uniform
        const int init = L.min_len;
        while (n > init && init < L.max_len) {
            const bool valid = n - lane > init;  // start positions ch + 0 .. that the sequential loop would still visit
            bool ev = false, ok = false;
            float la = 0, lb = 0, lc = 0, lel = inf, lal = 0, lbe = 0, led = inf;
            if (valid) {
                const uint32_t* w = ch + lane;
                if (lf_count(Q, w, 1) >= 1 && lf_count(Q, w, init) >= L.init_depth_count) {
                    ev = true;
                    lf_line(w, init, la, lb, lc, lel);
                    lf_depth(Q, L, w, init, la, lb, lc, lal, lbe, led);
                    ok = (double)lel <= 1.0 && (double)led <= 1.0;
                }
            }
            const unsigned b_ok = __ballot_sync(SDM_FULL, ok), b_ev = __ballot_sync(SDM_FULL, ev);
            const unsigned b_val = __ballot_sync(SDM_FULL, valid);
            // the first accepted position ends the search there; otherwise every valid position was consumed and the
            // state is the one of the last position that was fitted (if any)
            const int src = b_ok ? __ffs(b_ok) - 1 : (b_ev ? 31 - __clz(b_ev) : -1);
            const int adv = b_ok ? src : __popc(b_val);
            if (src >= 0) {
                a = __shfl_sync(SDM_FULL, la, src); b = __shfl_sync(SDM_FULL, lb, src); c = __shfl_sync(SDM_FULL, lc, src);
                alpha = __shfl_sync(SDM_FULL, lal, src); beta = __shfl_sync(SDM_FULL, lbe, src);
                err_l = __shfl_sync(SDM_FULL, lel, src); err_d = __shfl_sync(SDM_FULL, led, src);
            }
            ch += adv;
            n -= adv;
            if (b_ok) break;
        }
        if (err_l > L.e1 || err_d > L.e2) break;  // (inf when no window was ever fitted; NaN compares false like :739-741)
        int len = init;
        if (lane == 0) {
        int interval = 0;
        while (len < L.max_len && len < n) {
            const uint32_t rc = ch[len];
            const float dist = fabsf(a * (float)(rc & 0xffffu) + b * (float)(rc >> 16) + c) / sqrtf(a * a + b * b);
            if (dist > L.e1) break;
            ++len; ++interval;
            if (interval >= L.min_len) {
                interval = 0;
                if (lf_count(Q, ch + len - L.min_len, L.min_len) < 1) { len -= L.min_len; break; }
                int prev = len - L.min_len;
                bool stop = false;
                for (int i = prev, i_end = len; i < i_end; ++i) {
                    const float dept = lf_point_depth(Q, L, ch, a, b, c, alpha, beta, ch[i]);
                    if (dept > L.e2) { len = prev; stop = true; break; }
                    if ((double)dept >= 0.0) prev = i;
                }
                if (stop) break;
            }
        }
        lf_line(ch, len, a, b, c, err_l);
        if ((float)lf_count(Q, ch, len) / (float)len > (float)L.init_depth_count / (float)L.min_len) {
            lf_depth(Q, L, ch, len, a, b, c, alpha, beta, err_d);
            float sx, sy, ex, ey;
            lf_closest(a, b, c, (int)(ch[0] & 0xffffu), (int)(ch[0] >> 16), sx, sy);
            lf_closest(a, b, c, (int)(ch[len - 1] & 0xffffu), (int)(ch[len - 1] >> 16), ex, ey);
            const float half = (L.fx + L.fy) / 2;
            float Zs = beta;
            Zs /= half;
            const float Xs = Zs * (sx - L.cx) / L.fx, Ys = Zs * (sy - L.cy) / L.fy;
            float Ze = alpha * lf_norm2(ex, ey, sx, sy) + beta;
            Ze /= half;
            const float Xe = Ze * (ex - L.cx) / L.fx, Ye = Ze * (ey - L.cy) / L.fy;
            const double dx = (double)(Xe - Xs), dy = (double)(Ye - Ys), dz = (double)(Ze - Zs);
            const double nd = sqrt(dx * dx + dy * dy + dz * dz);
            const double ns = sqrt((double)Xs * Xs + (double)Ys * Ys + (double)Zs * Zs);
            const double ne = sqrt((double)Xe * Xe + (double)Ye * Ye + (double)Ze * Ze);
            const double cosS = (dx * Xs + dy * Ys + dz * Zs) / nd / ns, cosE = (dx * Xe + dy * Ye + dz * Ze) / nd / ne;
            const double angS = acos(fabs(cosS)) * 180.0 / 3.1415926535897932384626433832795;
            const double angE = acos(fabs(cosE)) * 180.0 / 3.1415926535897932384626433832795;
            if (angS > (double)L.min_angle && angE > (double)L.min_angle) {
                DevLine d;
                d.seg[0] = sx; d.seg[1] = sy; d.seg[2] = ex; d.seg[3] = ey;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    d.xyz[r] = L.Twc[4 * r] * Xs + L.Twc[4 * r + 1] * Ys + L.Twc[4 * r + 2] * Zs + L.Twc[4 * r + 3] * 1.0f;
                    d.xyz[3 + r] = L.Twc[4 * r] * Xe + L.Twc[4 * r + 1] * Ye + L.Twc[4 * r + 2] * Ze + L.Twc[4 * r + 3] * 1.0f;
                }
                d.chain = k - F.chain0;
                d.kf_index = ki;
                o[no++] = d;
            }
        }
        }
        len = __shfl_sync(SDM_FULL, len, 0);
        ch += len;
        n -= len;
    }
    if (lane == 0) n_out[k] = no;
}

// exclusive scan of the per-chain line counts (one block) + per-keyframe totals; then the compaction: chain k copies its
// lines from the worst-case slots to their final place (chain order = the reference's push_back order, :822-823)
__global__ void __launch_bounds__(1024) k_line_scan(const int* __restrict__ n_out, int n_chains, const LineFitKf* __restrict__ kfs,
                                                    int n_kf, unsigned long long* __restrict__ offs,
                                                    unsigned long long* __restrict__ kf_totals)
{
    __shared__ unsigned long long part[1024];
    const int t = threadIdx.x, per = (n_chains + 1023) / 1024;
    const int i0 = min(t * per, n_chains), i1 = min(i0 + per, n_chains);
    unsigned long long s = 0;
    for (int i = i0; i < i1; ++i) s += (unsigned long long)n_out[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned long long v = t >= d ? part[t - d] : 0ull;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = part[t] - s;
    for (int i = i0; i < i1; ++i) { offs[i] = run; run += (unsigned long long)n_out[i]; }
    if (t == 1023) offs[n_chains] = part[1023];
    __syncthreads();
    for (int i = t; i < n_kf; i += 1024) {
        const int c0 = kfs[i].chain0, c1 = i + 1 < n_kf ? kfs[i + 1].chain0 : n_chains;
        kf_totals[i] = offs[c1] - offs[c0];
    }
}
__global__ void __launch_bounds__(256) k_line_compact(const DevLine* __restrict__ in, const int* __restrict__ slot0,
                                                     const int* __restrict__ n_out, const unsigned long long* __restrict__ offs,
                                                     int n_chains, DevLine* __restrict__ out, unsigned long long capacity)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_chains) return;
    const DevLine* src = in + slot0[k];
    const unsigned long long o = offs[k];
    for (int i = 0; i < n_out[k]; ++i)
        if (o + i < capacity) out[o + i] = src[i];
}

}  // namespace sdm
