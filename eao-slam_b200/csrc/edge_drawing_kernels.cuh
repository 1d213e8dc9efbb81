// edge_drawing_kernels.cuh - stage 1 of the Edge Drawing detector on the device (sm_100a), for a batch of keyframes.
//
// The reference calls the closed-source EDLib once per keyframe inside pass 1 (LineDetector::DetectEdgeMap,
// /root/reference/src/LineDetector.cc:843-881: DetectEdgesByED(srcImg, width, height, SOBEL_OPERATOR, 36, 8, 1.0)); its chains
// are the mEdgeIndex candidate mask of the hot loop (ProbabilityMapping.cc:454) and the input of the 3-D line fitting
// (:884-900).  eao-slam_b200/host/edge_drawing.h documents the open implementation and how every open detail was pinned to the
// library's output.  The detector has two stages: per-pixel image work (smoothing, Sobel gradient, edge direction, anchor test)
// and a sequential walk from the anchors.  These kernels are the first stage, the walk runs on host threads (sdm_edge_drawing in
// sdm_b200.cu) while the next chunk of keyframes is on the device.
//
//   in : im  [n][H][W] u8                                                  1 byte per pixel
//   out: G   [n][H][W] i16  |gx| + |gy| of Sobel on the smoothed image; border pixels = grad_thresh - 1
//        F   [n][H][W] u8   bits 0-1: 0 below grad_thresh, 1 vertical edge pixel (|gx| >= |gy|), 2 horizontal;
//                           bits 2-3 / 4-5: where an unguided walk goes from the pixel backwards / forwards along its direction
//                           (the largest of the three gradients ahead: the smart-routing step, decided here for all pixels at
//                           once so that the sequential walk never touches the gradient plane);
//                           bit 7: anchor (rows / columns 2 .. size-3: G exceeds both neighbours across the edge by
//                           anchor_thresh)                                 3 bytes per pixel
// Integer arithmetic throughout: the planes are bit-identical to sdm_host::EdPlanesHost.
//
// One block = one 64 x 16 tile of one keyframe, everything between the image bytes and the two outputs stays in shared
// memory: image tile with a 4-pixel apron (2 smoothing + 1 Sobel + 1 anchor test) -> separable 1-4-6-4-1 smoothing (row sums
// as u16, column pass with the library's rounding: half to even where its 4-wide vector loop runs, half up in the last W % 4
// columns) -> gradient + direction (packed in one u16) -> anchor test -> 8-byte (G) and 4-byte (F) stores per thread.
// HBM-bound by construction (4 algorithmic bytes per pixel, no re-reads besides the apron, which L2 serves).
//
// Two kernels.  k_ed_planes4 (widths that are multiples of four: every camera format) works on FOUR pixels per thread in
// every phase with two 16-bit lanes per 32-bit register: 32-bit image loads, the 1-4-6-4-1 sums as packed multiply-adds (a
// row sum is at most 16 * 255 and a column sum at most 256 * 255, both fit 16 bits, so no carry crosses the lanes), the
// rounding, the Sobel taps and |a - b| = max - min (VIMNMX.U16x2) all lane-wise - 3.2x fewer issued instructions than the
// byte-per-thread form, which was issue-bound (profiles/r02l_ed_planes.md: 80 % issue, 242 thread instructions per pixel).
// The phases are host-callable functions of (tile, thread index), so tests/cpp/test_edge_drawing.cpp replays the kernel
// thread by thread on the CPU against EdPlanesHost.  k_ed_planes (any width) is the byte-per-thread form.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "../host/edge_drawing.h"

#if defined(__CUDACC__)
#define SDM_ED_HD __host__ __device__ __forceinline__
#define SDM_ED_UNROLL _Pragma("unroll")
#else
#define SDM_ED_HD inline
#define SDM_ED_UNROLL
#endif

namespace sdm {

#if !defined(__CUDACC__)
using std::max;
using std::min;
#endif

constexpr int kEdTW = 64, kEdTH = 16, kEdThreads = 256;
constexpr int kEdFlagAnchor = 0x80;

// ---- four pixels per thread (W % 4 == 0) ----------------------------------------------------------------------------------
// Two 16-bit lanes per register.  A group of four neighbouring values v0 v1 v2 v3 is kept as the pair of registers
// (v0 | v2 << 16, v1 | v3 << 16): byte b of a 32-bit image word lands in lane (b & 1) of register (b >> 1) with one AND / shift.
constexpr int kEdImRows = kEdTH + 8;   // image rows y0 - 4 .. y0 + TH + 3
constexpr int kEdImWords = 18;         // image columns x0 - 4 .. x0 + 67 as 32-bit words
constexpr int kEdGroups = 17;          // groups of four columns: from x0 - 2 (row sums, smoothed image), from x0 - 1 (gradient)
constexpr int kEdGWords = 2 * kEdGroups;  // gradient row: 68 u16 = 34 words, columns x0 - 1 .. x0 + 66

struct EdPair { uint32_t x, y; };
struct alignas(16) EdTile {
    uint32_t im[kEdImRows][kEdImWords];
    EdPair row[kEdImRows][kEdGroups];          // row sums 1 4 6 4 1, lanes (o0, o2) / (o1, o3)
    uint32_t sm[kEdTH + 4][kEdImWords];        // smoothed bytes in natural order; word 17 is padding the last gradient group reads
    uint32_t g[kEdTH + 2][kEdGWords];          // gradient | direction << 12, u16 in natural order
};

SDM_ED_HD uint32_t ed_funnel_r(uint32_t lo, uint32_t hi, int s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> s);
#endif
}
SDM_ED_HD uint32_t ed_absdiff2(uint32_t a, uint32_t b)  // |a - b| in both 16-bit lanes
{
#if defined(__CUDA_ARCH__)
    return __vmaxu2(a, b) - __vminu2(a, b);
#else
    const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
    return (al > bl ? al - bl : bl - al) | ((ah > bh ? ah - bh : bh - ah) << 16);
#endif
}
SDM_ED_HD uint32_t ed_lo_pair(uint32_t a, uint32_t b) { return (a & 0xffffu) | (b << 16); }        // (a.lo, b.lo)
SDM_ED_HD uint32_t ed_hi_pair(uint32_t a, uint32_t b) { return (a >> 16) | (b & 0xffff0000u); }    // (a.hi, b.hi)

struct EdArgs {
    const uint8_t* src;  // the keyframe's image, dense W x H
    int16_t* G;
    uint8_t* F;
    int W, H, x0, y0, grad_thresh, anchor_thresh;
};

// image tile, replicated border (the smoothing's border rule); items: kEdImRows * kEdImWords
SDM_ED_HD void ed4_load(EdTile& t, const EdArgs& a, int i)
{
    const int j = i / kEdImWords, w = i - j * kEdImWords;
    const int y = min(max(a.y0 - 4 + j, 0), a.H - 1), x = a.x0 - 4 + 4 * w;
    const uint8_t* r = a.src + (size_t)y * a.W;
    uint32_t v;
    if (x < 0) v = r[0] * 0x01010101u;
    else if (x >= a.W) v = r[a.W - 1] * 0x01010101u;
    else {
#if defined(__CUDA_ARCH__)
        v = __ldg(reinterpret_cast<const uint32_t*>(r + x));
#else
        v = (uint32_t)r[x] | ((uint32_t)r[x + 1] << 8) | ((uint32_t)r[x + 2] << 16) | ((uint32_t)r[x + 3] << 24);
#endif
    }
    t.im[j][w] = v;
}

// row sums at x = x0 - 2 + 4 g + q from image bytes b0 .. b7 = columns x0 - 4 + 4 g ..; items: kEdImRows * kEdGroups
SDM_ED_HD void ed4_rows(EdTile& t, int i)
{
    const int j = i / kEdGroups, g = i - j * kEdGroups;
    const uint32_t m = 0x00ff00ffu;
    const uint32_t w0 = t.im[j][g], w1 = t.im[j][g + 1];
    const uint32_t E = w0 & m, O = (w0 >> 8) & m;                                  // (b0, b2) (b1, b3)
    const uint32_t S16 = ed_funnel_r(w0, w1, 16) & m, S24 = ed_funnel_r(w0, w1, 24) & m;  // (b2, b4) (b3, b5)
    const uint32_t E1 = w1 & m, O1 = (w1 >> 8) & m;                                // (b4, b6) (b5, b7)
    EdPair r;
    r.x = E + 4 * O + 6 * S16 + 4 * S24 + E1;   // o0 = b0 + 4 b1 + 6 b2 + 4 b3 + b4,  o2 = b2 + .. + b6
    r.y = O + 4 * S16 + 6 * S24 + 4 * E1 + O1;  // o1, o3
    t.row[j][g] = r;
}

// column pass at y = y0 - 2 + j with the library's half-to-even rounding of v / 256; items: (kEdTH + 4) * kEdGroups
SDM_ED_HD uint32_t ed4_round(uint32_t v) { return ((v + 0x007f007fu + ((v >> 8) & 0x00010001u)) >> 8) & 0x00ff00ffu; }
SDM_ED_HD void ed4_cols(EdTile& t, int i)
{
    const int j = i / kEdGroups, g = i - j * kEdGroups;
    const EdPair a = t.row[j][g], b = t.row[j + 1][g], c = t.row[j + 2][g], d = t.row[j + 3][g], e = t.row[j + 4][g];
    const uint32_t s02 = ed4_round(a.x + 4 * b.x + 6 * c.x + 4 * d.x + e.x);
    const uint32_t s13 = ed4_round(a.y + 4 * b.y + 6 * c.y + 4 * d.y + e.y);
    t.sm[j][g] = s02 | (s13 << 8);
    if (g == kEdGroups - 1) t.sm[j][kEdGroups] = 0;
}

// gradient + direction at y = y0 - 1 + j, x = x0 - 1 + 4 g + q; items: (kEdTH + 2) * kEdGroups
SDM_ED_HD void ed4_grad(EdTile& t, const EdArgs& a, int i)
{
    const int j = i / kEdGroups, g = i - j * kEdGroups;
    const uint32_t m = 0x00ff00ffu;
    uint32_t E[3], O[3], S16[3], S24[3];
    SDM_ED_UNROLL
    for (int r = 0; r < 3; ++r) {
        const uint32_t w0 = t.sm[j + r][g], w1 = t.sm[j + r][g + 1];
        E[r] = w0 & m;
        O[r] = (w0 >> 8) & m;
        S16[r] = ed_funnel_r(w0, w1, 16) & m;
        S24[r] = ed_funnel_r(w0, w1, 24) & m;
    }
    // |gy|: the 1 2 1 row taps of the row below minus those of the row above, centres at bytes 1 .. 4
    const uint32_t gy02 = ed_absdiff2(E[2] + 2 * O[2] + S16[2], E[0] + 2 * O[0] + S16[0]);
    const uint32_t gy13 = ed_absdiff2(O[2] + 2 * S16[2] + S24[2], O[0] + 2 * S16[0] + S24[0]);
    // |gx|: the 1 2 1 column taps of the column to the right minus those of the column to the left
    const uint32_t cE = E[0] + 2 * E[1] + E[2], cO = O[0] + 2 * O[1] + O[2];
    const uint32_t cS16 = S16[0] + 2 * S16[1] + S16[2], cS24 = S24[0] + 2 * S24[1] + S24[2];
    const uint32_t gx02 = ed_absdiff2(cS16, cE), gx13 = ed_absdiff2(cS24, cO);
    const uint32_t th = (uint32_t)a.grad_thresh * 0x00010001u;
    uint32_t code[2];
    {
        const uint32_t gx[2] = {gx02, gx13}, gy[2] = {gy02, gy13};
    SDM_ED_UNROLL
        for (int k = 0; k < 2; ++k) {
            const uint32_t v = gx[k] + gy[k];                                                   // <= 2040 per lane
            const uint32_t ge = (((gx[k] | 0x80008000u) - gy[k]) >> 15) & 0x00010001u;          // |gx| >= |gy|
            const uint32_t on = (((v | 0x80008000u) - th) >> 15) & 0x00010001u;                 // v >= grad_thresh
            code[k] = v | (((0x00020002u - ge) & (on * 3u)) << 12);                             // direction 1 vertical, 2 horizontal
        }
    }
    uint32_t lo = ed_lo_pair(code[0], code[1]), hi = ed_hi_pair(code[0], code[1]);  // (o0, o1) (o2, o3)
    const int y = a.y0 - 1 + j, xs = a.x0 - 1 + 4 * g;
    const uint32_t border = (uint32_t)(a.grad_thresh - 1);
    if (y < 1 || y > a.H - 2) lo = hi = border * 0x00010001u;
    else if (xs < 1 || xs + 3 > a.W - 2) {
        if (xs < 1 || xs > a.W - 2) lo = (lo & 0xffff0000u) | border;
        if (xs + 1 < 1 || xs + 1 > a.W - 2) lo = (lo & 0x0000ffffu) | (border << 16);
        if (xs + 2 < 1 || xs + 2 > a.W - 2) hi = (hi & 0xffff0000u) | border;
        if (xs + 3 < 1 || xs + 3 > a.W - 2) hi = (hi & 0x0000ffffu) | (border << 16);
    }
    t.g[j][2 * g] = lo;
    t.g[j][2 * g + 1] = hi;
}

// anchor test and stores: thread i -> row i / 16 of the tile, four pixels from column 4 * (i % 16)
SDM_ED_HD void ed4_store(const EdTile& t, const EdArgs& a, int i)
{
    const int ty = i >> 4, tw = (i & 15) << 1;  // first gradient word of the group (u16 index 4 * (i % 16))
    const int y = a.y0 + ty, x = a.x0 + 2 * tw;
    if (y >= a.H || x >= a.W) return;
    // u16 lanes c[0 .. 5] of the pixel's row = columns x - 1 .. x + 4, u[..] / d[..] the rows above / below
    const uint32_t c0 = t.g[ty + 1][tw], c1 = t.g[ty + 1][tw + 1], c2 = t.g[ty + 1][tw + 2];
    const uint32_t u0 = t.g[ty][tw], u1 = t.g[ty][tw + 1], u2 = t.g[ty][tw + 2];
    const uint32_t d0 = t.g[ty + 2][tw], d1 = t.g[ty + 2][tw + 1], d2 = t.g[ty + 2][tw + 2];
    const uint32_t cw[3] = {c0, c1, c2}, uw[3] = {u0, u1, u2}, dw[3] = {d0, d1, d2};
    uint32_t f = 0;
    SDM_ED_UNROLL
    for (int q = 0; q < 4; ++q) {
#define SDM_ED_LANE(w, k) (((w)[(k) >> 1] >> (((k) & 1) * 16)) & 0xffffu)
        const int v = (int)SDM_ED_LANE(cw, q + 1);
        const int gq = v & 0xfff, dir = v >> 12;
        int fq = dir;
        const int xx = x + q;
        if (dir != 0) {
            // the eight neighbours' gradients (an edge pixel is never on the image border, so all of them exist)
            const int ul = (int)SDM_ED_LANE(uw, q) & 0xfff, uc = (int)SDM_ED_LANE(uw, q + 1) & 0xfff, ur = (int)SDM_ED_LANE(uw, q + 2) & 0xfff;
            const int cl = (int)SDM_ED_LANE(cw, q) & 0xfff, cr = (int)SDM_ED_LANE(cw, q + 2) & 0xfff;
            const int dl = (int)SDM_ED_LANE(dw, q) & 0xfff, dc = (int)SDM_ED_LANE(dw, q + 1) & 0xfff, dr = (int)SDM_ED_LANE(dw, q + 2) & 0xfff;
            // where an unguided walk goes from here, backwards and forwards along the pixel's direction (sdm_host::ed_route_code)
            const int back = dir == 2 ? sdm_host::ed_route_code(ul, cl, dl) : sdm_host::ed_route_code(ul, uc, ur);
            const int fwd = dir == 2 ? sdm_host::ed_route_code(ur, cr, dr) : sdm_host::ed_route_code(dl, dc, dr);
            fq |= (back << sdm_host::kEdRouteShiftBack) | (fwd << sdm_host::kEdRouteShiftFwd);
            if (y >= 2 && y <= a.H - 3 && xx >= 2 && xx <= a.W - 3) {
                const int n0 = dir == 1 ? cl : uc, n1 = dir == 1 ? cr : dc;
                if (gq - n0 >= a.anchor_thresh && gq - n1 >= a.anchor_thresh) fq |= kEdFlagAnchor;
            }
        }
#undef SDM_ED_LANE
        f |= (uint32_t)fq << (8 * q);
    }
    const uint32_t g01 = ed_funnel_r(c0, c1, 16) & 0x0fff0fffu, g23 = ed_funnel_r(c1, c2, 16) & 0x0fff0fffu;
    const size_t o = (size_t)y * a.W + x;
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2*>(a.G + o) = make_uint2(g01, g23);
    *reinterpret_cast<uint32_t*>(a.F + o) = f;
#else
    a.G[o] = (int16_t)(g01 & 0xffffu); a.G[o + 1] = (int16_t)(g01 >> 16);
    a.G[o + 2] = (int16_t)(g23 & 0xffffu); a.G[o + 3] = (int16_t)(g23 >> 16);
    for (int q = 0; q < 4; ++q) a.F[o + q] = (uint8_t)(f >> (8 * q));
#endif
}

constexpr int kEd4Items0 = kEdImRows * kEdImWords, kEd4Items1 = kEdImRows * kEdGroups, kEd4Items2 = (kEdTH + 4) * kEdGroups,
              kEd4Items3 = (kEdTH + 2) * kEdGroups;

#if defined(__CUDACC__)
__global__ void __launch_bounds__(kEdThreads)
k_ed_planes4(const uint8_t* __restrict__ im, int W, int H, int grad_thresh, int anchor_thresh, int16_t* __restrict__ G,
             uint8_t* __restrict__ F)
{
    __shared__ EdTile t;
    const int tid = threadIdx.x;
    const size_t plane = (size_t)blockIdx.z * (size_t)W * (size_t)H;
    const EdArgs a = {im + plane, G + plane, F + plane, W, H, (int)blockIdx.x * kEdTW, (int)blockIdx.y * kEdTH, grad_thresh, anchor_thresh};
    for (int i = tid; i < kEd4Items0; i += kEdThreads) ed4_load(t, a, i);
    __syncthreads();
    for (int i = tid; i < kEd4Items1; i += kEdThreads) ed4_rows(t, i);
    __syncthreads();
    for (int i = tid; i < kEd4Items2; i += kEdThreads) ed4_cols(t, i);
    __syncthreads();
    for (int i = tid; i < kEd4Items3; i += kEdThreads) ed4_grad(t, a, i);
    __syncthreads();
    ed4_store(t, a, tid);
}
#endif

// the same kernel thread by thread on the host (tests): every phase for every thread of every tile, in order
inline void EdPlanes4Replay(const uint8_t* im, int W, int H, int grad_thresh, int anchor_thresh, int16_t* G, uint8_t* F)
{
    static EdTile t;
    for (int by = 0; by < (H + kEdTH - 1) / kEdTH; ++by)
        for (int bx = 0; bx < (W + kEdTW - 1) / kEdTW; ++bx) {
            const EdArgs a = {im, G, F, W, H, bx * kEdTW, by * kEdTH, grad_thresh, anchor_thresh};
            for (int i = 0; i < kEd4Items0; ++i) ed4_load(t, a, i);
            for (int i = 0; i < kEd4Items1; ++i) ed4_rows(t, i);
            for (int i = 0; i < kEd4Items2; ++i) ed4_cols(t, i);
            for (int i = 0; i < kEd4Items3; ++i) ed4_grad(t, a, i);
            for (int i = 0; i < kEdThreads; ++i) ed4_store(t, a, i);
        }
}

// ---- one byte per thread (any width) ----------------------------------------------------------------------------------------
#if defined(__CUDACC__)
__global__ void __launch_bounds__(kEdThreads)
k_ed_planes(const uint8_t* __restrict__ im, int W, int H, int grad_thresh, int anchor_thresh, int16_t* __restrict__ G,
            uint8_t* __restrict__ F)
{
    __shared__ __align__(16) uint8_t s_im[kEdTH + 8][kEdTW + 8];
    __shared__ __align__(16) uint16_t s_row[kEdTH + 8][kEdTW + 4];
    __shared__ __align__(16) uint8_t s_sm[kEdTH + 4][kEdTW + 4];
    __shared__ __align__(16) uint16_t s_g[kEdTH + 2][kEdTW + 2];  // gradient | direction << 12

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kEdTW, y0 = blockIdx.y * kEdTH;
    const size_t plane = (size_t)blockIdx.z * (size_t)W * (size_t)H;
    const uint8_t* src = im + plane;

    // image tile, replicated border (the smoothing's border rule)
    for (int i = tid; i < (kEdTH + 8) * (kEdTW + 8); i += kEdThreads) {
        const int j = i / (kEdTW + 8), k = i - j * (kEdTW + 8);
        const int y = min(max(y0 - 4 + j, 0), H - 1), x = min(max(x0 - 4 + k, 0), W - 1);
        s_im[j][k] = __ldg(src + (size_t)y * W + x);
    }
    __syncthreads();
    // row sums 1 4 6 4 1 at x = x0 - 2 + k
    for (int i = tid; i < (kEdTH + 8) * (kEdTW + 4); i += kEdThreads) {
        const int j = i / (kEdTW + 4), k = i - j * (kEdTW + 4);
        const uint8_t* p = &s_im[j][k];
        s_row[j][k] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();
    // column pass at y = y0 - 2 + j, rounding of v / 256 as the library does it
    const int wsimd = W & ~3;
    for (int i = tid; i < (kEdTH + 4) * (kEdTW + 4); i += kEdThreads) {
        const int j = i / (kEdTW + 4), k = i - j * (kEdTW + 4);
        const int v = s_row[j][k] + 4 * s_row[j + 1][k] + 6 * s_row[j + 2][k] + 4 * s_row[j + 3][k] + s_row[j + 4][k];
        const int q = v >> 8, rem = v & 255;
        const int even = q + ((rem > 128) | ((rem == 128) & (q & 1)));
        const int up = (v + 128) >> 8;
        s_sm[j][k] = (uint8_t)((x0 - 2 + k) < wsimd ? even : up);
    }
    __syncthreads();
    // gradient + direction at y = y0 - 1 + j, x = x0 - 1 + k
    for (int i = tid; i < (kEdTH + 2) * (kEdTW + 2); i += kEdThreads) {
        const int j = i / (kEdTW + 2), k = i - j * (kEdTW + 2);
        const int y = y0 - 1 + j, x = x0 - 1 + k;
        int v = grad_thresh - 1;
        if (y >= 1 && y <= H - 2 && x >= 1 && x <= W - 2) {
            const int a = s_sm[j][k], b = s_sm[j][k + 1], c = s_sm[j][k + 2];
            const int d = s_sm[j + 1][k], f = s_sm[j + 1][k + 2];
            const int g = s_sm[j + 2][k], h = s_sm[j + 2][k + 1], l = s_sm[j + 2][k + 2];
            const int com1 = l - a, com2 = c - g;
            const int gx = abs(com1 + com2 + 2 * (f - d));
            const int gy = abs(com1 - com2 + 2 * (h - b));
            v = gx + gy;
            if (v >= grad_thresh) v |= (gx >= gy ? 1 : 2) << 12;
        }
        s_g[j][k] = (uint16_t)v;
    }
    __syncthreads();
    // anchor test and stores: thread t -> row t / 16 of the tile, four pixels from column 4 * (t % 16)
    {
        const int ty = tid >> 4, tx = (tid & 15) << 2;
        const int y = y0 + ty, x = x0 + tx;
        if (y < H && x < W) {
            short gq[4];
            uint8_t fq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int v = s_g[ty + 1][tx + q + 1];
                const int g = v & 0xfff, dir = v >> 12;
                int f = dir;
                const int xx = x + q;
                if (dir != 0) {
                    const int ul = s_g[ty][tx + q] & 0xfff, uc = s_g[ty][tx + q + 1] & 0xfff, ur = s_g[ty][tx + q + 2] & 0xfff;
                    const int cl = s_g[ty + 1][tx + q] & 0xfff, cr = s_g[ty + 1][tx + q + 2] & 0xfff;
                    const int dl = s_g[ty + 2][tx + q] & 0xfff, dc = s_g[ty + 2][tx + q + 1] & 0xfff, dr = s_g[ty + 2][tx + q + 2] & 0xfff;
                    const int back = dir == 2 ? sdm_host::ed_route_code(ul, cl, dl) : sdm_host::ed_route_code(ul, uc, ur);
                    const int fwd = dir == 2 ? sdm_host::ed_route_code(ur, cr, dr) : sdm_host::ed_route_code(dl, dc, dr);
                    f |= (back << sdm_host::kEdRouteShiftBack) | (fwd << sdm_host::kEdRouteShiftFwd);
                    if (y >= 2 && y <= H - 3 && xx >= 2 && xx <= W - 3) {
                        const int n0 = dir == 1 ? cl : uc, n1 = dir == 1 ? cr : dc;
                        if (g - n0 >= anchor_thresh && g - n1 >= anchor_thresh) f |= kEdFlagAnchor;
                    }
                }
                gq[q] = (short)g;
                fq[q] = (uint8_t)f;
            }
            int16_t* go = G + plane + (size_t)y * W + x;
            uint8_t* fo = F + plane + (size_t)y * W + x;
            if (x + 3 < W && (W & 3) == 0) {
                *reinterpret_cast<short4*>(go) = make_short4(gq[0], gq[1], gq[2], gq[3]);
                *reinterpret_cast<uchar4*>(fo) = make_uchar4(fq[0], fq[1], fq[2], fq[3]);
            } else {
                for (int q = 0; q < 4 && x + q < W; ++q) { go[q] = gq[q]; fo[q] = fq[q]; }
            }
        }
    }
}

// ---- the anchors of every image in walking order: decreasing gradient, raster order among equals ---------------------------
// A stable counting sort over the 2048 gradient values, one block of eight warps per image (the walk's own thread would pay two
// dependent memory round trips per anchor, twice: a quarter of the routing kernel in its first form; one warp: 3 ms per image).
// The image is cut into eight consecutive pieces, warp w counts piece w into its own histogram, the block turns the eight
// histograms into start positions (all smaller keys of all pieces first, then the same key of the earlier pieces), and every
// warp places the anchors of its piece.  Inside a piece, lane l holds 16 consecutive pixels of a 512-pixel window (one 16-byte
// load of the flags, two of the gradients, in flight together) and the lanes take turns in lane order, so raster order is kept.
constexpr int kEdSortWarps = 8, kEdSortThreads = 32 * kEdSortWarps, kEdSortSmem = kEdSortWarps * 2048 * (int)sizeof(int);

__device__ __forceinline__ int ed_anchor_key(uint32_t gword, int half) { return 2047 - min((int)(short)((gword >> (16 * half)) & 0xffffu), 2047); }

// pass 1 (count) or pass 2 (place) of one warp over pixels [begin, end) of one image; hist = this warp's 2048 counters
template <bool kPlace>
__device__ __forceinline__ void ed_warp_anchor_pass(const uint8_t* __restrict__ F, const int16_t* __restrict__ G, size_t begin, size_t end, bool wide,
                                                    int* hist, int* __restrict__ anchors, int lane)
{
    const size_t win = wide ? 512 : 32;
    for (size_t i0 = begin; i0 < end; i0 += win) {
        if (wide) {
            const size_t i = i0 + 16 * (size_t)lane;
            uint32_t fw[4] = {0, 0, 0, 0}, gw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            bool any = false;
            if (i < end) {
                const uint4 f = *reinterpret_cast<const uint4*>(F + i);
                const uint4 g0 = *reinterpret_cast<const uint4*>(G + i), g1 = *reinterpret_cast<const uint4*>(G + i + 8);
                fw[0] = f.x; fw[1] = f.y; fw[2] = f.z; fw[3] = f.w;
                any = ((f.x | f.y | f.z | f.w) & 0x80808080u) != 0;
                gw[0] = g0.x; gw[1] = g0.y; gw[2] = g0.z; gw[3] = g0.w; gw[4] = g1.x; gw[5] = g1.y; gw[6] = g1.z; gw[7] = g1.w;
            }
            if (!kPlace) {
                if (any) {
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if ((fw[k >> 2] >> (8 * (k & 3))) & 0x80u) atomicAdd(&hist[ed_anchor_key(gw[k >> 1], k & 1)], 1);
                }
            } else {
                // the lanes that hold anchors take their turn in lane order; a lane places its own anchors in byte order
                unsigned m = __ballot_sync(0xffffffffu, any);
                while (m) {
                    const int turn = __ffs(m) - 1;
                    if (lane == turn) {
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            if ((fw[k >> 2] >> (8 * (k & 3))) & 0x80u) anchors[hist[ed_anchor_key(gw[k >> 1], k & 1)]++] = (int)(i + k);
                    }
                    __syncwarp();
                    m &= m - 1;
                }
            }
        } else {
            const size_t i = i0 + lane;
            const bool is_anchor = i < end && (F[i] & kEdFlagAnchor);
            if (!kPlace) {
                if (is_anchor) atomicAdd(&hist[2047 - min((int)G[i], 2047)], 1);
            } else {
                const unsigned m = __ballot_sync(0xffffffffu, is_anchor);
                if (is_anchor) {
                    const int key = 2047 - min((int)G[i], 2047);
                    const unsigned peers = __match_any_sync(m, key);   // the lanes of this group with the same gradient
                    const int leader = __ffs(peers) - 1;
                    int pos = 0;
                    if (lane == leader) { pos = hist[key]; hist[key] = pos + __popc(peers); }
                    pos = __shfl_sync(peers, pos, leader);
                    anchors[pos + __popc(peers & ((1u << lane) - 1u))] = (int)i;
                }
                __syncwarp();
            }
        }
    }
}

// list + img * stride: the image's anchors (room for `cap`); count[img] = their number, or -1 (nothing written) if more than cap
__global__ void __launch_bounds__(kEdSortThreads)
k_ed_sort(const int16_t* __restrict__ Gall, const uint8_t* __restrict__ Fall, int W, int H, int* __restrict__ list, size_t stride, int cap,
          int* __restrict__ count)
{
    extern __shared__ int hist[];  // [kEdSortWarps][2048]
    __shared__ int warp_tot[kEdSortWarps];
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const size_t P = (size_t)W * H;
    const uint8_t* F = Fall + (size_t)img * P;
    const int16_t* G = Gall + (size_t)img * P;
    const bool wide = (P & 15) == 0;  // (plane bases are 256-byte aligned and P is a multiple of 16: aligned 16-byte loads)
    const size_t unit = wide ? 512 : 32;
    const size_t piece = ((P + kEdSortWarps - 1) / kEdSortWarps + unit - 1) / unit * unit;
    const size_t begin = min((size_t)w * piece, P), end = min(begin + piece, P);
    for (int k = tid; k < kEdSortWarps * 2048; k += kEdSortThreads) hist[k] = 0;
    __syncthreads();
    ed_warp_anchor_pass<false>(F, G, begin, end, wide, hist + w * 2048, nullptr, lane);
    __syncthreads();
    // start positions: thread t owns keys 8 t .. 8 t + 7
    int tot[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int c = 0;
        for (int q = 0; q < kEdSortWarps; ++q) c += hist[q * 2048 + 8 * tid + k];
        tot[k] = c;
        sum += c;
    }
    int incl = sum;
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int base = incl - sum, total = 0;
    for (int q = 0; q < kEdSortWarps; ++q) { if (q < w) base += warp_tot[q]; total += warp_tot[q]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int run = base;
        for (int q = 0; q < kEdSortWarps; ++q) { const int c = hist[q * 2048 + 8 * tid + k]; hist[q * 2048 + 8 * tid + k] = run; run += c; }
        base += tot[k];
    }
    __syncthreads();
    if (tid == 0) count[img] = total <= cap ? total : -1;
    if (total > cap) return;
    ed_warp_anchor_pass<true>(F, G, begin, end, wide, hist + w * 2048, list + (size_t)img * stride, lane);
}

// ---- stage 2 on the device: one warp per image ------------------------------------------------------------------------------
// The routing walk is sequential per image (a step depends on the marks the previous walks left), so an image is one thread's
// work: lane 0 of the image's warp runs sdm_host::EdRouteFixed - the same source the host threads run, on fixed arrays in
// global memory - on the anchors k_ed_sort has put in order, after the 32 lanes have filled the image's edge-index plane with -1.  One warp per block: the device takes
// 32 resident images per SM, 4736 in flight on a B200; a step costs a few dependent L2 / HBM round trips (about 1 us), which
// only pays when many images are routed at once - the host threads need 1.4 ms per image and thread, but there are 16-32 of
// them against thousands of warps.  result[img] = {chains, chain pixels, 1 if complete (0: a capacity ran out, the host routes
// the image instead), 0}.
struct EdRouteBatch {
    int W, H, grad_thresh;
    const int16_t* G;      // [n][H][W] stage-1 planes
    uint8_t* F;            // [n][H][W], modified in place
    uint8_t* scratch;      // n * scratch_stride bytes
    size_t scratch_stride;
    sdm_host::EdRouteCaps caps;
    int32_t* offsets;      // [n][caps.offsets]
    uint32_t* pixels;      // [n][caps.out_pixels]
    int32_t* edge_index;   // [n][H][W] or NULL
    int4* result;          // [n]
    const int* n_anchors;  // [n] from k_ed_sort
    long long* prof;       // [n][16] or NULL: cycles of {edge fill + anchor sort, pass over the anchors, walks, extraction}, walked pixels, trees
};

__global__ void __launch_bounds__(32) k_ed_route(EdRouteBatch b)
{
    const int img = blockIdx.x, lane = threadIdx.x;
    const size_t P = (size_t)b.W * b.H;
    const int16_t* G = b.G + (size_t)img * P;
    uint8_t* F = b.F + (size_t)img * P;
    uint8_t* scratch = b.scratch + (size_t)img * b.scratch_stride;
    int32_t* edge = b.edge_index ? b.edge_index + (size_t)img * P : nullptr;
    const long long t_start = clock64();
    if (edge) {
        if ((P & 3) == 0) {
            int4* e4 = reinterpret_cast<int4*>(edge);
            for (size_t i = lane; i < P / 4; i += 32) e4[i] = make_int4(-1, -1, -1, -1);
        } else {
            for (size_t i = lane; i < P; i += 32) edge[i] = -1;
        }
    }
    const int total = b.n_anchors[img];  // k_ed_sort left the anchors in the scratch block's anchor slots, in walking order
    if (total < 0) {  // (cannot happen with EdRouteCapsFor's P / 2; the host routes the image if it does)
        if (lane == 0) b.result[img] = make_int4(0, 0, 0, 0);
        return;
    }
    __syncwarp();
    if (lane == 0) {
        int nc = 0, np = 0;
        long long* prof = b.prof ? b.prof + (size_t)img * 16 : nullptr;
        if (prof) prof[5] = clock64() - t_start;
        const bool ok = sdm_host::EdRouteFixed(b.W, b.H, G, F, b.grad_thresh, scratch, b.caps, b.offsets + (size_t)img * b.caps.offsets,
                                               b.pixels + (size_t)img * b.caps.out_pixels, edge, (size_t)b.W * 4, &nc, &np, total, prof);
        if (prof) prof[6] = clock64() - t_start;
        b.result[img] = make_int4(nc, np, ok ? 1 : 0, 0);
    }
}

// kf->mEdgeIndex of a batch on the device from its chain lists (host routing, masks kept on the device): the planes hold -1
// already; pixel j of image `img` gets the number of the chain it belongs to (binary search in the image's chain offsets)
struct EdMaskImage { unsigned long long pix_begin, off_begin; int n_chains, n_pixels; };
__global__ void __launch_bounds__(256) k_ed_mask(const EdMaskImage* __restrict__ images, const uint32_t* __restrict__ pixels,
                                                 const int32_t* __restrict__ offsets, int W, int H, int32_t* __restrict__ masks)
{
    const EdMaskImage im = images[blockIdx.y];
    const uint32_t* px = pixels + im.pix_begin;
    const int32_t* off = offsets + im.off_begin;  // n_chains + 1 entries, off[0] = 0
    int32_t* mask = masks + (size_t)blockIdx.y * (size_t)W * (size_t)H;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < im.n_pixels; j += gridDim.x * blockDim.x) {
        int lo = 0, hi = im.n_chains;  // the chain with off[c] <= j < off[c + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (off[mid] <= j) lo = mid; else hi = mid;
        }
        const uint32_t p = px[j];
        mask[(size_t)(p >> 16) * W + (p & 0xffffu)] = lo;
    }
}

// chain lists of a batch as one contiguous block (one D2H copy instead of two per image): at[i] = first int32 of image i's
// list - its offsets (chains + 1) followed by its pixels - at[n] = total; images that did not complete take no room
__global__ void __launch_bounds__(1024) k_ed_chain_offsets(const int4* __restrict__ result, int n, unsigned long long* __restrict__ at)
{
    __shared__ unsigned long long warp_sum[32];
    const int i = threadIdx.x, lane = i & 31, w = i >> 5;
    unsigned long long v = 0;
    if (i < n) { const int4 r = result[i]; v = r.z ? (unsigned long long)r.x + 1ull + (unsigned long long)r.y : 0ull; }
    unsigned long long inc = v;
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warp_sum[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned long long s = warp_sum[lane], t = s;
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += o; }
        warp_sum[lane] = t - s;  // exclusive
    }
    __syncthreads();
    const unsigned long long excl = warp_sum[w] + inc - v;
    if (i < n) at[i] = excl;
    if (i == n - 1) at[n] = excl + v;
}

__global__ void __launch_bounds__(256) k_ed_chain_gather(EdRouteBatch b, const unsigned long long* __restrict__ at, int32_t* __restrict__ dst,
                                                         unsigned long long room /* int32 the block holds */)
{
    const int img = blockIdx.y;
    const int4 r = b.result[img];
    if (!r.z || at[img] + (unsigned long long)r.x + 1ull + (unsigned long long)r.y > room) return;  // (the host checks the total)
    const int32_t* off = b.offsets + (size_t)img * b.caps.offsets;
    const uint32_t* px = b.pixels + (size_t)img * b.caps.out_pixels;
    int32_t* d = dst + at[img];
    const int n_off = r.x + 1, total = n_off + r.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        d[i] = i < n_off ? off[i] : (int32_t)px[i - n_off];
}
#endif

}  // namespace sdm
