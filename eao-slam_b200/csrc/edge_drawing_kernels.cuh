// edge_drawing_kernels.cuh - stage 1 of the Edge Drawing detector on the device (sm_100a), for a batch of keyframes.
//
// The reference calls the closed-source EDLib once per keyframe inside pass 1 (LineDetector::DetectEdgeMap,
// /root/reference/src/LineDetector.cc:843-881: DetectEdgesByED(srcImg, width, height, SOBEL_OPERATOR, 36, 8, 1.0)); its chains
// are the mEdgeIndex candidate mask of the hot loop (ProbabilityMapping.cc:454) and the input of the 3-D line fitting
// (:884-900).  eao-slam_b200/host/edge_drawing.h documents the open implementation and how every open detail was pinned to the
// library's output.  The detector has two stages: per-pixel image work (smoothing, Sobel gradient, edge direction, anchor test)
// and a sequential walk from the anchors.  This kernel is the first stage, the walk runs on host threads (sdm_edge_drawing in
// sdm_b200.cu) while the next chunk of keyframes is on the device.
//
//   in : im  [n][H][W] u8                                                  1 byte per pixel
//   out: G   [n][H][W] i16  |gx| + |gy| of Sobel on the smoothed image; border pixels = grad_thresh - 1
//        F   [n][H][W] u8   bits 0-1: 0 below grad_thresh, 1 vertical edge pixel (|gx| >= |gy|), 2 horizontal;
//                           bit 7: anchor (rows / columns 2 .. size-3: G exceeds both neighbours across the edge by
//                           anchor_thresh)                                 3 bytes per pixel
// Integer arithmetic throughout: the planes are bit-identical to sdm_host::EdPlanesHost.
//
// One block = one 64 x 16 tile of one keyframe, everything between the image bytes and the two outputs stays in shared
// memory: image tile with a 4-pixel apron (2 smoothing + 1 Sobel + 1 anchor test) -> separable 1-4-6-4-1 smoothing (row sums
// as u16, column pass with the library's rounding: half to even where its 4-wide vector loop runs, half up in the last W % 4
// columns) -> gradient + direction (packed in one u16) -> anchor test -> 8-byte (G) and 4-byte (F) stores per thread.
// HBM-bound by construction (4 algorithmic bytes per pixel, no re-reads besides the apron, which L2 serves).
#pragma once

#include <cstdint>

namespace sdm {

constexpr int kEdTW = 64, kEdTH = 16, kEdThreads = 256;
constexpr int kEdFlagAnchor = 0x80;

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
                if (dir != 0 && y >= 2 && y <= H - 3 && xx >= 2 && xx <= W - 3) {
                    const int n0 = dir == 1 ? (s_g[ty + 1][tx + q] & 0xfff) : (s_g[ty][tx + q + 1] & 0xfff);
                    const int n1 = dir == 1 ? (s_g[ty + 1][tx + q + 2] & 0xfff) : (s_g[ty + 2][tx + q + 1] & 0xfff);
                    if (g - n0 >= anchor_thresh && g - n1 >= anchor_thresh) f |= kEdFlagAnchor;
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

}  // namespace sdm
