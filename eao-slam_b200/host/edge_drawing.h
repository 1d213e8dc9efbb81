// edge_drawing.h - open implementation of the Edge Drawing edge-segment detector that EAO-SLAM calls through the
// closed-source Thirdparty/EDTest/EDLib.a:
//     EdgeMap* map = DetectEdgesByED(srcImg, width, height, SOBEL_OPERATOR, 36, 8, 1.0);      LineDetector.cc:855
// (LineDetector::DetectEdgeMap, :843-881: the edge chains feed LineFitting, :884-900, and their pixels are the mEdgeIndex mask
// of the semi-dense hot loop, ProbabilityMapping.cc:454).  SURVEY.md 8(a17) / 8(f-2): "the upstream DetectEdgeMap mask (would
// need an open ED re-implementation)".
//
// The algorithm is the published one (C. Topal, C. Akinlar, "Edge Drawing: A Combined Real-Time Edge and Segment Detector",
// JVCIR 23(6), 2012): smooth, gradient magnitude + direction, anchors, smart routing from the anchors in decreasing gradient
// order into a tree of chains, the longest path of the tree as the segment, the other long chains as further segments.  The
// details that a black-box library leaves open were fixed by matching its OUTPUT: the binary runs in the build container,
// so every stage was compared against it (oracle/ed_chains.cpp; tests/golden/ed_chains_*.npz hold its chains) until the
// chains were identical - pixel for pixel, in order - on every test image: the six + 32 synthetic keyframes at 320x240 / 640x480,
// the library's own lena.pgm (flipped, transposed, cropped, subsampled), and some 500 random images - noise, drawn shapes, smooth
// blobs quantised to a few grey levels, sizes 12 .. 700 not multiples of four (tests/test_edge_drawing.py; fixtures
// tests/golden/ed_chains_small.npz and ed_chains_misc.npz by oracle/make_ed_golden.py):
//   * smoothing  = 5x5 binomial [1 4 6 4 1]^2 / 256, replicated border; rounding as cvSmooth(CV_GAUSSIAN, 5, 5) of the bundled
//                  OpenCV 2.4.5 does it: half to even in the columns its 4-wide vector loop covers (x < W & ~3), half up in
//                  the scalar tail (checked against that cvSmooth itself through its C API, tests/test_edge_drawing.py);
//   * gradient   = |gx| + |gy| of the Sobel operator on the smoothed image, border pixels = threshold - 1; a pixel at or
//                  above the threshold is a VERTICAL edge pixel if |gx| >= |gy|, else HORIZONTAL;
//   * anchors    = every pixel of rows / columns 2 .. size-3 whose gradient exceeds both neighbours across the edge by >= 8;
//   * routing    = from the anchors in decreasing gradient order (ties in raster order), depth first, the more recently
//                  opened direction first (up before down, left before right); a walk looks for an already marked anchor /
//                  edge pixel among its three forward neighbours - straight first, then the diagonals in an order that is
//                  rotation-symmetric: up-left for LEFT and UP, down-right for RIGHT and DOWN - before it follows the largest
//                  gradient; anchors beside a walked pixel (across the edge) are cleared;
//   * extraction = trees with fewer than 10 new pixels are erased; the main segment is the longest path through the anchor,
//                  further segments are the remaining chains whose longest path has >= 10 pixels; where two chains meet, pixels
//                  that double back are dropped, and the first chain of a segment is compared with the LAST PIXEL OF THE
//                  PREVIOUS SEGMENT (the library keeps all segments in one pixel array and reads one element before the
//                  segment it is filling) - reproduced because it decides whether that chain's first pixel is kept.
// Host code like the reference's call (it runs once per keyframe before the planes are uploaded; the routing is a sequential
// walk).  Header-only, C++11, no dependencies.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace sdm_host {

struct EdgeChains {
    std::vector<int32_t> offsets;  // n_chains + 1, offsets[0] = 0
    std::vector<uint32_t> pixels;  // (row << 16) | col, chain k = pixels[offsets[k] .. offsets[k + 1])
    int n_chains() const { return offsets.empty() ? 0 : (int)offsets.size() - 1; }
};

#if defined(__CUDACC__)
#define SDM_EDR_HD __host__ __device__
#define SDM_EDR_UNROLL4 _Pragma("unroll 4")
#define SDM_EDR_UNROLL8 _Pragma("unroll 8")
#if defined(__CUDA_ARCH__)
#define SDM_EDR_CLOCK() clock64()
#else
#define SDM_EDR_CLOCK() 0ll
#endif
// the routing templates are __host__ __device__ and are instantiated on std::vector for the host threads: nvcc warns about the
// host-only calls of that (host-only) instantiation
#pragma nv_diag_suppress 20011, 20014
#else
#define SDM_EDR_HD
#define SDM_EDR_UNROLL4
#define SDM_EDR_UNROLL8
#define SDM_EDR_CLOCK() 0ll
#endif

namespace ed_detail {

enum { kEdgeVertical = 1, kEdgeHorizontal = 2, kLeft = 1, kRight = 2, kUp = 3, kDown = 4 };

struct Px { int r, c; };
struct Chain { int len, parent, dir, child[2], start; };
struct Todo { int r, c, dir, parent; };

SDM_EDR_HD inline int iabs(int v) { return v < 0 ? -v : v; }
SDM_EDR_HD inline bool adjacent(const Px& a, const Px& b) { return iabs(a.r - b.r) <= 1 && iabs(a.c - b.c) <= 1; }

// The routing core (EdRouteCore below) is written once against this small vector interface: on the host it runs on
// std::vector (HostVec), on the device - one image per warp, k_ed_route in csrc/edge_drawing_kernels.cuh - on fixed arrays in
// global memory (FixedVec), where a push beyond the capacity is dropped and flagged; the core gives up on the image as soon as
// a flag is up, and the library routes that image on the host instead.
template <class T>
struct HostVec {
    std::vector<T> v;
    int size() const { return (int)v.size(); }
    bool empty() const { return v.empty(); }
    void clear() { v.clear(); }
    void push_back(const T& x) { v.push_back(x); }
    void pop_back() { v.pop_back(); }
    T& back() { return v.back(); }
    T& operator[](int i) { return v[(size_t)i]; }
    const T& operator[](int i) const { return v[(size_t)i]; }
    void resize(int n) { v.resize((size_t)n); }
    void append(const T* src, int count, int step)  // count elements from src, walking by step (+1 / -1)
    {
        const size_t n0 = v.size();
        v.resize(n0 + (size_t)count);
        T* d = v.data() + n0;
        for (int i = 0; i < count; ++i) d[i] = src[(ptrdiff_t)i * step];
    }
    bool overflow() const { return false; }
};
template <class T>
struct FixedVec {
    T* p;
    int n, cap;
    bool ovf;
    SDM_EDR_HD FixedVec(T* mem, int capacity) : p(mem), n(0), cap(capacity), ovf(false) {}
    SDM_EDR_HD int size() const { return n; }
    SDM_EDR_HD bool empty() const { return n == 0; }
    SDM_EDR_HD void clear() { n = 0; }
    SDM_EDR_HD void push_back(const T& x) { if (n < cap) p[n++] = x; else ovf = true; }
    SDM_EDR_HD void pop_back() { if (n > 0) --n; }
    SDM_EDR_HD T& back() { return p[n > 0 ? n - 1 : 0]; }
    SDM_EDR_HD T& operator[](int i) { return p[i]; }
    SDM_EDR_HD const T& operator[](int i) const { return p[i]; }
    SDM_EDR_HD void resize(int m) { if (m <= cap) n = m; else { n = cap; ovf = true; } }
    SDM_EDR_HD void append(const T* src, int count, int step)  // independent iterations: the loads of four are in flight together
    {
        if (count > cap - n) { ovf = true; return; }
        T* d = p + n;
        int i = 0;
        for (; i + 8 <= count; i += 8) {  // eight loads, then eight stores (src and d may alias for all the compiler knows)
            T v[8];
            SDM_EDR_UNROLL8
            for (int k = 0; k < 8; ++k) v[k] = src[(ptrdiff_t)(i + k) * step];
            SDM_EDR_UNROLL8
            for (int k = 0; k < 8; ++k) d[i + k] = v[k];
        }
        for (; i < count; ++i) d[i] = src[(ptrdiff_t)i * step];
        n += count;
    }
    SDM_EDR_HD bool overflow() const { return ovf; }
};

// best[k] for every chain of a finished tree: the chain's length plus the longer of the two paths below it, a zero-length child
// counting - with everything below it - as absent.  A child is created after its parent, so one pass from the last chain down
// sees every child before its parent: no traversal, no pointer chasing (the recursive descent this replaces - one per queried
// chain, each visiting the whole subtree - was two fifths of the segment extraction).
template <class ChainVec, class IntVec>
SDM_EDR_HD inline void longest_paths(const ChainVec& ch, IntVec& best)
{
    for (int k = ch.size() - 1; k >= 1; --k) {
        const Chain& c = ch[k];
        const int l0 = (c.child[0] != -1 && ch[c.child[0]].len != 0) ? best[c.child[0]] : 0;
        const int l1 = (c.child[1] != -1 && ch[c.child[1]].len != 0) ? best[c.child[1]] : 0;
        best[k] = c.len + (l0 >= l1 ? l0 : l1);
    }
}

// Length of the longest path below `root` (from longest_paths); every chain ON that path keeps only the child the path continues
// with (first child on ties; the last chain keeps its first child link whatever it points to) - what retrieve_chain_nos follows.
// Valid as long as no chain below `root` has been copied since longest_paths ran, which holds for every query the extraction
// makes: the chains it copies in between lie on paths that do not meet the queried subtree.
template <class ChainVec, class IntVec>
SDM_EDR_HD inline int longest_chain(ChainVec& ch, int root, const IntVec& best)
{
    if (root == -1 || ch[root].len == 0) return 0;
    for (int k = root; k != -1;) {
        Chain& c = ch[k];
        const int l0 = (c.child[0] != -1 && ch[c.child[0]].len != 0) ? best[c.child[0]] : 0;
        const int l1 = (c.child[1] != -1 && ch[c.child[1]].len != 0) ? best[c.child[1]] : 0;
        if (l0 >= l1) { c.child[1] = -1; k = l0 > 0 ? c.child[0] : -1; }
        else { c.child[0] = -1; k = c.child[1]; }
    }
    return best[root];
}

template <class ChainVec, class IntVec>
SDM_EDR_HD inline void retrieve_chain_nos(const ChainVec& ch, int root, IntVec& nos)
{
    nos.clear();
    while (root != -1) {
        nos.push_back(root);
        root = ch[root].child[0] != -1 ? ch[root].child[0] : ch[root].child[1];
    }
}

}  // namespace ed_detail

// The detector runs in two stages.  Stage 1 (EdPlanesHost) is per-pixel image work - smoothing, gradient, direction, the routing
// choice of every edge pixel, anchor test - and produces two planes: G (int16 gradient magnitude) and F (one flag byte per
// pixel).  Stage 2 (EdRouteChains / EdRouteCore) is the sequential part: anchor order, smart routing, segment extraction.
// libsdm_b200.so runs stage 1 on the device for a batch of keyframes (k_ed_planes4 / k_ed_planes, k_ed_sort:
// csrc/edge_drawing_kernels.cuh; C-ABI sdm_edge_drawing) and stage 2 either here on host threads while the next keyframes are
// on the device, or on the device as well (k_ed_route: one warp per keyframe running EdRouteFixed, the same source); both
// stage-1 forms produce identical planes and both stage-2 forms identical chains (tests/test_gpu_edge_drawing.py).
enum : uint8_t {
    kEdDirMask = 3,       // F & 3: 0 = below the gradient threshold, 1 = vertical edge pixel, 2 = horizontal edge pixel
    // bits 2-3 / 4-5: where a walk that finds no marked pixel ahead goes from this pixel when it walks backwards (left / up) /
    // forwards (right / down) along the pixel's own edge direction - the largest of the three gradients ahead, decided in
    // stage 1 where all pixels are decided at once (ed_route_code); the walk itself then never reads the gradient plane
    kEdRouteShiftBack = 2, kEdRouteShiftFwd = 4, kEdRouteMask = 3,
    kEdStaticMask = 0x3f, // direction + routes: what stage 1 writes and the routing never changes
    kEdFlagEdge = 0x40,   // set by the routing: pixel belongs to a walked chain
    kEdFlagAnchor = 0x80  // set by stage 1: anchor; cleared by the routing when the pixel is walked or suppressed
};
// A, B, C = gradients of the three pixels ahead, A on the up / left side, B straight, C on the down / right side:
// 0 = straight, 1 = to A's side, 2 = to C's side (the rule of the smart-routing step)
SDM_EDR_HD inline int ed_route_code(int A, int B, int C) { return A > B ? (A > C ? 1 : 2) : (C > B ? 2 : 0); }

// Stage 1 on the host.  G and F are dense W x H planes.  grad_thresh / anchor_thresh = the GRADIENT_THRESH / ANCHOR_THRESH
// arguments of DetectEdgesByED (36 / 8 at LineDetector.cc:855; SOBEL_OPERATOR, sigma 1.0).  Needs W >= 5 and H >= 5.
inline void EdPlanesHost(const uint8_t* im, size_t step, int W, int H, int grad_thresh, int anchor_thresh, int16_t* G, uint8_t* F)
{
    using namespace ed_detail;
    const size_t P = (size_t)W * H;
    // ---- smoothing: [1 4 6 4 1] x [1 4 6 4 1] / 256, replicated border
    static thread_local std::vector<uint16_t> tmp;  // (work buffers are kept between calls: one call per keyframe)
    static thread_local std::vector<uint8_t> sm;
    tmp.resize(P);
    sm.resize(P);
    for (int y = 0; y < H; ++y) {
        const uint8_t* s = im + (size_t)y * step;
        uint16_t* t = &tmp[(size_t)y * W];
        for (int x = 0; x < 2; ++x) t[x] = (uint16_t)(s[0] + 4 * s[std::max(x - 1, 0)] + 6 * s[x] + 4 * s[x + 1] + s[x + 2]);
        for (int x = 2; x < W - 2; ++x) t[x] = (uint16_t)(s[x - 2] + 4 * s[x - 1] + 6 * s[x] + 4 * s[x + 1] + s[x + 2]);
        for (int x = W - 2; x < W; ++x) t[x] = (uint16_t)(s[x - 2] + 4 * s[x - 1] + 6 * s[x] + 4 * s[std::min(x + 1, W - 1)] + s[W - 1]);
    }
    const int wsimd = W & ~3;
    for (int y = 0; y < H; ++y) {
        const uint16_t* r0 = &tmp[(size_t)std::max(y - 2, 0) * W];
        const uint16_t* r1 = &tmp[(size_t)std::max(y - 1, 0) * W];
        const uint16_t* r2 = &tmp[(size_t)y * W];
        const uint16_t* r3 = &tmp[(size_t)std::min(y + 1, H - 1) * W];
        const uint16_t* r4 = &tmp[(size_t)std::min(y + 2, H - 1) * W];
        uint8_t* o = &sm[(size_t)y * W];
        for (int x = 0; x < wsimd; ++x) {  // half to even (the library's vector loop)
            const int v = r0[x] + 4 * r1[x] + 6 * r2[x] + 4 * r3[x] + r4[x];
            const int q = v >> 8, rem = v & 255;
            o[x] = (uint8_t)(q + ((rem > 128) | ((rem == 128) & (q & 1))));
        }
        for (int x = wsimd; x < W; ++x)  // half up (its scalar tail)
            o[x] = (uint8_t)((r0[x] + 4 * r1[x] + 6 * r2[x] + 4 * r3[x] + r4[x] + 128) >> 8);
    }
    // ---- gradient magnitude (Sobel, |gx| + |gy|) and edge direction
    std::fill(G, G + P, (int16_t)(grad_thresh - 1));
    std::memset(F, 0, P);
#define SM(y, x) ((int)sm[(size_t)(y) * W + (x)])
    for (int y = 1; y < H - 1; ++y)
        for (int x = 1; x < W - 1; ++x) {
            const int com1 = SM(y + 1, x + 1) - SM(y - 1, x - 1), com2 = SM(y - 1, x + 1) - SM(y + 1, x - 1);
            const int gx = std::abs(com1 + com2 + 2 * (SM(y, x + 1) - SM(y, x - 1)));
            const int gy = std::abs(com1 - com2 + 2 * (SM(y + 1, x) - SM(y - 1, x)));
            const int sum = gx + gy;
            G[(size_t)y * W + x] = (int16_t)sum;
            if (sum >= grad_thresh) F[(size_t)y * W + x] = gx >= gy ? kEdgeVertical : kEdgeHorizontal;
        }
#undef SM
    // ---- where an unguided walk goes from each edge pixel (both senses along its direction)
    for (int y = 1; y < H - 1; ++y)
        for (int x = 1; x < W - 1; ++x) {
            uint8_t& f = F[(size_t)y * W + x];
            if (f == 0) continue;
            const int16_t* g = G + (size_t)y * W + x;
            int back, fwd;
            if (f == kEdgeHorizontal) { back = ed_route_code(g[-W - 1], g[-1], g[W - 1]); fwd = ed_route_code(g[-W + 1], g[1], g[W + 1]); }
            else { back = ed_route_code(g[-W - 1], g[-W], g[-W + 1]); fwd = ed_route_code(g[W - 1], g[W], g[W + 1]); }
            f = (uint8_t)(f | (back << kEdRouteShiftBack) | (fwd << kEdRouteShiftFwd));
        }
    // ---- anchors: gradient exceeds both neighbours across the edge by anchor_thresh
    for (int y = 2; y < H - 2; ++y) {
        const int16_t* g0 = G + (size_t)(y - 1) * W;
        const int16_t* g1 = G + (size_t)y * W;
        const int16_t* g2 = G + (size_t)(y + 1) * W;
        uint8_t* f1 = F + (size_t)y * W;
        for (int x = 2; x < W - 2; ++x) {
            const int g = g1[x];
            if (g < grad_thresh) continue;
            const bool is_anchor = (f1[x] & kEdDirMask) == kEdgeVertical ? (g - g1[x - 1] >= anchor_thresh && g - g1[x + 1] >= anchor_thresh)
                                                          : (g - g0[x] >= anchor_thresh && g - g2[x] >= anchor_thresh);
            if (is_anchor) f1[x] |= kEdFlagAnchor;
        }
    }
}

// Stage 2: anchors in decreasing gradient order, smart routing, segment extraction.  F is modified in place (flags of the
// walked pixels).  If edge_index is given (int32 plane, row pitch edge_step bytes, ALREADY FILLED WITH -1 by the caller:
// KeyFrame.cc:87) every chain pixel receives its chain number in chain order (LineDetector.cc:857-866).
// Storage (see HostVec / FixedVec): found / anchors: anchor positions; chains, pixels, stack, seg: the tree being walked;
// best / order / nos: work lists of the segment extraction; out_offsets / out_pixels: the result, offsets[0] = 0 pushed here.
// Returns false if a FixedVec ran out of capacity (the outputs are then incomplete and must be discarded).
template <class IntVec, class ListVec, class ChainVec, class PxVec, class TodoVec, class OffVec, class PixVec>
SDM_EDR_HD inline bool EdRouteCore(int W, int H, const int16_t* G, uint8_t* F, int grad_thresh, IntVec& found, IntVec& anchors,
                                   ChainVec& chains, PxVec& pixels, PxVec& seg, TodoVec& stack, ListVec& best, ListVec& order, ListVec& nos,
                                   OffVec& out_offsets, PixVec& out_pixels, int32_t* edge_index, size_t edge_step, int* hist /* [2049] */,
                                   int presorted = -1 /* >= 0: `anchors` already holds that many positions in walking order */,
                                   long long* prof = nullptr /* device: cycles {anchor scan, walks, extraction}, walked pixels, trees */)
{
    using namespace ed_detail;
    (void)grad_thresh;  // (direction 0 in F = below the threshold: the walk reads nothing else)
    (void)order;        // (work list of the recursive form of longest_chain; kept in the interface)
    out_offsets.clear();
    out_offsets.push_back(0);
    out_pixels.clear();
    if (W < 5 || H < 5) return true;
    const size_t P = (size_t)W * H;
#define GR(y, x) ((int)G[(size_t)(y) * W + (x)])
#define FL(y, x) F[(size_t)(y) * W + (x)]
#define DI(y, x) (FL(y, x) & kEdDirMask)
#define MARKED(y, x) ((FL(y, x) & (kEdFlagAnchor | kEdFlagEdge)) != 0)
#define IS_ANCHOR(y, x) ((FL(y, x) & kEdFlagAnchor) != 0)
#define IS_EDGE(y, x) ((FL(y, x) & kEdFlagEdge) != 0)
#define SDM_EDR_OVERFLOW() (found.overflow() || anchors.overflow() || chains.overflow() || pixels.overflow() || seg.overflow() || \
                            stack.overflow() || best.overflow() || order.overflow() || nos.overflow() || out_offsets.overflow() ||  \
                            out_pixels.overflow())
    // ---- anchor order: decreasing gradient, raster order among equals (counting sort: |gx| + |gy| <= 2040)
    if (presorted >= 0) {
        anchors.resize(presorted);
        if (anchors.overflow()) return false;
    } else {
        found.clear();
        for (int i = 0; i <= 2048; ++i) hist[i] = 0;
        {
            size_t i = 0;
            while (i < P && ((size_t)(F + i) & 7) != 0) {
                if (F[i] & kEdFlagAnchor) { found.push_back((int)i); ++hist[2047 - (G[i] < 2047 ? (int)G[i] : 2047)]; }
                ++i;
            }
            for (; i + 8 <= P; i += 8) {  // eight flag bytes at a time
                const uint64_t w = *reinterpret_cast<const uint64_t*>(F + i);
                if ((w & 0x8080808080808080ull) == 0) continue;
                for (int k = 0; k < 8; ++k)
                    if (F[i + k] & kEdFlagAnchor) { found.push_back((int)(i + k)); ++hist[2047 - (G[i + k] < 2047 ? (int)G[i + k] : 2047)]; }
            }
            for (; i < P; ++i)
                if (F[i] & kEdFlagAnchor) { found.push_back((int)i); ++hist[2047 - (G[i] < 2047 ? (int)G[i] : 2047)]; }
        }
        if (found.overflow()) return false;
        {
            int run = 0;
            for (int i = 0; i <= 2047; ++i) { const int c = hist[i]; hist[i] = run; run += c; }
            anchors.resize(found.size());
            if (anchors.overflow()) return false;
            for (int i = 0; i < found.size(); ++i) { const int g = G[found[i]]; anchors[hist[2047 - (g < 2047 ? g : 2047)]++] = found[i]; }
        }

    }

    bool have_prev = false;
    Px prev_last = {0, 0};  // last pixel of the previously emitted segment (see the header comment)

    // emits `seg` as the next chain
#define SDM_EDR_EMIT()                                                                                                       \
    do {                                                                                                                     \
        const int32_t id = (int32_t)out_offsets.size() - 1;                                                                  \
        const int base_ = out_pixels.size(), cnt_ = seg.size();                                                              \
        out_pixels.resize(base_ + cnt_);                                                                                     \
        if (!out_pixels.overflow()) {                                                                                        \
            int i_ = 0;                                                                                                      \
            for (; i_ + 8 <= cnt_; i_ += 8) { /* eight loads, then the stores */                                             \
                Px v_[8];                                                                                                    \
                SDM_EDR_UNROLL8                                                                                              \
                for (int k_ = 0; k_ < 8; ++k_) v_[k_] = seg[i_ + k_];                                                        \
                SDM_EDR_UNROLL8                                                                                              \
                for (int k_ = 0; k_ < 8; ++k_) {                                                                             \
                    out_pixels[base_ + i_ + k_] = ((uint32_t)v_[k_].r << 16) | (uint32_t)v_[k_].c;                           \
                    if (edge_index)                                                                                          \
                        reinterpret_cast<int32_t*>(reinterpret_cast<char*>(edge_index) + (size_t)v_[k_].r * edge_step)[v_[k_].c] = id; \
                }                                                                                                            \
            }                                                                                                                \
            for (; i_ < cnt_; ++i_) {                                                                                        \
                const Px v_ = seg[i_];                                                                                       \
                out_pixels[base_ + i_] = ((uint32_t)v_.r << 16) | (uint32_t)v_.c;                                            \
                if (edge_index) reinterpret_cast<int32_t*>(reinterpret_cast<char*>(edge_index) + (size_t)v_.r * edge_step)[v_.c] = id; \
            }                                                                                                                \
        }                                                                                                                    \
        out_offsets.push_back((int32_t)out_pixels.size());                                                                   \
        if (!seg.empty()) { prev_last = seg.back(); have_prev = true; }                                                      \
    } while (0)
    // appends chain `cno_` to `seg` in walking order, dropping the pixels that double back at the joint
#define SDM_EDR_APPEND_FORWARD(cno_)                                                                                         \
    do {                                                                                                                     \
        Chain& c_ = chains[cno_];                                                                                            \
        const Px first_ = pixels[c_.start];                                                                                  \
        int idx_ = seg.size() - 2;                                                                                           \
        while (idx_ >= 0 && adjacent(first_, seg[idx_])) { seg.pop_back(); --idx_; }                                         \
        int start_ = 0;                                                                                                      \
        if (c_.len > 1 && (!seg.empty() || have_prev)) {                                                                     \
            const Px last_ = seg.empty() ? prev_last : seg.back();                                                           \
            if (adjacent(pixels[c_.start + 1], last_)) start_ = 1;                                                           \
        }                                                                                                                    \
        if (c_.len > start_) seg.append(&pixels[c_.start + start_], c_.len - start_, 1);                                     \
        c_.len = 0; /* copied */                                                                                             \
    } while (0)

    long long t_walk = 0, t_extract = 0, n_walked = 0, n_trees = 0, t_x[5] = {0, 0, 0, 0, 0};
    const long long t_begin = SDM_EDR_CLOCK();
    // The pass over the anchor list reads eight positions and their flags at a time (independent loads, in flight together).
    // An anchor flag is only ever cleared: a clear bit in the batch's copy is final; a set bit is read again if a walk has
    // run since the batch was loaded.
    int batch_pos[8];
    uint8_t batch_flag[8];
    bool walked_since = false;
    for (int a = 0; a < anchors.size(); ++a) {
        const int ab = a & 7;
        if (ab == 0) {
            const int nb8 = anchors.size() - a < 8 ? anchors.size() - a : 8;
            SDM_EDR_UNROLL8
            for (int k = 0; k < 8; ++k) batch_pos[k] = k < nb8 ? anchors[a + k] : 0;
            SDM_EDR_UNROLL8
            for (int k = 0; k < 8; ++k) batch_flag[k] = F[batch_pos[k]];
            walked_since = false;
        }
        if (!(batch_flag[ab] & kEdFlagAnchor)) continue;
        const int ay = batch_pos[ab] / W, ax = batch_pos[ab] % W;
        if (walked_since && !IS_ANCHOR(ay, ax)) continue;
        walked_since = true;
        const long long t_a = SDM_EDR_CLOCK();
        chains.clear();
        pixels.clear();
        stack.clear();
        Chain root = {0, -1, 0, {-1, -1}, 0};
        chains.push_back(root);
        int dup = 0;
        if (DI(ay, ax) == kEdgeVertical) {
            Todo t0 = {ay, ax, kDown, 0}, t1 = {ay, ax, kUp, 0};
            stack.push_back(t0);
            stack.push_back(t1);
        } else {
            Todo t0 = {ay, ax, kRight, 0}, t1 = {ay, ax, kLeft, 0};
            stack.push_back(t0);
            stack.push_back(t1);
        }
        while (!stack.empty()) {
            const Todo t = stack.back();
            stack.pop_back();
            int r = t.r, c = t.c;
            const int dir = t.dir;
            if (!IS_EDGE(r, c)) ++dup;
            Chain ch = {0, t.parent, dir, {-1, -1}, pixels.size()};
            const int no = chains.size();
            chains.push_back(ch);
            Px p0 = {r, c};
            pixels.push_back(p0);
            if (chains.overflow() || pixels.overflow() || stack.overflow()) return false;
            int clen = 1;
            const bool horizontal = dir == kLeft || dir == kRight;
            const int child = (dir == kLeft || dir == kUp) ? 0 : 1;
            const int fwd = (dir == kLeft || dir == kUp) ? -1 : 1;  // step along the walk; also the diagonal looked at first
            bool ended = false;
            // One step: mark the pixel, clear the anchors beside it (across the edge), look at the three forward neighbours.
            // Their flags and gradients are loaded together up front and the chosen neighbour's values are carried into the
            // next step (nothing this step writes touches them): one round trip to memory per step instead of a chain of
            // dependent ones - what a step costs on the device, where a load is an L2 access.
            const int want = horizontal ? kEdgeHorizontal : kEdgeVertical;
            // offsets in pixels (the planes hold fewer than 2^31): one pixel forward, one pixel down / right, the two diagonals
            const int along = horizontal ? fwd : fwd * W, across = horizontal ? W : 1;
            const int o_f1 = along + fwd * across, o_f2 = along - fwd * across;
            const int route_shift = fwd < 0 ? kEdRouteShiftBack : kEdRouteShiftFwd;
            uint8_t* pf = F + ((size_t)r * W + c);
            uint8_t fcur = *pf;
            bool full = false;
            while ((fcur & kEdDirMask) == want) {
                // forward neighbours: straight, the diagonal on the `fwd` side, the other diagonal; the two pixels beside this one
                const uint8_t f0 = pf[along], f1 = pf[o_f1], f2 = pf[o_f2];
                const uint8_t s0 = pf[-across], s1 = pf[across];  // (all five loads before the first store: one round trip)
                pf[0] = (uint8_t)((fcur & kEdStaticMask) | kEdFlagEdge);
                pf[-across] = (uint8_t)(s0 & ~kEdFlagAnchor);
                pf[across] = (uint8_t)(s1 & ~kEdFlagAnchor);
                int side;  // -1 / 0 / +1: offset across the walk of the pixel taken
                if (f0 & (kEdFlagAnchor | kEdFlagEdge)) side = 0;
                else if (f1 & (kEdFlagAnchor | kEdFlagEdge)) side = fwd;
                else if (f2 & (kEdFlagAnchor | kEdFlagEdge)) side = -fwd;
                else {  // the largest gradient ahead, as stage 1 found it
                    const int code = (fcur >> route_shift) & kEdRouteMask;
                    side = code == 0 ? 0 : (code == 1 ? -1 : 1);
                }
                pf += along + side * across;
                if (horizontal) { r += side; c += fwd; } else { r += fwd; c += side; }
                fcur = side == 0 ? f0 : (side == fwd ? f1 : f2);
                if ((fcur & kEdFlagEdge) || (fcur & kEdDirMask) == 0) {  // met an edge or left the gradient ridge (direction 0 = below the threshold)
                    ended = true;
                    break;
                }
                Px pn = {r, c};
                pixels.push_back(pn);
                if (pixels.overflow()) { full = true; break; }  // (one-level exits only inside the step loop: it is the hot path)
                ++clen;
            }
            if (full) return false;
            if (ended) {
                chains[no].len = clen;
                chains[t.parent].child[child] = no;
            }
            if (ended) continue;
            // the edge turns: the last pixel opens the two walks across the old direction and belongs to them
            if (horizontal) {
                Todo t0 = {r, c, kDown, no}, t1 = {r, c, kUp, no};
                stack.push_back(t0);
                stack.push_back(t1);
            } else {
                Todo t0 = {r, c, kRight, no}, t1 = {r, c, kLeft, no};
                stack.push_back(t0);
                stack.push_back(t1);
            }
            pixels.pop_back();
            --clen;
            chains[no].len = clen;
            chains[t.parent].child[child] = no;
        }
        const long long t_b = SDM_EDR_CLOCK();
        t_walk += t_b - t_a;
        n_walked += pixels.size();
        ++n_trees;
        if (pixels.size() - dup < 10) {  // too short: take the walk back
            {
                int i = 0;
                for (; i + 4 <= pixels.size(); i += 4) {  // positions, then flags, then the stores: two round trips per four pixels
                    size_t q[4];
                    uint8_t v[4];
                    SDM_EDR_UNROLL4
                    for (int k = 0; k < 4; ++k) q[k] = (size_t)pixels[i + k].r * W + pixels[i + k].c;
                    SDM_EDR_UNROLL4
                    for (int k = 0; k < 4; ++k) v[k] = F[q[k]];
                    // (a pixel may be in the list twice - the pixel a chain turned at starts both children: the stores are idempotent)
                    SDM_EDR_UNROLL4
                    for (int k = 0; k < 4; ++k) F[q[k]] = (uint8_t)(v[k] & kEdStaticMask);
                }
                for (; i < pixels.size(); ++i) FL(pixels[i].r, pixels[i].c) &= kEdStaticMask;
            }
            t_extract += SDM_EDR_CLOCK() - t_b;
            continue;
        }
        best.resize(chains.size());
        if (best.overflow()) return false;
        longest_paths(chains, best);
        // ---- main segment: longest path of the second direction backwards, the anchor, longest path of the first direction
        seg.clear();
        const long long t_x0 = SDM_EDR_CLOCK();
        const int l_second = longest_chain(chains, chains[0].child[1], best);
        t_x[0] += SDM_EDR_CLOCK() - t_x0;
        if (l_second > 0) {
            retrieve_chain_nos(chains, chains[0].child[1], nos);
            for (int k = nos.size(); k-- > 0;) {
                Chain& c = chains[nos[k]];
                const Px lastp = pixels[c.start + c.len - 1];
                int idx = seg.size() - 2;
                while (idx >= 0 && adjacent(lastp, seg[idx])) { seg.pop_back(); --idx; }
                if (c.len > 1 && (!seg.empty() || have_prev) &&
                    adjacent(pixels[c.start + c.len - 2], seg.empty() ? prev_last : seg.back()))
                    --c.len;
                if (c.len > 0) seg.append(&pixels[c.start + c.len - 1], c.len, -1);
                c.len = 0;
            }
        }
        const long long t_x1 = SDM_EDR_CLOCK();
        t_x[1] += t_x1 - t_x0;  // (selection of the second direction + its copies)
        if (longest_chain(chains, chains[0].child[0], best) > 1) {
            retrieve_chain_nos(chains, chains[0].child[0], nos);
            ++chains[nos[0]].start;  // the anchor is already there
            --chains[nos[0]].len;
            for (int k = 0; k < nos.size(); ++k) SDM_EDR_APPEND_FORWARD(nos[k]);
        }
        if (seg.size() > 1 && adjacent(seg[1], seg.back())) {  // drop the first pixel
            for (int i = 1; i < seg.size(); ++i) seg[i - 1] = seg[i];
            seg.pop_back();
        }
        const long long t_x2 = SDM_EDR_CLOCK();
        t_x[2] += t_x2 - t_x1;  // (first direction + the closing test)
        SDM_EDR_EMIT();
        const long long t_x3 = SDM_EDR_CLOCK();
        t_x[3] += t_x3 - t_x2;  // (emit of the main segment)
        // ---- the other long chains of the tree
        for (int k = 2; k < chains.size(); ++k) {
            if (chains[k].len < 2) continue;
            if (longest_chain(chains, k, best) >= 10) {
                retrieve_chain_nos(chains, k, nos);
                seg.clear();
                for (int q = 0; q < nos.size(); ++q) SDM_EDR_APPEND_FORWARD(nos[q]);
                SDM_EDR_EMIT();
            }
        }
        t_x[4] += SDM_EDR_CLOCK() - t_x3;  // (the other chains)
        t_extract += SDM_EDR_CLOCK() - t_b;
        if (SDM_EDR_OVERFLOW()) return false;
    }
    if (prof) {
        prof[0] = SDM_EDR_CLOCK() - t_begin - t_walk - t_extract;  // the pass over the anchor list
        prof[1] = t_walk;
        prof[2] = t_extract;
        prof[3] = n_walked;
        prof[4] = n_trees;
        for (int k = 0; k < 5; ++k) prof[8 + k] = t_x[k];  // extraction: longest path 2nd direction / + its copies / 1st direction / emit / other chains
    }
    return !SDM_EDR_OVERFLOW();
#undef SDM_EDR_EMIT
#undef SDM_EDR_APPEND_FORWARD
#undef SDM_EDR_OVERFLOW
#undef GR
#undef FL
#undef DI
#undef MARKED
#undef IS_ANCHOR
#undef IS_EDGE
}

// The core on std::vector: what the routing threads of sdm_edge_drawing and the host-only detector call.  Fills edge_index
// with -1 first (KeyFrame.cc:87).
// sorted_anchors / n_sorted (>= 0): the anchors already in walking order (k_ed_sort on the device), else they are sorted here.
inline void EdRouteChains(int W, int H, const int16_t* G, uint8_t* F, int grad_thresh, EdgeChains& out,
                          int32_t* edge_index = nullptr, size_t edge_step = 0, const int32_t* sorted_anchors = nullptr, int n_sorted = -1)
{
    using namespace ed_detail;
    if (edge_index)
        for (int y = 0; y < H; ++y) {
            int32_t* row = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(edge_index) + (size_t)y * edge_step);
            for (int x = 0; x < W; ++x) row[x] = -1;
        }
    static thread_local HostVec<int> found, anchors, best, order, nos;  // (work lists are kept between calls)
    static thread_local HostVec<Chain> chains;
    static thread_local HostVec<Px> pixels, seg;
    static thread_local HostVec<Todo> stack;
    HostVec<int32_t> offsets;
    HostVec<uint32_t> px;
    int hist[2048 + 1];
    offsets.v.swap(out.offsets);
    px.v.swap(out.pixels);
    if (sorted_anchors && n_sorted >= 0) anchors.v.assign(sorted_anchors, sorted_anchors + n_sorted);
    else n_sorted = -1;
    EdRouteCore(W, H, G, F, grad_thresh, found, anchors, chains, pixels, seg, stack, best, order, nos, offsets, px, edge_index, edge_step, hist,
                n_sorted);
    offsets.v.swap(out.offsets);
    px.v.swap(out.pixels);
}

// The core on fixed arrays: what one warp of k_ed_route runs per image (lane 0), and what tests/cpp/test_edge_drawing.cpp
// replays on the CPU against the std::vector form.  Capacities are generous for real images (a VGA keyframe of the bench
// scene needs 12 k anchors, ~30 chains and ~300 pixels per tree, 31 k chain pixels); an image that exceeds one returns false.
struct EdRouteCaps {
    int anchors, chains, pixels, offsets, out_pixels;
};
SDM_EDR_HD inline EdRouteCaps EdRouteCapsFor(size_t P)
{
    EdRouteCaps c;
    c.anchors = (int)(P / 2) + 16;     // an anchor exceeds both neighbours across its edge: no two are adjacent in that direction
    c.chains = (int)(P / 8) + 16;      // chains / open walks of ONE tree
    c.pixels = (int)(P / 2) + 16;      // walked pixels of ONE tree
    c.offsets = (int)(P / 8) + 16;     // chains of the image + 1
    c.out_pixels = (int)(P / 2) + 16;  // chain pixels of the image
    return c;
}
SDM_EDR_HD inline size_t EdRouteAlign(size_t v) { return (v + 15) & ~(size_t)15; }
SDM_EDR_HD inline size_t EdRouteScratchBytes(const EdRouteCaps& c)
{
    using namespace ed_detail;
    return 2 * EdRouteAlign((size_t)c.anchors * sizeof(int)) + EdRouteAlign((size_t)c.chains * sizeof(Chain)) +
           2 * EdRouteAlign((size_t)c.pixels * sizeof(Px)) + EdRouteAlign((size_t)c.chains * sizeof(Todo)) +
           3 * EdRouteAlign((size_t)c.chains * sizeof(int)) + EdRouteAlign(2049 * sizeof(int));
}
// where a caller that sorts the anchors itself (k_ed_route: all 32 lanes of the warp) leaves them: caps.anchors ints
SDM_EDR_HD inline int* EdRouteAnchorSlots(uint8_t* scratch, const EdRouteCaps& caps)
{
    return reinterpret_cast<int*>(scratch + EdRouteAlign((size_t)caps.anchors * sizeof(int)));
}
// scratch: EdRouteScratchBytes(caps) bytes, 16-byte aligned; out_offsets[caps.offsets], out_pixels[caps.out_pixels];
// edge_index (may be NULL) already filled with -1.  *n_chains / *n_pixels are valid when true is returned.
SDM_EDR_HD inline bool EdRouteFixed(int W, int H, const int16_t* G, uint8_t* F, int grad_thresh, uint8_t* scratch, const EdRouteCaps& caps,
                                    int32_t* out_offsets, uint32_t* out_pixels, int32_t* edge_index, size_t edge_step, int* n_chains,
                                    int* n_pixels, int presorted = -1 /* see EdRouteAnchorSlots */, long long* prof = nullptr)
{
    using namespace ed_detail;
    uint8_t* p = scratch;
    FixedVec<int> found(reinterpret_cast<int*>(p), caps.anchors);      p += EdRouteAlign((size_t)caps.anchors * sizeof(int));
    FixedVec<int> anchors(reinterpret_cast<int*>(p), caps.anchors);    p += EdRouteAlign((size_t)caps.anchors * sizeof(int));
    FixedVec<Chain> chains(reinterpret_cast<Chain*>(p), caps.chains);  p += EdRouteAlign((size_t)caps.chains * sizeof(Chain));
    FixedVec<Px> pixels(reinterpret_cast<Px*>(p), caps.pixels);        p += EdRouteAlign((size_t)caps.pixels * sizeof(Px));
    FixedVec<Px> seg(reinterpret_cast<Px*>(p), caps.pixels);           p += EdRouteAlign((size_t)caps.pixels * sizeof(Px));
    FixedVec<Todo> stack(reinterpret_cast<Todo*>(p), caps.chains);     p += EdRouteAlign((size_t)caps.chains * sizeof(Todo));
    FixedVec<int> best(reinterpret_cast<int*>(p), caps.chains);        p += EdRouteAlign((size_t)caps.chains * sizeof(int));
    FixedVec<int> order(reinterpret_cast<int*>(p), caps.chains);       p += EdRouteAlign((size_t)caps.chains * sizeof(int));
    FixedVec<int> nos(reinterpret_cast<int*>(p), caps.chains);         p += EdRouteAlign((size_t)caps.chains * sizeof(int));
    int* hist = reinterpret_cast<int*>(p);
    FixedVec<int32_t> offs(out_offsets, caps.offsets);
    FixedVec<uint32_t> px(out_pixels, caps.out_pixels);
    const bool ok = EdRouteCore(W, H, G, F, grad_thresh, found, anchors, chains, pixels, seg, stack, best, order, nos, offs, px,
                                edge_index, edge_step, hist, presorted, prof);
    *n_chains = offs.size() - 1;
    *n_pixels = px.size();
    return ok;
}

// Both stages on the host: what the reference's call does per keyframe (image row pitch `step` bytes).  Images up to
// 65535 x 65535.
inline void DetectEdgesByED(const uint8_t* im, size_t step, int W, int H, int grad_thresh, int anchor_thresh, EdgeChains& out,
                            int32_t* edge_index = nullptr, size_t edge_step = 0)
{
    static thread_local std::vector<int16_t> G;
    static thread_local std::vector<uint8_t> F;
    if (W >= 5 && H >= 5) {
        G.resize((size_t)W * H);
        F.resize((size_t)W * H);
        EdPlanesHost(im, step, W, H, grad_thresh, anchor_thresh, G.data(), F.data());
    }
    EdRouteChains(W, H, G.data(), F.data(), grad_thresh, out, edge_index, edge_step);
}

}  // namespace sdm_host
