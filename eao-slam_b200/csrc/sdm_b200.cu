// sdm_b200.cu — C-ABI implementation of libsdm_b200.so (see include/sdm_b200.h).
//
// Host side of the B200-native semi-dense mapping path: device arena, keyframe upload, batched
// persistent launches of the pass-1 / pass-2 kernels (sdm_kernels.cuh), downloads, CUDA-IPC halo
// pulls.  Replaces the CPU loops of yanmin-wu/EAO-SLAM src/ProbabilityMapping.cc:348-597.
// No CPU fallback: every entry point that computes needs a CUDA device.
//
// Streams: H2D copies on s_copy (into a ring of staging sets), every kernel — k_pack, the passes — on
// s_compute, D2H copies on s_down.  The two copy streams carry DMA only, so they never compete with the
// persistent pass kernels for SM slots.  Ordering is tracked per keyframe slot with event ids (see
// EventRing): k_pack waits for its staging set's H2D and for downloads still reading the slot, a pass
// waits for downloads of the planes it overwrites, a download waits for the last kernel on its slot.
// Nothing on these paths blocks the host, so a caller keeps all three engines busy by issuing
// upload / pass / download calls on chunks of keyframes (bench.py e2e).
//
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false
//        -Xcompiler -fPIC,-ffp-contract=off -shared   (see __graft_entry__.build()).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sdm_b200.h"
#include "pair_geometry.h"
#include "sdm_kernels.cuh"
#include "edge_drawing_kernels.cuh"
#include "../host/edge_drawing.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SDM_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define RC(call)                    \
    do {                            \
        int rc_ = (call);           \
        if (rc_ != SDM_OK) return rc_; \
    } while (0)

#ifndef SDM_DEFAULT_SCAN_GEN
#define SDM_DEFAULT_SCAN_GEN 3  // scan loop generation of a context when env SDM_SCAN does not say (lane1 | lane2 | lane3 | warp)
#endif
constexpr int kUpStages = 48;   // upload staging sets (im, grad, theta, edge): H2D runs this far ahead of k_pack
constexpr int kDownStages = 6;  // (rho | sigma) split staging sets
constexpr int kItemStages = 32;  // pinned work-order staging buffers (how many passes the host may run ahead)
constexpr int kMaxPeers = 16;
constexpr int kLongScanColumns = 128;  // mean search range (columns) above which sdm_pass1 launches k_pass1_lane's long-scan build
constexpr int kIntraChunk = 64;  // keyframes per batched intra launch (bounds the tmp arena)

// A ring of CUDA events addressed by monotonically increasing ids.  id 0 = "never".  An id that has
// fallen out of the ring is complete by construction (its event was synchronised when recycled).
struct EventRing {
    std::vector<cudaEvent_t> ev;
    uint64_t next = 1;
    int init(int n)
    {
        ev.resize(n);
        for (auto& e : ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return SDM_OK;
    }
    void destroy()
    {
        for (auto e : ev)
            if (e) cudaEventDestroy(e);
        ev.clear();
    }
    int record(cudaStream_t s, uint64_t* id)
    {
        const uint64_t i = next++;
        cudaEvent_t e = ev[i % ev.size()];
        if (i > ev.size()) CU(cudaEventSynchronize(e));  // recycle: the old occupant must be complete
        CU(cudaEventRecord(e, s));
        *id = i;
        return SDM_OK;
    }
    bool live(uint64_t id) const { return id != 0 && next - id <= ev.size(); }
    int wait(cudaStream_t s, uint64_t id)
    {
        if (live(id)) CU(cudaStreamWaitEvent(s, ev[id % ev.size()], 0));
        return SDM_OK;
    }
    int host_sync(uint64_t id)
    {
        if (live(id)) CU(cudaEventSynchronize(ev[id % ev.size()]));
        return SDM_OK;
    }
    bool done(uint64_t id) { return !live(id) || cudaEventQuery(ev[id % ev.size()]) == cudaSuccess; }
};

// developer timeline (env SDM_TRACE=1): timing events around H2D batches / packs / passes, printed by sdm_synchronize
struct TraceRec {
    const char* what;
    int n;
    cudaEvent_t a, b;
};

struct KfState {
    bool uploaded = false;
    bool pass1_done = false;
    bool rs_dense = false;  // (rho,sigma) plane did not come from pass 1: pass 2 must visit every pixel
    bool split_stale = false;  // dpl / spl do not mirror rs (external write): downloads de-interleave rs instead
    float K[4] = {0, 0, 0, 0};
    float Tcw[12] = {0};
    uint64_t comp_id = 0;     // last kernel (k_pack, pass, pull) touching the slot
    uint64_t down_ds_id = 0;  // last download of depth_map_ / depth_sigma_ (pass 1 and the intra stencils overwrite them)
    uint64_t down_cp_id = 0;  // last download of depth_map_checked_ / points (pass 2 overwrites them)
    uint64_t count_id = 0;    // compute-ring id after which h_cand[slot] holds the slot's candidate count
    uint64_t pull_id = 0;     // compute-ring id of a halo pull (sdm_exchange) into the slot, recorded on the halo stream
};

struct UpStage {
    uint8_t* im = nullptr;
    float* grad = nullptr;
    float* theta = nullptr;
    int32_t* edge = nullptr;
    uint64_t busy = 0;  // compute-ring id of the k_pack that reads it
};
struct DownStage {
    float* planes = nullptr;  // 2 * P floats
    uint64_t busy = 0;        // down-ring id
};
// Work orders of one pass: pinned host staging + its own device copy.  The H2D goes out on the COPY stream at
// call time (in host order, i.e. ahead of the bulk uploads of later chunks; on the compute stream it would sit in
// the copy engine behind them until the previous pass finished); the pass's kernels wait for it by event.
struct ItemStage {
    void* host = nullptr;  // pinned: DevItem[cap] followed by int aux[3 * cap]
    sdm::DevItem* d_items = nullptr;  // = the current pass's ItemStage buffers
    int* d_aux = nullptr;
    int cap = 0;
    uint64_t busy = 0;  // compute-ring id after the last kernel that reads the device copy
};

// sdm_scatter_keyframes: sparse D2H + host-side scatter.  The calling thread enqueues k_gather_sparse and ONE
// contiguous D2H per keyframe into a ring of staging buffers; worker threads wait for the copy's event and write
// the records into the caller's planes.
constexpr int kScatStages = 16;
struct ScatStage {
    uint32_t* dev = nullptr;
    uint32_t* host = nullptr;  // pinned
    cudaEvent_t ev = nullptr;
    bool free_ = true;
};
struct ScatJob {
    int stage;
    int n;
    int W;
    sdm_download_desc d;
};
struct Scatter {
    ScatStage st[kScatStages];
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_free, cv_idle;
    std::deque<ScatJob> jobs;
    int next = 0;
    long pending = 0;
    bool stop = false;
    std::string err;
};

void scatter_records(const ScatJob& j, const uint32_t* rec)
{
    const size_t n = (size_t)j.n;
    const uint32_t* pix = rec;
    const float* f = reinterpret_cast<const float*>(rec);
    const float *rho = f + n, *sig = f + 2 * n, *chk = f + 3 * n, *pts = f + 4 * n;
    const sdm_download_desc& d = j.d;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t p = pix[i];
        const size_t x = p & 0xffffu, y = p >> 16;
        if (d.depth) *reinterpret_cast<float*>(reinterpret_cast<char*>(d.depth) + y * d.depth_step + 4 * x) = rho[i];
        if (d.sigma) *reinterpret_cast<float*>(reinterpret_cast<char*>(d.sigma) + y * d.sigma_step + 4 * x) = sig[i];
        if (d.checked) *reinterpret_cast<float*>(reinterpret_cast<char*>(d.checked) + y * d.checked_step + 4 * x) = chk[i];
        if (d.points) {
            float* q = reinterpret_cast<float*>(reinterpret_cast<char*>(d.points) + y * d.points_step + 12 * x);
            q[0] = pts[3 * i]; q[1] = pts[3 * i + 1]; q[2] = pts[3 * i + 2];
        }
    }
}

void scatter_worker(Scatter* S, int device)
{
    cudaSetDevice(device);
    for (;;) {
        ScatJob j;
        {
            std::unique_lock<std::mutex> lk(S->mu);
            S->cv_job.wait(lk, [&] { return S->stop || !S->jobs.empty(); });
            if (S->jobs.empty()) return;  // stop requested and nothing left
            j = S->jobs.front();
            S->jobs.pop_front();
        }
        const cudaError_t e = cudaEventSynchronize(S->st[j.stage].ev);
        if (e == cudaSuccess) scatter_records(j, S->st[j.stage].host);
        {
            std::lock_guard<std::mutex> lk(S->mu);
            if (e != cudaSuccess && S->err.empty()) S->err = cudaGetErrorString(e);
            S->st[j.stage].free_ = true;
            --S->pending;
        }
        S->cv_free.notify_all();
        S->cv_idle.notify_all();
    }
}

}  // namespace

struct sdm_ctx {
    sdm_config cfg;
    sdm::DevParams P;
    sdm::DevArena A;
    size_t npix = 0;
    std::vector<KfState> kf;
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_down = nullptr;
    EventRing r_copy, r_compute, r_down;
    cudaEvent_t ev_p1[2] = {nullptr, nullptr}, ev_p2[2] = {nullptr, nullptr}, ev_p1_scan = nullptr;
    std::vector<cudaEvent_t> ev_pack;  // start / stop pairs around the pack kernels of the last sdm_upload_keyframes
    int n_pack_ev = 0;
    cudaEvent_t marks[SDM_N_MARKS] = {nullptr};
    bool mark_set[SDM_N_MARKS] = {false};
    bool p1_timed = false, p2_timed = false;
    UpStage up[kUpStages];
    int up_next = 0;
    DownStage down[kDownStages];
    int down_next = 0;
    ItemStage ist[kItemStages];
    int ist_next = 0;
    int* h_count = nullptr;  // pinned, sdm_candidate_count
    int* h_cand = nullptr;   // pinned + device-visible: candidate count per slot, published by k_publish_counts
    Scatter* scat = nullptr;  // created by the first sdm_scatter_keyframes
    // device work orders of the current pass
    sdm::DevItem* d_items = nullptr;  // = the current pass's ItemStage buffers
    int* d_aux = nullptr;
    int* d_chunk_off = nullptr;
    int* d_counter = nullptr;
    int items_cap = 0;
    sdm::DevStats* d_stats = nullptr;
    float2* tmp_rs = nullptr;  // kIntraChunk planes: the other side of the intra ping-pong
    int tmp_planes = 0;
    float* xfer = nullptr;  // 2 dense float planes (sdm_upload_depth staging)
    float* dbg = nullptr;   // 4 float planes + 1 byte plane (sdm_epipolar_search_plane)
    void* exp_buf = nullptr;  // sdm_export_points: block counts | offsets | keyframe totals, then the compacted points
    size_t exp_cap = 0;
    sdm_point* exp_pts = nullptr;
    size_t exp_pts_cap = 0;
    void* lf_buf = nullptr;  // sdm_line_fit: chain lists, worst-case line slots, compacted lines
    size_t lf_cap = 0;
    void* lf_host = nullptr;  // pinned staging of the chain lists
    size_t lf_host_cap = 0;
    cudaEvent_t lf_ev[2] = {nullptr, nullptr};
    float lf_ms = 0.f;
    // sdm_edge_drawing: device planes (im | G | F) and their pinned mirrors for up to ed_cap keyframes
    uint8_t* ed_dev = nullptr;
    uint8_t* ed_host = nullptr;
    int ed_cap = 0;
    cudaStream_t s_ed = nullptr;
    std::vector<cudaEvent_t> ed_ev;  // per chunk: kernel start, kernel stop, planes on the host
    float ed_kernel_ms = 0.f, ed_wall_ms = 0.f, ed_route_ms = 0.f;
    // device routing (sdm_set_edge_drawing_route): per-image scratch, chain lists, edge-index planes, results
    int ed_route_mode = 0;
    uint8_t* edr_dev = nullptr;
    int4* edr_result_host = nullptr;
    int edr_cap = 0;
    int edr_fallbacks = 0;
    int edr_last_base = 0, edr_last_n = 0;  // images of the last device-routed batch whose edge-index planes are still in edr_dev
    // SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE: edge-index planes of the last batch, scattered from the chains by k_ed_mask
    uint8_t* edm_dev = nullptr;
    size_t edm_bytes = 0;
    int edm_last_base = 0, edm_last_n = 0;
    cudaEvent_t edr_ev[2] = {nullptr, nullptr};
    void* peer_rs[kMaxPeers] = {nullptr};
    // sdm_exchange: flag blocks (own + IPC-mapped peers), halo plan, step counter
    sdm::XFlags* xflags = nullptr;
    sdm::XFlags* peer_flags[kMaxPeers] = {nullptr};
    void* peer_rs2[kMaxPeers] = {nullptr};  // peers' (rho,sigma) arenas opened by sdm_import_peer
    int peer_nslots[kMaxPeers] = {0};
    int my_rank = -1;
    unsigned xstep = 0, acks_waited = 0;
    std::vector<int32_t> halo_local, halo_rank, halo_slot;
    std::vector<char> is_halo;  // per slot
    cudaStream_t s_halo = nullptr;
    cudaStream_t s_down_k = nullptr;  // sparse_download = 2: the block-sparse kernel for SemiDensePointSets_ runs here, next to the DMAs of s_down
    cudaEvent_t ev_down_k = nullptr;
    int grid_pass1_warp = 0, grid_pass2 = 0, grid_intra = 0, n_sm = 0;  // persistent grids (blocks)
    long long launches = 0;
    bool trace = false;
    cudaEvent_t trace_base = nullptr;
    std::vector<TraceRec> trace_recs;
    bool last_scan_long = false;  // the last sdm_pass1 ran the long-scan instantiation of k_pass1_lane
    int scan_warp_per_pixel = 0;  // developer A/B knob (env SDM_SCAN=warp): the warp-per-pixel scan kernel; 2 = warp_tma:
                                  // the same with neighbour tiles staged in shared memory by TMA (k_pass1_tma)
    CUtensorMap tex_map;          // the texel arena as a TMA tensor {4 floats, W, slots * H} (warp_tma only)
    int grid_pass1_tma = 0;
};

namespace {

bool slot_ok(const sdm_ctx* c, int s) { return s >= 0 && s < (int)c->kf.size(); }

void trace_begin(sdm_ctx* c, cudaStream_t s, const char* what, int n)
{
    if (!c->trace) return;
    TraceRec r{what, n, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    c->trace_recs.push_back(r);
}
void trace_end(sdm_ctx* c, cudaStream_t s)
{
    if (!c->trace) return;
    cudaEventRecord(c->trace_recs.back().b, s);
}
void trace_dump(sdm_ctx* c)
{
    if (!c->trace || c->trace_recs.empty()) return;
    for (auto& r : c->trace_recs) {
        float t0 = 0, t1 = 0;
        cudaEventElapsedTime(&t0, c->trace_base, r.a);
        cudaEventElapsedTime(&t1, c->trace_base, r.b);
        fprintf(stderr, "[sdm trace] %-8s n=%3d  %9.3f -> %9.3f ms\n", r.what, r.n, t0, t1);
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    c->trace_recs.clear();
    cudaEventRecord(c->trace_base, c->s_compute);
}

// pitched plane copy; rows that are contiguous on both sides go out as ONE 1-D copy (cheaper to enqueue
// and a single DMA descriptor), which is the common case for cv::Mat planes that are not ROIs
int copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind,
           cudaStream_t s)
{
    if (dpitch == width && spitch == width)
        CU(cudaMemcpyAsync(dst, src, width * height, kind, s));
    else
        CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, s));
    return SDM_OK;
}
dim3 tile_grid(const sdm_ctx* c, int z = 1) { return dim3((c->cfg.width + 31) / 32, (c->cfg.height + 7) / 8, z); }

int ensure_items(sdm_ctx* c, int n)
{
    if (n <= c->items_cap) return SDM_OK;
    const int cap = std::max(n, std::max(64, c->items_cap * 2));
    CU(cudaStreamSynchronize(c->s_compute));
    cudaFree(c->d_chunk_off);
    c->d_chunk_off = nullptr;
    c->items_cap = 0;
    CU(cudaMalloc(&c->d_chunk_off, sizeof(int) * (cap + 1)));
    c->items_cap = cap;
    return SDM_OK;
}

// staging for n work orders whose previous user (a pass) has finished
int acquire_item_stage(sdm_ctx* c, int n, ItemStage** out)
{
    // a free stage that is already large enough, else any free one (allocated below): in steady state a loop cycles
    // through the two or three stages its passes keep in flight instead of growing all kItemStages (each allocation is a
    // cudaMallocHost + two cudaMallocs, milliseconds with the device idle, for work orders of a thousand keyframes)
    ItemStage* st = nullptr;
    for (int pass = 0; pass < 2 && !st; ++pass)
        for (int k = 0; k < kItemStages; ++k) {
            ItemStage& cand = c->ist[(c->ist_next + k) % kItemStages];
            if ((pass == 1 || cand.cap >= n) && c->r_compute.done(cand.busy)) { st = &cand; break; }
        }
    if (!st) {
        st = &c->ist[c->ist_next];
        RC(c->r_compute.host_sync(st->busy));
    }
    c->ist_next = (int)((st - c->ist) + 1) % kItemStages;
    if (st->cap < n) {
        if (st->host) CU(cudaFreeHost(st->host));
        cudaFree(st->d_items);
        cudaFree(st->d_aux);
        st->host = nullptr; st->d_items = nullptr; st->d_aux = nullptr;
        st->cap = 0;
        const int cap = std::max(n, 32);
        CU(cudaMallocHost(&st->host, (sizeof(sdm::DevItem) + 3 * sizeof(int)) * cap));
        CU(cudaMalloc(&st->d_items, sizeof(sdm::DevItem) * cap));
        CU(cudaMalloc(&st->d_aux, sizeof(int) * 3 * cap));
        st->cap = cap;
    }
    *out = st;
    return SDM_OK;
}

// SemiDenseLoop :424-438 per keyframe: R21/t21/F12 for every neighbour, Twc for the point set
int build_item(sdm_ctx* c, const sdm_item& in, sdm::DevItem& out, bool pass2, bool points_only)
{
    if (!slot_ok(c, in.kf)) return fail(SDM_ERR_ARG, "item keyframe slot %d out of range", in.kf);
    if (in.n_nbr < 0 || in.n_nbr > SDM_MAX_NBR) return fail(SDM_ERR_ARG, "n_nbr %d out of [0,%d]", in.n_nbr, SDM_MAX_NBR);
    const KfState& k1 = c->kf[in.kf];
    if (!k1.uploaded) return fail(SDM_ERR_STATE, "keyframe slot %d not uploaded", in.kf);
    if (pass2 && !points_only && !k1.pass1_done) return fail(SDM_ERR_STATE, "pass 2 before pass 1 for slot %d", in.kf);
    memset(&out, 0, offsetof(sdm::DevItem, pair));
    out.kf = in.kf;
    out.n_nbr = in.n_nbr;
    out.min_depth = in.min_depth;
    out.max_depth = in.max_depth;
    memcpy(out.K, k1.K, sizeof(out.K));
    sdm::pose_inverse(k1.Tcw, out.Twc);
    for (int j = 0; j < in.n_nbr; ++j) {
        const int s2 = in.nbr[j];
        if (!slot_ok(c, s2)) return fail(SDM_ERR_ARG, "neighbour slot %d out of range", s2);
        const KfState& k2 = c->kf[s2];
        if (pass2) {
            if (!k2.pass1_done) return fail(SDM_ERR_STATE, "neighbour slot %d has no pass-1 planes", s2);
        } else if (!k2.uploaded) {
            return fail(SDM_ERR_STATE, "neighbour slot %d not uploaded", s2);
        }
        const sdm::PairGeometry g = sdm::pair_geometry(k1.K, k1.Tcw, k2.K, k2.Tcw);
        sdm::DevPair& p = out.pair[j];
        memcpy(p.F, g.F12.m, sizeof(p.F));
        memcpy(p.R, g.R21.m, sizeof(p.R));
        memcpy(p.t, g.t21.v, sizeof(p.t));
        p.rot = in.rot_deg[j];
        p.slot = s2;
        memcpy(p.K2, k2.K, sizeof(p.K2));
    }
    return SDM_OK;
}

struct Batch {
    int n = 0;
    int n_sparse = 0, n_dense = 0;
    int* d_slots = nullptr;
    int* d_order_sparse = nullptr;
    int* d_order_dense = nullptr;
    ItemStage* stage = nullptr;
};

// Build the device work orders of a pass and make s_compute wait for everything the pass depends on.
int prepare_batch(sdm_ctx* c, int n, const sdm_item* items, bool pass2, bool points_only, Batch* b)
{
    RC(ensure_items(c, n));
    ItemStage* st;
    RC(acquire_item_stage(c, n, &st));
    sdm::DevItem* h_items = (sdm::DevItem*)st->host;
    int* h_aux = (int*)(h_items + st->cap);
    int* h_slots = h_aux;
    int* h_sparse = h_aux + n;
    int* h_dense = h_aux + 2 * n;
    uint64_t need_down = 0;
    b->n = n;
    for (int i = 0; i < n; ++i) {
        RC(build_item(c, items[i], h_items[i], pass2, points_only));
        const KfState& k1 = c->kf[items[i].kf];
        h_slots[i] = items[i].kf;
        if (pass2 && (k1.rs_dense || points_only)) h_dense[b->n_dense++] = i; else h_sparse[b->n_sparse++] = i;
        need_down = std::max(need_down, pass2 ? k1.down_cp_id : k1.down_ds_id);  // planes this pass overwrites
    }
    RC(c->r_down.wait(c->s_compute, need_down));  // (uploads need no wait: k_pack runs on s_compute itself)
    CU(cudaMemcpyAsync(st->d_items, h_items, sizeof(sdm::DevItem) * n, cudaMemcpyHostToDevice, c->s_copy));
    CU(cudaMemcpyAsync(st->d_aux, h_aux, sizeof(int) * 3 * n, cudaMemcpyHostToDevice, c->s_copy));
    uint64_t h2d = 0;
    RC(c->r_copy.record(c->s_copy, &h2d));
    RC(c->r_copy.wait(c->s_compute, h2d));
    c->d_items = st->d_items;
    c->d_aux = st->d_aux;
    b->d_slots = c->d_aux;
    b->d_order_sparse = c->d_aux + n;
    b->d_order_dense = c->d_aux + 2 * n;
    b->stage = st;
    return SDM_OK;
}

// after the pass's kernels are enqueued: every referenced slot (and the work-order staging) is busy until this
// point of s_compute
int finish_batch(sdm_ctx* c, int n, const sdm_item* items, const Batch& b)
{
    uint64_t id = 0;
    RC(c->r_compute.record(c->s_compute, &id));
    b.stage->busy = id;
    for (int i = 0; i < n; ++i) {
        c->kf[items[i].kf].comp_id = id;
        for (int j = 0; j < items[i].n_nbr; ++j) c->kf[items[i].nbr[j]].comp_id = id;
    }
    return SDM_OK;
}

sdm::DevPlan make_plan(sdm_ctx* c, const int* d_order, int n)
{
    sdm::DevPlan p;
    p.chunk_off = c->d_chunk_off;
    p.counter = c->d_counter;
    p.order = d_order;
    p.n_items = n;
    return p;
}

// batched intra stencils on the (rho,sigma) planes of `n` slots (device array d_slots)
int run_intra(sdm_ctx* c, const int* d_slots, int n, bool check, bool grow)
{
    if (!check && !grow) return SDM_OK;
    const int want = std::min(n, kIntraChunk);
    if (c->tmp_planes < want) {
        CU(cudaStreamSynchronize(c->s_compute));
        cudaFree(c->tmp_rs);
        c->tmp_rs = nullptr;
        c->tmp_planes = 0;
        CU(cudaMalloc(&c->tmp_rs, (size_t)want * c->npix * sizeof(float2)));
        c->tmp_planes = want;
    }
    for (int i0 = 0; i0 < n; i0 += c->tmp_planes) {
        const int m = std::min(c->tmp_planes, n - i0);
        sdm::k_intra_check<<<tile_grid(c, m), dim3(32, 8), 0, c->s_compute>>>(c->P, c->A.rs, c->tmp_rs, d_slots + i0, c->npix,
                                                                              check ? 0 : 1);
        sdm::k_intra_grow<<<tile_grid(c, m), dim3(32, 8), 0, c->s_compute>>>(c->P, c->tmp_rs, c->A.rs, c->A.dpl, c->A.spl, c->A.tex,
                                                                             d_slots + i0, c->npix, grow ? 0 : 1);
        c->launches += 2;
    }
    CU(cudaGetLastError());
    return SDM_OK;
}

int single_slot_array(sdm_ctx* c, int slot, int** d_slots, ItemStage** stage)
{
    ItemStage* st;
    RC(acquire_item_stage(c, 1, &st));
    int* h = (int*)st->host;
    h[0] = slot;
    CU(cudaMemcpyAsync(st->d_aux, h, sizeof(int), cudaMemcpyHostToDevice, c->s_copy));
    uint64_t h2d = 0;
    RC(c->r_copy.record(c->s_copy, &h2d));
    RC(c->r_copy.wait(c->s_compute, h2d));
    *d_slots = st->d_aux;
    *stage = st;
    return SDM_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

void sdm_default_config(sdm_config* cfg)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->width = 640;
    cfg->height = 480;
    cfg->max_keyframes = 16;
    cfg->lambdaG = 8;
    cfg->lambdaL = 80;
    cfg->lambdaTheta = 45;
    cfg->lambdaN = 3;
    cfg->theta = (float)0.23;
    cfg->sigmaI = 20.0f;
    cfg->chi2_fusion = 5.99;
    cfg->chi2_inter = 3.84;
    cfg->eps = 0.000001;
    cfg->slope_max = 4.0f;
    cfg->intra_check = 0;
    cfg->intra_grow = 0;
    cfg->device = 0;
}

const char* sdm_last_error(void) { return g_last_error.c_str(); }
const char* sdm_version(void) { return "sdm_b200 0.2 (sm_100a)"; }

void sdm_destroy(sdm_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kMaxPeers; ++i) {
        if (c->peer_rs[i]) cudaIpcCloseMemHandle(c->peer_rs[i]);
        if (c->peer_rs2[i]) cudaIpcCloseMemHandle(c->peer_rs2[i]);
        if (c->peer_flags[i]) cudaIpcCloseMemHandle(c->peer_flags[i]);
    }
    cudaFree(c->xflags);
    if (c->s_halo) cudaStreamDestroy(c->s_halo);
    if (c->s_down_k) cudaStreamDestroy(c->s_down_k);
    if (c->ev_down_k) cudaEventDestroy(c->ev_down_k);
    cudaFree(c->A.tex); cudaFree(c->A.ipair); cudaFree(c->A.texw); cudaFree(c->A.cand); cudaFree(c->A.cand_count);
    cudaFree(c->A.plane_irregular); cudaFree(c->A.skip); cudaFree(c->A.blk);
    cudaFree(c->A.rs); cudaFree(c->A.chk); cudaFree(c->A.pts); cudaFree(c->A.dpl); cudaFree(c->A.spl); cudaFree(c->A.rs2);
    for (auto& s : c->up) { cudaFree(s.im); cudaFree(s.grad); cudaFree(s.theta); cudaFree(s.edge); }
    for (auto& s : c->down) cudaFree(s.planes);
    for (auto& s : c->ist) {
        if (s.host) cudaFreeHost(s.host);
        cudaFree(s.d_items);
        cudaFree(s.d_aux);
    }
    if (c->scat) {
        {
            std::lock_guard<std::mutex> lk(c->scat->mu);
            c->scat->stop = true;
        }
        c->scat->cv_job.notify_all();
        for (auto& t : c->scat->workers) t.join();
        for (auto& st : c->scat->st) {
            cudaFree(st.dev);
            if (st.host) cudaFreeHost(st.host);
            if (st.ev) cudaEventDestroy(st.ev);
        }
        delete c->scat;
    }
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->h_cand) cudaFreeHost(c->h_cand);
    cudaFree(c->d_chunk_off); cudaFree(c->d_counter); cudaFree(c->d_stats);
    cudaFree(c->tmp_rs); cudaFree(c->xfer); cudaFree(c->dbg); cudaFree(c->exp_buf); cudaFree(c->exp_pts); cudaFree(c->lf_buf);
    if (c->lf_host) cudaFreeHost(c->lf_host);
    cudaFree(c->ed_dev);
    cudaFree(c->edr_dev);
    cudaFree(c->edm_dev);
    if (c->edr_result_host) cudaFreeHost(c->edr_result_host);
    for (auto& e : c->edr_ev) if (e) cudaEventDestroy(e);
    if (c->ed_host) cudaFreeHost(c->ed_host);
    if (c->s_ed) cudaStreamDestroy(c->s_ed);
    for (auto& e : c->ed_ev) cudaEventDestroy(e);
    for (auto& e : c->lf_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {c->ev_p1[0], c->ev_p1[1], c->ev_p2[0], c->ev_p2[1], c->ev_p1_scan})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->marks)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_pack) cudaEventDestroy(e);
    c->r_copy.destroy(); c->r_compute.destroy(); c->r_down.destroy();
    for (cudaStream_t s : {c->s_compute, c->s_copy, c->s_down})
        if (s) cudaStreamDestroy(s);
    delete c;
}

static int create_impl(sdm_ctx* c)
{
    const sdm_config& cfg = c->cfg;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SDM_ERR_CUDA, "no CUDA device: libsdm_b200 has no CPU fallback");
    if (cfg.device < 0 || cfg.device >= ndev) return fail(SDM_ERR_ARG, "device %d out of range (%d devices)", cfg.device, ndev);
    CU(cudaSetDevice(cfg.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg.device));
    if (prop.major < 10) return fail(SDM_ERR_CUDA, "device sm_%d%d: this library is built for sm_100a only", prop.major, prop.minor);

    const size_t P = (size_t)cfg.width * cfg.height;
    const size_t n = (size_t)cfg.max_keyframes;
    c->npix = P;
    if (const char* e = getenv("SDM_SCAN")) c->scan_warp_per_pixel = strcmp(e, "warp") == 0 ? 1 : (strcmp(e, "warp_tma") == 0 ? 2 : 0);
    if (const char* e = getenv("SDM_TRACE")) c->trace = (e[0] == '1');
    c->kf.assign(n, KfState());
    sdm::DevParams& D = c->P;
    D.W = cfg.width;
    D.H = cfg.height;
    D.lambdaG = (float)cfg.lambdaG;
    D.lambdaL = (float)cfg.lambdaL;
    D.lambdaTheta = (float)cfg.lambdaTheta;
    D.lambdaN = cfg.lambdaN;
    D.theta = cfg.theta;
    D.inv_theta = 1 / cfg.theta;
    D.var_num = 2 * cfg.sigmaI * cfg.sigmaI;
    D.chi_fusion_lt = sdm::thr_lt(cfg.chi2_fusion);
    D.chi_inter_lt = sdm::thr_lt(cfg.chi2_inter);
    D.chi_inter_lo = D.chi_inter_lt * (1.0f - 0x1p-16f);
    D.chi_inter_hi = D.chi_inter_lt * (1.0f + 0x1p-16f);
    D.eps_gt = sdm::thr_gt(cfg.eps);
    D.eps_lt = sdm::thr_lt(cfg.eps);
    D.slope_max = cfg.slope_max;

    CU(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    {   // the download stream also runs the small block-sparse copy kernel: highest priority, so that its blocks are placed
        // as soon as a persistent pass kernel's block retires instead of queueing behind the next pass
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->s_down, cudaStreamNonBlocking, hi));
        CU(cudaStreamCreateWithPriority(&c->s_down_k, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&c->ev_down_k, cudaEventDisableTiming));
    }
    CU(cudaStreamCreateWithFlags(&c->s_halo, cudaStreamNonBlocking));
    CU(cudaMalloc(&c->xflags, sizeof(sdm::XFlags)));
    CU(cudaMemset(c->xflags, 0, sizeof(sdm::XFlags)));
    c->is_halo.assign(n, 0);
    RC(c->r_copy.init(1024));
    RC(c->r_compute.init(1024));
    RC(c->r_down.init(1024));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreate(&c->ev_p1[i]));
        CU(cudaEventCreate(&c->ev_p2[i]));
    }
    CU(cudaEventCreate(&c->ev_p1_scan));
    for (auto& e : c->marks) CU(cudaEventCreate(&e));
    sdm::DevArena& A = c->A;
    memset(&A, 0, sizeof(A));
    A.P = P;
    CU(cudaMalloc(&A.tex, n * P * sizeof(float4)));
    CU(cudaMalloc(&A.ipair, n * P * sizeof(uchar2)));
    CU(cudaMalloc(&A.cand, n * P * sizeof(uint32_t)));
    CU(cudaMalloc(&A.cand_count, n * sizeof(int)));
    CU(cudaMalloc(&A.plane_irregular, n * sizeof(int)));
    CU(cudaMemsetAsync(A.plane_irregular, 0, n * sizeof(int), c->s_compute));
    A.blk_words = ((cfg.width + sdm::kBlkPx - 1) / sdm::kBlkPx + 31) / 32;
    CU(cudaMalloc(&A.blk, n * (size_t)cfg.height * A.blk_words * sizeof(uint32_t)));
    CU(cudaMemsetAsync(A.blk, 0, n * (size_t)cfg.height * A.blk_words * sizeof(uint32_t), c->s_compute));
    CU(cudaMalloc(&A.rs, n * P * sizeof(float2)));
    CU(cudaMalloc(&A.chk, n * P * sizeof(float)));
    CU(cudaMalloc(&A.pts, n * P * 3 * sizeof(float)));
    CU(cudaMalloc(&A.dpl, n * P * sizeof(float)));
    CU(cudaMalloc(&A.spl, n * P * sizeof(float)));
    CU(cudaMemsetAsync(A.dpl, 0, n * P * sizeof(float), c->s_compute));
    CU(cudaMemsetAsync(A.spl, 0, n * P * sizeof(float), c->s_compute));
    if (cfg.intra_check || cfg.intra_grow) {  // second plane of the intra ping-pong (candidate-list stencils)
        CU(cudaMalloc(&A.rs2, n * P * sizeof(float2)));
        CU(cudaMemsetAsync(A.rs2, 0, n * P * sizeof(float2), c->s_compute));
    }
    CU(cudaMemsetAsync(A.cand_count, 0, n * sizeof(int), c->s_compute));
    CU(cudaMemsetAsync(A.rs, 0, n * P * sizeof(float2), c->s_compute));
    CU(cudaMemsetAsync(A.chk, 0, n * P * sizeof(float), c->s_compute));
    CU(cudaMemsetAsync(A.pts, 0, n * P * 3 * sizeof(float), c->s_compute));
    for (auto& s : c->up) {
        CU(cudaMalloc(&s.im, P));
        CU(cudaMalloc(&s.grad, P * sizeof(float)));
        CU(cudaMalloc(&s.theta, P * sizeof(float)));
        CU(cudaMalloc(&s.edge, P * sizeof(int32_t)));
    }
    for (auto& s : c->down) CU(cudaMalloc(&s.planes, 2 * P * sizeof(float)));
    CU(cudaMallocHost(&c->h_count, sizeof(int)));
    CU(cudaMallocHost(&c->h_cand, n * sizeof(int)));
    memset(c->h_cand, 0, n * sizeof(int));
    CU(cudaMalloc(&c->d_counter, sizeof(int)));
    CU(cudaMalloc(&c->d_stats, sizeof(sdm::DevStats)));
    CU(cudaMemsetAsync(c->d_stats, 0, sizeof(sdm::DevStats), c->s_compute));
    CU(cudaMalloc(&c->xfer, 2 * P * sizeof(float)));
    // persistent grids: every SM filled to the kernels' occupancy
    int occ = 0;
    c->n_sm = prop.multiProcessorCount;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sdm::k_pass1, sdm::kPass1Warps * 32, 0));
    c->grid_pass1_warp = std::max(1, occ) * prop.multiProcessorCount;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sdm::k_pass2_cand, sdm::kLaneBlock, 0));
    c->grid_pass2 = std::max(1, occ) * prop.multiProcessorCount;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sdm::k_intra_cand, sdm::kChunk, 0));
    c->grid_intra = std::max(1, occ) * prop.multiProcessorCount;
    if (c->scan_warp_per_pixel == 2) {
        // TMA descriptor of the texel arena; cuTensorMapEncodeTiled comes from the driver through the runtime (no -lcuda)
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(SDM_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        const cuuint64_t dims[3] = {4, (cuuint64_t)cfg.width, (cuuint64_t)n * cfg.height};
        const cuuint64_t strides[2] = {16, (cuuint64_t)16 * cfg.width};  // bytes, dims 1 and 2
        const cuuint32_t box[3] = {4, (cuuint32_t)sdm::kTmaBW, (cuuint32_t)sdm::kTmaBH};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = ((EncodeFn)fn)(&c->tex_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, A.tex, dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SDM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        const int smem = sdm::kTmaWarps * sdm::kTmaTileBytes;
        CU(cudaFuncSetAttribute(sdm::k_pass1_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sdm::k_pass1_tma, sdm::kTmaWarps * 32, smem));
        c->grid_pass1_tma = std::max(1, occ) * prop.multiProcessorCount;
    }
    // second-generation scan loop (scan_pixel_lane<2>): the reference's constants only, and only if its reciprocal
    // form of x / 0.23f is IEEE-exact over the whole range the loop can produce (checked here, ~2 ms)
    D.scan2 = 0;
    const char* scan_env = getenv("SDM_SCAN");
    if (cfg.lambdaG == 8 && cfg.lambdaL == 80 && cfg.lambdaTheta == 45 && cfg.theta == sdm::kTheta2 &&
        !(scan_env && strcmp(scan_env, "lane1") == 0)) {
        unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(&c->d_stats->checked);
        sdm::k_verify_div<<<c->n_sm * 8, 256, 0, c->s_compute>>>(d_bad);
        CU(cudaGetLastError());
        unsigned long long bad = 1;
        CU(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, c->s_compute));
        CU(cudaStreamSynchronize(c->s_compute));
        CU(cudaMemsetAsync(d_bad, 0, sizeof(bad), c->s_compute));
        D.scan2 = (bad == 0) ? 2 : 0;
        // third generation: second-generation arithmetic on the columns the skip planes (k_skip) leave.  Its row-exit
        // estimate is sized for images up to 8192 x 8192 (scan_columns3)
        const bool want3 = SDM_DEFAULT_SCAN_GEN == 3 ? !(scan_env && strcmp(scan_env, "lane2") == 0)
                                                     : (scan_env && strcmp(scan_env, "lane3") == 0);
        if (D.scan2 && want3 && cfg.width <= 8192 && cfg.height <= 8192) {
            CU(cudaMalloc(&A.skip, n * P * sdm::kSkipBins));
            D.scan2 = 3;
        }
    }
    // the wrap-encoded texel plane of the second / third generation (irregular keyframes run the first generation on tex)
    if (D.scan2) CU(cudaMalloc(&A.texw, n * P * sizeof(float4)));
    CU(cudaStreamSynchronize(c->s_compute));
    if (c->trace) {
        CU(cudaEventCreate(&c->trace_base));
        CU(cudaEventRecord(c->trace_base, c->s_compute));
    }
    return SDM_OK;
}

int sdm_create(const sdm_config* cfg, sdm_ctx** out)
{
    if (!cfg || !out) return fail(SDM_ERR_ARG, "sdm_create: null argument");
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->width > 65535 || cfg->height > 65535)
        return fail(SDM_ERR_ARG, "image size %dx%d unsupported", cfg->width, cfg->height);
    if (cfg->max_keyframes < 1) return fail(SDM_ERR_ARG, "max_keyframes must be >= 1");
    if (!(cfg->theta > 0.f)) return fail(SDM_ERR_ARG, "theta must be > 0");
    sdm_ctx* c = new (std::nothrow) sdm_ctx();
    if (!c) return fail(SDM_ERR_NOMEM, "out of host memory");
    c->cfg = *cfg;
    int rc = create_impl(c);
    if (rc != SDM_OK) {
        std::string keep = g_last_error;
        sdm_destroy(c);
        g_last_error = keep;
        return rc;
    }
    *out = c;
    return SDM_OK;
}

int sdm_synchronize(sdm_ctx* c)
{
    if (!c) return fail(SDM_ERR_ARG, "null context");
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaStreamSynchronize(c->s_copy));
    CU(cudaStreamSynchronize(c->s_compute));
    CU(cudaStreamSynchronize(c->s_down));
    CU(cudaStreamSynchronize(c->s_halo));
    if (c->xstep) {  // a device-side wait of sdm_exchange / sdm_pass1 that gave up
        unsigned err = 0;
        CU(cudaMemcpy(&err, &c->xflags->err, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            CU(cudaMemset(&c->xflags->err, 0, sizeof(err)));
            return fail(SDM_ERR_STATE, "sdm_exchange: device-side wait for %s timed out (a peer did not reach the same step)",
                        err == 1 ? "an owner's pass-1 planes" : "the pullers' acknowledgements");
        }
    }
    if (c->scat) {  // host-side scatters of sdm_scatter_keyframes
        std::unique_lock<std::mutex> lk(c->scat->mu);
        c->scat->cv_idle.wait(lk, [&] { return c->scat->pending == 0; });
        if (!c->scat->err.empty()) {
            const std::string e = c->scat->err;
            c->scat->err.clear();
            return fail(SDM_ERR_CUDA, "scatter worker: %s", e.c_str());
        }
    }
    trace_dump(c);
    return SDM_OK;
}

int sdm_scan_generation(sdm_ctx* c) { return !c || c->scan_warp_per_pixel ? 0 : (c->P.scan2 ? c->P.scan2 : 1); }
int sdm_last_scan_long(sdm_ctx* c) { return c && c->last_scan_long ? 1 : 0; }

int sdm_get_stats(sdm_ctx* c, sdm_stats* out)
{
    if (!c || !out) return fail(SDM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    sdm::DevStats h;
    CU(cudaMemcpyAsync(&h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    out->candidates = (long long)h.candidates;
    out->fused = (long long)h.fused;
    out->checked = (long long)h.checked;
    return SDM_OK;
}

int sdm_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return fail(SDM_ERR_ARG, "null argument");
    CU(cudaMallocHost(ptr, bytes));
    return SDM_OK;
}
int sdm_host_free(void* ptr)
{
    CU(cudaFreeHost(ptr));
    return SDM_OK;
}

// ---- keyframe planes -----------------------------------------------------------------------------
int sdm_upload_keyframes(sdm_ctx* c, int n, const sdm_upload_desc* d)
{
    if (!c || (n > 0 && !d)) return fail(SDM_ERR_ARG, "sdm_upload_keyframes: null argument");
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t row = (size_t)W * 4;
    for (int i = 0; i < n; ++i) {
        if (!d[i].im || (d[i].grad == nullptr) != (d[i].theta == nullptr))
            return fail(SDM_ERR_ARG, "sdm_upload_keyframes: entry %d needs im and either both or neither of grad / theta", i);
        if (!slot_ok(c, d[i].kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range [0,%d)", d[i].kf, (int)c->kf.size());
        if (d[i].im_step < (size_t)W || (d[i].grad && (d[i].grad_step < row || d[i].theta_step < row)) ||
            (d[i].edge && d[i].edge_step < row))
            return fail(SDM_ERR_ARG, "row step smaller than a row");
    }
    CU(cudaSetDevice(c->cfg.device));
    c->n_pack_ev = 0;
    for (int i0 = 0; i0 < n; i0 += kUpStages / 2) {
        const int m = std::min(kUpStages / 2, n - i0);
        const int first_stage = c->up_next;
        // H2D on the copy stream as soon as the staging sets are free (the k_packs that last read them have run) ...
        uint64_t busy = 0, need_down = 0;
        for (int i = 0; i < m; ++i) busy = std::max(busy, c->up[(first_stage + i) % kUpStages].busy);
        RC(c->r_compute.wait(c->s_copy, busy));
        trace_begin(c, c->s_copy, "h2d", m);
        for (int i = 0; i < m; ++i) {
            const sdm_upload_desc& u = d[i0 + i];
            UpStage& st = c->up[(first_stage + i) % kUpStages];
            RC(copy2d(st.im, W, u.im, u.im_step, W, H, cudaMemcpyHostToDevice, c->s_copy));
            if (u.grad) {
                RC(copy2d(st.grad, row, u.grad, u.grad_step, row, H, cudaMemcpyHostToDevice, c->s_copy));
                RC(copy2d(st.theta, row, u.theta, u.theta_step, row, H, cudaMemcpyHostToDevice, c->s_copy));
            }
            if (u.edge) RC(copy2d(st.edge, row, u.edge, u.edge_step, row, H, cudaMemcpyDefault, c->s_copy));  // host or device plane
            need_down = std::max(need_down, std::max(c->kf[u.kf].down_ds_id, c->kf[u.kf].down_cp_id));
        }
        trace_end(c, c->s_copy);
        uint64_t h2d = 0;
        RC(c->r_copy.record(c->s_copy, &h2d));
        // ... packing + candidate compaction on the compute stream, behind every pass already queued on the slots
        RC(c->r_copy.wait(c->s_compute, h2d));
        RC(c->r_down.wait(c->s_compute, need_down));
        trace_begin(c, c->s_compute, "pack", m);
        while ((int)c->ev_pack.size() < c->n_pack_ev + 2) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            c->ev_pack.push_back(e);
        }
        CU(cudaEventRecord(c->ev_pack[c->n_pack_ev], c->s_compute));
        {   // the whole batch with one launch per kernel (grid.z = keyframe)
            static_assert(kUpStages / 2 <= sdm::kPackBatch, "pack batch");
            sdm::PackBatch B;
            memset(&B, 0, sizeof(B));
            bool any_planes = false, any_image = false;
            for (int i = 0; i < m; ++i) {
                const sdm_upload_desc& u = d[i0 + i];
                UpStage& st = c->up[(first_stage + i) % kUpStages];
                B.slot[i] = u.kf;
                B.im[i] = st.im;
                B.grad[i] = u.grad ? st.grad : nullptr;
                B.theta[i] = u.grad ? st.theta : nullptr;
                B.edge[i] = u.edge ? st.edge : nullptr;
                (u.grad ? any_planes : any_image) = true;
            }
            sdm::k_pack_reset<<<m, 256, 0, c->s_compute>>>(c->A, c->P, B);
            if (any_planes)
                sdm::k_pack<<<dim3((W + sdm::kTileW - 1) / sdm::kTileW, (H + sdm::kTileH - 1) / sdm::kTileH, m), dim3(32, 8), 0,
                              c->s_compute>>>(c->A, c->P, B);
            if (any_image)  // GradImg / GradTheta produced on the device from im_ (KeyFrame.cc:69-74)
                sdm::k_pack_image<<<tile_grid(c, m), dim3(32, 8), 0, c->s_compute>>>(c->A, c->P, B);
            if (c->A.skip)  // skip distances of the third-generation scan loop, from the texels just packed
                sdm::k_skip<<<dim3(1, H, m), sdm::kSkipSpan, 0, c->s_compute>>>(c->A, c->P, B);
            c->launches += 1 + (any_planes ? 1 : 0) + (any_image ? 1 : 0) + (c->A.skip ? 1 : 0);
        }
        static_assert(kUpStages / 2 <= sdm::kSlotList, "slot list of k_publish_counts");
        sdm::SlotList sl;
        for (int i = 0; i < m; ++i) sl.s[i] = d[i0 + i].kf;
        sdm::k_publish_counts<<<1, 32, 0, c->s_compute>>>(c->A.cand_count, sl, m, c->h_cand);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->ev_pack[c->n_pack_ev + 1], c->s_compute));
        c->n_pack_ev += 2;
        trace_end(c, c->s_compute);
        c->launches += 1;
        uint64_t id = 0;
        RC(c->r_compute.record(c->s_compute, &id));
        for (int i = 0; i < m; ++i) {
            const sdm_upload_desc& u = d[i0 + i];
            c->up[(first_stage + i) % kUpStages].busy = id;
            KfState& k = c->kf[u.kf];
            k.comp_id = id;
            k.count_id = id;
            k.uploaded = true;
            k.pass1_done = false;
            k.rs_dense = false;
            k.split_stale = false;
            memcpy(k.K, u.K, sizeof(k.K));
            memcpy(k.Tcw, u.Tcw, sizeof(k.Tcw));
        }
        c->up_next = (first_stage + m) % kUpStages;
    }
    return SDM_OK;
}

int sdm_upload_keyframe(sdm_ctx* c, int kf, const uint8_t* im, size_t im_step, const float* grad, size_t grad_step,
                        const float* theta, size_t theta_step, const int32_t* edge, size_t edge_step, const float K[4],
                        const float Tcw[12])
{
    if (!c || !im || !K || !Tcw) return fail(SDM_ERR_ARG, "sdm_upload_keyframe: null argument");
    sdm_upload_desc u;
    u.kf = kf;
    u.im = im; u.im_step = im_step;
    u.grad = grad; u.grad_step = grad_step;
    u.theta = theta; u.theta_step = theta_step;
    u.edge = edge; u.edge_step = edge_step;
    memcpy(u.K, K, sizeof(u.K));
    memcpy(u.Tcw, Tcw, sizeof(u.Tcw));
    return sdm_upload_keyframes(c, 1, &u);
}

int sdm_set_pose(sdm_ctx* c, int kf, const float Tcw[12])
{
    if (!c || !Tcw) return fail(SDM_ERR_ARG, "null argument");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    memcpy(c->kf[kf].Tcw, Tcw, sizeof(float) * 12);
    return SDM_OK;
}

int sdm_set_intrinsics(sdm_ctx* c, int kf, const float K[4])
{
    if (!c || !K) return fail(SDM_ERR_ARG, "null argument");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    memcpy(c->kf[kf].K, K, sizeof(float) * 4);
    return SDM_OK;
}

int sdm_candidate_count(sdm_ctx* c, int kf, int* count)
{
    if (!c || !count) return fail(SDM_ERR_ARG, "null argument");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    if (!c->kf[kf].uploaded) return fail(SDM_ERR_STATE, "keyframe slot %d not uploaded", kf);
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaMemcpyAsync(c->h_count, c->A.cand_count + kf, sizeof(int), cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    *count = *c->h_count;
    return SDM_OK;
}

int sdm_candidate_blocks(sdm_ctx* c, int n, const int32_t* kfs, uint64_t* blocks)
{
    if (!c || (n > 0 && !kfs) || !blocks) return fail(SDM_ERR_ARG, "null argument");
    *blocks = 0;
    CU(cudaSetDevice(c->cfg.device));
    const size_t words = (size_t)c->cfg.height * c->A.blk_words;
    std::vector<uint32_t> h(words);
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(c, kfs[i])) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kfs[i]);
        if (!c->kf[kfs[i]].uploaded) return fail(SDM_ERR_STATE, "keyframe slot %d not uploaded", kfs[i]);
        CU(cudaMemcpyAsync(h.data(), c->A.blk + (size_t)kfs[i] * words, words * 4, cudaMemcpyDeviceToHost, c->s_compute));
        CU(cudaStreamSynchronize(c->s_compute));
        for (size_t w = 0; w < words; ++w) *blocks += (uint64_t)__builtin_popcount(h[w]);
    }
    return SDM_OK;
}

// ---- the two hot loops ---------------------------------------------------------------------------
int sdm_pass1(sdm_ctx* c, int n, const sdm_item* items)
{
    if (!c || (n > 0 && !items)) return fail(SDM_ERR_ARG, "sdm_pass1: null argument");
    if (n <= 0) return SDM_OK;
    if (n > 65535) return fail(SDM_ERR_ARG, "pass-1 batch limited to 65535 keyframes per call");
    CU(cudaSetDevice(c->cfg.device));
    Batch b;
    RC(prepare_batch(c, n, items, false, false, &b));
    for (int i = 0; i < n; ++i) {
        KfState& k = c->kf[items[i].kf];
        if (k.rs_dense) {  // the slot's outputs were produced by the dense pass 2: drop stale non-candidate values
            const size_t P = c->npix, s = (size_t)items[i].kf;
            RC(c->r_down.wait(c->s_compute, k.down_cp_id));  // a download of chk / pts may still be in flight (prepare_batch waited for rho / sigma only)
            CU(cudaMemsetAsync(c->A.rs + s * P, 0, P * sizeof(float2), c->s_compute));
            CU(cudaMemsetAsync(c->A.dpl + s * P, 0, P * sizeof(float), c->s_compute));
            CU(cudaMemsetAsync(c->A.spl + s * P, 0, P * sizeof(float), c->s_compute));
            CU(cudaMemsetAsync(c->A.chk + s * P, 0, P * sizeof(float), c->s_compute));
            CU(cudaMemsetAsync(c->A.pts + s * P * 3, 0, P * 3 * sizeof(float), c->s_compute));
            k.rs_dense = false;
            k.split_stale = false;
        }
    }
    if (c->xstep && c->acks_waited != c->xstep) {
        // the planes about to be overwritten may still be read by peers pulling the previous step (sdm_exchange)
        sdm::k_xwait_acks<<<1, 32, 0, c->s_compute>>>(c->xflags, c->xstep);
        c->acks_waited = c->xstep;
        c->launches++;
    }
    CU(cudaMemsetAsync(&c->d_stats->fused, 0, sizeof(unsigned long long), c->s_compute));
    const sdm::DevPlan plan = make_plan(c, nullptr, n);
    sdm::k_plan<<<1, 1024, 0, c->s_compute>>>(plan, c->d_items, c->A.cand_count, c->d_stats);
    CU(cudaEventRecord(c->ev_p1[0], c->s_compute));
    trace_begin(c, c->s_compute, "pass1", n);
    if (c->scan_warp_per_pixel == 2)
        sdm::k_pass1_tma<<<c->grid_pass1_tma, sdm::kTmaWarps * 32, sdm::kTmaWarps * sdm::kTmaTileBytes, c->s_compute>>>(
            c->A, c->P, c->d_items, plan, c->d_stats, c->tex_map);
    else if (c->scan_warp_per_pixel)
        sdm::k_pass1<<<c->grid_pass1_warp, sdm::kPass1Warps * 32, 0, c->s_compute>>>(c->A, c->P, c->d_items, plan, c->d_stats);
    else {
        int max_n = 1;
        for (int i = 0; i < n; ++i) max_n = std::max(max_n, (int)items[i].n_nbr);
        // per warp: the work order with max_n neighbour records (16-byte aligned); then the hypotheses [max_n][block]
        const int item_bytes = (int)((offsetof(sdm::DevItem, pair) + (size_t)max_n * sizeof(sdm::DevPair) + 15) & ~(size_t)15);
        const size_t smem = (size_t)(sdm::kLaneBlock / 32) * item_bytes + (size_t)max_n * sdm::kLaneBlock * sizeof(float2);
        int occ = 0;  // persistent grid: fill every SM to this launch's occupancy
        // the exact short forms of the two direction gates exist for the reference's thresholds only
        const bool fast = (c->cfg.lambdaL == 80 && c->cfg.lambdaTheta == 45);
        // long scans: the 64-register instantiation (k_pass1_lane's kLong).  Mean search range of the batch's pairs, taken at
        // the principal point: u(d) = fx * X / Z + cx of R21 * (0, 0, 1) * d + t21 at the two depth bounds (GetSearchRange, :1598-1631)
        bool long_scans = false;
        if (fast && c->P.scan2 == 3) {
            const char* env = getenv("SDM_SCAN_LONG");  // developer knob: 0 / 1 forces the choice
            if (env && (env[0] == '0' || env[0] == '1')) {
                long_scans = env[0] == '1';
            } else {
                const sdm::DevItem* h_items = (const sdm::DevItem*)b.stage->host;
                double sum = 0;
                long cnt = 0;
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < h_items[i].n_nbr; ++j) {
                        const sdm::DevPair& p = h_items[i].pair[j];
                        const float z0 = p.R[8] * h_items[i].min_depth + p.t[2], z1 = p.R[8] * h_items[i].max_depth + p.t[2];
                        if (!(z0 > 0.f) || !(z1 > 0.f)) continue;
                        const float u0 = h_items[i].K[0] * (p.R[2] * h_items[i].min_depth + p.t[0]) / z0;
                        const float u1 = h_items[i].K[0] * (p.R[2] * h_items[i].max_depth + p.t[0]) / z1;
                        sum += std::min((double)fabsf(u1 - u0), (double)c->cfg.width);
                        ++cnt;
                    }
                long_scans = cnt > 0 && sum / (double)cnt > (double)kLongScanColumns;
            }
        }
        c->last_scan_long = long_scans;
        auto kern = !fast ? sdm::k_pass1_lane<false, 2>
                          : (c->P.scan2 == 3 ? (long_scans ? sdm::k_pass1_lane<true, 3, true> : sdm::k_pass1_lane<true, 3, false>)
                                             : sdm::k_pass1_lane<true, 2>);
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, sdm::kLaneBlock, smem));
        kern<<<std::max(1, occ) * c->n_sm, sdm::kLaneBlock, smem, c->s_compute>>>(c->A, c->P, c->d_items, plan, c->d_stats, item_bytes);
    }
    CU(cudaGetLastError());
    c->launches += 2;
    CU(cudaEventRecord(c->ev_p1_scan, c->s_compute));
    if (c->cfg.intra_check || c->cfg.intra_grow) {
        // candidate-list stencils over the same work plan: check rs -> rs2, grow rs2 -> rs (+ dense copies)
        for (int stage = 0; stage < 2; ++stage) {
            CU(cudaMemsetAsync(c->d_counter, 0, sizeof(int), c->s_compute));
            const int copy_only = stage == 0 ? !c->cfg.intra_check : !c->cfg.intra_grow;
            sdm::k_intra_cand<<<c->grid_intra, sdm::kChunk, 0, c->s_compute>>>(c->A, c->P, c->d_items, plan, stage, copy_only);
        }
        CU(cudaGetLastError());
        c->launches += 2;
    }
    CU(cudaEventRecord(c->ev_p1[1], c->s_compute));
    trace_end(c, c->s_compute);
    c->p1_timed = true;
    RC(finish_batch(c, n, items, b));
    for (int i = 0; i < n; ++i) c->kf[items[i].kf].pass1_done = true;
    return SDM_OK;
}

// continuation = second half of a pass split around a halo pull (sdm_pass2): keeps the counters and the start event
static int pass2_impl(sdm_ctx* c, int n, const sdm_item* items, int points_only, bool continuation = false)
{
    CU(cudaSetDevice(c->cfg.device));
    Batch b;
    RC(prepare_batch(c, n, items, true, points_only != 0, &b));
    if (!points_only && !continuation) CU(cudaMemsetAsync(&c->d_stats->checked, 0, sizeof(unsigned long long), c->s_compute));
    if (!continuation) CU(cudaEventRecord(c->ev_p2[0], c->s_compute));
    if (b.n_sparse > 0) {
        const sdm::DevPlan plan = make_plan(c, b.d_order_sparse, b.n_sparse);
        sdm::k_plan<<<1, 1024, 0, c->s_compute>>>(plan, c->d_items, c->A.cand_count, nullptr);
        sdm::k_pass2_cand<<<c->grid_pass2, sdm::kLaneBlock, 0, c->s_compute>>>(c->A, c->P, c->d_items, plan, c->d_stats);
        c->launches += 2;
    }
    if (b.n_dense > 0) {
        dim3 grid((unsigned)((c->npix + 255) / 256), (unsigned)b.n_dense);
        sdm::k_pass2<<<grid, 256, 0, c->s_compute>>>(c->A, c->P, c->d_items, b.d_order_dense, points_only ? nullptr : c->d_stats,
                                                     points_only);
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_p2[1], c->s_compute));
    c->p2_timed = true;
    return finish_batch(c, n, items, b);
}

int sdm_pass2(sdm_ctx* c, int n, const sdm_item* items)
{
    if (!c || (n > 0 && !items)) return fail(SDM_ERR_ARG, "sdm_pass2: null argument");
    if (n <= 0) return SDM_OK;
    if (n > 65535) return fail(SDM_ERR_ARG, "pass-2 batch limited to 65535 keyframes per call");
    // keyframes that read a plane still being pulled from a peer (sdm_exchange) go second, behind the pull;
    // the others run at once, so the NVLink copies overlap pass 2 of the interior keyframes
    uint64_t need_pull = 0;
    std::vector<sdm_item> indep, dep;
    for (int i = 0; i < n; ++i) {
        bool d = false;
        if (items[i].n_nbr >= 0 && items[i].n_nbr <= SDM_MAX_NBR)
            for (int j = -1; j < items[i].n_nbr; ++j) {
                const int slot = j < 0 ? items[i].kf : items[i].nbr[j];
                if (!slot_ok(c, slot)) continue;  // reported by build_item
                const uint64_t id = c->kf[slot].pull_id;
                if (id && !c->r_compute.done(id)) { d = true; need_pull = std::max(need_pull, id); }
            }
        if (need_pull == 0) continue;  // nothing in flight so far: no copies needed yet
        if (indep.empty() && dep.empty()) indep.assign(items, items + i);
        (d ? dep : indep).push_back(items[i]);
    }
    if (need_pull == 0) return pass2_impl(c, n, items, 0);
    if (!indep.empty()) RC(pass2_impl(c, (int)indep.size(), indep.data(), 0));
    RC(c->r_compute.wait(c->s_compute, need_pull));
    return pass2_impl(c, (int)dep.size(), dep.data(), 0, !indep.empty());
}

int sdm_update_points(sdm_ctx* c, int n, const int32_t* kfs)
{
    if (!c || (n > 0 && !kfs)) return fail(SDM_ERR_ARG, "sdm_update_points: null argument");
    if (n <= 0) return SDM_OK;
    if (n > 65535) return fail(SDM_ERR_ARG, "batch limited to 65535 keyframes per call");
    std::vector<sdm_item> items((size_t)n);
    for (int i = 0; i < n; ++i) {
        memset(&items[i], 0, sizeof(sdm_item));
        items[i].kf = kfs[i];
    }
    return pass2_impl(c, n, items.data(), 1);
}

int sdm_download_keyframes(sdm_ctx* c, int n, const sdm_download_desc* d)
{
    if (!c || (n > 0 && !d)) return fail(SDM_ERR_ARG, "sdm_download_keyframes: null argument");
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t P = c->npix, row = (size_t)W * 4;
    uint64_t need = 0;
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(c, d[i].kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", d[i].kf);
        if ((d[i].depth && d[i].depth_step < row) || (d[i].sigma && d[i].sigma_step < row) ||
            (d[i].checked && d[i].checked_step < row) || (d[i].points && d[i].points_step < 3 * row))
            return fail(SDM_ERR_ARG, "row step smaller than a row");
        need = std::max(need, c->kf[d[i].kf].comp_id);
    }
    if (n <= 0) return SDM_OK;
    CU(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->s_down;
    RC(c->r_compute.wait(s, need));
    for (int i = 0; i < n; ++i) {
        const sdm_download_desc& q = d[i];
        const KfState& k = c->kf[q.kf];
        const size_t off = (size_t)q.kf * P;
        if ((q.depth || q.sigma) && !k.split_stale) {
            if (q.depth) RC(copy2d(q.depth, q.depth_step, c->A.dpl + off, row, row, H, cudaMemcpyDeviceToHost, s));
            if (q.sigma) RC(copy2d(q.sigma, q.sigma_step, c->A.spl + off, row, row, H, cudaMemcpyDeviceToHost, s));
        } else if (q.depth || q.sigma) {  // (rho,sigma) written from outside: de-interleave on the fly
            DownStage& st = c->down[c->down_next];
            c->down_next = (c->down_next + 1) % kDownStages;
            // same stream: the previous user's D2H copies are ordered before this kernel
            sdm::k_split_rs<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(c->A.rs + off, st.planes, st.planes + P, P);
            CU(cudaGetLastError());
            c->launches++;
            if (q.depth) RC(copy2d(q.depth, q.depth_step, st.planes, row, row, H, cudaMemcpyDeviceToHost, s));
            if (q.sigma) RC(copy2d(q.sigma, q.sigma_step, st.planes + P, row, row, H, cudaMemcpyDeviceToHost, s));
        }
        if (q.checked) RC(copy2d(q.checked, q.checked_step, c->A.chk + off, row, row, H, cudaMemcpyDeviceToHost, s));
        if (q.points) RC(copy2d(q.points, q.points_step, c->A.pts + off * 3, 3 * row, 3 * row, H, cudaMemcpyDeviceToHost, s));
    }
    uint64_t id = 0;
    RC(c->r_down.record(s, &id));
    for (int i = 0; i < n; ++i) {
        if (d[i].depth || d[i].sigma) c->kf[d[i].kf].down_ds_id = id;
        if (d[i].checked || d[i].points) c->kf[d[i].kf].down_cp_id = id;
    }
    return SDM_OK;
}

// device-visible address of a host pointer if it lies in pinned memory (cudaMallocHost / cudaHostRegister), else nullptr
static float* pinned_dev_ptr(float* host)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return (float*)at.devicePointer;
}

// returns 1 when a destination plane is not pinned (nothing was enqueued), else SDM_OK / an error
static int sparse_to_pinned(sdm_ctx* c, int n, const sdm_download_desc* d, cudaStream_t s)
{
    std::vector<sdm::SparseDst> dst((size_t)n);
    uint64_t need = 0;
    for (int i = 0; i < n; ++i) {
        sdm::SparseDst& t = dst[i];
        t.slot = d[i].kf;
        t.depth = t.sigma = t.checked = t.points = nullptr;
        t.depth_step = d[i].depth_step; t.sigma_step = d[i].sigma_step;
        t.checked_step = d[i].checked_step; t.points_step = d[i].points_step;
        if (d[i].depth && !(t.depth = pinned_dev_ptr(d[i].depth))) return 1;
        if (d[i].sigma && !(t.sigma = pinned_dev_ptr(d[i].sigma))) return 1;
        if (d[i].checked && !(t.checked = pinned_dev_ptr(d[i].checked))) return 1;
        if (d[i].points && !(t.points = pinned_dev_ptr(d[i].points))) return 1;
        need = std::max(need, c->kf[d[i].kf].comp_id);
    }
    RC(c->r_compute.wait(s, need));
    const int H = c->cfg.height;
    for (int i0 = 0; i0 < n; i0 += sdm::kSparseBatch) {
        const int m = std::min(sdm::kSparseBatch, n - i0);
        sdm::SparseBatch B;
        memset(&B, 0, sizeof(B));
        for (int i = 0; i < m; ++i) B.d[i] = dst[i0 + i];
        sdm::k_sparse_rows<<<dim3(H, m), 32, 0, s>>>(c->A, c->P, B);
        c->launches++;
    }
    CU(cudaGetLastError());
    if (s != c->s_down) {  // the download ring lives on s_down: it continues behind this stream's kernels
        CU(cudaEventRecord(c->ev_down_k, s));
        CU(cudaStreamWaitEvent(c->s_down, c->ev_down_k, 0));
    }
    uint64_t id = 0;
    RC(c->r_down.record(c->s_down, &id));
    for (int i = 0; i < n; ++i) {
        if (d[i].depth || d[i].sigma) c->kf[d[i].kf].down_ds_id = id;
        if (d[i].checked || d[i].points) c->kf[d[i].kf].down_cp_id = id;
    }
    return SDM_OK;
}

static int scatter_init(sdm_ctx* c)
{
    Scatter* S = new (std::nothrow) Scatter();
    if (!S) return fail(SDM_ERR_NOMEM, "out of host memory");
    c->scat = S;
    const size_t bytes = 28 * c->npix;  // worst case: every pixel a candidate
    for (auto& st : S->st) {
        CU(cudaMalloc(&st.dev, bytes));
        CU(cudaMallocHost(&st.host, bytes));
        CU(cudaEventCreateWithFlags(&st.ev, cudaEventDisableTiming | cudaEventBlockingSync));
    }
    int nt = 6;
    if (const char* e = getenv("SDM_SCATTER_THREADS")) nt = std::max(1, std::min(64, atoi(e)));
    for (int i = 0; i < nt; ++i) S->workers.emplace_back(scatter_worker, S, c->cfg.device);
    return SDM_OK;
}

int sdm_scatter_keyframes(sdm_ctx* c, int n, const sdm_download_desc* d)
{
    if (!c || (n > 0 && !d)) return fail(SDM_ERR_ARG, "sdm_scatter_keyframes: null argument");
    const int W = c->cfg.width;
    const size_t row = (size_t)W * 4;
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(c, d[i].kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", d[i].kf);
        if ((d[i].depth && d[i].depth_step < row) || (d[i].sigma && d[i].sigma_step < row) ||
            (d[i].checked && d[i].checked_step < row) || (d[i].points && d[i].points_step < 3 * row))
            return fail(SDM_ERR_ARG, "row step smaller than a row");
        const KfState& k = c->kf[d[i].kf];
        if (!k.uploaded || !k.pass1_done) return fail(SDM_ERR_STATE, "keyframe slot %d has no results", d[i].kf);
        if (k.rs_dense || k.split_stale)
            return fail(SDM_ERR_STATE, "slot %d holds planes written from outside (not confined to the candidate pixels): "
                                        "use sdm_download_keyframes", d[i].kf);
    }
    if (n <= 0) return SDM_OK;
    CU(cudaSetDevice(c->cfg.device));
    {   // destination planes in pinned (device-visible) host memory: a kernel writes the candidate blocks of every row
        // straight into them over PCIe - no staging, no host threads
        int rc = sparse_to_pinned(c, n, d, c->s_down);
        if (rc != 1) return rc;  // 1 = some plane is pageable: candidate records + host-side scatter below
    }
    if (!c->scat) RC(scatter_init(c));
    Scatter* S = c->scat;
    cudaStream_t s = c->s_down;
    for (int i = 0; i < n; ++i) {
        const sdm_download_desc& q = d[i];
        KfState& k = c->kf[q.kf];
        RC(c->r_compute.host_sync(k.count_id));  // k_pack of the slot has run (it is far behind the caller's issue point)
        const int cnt = c->h_cand[q.kf];
        if (cnt <= 0) continue;
        int stage;
        {
            std::unique_lock<std::mutex> lk(S->mu);
            stage = S->next;
            S->cv_free.wait(lk, [&] { return S->st[stage].free_; });
            S->st[stage].free_ = false;
            S->next = (stage + 1) % kScatStages;
        }
        ScatStage& st = S->st[stage];
        // from here until the job is queued a failure must hand the stage back, or the ring blocks when it wraps
        auto release = [&](int rc) {
            {
                std::lock_guard<std::mutex> lk(S->mu);
                st.free_ = true;
            }
            S->cv_free.notify_all();
            return rc;
        };
        uint64_t id = 0;
        int rc = c->r_compute.wait(s, k.comp_id);
        if (rc == SDM_OK) {
            sdm::k_gather_sparse<<<(cnt + 255) / 256, 256, 0, s>>>(c->A, c->P, q.kf, cnt, st.dev);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(st.host, st.dev, (size_t)cnt * 28, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaEventRecord(st.ev, s);
            if (e != cudaSuccess) rc = fail(SDM_ERR_CUDA, "sdm_scatter_keyframes: %s", cudaGetErrorString(e));
        }
        if (rc == SDM_OK) rc = c->r_down.record(s, &id);
        if (rc != SDM_OK) return release(rc);
        c->launches++;
        k.down_ds_id = id;
        k.down_cp_id = id;
        {
            std::lock_guard<std::mutex> lk(S->mu);
            S->jobs.push_back(ScatJob{stage, cnt, W, q});
            ++S->pending;
        }
        S->cv_job.notify_one();
    }
    return SDM_OK;
}

int sdm_download_async(sdm_ctx* c, int kf, float* depth, size_t depth_step, float* sigma, size_t sigma_step, float* checked,
                       size_t checked_step, float* points, size_t points_step)
{
    sdm_download_desc q;
    q.kf = kf;
    q.depth = depth; q.depth_step = depth_step;
    q.sigma = sigma; q.sigma_step = sigma_step;
    q.checked = checked; q.checked_step = checked_step;
    q.points = points; q.points_step = points_step;
    return sdm_download_keyframes(c, 1, &q);
}

int sdm_download(sdm_ctx* c, int kf, float* depth, size_t depth_step, float* sigma, size_t sigma_step, float* checked,
                 size_t checked_step, float* points, size_t points_step)
{
    RC(sdm_download_async(c, kf, depth, depth_step, sigma, sigma_step, checked, checked_step, points, points_step));
    CU(cudaStreamSynchronize(c->s_down));
    return SDM_OK;
}

int sdm_download_planes(sdm_ctx* c, int kf, float* grad, size_t grad_step, float* theta, size_t theta_step)
{
    if (!c || !grad || !theta) return fail(SDM_ERR_ARG, "null argument");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    if (!c->kf[kf].uploaded) return fail(SDM_ERR_STATE, "keyframe slot %d not uploaded", kf);
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t P = c->npix, row = (size_t)W * 4;
    if (grad_step < row || theta_step < row) return fail(SDM_ERR_ARG, "row step smaller than a row");
    CU(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->s_down;
    RC(c->r_compute.wait(s, c->kf[kf].comp_id));
    DownStage& st = c->down[c->down_next];
    c->down_next = (c->down_next + 1) % kDownStages;
    sdm::k_split_tex<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(c->A.tex + (size_t)kf * P, st.planes, st.planes + P, P);
    CU(cudaGetLastError());
    c->launches++;
    RC(copy2d(grad, grad_step, st.planes, row, row, H, cudaMemcpyDeviceToHost, s));
    RC(copy2d(theta, theta_step, st.planes + P, row, row, H, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return SDM_OK;
}

int sdm_upload_depth(sdm_ctx* c, int kf, const float* depth, size_t depth_step, const float* sigma, size_t sigma_step)
{
    if (!c || !depth || !sigma) return fail(SDM_ERR_ARG, "null argument");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t P = c->npix, row = (size_t)W * 4;
    if (depth_step < row || sigma_step < row) return fail(SDM_ERR_ARG, "row step smaller than a row");
    CU(cudaSetDevice(c->cfg.device));
    KfState& k = c->kf[kf];
    cudaStream_t s = c->s_compute;
    RC(c->r_down.wait(s, k.down_ds_id));
    RC(copy2d(c->xfer, row, depth, depth_step, row, H, cudaMemcpyHostToDevice, s));
    RC(copy2d(c->xfer + P, row, sigma, sigma_step, row, H, cudaMemcpyHostToDevice, s));
    sdm::k_merge_rs<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(c->A.rs + (size_t)kf * P, c->xfer, c->xfer + P, P);
    CU(cudaGetLastError());
    c->launches++;
    RC(c->r_compute.record(s, &k.comp_id));
    CU(cudaStreamSynchronize(s));
    k.pass1_done = true;
    k.rs_dense = true;
    k.split_stale = true;
    return SDM_OK;
}

int sdm_upload_checked(sdm_ctx* c, int kf, const float* checked, size_t checked_step)
{
    if (!c || !checked) return fail(SDM_ERR_ARG, "null argument");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t P = c->npix, row = (size_t)W * 4;
    if (checked_step < row) return fail(SDM_ERR_ARG, "row step smaller than a row");
    CU(cudaSetDevice(c->cfg.device));
    KfState& k = c->kf[kf];
    cudaStream_t s = c->s_compute;
    RC(c->r_down.wait(s, k.down_cp_id));
    RC(copy2d(c->A.chk + (size_t)kf * P, row, checked, checked_step, row, H, cudaMemcpyHostToDevice, s));
    RC(c->r_compute.record(s, &k.comp_id));
    CU(cudaStreamSynchronize(s));
    k.rs_dense = true;  // the plane is not confined to the slot's candidate pixels: pass 2 / the point set visit every pixel
    return SDM_OK;
}

// ---- point-cloud export --------------------------------------------------------------------------
int sdm_export_points(sdm_ctx* c, int n, const int32_t* kfs, double sigma_max, sdm_point* out, size_t capacity,
                      uint64_t* counts, uint64_t* total)
{
    if (!c || (n > 0 && !kfs) || !total || (capacity > 0 && !out)) return fail(SDM_ERR_ARG, "sdm_export_points: null argument");
    *total = 0;
    if (n <= 0) return SDM_OK;
    uint64_t need = 0;
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(c, kfs[i])) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kfs[i]);
        if (!c->kf[kfs[i]].pass1_done) return fail(SDM_ERR_STATE, "slot %d has no depth planes", kfs[i]);
        need = std::max(need, c->kf[kfs[i]].comp_id);
    }
    CU(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->s_down;  // reads only; ordered after the last kernel on the slots
    RC(c->r_compute.wait(s, need));
    const int bpk = (int)((c->npix + sdm::kExportBlock - 1) / sdm::kExportBlock);
    const size_t nb = (size_t)n * bpk;
    if (nb > 0x7fffffffULL) return fail(SDM_ERR_ARG, "export batch too large");
    // scratch layout: slots[n] | counts[nb] | offsets[nb + 1] (u64) | kf_totals[n] (u64)
    const size_t off_counts = ((size_t)n * 4 + 7) & ~(size_t)7;
    const size_t off_offsets = (off_counts + nb * 4 + 7) & ~(size_t)7;
    const size_t off_totals = off_offsets + (nb + 1) * 8;
    const size_t bytes = off_totals + (size_t)n * 8;
    if (c->exp_cap < bytes) {
        CU(cudaStreamSynchronize(s));
        cudaFree(c->exp_buf);
        c->exp_buf = nullptr; c->exp_cap = 0;
        CU(cudaMalloc(&c->exp_buf, bytes));
        c->exp_cap = bytes;
    }
    char* base = (char*)c->exp_buf;
    int* d_slots = (int*)base;
    int* d_counts = (int*)(base + off_counts);
    unsigned long long* d_offsets = (unsigned long long*)(base + off_offsets);
    unsigned long long* d_totals = (unsigned long long*)(base + off_totals);
    const float sigma_gt = sdm::thr_gt(sigma_max);
    CU(cudaMemcpyAsync(d_slots, kfs, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    sdm::k_export_count<<<(unsigned)nb, sdm::kExportBlock, 0, s>>>(c->A, c->P, d_slots, sigma_gt, bpk, d_counts);
    sdm::k_export_scan<<<1, 1024, 0, s>>>(d_counts, (int)nb, bpk, d_offsets, d_totals);
    CU(cudaGetLastError());
    c->launches += 2;
    std::vector<unsigned long long> h_tot((size_t)n + 1);
    CU(cudaMemcpyAsync(h_tot.data(), d_totals, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&h_tot[n], d_offsets + nb, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *total = (uint64_t)h_tot[n];
    if (counts)
        for (int i = 0; i < n; ++i) counts[i] = (uint64_t)h_tot[i];
    const size_t m = (size_t)std::min<uint64_t>(*total, capacity);
    if (m > 0) {
        if (c->exp_pts_cap < m) {
            cudaFree(c->exp_pts);
            c->exp_pts = nullptr; c->exp_pts_cap = 0;
            CU(cudaMalloc(&c->exp_pts, m * sizeof(sdm_point)));
            c->exp_pts_cap = m;
        }
        sdm::k_export_scatter<<<(unsigned)nb, sdm::kExportBlock, 0, s>>>(c->A, c->P, d_slots, sigma_gt, bpk, d_offsets, c->exp_pts,
                                                                        (unsigned long long)m);
        CU(cudaGetLastError());
        c->launches++;
        CU(cudaMemcpyAsync(out, c->exp_pts, m * sizeof(sdm_point), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    uint64_t id = 0;
    RC(c->r_down.record(s, &id));
    for (int i = 0; i < n; ++i) c->kf[kfs[i]].down_ds_id = c->kf[kfs[i]].down_cp_id = id;
    return SDM_OK;
}

// ---- LineDetector::LineFitting over caller-supplied edge chains (LineDetector.cc:884-900) ----
int sdm_line_fit(sdm_ctx* c, int n, const sdm_edge_chains* sets, sdm_line3d* out, size_t capacity, uint64_t* counts, uint64_t* total)
{
    if (!c || (n > 0 && !sets) || !total || (capacity > 0 && !out)) return fail(SDM_ERR_ARG, "sdm_line_fit: null argument");
    *total = 0;
    if (n <= 0) return SDM_OK;
    if (c->cfg.width > 65535 || c->cfg.height > 65535) return fail(SDM_ERR_ARG, "sdm_line_fit: planes larger than 65535");
    const int min_len = 10;  // MIN_LINE_LENGTH (LineDetector.cc:20)
    size_t n_chains = 0, n_pix = 0, n_slots = 0;
    uint64_t need = 0;
    for (int i = 0; i < n; ++i) {
        const sdm_edge_chains& e = sets[i];
        if (!slot_ok(c, e.kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", e.kf);
        if (!c->kf[e.kf].pass1_done) return fail(SDM_ERR_STATE, "slot %d has no depth planes", e.kf);
        if (e.n_chains < 0 || (e.n_chains > 0 && (!e.offsets || !e.pixels))) return fail(SDM_ERR_ARG, "chain set %d: null list", i);
        for (int k = 0; k < e.n_chains; ++k) {
            const int len = e.offsets[k + 1] - e.offsets[k];
            if (len < 0) return fail(SDM_ERR_ARG, "chain set %d: offsets not ascending at %d", i, k);
            n_slots += (size_t)(len / min_len);
        }
        n_chains += (size_t)e.n_chains;
        n_pix += e.n_chains > 0 ? (size_t)e.offsets[e.n_chains] : 0;
        need = std::max(need, c->kf[e.kf].comp_id);
    }
    if (counts) for (int i = 0; i < n; ++i) counts[i] = 0;
    if (n_chains == 0) return SDM_OK;
    if (n_chains > 0x7fffffffULL || n_pix > 0x7fffffffULL || n_slots > 0x7fffffffULL) return fail(SDM_ERR_ARG, "line-fit batch too large");
    CU(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->s_down;  // reads the planes only; ordered after the last kernel on the slots
    RC(c->r_compute.wait(s, need));
    // host staging: kfs[n] | off[n_chains + 1] | kfi[n_chains] | slot0[n_chains] | pix[n_pix]
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_kfs = 0, o_off = al(o_kfs + (size_t)n * sizeof(sdm::LineFitKf)), o_kfi = al(o_off + (n_chains + 1) * 4),
                 o_slot0 = al(o_kfi + n_chains * 4), o_pix = al(o_slot0 + n_chains * 4), in_bytes = al(o_pix + n_pix * 4);
    // device only: n_out[n_chains] | offs[n_chains + 1] u64 | kf_totals[n] u64 | raw lines[n_slots] | compacted lines
    const size_t o_nout = in_bytes, o_offs = al(o_nout + n_chains * 4), o_tot = al(o_offs + (n_chains + 1) * 8),
                 o_raw = al(o_tot + (size_t)n * 8), o_cmp = al(o_raw + n_slots * sizeof(sdm::DevLine)),
                 bytes = o_cmp + n_slots * sizeof(sdm::DevLine);
    if (c->lf_host_cap < in_bytes) {
        CU(cudaStreamSynchronize(s));
        if (c->lf_host) cudaFreeHost(c->lf_host);
        c->lf_host = nullptr; c->lf_host_cap = 0;
        CU(cudaMallocHost(&c->lf_host, in_bytes));
        c->lf_host_cap = in_bytes;
    }
    if (c->lf_cap < bytes) {
        CU(cudaStreamSynchronize(s));
        cudaFree(c->lf_buf);
        c->lf_buf = nullptr; c->lf_cap = 0;
        CU(cudaMalloc(&c->lf_buf, bytes));
        c->lf_cap = bytes;
    }
    for (auto& e : c->lf_ev) if (!e) CU(cudaEventCreate(&e));
    char* h = (char*)c->lf_host;
    sdm::LineFitKf* h_kfs = (sdm::LineFitKf*)(h + o_kfs);
    int* h_off = (int*)(h + o_off);
    int* h_kfi = (int*)(h + o_kfi);
    int* h_slot0 = (int*)(h + o_slot0);
    uint32_t* h_pix = (uint32_t*)(h + o_pix);
    size_t ck = 0, cp = 0, cs = 0;
    for (int i = 0; i < n; ++i) {
        const sdm_edge_chains& e = sets[i];
        const KfState& k = c->kf[e.kf];
        h_kfs[i].slot = e.kf;
        h_kfs[i].chain0 = (int)ck;
        memcpy(h_kfs[i].K, k.K, sizeof(k.K));
        sdm::pose_inverse(k.Tcw, h_kfs[i].Twc);
        for (int j = 0; j < e.n_chains; ++j, ++ck) {
            const int len = e.offsets[j + 1] - e.offsets[j];
            h_off[ck] = (int)cp + (e.offsets[j] - e.offsets[0]);
            h_kfi[ck] = i;
            h_slot0[ck] = (int)cs;
            cs += (size_t)(len / min_len);
        }
        const size_t m = e.n_chains > 0 ? (size_t)(e.offsets[e.n_chains] - e.offsets[0]) : 0;
        for (size_t q = 0; q < m; ++q) {
            const uint32_t rc = e.pixels[e.offsets[0] + q];
            if ((int)(rc >> 16) >= c->cfg.height || (int)(rc & 0xffffu) >= c->cfg.width)
                return fail(SDM_ERR_ARG, "chain set %d: pixel (%u, %u) outside the plane", i, rc >> 16, rc & 0xffffu);
            h_pix[cp + q] = rc;
        }
        cp += m;
    }
    h_off[ck] = (int)cp;
    char* d = (char*)c->lf_buf;
    CU(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s));
    sdm::LineFitParams L;
    memset(&L, 0, sizeof(L));
    L.min_len = min_len; L.max_len = 1000; L.init_depth_count = 3;  // LineDetector.cc:20-22
    L.min_angle = 30.f; L.e1 = 1.0f; L.e2 = 1.5f;                   // :23-25
    L.sigma_lt = 0.02f;                                             // :29 (float against float)
    const int nc = (int)n_chains;
    CU(cudaEventRecord(c->lf_ev[0], s));
    sdm::k_line_fit<<<(unsigned)(((size_t)nc * 32 + sdm::kLineFitBlock - 1) / sdm::kLineFitBlock), sdm::kLineFitBlock, 0, s>>>(c->A, c->P, L, (const sdm::LineFitKf*)(d + o_kfs), nc, (const int*)(d + o_off),
                                                  (const int*)(d + o_kfi), (const uint32_t*)(d + o_pix), (const int*)(d + o_slot0),
                                                  (sdm::DevLine*)(d + o_raw), (int*)(d + o_nout));
    sdm::k_line_scan<<<1, 1024, 0, s>>>((const int*)(d + o_nout), nc, (const sdm::LineFitKf*)(d + o_kfs), n,
                                        (unsigned long long*)(d + o_offs), (unsigned long long*)(d + o_tot));
    sdm::k_line_compact<<<(nc + 255) / 256, 256, 0, s>>>((const sdm::DevLine*)(d + o_raw), (const int*)(d + o_slot0),
                                                         (const int*)(d + o_nout), (const unsigned long long*)(d + o_offs), nc,
                                                         (sdm::DevLine*)(d + o_cmp), (unsigned long long)n_slots);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->lf_ev[1], s));
    c->launches += 3;
    std::vector<unsigned long long> h_tot((size_t)n + 1);
    CU(cudaMemcpyAsync(h_tot.data(), d + o_tot, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&h_tot[n], d + o_offs + (size_t)nc * 8, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaEventElapsedTime(&c->lf_ms, c->lf_ev[0], c->lf_ev[1]));
    *total = (uint64_t)h_tot[n];
    if (counts) for (int i = 0; i < n; ++i) counts[i] = (uint64_t)h_tot[i];
    const size_t m = (size_t)std::min<uint64_t>(*total, capacity);
    if (m > 0) {
        CU(cudaMemcpyAsync(out, d + o_cmp, m * sizeof(sdm_line3d), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    uint64_t id = 0;
    RC(c->r_down.record(s, &id));
    for (int i = 0; i < n; ++i) c->kf[sets[i].kf].down_ds_id = c->kf[sets[i].kf].down_cp_id = id;
    return SDM_OK;
}
int sdm_last_line_fit_ms(sdm_ctx* c, float* ms)
{
    if (!c || !ms) return fail(SDM_ERR_ARG, "null argument");
    *ms = c->lf_ms;
    return SDM_OK;
}

// ---- Edge Drawing: stage 1 on the device, the routing walk on host threads ------------------------
struct sdm_ed_result {
    std::vector<sdm_host::EdgeChains> chains;  // host routing (and device-mode fall-backs): one pair of lists per image
    // device routing: the lists of all images in one block (image i: offsets at blob[at[i]], n_chains + 1 of them, then its pixels)
    std::vector<int32_t> blob;
    std::vector<size_t> at;
    std::vector<int32_t> blob_chains;  // chains of image i, -1: the image is in `chains`
};

namespace {

// keyframes per device chunk (H2D, stage-1 kernel, D2H; the unit the routing threads wait for): two chunks of 8 so that the
// host threads start walking after ~100 us, 16, then 32 per launch (39 MB of planes at VGA: enough to fill the device)
inline int ed_chunk_first(int ch) { return ch < 2 ? 8 * ch : (ch == 2 ? 16 : 32 * (ch - 2)); }
inline int ed_chunk_of(int i) { return i < 16 ? i / 8 : (i < 32 ? 2 : 2 + i / 32); }
inline int ed_chunk_count(int n) { return n <= 0 ? 0 : ed_chunk_of(n - 1) + 1; }
constexpr int kEdMaxBatch = 256;  // keyframes whose planes are resident at once (pinned + device: 4 bytes per pixel each)

size_t ed_bytes(int cap, size_t P);

int ed_reserve(sdm_ctx* c, int n_kf)
{
    if (!c->s_ed) CU(cudaStreamCreateWithFlags(&c->s_ed, cudaStreamNonBlocking));
    if (c->ed_cap >= n_kf) return SDM_OK;
    CU(cudaStreamSynchronize(c->s_ed));
    cudaFree(c->ed_dev);
    if (c->ed_host) cudaFreeHost(c->ed_host);
    c->ed_dev = nullptr; c->ed_host = nullptr; c->ed_cap = 0;
    const size_t bytes = ed_bytes(n_kf, c->npix);
    CU(cudaMalloc((void**)&c->ed_dev, bytes));
    CU(cudaMallocHost((void**)&c->ed_host, bytes));
    c->ed_cap = n_kf;
    return SDM_OK;
}

// planes of up to `cap` keyframes inside the (device or pinned) block: im[cap][P] u8 | G[cap][P] i16 | F[cap][P] u8, every
// section on a 256-byte boundary
// ... | NA[cap] int32 (rounded up to 64): number of anchors per keyframe (k_ed_sort; -1: more than fit) | A[cap][ed_anchor_stride]
// int32: their positions in walking order (host routing: the host thread's own counting sort is a fifth of its time)
inline size_t ed_align(size_t v) { return (v + 255) & ~(size_t)255; }
inline size_t ed_anchor_stride(size_t P) { return (P / 8 + 16 + 3) & ~(size_t)3; }  // ints; a VGA keyframe of the bench scene has 12 k anchors
size_t ed_bytes(int cap, size_t P)
{
    return ed_align((size_t)cap * P) + ed_align((size_t)cap * P * 2) + ed_align((size_t)cap * P) +
           ed_align(((((size_t)cap + 63) & ~(size_t)63) + (size_t)cap * ed_anchor_stride(P)) * 4);
}
struct EdLayout {
    uint8_t* im; int16_t* G; uint8_t* F; int32_t* NA; int32_t* A;
    EdLayout(uint8_t* base, int cap, size_t P)
        : im(base), G((int16_t*)(base + ed_align((size_t)cap * P))), F(base + ed_align((size_t)cap * P) + ed_align((size_t)cap * P * 2)),
          NA((int32_t*)(base + ed_align((size_t)cap * P) + ed_align((size_t)cap * P * 2) + ed_align((size_t)cap * P))),
          A(NA + (((size_t)cap + 63) & ~(size_t)63)) {}
};

// k_ed_sort over `count` keyframes from `first`: lists of `cap` positions at list + i * stride, numbers at n_anchors + i
int ed_launch_sort(sdm_ctx* c, const EdLayout& d, int first, int count, int32_t* list, size_t stride, int cap, int32_t* n_anchors)
{
    // (per device, and cheap: set before every launch rather than remembered per process)
    CU(cudaFuncSetAttribute(sdm::k_ed_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, sdm::kEdSortSmem));
    sdm::k_ed_sort<<<count, sdm::kEdSortThreads, sdm::kEdSortSmem, c->s_ed>>>(d.G + (size_t)first * c->npix, d.F + (size_t)first * c->npix, c->cfg.width,
                                                                               c->cfg.height, list, stride, cap, n_anchors);
    CU(cudaGetLastError());
    ++c->launches;
    return SDM_OK;
}

int ed_launch(sdm_ctx* c, const EdLayout& d, int first, int count, int grad_thresh, int anchor_thresh)
{
    const int W = c->cfg.width, H = c->cfg.height;
    const dim3 grid((W + sdm::kEdTW - 1) / sdm::kEdTW, (H + sdm::kEdTH - 1) / sdm::kEdTH, count);
    const uint8_t* im = d.im + (size_t)first * c->npix;
    int16_t* G = d.G + (size_t)first * c->npix;
    uint8_t* F = d.F + (size_t)first * c->npix;
    if ((W & 3) == 0) sdm::k_ed_planes4<<<grid, sdm::kEdThreads, 0, c->s_ed>>>(im, W, H, grad_thresh, anchor_thresh, G, F);
    else sdm::k_ed_planes<<<grid, sdm::kEdThreads, 0, c->s_ed>>>(im, W, H, grad_thresh, anchor_thresh, G, F);
    CU(cudaGetLastError());
    ++c->launches;
    return SDM_OK;
}

}  // namespace

namespace {

constexpr int kEdDevBatch = 1024;  // images routed by one k_ed_route launch (27 bytes of scratch + 9 of results per pixel each)

// sections of the device block of the routing kernel for `cap` images
struct EdRouteLayout {
    sdm_host::EdRouteCaps caps;
    size_t scratch_stride, o_scratch, o_off, o_px, o_edge, o_res, o_at, o_prof, o_na, bytes;
    EdRouteLayout(int cap, size_t P)
    {
        caps = sdm_host::EdRouteCapsFor(P);
        if (const char* e = getenv("SDM_ED_ROUTE_TEST_CAPS")) {  // test hook: tiny per-tree capacities force the host fall-back
            const int v = atoi(e);
            if (v > 0) { caps.pixels = std::min(caps.pixels, v); caps.chains = std::min(caps.chains, std::max(4, v / 4)); }
        }
        scratch_stride = ed_align(sdm_host::EdRouteScratchBytes(caps));
        o_scratch = 0;
        o_off = o_scratch + (size_t)cap * scratch_stride;
        o_px = ed_align(o_off + (size_t)cap * caps.offsets * 4);
        o_edge = ed_align(o_px + (size_t)cap * caps.out_pixels * 4);
        o_res = ed_align(o_edge + (size_t)cap * P * 4);
        o_at = ed_align(o_res + (size_t)cap * sizeof(int4));
        o_prof = ed_align(o_at + ((size_t)cap + 1) * sizeof(unsigned long long));
        o_na = ed_align(o_prof + (size_t)cap * 16 * sizeof(long long));
        bytes = ed_align(o_na + (size_t)cap * sizeof(int32_t));
    }
};

int edr_reserve(sdm_ctx* c, int n_img)
{
    for (auto& e : c->edr_ev) if (!e) CU(cudaEventCreate(&e));
    if (c->edr_cap >= n_img) return SDM_OK;
    CU(cudaStreamSynchronize(c->s_ed));
    cudaFree(c->edr_dev);
    if (c->edr_result_host) cudaFreeHost(c->edr_result_host);
    c->edr_dev = nullptr; c->edr_result_host = nullptr; c->edr_cap = 0;
    const EdRouteLayout L(n_img, c->npix);
    if (cudaMalloc((void**)&c->edr_dev, L.bytes) != cudaSuccess) {  // the caller retries with a smaller batch
        cudaGetLastError();
        c->edr_dev = nullptr;
        return fail(SDM_ERR_NOMEM, "sdm_edge_drawing: %zu bytes of routing buffers for %d images do not fit the device", L.bytes, n_img);
    }
    CU(cudaMallocHost((void**)&c->edr_result_host, (size_t)n_img * sizeof(int4) + ((size_t)n_img + 1) * sizeof(unsigned long long)));
    c->edr_cap = n_img;
    return SDM_OK;
}

// sdm_edge_drawing with both stages on the device: batches of up to kEdDevBatch images - images up, k_ed_planes(4) over the
// batch, k_ed_route (one warp per image), the chain lists gathered into one block (k_ed_chain_offsets / k_ed_chain_gather,
// into the G / F planes, which are spent by then) and brought down in ONE copy together with the per-image counts, then the
// requested edge-index planes.  Images whose routing ran out of a capacity are routed on the host.
int ed_run_device(sdm_ctx* c, int n, const sdm_ed_image* images, int grad_thresh, int anchor_thresh, sdm_ed_result* res)
{
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t P = c->npix;
    CU(cudaSetDevice(c->cfg.device));
    int cap = std::min(n, kEdDevBatch);
    RC(ed_reserve(c, cap));
    for (;;) {  // 25 bytes per pixel and image: halve the batch until the buffers fit beside the keyframe arena
        const int rrc = edr_reserve(c, cap);
        if (rrc == SDM_OK) break;
        if (rrc != SDM_ERR_NOMEM || cap == 1) return rrc;
        cap = (cap + 1) / 2;
    }
    while ((int)c->ed_ev.size() < 3) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        c->ed_ev.push_back(e);
    }
    const EdLayout dv(c->ed_dev, c->ed_cap, P), hv(c->ed_host, c->ed_cap, P);
    const EdRouteLayout L(c->edr_cap, P);
    const size_t list_room = (size_t)((uint8_t*)dv.NA - (uint8_t*)dv.G) / 4;  // int32 the G / F planes hold
    unsigned long long* at_host = reinterpret_cast<unsigned long long*>(c->edr_result_host + c->edr_cap);
    c->edr_fallbacks = 0;
    c->edr_last_n = 0;
    res->blob_chains.assign((size_t)n, -1);
    res->at.assign((size_t)n, 0);
    for (int base = 0; base < n; base += cap) {
        const int nb = std::min(cap, n - base);
        bool want_edge = false;
        for (int i = 0; i < nb; ++i) want_edge = want_edge || images[base + i].edge_index != nullptr;
        {   // rows into the pinned mirror, a few threads for a large batch (a single memcpy stream runs at ~10 GB/s)
            auto pack = [&](int i0, int i1) {
                for (int i = i0; i < i1; ++i) {
                    const sdm_ed_image& im = images[base + i];
                    uint8_t* dst = hv.im + (size_t)i * P;
                    if (im.im_step == (size_t)W) std::memcpy(dst, im.im, P);
                    else for (int y = 0; y < H; ++y) std::memcpy(dst + (size_t)y * W, im.im + (size_t)y * im.im_step, (size_t)W);
                }
            };
            const int pt = nb >= 64 ? (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1;
            if (pt == 1) pack(0, nb);
            else {
                std::vector<std::thread> th;
                for (int t = 0; t < pt; ++t) th.emplace_back(pack, (int)((long long)nb * t / pt), (int)((long long)nb * (t + 1) / pt));
                for (auto& t : th) t.join();
            }
        }
        CU(cudaMemcpyAsync(dv.im, hv.im, (size_t)nb * P, cudaMemcpyHostToDevice, c->s_ed));
        CU(cudaEventRecord(c->ed_ev[0], c->s_ed));
        RC(ed_launch(c, dv, 0, nb, grad_thresh, anchor_thresh));
        CU(cudaEventRecord(c->ed_ev[1], c->s_ed));
        sdm::EdRouteBatch b;
        b.W = W; b.H = H; b.grad_thresh = grad_thresh;
        b.G = dv.G; b.F = dv.F;
        b.scratch = c->edr_dev + L.o_scratch; b.scratch_stride = L.scratch_stride;
        b.caps = L.caps;
        b.offsets = (int32_t*)(c->edr_dev + L.o_off);
        b.pixels = (uint32_t*)(c->edr_dev + L.o_px);
        b.edge_index = (int32_t*)(c->edr_dev + L.o_edge);  // always: sdm_ed_device_edge_plane hands them to sdm_upload_keyframes
        b.result = (int4*)(c->edr_dev + L.o_res);
        const bool prof = getenv("SDM_ED_ROUTE_PROF") != nullptr;  // per-image cycle counts of the routing kernel on stderr
        b.prof = prof ? (long long*)(c->edr_dev + L.o_prof) : nullptr;
        unsigned long long* at_dev = (unsigned long long*)(c->edr_dev + L.o_at);
        int32_t* na_dev = (int32_t*)(c->edr_dev + L.o_na);
        b.n_anchors = na_dev;
        CU(cudaEventRecord(c->edr_ev[0], c->s_ed));
        static_assert(sizeof(int) == 4, "anchor slots");
        RC(ed_launch_sort(c, dv, 0, nb, (int32_t*)sdm_host::EdRouteAnchorSlots(b.scratch, L.caps), L.scratch_stride / 4, L.caps.anchors, na_dev));
        sdm::k_ed_route<<<nb, 32, 0, c->s_ed>>>(b);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->edr_ev[1], c->s_ed));
        sdm::k_ed_chain_offsets<<<1, 1024, 0, c->s_ed>>>(b.result, nb, at_dev);
        CU(cudaGetLastError());
        sdm::k_ed_chain_gather<<<dim3(8, nb), 256, 0, c->s_ed>>>(b, at_dev, (int32_t*)dv.G, (unsigned long long)list_room);
        CU(cudaGetLastError());
        c->launches += 3;
        CU(cudaMemcpyAsync(c->edr_result_host, b.result, (size_t)nb * sizeof(int4), cudaMemcpyDeviceToHost, c->s_ed));
        CU(cudaMemcpyAsync(at_host, at_dev, ((size_t)nb + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->s_ed));
        // the edge-index planes: one copy when the caller's planes are dense and consecutive, one per image otherwise (the
        // planes of images that fell back are overwritten by the host routing below)
        bool dense_edges = want_edge;
        for (int i = 0; i < nb && dense_edges; ++i) {
            const sdm_ed_image& im = images[base + i];
            dense_edges = im.edge_index && im.edge_step == (size_t)W * 4 && (i == 0 || im.edge_index == images[base + i - 1].edge_index + P);
        }
        if (dense_edges)
            CU(cudaMemcpyAsync(images[base].edge_index, b.edge_index, (size_t)nb * P * 4, cudaMemcpyDeviceToHost, c->s_ed));
        else if (want_edge)
            for (int i = 0; i < nb; ++i) {
                const sdm_ed_image& im = images[base + i];
                if (!im.edge_index) continue;
                CU(cudaMemcpy2DAsync(im.edge_index, im.edge_step, b.edge_index + (size_t)i * P, (size_t)W * 4, (size_t)W * 4, (size_t)H,
                                     cudaMemcpyDeviceToHost, c->s_ed));
            }
        CU(cudaStreamSynchronize(c->s_ed));  // (the gather ran before the copies on the same stream; its writes are bounded by the
                                             //  capacities: at most (P / 8 + P / 2 + 32) int32 per image)
        const size_t total = (size_t)at_host[nb];
        if (total > list_room) return fail(SDM_ERR_NOMEM, "sdm_edge_drawing: chain lists (%zu words) exceed the staging planes", total);
        if (total > 0) CU(cudaMemcpyAsync(hv.G, dv.G, total * 4, cudaMemcpyDeviceToHost, c->s_ed));
        CU(cudaStreamSynchronize(c->s_ed));
        const size_t blob0 = res->blob.size();
        if (total > 0)  // (one pass: no zero-fill before the copy)
            res->blob.insert(res->blob.end(), reinterpret_cast<const int32_t*>(hv.G), reinterpret_cast<const int32_t*>(hv.G) + total);
        for (int i = 0; i < nb; ++i) {
            const int4 r = c->edr_result_host[i];
            if (!r.z) continue;
            res->at[(size_t)(base + i)] = blob0 + (size_t)at_host[i];
            res->blob_chains[(size_t)(base + i)] = r.x;
        }
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ed_ev[0], c->ed_ev[1]));
        c->ed_kernel_ms += ms;
        CU(cudaEventElapsedTime(&ms, c->edr_ev[0], c->edr_ev[1]));
        c->ed_route_ms += ms;
        if (prof) {
            std::vector<long long> h((size_t)nb * 16);
            CU(cudaMemcpy(h.data(), b.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            double mean[16] = {0}, mx[16] = {0};
            for (int i = 0; i < nb; ++i)
                for (int k = 0; k < 16; ++k) { mean[k] += (double)h[(size_t)i * 16 + k] / nb; mx[k] = std::max(mx[k], (double)h[(size_t)i * 16 + k]); }
            fprintf(stderr, "k_ed_route %d images, %.2f ms: mean / max kcycles: sort %.0f / %.0f, anchor pass %.0f / %.0f, walks %.0f / %.0f, extraction %.0f / %.0f, "
                            "whole %.0f / %.0f; walked pixels %.0f / %.0f, trees %.0f / %.0f; extraction split (mean): 2nd-direction path %.0f, + copies %.0f, "
                            "1st direction %.0f, emit %.0f, other chains %.0f\n", nb, ms, mean[5] / 1e3, mx[5] / 1e3, mean[0] / 1e3, mx[0] / 1e3,
                    mean[1] / 1e3, mx[1] / 1e3, mean[2] / 1e3, mx[2] / 1e3, mean[6] / 1e3, mx[6] / 1e3, mean[3], mx[3], mean[4], mx[4], mean[8] / 1e3,
                    (mean[9] - mean[8]) / 1e3, mean[10] / 1e3, mean[11] / 1e3, mean[12] / 1e3);
        }
        // images that ran out of a capacity on the device: stage 1 again (the planes are spent), stage 2 on the host
        for (int i = 0; i < nb; ++i) {
            if (c->edr_result_host[i].z) continue;
            ++c->edr_fallbacks;
            RC(ed_launch(c, dv, i, 1, grad_thresh, anchor_thresh));
            CU(cudaMemcpyAsync(hv.G + (size_t)i * P, dv.G + (size_t)i * P, P * 2, cudaMemcpyDeviceToHost, c->s_ed));
            CU(cudaMemcpyAsync(hv.F + (size_t)i * P, dv.F + (size_t)i * P, P, cudaMemcpyDeviceToHost, c->s_ed));
            CU(cudaStreamSynchronize(c->s_ed));
            const sdm_ed_image& im = images[base + i];
            std::vector<int32_t> plane((size_t)P);  // the device's copy of the plane comes from the host routing as well
            sdm_host::EdRouteChains(W, H, hv.G + (size_t)i * P, hv.F + (size_t)i * P, grad_thresh, res->chains[(size_t)(base + i)],
                                    plane.data(), (size_t)W * 4);
            if (im.edge_index)
                for (int y = 0; y < H; ++y)
                    std::memcpy(reinterpret_cast<char*>(im.edge_index) + (size_t)y * im.edge_step, plane.data() + (size_t)y * W, (size_t)W * 4);
            CU(cudaMemcpy(b.edge_index + (size_t)i * P, plane.data(), P * 4, cudaMemcpyHostToDevice));
        }
        c->edr_last_base = base;
        c->edr_last_n = nb;
    }
    return SDM_OK;
}

}  // namespace

int sdm_set_edge_drawing_route(sdm_ctx* c, int mode)
{
    if (!c) return fail(SDM_ERR_ARG, "null context");
    if (mode != SDM_ED_ROUTE_HOST && mode != SDM_ED_ROUTE_DEVICE && mode != SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE)
        return fail(SDM_ERR_ARG, "sdm_set_edge_drawing_route: unknown mode %d", mode);
    c->ed_route_mode = mode;
    return SDM_OK;
}

int sdm_last_edge_drawing_fallbacks(sdm_ctx* c) { return c ? c->edr_fallbacks : 0; }

int sdm_ed_device_edge_plane(sdm_ctx* c, int i, const int32_t** dev_plane)
{
    if (!c || !dev_plane) return fail(SDM_ERR_ARG, "sdm_ed_device_edge_plane: null argument");
    *dev_plane = nullptr;
    if (c->edm_dev && i >= c->edm_last_base && i < c->edm_last_base + c->edm_last_n) {  // host routing, masks scattered on the device
        *dev_plane = (const int32_t*)c->edm_dev + (size_t)(i - c->edm_last_base) * c->npix;
        return SDM_OK;
    }
    if (!c->edr_dev || i < c->edr_last_base || i >= c->edr_last_base + c->edr_last_n)
        return fail(SDM_ERR_STATE, "sdm_ed_device_edge_plane: image %d is not in the last batch whose masks are on the device (%d .. %d)", i,
                    c->edr_last_base, c->edr_last_base + c->edr_last_n - 1);
    const EdRouteLayout L(c->edr_cap, c->npix);
    *dev_plane = (const int32_t*)(c->edr_dev + L.o_edge) + (size_t)(i - c->edr_last_base) * c->npix;
    return SDM_OK;
}

int sdm_edge_drawing(sdm_ctx* c, int n, const sdm_ed_image* images, int grad_thresh, int anchor_thresh, int n_threads,
                     sdm_ed_result** result)
{
    if (!c || !result || (n > 0 && !images)) return fail(SDM_ERR_ARG, "sdm_edge_drawing: null argument");
    *result = nullptr;
    if (n < 0) return fail(SDM_ERR_ARG, "sdm_edge_drawing: negative count");
    if (grad_thresh < 1 || grad_thresh > 2047 || anchor_thresh < 0) return fail(SDM_ERR_ARG, "sdm_edge_drawing: thresholds out of range");
    for (int i = 0; i < n; ++i)
        if (!images[i].im || images[i].im_step < (size_t)c->cfg.width ||
            (images[i].edge_index && images[i].edge_step < (size_t)c->cfg.width * 4))
            return fail(SDM_ERR_ARG, "sdm_edge_drawing: image %d: null plane or row pitch below the width", i);
    sdm_ed_result* res = new (std::nothrow) sdm_ed_result();
    if (!res) return fail(SDM_ERR_NOMEM, "out of host memory");
    res->chains.resize((size_t)n);
    c->ed_kernel_ms = c->ed_wall_ms = c->ed_route_ms = 0.f;
    c->edr_last_n = c->edm_last_n = 0;  // (sdm_ed_device_edge_plane: the planes of an earlier call are no longer handed out)
    c->edr_fallbacks = 0;
    if (n == 0) { *result = res; return SDM_OK; }
    const auto wall0 = std::chrono::steady_clock::now();
    if (c->ed_route_mode == SDM_ED_ROUTE_DEVICE) {
        if (!c->s_ed) CU(cudaStreamCreateWithFlags(&c->s_ed, cudaStreamNonBlocking));
        const int drc = ed_run_device(c, n, images, grad_thresh, anchor_thresh, res);
        c->ed_wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        if (drc != SDM_OK) { delete res; return drc; }
        *result = res;
        return SDM_OK;
    }
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t P = c->npix;
    int rc = SDM_OK;
    auto body = [&]() -> int {
        CU(cudaSetDevice(c->cfg.device));
        RC(ed_reserve(c, std::min(n, kEdMaxBatch)));
        const int n_chunks_max = ed_chunk_count(std::min(n, kEdMaxBatch));
        while ((int)c->ed_ev.size() < 3 * n_chunks_max) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            c->ed_ev.push_back(e);
        }
        unsigned hw = std::thread::hardware_concurrency();
        if (hw == 0) hw = 1;
        const int nt = std::max(1, std::min(n_threads > 0 ? n_threads : (int)std::min(hw, 32u), n));
        const EdLayout dv(c->ed_dev, c->ed_cap, P), hv(c->ed_host, c->ed_cap, P);
        const size_t sa = ed_anchor_stride(P);
        for (int base = 0; base < n; base += kEdMaxBatch) {
            const int nb = std::min(kEdMaxBatch, n - base), n_chunks = ed_chunk_count(nb);
            std::atomic<int> next(0), issued(0), failed(0);
            std::atomic<long long> route_ns(0);
            std::mutex mu;
            std::condition_variable cv;
            std::vector<std::thread> workers;
            for (int t = 0; t < nt; ++t)
                workers.emplace_back([&]() {
                    cudaSetDevice(c->cfg.device);
                    for (;;) {
                        const int i = next.fetch_add(1);
                        if (i >= nb) return;
                        const int chunk = ed_chunk_of(i);
                        {
                            std::unique_lock<std::mutex> lk(mu);
                            cv.wait(lk, [&] { return issued.load() > chunk || failed.load(); });
                        }
                        if (failed.load()) return;
                        if (cudaEventSynchronize(c->ed_ev[3 * chunk + 2]) != cudaSuccess) { failed.store(1); return; }
                        const auto t0 = std::chrono::steady_clock::now();
                        const sdm_ed_image& im = images[base + i];
                        const int n_sorted = chunk >= 2 ? hv.NA[i] : -1;  // (-1: the host sorts the anchors)
                        sdm_host::EdRouteChains(W, H, hv.G + (size_t)i * P, hv.F + (size_t)i * P, grad_thresh,
                                                res->chains[(size_t)(base + i)], im.edge_index, im.edge_step, hv.A + (size_t)i * sa, n_sorted);
                        route_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
                    }
                });
            auto issue = [&]() -> int {
                for (int ch = 0; ch < n_chunks; ++ch) {
                    const int first = ed_chunk_first(ch), count = std::min(ed_chunk_first(ch + 1), nb) - first;
                    for (int i = first; i < first + count; ++i) {  // pack the rows into the pinned mirror
                        const sdm_ed_image& im = images[base + i];
                        uint8_t* dst = hv.im + (size_t)i * P;
                        if (im.im_step == (size_t)W) std::memcpy(dst, im.im, P);
                        else for (int y = 0; y < H; ++y) std::memcpy(dst + (size_t)y * W, im.im + (size_t)y * im.im_step, (size_t)W);
                    }
                    CU(cudaMemcpyAsync(dv.im + (size_t)first * P, hv.im + (size_t)first * P, (size_t)count * P, cudaMemcpyHostToDevice, c->s_ed));
                    CU(cudaEventRecord(c->ed_ev[3 * ch], c->s_ed));
                    RC(ed_launch(c, dv, first, count, grad_thresh, anchor_thresh));
                    CU(cudaEventRecord(c->ed_ev[3 * ch + 1], c->s_ed));
                    CU(cudaMemcpyAsync(hv.G + (size_t)first * P, dv.G + (size_t)first * P, (size_t)count * P * 2, cudaMemcpyDeviceToHost, c->s_ed));
                    CU(cudaMemcpyAsync(hv.F + (size_t)first * P, dv.F + (size_t)first * P, (size_t)count * P, cudaMemcpyDeviceToHost, c->s_ed));
                    if (ch >= 2) {  // (the first 16 keyframes go to the host threads at once: they sort their anchors themselves)
                        RC(ed_launch_sort(c, dv, first, count, dv.A + (size_t)first * sa, sa, (int)sa, dv.NA + first));
                        CU(cudaMemcpyAsync(hv.NA + first, dv.NA + first, (size_t)count * 4, cudaMemcpyDeviceToHost, c->s_ed));
                        CU(cudaMemcpyAsync(hv.A + (size_t)first * sa, dv.A + (size_t)first * sa, (size_t)count * sa * 4, cudaMemcpyDeviceToHost, c->s_ed));
                    }
                    CU(cudaEventRecord(c->ed_ev[3 * ch + 2], c->s_ed));
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        issued.store(ch + 1);
                    }
                    cv.notify_all();
                }
                return SDM_OK;
            };
            int irc = issue();
            if (irc != SDM_OK) {
                {
                    std::lock_guard<std::mutex> lk(mu);
                    failed.store(1);
                }
                cv.notify_all();
            }
            for (auto& t : workers) t.join();
            if (irc != SDM_OK) return irc;
            if (failed.load()) return fail(SDM_ERR_CUDA, "sdm_edge_drawing: device stage failed: %s", cudaGetErrorString(cudaGetLastError()));
            CU(cudaStreamSynchronize(c->s_ed));
            for (int ch = 0; ch < n_chunks; ++ch) {
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, c->ed_ev[3 * ch], c->ed_ev[3 * ch + 1]));
                c->ed_kernel_ms += ms;
            }
            c->ed_route_ms += (float)((double)route_ns.load() * 1e-6);
            if (c->ed_route_mode == SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE) {
                // kf->mEdgeIndex of this batch on the device: the chain lists go up (a tenth of the planes' bytes), k_ed_mask
                // writes the chain numbers into planes of -1
                std::vector<sdm::EdMaskImage> desc((size_t)nb);
                size_t n_pix = 0, n_off = 0;
                for (int i = 0; i < nb; ++i) {
                    const sdm_host::EdgeChains& e = res->chains[(size_t)(base + i)];
                    desc[(size_t)i].pix_begin = n_pix; desc[(size_t)i].off_begin = n_off;
                    desc[(size_t)i].n_chains = e.n_chains(); desc[(size_t)i].n_pixels = (int)e.pixels.size();
                    n_pix += e.pixels.size(); n_off += e.offsets.size();
                }
                const size_t o_pix = ed_align((size_t)nb * P * 4), o_off = ed_align(o_pix + n_pix * 4), o_desc = ed_align(o_off + n_off * 4),
                             bytes = ed_align(o_desc + (size_t)nb * sizeof(sdm::EdMaskImage));
                if (bytes > c->edm_bytes) {
                    cudaFree(c->edm_dev);
                    c->edm_dev = nullptr; c->edm_bytes = 0;
                    CU(cudaMalloc((void**)&c->edm_dev, bytes + bytes / 4));
                    c->edm_bytes = bytes + bytes / 4;
                }
                CU(cudaMemsetAsync(c->edm_dev, 0xff, (size_t)nb * P * 4, c->s_ed));
                int max_pix = 0;
                for (int i = 0; i < nb; ++i) {
                    const sdm_host::EdgeChains& e = res->chains[(size_t)(base + i)];
                    if (!e.pixels.empty())
                        CU(cudaMemcpyAsync(c->edm_dev + o_pix + desc[(size_t)i].pix_begin * 4, e.pixels.data(), e.pixels.size() * 4, cudaMemcpyHostToDevice, c->s_ed));
                    CU(cudaMemcpyAsync(c->edm_dev + o_off + desc[(size_t)i].off_begin * 4, e.offsets.data(), e.offsets.size() * 4, cudaMemcpyHostToDevice, c->s_ed));
                    max_pix = std::max(max_pix, desc[(size_t)i].n_pixels);
                }
                CU(cudaMemcpyAsync(c->edm_dev + o_desc, desc.data(), (size_t)nb * sizeof(sdm::EdMaskImage), cudaMemcpyHostToDevice, c->s_ed));
                if (max_pix > 0) {
                    sdm::k_ed_mask<<<dim3((unsigned)std::min(64, (max_pix + 255) / 256), (unsigned)nb), 256, 0, c->s_ed>>>(
                        (const sdm::EdMaskImage*)(c->edm_dev + o_desc), (const uint32_t*)(c->edm_dev + o_pix), (const int32_t*)(c->edm_dev + o_off), W, H,
                        (int32_t*)c->edm_dev);
                    CU(cudaGetLastError());
                    ++c->launches;
                }
                CU(cudaStreamSynchronize(c->s_ed));  // (the copies read the result's own vectors: done before the call returns)
                c->edm_last_base = base;
                c->edm_last_n = nb;
            }
        }
        return SDM_OK;
    };
    rc = body();
    c->ed_wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (rc != SDM_OK) { delete res; return rc; }
    *result = res;
    return SDM_OK;
}

int sdm_ed_chains(const sdm_ed_result* r, int i, int32_t* n_chains, const int32_t** offsets, const uint32_t** pixels)
{
    if (!r || !n_chains || !offsets || !pixels) return fail(SDM_ERR_ARG, "sdm_ed_chains: null argument");
    if (i < 0 || i >= (int)r->chains.size()) return fail(SDM_ERR_ARG, "sdm_ed_chains: keyframe %d out of range", i);
    if ((size_t)i < r->blob_chains.size() && r->blob_chains[(size_t)i] >= 0) {
        *n_chains = r->blob_chains[(size_t)i];
        *offsets = r->blob.data() + r->at[(size_t)i];
        *pixels = reinterpret_cast<const uint32_t*>(*offsets + *n_chains + 1);
        return SDM_OK;
    }
    const sdm_host::EdgeChains& e = r->chains[(size_t)i];
    *n_chains = e.n_chains();
    *offsets = e.offsets.data();
    *pixels = e.pixels.data();
    return SDM_OK;
}

void sdm_ed_free(sdm_ed_result* r) { delete r; }

int sdm_last_edge_drawing_ms(sdm_ctx* c, float* kernel_ms, float* wall_ms, float* route_thread_ms)
{
    if (!c) return fail(SDM_ERR_ARG, "null argument");
    if (kernel_ms) *kernel_ms = c->ed_kernel_ms;
    if (wall_ms) *wall_ms = c->ed_wall_ms;
    if (route_thread_ms) *route_thread_ms = c->ed_route_ms;
    return SDM_OK;
}

int sdm_ed_planes(sdm_ctx* c, const uint8_t* im, size_t im_step, int grad_thresh, int anchor_thresh, int16_t* G, uint8_t* F)
{
    if (!c || !im || !G || !F) return fail(SDM_ERR_ARG, "sdm_ed_planes: null argument");
    if (im_step < (size_t)c->cfg.width) return fail(SDM_ERR_ARG, "sdm_ed_planes: row pitch below the width");
    if (grad_thresh < 1 || grad_thresh > 2047 || anchor_thresh < 0) return fail(SDM_ERR_ARG, "sdm_ed_planes: thresholds out of range");
    CU(cudaSetDevice(c->cfg.device));
    RC(ed_reserve(c, 1));
    const size_t P = c->npix;
    const EdLayout dv(c->ed_dev, c->ed_cap, P);
    CU(cudaMemcpy2DAsync(dv.im, (size_t)c->cfg.width, im, im_step, (size_t)c->cfg.width, (size_t)c->cfg.height, cudaMemcpyHostToDevice, c->s_ed));
    RC(ed_launch(c, dv, 0, 1, grad_thresh, anchor_thresh));
    CU(cudaMemcpyAsync(G, dv.G, P * 2, cudaMemcpyDeviceToHost, c->s_ed));
    CU(cudaMemcpyAsync(F, dv.F, P, cudaMemcpyDeviceToHost, c->s_ed));
    CU(cudaStreamSynchronize(c->s_ed));
    return SDM_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------
int sdm_depth_plane_ptr(sdm_ctx* c, int kf, void** dev_ptr, size_t* bytes)
{
    if (!c || !dev_ptr) return fail(SDM_ERR_ARG, "null argument");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    *dev_ptr = (void*)(c->A.rs + (size_t)kf * c->npix);
    if (bytes) *bytes = c->npix * sizeof(float2);
    return SDM_OK;
}

int sdm_export_arena(sdm_ctx* c, void* handle64, size_t* slot_bytes)
{
    if (!c || !handle64) return fail(SDM_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(c->cfg.device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->A.rs));
    memcpy(handle64, &h, 64);
    if (slot_bytes) *slot_bytes = c->npix * sizeof(float2);
    return SDM_OK;
}

int sdm_import_peer_arena(sdm_ctx* c, int peer_rank, const void* handle64)
{
    if (!c || !handle64) return fail(SDM_ERR_ARG, "null argument");
    if (peer_rank < 0 || peer_rank >= kMaxPeers) return fail(SDM_ERR_ARG, "peer rank %d out of [0,%d)", peer_rank, kMaxPeers);
    CU(cudaSetDevice(c->cfg.device));
    if (c->peer_rs[peer_rank]) {
        CU(cudaIpcCloseMemHandle(c->peer_rs[peer_rank]));
        c->peer_rs[peer_rank] = nullptr;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_rs[peer_rank] = p;
    return SDM_OK;
}

int sdm_pull_halo(sdm_ctx* c, int n, const int32_t* local_slot, const int32_t* peer_rank, const int32_t* peer_slot)
{
    if (!c || (n > 0 && (!local_slot || !peer_rank || !peer_slot))) return fail(SDM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    const size_t bytes = c->npix * sizeof(float2);
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(c, local_slot[i])) return fail(SDM_ERR_ARG, "local slot %d out of range", local_slot[i]);
        const int pr = peer_rank[i];
        if (pr < 0 || pr >= kMaxPeers || !c->peer_rs[pr]) return fail(SDM_ERR_STATE, "peer %d arena not imported", pr);
        if (peer_slot[i] < 0) return fail(SDM_ERR_ARG, "peer slot %d negative", peer_slot[i]);
        KfState& k = c->kf[local_slot[i]];
        RC(c->r_down.wait(c->s_compute, k.down_ds_id));  // (k_pack of the slot is earlier on this same stream)
        const char* src = (const char*)c->peer_rs[pr] + (size_t)peer_slot[i] * bytes;
        CU(cudaMemcpyAsync(c->A.rs + (size_t)local_slot[i] * c->npix, src, bytes, cudaMemcpyDeviceToDevice, c->s_compute));
        RC(c->r_compute.record(c->s_compute, &k.comp_id));
        k.pass1_done = true;
        k.rs_dense = true;
        k.split_stale = true;
    }
    return SDM_OK;
}

int sdm_mark_pass1_done(sdm_ctx* c, int kf)
{
    if (!c) return fail(SDM_ERR_ARG, "null context");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    c->kf[kf].pass1_done = true;
    c->kf[kf].rs_dense = true;
    c->kf[kf].split_stale = true;
    return SDM_OK;
}

// ---- exchange ordered on the devices --------------------------------------------------------------
namespace {
struct PeerHandleLayout {
    cudaIpcMemHandle_t rs, flags;
    int32_t n_slots, rank;
    uint64_t slot_bytes;
};
static_assert(sizeof(PeerHandleLayout) <= SDM_PEER_HANDLE_BYTES, "sdm_peer_handle too small");
}  // namespace

int sdm_export_peer_handle(sdm_ctx* c, int my_rank, sdm_peer_handle* out)
{
    if (!c || !out) return fail(SDM_ERR_ARG, "null argument");
    if (my_rank < 0 || my_rank >= kMaxPeers) return fail(SDM_ERR_ARG, "rank %d out of [0,%d)", my_rank, kMaxPeers);
    CU(cudaSetDevice(c->cfg.device));
    PeerHandleLayout h;
    memset(&h, 0, sizeof(h));
    CU(cudaIpcGetMemHandle(&h.rs, c->A.rs));
    CU(cudaIpcGetMemHandle(&h.flags, c->xflags));
    h.n_slots = (int32_t)c->kf.size();
    h.rank = my_rank;
    h.slot_bytes = c->npix * sizeof(float2);
    memset(out, 0, sizeof(*out));
    memcpy(out->bytes, &h, sizeof(h));
    c->my_rank = my_rank;
    return SDM_OK;
}

int sdm_import_peer(sdm_ctx* c, int peer_rank, const sdm_peer_handle* handle)
{
    if (!c || !handle) return fail(SDM_ERR_ARG, "null argument");
    if (peer_rank < 0 || peer_rank >= kMaxPeers) return fail(SDM_ERR_ARG, "peer rank %d out of [0,%d)", peer_rank, kMaxPeers);
    if (c->my_rank < 0) return fail(SDM_ERR_STATE, "sdm_import_peer before sdm_export_peer_handle (this rank's number is unknown)");
    PeerHandleLayout h;
    memcpy(&h, handle->bytes, sizeof(h));
    if (h.rank != peer_rank) return fail(SDM_ERR_ARG, "handle belongs to rank %d, not %d", h.rank, peer_rank);
    if (h.slot_bytes != c->npix * sizeof(float2)) return fail(SDM_ERR_ARG, "peer %d uses another image size", peer_rank);
    CU(cudaSetDevice(c->cfg.device));
    if (c->peer_rs2[peer_rank]) { CU(cudaIpcCloseMemHandle(c->peer_rs2[peer_rank])); c->peer_rs2[peer_rank] = nullptr; }
    if (c->peer_flags[peer_rank]) { CU(cudaIpcCloseMemHandle(c->peer_flags[peer_rank])); c->peer_flags[peer_rank] = nullptr; }
    void *p = nullptr, *f = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h.rs, cudaIpcMemLazyEnablePeerAccess));
    c->peer_rs2[peer_rank] = p;
    CU(cudaIpcOpenMemHandle(&f, h.flags, cudaIpcMemLazyEnablePeerAccess));
    c->peer_flags[peer_rank] = (sdm::XFlags*)f;
    c->peer_nslots[peer_rank] = h.n_slots;
    // tell the owner that this rank pulls from it: its next pass 1 waits for our acknowledgement
    sdm::k_xregister<<<1, 1, 0, c->s_compute>>>(c->peer_flags[peer_rank], c->my_rank);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->s_compute));
    return SDM_OK;
}

int sdm_set_halo(sdm_ctx* c, int n, const int32_t* local_slot, const int32_t* peer_rank, const int32_t* peer_slot)
{
    if (!c || (n > 0 && (!local_slot || !peer_rank || !peer_slot))) return fail(SDM_ERR_ARG, "null argument");
    for (int i = 0; i < n; ++i) {
        if (!slot_ok(c, local_slot[i])) return fail(SDM_ERR_ARG, "local slot %d out of range", local_slot[i]);
        const int pr = peer_rank[i];
        if (pr < 0 || pr >= kMaxPeers || !c->peer_flags[pr]) return fail(SDM_ERR_STATE, "peer %d not imported (sdm_import_peer)", pr);
        if (peer_slot[i] < 0 || peer_slot[i] >= c->peer_nslots[pr])
            return fail(SDM_ERR_ARG, "peer slot %d out of [0,%d) of rank %d", peer_slot[i], c->peer_nslots[pr], pr);
    }
    c->halo_local.assign(local_slot, local_slot + std::max(n, 0));
    c->halo_rank.assign(peer_rank, peer_rank + std::max(n, 0));
    c->halo_slot.assign(peer_slot, peer_slot + std::max(n, 0));
    std::fill(c->is_halo.begin(), c->is_halo.end(), 0);
    for (int i = 0; i < n; ++i) c->is_halo[local_slot[i]] = 1;
    return SDM_OK;
}

int sdm_exchange(sdm_ctx* c)
{
    if (!c) return fail(SDM_ERR_ARG, "null context");
    if (c->my_rank < 0) return SDM_OK;  // single GPU: nothing to exchange
    CU(cudaSetDevice(c->cfg.device));
    const unsigned step = ++c->xstep;
    // everything queued on the compute stream so far (pass 1 of this step) precedes the flag
    sdm::k_xpublish<<<1, 1, 0, c->s_compute>>>(c->xflags, step);
    c->launches++;
    const int n = (int)c->halo_local.size();
    if (n == 0) { CU(cudaGetLastError()); return SDM_OK; }
    cudaStream_t s = c->s_halo;
    uint64_t need_c = 0, need_d = 0;
    unsigned peers = 0;
    for (int i = 0; i < n; ++i) {
        const KfState& k = c->kf[c->halo_local[i]];
        need_c = std::max(need_c, k.comp_id);   // last kernel that read the slot (pass 2 of the previous step) or packed it
        need_d = std::max(need_d, k.down_ds_id);
        peers |= 1u << c->halo_rank[i];
    }
    RC(c->r_compute.wait(s, need_c));
    RC(c->r_down.wait(s, need_d));
    for (int r = 0; r < kMaxPeers; ++r)
        if (peers >> r & 1u) { sdm::k_xwait_done<<<1, 1, 0, s>>>(c->peer_flags[r], step, c->xflags); c->launches++; }
    const size_t bytes = c->npix * sizeof(float2);
    for (int i = 0; i < n; ++i) {
        const char* src = (const char*)c->peer_rs2[c->halo_rank[i]] + (size_t)c->halo_slot[i] * bytes;
        CU(cudaMemcpyAsync(c->A.rs + (size_t)c->halo_local[i] * c->npix, src, bytes, cudaMemcpyDeviceToDevice, s));
    }
    for (int r = 0; r < kMaxPeers; ++r)
        if (peers >> r & 1u) { sdm::k_xack<<<1, 1, 0, s>>>(c->peer_flags[r], c->my_rank, step); c->launches++; }
    CU(cudaGetLastError());
    uint64_t id = 0;
    RC(c->r_compute.record(s, &id));
    for (int i = 0; i < n; ++i) {
        KfState& k = c->kf[c->halo_local[i]];
        k.comp_id = id;
        k.pull_id = id;
        k.pass1_done = true;
        k.rs_dense = true;
        k.split_stale = true;
    }
    return SDM_OK;
}

// results of a chunk to the host: the block-sparse kernel when the caller vouches for zero-initialised destination planes
// and they are pinned, the dense DMA otherwise
// mode 0: dense DMA of every plane; 1: block-sparse kernel for every plane (pinned, zero-initialised destinations);
// 2: dense DMA of the three 4-byte planes + block-sparse kernel for SemiDensePointSets_ (61 % of the dense bytes) on its
// own stream, so that the copy engine and the SM-issued writes share the link
static int download(sdm_ctx* c, int n, const sdm_download_desc* d, int mode)
{
    if (mode) {
        bool ok = true;
        for (int i = 0; i < n && ok; ++i) ok = slot_ok(c, d[i].kf) && !c->kf[d[i].kf].rs_dense && !c->kf[d[i].kf].split_stale;
        if (ok && mode == 1) {
            CU(cudaSetDevice(c->cfg.device));
            const int rc = sparse_to_pinned(c, n, d, c->s_down);
            if (rc != 1) return rc;
        } else if (ok) {
            bool any_pts = false;
            std::vector<sdm_download_desc> dense(d, d + n), pts(d, d + n);
            for (int i = 0; i < n; ++i) {
                any_pts |= d[i].points != nullptr;
                dense[i].points = nullptr;
                pts[i].depth = pts[i].sigma = pts[i].checked = nullptr;
            }
            if (any_pts) {
                CU(cudaSetDevice(c->cfg.device));
                const int rc = sparse_to_pinned(c, n, pts.data(), c->s_down_k);
                if (rc < 0) return rc;
                if (rc == 0) return sdm_download_keyframes(c, n, dense.data());
            }
        }
    }
    return sdm_download_keyframes(c, n, d);
}

// ---- one SemiDenseLoop as a pipeline ---------------------------------------------------------------
int sdm_run_loop(sdm_ctx* c, const sdm_loop* L)
{
    if (!c || !L) return fail(SDM_ERR_ARG, "sdm_run_loop: null argument");
    if ((L->n_upload > 0 && !L->upload) || (L->n_pass1 > 0 && !L->pass1) || (L->n_pass2 > 0 && !L->pass2))
        return fail(SDM_ERR_ARG, "sdm_run_loop: null array");
    const int CH = L->chunk > 0 ? L->chunk : 4;
    const int n1 = std::max(L->n_pass1, 0), n2 = std::max(L->n_pass2, 0), nu = std::max(L->n_upload, 0);
    const int nslots = (int)c->kf.size();
    std::vector<int> up_pos(nslots, -1), p1_pos(nslots, -1);
    for (int i = 0; i < nu; ++i) {
        if (!slot_ok(c, L->upload[i].kf)) return fail(SDM_ERR_ARG, "upload slot %d out of range", L->upload[i].kf);
        up_pos[L->upload[i].kf] = i;
    }
    auto slots_ok = [&](const sdm_item& it) {
        if (!slot_ok(c, it.kf) || it.n_nbr < 0 || it.n_nbr > SDM_MAX_NBR) return false;
        for (int j = 0; j < it.n_nbr; ++j)
            if (!slot_ok(c, it.nbr[j])) return false;
        return true;
    };
    for (int i = 0; i < n1; ++i) {
        if (!slots_ok(L->pass1[i])) return fail(SDM_ERR_ARG, "pass-1 work order %d names a slot out of range", i);
        p1_pos[L->pass1[i].kf] = i;
    }
    // pass-2 work order i may be queued once pass 1 of work order ready[i] has been (-1: at once); work orders that
    // read a halo plane wait for the exchange
    std::vector<int> ready(n2), order(n2);
    std::vector<char> after_x(n2, 0);
    for (int i = 0; i < n2; ++i) {
        const sdm_item& it = L->pass2[i];
        if (!slots_ok(it)) return fail(SDM_ERR_ARG, "pass-2 work order %d names a slot out of range", i);
        int r = p1_pos[it.kf];
        bool halo = L->exchange && c->is_halo[it.kf];
        for (int j = 0; j < it.n_nbr; ++j) {
            r = std::max(r, p1_pos[it.nbr[j]]);
            halo = halo || (L->exchange && c->is_halo[it.nbr[j]]);
        }
        ready[i] = r;
        after_x[i] = halo;
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ready[a] < ready[b]; });
    std::vector<sdm_item> batch;
    std::vector<sdm_download_desc> dl;
    std::vector<int> deferred;
    int next2 = 0, next_up = 0;
    auto run2 = [&](const std::vector<int>& idx) -> int {
        for (size_t i0 = 0; i0 < idx.size(); i0 += (size_t)CH) {
            const size_t m = std::min<size_t>(CH, idx.size() - i0);
            batch.clear(); dl.clear();
            for (size_t i = 0; i < m; ++i) {
                batch.push_back(L->pass2[idx[i0 + i]]);
                if (L->down2) dl.push_back(L->down2[idx[i0 + i]]);
            }
            RC(sdm_pass2(c, (int)m, batch.data()));
            if (!dl.empty()) RC(download(c, (int)dl.size(), dl.data(), L->sparse_download));
        }
        return SDM_OK;
    };
    std::vector<int> now;
    for (int i0 = 0; i0 < n1 || i0 == 0; i0 += CH) {
        const int m = std::min(CH, n1 - i0);
        if (m > 0) {
            int need = -1;  // uploads go out in list order up to the last one this chunk needs
            for (int i = 0; i < m; ++i) {
                const sdm_item& it = L->pass1[i0 + i];
                need = std::max(need, up_pos[it.kf]);
                for (int j = 0; j < it.n_nbr; ++j) need = std::max(need, up_pos[it.nbr[j]]);
            }
            if (need >= next_up) {
                RC(sdm_upload_keyframes(c, need + 1 - next_up, L->upload + next_up));
                next_up = need + 1;
            }
            RC(sdm_pass1(c, m, L->pass1 + i0));
            if (L->down1) RC(download(c, m, L->down1 + i0, L->sparse_download));
        }
        // pass 2 of the work orders whose pass-1 inputs are all queued, one chunk behind: the newest chunk's pass 1 is
        // queued first so that the SMs never wait for a download
        now.clear();
        const int done_before = i0;  // pass-1 work orders [0, i0) were queued before this chunk
        while (next2 < n2 && ready[order[next2]] < done_before) {
            const int i = order[next2++];
            if (after_x[i]) deferred.push_back(i); else now.push_back(i);
        }
        RC(run2(now));
        if (m <= 0) break;
    }
    if (next_up < nu) RC(sdm_upload_keyframes(c, nu - next_up, L->upload + next_up));
    if (L->exchange) RC(sdm_exchange(c));
    now.clear();
    while (next2 < n2) {
        const int i = order[next2++];
        if (after_x[i]) deferred.push_back(i); else now.push_back(i);
    }
    RC(run2(now));
    std::sort(deferred.begin(), deferred.end());
    return run2(deferred);
}

// ---- per-method entry points ---------------------------------------------------------------------
int sdm_pair_geometry(const float K1[4], const float Tcw1[12], const float K2[4], const float Tcw2[12],
                      sdm_pair_geometry_t* out)
{
    if (!K1 || !Tcw1 || !K2 || !Tcw2 || !out) return fail(SDM_ERR_ARG, "null argument");
    const sdm::PairGeometry g = sdm::pair_geometry(K1, Tcw1, K2, Tcw2);
    memcpy(out->R21, g.R21.m, sizeof(out->R21));
    memcpy(out->t21, g.t21.v, sizeof(out->t21));
    memcpy(out->F12, g.F12.m, sizeof(out->F12));
    return SDM_OK;
}

int sdm_stereo_search_constraints(const float* inv_depths, int n, float* min_depth, float* max_depth)
{
    if (!inv_depths || !min_depth || !max_depth || n <= 0) return fail(SDM_ERR_ARG, "bad argument");
    sdm::stereo_search_constraints(inv_depths, n, min_depth, max_depth);
    return SDM_OK;
}

static int single_pair_item(sdm_ctx* c, int kf1, int kf2, float mind, float maxd, float rot)
{
    sdm_item it;
    memset(&it, 0, sizeof(it));
    it.kf = kf1;
    it.n_nbr = 1;
    it.nbr[0] = kf2;
    it.rot_deg[0] = rot;
    it.min_depth = mind;
    it.max_depth = maxd;
    Batch b;
    RC(prepare_batch(c, 1, &it, false, false, &b));
    return finish_batch(c, 1, &it, b);  // the debug kernels follow on s_compute right away (callers synchronise)
}

static int ensure_dbg(sdm_ctx* c)
{
    if (!c->dbg) CU(cudaMalloc(&c->dbg, c->npix * (4 * sizeof(float) + 1)));
    return SDM_OK;
}

int sdm_inter_chi_test(sdm_ctx* c, int n, const float* diff, const float* sigma, uint8_t* accept)
{
    if (!c || (n > 0 && (!diff || !sigma || !accept))) return fail(SDM_ERR_ARG, "null argument");
    if (n <= 0) return SDM_OK;
    CU(cudaSetDevice(c->cfg.device));
    float* d = nullptr;
    CU(cudaMalloc(&d, (size_t)n * 9));
    uint8_t* da = reinterpret_cast<uint8_t*>(d + 2 * (size_t)n);
    cudaStream_t s = c->s_compute;
    cudaError_t e = cudaMemcpyAsync(d, diff, (size_t)n * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, sigma, (size_t)n * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        sdm::k_chi_inter<<<(n + 255) / 256, 256, 0, s>>>(c->P, n, d, d + n, da);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(accept, da, (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    c->launches++;
    if (e != cudaSuccess) return fail(SDM_ERR_CUDA, "sdm_inter_chi_test: %s", cudaGetErrorString(e));
    return SDM_OK;
}

int sdm_search_range(sdm_ctx* c, int kf1, int kf2, int px, int py, float mind, float maxd, float* umin, float* umax)
{
    if (!c || !umin || !umax) return fail(SDM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    RC(single_pair_item(c, kf1, kf2, mind, maxd, 0.f));
    RC(ensure_dbg(c));
    sdm::k_search_range<<<1, 1, 0, c->s_compute>>>(c->P, c->d_items, px, py, c->dbg);
    CU(cudaGetLastError());
    c->launches++;
    float h[2];
    CU(cudaMemcpyAsync(h, c->dbg, sizeof(h), cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    *umin = h[0];
    *umax = h[1];
    return SDM_OK;
}

int sdm_epipolar_search(sdm_ctx* c, int kf1, int kf2, int x, int y, float pixel, float min_depth, float max_depth,
                        float th_pi, float rot_deg, sdm_hypothesis* out)
{
    if (!c || !out) return fail(SDM_ERR_ARG, "null argument");
    if (x < 0 || y < 0 || x >= c->cfg.width || y >= c->cfg.height) return fail(SDM_ERR_ARG, "pixel (%d,%d) outside the image", x, y);
    CU(cudaSetDevice(c->cfg.device));
    RC(single_pair_item(c, kf1, kf2, min_depth, max_depth, rot_deg));
    RC(ensure_dbg(c));
    float* d = c->dbg;
    uint8_t* ok = (uint8_t*)(c->dbg + 4);
    sdm::k_pair_hypotheses<<<1, sdm::kPass1Warps * 32, 0, c->s_compute>>>(c->A, c->P, c->d_items, d, d + 1, d + 2, d + 3, ok,
                                                                          (y << 16) | x, pixel, th_pi);
    CU(cudaGetLastError());
    c->launches++;
    float h[5];
    CU(cudaMemcpyAsync(h, c->dbg, sizeof(h), cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    uint8_t okh;
    memcpy(&okh, &h[4], 1);
    out->depth = h[0];
    out->sigma = h[1];
    out->best_u = h[2];
    out->best_v = h[3];
    out->supported = okh ? 1 : 0;
    return SDM_OK;
}

int sdm_epipolar_search_plane(sdm_ctx* c, int kf1, int kf2, float min_depth, float max_depth, float rot_deg,
                              float* hyp_depth, float* hyp_sigma, float* hyp_u, uint8_t* ok)
{
    if (!c || !hyp_depth || !hyp_sigma || !hyp_u || !ok) return fail(SDM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    RC(single_pair_item(c, kf1, kf2, min_depth, max_depth, rot_deg));
    RC(ensure_dbg(c));
    const size_t P = c->npix;
    float* d = c->dbg;
    uint8_t* dok = (uint8_t*)(c->dbg + 4 * P);
    CU(cudaMemsetAsync(c->dbg, 0, P * (4 * sizeof(float) + 1), c->s_compute));
    // upper bound of the candidate count (the kernel exits above the real one): no host round trip
    const unsigned blocks = (unsigned)((P + sdm::kPass1Warps - 1) / sdm::kPass1Warps);
    sdm::k_pair_hypotheses<<<blocks, sdm::kPass1Warps * 32, 0, c->s_compute>>>(c->A, c->P, c->d_items, d, d + P, d + 2 * P, nullptr,
                                                                               dok, -1, 0.f, 0.f);
    CU(cudaGetLastError());
    c->launches++;
    CU(cudaMemcpyAsync(hyp_depth, d, P * 4, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaMemcpyAsync(hyp_sigma, d + P, P * 4, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaMemcpyAsync(hyp_u, d + 2 * P, P * 4, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaMemcpyAsync(ok, dok, P, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    return SDM_OK;
}

int sdm_fuse(sdm_ctx* c, int m, int n, const float* depth, const float* sigma, const int32_t* count, float* out_depth,
             float* out_sigma, int32_t* out_supported)
{
    if (!c || !depth || !sigma || !count || !out_depth || !out_sigma || !out_supported) return fail(SDM_ERR_ARG, "null argument");
    if (m <= 0) return SDM_OK;
    if (n < 1 || n > 32) return fail(SDM_ERR_ARG, "set size %d out of [1,32]", n);
    for (int i = 0; i < m; ++i)
        if (count[i] < 0 || count[i] > n) return fail(SDM_ERR_ARG, "count[%d]=%d out of [0,%d]", i, count[i], n);
    CU(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->s_compute;
    float *dd = nullptr, *ds = nullptr, *od = nullptr, *os = nullptr;
    int *dc = nullptr, *ok = nullptr;
    const size_t mn = (size_t)m * n;
    CU(cudaMalloc(&dd, mn * 4)); CU(cudaMalloc(&ds, mn * 4)); CU(cudaMalloc(&dc, (size_t)m * 4));
    CU(cudaMalloc(&od, (size_t)m * 4)); CU(cudaMalloc(&os, (size_t)m * 4)); CU(cudaMalloc(&ok, (size_t)m * 4));
    CU(cudaMemcpyAsync(dd, depth, mn * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ds, sigma, mn * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dc, count, (size_t)m * 4, cudaMemcpyHostToDevice, s));
    sdm::k_fuse_sets<<<(m + 7) / 8, 256, 0, s>>>(c->P, m, n, dd, ds, dc, od, os, ok);
    CU(cudaGetLastError());
    c->launches++;
    CU(cudaMemcpyAsync(out_depth, od, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(out_sigma, os, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(out_supported, ok, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    cudaFree(dd); cudaFree(ds); cudaFree(dc); cudaFree(od); cudaFree(os); cudaFree(ok);
    return SDM_OK;
}

static int intra_single(sdm_ctx* c, int kf, bool check, bool grow)
{
    if (!c) return fail(SDM_ERR_ARG, "null context");
    if (!slot_ok(c, kf)) return fail(SDM_ERR_ARG, "keyframe slot %d out of range", kf);
    KfState& k = c->kf[kf];
    if (!k.pass1_done || (grow && !k.uploaded)) return fail(SDM_ERR_STATE, "slot %d has no pass-1 planes", kf);
    CU(cudaSetDevice(c->cfg.device));
    RC(c->r_down.wait(c->s_compute, k.down_ds_id));
    int* d_slots;
    ItemStage* st;
    RC(single_slot_array(c, kf, &d_slots, &st));
    RC(run_intra(c, d_slots, 1, check, grow));
    k.split_stale = false;  // the second stencil stage rewrites dpl / spl from the final (rho,sigma) plane
    RC(c->r_compute.record(c->s_compute, &k.comp_id));
    st->busy = k.comp_id;
    return SDM_OK;
}

int sdm_intra_check(sdm_ctx* c, int kf) { return intra_single(c, kf, true, false); }
int sdm_intra_grow(sdm_ctx* c, int kf) { return intra_single(c, kf, false, true); }

int sdm_inter_check(sdm_ctx* c, const sdm_item* item)
{
    if (!c || !item) return fail(SDM_ERR_ARG, "null argument");
    return sdm_pass2(c, 1, item);
}

// ---- measurement ---------------------------------------------------------------------------------
int sdm_last_pass_ms(sdm_ctx* c, float* pass1_ms, float* pass2_ms)
{
    if (!c) return fail(SDM_ERR_ARG, "null context");
    CU(cudaSetDevice(c->cfg.device));
    if (pass1_ms) {
        *pass1_ms = 0.f;
        if (c->p1_timed) {
            CU(cudaEventSynchronize(c->ev_p1[1]));
            CU(cudaEventElapsedTime(pass1_ms, c->ev_p1[0], c->ev_p1[1]));
        }
    }
    if (pass2_ms) {
        *pass2_ms = 0.f;
        if (c->p2_timed) {
            CU(cudaEventSynchronize(c->ev_p2[1]));
            CU(cudaEventElapsedTime(pass2_ms, c->ev_p2[0], c->ev_p2[1]));
        }
    }
    return SDM_OK;
}

long long sdm_launch_count(sdm_ctx* c) { return c ? c->launches : 0; }

int sdm_last_timing(sdm_ctx* c, sdm_timing* out)
{
    if (!c || !out) return fail(SDM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    memset(out, 0, sizeof(*out));
    if (c->p1_timed) {
        CU(cudaEventSynchronize(c->ev_p1[1]));
        CU(cudaEventElapsedTime(&out->pass1_scan_ms, c->ev_p1[0], c->ev_p1_scan));
        CU(cudaEventElapsedTime(&out->pass1_intra_ms, c->ev_p1_scan, c->ev_p1[1]));
    }
    if (c->p2_timed) {
        CU(cudaEventSynchronize(c->ev_p2[1]));
        CU(cudaEventElapsedTime(&out->pass2_ms, c->ev_p2[0], c->ev_p2[1]));
    }
    return SDM_OK;
}

int sdm_last_pack_ms(sdm_ctx* c, float* pack_ms)
{
    if (!c || !pack_ms) return fail(SDM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    *pack_ms = 0.f;
    for (int i = 0; i + 1 < c->n_pack_ev; i += 2) {
        float ms = 0.f;
        CU(cudaEventSynchronize(c->ev_pack[i + 1]));
        CU(cudaEventElapsedTime(&ms, c->ev_pack[i], c->ev_pack[i + 1]));
        *pack_ms += ms;
    }
    return SDM_OK;
}

int sdm_mark(sdm_ctx* c, int idx)
{
    if (!c) return fail(SDM_ERR_ARG, "null context");
    if (idx < 0 || idx >= SDM_N_MARKS) return fail(SDM_ERR_ARG, "mark %d out of [0,%d)", idx, SDM_N_MARKS);
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaEventRecord(c->marks[idx], c->s_compute));
    c->mark_set[idx] = true;
    return SDM_OK;
}

int sdm_elapsed_ms(sdm_ctx* c, int from, int to, float* ms)
{
    if (!c || !ms) return fail(SDM_ERR_ARG, "null argument");
    if (from < 0 || from >= SDM_N_MARKS || to < 0 || to >= SDM_N_MARKS) return fail(SDM_ERR_ARG, "mark out of range");
    if (!c->mark_set[from] || !c->mark_set[to]) return fail(SDM_ERR_STATE, "mark not recorded");
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaEventSynchronize(c->marks[to]));
    CU(cudaEventElapsedTime(ms, c->marks[from], c->marks[to]));
    return SDM_OK;
}

}  // extern "C"
