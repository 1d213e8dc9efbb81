/*
 * sdm_b200.h — C-ABI of the B200-native semi-dense mapping library (libsdm_b200.so).
 *
 * This is the drop-in boundary for EAO-SLAM's ProbabilityMapping hot path.  The reference has no
 * FFI layer: ProbabilityMapping (include/ProbabilityMapping.h:71-158) is a concrete C++ class
 * that reads/writes public cv::Mat members of ORB_SLAM2::KeyFrame (include/KeyFrame.h:155-175).
 * Each entry point below names the reference code it replaces; the C++ class shim that keeps the
 * reference's method names on top of these calls is eao-slam_b200/host/ProbabilityMapping.h, and
 * INTEGRATION.md shows the binding a maintainer adds to the reference.
 *
 * Conventions: extern "C", plain pointers and sizes, no C++/torch types.  Every function returns
 * 0 on success or a negative sdm_status; sdm_last_error() gives the message of the last failure
 * on the calling thread.  One context per CUDA device; a context is NOT thread-safe (the
 * reference runs the whole path on its single semi-dense thread, System.cc:124).  There is no CPU
 * fallback: without a CUDA device sdm_create fails with SDM_ERR_CUDA.
 *
 * All citations are relative to the reference root (yanmin-wu/EAO-SLAM).
 */
#ifndef SDM_B200_H
#define SDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDM_MAX_NBR 16 /* covisN is 7 in the reference (ProbabilityMapping.h:45); BASELINE uses 6 and 10 */

typedef enum {
    SDM_OK = 0,
    SDM_ERR_ARG = -1,   /* bad argument (null pointer, slot out of range, N > SDM_MAX_NBR, ...) */
    SDM_ERR_CUDA = -2,  /* CUDA runtime error, no device */
    SDM_ERR_STATE = -3, /* keyframe not uploaded, pass 2 before pass 1, ... */
    SDM_ERR_NOMEM = -4
} sdm_status;

typedef struct sdm_ctx sdm_ctx;

/* The #defines of include/ProbabilityMapping.h:45-56 and the literals of
 * src/ProbabilityMapping.cc (:757 slope, :1638 5.99, :1207 3.84, :877 0.000001), made runtime. */
typedef struct {
    int width, height;   /* im_.cols, im_.rows: one size per context                          */
    int max_keyframes;   /* device slots to reserve (owned + halo keyframes of this rank)      */
    int lambdaG;         /* 8   */
    int lambdaL;         /* 80  */
    int lambdaTheta;     /* 45  */
    int lambdaN;         /* 3   */
    float theta;         /* (float)0.23 = THETA */
    float sigmaI;        /* 20 = I_stddev, KeyFrame.cc:65 */
    double chi2_fusion;  /* 5.99 */
    double chi2_inter;   /* 3.84 */
    double eps;          /* 0.000001 */
    float slope_max;     /* 4 */
    int intra_check;     /* 0: shipped loop (:491-494 commented out); 1: IntraKeyFrameDepthChecking on */
    int intra_grow;      /* same for IntraKeyFrameDepthGrowing                                        */
    int device;          /* CUDA device ordinal */
} sdm_config;

/* One keyframe's work order for a pass: the per-keyframe state SemiDenseLoop assembles at
 * :365-438 (closestMatches, rotIs, min/max depth).  kf and nbr[] are device slots. */
typedef struct {
    int32_t kf;
    int32_t n_nbr;
    int32_t nbr[SDM_MAX_NBR];
    float rot_deg[SDM_MAX_NBR]; /* rotIs[kf2], :406-415 */
    float min_depth, max_depth; /* StereoSearchConstraints, :427 */
} sdm_item;

typedef struct {
    float R21[9];
    float t21[3];
    float F12[9];
} sdm_pair_geometry_t;

/* replaces: depthHo (ProbabilityMapping.h:74-80) as returned by EpipolarSearch */
typedef struct {
    float depth; /* inverse depth */
    float sigma;
    int32_t supported;
    float best_u, best_v;
} sdm_hypothesis;

/* counters of the last pass (device-side atomics; optional diagnostics) */
typedef struct {
    long long candidates; /* pixels passing :454-456 in the keyframes of the last sdm_pass1 */
    long long fused;      /* pixels written at :483                                          */
    long long checked;    /* pixels with depth_map_checked_ > 0 after the last sdm_pass2      */
} sdm_stats;

/* ---- context ------------------------------------------------------------------------------- */
void sdm_default_config(sdm_config* cfg);            /* ProbabilityMapping.h:45-56 defaults */
int sdm_create(const sdm_config* cfg, sdm_ctx** out); /* replaces ProbabilityMapping ctor, :195-202 */
void sdm_destroy(sdm_ctx* ctx);
const char* sdm_last_error(void);
const char* sdm_version(void);
int sdm_synchronize(sdm_ctx* ctx);
int sdm_get_stats(sdm_ctx* ctx, sdm_stats* out);
/* which column loop sdm_pass1 runs for keyframes with regular planes: 3 = third generation (the second generation's
 * arithmetic, but a lane jumps over columns that its skip-distance planes rule out), 2 = second generation (reference
 * thresholds lambdaL = 80 / lambdaTheta = 45 and the reciprocal form of x / theta verified for this theta at
 * sdm_create), 1 = first generation (any thresholds; also used per keyframe when an orientation plane holds values
 * outside [0, 360] or |rot| > 360).  env SDM_SCAN=lane1 | lane2 | lane3 forces one; SDM_SCAN=warp selects the
 * warp-per-pixel kernel of the survey's plan (kept for A/B, 2-4x slower; reported as 0).  Results are bit-identical. */
int sdm_scan_generation(sdm_ctx* ctx);
/* 1 if the last sdm_pass1 ran the scan kernel's build for long scans (mean search range of the batch above 128 columns:
 * 64 registers / 8 blocks per SM instead of 48 / 10; same arithmetic, same results); env SDM_SCAN_LONG=0|1 forces it */
int sdm_last_scan_long(sdm_ctx* ctx);

/* pinned host memory for asynchronous uploads/downloads (optional; any host pointer is accepted) */
int sdm_host_alloc(void** ptr, size_t bytes);
int sdm_host_free(void* ptr);

/* ---- keyframe planes ----------------------------------------------------------------------- */
/* replaces: the KeyFrame ctor's plane set-up (KeyFrame.cc:63-88) + mEdgeIndex from
 * LineDetector::DetectEdgeMap (LineDetector.cc:843-881), as device-resident packed buffers, and the
 * candidate test of :454-456 (compaction).  Steps are in BYTES (cv::Mat::step).  edge may be NULL
 * (every pixel passes :454).  K = {fx, fy, cx, cy}; Tcw = rows 0..2 of the 4x4 pose, row-major.
 * grad and theta may BOTH be NULL: GradImg / GradTheta are then produced on the device from im (SURVEY.md
 * 8f-1; replaces KeyFrame.cc:69-74: cv::Scharr x2 with scale 1/32 [exact], magnitude = sqrtf(gx*gx + gy*gy),
 * phase = the scalar cv::fastAtan2 in degrees).  OpenCV's SIMD magnitude / phase differ from these scalar forms
 * by <= 1 ulp / 3e-5 deg and are not reproducible between calls on a multi-threaded host, so device-produced
 * planes are as valid as a host run's but not bit-identical to a particular one; sdm_download_planes returns them.
 * Asynchronous: host buffers must stay valid until sdm_synchronize / the next blocking call. */
int sdm_upload_keyframe(sdm_ctx* ctx, int kf,
                        const uint8_t* im, size_t im_step,
                        const float* grad, size_t grad_step,
                        const float* theta, size_t theta_step,
                        const int32_t* edge, size_t edge_step,
                        const float K[4], const float Tcw[12]);
/* the same for n keyframes in one call (one work order per keyframe): fewer host calls and one
 * stream hand-over per batch.  This is what a host loop over all keyframes should use. */
typedef struct {
    int32_t kf;
    const uint8_t* im;    size_t im_step;
    const float* grad;    size_t grad_step;
    const float* theta;   size_t theta_step;
    const int32_t* edge;  size_t edge_step; /* edge may be NULL; host plane or device plane (sdm_ed_device_edge_plane) */
    float K[4];
    float Tcw[12];
} sdm_upload_desc;
int sdm_upload_keyframes(sdm_ctx* ctx, int n, const sdm_upload_desc* desc);
/* GradImg / GradTheta of a slot as dense planes (what KeyFrame::GradImg / GradTheta hold, KeyFrame.h:160); blocking */
int sdm_download_planes(sdm_ctx* ctx, int kf, float* grad, size_t grad_step, float* theta, size_t theta_step);
/* replaces: KeyFrame::SetPose (KeyFrame.cc:108-124) for PoseChanged refresh (:691-694) */
int sdm_set_pose(sdm_ctx* ctx, int kf, const float Tcw[12]);
/* calibration of a slot (KeyFrame::fx,fy,cx,cy; Frame.cc:584-590).  With sdm_set_pose this is all a
 * halo slot needs when only its pass-1 planes arrive over NVLink (no image upload) */
int sdm_set_intrinsics(sdm_ctx* ctx, int kf, const float K[4]);
int sdm_candidate_count(sdm_ctx* ctx, int kf, int* count); /* blocking */
/* number of 16-pixel row blocks of the n keyframes that hold at least one candidate pixel: what a block-sparse download
 * (sdm_scatter_keyframes into pinned planes, sdm_loop.sparse_download) moves, 16 * 24 bytes per block for the four planes;
 * blocking */
int sdm_candidate_blocks(sdm_ctx* ctx, int n, const int32_t* kfs, uint64_t* blocks);

/* ---- the two hot loops --------------------------------------------------------------------- */
/* replaces: SemiDenseLoop pass 1 body, :424-497 — ComputeFundamental, the omp pixel loop with
 * EpipolarSearch / InverseDepthHypothesisFusion, [IntraKeyFrameDepthChecking/Growing] — for n
 * keyframes in one batched launch.  Writes depth_map_/depth_sigma_ of every item's keyframe. */
int sdm_pass1(sdm_ctx* ctx, int n, const sdm_item* items);
/* replaces: SemiDenseLoop pass 2 body, :550-552 — InterKeyFrameDepthChecking(kf, neighbours)
 * (:1121-1296) + UpdateSemiDensePointSet (:700-731).  Reads the neighbours' pass-1 planes. */
int sdm_pass2(sdm_ctx* ctx, int n, const sdm_item* items);
/* replaces: UpdateSemiDensePointSet alone (UpdateAllSemiDensePointSet, :678-697) */
int sdm_update_points(sdm_ctx* ctx, int n, const int32_t* kfs);

/* replaces: reading KeyFrame::depth_map_, depth_sigma_, depth_map_checked_, SemiDensePointSets_.
 * Any pointer may be NULL.  Steps in bytes.  Blocking. */
int sdm_download(sdm_ctx* ctx, int kf,
                 float* depth, size_t depth_step, float* sigma, size_t sigma_step,
                 float* checked, size_t checked_step, float* points, size_t points_step);
/* same, but only enqueued (on the library's download stream, ordered after the last pass that wrote the
 * slot): the host buffers are valid after the next sdm_synchronize.  Lets a caller overlap the D2H of one
 * chunk of keyframes with the passes of the next (pinned buffers from sdm_host_alloc recommended). */
int sdm_download_async(sdm_ctx* ctx, int kf,
                       float* depth, size_t depth_step, float* sigma, size_t sigma_step,
                       float* checked, size_t checked_step, float* points, size_t points_step);
/* sdm_download_async for n keyframes in one call (any plane pointer of an entry may be NULL) */
typedef struct {
    int32_t kf;
    float* depth;    size_t depth_step;
    float* sigma;    size_t sigma_step;
    float* checked;  size_t checked_step;
    float* points;   size_t points_step;
} sdm_download_desc;
int sdm_download_keyframes(sdm_ctx* ctx, int n, const sdm_download_desc* desc);
/* Sparse form of sdm_download_keyframes for planes that are ZERO-INITIALISED the way KeyFrame's constructors leave
 * them (KeyFrame.cc:78-81): SemiDenseLoop only ever writes a non-zero value at a keyframe's candidate pixels
 * (the stores of :483-484, :1290 and :725-727 all sit behind the candidate test of :454-456), so only those
 * pixels' records (28 bytes each instead of 24 bytes for every pixel) cross PCIe, and worker threads of the library
 * (env SDM_SCATTER_THREADS, default 6) write them into the caller's planes; every other element is left untouched.
 * Destination planes in PINNED host memory (sdm_host_alloc, cudaHostRegister) take a faster route with the same
 * contract: a kernel writes the 16-pixel blocks of every row that hold a candidate straight into them over PCIe (no
 * records, no host threads).
 * On zero-initialised planes the result is identical to sdm_download_keyframes.  Asynchronous: the planes are
 * complete after sdm_synchronize.  Slots whose planes were written from outside (sdm_upload_depth) are refused
 * (SDM_ERR_STATE).  Blocks the caller only until the slot's candidate count is known (its upload has been packed). */
int sdm_scatter_keyframes(sdm_ctx* ctx, int n, const sdm_download_desc* desc);
/* writes pass-1 planes of a keyframe (used to seed halo keyframes / tests); blocking */
int sdm_upload_depth(sdm_ctx* ctx, int kf, const float* depth, size_t depth_step,
                     const float* sigma, size_t sigma_step);

/* writes depth_map_checked_ of a keyframe (the pass-2 plane).  For hosts that keep the planes in their own memory and
 * have to re-create a slot (a larger arena, a reset): with sdm_upload_depth and sdm_update_points this restores
 * everything a finished keyframe's slot held (SemiDensePointSets_ is a function of this plane and the pose, :700-731);
 * blocking */
int sdm_upload_checked(sdm_ctx* ctx, int kf, const float* checked, size_t checked_step);

/* ---- point-cloud export ---------------------------------------------------- */
/* replaces: the filter loop of SaveSemiDensePoints (ProbabilityMapping.cc:159-186), MapDrawer::DrawSemiDense
 * (MapDrawer.cc:99-117) and the CARV point entry (SFMTranscriptInterface_ORBSLAM.cpp:268-290): for the n
 * keyframes in list order and their pixels in raster order, `if (depth_sigma_ > sigma_max) continue;
 * if (depth_map_checked_ > 0.000001) emit SemiDensePointSets_(y,x)`.  Compacted on the device, so only the
 * surviving points cross PCIe.  `pixel` = (y << 16) | x lets the caller fetch the colour from its own rgb_.
 * Writes min(total, capacity) points to out (out may be NULL with capacity 0 to count only), the surviving
 * count of every keyframe to counts[n] (may be NULL) and their sum to *total.  Blocking. */
typedef struct {
    float x, y, z;
    uint32_t pixel;
} sdm_point;
int sdm_export_points(sdm_ctx* ctx, int n, const int32_t* kfs, double sigma_max, sdm_point* out, size_t capacity,
                      uint64_t* counts, uint64_t* total);

/* ---- edge-aided 3-D line fitting -------------------------------------------- */
/* replaces: LineDetector::LineFitting / LineFit / LeastSquaresLineFit / LeastSquaresDepthFit / CountDepth /
 * ComputePointDistance2Line / ComputePointDepth2Line (LineDetector.cc:578-840, :884-900; called from
 * ProbabilityMapping.cc:262-268 once the loop is done) for a batch of keyframes, reading depth_map_checked_ and
 * depth_sigma_ where pass 2 left them on the device.  The edge chains are an INPUT: the reference gets them from the
 * closed-source EDLib (DetectEdgesByED, LineDetector.cc:855) on the host and keeps them in kf->mEdgeMap; a set lists
 * the pixels of EdgeMap::segments[i] in order, packed (row << 16) | col, offsets[i] .. offsets[i + 1].
 * Output = the rows the reference appends to kf->mLinesSeg (seg) and kf->mLines3D (xyz), in its order (keyframes in
 * list order, chains in order, lines along the chain); `chain` / `kf_index` say where a row came from.
 * Writes min(total, capacity) rows (out may be NULL with capacity 0 to count only), per-set counts to counts[n] (may be
 * NULL), their sum to *total.  Blocking.  The two least-squares fits are closed-form double precision where OpenCV runs a
 * float Jacobi SVD: results agree to rounding, decisions that sit exactly on a threshold may differ. */
typedef struct {
    int32_t kf;              /* slot whose planes are read (pass 2 done) */
    int32_t n_chains;
    const int32_t* offsets;  /* n_chains + 1 entries, offsets[0] = 0 */
    const uint32_t* pixels;  /* offsets[n_chains] entries */
} sdm_edge_chains;
typedef struct {
    float seg[4];   /* s.x s.y e.x e.y */
    float xyz[6];   /* Pws, Pwe */
    int32_t chain, kf_index;
} sdm_line3d;
int sdm_line_fit(sdm_ctx* ctx, int n, const sdm_edge_chains* sets, sdm_line3d* out, size_t capacity, uint64_t* counts,
                 uint64_t* total);
/* device time of the last sdm_line_fit (its three kernels), ms */
int sdm_last_line_fit_ms(sdm_ctx* ctx, float* ms);

/* ---- Edge Drawing: the candidate mask and the edge chains --------------- */
/* replaces: LineDetector::DetectEdgeMap (LineDetector.cc:843-881), which the reference calls per keyframe inside pass 1
 * (ProbabilityMapping.cc:394) and which hands the image to the closed-source EDLib.a:
 *     EdgeMap* map = DetectEdgesByED(srcImg, width, height, SOBEL_OPERATOR, 36, 8, 1.0);                 (:855)
 *     kf->mEdgeIndex.at<int>(r, c) = i for every pixel of map->segments[i];  kf->mEdgeMap = map;          (:857-869)
 * for a batch of keyframes.  Stage 1 (smoothing, Sobel gradient, direction, the routing choice of every edge pixel,
 * anchors: k_ed_planes4 / k_ed_planes; the anchors in walking order: k_ed_sort) runs on the device in chunks of keyframes,
 * stage 2 (the sequential smart-routing walk from the anchors, host/edge_drawing.h) on n_threads host threads while the
 * next chunk is on the device - or on the device as well, see sdm_set_edge_drawing_route.  The chains are the library's, pixel for pixel and in order
 * (tests/test_edge_drawing.py, tests/test_gpu_edge_drawing.py).
 * images[i].im = kf->im_ (8-bit, ctx width x height, row pitch im_step bytes); images[i].edge_index, if not NULL,
 * receives kf->mEdgeIndex: -1, then the chain number of every chain pixel (int32 plane, row pitch edge_step bytes) -
 * the `edge` plane of sdm_upload_keyframes.  *result owns the chains of the batch until sdm_ed_free; sdm_ed_chains
 * returns keyframe i's lists in the layout sdm_line_fit takes (sdm_edge_chains.offsets / .pixels).
 * grad_thresh in 1 .. 2047 (the reference: 36), anchor_thresh >= 0 (8), n_threads <= 0: one per host core, at most 32.
 * Blocking. */
typedef struct {
    const uint8_t* im;    size_t im_step;
    int32_t* edge_index;  size_t edge_step;  /* edge_index may be NULL */
} sdm_ed_image;
typedef struct sdm_ed_result sdm_ed_result;
int sdm_edge_drawing(sdm_ctx* ctx, int n, const sdm_ed_image* images, int grad_thresh, int anchor_thresh, int n_threads,
                     sdm_ed_result** result);
int sdm_ed_chains(const sdm_ed_result* result, int i, int32_t* n_chains, const int32_t** offsets, const uint32_t** pixels);
void sdm_ed_free(sdm_ed_result* result);
/* where stage 2 (the routing walk) of sdm_edge_drawing runs.  SDM_ED_ROUTE_HOST (default): host threads, overlapped with
 * stage 1 of the following chunks - the fast choice for a few hundred keyframes on one GPU (0.12 ms per VGA keyframe on 16
 * cores).  SDM_ED_ROUTE_DEVICE: k_ed_route, one warp per keyframe, up to 1024 keyframes per launch - the walk is sequential
 * per image and a device thread walks ~15 x slower than a host core, but thousands of images walk at once and the host
 * cores stay free (one process per GPU shares them); chains and edge index are identical in both modes.  An image whose
 * walk exceeds the kernel's fixed capacities is routed on the host (sdm_last_edge_drawing_fallbacks counts them). */
/* SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE: the walks on host threads as in SDM_ED_ROUTE_HOST, but kf->mEdgeIndex is (also) built on
 * the device from the chain lists (k_ed_mask) for sdm_ed_device_edge_plane: with images[i].edge_index = NULL the host threads
 * skip filling and scattering the 4-byte planes and the masks reach sdm_upload_keyframes without crossing PCIe as planes. */
enum { SDM_ED_ROUTE_HOST = 0, SDM_ED_ROUTE_DEVICE = 1, SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE = 2 };
int sdm_set_edge_drawing_route(sdm_ctx* ctx, int mode);
int sdm_last_edge_drawing_fallbacks(sdm_ctx* ctx);
/* SDM_ED_ROUTE_DEVICE and SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE keep kf->mEdgeIndex of every image of their last batch on the device (dense int32 plane, row pitch
 * 4 * width): *dev_plane is a DEVICE pointer that sdm_upload_desc.edge / sdm_upload_keyframe's edge accept like a host
 * plane (edge_step = 4 * width), so the candidate mask of :454 goes from the detector to the packing kernel without
 * crossing PCIe (images[i].edge_index may then be NULL).  Valid until the next sdm_edge_drawing or sdm_destroy on this
 * context; calls with more than 1024 images (256 with the walks on host threads) keep the planes of the last batch only. */
int sdm_ed_device_edge_plane(sdm_ctx* ctx, int i, const int32_t** dev_plane);
/* timing of the last sdm_edge_drawing: device time of its k_ed_planes launches, host wall time of the call, and the
 * summed thread time of the routing walks - in SDM_ED_ROUTE_DEVICE mode the device time of k_ed_sort + k_ed_route (all
 * ms; any pointer may be NULL) */
int sdm_last_edge_drawing_ms(sdm_ctx* ctx, float* kernel_ms, float* wall_ms, float* route_thread_ms);
/* the stage-1 planes of one image as the device computes them (G int16, F uint8, dense width x height; see
 * csrc/edge_drawing_kernels.cuh) - for tests and tools */
int sdm_ed_planes(sdm_ctx* ctx, const uint8_t* im, size_t im_step, int grad_thresh, int anchor_thresh, int16_t* G, uint8_t* F);

/* ---- multi-GPU: pass-1 planes of halo keyframes over NVLink (the dependency of :1202-1249) -- */
/* device pointer + byte size of the (rho, sigma) float2 plane of a slot, for NCCL / peer copies */
int sdm_depth_plane_ptr(sdm_ctx* ctx, int kf, void** dev_ptr, size_t* bytes);
/* CUDA IPC handle (64 bytes) of the whole (rho, sigma) arena, to be opened by a peer process */
int sdm_export_arena(sdm_ctx* ctx, void* handle64, size_t* slot_bytes);
/* open a peer's arena; afterwards sdm_pull_halo copies peer slots into local slots over NVLink */
int sdm_import_peer_arena(sdm_ctx* ctx, int peer_rank, const void* handle64);
int sdm_pull_halo(sdm_ctx* ctx, int n, const int32_t* local_slot, const int32_t* peer_rank,
                  const int32_t* peer_slot);
/* mark a slot's pass-1 planes as valid (after an external NCCL receive into sdm_depth_plane_ptr) */
int sdm_mark_pass1_done(sdm_ctx* ctx, int kf);
/* PROTOCOL of the three calls above (sdm_export_arena / sdm_import_peer_arena / sdm_pull_halo): they carry no ordering
 * between ranks.  The caller must (1) sdm_synchronize() + a cross-rank barrier after its sdm_pass1 and before anybody's
 * sdm_pull_halo, and (2) sdm_synchronize() + a barrier after the pulls and before the next sdm_pass1 / upload on an
 * owning rank; peer_slot is not range-checked (the peer's slot count is unknown).  sdm_exchange below needs neither. */

/* ---- multi-GPU exchange ordered on the devices (SURVEY.md 8b `sdm_exchange`, 8e-ii) ------------------------------
 * The only cross-keyframe dependency of the path is "pass 1 of all neighbours before pass 2 of a keyframe"
 * (ProbabilityMapping.cc:1202-1249).  With keyframes sharded over GPUs, each rank pulls the (rho, sigma) planes of its
 * halo keyframes from their owners between the passes.  One-time set-up per rank:
 *     sdm_export_peer_handle(ctx, my_rank, &h)   -> all-gather the handles by any transport
 *     sdm_import_peer(ctx, r, &handle_of_r)      for every rank r this rank pulls from
 *     sdm_set_halo(ctx, n, local_slot, peer_rank, peer_slot)   (peer_slot is validated against the peer's slot count)
 *     [one host barrier: every import done before the first step]
 * Per step, on every rank, with NO host synchronisation or barrier:
 *     sdm_pass1(...); sdm_exchange(ctx); sdm_pass2(...);
 * sdm_exchange publishes "my pass-1 planes of step s are complete" with a kernel behind pass 1, and on a separate
 * stream waits (device-side spin on the owner's flag over NVLink) for each owner, copies the halo planes peer-to-peer
 * and acknowledges to the owner.  sdm_pass2 launches the keyframes that do not read a halo plane at once and the
 * others behind the pull (the pull overlaps pass 2 of the interior keyframes); the next sdm_pass1 waits on the device
 * for the acknowledgements of every rank that pulls from this one before it overwrites the planes.  All ranks must call
 * sdm_exchange the same number of times.  A wait that exceeds 20 s flags an error that the next sdm_synchronize
 * returns as SDM_ERR_STATE.  With an empty halo plan (single GPU) sdm_exchange is a no-op. */
#define SDM_PEER_HANDLE_BYTES 192
typedef struct {
    unsigned char bytes[SDM_PEER_HANDLE_BYTES]; /* IPC handles of the (rho,sigma) arena and the flag block, slot count, rank */
} sdm_peer_handle;
int sdm_export_peer_handle(sdm_ctx* ctx, int my_rank, sdm_peer_handle* out);
int sdm_import_peer(sdm_ctx* ctx, int peer_rank, const sdm_peer_handle* handle);
int sdm_set_halo(sdm_ctx* ctx, int n, const int32_t* local_slot, const int32_t* peer_rank, const int32_t* peer_slot);
int sdm_exchange(sdm_ctx* ctx);

/* ---- one whole SemiDenseLoop as a pipeline (ProbabilityMapping.cc:348-597) --------------------------------------
 * Issues uploads, pass 1, the downloads of depth_map_/depth_sigma_, [the exchange], pass 2 and the downloads of
 * depth_map_checked_/SemiDensePointSets_ in chunks of `chunk` keyframes, in the order that keeps the copy engines and
 * the SMs busy at the same time: the planes of chunk k+1 go up while pass 1 of chunk k runs, a chunk's pass-1 planes
 * leave as soon as its pass 1 is queued, and a keyframe's pass 2 is queued as soon as pass 1 of all its neighbours is.
 * pass1[] and pass2[] are the work orders of the two passes in processing order (they differ when the gating of
 * :365-384 and :523-542 differs); down1[i] / down2[i] receive the planes of pass1[i].kf / pass2[i].kf (either array
 * may be NULL, as may any plane pointer).  upload[] lists the keyframes whose planes are not resident yet, in the
 * order of first use.  Asynchronous like its parts: host buffers are valid after sdm_synchronize. */
typedef struct {
    int32_t n_upload;
    const sdm_upload_desc* upload;
    int32_t n_pass1;
    const sdm_item* pass1;
    const sdm_download_desc* down1;
    int32_t n_pass2;
    const sdm_item* pass2;
    const sdm_download_desc* down2;
    int32_t chunk;    /* keyframes per pipeline chunk; <= 0: 4 (measured optimum on B200 + PCIe 5) */
    int32_t exchange; /* != 0: sdm_exchange between the passes; pass-2 work orders that read halo planes wait for it */
    int32_t sparse_download; /* != 0: the destination planes are ZERO-INITIALISED the way KeyFrame's constructor leaves them
                              * (KeyFrame.cc:78-81) - the contract of sdm_scatter_keyframes.  Planes in pinned host memory then
                              * receive only the 16-pixel blocks that hold a candidate pixel, written by a kernel over PCIe
                              * (about half the bytes of the dense DMA on textured scenes, far less with an edge mask);
                              * pageable planes get the dense DMA.  Identical host planes either way.
                              * 2: the same contract for SemiDensePointSets_ only (61 % of the dense bytes): the three 4-byte
                              * planes leave by DMA while a kernel on a second stream writes the point blocks, so the copy
                              * engine and the SM-issued writes share the link. */
} sdm_loop;
int sdm_run_loop(sdm_ctx* ctx, const sdm_loop* loop);

/* ---- per-method entry points (each named class method stays individually callable) ---------- */
/* replaces: ComputeFundamental (:1694-1709) and the R21/t21 of :1136-1137; host arithmetic with
 * OpenCV's evaluation rules */
int sdm_pair_geometry(const float K1[4], const float Tcw1[12], const float K2[4], const float Tcw2[12],
                      sdm_pair_geometry_t* out);
/* replaces: StereoSearchConstraints (:734-747); inv_depths = KeyFrame::GetAllPointDepths */
int sdm_stereo_search_constraints(const float* inv_depths, int n, float* min_depth, float* max_depth);
/* the chi-square gate of InterKeyFrameDepthChecking alone (:1204-1211 and its three copies): accept[i] = 1 iff
 * (float)(dd*dd / (sigma*sigma)) < 3.84 with dd = (double)diff[i], diff = depthj - depth_map_(neighbour) as a float */
int sdm_inter_chi_test(sdm_ctx* ctx, int n, const float* diff, const float* sigma, uint8_t* accept);
/* replaces: GetSearchRange (:1598-1631) */
int sdm_search_range(sdm_ctx* ctx, int kf1, int kf2, int px, int py, float mind, float maxd,
                     float* umin, float* umax);
/* replaces: EpipolarSearch (:749-845) for one pixel / one neighbour, same argument meaning and order as
 * the reference (kf1, kf2, x, y, pixel, min_depth, max_depth, dh, [F12: recomputed from the poses],
 * best_u, best_v, th_pi, rot): `pixel` = kf1->im_(y,x) and `th_pi` = kf1->GradTheta(y,x) at the
 * reference's call site (:457, :466-470) */
int sdm_epipolar_search(sdm_ctx* ctx, int kf1, int kf2, int x, int y, float pixel, float min_depth, float max_depth,
                        float th_pi, float rot_deg, sdm_hypothesis* out);
/* EpipolarSearch over every candidate pixel of kf1 against kf2; planes are dense W*H; ok: 0 = no
 * hypothesis, 1 = hypothesis failing the keep test of :472, 2 = kept */
int sdm_epipolar_search_plane(sdm_ctx* ctx, int kf1, int kf2, float min_depth, float max_depth, float rot_deg,
                              float* hyp_depth, float* hyp_sigma, float* hyp_u, uint8_t* ok);
/* replaces: InverseDepthHypothesisFusion (:978-1009) for m independent hypothesis sets of n each
 * (n <= SDM_MAX_NBR; depth/sigma are [m][n]; count[m] valid entries per set) */
int sdm_fuse(sdm_ctx* ctx, int m, int n, const float* depth, const float* sigma, const int32_t* count,
             float* out_depth, float* out_sigma, int32_t* out_supported);
/* replaces: IntraKeyFrameDepthChecking(cv::Mat&, cv::Mat&, cv::Mat) (:866-927) on a keyframe's planes */
int sdm_intra_check(sdm_ctx* ctx, int kf);
/* replaces: IntraKeyFrameDepthGrowing (:929-976) */
int sdm_intra_grow(sdm_ctx* ctx, int kf);
/* replaces: InterKeyFrameDepthChecking(KeyFrame*, vector<KeyFrame*>) (:1121-1296) alone */
int sdm_inter_check(sdm_ctx* ctx, const sdm_item* item);

/* ---- measurement --------------------------------------------------------------------------- */
/* CUDA-event time (ms) of the kernels of the last sdm_pass1 / sdm_pass2 call on the context's
 * compute stream (0 if none); number of kernel launches issued since context creation */
int sdm_last_pass_ms(sdm_ctx* ctx, float* pass1_ms, float* pass2_ms);
long long sdm_launch_count(sdm_ctx* ctx);
/* per-kernel split of the same: the epipolar scan + fusion kernel, the optional intra stencils, pass 2
 * (the per-stage "took ... ms" prints of ProbabilityMapping.cc:389-443, :505-508, :545-565) */
typedef struct {
    float pass1_scan_ms;
    float pass1_intra_ms;
    float pass2_ms;
} sdm_timing;
int sdm_last_timing(sdm_ctx* ctx, sdm_timing* out);
/* device time (ms) of the packing / candidate-compaction kernels (k_pack | k_pack_image, k_skip: SURVEY.md 8d counts the
 * compaction as part of the device-resident loop) of the last sdm_upload_keyframes call, H2D copies excluded */
int sdm_last_pack_ms(sdm_ctx* ctx, float* pack_ms);
/* user marks on the context's compute stream (CUDA events), for timing a whole SemiDenseLoop the way
 * :246-254 does with clock_gettime: sdm_mark(idx) records, sdm_elapsed_ms blocks until `to` completed */
#define SDM_N_MARKS 8
int sdm_mark(sdm_ctx* ctx, int idx);
int sdm_elapsed_ms(sdm_ctx* ctx, int from, int to, float* ms);

#ifdef __cplusplus
}
#endif
#endif
