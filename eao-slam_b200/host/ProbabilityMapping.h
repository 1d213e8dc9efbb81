// ProbabilityMapping.h — host-side drop-in for EAO-SLAM's semi-dense mapping class, on top of the
// C-ABI of libsdm_b200.so (include/sdm_b200.h).
//
// Mirrors include/ProbabilityMapping.h:71-158 of yanmin-wu/EAO-SLAM: same class name, same method
// names / argument order / meaning (SemiDenseLoop, StereoSearchConstraints, EpipolarSearch,
// GetSearchRange, InverseDepthHypothesisFusion, IntraKeyFrameDepthChecking/Growing,
// InterKeyFrameDepthChecking, UpdateSemiDensePointSet, the finish / reset handshakes), same error
// behaviour (void methods, ineligible keyframes silently skipped, problems to std::cerr), but the
// pixel loops run as CUDA kernels and the keyframes of a pass are batched into one launch.
// Header-only, C++11 (the reference's dialect, CMakeLists.txt:22).
//
// Two build modes:
//   * inside EAO-SLAM: define SDM_HOST_WITH_ORBSLAM2 before including; KeyFrame / Map / cv::Mat are
//     the reference's own types (INTEGRATION.md shows the three-line change to System.cc); the class then also owns
//     the Modeler and the LineDetector like the reference's (:195-202) and drives them where the reference does
//     (edge map + line segments per keyframe inside pass 1, :394-397; the consumers after the loop, :256-297).
//     tests/test_shim_orbslam2_mode.py compiles this mode against the stand-in headers of oracle/refshim/;
//   * stand-alone (this repository, no OpenCV available): host_types.h supplies minimal stand-ins
//     with the same member names, which is what tests/cpp/test_shim.cpp compiles against.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <sys/stat.h>

#include "../../include/sdm_b200.h"

#ifdef SDM_HOST_WITH_ORBSLAM2
#include <opencv2/core/core.hpp>
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "Modeler.h"        // the consumers the reference's class owns / drives after the loop (:195-202, :256-297):
#include "LineDetector.h"   // unchanged reference code, called at the same places
namespace sdm_host {
typedef cv::Mat Mat;
typedef ORB_SLAM2::KeyFrame KeyFrame;
typedef ORB_SLAM2::Map Map;
inline Mat zeros32f(int rows, int cols) { return cv::Mat::zeros(rows, cols, CV_32F); }
inline Mat plane32s(int rows, int cols) { return cv::Mat(rows, cols, CV_32S); }
}  // namespace sdm_host
#else
#include "host_types.h"
namespace sdm_host {
inline Mat plane32s(int rows, int cols) { return Mat(rows, cols, sizeof(int32_t)); }
}
#endif
#include "edge_drawing.h"

namespace sdm_host {
// ProbabilityMapping::GetRotInPlane (:847-864): angle2 - angle1 of the ORB keypoints of every map point
// both keyframes observe.  A hash join replaces the reference's O(n1*n2) pointer scan; the caller only
// uses the sorted multiset (lower median, :408-413), which is identical.
template <class KF>
std::vector<float> RotInPlane(KF* kf1, KF* kf2)
{
    std::vector<float> rot;
    const auto vMPs1 = kf1->GetMapPointMatches();
    const auto vMPs2 = kf2->GetMapPointMatches();
    const auto kp1 = kf1->GetKeyPointsUn();
    const auto kp2 = kf2->GetKeyPointsUn();
    std::unordered_multimap<const void*, size_t> where2;
    for (size_t i2 = 0; i2 < vMPs2.size(); i2++)
        if (vMPs2[i2]) where2.insert(std::make_pair((const void*)vMPs2[i2], i2));
    for (size_t i1 = 0; i1 < vMPs1.size(); i1++) {
        if (!vMPs1[i1]) continue;
        auto range = where2.equal_range((const void*)vMPs1[i1]);
        for (auto it = range.first; it != range.second; ++it) {
            const float angle1 = kp1[i1].angle, angle2 = kp2[it->second].angle;
            if (angle1 < 0 || angle2 < 0) continue;
            rot.push_back(angle2 - angle1);
        }
    }
    return rot;
}
}  // namespace sdm_host


#ifndef covisN
#define covisN 7   // ProbabilityMapping.h:45 (runtime-overridable through SetCovisN for BASELINE's 6 / 10)
#endif

class ProbabilityMapping {
public:
    typedef sdm_host::KeyFrame KeyFrame;
    typedef sdm_host::Map Map;
    typedef sdm_host::Mat Mat;

    // ProbabilityMapping.h:74-80 (Pw is unused on the path)
    struct depthHo {
        depthHo() : depth(0.0f), sigma(0.0f), supported(false) {}
        float depth;
        float sigma;
        bool supported;
    };

    explicit ProbabilityMapping(Map* pMap)
        : mMutexSemiDense(), mpMap(pMap), mCtx(NULL), mN(covisN), mW(0), mH(0), mCapacity(0), mChunk(0), mHeadroom(-1), mDevicePlanes(false), mSparse(false),
          mOnline(false), mEdgeDrawing(false), mEdRouteDevice(false), mEdThreads(0), mEdGrad(36), mEdAnchor(8), mbFinishRequested(false), mbFinished(false), mbResetRequested(false)
    {
        sdm_default_config(&mCfg);
#ifdef SDM_HOST_WITH_ORBSLAM2
        mpModeler = new Modeler(mpMap);  // :195-202
        mpMap->SetModeler(mpModeler);
#endif
    }
    ~ProbabilityMapping() { sdm_destroy(mCtx); }

    // run-time versions of the #defines of ProbabilityMapping.h:45-56; call before the first loop
    sdm_config& Config() { return mCfg; }
    void SetCovisN(int n) { mN = n; }
    // upload im_ only and let the device produce GradImg / GradTheta (KeyFrame.cc:69-74: scalar-form magnitude / phase,
    // <= 1 ulp / 3e-5 deg from OpenCV's SIMD kernels) instead of uploading the keyframe's own planes; 1 B/px instead of 9
    void SetProducePlanesOnDevice(bool on) { mDevicePlanes = on; }
    // keyframes per chunk of the pipelined loop (sdm_run_loop); 0 = the library's default
    void SetPipelineChunk(int n) { mChunk = n; }
    // Opt-in: results come back block-sparse.  KeyFrame's constructor zero-initialises depth_map_, depth_sigma_,
    // depth_map_checked_ and SemiDensePointSets_ (KeyFrame.cc:78-81) and the loop writes a keyframe's planes once, so only
    // the 16-pixel blocks that hold a candidate pixel have to cross PCIe (planes in pinned memory; pageable planes get the
    // dense DMA either way).  Halves the bytes on textured scenes but is written by a kernel, whose PCIe writes run at
    // about half the copy engines' rate: measured slower than the dense DMA at 23 % candidate density (38.1 vs 34.9 ms per
    // 200 keyframes), a gain only for sparse candidate sets (edge masks).  Identical planes either way.
    void SetSparseDownloads(bool on) { mSparse = on; }
    // device slots reserved beyond the keyframes of the map when an arena is created; < 0 = half the map + 16.  A slot
    // costs 78-86 bytes per pixel, a larger arena is created (and finished keyframes re-seeded) when the map outgrows it
    void SetArenaHeadroom(int slots) { mHeadroom = slots; }
    // What the reference does per keyframe right before its pixel loop (:394-397): mLineDetector.DetectEdgeMap(kf) fills
    // kf->mEdgeIndex (the candidate mask of :454), DetectLineSegments(kf) feeds the line fitting.  With
    // SDM_HOST_WITH_ORBSLAM2 the class calls its own mLineDetector exactly there; a hook replaces that call (stand-alone
    // builds have no LineDetector: without a hook mEdgeIndex is used as the caller left it, empty = every pixel passes).
    void SetEdgeMapHook(const std::function<void(KeyFrame*)>& f) { mEdgeHook = f; }
    // The open Edge Drawing implementation instead of the closed-source EDLib.a behind LineDetector::DetectEdgeMap
    // (LineDetector.cc:843-881): the edge maps of ALL keyframes pass 1 is about to process in one sdm_edge_drawing call
    // (smoothing / gradient / anchors on the device, the routing walks on `threads` host threads) right where the
    // reference detects them one by one (:394).  Fills kf->mEdgeIndex (allocated if empty) exactly as :857-866 does - the
    // chains are the library's, pixel for pixel (tests/test_edge_drawing.py) - and keeps the chains for FitLines
    // (EdgeChainsOf); with SDM_HOST_WITH_ORBSLAM2 kf->mEdgeMap is built from them as well (:869), so the reference's own
    // consumers of the edge map run unchanged.  Off by default: a hook, then the reference's LineDetector, come first.
    // route_on_device: the routing walks in k_ed_route (one warp per keyframe) instead of host threads - for maps of a
    // thousand keyframes and more, or when the host cores are shared by several GPU processes (sdm_set_edge_drawing_route).
    void SetEdgeDrawing(bool on, int threads = 0, int grad_thresh = 36, int anchor_thresh = 8, bool route_on_device = false)
    {
        mEdgeDrawing = on; mEdThreads = threads; mEdGrad = grad_thresh; mEdAnchor = anchor_thresh; mEdRouteDevice = route_on_device;
    }
    const sdm_host::EdgeChains* EdgeChainsOf(KeyFrame* kf) const
    {
        std::unordered_map<KeyFrame*, sdm_host::EdgeChains>::const_iterator it = mEdgeChains.find(kf);
        return it == mEdgeChains.end() ? NULL : &it->second;
    }
    // wall / device / routing-thread time of the last batch (ms)
    void LastEdgeDrawingMs(float* kernel_ms, float* wall_ms, float* route_thread_ms)
    {
        if (mCtx) sdm_last_edge_drawing_ms(mCtx, kernel_ms, wall_ms, route_thread_ms);
    }
    // where SaveSemiDensePoints / WriteModel put their files; default "results_line_segments/<date-time>" (:138, :121)
    void SetResultsDir(const std::string& d) { mResultsDir = d; }

    // ProbabilityMapping.cc:204-300.  Offline mode (the reference's build, `#define OnlineLoop` commented out at
    // ProbabilityMapping.cc:42): idle until finish is requested, then one loop and the consumers of :256-297.  Online
    // mode (:223-234): every 5 ms the keyframes that became eligible are processed (the gating of SemiDenseLoop skips
    // finished ones, their planes stay resident on the device) and the point sets of keyframes whose pose changed are
    // refreshed.
    void SetOnline(bool on) { mOnline = on; }
    void Run()
    {
        while (true) {
            if (CheckFinish()) break;
            if (mOnline) {
                SemiDenseLoop();
                UpdateAllSemiDensePointSet();  // make point position dependent to kf position (:226)
#ifdef SDM_HOST_WITH_ORBSLAM2
                if (mpModeler->CheckNewTranscriptEntry()) { mpModeler->RunRemainder(); mpModeler->UpdateModel(); }
#endif
            }
            ResetIfRequested();
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
        const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        SemiDenseLoop();
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const size_t nkf = mpMap->GetAllKeyFrames().size();
        std::cout << "semi dense mapping took total: " << ms << "ms  avg:" << (nkf ? ms / nkf : 0.0) << "ms  #KF:" << nkf << std::endl;  // :254
        SaveSemiDensePoints();  // :256
#ifdef SDM_HOST_WITH_ORBSLAM2
        {   // the line / surface consumers of :258-295, unchanged reference code
            std::vector<KeyFrame*> vpKFs = mpMap->GetAllKeyFrames();
            mLineDetector.RunLine3Dpp(vpKFs);
            mLineDetector.LineFittingEDLinesOffline(vpKFs);
            if (!mOnline) {
                mLineDetector.LineFittingOffline(vpKFs, mpModeler);
                for (size_t i = 0; i < vpKFs.size(); i++) mpModeler->AddLineSegmentKeyFrameEntry(vpKFs[i]);
            }
            mLineDetector.SaveAllLineSegments();
            mLineDetector.SaveClusteredSegments();
            if (mpModeler->CheckNewTranscriptEntry()) { mpModeler->RunOnce(); mpModeler->UpdateModel(); }
            mLineDetector.Summary();
        }
#endif
        SetFinish();
    }

    // ProbabilityMapping.cc:348-597.  Same gating, same two passes.  Pass 1 of a keyframe reads only immutable inputs of
    // its neighbours and pass 2 only pass-1 planes (:1202-1249), and which keyframes pass 2 will process follows from
    // the flags pass 1 is going to set (:497), so both gating loops run first, on the host, and the device work of the
    // whole loop - uploads, pass 1, depth_map_/depth_sigma_ back, pass 2, depth_map_checked_/SemiDensePointSets_ back -
    // goes to the library as ONE pipelined call (sdm_run_loop) whose copies overlap its kernels.
    void SemiDenseLoop()
    {
        std::unique_lock<std::mutex> lock(mMutexSemiDense);
        std::vector<KeyFrame*> vpKFs = mpMap->GetAllKeyFrames();
        if (vpKFs.size() < 10) return;  // :351
        if (!EnsureContext(vpKFs)) return;

        // ---- gating of pass 1 (:353-384) and, with the flags pass 1 will leave, of pass 2 (:512-542)
        std::vector<Work> w1, w2;
        std::unordered_set<KeyFrame*> will_be_mapped;
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            Pin(kf);
            if (kf->isBad() || kf->semidense_flag_ || !kf->MappingIdDelay()) continue;  // :359
            Work w;
            w.kf = kf;
            if (!ClosestMatches(kf, false, will_be_mapped, w.nbrs)) continue;  // :365-384
            if (!BatchedEdgeDrawing()) DetectEdgeMap(kf);                      // :394-397
            w1.push_back(w);
        }
        if (BatchedEdgeDrawing() && !DetectEdgeMaps(w1)) { UnpinAll(); return; }  // :394 for every keyframe of pass 1 at once
        for (size_t i = 0; i < w1.size(); i++) will_be_mapped.insert(w1[i].kf);
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            const bool mapped = kf->semidense_flag_ || will_be_mapped.count(kf);
            if (kf->isBad() || kf->interKF_depth_flag_ || !kf->MappingIdDelay() || !mapped) continue;  // :518
            Work w;
            w.kf = kf;
            if (!ClosestMatches(kf, true, will_be_mapped, w.nbrs)) continue;  // :523-542
            w2.push_back(w);
        }
        if (!w1.empty() || !w2.empty()) RunLoop(w1, w2);
        UnpinAll();
    }

    // ProbabilityMapping.cc:734-747
    void StereoSearchConstraints(KeyFrame* kf, float* min_depth, float* max_depth)
    {
        std::vector<float> orb_depths = kf->GetAllPointDepths();
        if (orb_depths.empty() ||
            !Check(sdm_stereo_search_constraints(orb_depths.data(), (int)orb_depths.size(), min_depth, max_depth), "stereo")) {
            *min_depth = *max_depth = 0.0f;
        }
    }

    // ProbabilityMapping.cc:749-845.  F12 is accepted for signature compatibility; the library derives
    // it from the two poses with the same arithmetic as ComputeFundamental (:1694-1709).
    void EpipolarSearch(KeyFrame* kf1, KeyFrame* kf2, const int x, const int y, float pixel, float min_depth, float max_depth,
                        depthHo* dh, Mat /*F12*/, float& best_u, float& best_v, float th_pi, float rot)
    {
        int s1, s2;
        if (!SlotsFor(kf1, kf2, s1, s2)) return;
        sdm_hypothesis h;
        if (!Check(sdm_epipolar_search(mCtx, s1, s2, x, y, pixel, min_depth, max_depth, th_pi, rot, &h), "sdm_epipolar_search"))
            return;
        if (h.supported) {  // the reference leaves *dh untouched when no candidate survives
            dh->depth = h.depth;
            dh->sigma = h.sigma;
            dh->supported = true;
            best_u = h.best_u;
            best_v = h.best_v;
        }
    }

    // ProbabilityMapping.cc:1598-1631
    void GetSearchRange(float& umin, float& umax, int px, int py, float mind, float maxd, KeyFrame* kf, KeyFrame* kf2)
    {
        int s1, s2;
        if (!SlotsFor(kf, kf2, s1, s2)) return;
        Check(sdm_search_range(mCtx, s1, s2, px, py, mind, maxd, &umin, &umax), "sdm_search_range");
    }

    // ProbabilityMapping.cc:978-1009
    void InverseDepthHypothesisFusion(const std::vector<depthHo>& h, depthHo& dist)
    {
        dist.depth = 0;
        dist.sigma = 0;
        dist.supported = false;
        const int n = (int)h.size();
        if (n == 0 || n > 32 || !EnsureScratchContext()) return;
        std::vector<float> d(n), s(n);
        for (int i = 0; i < n; i++) { d[i] = h[i].depth; s[i] = h[i].sigma; }
        int32_t cnt = n, ok = 0;
        float od = 0, os = 0;
        if (!Check(sdm_fuse(mCtx, 1, n, d.data(), s.data(), &cnt, &od, &os, &ok), "sdm_fuse")) return;
        if (ok) { dist.depth = od; dist.sigma = os; dist.supported = true; }
    }

    // ProbabilityMapping.cc:866-927 / :929-976 on caller-owned planes (the reference passes kf's planes)
    void IntraKeyFrameDepthChecking(Mat& depth_map, Mat& depth_sigma, const Mat gradimg) { Intra(depth_map, depth_sigma, gradimg, true); }
    void IntraKeyFrameDepthGrowing(Mat& depth_map, Mat& depth_sigma, const Mat gradimg) { Intra(depth_map, depth_sigma, gradimg, false); }

    // ProbabilityMapping.cc:1121-1296
    void InterKeyFrameDepthChecking(KeyFrame* currentKf, std::vector<KeyFrame*> neighbors)
    {
        sdm_item it;
        if (!EnsureUploaded(currentKf) || !BuildItem(currentKf, neighbors, false, it)) return;
        // the neighbours' pass-1 planes live in their cv::Mat members: make the device copies current
        for (size_t j = 0; j < neighbors.size(); j++) PushDepth(neighbors[j]);
        PushDepth(currentKf);
        if (!Check(sdm_inter_check(mCtx, &it), "sdm_inter_check")) return;
        Check(sdm_download(mCtx, mSlot[currentKf], NULL, 0, NULL, 0, currentKf->depth_map_checked_.ptr<float>(0),
                           (size_t)currentKf->depth_map_checked_.step, NULL, 0),
              "sdm_download");
    }

    // ProbabilityMapping.cc:700-731
    void UpdateSemiDensePointSet(KeyFrame* kf)
    {
        std::unique_lock<std::mutex> lock(kf->mMutexSemiDensePoints);
        if (!EnsureResident(kf)) return;
        const int32_t s = mSlot[kf];
        float Tcw[12];
        PoseOf(kf, Tcw);
        sdm_set_pose(mCtx, s, Tcw);  // PoseChanged() refresh (:691-694)
        if (!Check(sdm_update_points(mCtx, 1, &s), "sdm_update_points")) return;
        Check(sdm_download(mCtx, s, NULL, 0, NULL, 0, NULL, 0, kf->SemiDensePointSets_.ptr<float>(0),
                           (size_t)kf->SemiDensePointSets_.step),
              "sdm_download");
    }

    // ProbabilityMapping.cc:678-697
    void UpdateAllSemiDensePointSet()
    {
        std::vector<KeyFrame*> vpKFs = mpMap->GetAllKeyFrames();
        if (vpKFs.size() < 10) return;  // :681
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            kf->SetNotEraseSemiDense();
            if (kf->isBad() || !kf->semidense_flag_ || !kf->interKF_depth_flag_ || !kf->MappingIdDelay()) {  // :686
                kf->SetEraseSemiDense();
                continue;
            }
            if (kf->PoseChanged()) {
                UpdateSemiDensePointSet(kf);
                kf->SetPoseChanged(false);
            }
            kf->SetEraseSemiDense();
        }
    }

    // ProbabilityMapping.cc:657-676 (no caller in the reference; host loop, nothing to accelerate)
    void ApplySigmaThreshold(KeyFrame* kf)
    {
        for (int y = 2; y < kf->im_.rows - 2; y++)
            for (int x = 2; x < kf->im_.cols - 2; x++) {
                if (kf->depth_map_.at<float>(y, x) < 0.000001) continue;
                if (kf->depth_sigma_.at<float>(y, x) > 0.025) kf->depth_map_.at<float>(y, x) = 0.0;
            }
    }

    // ProbabilityMapping.cc:306-345: marks every good keyframe as mapped (the constant-depth test points it builds are
    // discarded in the reference as well)
    void TestSemiDenseViewer()
    {
        std::vector<KeyFrame*> vpKFs = mpMap->GetAllKeyFrames();
        if (vpKFs.size() < 2) return;
        for (size_t i = 0; i < vpKFs.size(); i++)
            if (!vpKFs[i]->isBad() && !vpKFs[i]->semidense_flag_) vpKFs[i]->semidense_flag_ = true;
        std::cout << "semidense_Info:    vpKFs.size()--> " << vpKFs.size() << std::endl;
    }

    // The point filter of SaveSemiDensePoints (:156-186), MapDrawer::DrawSemiDense (MapDrawer.cc:88-117) and the CARV
    // entry: finished keyframes in Map order, pixels in raster order, `sigma > s -> skip; checked > 1e-6 -> emit`,
    // compacted on the device.  pts[k].pixel = (y << 16) | x indexes the keyframe's rgb_ for the colour; counts[i] is
    // the number of points of kfs[i].  Returns the number of points.
    size_t ExportSemiDensePoints(double sigma, std::vector<sdm_point>& pts, std::vector<KeyFrame*>* kfs_out = NULL,
                                 std::vector<uint64_t>* counts_out = NULL)
    {
        pts.clear();
        std::vector<KeyFrame*> vpKFs = mpMap->GetAllKeyFrames(), kfs;
        std::vector<int32_t> slots;
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            if (kf->isBad() || !kf->semidense_flag_ || !kf->interKF_depth_flag_) continue;  // :159
            if (!EnsureResident(kf)) return 0;  // a keyframe finished before the arena was rebuilt: planes go back up
            kfs.push_back(kf);
            slots.push_back(mSlot[kf]);
        }
        if (slots.empty() || !mCtx || !Flush()) return 0;
        std::vector<uint64_t> counts(slots.size());
        uint64_t total = 0;
        if (!Check(sdm_export_points(mCtx, (int)slots.size(), slots.data(), sigma, NULL, 0, counts.data(), &total), "sdm_export_points"))
            return 0;
        pts.resize((size_t)total);
        if (total && !Check(sdm_export_points(mCtx, (int)slots.size(), slots.data(), sigma, pts.data(), pts.size(), counts.data(), &total),
                            "sdm_export_points"))
            pts.clear();
        if (kfs_out) kfs_out->swap(kfs);
        if (counts_out) counts_out->swap(counts);
        return pts.size();
    }

    // ProbabilityMapping.cc:136-192: "v x y z r g b" per surviving point into <results dir>/semi_pointcloud.obj, same
    // order, same filter, same formatting (std::to_string); the points come compacted from the device instead of a
    // scan over every pixel of every keyframe.  A keyframe without rgb_ (stand-alone tests) writes its grey value.
    void SaveSemiDensePoints()
    {
        const std::string dir = ResultsDir();
        if (!MakeDirs(dir)) {
            std::cerr << "Failed to create directory: " << dir << std::endl;
            return;
        }
        const std::string strFileName(dir + "/semi_pointcloud.obj");
        std::ofstream fileOut(strFileName.c_str(), std::ios::out);
        if (!fileOut) {
            std::cerr << "Failed to save semi dense points" << std::endl;
            return;
        }
        std::vector<sdm_point> pts;
        std::vector<KeyFrame*> kfs;
        std::vector<uint64_t> counts;
        ExportSemiDensePoints(0.02, pts, &kfs, &counts);
        size_t k = 0;
        for (size_t i = 0; i < kfs.size(); i++) {
            KeyFrame* kf = kfs[i];
            const bool colour = !kf->rgb_.empty();
            for (uint64_t j = 0; j < counts[i]; j++, k++) {
                const int x = (int)(pts[k].pixel & 0xffffu), y = (int)(pts[k].pixel >> 16);
                float vr, vg, vb;
                if (colour) {
                    float b = kf->rgb_.at<unsigned char>(y, 3 * x) / 255.0;
                    float g = kf->rgb_.at<unsigned char>(y, 3 * x + 1) / 255.0;
                    float r = kf->rgb_.at<unsigned char>(y, 3 * x + 2) / 255.0;
                    vr = r; vg = g; vb = b;
                } else {
                    vr = vg = vb = kf->im_.at<unsigned char>(y, x) / 255.0;
                }
                fileOut << "v " + std::to_string(pts[k].x) + " " + std::to_string(pts[k].y) + " " + std::to_string(pts[k].z) + " "
                               + std::to_string(vr) + " " << std::to_string(vg) + " " + std::to_string(vb) << std::endl;
            }
        }
        fileOut.flush();
        fileOut.close();
        std::cout << "saved semi dense point cloud" << std::endl;
    }

    // the per-keyframe LineDetector::LineFitting of LineFittingOffline (LineDetector.cc:484-523, :884-900) for
    // every finished keyframe of the map in ONE device call, on the planes pass 2 left in the arena.  `chains(kf, offsets,
    // pixels)` supplies the keyframe's edge chains (EdgeMap::segments of kf->mEdgeMap in the reference: offsets.size() =
    // chains + 1 starting at 0, pixels packed (row << 16) | col).  lines[k] are the rows the reference appends to
    // kf->mLinesSeg / kf->mLines3D for kfs[i], counts[i] of them, in the reference's order.  Returns the number of lines.
    typedef std::function<void(KeyFrame*, std::vector<int32_t>&, std::vector<uint32_t>&)> ChainSource;
    size_t FitLines(const ChainSource& chains, std::vector<sdm_line3d>& lines, std::vector<KeyFrame*>* kfs_out = NULL,
                    std::vector<uint64_t>* counts_out = NULL)
    {
        lines.clear();
        std::vector<KeyFrame*> vpKFs = mpMap->GetAllKeyFrames(), kfs;
        std::vector<std::vector<int32_t> > offs;
        std::vector<std::vector<uint32_t> > pix;
        std::vector<sdm_edge_chains> sets;
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            if (kf->isBad() || !kf->semidense_flag_ || !kf->interKF_depth_flag_) continue;  // LineDetector.cc:495
            if (!EnsureResident(kf)) return 0;
            kfs.push_back(kf);
            offs.push_back(std::vector<int32_t>());
            pix.push_back(std::vector<uint32_t>());
            chains(kf, offs.back(), pix.back());
            if (offs.back().empty()) offs.back().push_back(0);
        }
        if (kfs.empty() || !mCtx || !Flush()) return 0;
        for (size_t i = 0; i < kfs.size(); i++) {
            sdm_edge_chains e;
            e.kf = mSlot[kfs[i]];
            e.n_chains = (int32_t)offs[i].size() - 1;
            e.offsets = offs[i].data();
            e.pixels = pix[i].data();
            sets.push_back(e);
        }
        size_t cap = 0;
        for (size_t i = 0; i < offs.size(); i++)
            for (size_t k = 0; k + 1 < offs[i].size(); k++) cap += (size_t)((offs[i][k + 1] - offs[i][k]) / 10);
        lines.resize(cap);
        std::vector<uint64_t> counts(kfs.size());
        uint64_t total = 0;
        if (!Check(sdm_line_fit(mCtx, (int)sets.size(), sets.data(), lines.data(), lines.size(), counts.data(), &total), "sdm_line_fit")) {
            lines.clear();
            return 0;
        }
        lines.resize((size_t)total);
        if (kfs_out) kfs_out->swap(kfs);
        if (counts_out) counts_out->swap(counts);
        return lines.size();
    }

#ifdef SDM_HOST_WITH_ORBSLAM2
    // the device counterpart of `mLineDetector.LineFittingOffline(vpKFs, mpModeler)` (:263): kf->mEdgeMap in, kf->mLinesSeg /
    // kf->mLines3D out (LineDetector.cc:822-823), then the reference's own MergeLines per keyframe (:514)
    void LineFittingOnDevice()
    {
        std::vector<sdm_line3d> lines;
        std::vector<KeyFrame*> kfs;
        std::vector<uint64_t> counts;
        FitLines([](KeyFrame* kf, std::vector<int32_t>& off, std::vector<uint32_t>& pix) {
            off.push_back(0);
            if (kf->mEdgeMap == NULL) { std::cerr << "error: no edge map" << std::endl; return; }  // :889-892
            for (int i = 0; i < kf->mEdgeMap->noSegments; i++) {
                for (int j = 0; j < kf->mEdgeMap->segments[i].noPixels; j++)
                    pix.push_back(((uint32_t)kf->mEdgeMap->segments[i].pixels[j].r << 16) | (uint32_t)kf->mEdgeMap->segments[i].pixels[j].c);
                off.push_back((int32_t)pix.size());
            }
        }, lines, &kfs, &counts);
        size_t k = 0;
        for (size_t i = 0; i < kfs.size(); i++) {
            KeyFrame* kf = kfs[i];
            kf->SetNotEraseSemiDense();
            for (uint64_t j = 0; j < counts[i]; j++, k++) {
                cv::Mat line(1, 6, CV_32F), line2D(1, 4, CV_32F);
                for (int q = 0; q < 6; q++) line.at<float>(0, q) = lines[k].xyz[q];
                for (int q = 0; q < 4; q++) line2D.at<float>(0, q) = lines[k].seg[q];
                kf->mLines3D.push_back(line);
                kf->mLinesSeg.push_back(line2D);
            }
            mLineDetector.MergeLines(kf, mpModeler);
            kf->SetEraseSemiDense();
        }
    }
#endif

#ifdef SDM_HOST_WITH_ORBSLAM2
    Modeler* GetModeler() { return mpModeler; }  // :113-116
    void WriteModel()                            // :119-134
    {
        const std::string dir = ResultsDir();
        if (!MakeDirs(dir)) {
            std::cerr << "Failed to create directory: " << dir << std::endl;
            return;
        }
        mpModeler->WriteModel(dir + "/model.obj");
        std::cout << "saved mesh model" << std::endl;
    }
#endif

    // ProbabilityMapping.cc:847-864 (hash join instead of the O(n1*n2) pointer scan; same multiset)
    std::vector<float> GetRotInPlane(KeyFrame* kf1, KeyFrame* kf2) { return sdm_host::RotInPlane(kf1, kf2); }

    // finish / reset handshakes, ProbabilityMapping.cc:600-654
    void RequestFinish() { std::unique_lock<std::mutex> l(mMutexFinish); mbFinishRequested = true; }
    bool CheckFinish() { std::unique_lock<std::mutex> l(mMutexFinish); return mbFinishRequested; }
    void SetFinish() { std::unique_lock<std::mutex> l(mMutexFinish); mbFinished = true; }
    bool isFinished() { std::unique_lock<std::mutex> l(mMutexFinish); return mbFinished; }
    void RequestReset()
    {
        { std::unique_lock<std::mutex> l(mMutexReset); mbResetRequested = true; }
        while (true) {
            { std::unique_lock<std::mutex> l(mMutexReset); if (!mbResetRequested) break; }
            std::this_thread::sleep_for(std::chrono::milliseconds(3));
        }
    }
    void ResetIfRequested()
    {
        std::unique_lock<std::mutex> l(mMutexReset);
        if (mbResetRequested) {
            DropContext();
#ifdef SDM_HOST_WITH_ORBSLAM2
            mLineDetector.Reset();  // :618-626
            delete mpModeler;
            mpModeler = new Modeler(mpMap);
            mpMap->SetModeler(mpModeler);
#endif
            mbResetRequested = false;
        }
    }

    // per-stage device times of the last loop (the "... took" prints of :389-443, :505-508, :545-565)
    sdm_timing LastTiming() { sdm_timing t = sdm_timing(); if (mCtx) sdm_last_timing(mCtx, &t); return t; }
    // forget the device copies of all keyframes (the arena stays allocated): the next call uploads what it needs again.
    // For hosts that rewrite keyframe planes in place, and for timing a cold loop.
    void ForgetResidentKeyFrames()
    {
        if (!mCtx) return;
        sdm_synchronize(mCtx);
        mSlot.clear();
        mPending.clear();
        mPendingKFs.clear();
        mFree.clear();
        for (int s = mCapacity - 2; s >= 0; s--) mFree.push_back(s);
    }
    // device slots currently reserved / in use (diagnostics; tests)
    int ArenaCapacity() const { return mCapacity; }
    int ResidentKeyFrames() const { return (int)mSlot.size(); }

    std::mutex mMutexSemiDense;  // ProbabilityMapping.h:117

private:
    struct Work {
        KeyFrame* kf;
        std::vector<KeyFrame*> nbrs;
    };

    bool Check(int rc, const char* what)
    {
        if (rc == SDM_OK) return true;
        std::cerr << "ProbabilityMapping(sdm_b200): " << what << " failed: " << sdm_last_error() << std::endl;  // cf. :124-127
        return false;
    }

    static void PoseOf(KeyFrame* kf, float Tcw[12])
    {
        Mat T = kf->GetPose();  // 4x4 CV_32F
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) Tcw[r * 4 + c] = T.at<float>(r, c);
    }

    // erase pins of the reference (SetNotEraseSemiDense / SetEraseSemiDense around every use, :357-594): held for the
    // whole pipelined loop, because the device reads a keyframe's host planes until the loop's synchronise
    void Pin(KeyFrame* kf) { kf->SetNotEraseSemiDense(); mPinned.push_back(kf); }
    void UnpinAll()
    {
        for (size_t i = 0; i < mPinned.size(); i++) mPinned[i]->SetEraseSemiDense();
        mPinned.clear();
    }

    bool BatchedEdgeDrawing() const { return mEdgeDrawing && !mEdgeHook; }
    bool DetectEdgeMaps(const std::vector<Work>& w1)
    {
        if (w1.empty()) return true;
        std::vector<sdm_ed_image> ims(w1.size());
        for (size_t i = 0; i < w1.size(); i++) {
            KeyFrame* kf = w1[i].kf;
            if (kf->mEdgeIndex.empty()) kf->mEdgeIndex = sdm_host::plane32s(kf->im_.rows, kf->im_.cols);  // KeyFrame.cc:87
            ims[i].im = kf->im_.ptr<uint8_t>(0);
            ims[i].im_step = (size_t)kf->im_.step;
            ims[i].edge_index = kf->mEdgeIndex.ptr<int32_t>(0);
            ims[i].edge_step = (size_t)kf->mEdgeIndex.step;
        }
        sdm_ed_result* res = NULL;
        if (!Check(sdm_set_edge_drawing_route(mCtx, mEdRouteDevice ? SDM_ED_ROUTE_DEVICE : SDM_ED_ROUTE_HOST), "sdm_set_edge_drawing_route"))
            return false;
        if (!Check(sdm_edge_drawing(mCtx, (int)ims.size(), ims.data(), mEdGrad, mEdAnchor, mEdThreads, &res), "sdm_edge_drawing"))
            return false;
        for (size_t i = 0; i < w1.size(); i++) {
            int32_t n = 0;
            const int32_t* off = NULL;
            const uint32_t* pix = NULL;
            sdm_ed_chains(res, (int)i, &n, &off, &pix);
            sdm_host::EdgeChains& e = mEdgeChains[w1[i].kf];
            e.offsets.assign(off, off + n + 1);
            e.pixels.assign(pix, pix + off[n]);
#ifdef SDM_HOST_WITH_ORBSLAM2
            {   // kf->mEdgeMap = map (:869), in the library's result type
                KeyFrame* kf = w1[i].kf;
                EdgeMap* map = new EdgeMap(kf->im_.cols, kf->im_.rows);
                for (int k = 0; k < n; k++) {
                    map->segments[k].pixels = map->pixels + off[k];
                    map->segments[k].noPixels = off[k + 1] - off[k];
                }
                for (int q = 0; q < off[n]; q++) { map->pixels[q].r = (int)(pix[q] >> 16); map->pixels[q].c = (int)(pix[q] & 0xffffu); }
                map->noSegments = n;
                kf->mEdgeMap = map;
                mLineDetector.DetectLineSegments(kf);  // :397
            }
#endif
        }
        sdm_ed_free(res);
        return true;
    }

    void DetectEdgeMap(KeyFrame* kf)
    {
        if (mEdgeHook) { mEdgeHook(kf); return; }
#ifdef SDM_HOST_WITH_ORBSLAM2
        mLineDetector.DetectEdgeMap(kf);       // :394
        mLineDetector.DetectLineSegments(kf);  // :397
#else
        (void)kf;
#endif
    }

    std::string ResultsDir()
    {
        if (!mResultsDir.empty()) return mResultsDir;
#ifdef SDM_HOST_WITH_ORBSLAM2
        return "results_line_segments/" + mLineDetector.GetStringDateTime();
#else
        char buf[64];
        const std::time_t t = std::time(NULL);
        std::strftime(buf, sizeof(buf), "%Y-%m-%d_%H-%M-%S", std::localtime(&t));
        return std::string("results_line_segments/") + buf;
#endif
    }
    static bool MakeDirs(const std::string& path)  // boost::filesystem::create_directories of :121-127 / :140-146
    {
        for (size_t i = 1; i <= path.size(); i++)
            if (i == path.size() || path[i] == '/') {
                const std::string p = path.substr(0, i);
                struct stat st;
                if (stat(p.c_str(), &st) == 0) continue;
                if (mkdir(p.c_str(), 0777) != 0 && stat(p.c_str(), &st) != 0) return false;
            }
        return true;
    }

    void DropContext()
    {
        sdm_destroy(mCtx);
        mCtx = NULL;
        mSlot.clear();
        mFree.clear();
        mPending.clear();
        mPendingKFs.clear();
        mCapacity = 0;
    }

    bool CreateContext(int W, int H, int capacity)
    {
        DropContext();
        mCfg.width = W;
        mCfg.height = H;
        mCfg.max_keyframes = capacity;
        if (!Check(sdm_create(&mCfg, &mCtx), "sdm_create")) return false;  // no CPU fallback: the loop is skipped
        mW = W; mH = H; mCapacity = capacity;
        for (int s = capacity - 2; s >= 0; s--) mFree.push_back(s);  // the last slot is the scratch slot of Intra()
        return true;
    }

    // One slot per good keyframe of the map + the scratch slot.  Slots of keyframes that went bad are recycled; when the
    // map has outgrown the arena a larger one is created - the device state of finished keyframes is not lost with it:
    // EnsureResident() re-seeds their planes from the cv::Mats the first time they are needed again.
    bool EnsureContext(const std::vector<KeyFrame*>& kfs)
    {
        int W = 0, H = 0;
        for (size_t i = 0; i < kfs.size(); i++)
            if (!kfs[i]->im_.empty()) { W = kfs[i]->im_.cols; H = kfs[i]->im_.rows; break; }
        if (W == 0) return false;
        if (mCtx && W == mW && H == mH) {
            for (std::unordered_map<KeyFrame*, int>::iterator it = mSlot.begin(); it != mSlot.end();)
                if (it->first->isBad()) { mFree.push_back(it->second); it = mSlot.erase(it); } else ++it;
            size_t need = 0;
            for (size_t i = 0; i < kfs.size(); i++)
                if (!kfs[i]->isBad() && !mSlot.count(kfs[i])) need++;
            if (need <= mFree.size()) return true;
        }
        return CreateContext(W, H, (int)kfs.size() + 1 + (mHeadroom >= 0 ? mHeadroom : (int)kfs.size() / 2 + 16));
    }

    bool EnsureScratchContext() { return mCtx || CreateContext(64, 64, 2); }

    // Reserve a device slot for kf and queue its planes for upload; the queue goes out as ONE
    // sdm_upload_keyframes call (Flush, or the upload list of sdm_run_loop) before the planes are needed.
    bool EnsureUploaded(KeyFrame* kf)
    {
        if (!mCtx && !CreateContext(kf->im_.cols, kf->im_.rows, 64)) return false;
        if (mSlot.count(kf)) return true;
        if (mFree.empty()) {
            std::cerr << "ProbabilityMapping(sdm_b200): device arena full (" << mCapacity << " keyframes)" << std::endl;
            return false;
        }
        sdm_upload_desc u;
        std::memset(&u, 0, sizeof(u));
        u.kf = (int32_t)mFree.back();
        mFree.pop_back();
        u.im = kf->im_.ptr<uint8_t>(0);        u.im_step = (size_t)kf->im_.step;
        if (!mDevicePlanes) {
            u.grad = kf->GradImg.ptr<float>(0);    u.grad_step = (size_t)kf->GradImg.step;
            u.theta = kf->GradTheta.ptr<float>(0); u.theta_step = (size_t)kf->GradTheta.step;
        }
        if (!kf->mEdgeIndex.empty()) { u.edge = kf->mEdgeIndex.ptr<int32_t>(0); u.edge_step = (size_t)kf->mEdgeIndex.step; }
        u.K[0] = kf->fx; u.K[1] = kf->fy; u.K[2] = kf->cx; u.K[3] = kf->cy;
        PoseOf(kf, u.Tcw);
        mSlot[kf] = u.kf;
        mPending.push_back(u);
        mPendingKFs.push_back(kf);
        return true;
    }

    bool Flush()
    {
        if (mPending.empty()) return true;
        const bool ok = Check(sdm_upload_keyframes(mCtx, (int)mPending.size(), mPending.data()), "sdm_upload_keyframes");
        std::vector<KeyFrame*> kfs;
        kfs.swap(mPendingKFs);
        mPending.clear();
        if (!ok) return false;
        for (size_t i = 0; i < kfs.size(); i++)
            if (!Reseed(kfs[i])) return false;
        return true;
    }

    // a keyframe processed in an earlier loop whose slot is new (first use by a single-method call, arena rebuilt,
    // reset): its results live in the cv::Mats - pass-1 planes for neighbours' pass 2, depth_map_checked_ for
    // UpdateSemiDensePointSet / the exporters (the point set is recomputed from it, :700-731)
    bool Reseed(KeyFrame* kf)
    {
        const int32_t s = mSlot[kf];
        if (kf->semidense_flag_ &&
            !Check(sdm_upload_depth(mCtx, s, kf->depth_map_.ptr<float>(0), (size_t)kf->depth_map_.step,
                                    kf->depth_sigma_.ptr<float>(0), (size_t)kf->depth_sigma_.step), "sdm_upload_depth"))
            return false;
        if (kf->interKF_depth_flag_) {
            if (!Check(sdm_upload_checked(mCtx, s, kf->depth_map_checked_.ptr<float>(0), (size_t)kf->depth_map_checked_.step),
                       "sdm_upload_checked") ||
                !Check(sdm_update_points(mCtx, 1, &s), "sdm_update_points"))
                return false;
        }
        return true;
    }

    bool EnsureResident(KeyFrame* kf) { return EnsureUploaded(kf) && Flush(); }

    void PushDepth(KeyFrame* kf)
    {
        if (!EnsureResident(kf)) return;
        Check(sdm_upload_depth(mCtx, mSlot[kf], kf->depth_map_.ptr<float>(0), (size_t)kf->depth_map_.step,
                               kf->depth_sigma_.ptr<float>(0), (size_t)kf->depth_sigma_.step),
              "sdm_upload_depth");
    }

    bool SlotsFor(KeyFrame* a, KeyFrame* b, int& sa, int& sb)
    {
        if (!EnsureUploaded(a) || !EnsureUploaded(b) || !Flush()) return false;
        sa = mSlot[a];
        sb = mSlot[b];
        return true;
    }

    // :365-384 (pass 1) and :523-542 (pass 2, additionally requires the neighbour's semidense_flag_, which pass 1 of
    // this same loop sets for the keyframes in `mapped_now`, :497)
    bool ClosestMatches(KeyFrame* kf, bool pass2, const std::unordered_set<KeyFrame*>& mapped_now, std::vector<KeyFrame*>& closestMatches)
    {
        std::vector<KeyFrame*> all = kf->GetVectorCovisibleKeyFrames();
        for (size_t i = 0; i < all.size(); i++) {
            if ((int)closestMatches.size() >= mN) break;
            KeyFrame* k = all[i];
            Pin(k);
            if (k->isBad() || !k->Mapped() || (pass2 && !(k->semidense_flag_ || mapped_now.count(k)))) continue;
            closestMatches.push_back(k);
        }
        return (int)closestMatches.size() >= mN;
    }

    // the per-keyframe set-up of :406-438: rotIs medians, StereoSearchConstraints (F12 is computed in the library from
    // the poses the slots hold, so resident slots get the keyframes' CURRENT poses first: the reference reads
    // GetRotation() / GetTranslation() afresh in every loop, and local BA / loop closing move keyframes in between)
    bool BuildItem(KeyFrame* kf, const std::vector<KeyFrame*>& nbrs, bool pass1, sdm_item& it)
    {
        if ((int)nbrs.size() > SDM_MAX_NBR) {
            std::cerr << "ProbabilityMapping(sdm_b200): " << nbrs.size() << " neighbours > SDM_MAX_NBR" << std::endl;
            return false;
        }
        if (!EnsureUploaded(kf)) return false;
        std::memset(&it, 0, sizeof(it));
        it.kf = mSlot[kf];
        it.n_nbr = (int32_t)nbrs.size();
        RefreshPose(kf);
        for (size_t j = 0; j < nbrs.size(); j++) {
            if (!EnsureUploaded(nbrs[j])) return false;
            it.nbr[j] = mSlot[nbrs[j]];
            RefreshPose(nbrs[j]);
            if (pass1) {
                std::vector<float> rot = GetRotInPlane(kf, nbrs[j]);
                std::sort(rot.begin(), rot.end());
                it.rot_deg[j] = rot.empty() ? 0.0f : rot[(rot.size() - 1) / 2];  // :408-413
            }
        }
        if (pass1) StereoSearchConstraints(kf, &it.min_depth, &it.max_depth);  // :427
        return true;
    }

    void RefreshPose(KeyFrame* kf)
    {
        float Tcw[12];
        PoseOf(kf, Tcw);
        sdm_set_pose(mCtx, mSlot[kf], Tcw);  // host-side bookkeeping of the slot: a 48-byte copy
        for (size_t i = 0; i < mPending.size(); i++)  // not uploaded yet: the queued descriptor carries the pose
            if (mPendingKFs[i] == kf) std::memcpy(mPending[i].Tcw, Tcw, sizeof(Tcw));
    }

    // the device work of one SemiDenseLoop as ONE pipelined library call
    void RunLoop(const std::vector<Work>& w1, const std::vector<Work>& w2)
    {
        // keyframes finished in an earlier loop whose slot is new go up first, with their planes (rare: arena rebuilt)
        for (size_t pass = 0; pass < 2; pass++) {
            const std::vector<Work>& w = pass ? w2 : w1;
            for (size_t i = 0; i < w.size(); i++)
                for (size_t j = 0; j < w[i].nbrs.size(); j++)
                    if (w[i].nbrs[j]->semidense_flag_ && !mSlot.count(w[i].nbrs[j]) && !EnsureResident(w[i].nbrs[j])) return;
        }
        for (size_t i = 0; i < w2.size(); i++)
            if (w2[i].kf->semidense_flag_ && !mSlot.count(w2[i].kf) && !EnsureResident(w2[i].kf)) return;
        if (!Flush()) return;
        std::vector<sdm_item> it1(w1.size()), it2(w2.size());
        std::vector<sdm_download_desc> d1(w1.size()), d2(w2.size());
        for (size_t i = 0; i < w1.size(); i++) {
            KeyFrame* kf = w1[i].kf;
            if (!BuildItem(kf, w1[i].nbrs, true, it1[i])) return;
            std::memset(&d1[i], 0, sizeof(d1[i]));
            d1[i].kf = it1[i].kf;
            d1[i].depth = kf->depth_map_.ptr<float>(0);   d1[i].depth_step = (size_t)kf->depth_map_.step;
            d1[i].sigma = kf->depth_sigma_.ptr<float>(0); d1[i].sigma_step = (size_t)kf->depth_sigma_.step;
        }
        for (size_t i = 0; i < w2.size(); i++) {
            KeyFrame* kf = w2[i].kf;
            if (!BuildItem(kf, w2[i].nbrs, false, it2[i])) return;
            std::memset(&d2[i], 0, sizeof(d2[i]));
            d2[i].kf = it2[i].kf;
            d2[i].checked = kf->depth_map_checked_.ptr<float>(0); d2[i].checked_step = (size_t)kf->depth_map_checked_.step;
            d2[i].points = kf->SemiDensePointSets_.ptr<float>(0); d2[i].points_step = (size_t)kf->SemiDensePointSets_.step;
        }
        // SemiDensePointSets_ is written by DMA while the loop runs; readers (MapDrawer.cc:92-97) look at a keyframe only
        // after its interKF_depth_flag_ is set, which happens below, after the synchronise - no keyframe mutex is needed
        sdm_loop L;
        std::memset(&L, 0, sizeof(L));
        std::vector<sdm_upload_desc> up;
        up.swap(mPending);
        mPendingKFs.clear();
        L.n_upload = (int32_t)up.size();  L.upload = up.empty() ? NULL : up.data();
        L.n_pass1 = (int32_t)it1.size();  L.pass1 = it1.empty() ? NULL : it1.data();  L.down1 = d1.empty() ? NULL : d1.data();
        L.n_pass2 = (int32_t)it2.size();  L.pass2 = it2.empty() ? NULL : it2.data();  L.down2 = d2.empty() ? NULL : d2.data();
        L.chunk = mChunk;
        L.sparse_download = mSparse ? 1 : 0;
        if (!Check(sdm_run_loop(mCtx, &L), "sdm_run_loop") || !Check(sdm_synchronize(mCtx), "sdm_synchronize")) return;
        for (size_t i = 0; i < w1.size(); i++) w1[i].kf->semidense_flag_ = true;       // :497
        for (size_t i = 0; i < w2.size(); i++) w2[i].kf->interKF_depth_flag_ = true;  // :554
    }

    void Intra(Mat& depth_map, Mat& depth_sigma, const Mat& gradimg, bool check)
    {
        const int W = depth_map.cols, H = depth_map.rows;
        if (!mCtx && !CreateContext(W, H, 64)) return;
        if (W != mW || H != mH) { std::cerr << "ProbabilityMapping(sdm_b200): plane size differs from the context" << std::endl; return; }
        // a scratch slot: the last one of the arena is reserved for caller-owned planes
        const int s = mCapacity - 1;
        if (!check) {  // growing reads GradImg (:942): the scratch slot needs the gradient plane
            std::vector<uint8_t> im((size_t)W * H, 0);
            const float K[4] = {1, 1, 0, 0};
            const float T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
            if (!Check(sdm_upload_keyframe(mCtx, s, im.data(), (size_t)W, gradimg.ptr<float>(0), (size_t)gradimg.step,
                                           gradimg.ptr<float>(0), (size_t)gradimg.step, NULL, 0, K, T), "sdm_upload_keyframe"))
                return;
            sdm_synchronize(mCtx);
        }
        if (!Check(sdm_upload_depth(mCtx, s, depth_map.ptr<float>(0), (size_t)depth_map.step,
                                    depth_sigma.ptr<float>(0), (size_t)depth_sigma.step), "sdm_upload_depth"))
            return;
        if (!Check(check ? sdm_intra_check(mCtx, s) : sdm_intra_grow(mCtx, s), "sdm_intra")) return;
        Check(sdm_download(mCtx, s, depth_map.ptr<float>(0), (size_t)depth_map.step, depth_sigma.ptr<float>(0),
                           (size_t)depth_sigma.step, NULL, 0, NULL, 0), "sdm_download");
    }

    Map* mpMap;
    sdm_ctx* mCtx;
    sdm_config mCfg;
    int mN, mW, mH, mCapacity, mChunk, mHeadroom;
    bool mDevicePlanes, mSparse, mOnline;
    std::unordered_map<KeyFrame*, int> mSlot;
    std::vector<int> mFree;                  // unused device slots
    std::vector<sdm_upload_desc> mPending;   // queued uploads (EnsureUploaded / Flush)
    std::vector<KeyFrame*> mPendingKFs;
    std::vector<KeyFrame*> mPinned;
    std::function<void(KeyFrame*)> mEdgeHook;
    bool mEdgeDrawing, mEdRouteDevice;
    int mEdThreads, mEdGrad, mEdAnchor;
    std::unordered_map<KeyFrame*, sdm_host::EdgeChains> mEdgeChains;
    std::string mResultsDir;
#ifdef SDM_HOST_WITH_ORBSLAM2
    Modeler* mpModeler;
    LineDetector mLineDetector;
#endif
    bool mbFinishRequested, mbFinished, mbResetRequested;
    std::mutex mMutexFinish, mMutexReset;
};
