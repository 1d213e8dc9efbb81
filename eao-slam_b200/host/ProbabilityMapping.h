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
//     the reference's own types (INTEGRATION.md shows the three-line change to System.cc);
//   * stand-alone (this repository, no OpenCV available): host_types.h supplies minimal stand-ins
//     with the same member names, which is what tests/cpp/test_shim.cpp compiles against.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/sdm_b200.h"

#ifdef SDM_HOST_WITH_ORBSLAM2
#include <opencv2/core/core.hpp>
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
namespace sdm_host {
typedef cv::Mat Mat;
typedef ORB_SLAM2::KeyFrame KeyFrame;
typedef ORB_SLAM2::Map Map;
inline Mat zeros32f(int rows, int cols) { return cv::Mat::zeros(rows, cols, CV_32F); }
}  // namespace sdm_host
#else
#include "host_types.h"
#endif

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
        : mMutexSemiDense(), mpMap(pMap), mCtx(NULL), mN(covisN), mW(0), mH(0), mCapacity(0), mDevicePlanes(false), mOnline(false), mbFinishRequested(false),
          mbFinished(false), mbResetRequested(false)
    {
        sdm_default_config(&mCfg);
    }
    ~ProbabilityMapping() { sdm_destroy(mCtx); }

    // run-time versions of the #defines of ProbabilityMapping.h:45-56; call before the first loop
    sdm_config& Config() { return mCfg; }
    void SetCovisN(int n) { mN = n; }
    // upload im_ only and let the device produce GradImg / GradTheta (KeyFrame.cc:69-74: scalar-form magnitude / phase,
    // <= 1 ulp / 3e-5 deg from OpenCV's SIMD kernels) instead of uploading the keyframe's own planes; 1 B/px instead of 9
    void SetProducePlanesOnDevice(bool on) { mDevicePlanes = on; }

    // ProbabilityMapping.cc:204-300.  Offline mode (the reference's build, `#define OnlineLoop` commented out at
    // ProbabilityMapping.h:42): idle until finish is requested, then one loop.  Online mode (:223-234): every 5 ms the
    // keyframes that became eligible are processed (the gating of SemiDenseLoop skips finished ones, their planes stay
    // resident on the device) and the point sets of keyframes whose pose changed are refreshed.
    void SetOnline(bool on) { mOnline = on; }
    void Run()
    {
        while (true) {
            if (CheckFinish()) break;
            if (mOnline) {
                SemiDenseLoop();
                UpdateAllSemiDensePointSet();  // make point position dependent to kf position (:226)
            }
            ResetIfRequested();
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
        SemiDenseLoop();
        SetFinish();
    }

    // ProbabilityMapping.cc:348-597.  Same gating, same two passes; the per-keyframe bodies of each pass
    // are collected and launched as ONE batch (pass 1 of a keyframe reads only immutable inputs of its
    // neighbours, pass 2 only pass-1 planes, so the batching does not change any result).
    void SemiDenseLoop()
    {
        std::unique_lock<std::mutex> lock(mMutexSemiDense);
        std::vector<KeyFrame*> vpKFs = mpMap->GetAllKeyFrames();
        if (vpKFs.size() < 10) return;  // :351
        if (!EnsureContext(vpKFs)) return;

        // ---- pass 1 (:353-510)
        std::vector<sdm_item> items;
        std::vector<KeyFrame*> owners;
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            kf->SetNotEraseSemiDense();
            if (kf->isBad() || kf->semidense_flag_ || !kf->MappingIdDelay()) {  // :359
                kf->SetEraseSemiDense();
                continue;
            }
            std::vector<KeyFrame*> closestMatches;
            if (!ClosestMatches(kf, false, closestMatches)) {  // :365-384
                kf->SetEraseSemiDense();
                continue;
            }
            sdm_item it;
            if (BuildItem(kf, closestMatches, true, it)) {
                items.push_back(it);
                owners.push_back(kf);
            }
            for (size_t j = 0; j < closestMatches.size(); j++) closestMatches[j]->SetEraseSemiDense();
            kf->SetEraseSemiDense();
        }
        if (!items.empty()) {
            if (!Flush() || !Check(sdm_pass1(mCtx, (int)items.size(), items.data()), "sdm_pass1")) return;
            // depth_map_ / depth_sigma_ of the whole batch come back with one call (DMA straight from dense planes)
            std::vector<sdm_download_desc> dl(owners.size());
            for (size_t i = 0; i < owners.size(); i++) {
                KeyFrame* kf = owners[i];
                std::memset(&dl[i], 0, sizeof(dl[i]));
                dl[i].kf = mSlot[kf];
                dl[i].depth = kf->depth_map_.ptr<float>(0);   dl[i].depth_step = (size_t)kf->depth_map_.step;
                dl[i].sigma = kf->depth_sigma_.ptr<float>(0); dl[i].sigma_step = (size_t)kf->depth_sigma_.step;
            }
            if (!Check(sdm_download_keyframes(mCtx, (int)dl.size(), dl.data()), "sdm_download_keyframes") ||
                !Check(sdm_synchronize(mCtx), "sdm_synchronize"))
                return;
            for (size_t i = 0; i < owners.size(); i++) owners[i]->semidense_flag_ = true;  // :497
        }

        // ---- pass 2 (:512-596)
        items.clear();
        owners.clear();
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            kf->SetNotEraseSemiDense();
            if (kf->isBad() || kf->interKF_depth_flag_ || !kf->MappingIdDelay() || !kf->semidense_flag_) {  // :518
                kf->SetEraseSemiDense();
                continue;
            }
            std::vector<KeyFrame*> closestMatches;
            if (!ClosestMatches(kf, true, closestMatches)) {  // :523-542
                kf->SetEraseSemiDense();
                continue;
            }
            sdm_item it;
            if (BuildItem(kf, closestMatches, false, it)) {
                items.push_back(it);
                owners.push_back(kf);
            }
            for (size_t j = 0; j < closestMatches.size(); j++) closestMatches[j]->SetEraseSemiDense();
            kf->SetEraseSemiDense();
        }
        if (!items.empty()) {
            if (!Flush() || !Check(sdm_pass2(mCtx, (int)items.size(), items.data()), "sdm_pass2")) return;
            // depth_map_checked_ / SemiDensePointSets_ in groups of 16 keyframes, each group under the keyframes'
            // mMutexSemiDensePoints like UpdateSemiDensePointSet (:701) so the viewer never sees a half-written set
            for (size_t i0 = 0; i0 < owners.size(); i0 += 16) {
                const size_t m = std::min<size_t>(16, owners.size() - i0);
                std::vector<std::unique_lock<std::mutex> > locks;
                std::vector<sdm_download_desc> dl(m);
                for (size_t i = 0; i < m; i++) {
                    KeyFrame* kf = owners[i0 + i];
                    locks.emplace_back(kf->mMutexSemiDensePoints);
                    std::memset(&dl[i], 0, sizeof(dl[i]));
                    dl[i].kf = mSlot[kf];
                    dl[i].checked = kf->depth_map_checked_.ptr<float>(0); dl[i].checked_step = (size_t)kf->depth_map_checked_.step;
                    dl[i].points = kf->SemiDensePointSets_.ptr<float>(0); dl[i].points_step = (size_t)kf->SemiDensePointSets_.step;
                }
                if (!Check(sdm_download_keyframes(mCtx, (int)m, dl.data()), "sdm_download_keyframes") ||
                    !Check(sdm_synchronize(mCtx), "sdm_synchronize"))
                    return;
                for (size_t i = 0; i < m; i++) owners[i0 + i]->interKF_depth_flag_ = true;  // :554
            }
        }
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
        if (!EnsureUploaded(kf) || !Flush()) return;
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
        for (size_t i = 0; i < vpKFs.size(); i++) {
            KeyFrame* kf = vpKFs[i];
            if (kf->isBad() || !kf->interKF_depth_flag_) continue;
            if (kf->PoseChanged()) {
                UpdateSemiDensePointSet(kf);
                kf->SetPoseChanged(false);
            }
        }
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
            if (kf->isBad() || !kf->semidense_flag_ || !kf->interKF_depth_flag_ || !mSlot.count(kf)) continue;  // :159
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
            sdm_destroy(mCtx);
            mCtx = NULL;
            mSlot.clear();
            mPending.clear();
            mPendingKFs.clear();
            mCapacity = 0;
            mbResetRequested = false;
        }
    }

    // per-stage device times of the last loop (the "... took" prints of :389-443, :505-508, :545-565)
    sdm_timing LastTiming() { sdm_timing t = sdm_timing(); if (mCtx) sdm_last_timing(mCtx, &t); return t; }

    std::mutex mMutexSemiDense;  // ProbabilityMapping.h:117

private:
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

    bool CreateContext(int W, int H, int capacity)
    {
        sdm_destroy(mCtx);
        mCtx = NULL;
        mSlot.clear();
        mPending.clear();
        mPendingKFs.clear();
        mCfg.width = W;
        mCfg.height = H;
        mCfg.max_keyframes = capacity;
        if (!Check(sdm_create(&mCfg, &mCtx), "sdm_create")) return false;  // no CPU fallback: the loop is skipped
        mW = W; mH = H; mCapacity = capacity;
        return true;
    }

    bool EnsureContext(const std::vector<KeyFrame*>& kfs)
    {
        int W = 0, H = 0;
        for (size_t i = 0; i < kfs.size(); i++)
            if (!kfs[i]->im_.empty()) { W = kfs[i]->im_.cols; H = kfs[i]->im_.rows; break; }
        if (W == 0) return false;
        if (mCtx && W == mW && H == mH && (int)kfs.size() <= mCapacity) return true;
        // (re)create with head-room; slots are re-filled on demand (the flags live in the keyframes)
        return CreateContext(W, H, (int)kfs.size() + (int)kfs.size() / 2 + 16);
    }

    bool EnsureScratchContext() { return mCtx || CreateContext(64, 64, 1); }

    // Reserve a device slot for kf and queue its planes for upload; the queue goes out as ONE
    // sdm_upload_keyframes call (Flush) before the next library call that needs the planes.
    bool EnsureUploaded(KeyFrame* kf)
    {
        if (!mCtx && !CreateContext(kf->im_.cols, kf->im_.rows, 64)) return false;
        if (mSlot.count(kf)) return true;
        if ((int)mSlot.size() >= mCapacity - 1) {  // the last slot is the scratch slot of Intra()
            std::cerr << "ProbabilityMapping(sdm_b200): device arena full (" << mCapacity << " keyframes)" << std::endl;
            return false;
        }
        sdm_upload_desc u;
        std::memset(&u, 0, sizeof(u));
        u.kf = (int32_t)mSlot.size();
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
        // keyframes mapped in an earlier loop: their pass-1 planes live in the cv::Mats
        for (size_t i = 0; i < kfs.size(); i++)
            if (kfs[i]->semidense_flag_) PushDepth(kfs[i]);
        return true;
    }

    void PushDepth(KeyFrame* kf)
    {
        if (!EnsureUploaded(kf) || !Flush()) return;
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

    // :365-384 (pass 1) and :523-542 (pass 2, additionally requires the neighbour's semidense_flag_)
    bool ClosestMatches(KeyFrame* kf, bool pass2, std::vector<KeyFrame*>& closestMatches)
    {
        std::vector<KeyFrame*> all = kf->GetVectorCovisibleKeyFrames();
        for (size_t i = 0; i < all.size(); i++) {
            if ((int)closestMatches.size() >= mN) break;
            KeyFrame* k = all[i];
            k->SetNotEraseSemiDense();
            if (k->isBad() || !k->Mapped() || (pass2 && !k->semidense_flag_)) {
                k->SetEraseSemiDense();
                continue;
            }
            closestMatches.push_back(k);
        }
        if ((int)closestMatches.size() < mN) {
            for (size_t i = 0; i < closestMatches.size(); i++) closestMatches[i]->SetEraseSemiDense();
            return false;
        }
        return true;
    }

    // the per-keyframe set-up of :406-438: rotIs medians, StereoSearchConstraints (F12 is computed in the library)
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
        for (size_t j = 0; j < nbrs.size(); j++) {
            if (!EnsureUploaded(nbrs[j])) return false;
            it.nbr[j] = mSlot[nbrs[j]];
            if (pass1) {
                std::vector<float> rot = GetRotInPlane(kf, nbrs[j]);
                std::sort(rot.begin(), rot.end());
                it.rot_deg[j] = rot.empty() ? 0.0f : rot[(rot.size() - 1) / 2];  // :408-413
            }
        }
        if (pass1) StereoSearchConstraints(kf, &it.min_depth, &it.max_depth);  // :427
        return true;
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
    int mN, mW, mH, mCapacity;
    bool mDevicePlanes, mOnline;
    std::unordered_map<KeyFrame*, int> mSlot;
    std::vector<sdm_upload_desc> mPending;   // queued uploads (EnsureUploaded / Flush)
    std::vector<KeyFrame*> mPendingKFs;
    bool mbFinishRequested, mbFinished, mbResetRequested;
    std::mutex mMutexFinish, mMutexReset;
};
