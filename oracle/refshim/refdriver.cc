// refdriver.cc — C entry point that runs the REFERENCE'S OWN ProbabilityMapping::SemiDenseLoop()
// (/root/reference/src/ProbabilityMapping.cc, compiled where it lies against the stand-in headers of this directory)
// on caller-supplied keyframes.  TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libref_pm.so by oracle/Makefile,
// used by tests/test_ref_vs_oracle.py to validate the C restatement (oracle/sdm_oracle.c) against the reference text.
#include <cstdint>
#include <iostream>
#include <memory>
#include <sstream>
#include <vector>

#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "ProbabilityMapping.h"  // the reference's header (include/ProbabilityMapping.h)

long unsigned int ORB_SLAM2::KeyFrame::nNextMappingId = 1;  // (0 in KeyFrame.cc:30: the first keyframe is then never "Mapped")

static cv::Mat wrap_copy(const void* src, int rows, int cols, int type)
{
    cv::Mat m(rows, cols, type);
    memcpy(m.data, src, (size_t)rows * cols * m.elemSize());
    return m;
}

extern "C" {

int ref_covisN(void) { return covisN; }

// Runs SemiDenseLoop() over n keyframes (dense row-major planes).  nbr_idx: [n][covisN] covisibility lists in
// order.  The in-plane rotation of a pair comes out of the reference's own GetRotInPlane + median (:406-415): the
// driver gives keyframe k ORB keypoints of angle kp_angle[k] on one map point per unordered covisible pair, so
// rotIs[kf2] = kp_angle[kf2] - kp_angle[kf] in float.  inv_depths: [n][n_inv] sorted inverse depths per keyframe
// (GetAllPointDepths).  Outputs: the four planes of every keyframe and the flags.
// _ex: n_cov = length of every covisibility list (>= covisN; the reference takes the first covisN GOOD ones, :365-384),
// first_id = value of KeyFrame::nNextMappingId before the first keyframe is mapped (0 in the reference: that keyframe
// is then never "Mapped", KeyFrame.cc:796-806), extra_ids = keyframes mapped after ours (MappingIdDelay needs > 10
// newer ones, :789-794), bad[n] = isBad() per keyframe.
int ref_semidense_loop_ex(int n, int W, int H, const uint8_t* im, const float* grad, const float* theta, const int32_t* edge,
                          const float* K4, const float* Tcw12, int n_cov, const int32_t* nbr_idx, const float* kp_angle,
                          const float* inv_depths, int n_inv, int first_id, int extra_ids, const int32_t* bad, float* depth,
                          float* sigma, float* checked, float* points, int32_t* flags)
{
    using namespace ORB_SLAM2;
    std::ostringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());  // the loop narrates to stdout
    const size_t P = (size_t)W * H;
    KeyFrame::nNextMappingId = (long unsigned int)first_id;
    std::vector<std::unique_ptr<KeyFrame> > kfs;
    std::vector<std::unique_ptr<MapPoint> > mps;
    Map map;
    for (int i = 0; i < n; i++) {
        kfs.emplace_back(new KeyFrame());
        KeyFrame* kf = kfs.back().get();
        kf->fx = K4[0]; kf->fy = K4[1]; kf->cx = K4[2]; kf->cy = K4[3];
        kf->mK = cv::Mat::zeros(3, 3, CV_32F);
        kf->mK.at<float>(0, 0) = K4[0]; kf->mK.at<float>(1, 1) = K4[1];
        kf->mK.at<float>(0, 2) = K4[2]; kf->mK.at<float>(1, 2) = K4[3]; kf->mK.at<float>(2, 2) = 1.0f;
        kf->mnMinX = 0; kf->mnMinY = 0; kf->mnMaxX = W; kf->mnMaxY = H;  // Frame.cc:584-590 without distortion
        cv::Mat T = cv::Mat::zeros(4, 4, CV_32F);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) T.at<float>(r, c) = Tcw12[(size_t)i * 12 + r * 4 + c];
        T.at<float>(3, 3) = 1.0f;
        kf->SetPose(T);
        // KeyFrame.cc:63-88
        kf->im_ = wrap_copy(im + i * P, H, W, CV_8U);
        kf->GradImg = wrap_copy(grad + i * P, H, W, CV_32F);
        kf->GradTheta = wrap_copy(theta + i * P, H, W, CV_32F);
        kf->mEdgeIndex = edge ? wrap_copy(edge + i * P, H, W, CV_32S) : cv::Mat::zeros(H, W, CV_32S);
        kf->depth_map_ = cv::Mat::zeros(H, W, CV_32F);
        kf->depth_sigma_ = cv::Mat::zeros(H, W, CV_32F);
        kf->depth_map_checked_ = cv::Mat::zeros(H, W, CV_32F);
        kf->SemiDensePointSets_ = cv::Mat::zeros(H, W * 3, CV_32F);
        kf->mvInvDepths.assign(inv_depths + (size_t)i * n_inv, inv_depths + (size_t)(i + 1) * n_inv);
        kf->mbBad = bad && bad[i];
        kf->IncreaseMappingId();
        map.mvKFs.push_back(kf);
    }
    KeyFrame::nNextMappingId += (long unsigned int)extra_ids;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n_cov; j++) kfs[i]->mvpOrderedConnectedKeyFrames.push_back(kfs[nbr_idx[(size_t)i * n_cov + j]].get());
    // one shared map point per unordered covisible pair
    std::vector<std::vector<char> > linked(n, std::vector<char>(n, 0));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n_cov; j++) {
            const int k = nbr_idx[(size_t)i * n_cov + j];
            if (linked[i][k]) continue;
            linked[i][k] = linked[k][i] = 1;
            mps.emplace_back(new MapPoint());
            for (int q : {i, k}) {
                kfs[q]->mvpMapPoints.push_back(mps.back().get());
                cv::KeyPoint kp;
                kp.angle = kp_angle[q];
                kfs[q]->mvKeysUn.push_back(kp);
            }
        }

    ProbabilityMapping pm(&map);
    pm.SemiDenseLoop();

    for (int i = 0; i < n; i++) {
        KeyFrame* kf = kfs[i].get();
        memcpy(depth + i * P, kf->depth_map_.data, P * 4);
        memcpy(sigma + i * P, kf->depth_sigma_.data, P * 4);
        memcpy(checked + i * P, kf->depth_map_checked_.data, P * 4);
        memcpy(points + i * P * 3, kf->SemiDensePointSets_.data, P * 12);
        flags[2 * i] = kf->semidense_flag_;
        flags[2 * i + 1] = kf->interKF_depth_flag_;
    }
    std::cout.rdbuf(old);
    return 0;
}

int ref_semidense_loop(int n, int W, int H, const uint8_t* im, const float* grad, const float* theta, const int32_t* edge,
                       const float* K4, const float* Tcw12, const int32_t* nbr_idx, const float* kp_angle,
                       const float* inv_depths, int n_inv, float* depth, float* sigma, float* checked, float* points,
                       int32_t* flags)
{
    // every keyframe eligible: ids 1..n, 11 newer keyframes, nobody bad, lists of exactly covisN
    return ref_semidense_loop_ex(n, W, H, im, grad, theta, edge, K4, Tcw12, covisN, nbr_idx, kp_angle, inv_depths, n_inv, 1, 11,
                                 NULL, depth, sigma, checked, points, flags);
}

// IntraKeyFrameDepthChecking / IntraKeyFrameDepthGrowing (:866-976; public, commented out of the shipped loop at
// :491-494) of the reference on caller planes, in place.  which: 0 = checking, 1 = growing.
int ref_intra(int which, int W, int H, float* depth, float* sigma, const float* grad)
{
    ORB_SLAM2::Map map;
    ProbabilityMapping pm(&map);
    const size_t P = (size_t)W * H;
    cv::Mat d = wrap_copy(depth, H, W, CV_32F), s = wrap_copy(sigma, H, W, CV_32F), g = wrap_copy(grad, H, W, CV_32F);
    if (which == 0) pm.IntraKeyFrameDepthChecking(d, s, g); else pm.IntraKeyFrameDepthGrowing(d, s, g);
    memcpy(depth, d.data, P * 4);
    memcpy(sigma, s.data, P * 4);
    return 0;
}

// ---- SaveSemiDensePoints (:136-192) of the reference on caller planes.  Writes results_line_segments/refshim/
// semi_pointcloud.obj below the current directory (the stand-in boost::filesystem creates nothing: the caller makes the
// directory).  rgb may be NULL (the keyframes then carry a grey rgb_ built from im).
int ref_save_semidense_points(int n, int W, int H, const uint8_t* im, const uint8_t* rgb, const float* sigma,
                              const float* checked, const float* points, const int32_t* flags, const int32_t* bad)
{
    using namespace ORB_SLAM2;
    std::ostringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    const size_t P = (size_t)W * H;
    std::vector<std::unique_ptr<KeyFrame> > kfs;
    Map map;
    for (int i = 0; i < n; i++) {
        kfs.emplace_back(new KeyFrame());
        KeyFrame* kf = kfs.back().get();
        kf->im_ = wrap_copy(im + i * P, H, W, CV_8U);
        kf->rgb_ = cv::Mat(H, W * 3, CV_8U);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W * 3; x++)
                kf->rgb_.at<uchar>(y, x) = rgb ? rgb[(i * P + (size_t)y * W) * 3 + x] : im[i * P + (size_t)y * W + x / 3];
        kf->depth_sigma_ = wrap_copy(sigma + i * P, H, W, CV_32F);
        kf->depth_map_checked_ = wrap_copy(checked + i * P, H, W, CV_32F);
        kf->SemiDensePointSets_ = wrap_copy(points + i * P * 3, H, W * 3, CV_32F);
        kf->semidense_flag_ = flags[2 * i] != 0;
        kf->interKF_depth_flag_ = flags[2 * i + 1] != 0;
        kf->mbBad = bad && bad[i];
        map.mvKFs.push_back(kf);
    }
    ProbabilityMapping pm(&map);
    pm.SaveSemiDensePoints();
    std::cout.rdbuf(old);
    return 0;
}

// ---- the online mode (#define OnlineLoop, :42; this file is also built with -DOnlineLoop into libref_pm_online.so).
// Keyframes arrive in three batches; after each batch the body of Run()'s loop (:223-226) is executed once:
//     SemiDenseLoop(); UpdateAllSemiDensePointSet();
// Between the batches the poses of the keyframes listed in moved[] are replaced (SetPose: local BA / loop closing,
// KeyFrame.cc:108-124 sets poseChanged).  After the last batch extra_ids more keyframes are "mapped" and the final
// SemiDenseLoop() of :244 runs.  n1 < n2 <= n: batch sizes are n1, n2 - n1, n - n2; Tcw_moved: [n_moved][12].
int ref_online_sequence(int n, int n1, int n2, int W, int H, const uint8_t* im, const float* grad, const float* theta,
                        const float* K4, const float* Tcw12, int n_cov, const int32_t* nbr_idx, const float* inv_depths,
                        int n_inv, int extra_ids, int n_moved, const int32_t* moved, const float* Tcw_moved, float* depth,
                        float* sigma, float* checked, float* points, int32_t* flags)
{
    using namespace ORB_SLAM2;
    std::ostringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    const size_t P = (size_t)W * H;
    KeyFrame::nNextMappingId = 1;
    std::vector<std::unique_ptr<KeyFrame> > kfs;
    Map map;
    auto pose = [](const float* T12) {
        cv::Mat T = cv::Mat::zeros(4, 4, CV_32F);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) T.at<float>(r, c) = T12[r * 4 + c];
        T.at<float>(3, 3) = 1.0f;
        return T;
    };
    for (int i = 0; i < n; i++) {
        kfs.emplace_back(new KeyFrame());
        KeyFrame* kf = kfs.back().get();
        kf->fx = K4[0]; kf->fy = K4[1]; kf->cx = K4[2]; kf->cy = K4[3];
        kf->mK = cv::Mat::zeros(3, 3, CV_32F);
        kf->mK.at<float>(0, 0) = K4[0]; kf->mK.at<float>(1, 1) = K4[1];
        kf->mK.at<float>(0, 2) = K4[2]; kf->mK.at<float>(1, 2) = K4[3]; kf->mK.at<float>(2, 2) = 1.0f;
        kf->mnMinX = 0; kf->mnMinY = 0; kf->mnMaxX = W; kf->mnMaxY = H;
        kf->SetPose(pose(Tcw12 + (size_t)i * 12));
        kf->im_ = wrap_copy(im + i * P, H, W, CV_8U);
        kf->GradImg = wrap_copy(grad + i * P, H, W, CV_32F);
        kf->GradTheta = wrap_copy(theta + i * P, H, W, CV_32F);
        kf->mEdgeIndex = cv::Mat::zeros(H, W, CV_32S);
        kf->depth_map_ = cv::Mat::zeros(H, W, CV_32F);
        kf->depth_sigma_ = cv::Mat::zeros(H, W, CV_32F);
        kf->depth_map_checked_ = cv::Mat::zeros(H, W, CV_32F);
        kf->SemiDensePointSets_ = cv::Mat::zeros(H, W * 3, CV_32F);
        kf->mvInvDepths.assign(inv_depths + (size_t)i * n_inv, inv_depths + (size_t)(i + 1) * n_inv);
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n_cov; j++) kfs[i]->mvpOrderedConnectedKeyFrames.push_back(kfs[nbr_idx[(size_t)i * n_cov + j]].get());
    ProbabilityMapping pm(&map);
    const int batch_end[3] = {n1, n2, n};
    int next = 0;
    for (int b = 0; b < 3; b++) {
        for (; next < batch_end[b]; next++) {
            kfs[next]->IncreaseMappingId();
            map.mvKFs.push_back(kfs[next].get());
        }
        if (b == 1)
            for (int m = 0; m < n_moved; m++) kfs[moved[m]]->SetPose(pose(Tcw_moved + (size_t)m * 12));
        pm.SemiDenseLoop();               // :224
        pm.UpdateAllSemiDensePointSet();  // :226
    }
    KeyFrame::nNextMappingId += (long unsigned int)extra_ids;
    pm.SemiDenseLoop();  // :244
    for (int i = 0; i < n; i++) {
        KeyFrame* kf = kfs[i].get();
        memcpy(depth + i * P, kf->depth_map_.data, P * 4);
        memcpy(sigma + i * P, kf->depth_sigma_.data, P * 4);
        memcpy(checked + i * P, kf->depth_map_checked_.data, P * 4);
        memcpy(points + i * P * 3, kf->SemiDensePointSets_.data, P * 12);
        flags[2 * i] = kf->semidense_flag_;
        flags[2 * i + 1] = kf->interKF_depth_flag_;
    }
    std::cout.rdbuf(old);
    return 0;
}

int ref_online_build(void)
{
#ifdef OnlineLoop
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
