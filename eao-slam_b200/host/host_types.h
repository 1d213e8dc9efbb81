// host_types.h — minimal stand-ins for the reference types the semi-dense path touches, with the SAME
// member names, so that ProbabilityMapping.h compiles unchanged against either these or the real
// ORB_SLAM2::KeyFrame / ORB_SLAM2::Map / cv::Mat (define SDM_HOST_WITH_ORBSLAM2 for the latter).
//   Mat       <- cv::Mat subset: rows, cols, step, data, empty(), ptr<T>(row), at<T>(r,c), clone()
//   KeyFrame  <- include/KeyFrame.h:125-175 (semi-dense planes and flags, pose, covisibility, mapping id)
//   Map       <- include/Map.h (GetAllKeyFrames)
// OpenCV / Eigen / Boost are not installed in this repository's build image, hence the stand-ins.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace sdm_host {

class Mat {
public:
    Mat() : rows(0), cols(0), step(0), data(NULL) {}
    Mat(int r, int c, size_t elem_bytes) : rows(r), cols(c), step((size_t)c * elem_bytes), data(NULL)
    {
        buf_.reset(new uint8_t[(size_t)r * step](), std::default_delete<uint8_t[]>());
        data = buf_.get();
    }
    // wrap caller memory (like cv::Mat(rows, cols, type, ptr, step))
    Mat(int r, int c, size_t elem_bytes, void* p, size_t step_bytes) : rows(r), cols(c), step(step_bytes ? step_bytes : (size_t)c * elem_bytes), data((uint8_t*)p) {}
    bool empty() const { return data == NULL || rows == 0 || cols == 0; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    Mat clone() const
    {
        if (empty()) return Mat();
        Mat m(rows, 1, step);
        m.cols = cols;
        std::memcpy(m.data, data, (size_t)rows * step);
        return m;
    }
    int rows, cols;
    size_t step;
    uint8_t* data;

private:
    std::shared_ptr<uint8_t> buf_;
};

inline Mat zeros32f(int rows, int cols) { return Mat(rows, cols, sizeof(float)); }

struct KeyPoint { float angle; };
struct MapPoint { int id; };

class KeyFrame {
public:
    KeyFrame() : fx(0), fy(0), cx(0), cy(0), semidense_flag_(false), interKF_depth_flag_(false), I_stddev(20.0f),
                 poseChanged(false), mnMappingId(0), mbBad(false), mPinned(0) {}

    // KeyFrame.cc:63-88: take the image planes, zero-initialise the depth planes
    void SetPlanes(const Mat& im, const Mat& grad, const Mat& theta)
    {
        im_ = im; GradImg = grad; GradTheta = theta;
        depth_map_ = zeros32f(im.rows, im.cols);
        depth_sigma_ = zeros32f(im.rows, im.cols);
        depth_map_checked_ = zeros32f(im.rows, im.cols);
        SemiDensePointSets_ = Mat(im.rows, im.cols * 3, sizeof(float));
    }

    Mat GetPose() { std::unique_lock<std::mutex> l(mMutexPose); return Tcw.clone(); }
    void SetPose(const float T[12])
    {
        std::unique_lock<std::mutex> l(mMutexPose);
        Tcw = Mat(4, 4, sizeof(float));
        for (int i = 0; i < 12; i++) Tcw.ptr<float>(0)[i] = T[i];
        Tcw.at<float>(3, 3) = 1.0f;
        poseChanged = true;  // KeyFrame.cc:123
    }
    bool isBad() const { return mbBad; }
    // KeyFrame.cc:789-806
    bool MappingIdDelay() { std::unique_lock<std::mutex> l(mMutexMappingId); return mnMappingId != 0 && (nNextMappingId() - mnMappingId) > 10; }
    void IncreaseMappingId() { std::unique_lock<std::mutex> l(mMutexMappingId); mnMappingId = nNextMappingId()++; }
    bool Mapped() { std::unique_lock<std::mutex> l(mMutexMappingId); return mnMappingId != 0; }
    bool PoseChanged() { std::unique_lock<std::mutex> l(mMutexPose); return poseChanged; }
    void SetPoseChanged(bool b) { std::unique_lock<std::mutex> l(mMutexPose); poseChanged = b; }
    void SetNotEraseSemiDense() { ++mPinned; }   // KeyFrame.cc:836-856 (erase pins)
    void SetEraseSemiDense() { if (mPinned > 0) --mPinned; }
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
    std::vector<float> GetAllPointDepths() { return mvInvDepths; }   // sorted inverse depths, KeyFrame.cc:756-787
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<KeyPoint> GetKeyPointsUn() const { return mvKeysUn; }

    // KeyFrame.h:155-175
    Mat im_, rgb_, GradImg, GradTheta, depth_map_, depth_sigma_, depth_map_checked_, SemiDensePointSets_, mEdgeIndex;
    float fx, fy, cx, cy;
    bool semidense_flag_, interKF_depth_flag_;
    float I_stddev;
    bool poseChanged;
    std::mutex mMutexSemiDensePoints;
    unsigned long mnMappingId;
    static unsigned long& nNextMappingId() { static unsigned long n = 1; return n; }

    // what the stand-alone harness fills in place of the ORB-SLAM2 graph
    std::vector<KeyFrame*> mvpOrderedConnectedKeyFrames;
    std::vector<float> mvInvDepths;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<KeyPoint> mvKeysUn;
    bool mbBad;

private:
    Mat Tcw;
    std::mutex mMutexPose, mMutexMappingId;
    int mPinned;
};

class Map {
public:
    void AddKeyFrame(KeyFrame* kf) { mspKeyFrames.push_back(kf); }
    std::vector<KeyFrame*> GetAllKeyFrames() { return mspKeyFrames; }

private:
    std::vector<KeyFrame*> mspKeyFrames;
};

}  // namespace sdm_host
