// stand-in for include/KeyFrame.h: the members and accessors of ORB_SLAM2::KeyFrame that ProbabilityMapping.cc
// uses, with the reference's names and behaviour (KeyFrame.cc:63-88, 108-124, 151-161, 756-856).
#pragma once
// (standard headers the real OpenCV / Eigen / Boost / DBoW2 headers bring in transitively)
#include <unistd.h>

#include <algorithm>
#include <cassert>
#include <climits>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <vector>

#include "Map.h"
#include "MapPoint.h"
#include "cvstub.h"

// The real KeyFrame.h includes Thirdparty/DBoW2, whose TemplatedVocabulary.h:36 has `using namespace std;`.
// ProbabilityMapping.cc relies on it (unqualified cout, mutex, unique_lock, and abs()/sqrt() resolving to the
// float overloads), so the compile environment must have it too.
using namespace std;

// stand-in for Thirdparty/EDTest/EdgeMap.h (the closed-source Edge Drawing library's result type, KeyFrame.h:175)
struct Pixel { int r, c; };
struct EdgeSegment { Pixel* pixels; int noPixels; };
struct EdgeMap {
    int width, height;
    unsigned char* edgeImg;
    Pixel* pixels;
    EdgeSegment* segments;
    int noSegments;
    EdgeMap(int w, int h) : width(w), height(h), edgeImg(new unsigned char[w * h]), pixels(new Pixel[w * h]),
                            segments(new EdgeSegment[w * h]), noSegments(0) {}
    ~EdgeMap() { delete[] edgeImg; delete[] pixels; delete[] segments; }
};

namespace DBoW2 {
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;
}

namespace ORB_SLAM2 {

class KeyFrame {
public:
    KeyFrame() : fx(0), fy(0), cx(0), cy(0), mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0), semidense_flag_(false),
                 interKF_depth_flag_(false), I_stddev(20.0f), poseChanged(false), mnMappingId(0), mEdgeMap(NULL), mbBad(false) {}

    void SetPose(const cv::Mat& Tcw_)  // KeyFrame.cc:108-124
    {
        unique_lock<mutex> lock(mMutexPose);
        Tcw = Tcw_.clone();
        cv::Mat Rcw = Tcw.rowRange(0, 3).colRange(0, 3);
        cv::Mat tcw = Tcw.rowRange(0, 3).col(3);
        cv::Mat Rwc = Rcw.t();
        Ow = -Rwc * tcw;
        Twc = cv::Mat::zeros(4, 4, CV_32F);
        Twc.at<float>(3, 3) = 1.0f;  // cv::Mat::eye
        Rwc.copyTo(Twc.rowRange(0, 3).colRange(0, 3));
        Ow.copyTo(Twc.rowRange(0, 3).col(3));
        poseChanged = true;  // KeyFrame.cc:123
    }
    cv::Mat GetPose() { unique_lock<mutex> lock(mMutexPose); return Tcw.clone(); }
    cv::Mat GetPoseInverse() { unique_lock<mutex> lock(mMutexPose); return Twc.clone(); }
    cv::Mat GetRotation() { unique_lock<mutex> lock(mMutexPose); return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { unique_lock<mutex> lock(mMutexPose); return Tcw.rowRange(0, 3).col(3).clone(); }
    cv::Mat GetCalibrationMatrix() const { return mK.clone(); }
    cv::Mat GetImage() { return im_.clone(); }
    cv::Mat GetDescriptors() { return cv::Mat(); }
    DBoW2::FeatureVector GetFeatureVector() { return DBoW2::FeatureVector(); }
    std::vector<cv::KeyPoint> GetKeyPointsUn() const { return mvKeysUn; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
    std::vector<KeyFrame*> GetBestCovisibilityKeyFrames(const int& N)
    {
        std::vector<KeyFrame*> v = mvpOrderedConnectedKeyFrames;
        if ((int)v.size() > N) v.resize(N);
        return v;
    }
    std::vector<float> GetAllPointDepths() { return mvInvDepths; }  // KeyFrame.cc:756-787: sorted inverse depths
    bool isBad() { return mbBad; }
    bool MappingIdDelay() { unique_lock<mutex> lock(mMutexMappingId); if (mnMappingId == 0) return false; return (nNextMappingId - mnMappingId) > 10; }
    void IncreaseMappingId() { unique_lock<mutex> lock(mMutexMappingId); mnMappingId = nNextMappingId++; }
    bool Mapped() { unique_lock<mutex> lock(mMutexMappingId); return mnMappingId != 0; }
    bool PoseChanged() { unique_lock<mutex> lock(mMutexPose); return poseChanged; }
    void SetPoseChanged(bool b) { unique_lock<mutex> lock(mMutexPose); poseChanged = b; }
    void SetNotEraseSemiDense() {}
    void SetEraseSemiDense() {}

    float fx, fy, cx, cy;
    int mnMinX, mnMinY, mnMaxX, mnMaxY;
    cv::Mat mK;
    // KeyFrame.h:155-175
    cv::Mat im_, rgb_;
    bool semidense_flag_, interKF_depth_flag_;
    cv::Mat GradImg, GradTheta;
    float I_stddev;
    cv::Mat depth_map_, depth_sigma_, depth_map_checked_;
    bool poseChanged;
    std::mutex mMutexSemiDensePoints;
    cv::Mat SemiDensePointSets_;
    static long unsigned int nNextMappingId;
    long unsigned int mnMappingId;
    std::mutex mMutexMappingId;
    cv::Mat mEdgeIndex;
    cv::Mat mLinesSeg, mLines3D;  // KeyFrame.h:172-173
    EdgeMap* mEdgeMap;            // KeyFrame.h:175 (Thirdparty/EDTest/EdgeMap.h)

    // filled by the driver in place of the ORB-SLAM2 graph
    std::vector<KeyFrame*> mvpOrderedConnectedKeyFrames;
    std::vector<float> mvInvDepths;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<cv::KeyPoint> mvKeysUn;
    bool mbBad;

private:
    cv::Mat Tcw, Twc, Ow;
    std::mutex mMutexPose;
};

}  // namespace ORB_SLAM2
