// stand-in for include/ORBmatcher.h: only named from the dead helper GetInPlaneRotation (:1353-1440)
#pragma once
#include "cvstub.h"
namespace ORB_SLAM2 {
class ORBmatcher {
public:
    static int DescriptorDistance(const cv::Mat&, const cv::Mat&) { cv::stub_dead("ORBmatcher::DescriptorDistance"); }
};
}  // namespace ORB_SLAM2
