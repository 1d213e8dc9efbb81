// stand-in for include/MapPoint.h
#pragma once
#include "cvstub.h"
namespace ORB_SLAM2 {
class MapPoint {
public:
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    bool isBad() { return false; }
    cv::Mat mWorldPos;
};
}  // namespace ORB_SLAM2
