// stand-in for include/Modeler.h (CARV consumer, out of scope): no-ops
#pragma once
#include <string>
namespace ORB_SLAM2 { class KeyFrame; class Map; }
class Modeler {
public:
    explicit Modeler(ORB_SLAM2::Map*) {}
    bool CheckNewTranscriptEntry() { return false; }
    void RunRemainder() {}
    void RunOnce() {}
    void UpdateModel() {}
    void WriteModel(const std::string&) {}
    void AddLineSegmentKeyFrameEntry(ORB_SLAM2::KeyFrame*) {}
};
