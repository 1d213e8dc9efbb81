// stand-in for include/Map.h
#pragma once
#include <vector>
class Modeler;
namespace ORB_SLAM2 {
class KeyFrame;
class Map {
public:
    std::vector<KeyFrame*> GetAllKeyFrames() { return mvKFs; }   // Map.cc:79-83 (a std::set there: pointer order)
    void SetModeler(Modeler*) {}
    std::vector<KeyFrame*> mvKFs;
};
}  // namespace ORB_SLAM2
