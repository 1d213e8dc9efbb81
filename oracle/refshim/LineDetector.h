// stand-in for include/LineDetector.h: every entry point ProbabilityMapping.cc calls, as a no-op.  The real
// DetectEdgeMap fills kf->mEdgeIndex with edge-chain ids from the closed-source EDLib; here the driver pre-fills
// mEdgeIndex (all zeros = every pixel passes :454, or the test's own mask), so DetectEdgeMap leaves it alone.
#pragma once
#include <string>
#include <vector>

namespace ORB_SLAM2 { class KeyFrame; }
class Modeler;

// the real header pulls in boost::filesystem (used by SaveSemiDensePoints / WriteModel, which the driver never calls)
namespace boost { namespace filesystem {
struct path {
    std::string s;
    path() {}
    path(const std::string& s_) : s(s_) {}
    path(const char* s_) : s(s_) {}
};
inline path operator/(const path& a, const path& b) { return path(a.s + "/" + b.s); }
inline path current_path() { return path("."); }
inline bool exists(const path&) { return true; }
inline bool create_directories(const path&) { return true; }
template <class S> S& operator<<(S& o, const path& p) { o << p.s; return o; }
}}  // namespace boost::filesystem

class LineDetector {
public:
    std::vector<double> time_modeling;
    void DetectEdgeMap(ORB_SLAM2::KeyFrame*) {}
    void DetectLineSegments(ORB_SLAM2::KeyFrame*) {}
    void LineFitting(ORB_SLAM2::KeyFrame*) {}
    void MergeLines(ORB_SLAM2::KeyFrame*, Modeler*) {}
    void LineFittingOffline(std::vector<ORB_SLAM2::KeyFrame*>&, Modeler*) {}
    void LineFittingEDLinesOffline(std::vector<ORB_SLAM2::KeyFrame*>&) {}
    void RunLine3Dpp(std::vector<ORB_SLAM2::KeyFrame*>&) {}
    void SaveAllLineSegments() {}
    void SaveClusteredSegments() {}
    void Summary() {}
    void Reset() {}
    std::string GetStringDateTime() { return "refshim"; }
};
