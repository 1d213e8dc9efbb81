// test_shim.cpp — drives the C++ drop-in class (eao-slam_b200/host/ProbabilityMapping.h) the way
// System.cc / the reference's own loop would: keyframes with cv::Mat-like planes in a Map, then
// SemiDenseLoop(); plus a few single-method calls.  Reads a scene dump written by
// tests/test_cpp_shim.py, writes the resulting planes back for comparison with the oracle.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../eao-slam_b200/host/ProbabilityMapping.h"

using sdm_host::KeyFrame;
using sdm_host::Map;
using sdm_host::Mat;

static void rd(FILE* f, void* p, size_t n)
{
    if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: test_shim scene.bin out.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("scene"); return 2; }
    // n, W, H, covisN, pitch_pad, n_probe, n_cov (covisibility list length >= covisN), first mapping id, keyframes
    // mapped after ours (the reference's gating: KeyFrame.cc:789-806)
    int32_t hdr[9];
    rd(f, hdr, sizeof(hdr));
    const int n = hdr[0], W = hdr[1], H = hdr[2], N = hdr[3], pad = hdr[4], n_probe = hdr[5];
    const int n_cov = hdr[6], first_id = hdr[7], extra_ids = hdr[8];
    KeyFrame::nNextMappingId() = (unsigned long)first_id;
    float K[4];
    rd(f, K, sizeof(K));
    std::vector<std::unique_ptr<KeyFrame>> kfs;
    std::vector<std::vector<int32_t>> nbr(n);
    Map map;
    for (int i = 0; i < n; i++) {
        kfs.emplace_back(new KeyFrame());
        KeyFrame* kf = kfs.back().get();
        float Tcw[12];
        rd(f, Tcw, sizeof(Tcw));
        kf->SetPose(Tcw);
        kf->fx = K[0]; kf->fy = K[1]; kf->cx = K[2]; kf->cy = K[3];
        // pitched planes, like a cv::Mat ROI: W + pad elements per row
        Mat im(H, W + pad, 1), g(H, W + pad, 4), t(H, W + pad, 4);
        for (int y = 0; y < H; y++) rd(f, im.ptr<uint8_t>(y), (size_t)W);
        for (int y = 0; y < H; y++) rd(f, g.ptr<float>(y), (size_t)W * 4);
        for (int y = 0; y < H; y++) rd(f, t.ptr<float>(y), (size_t)W * 4);
        im.cols = g.cols = t.cols = W;
        kf->SetPlanes(im, g, t);
        int32_t nd;
        rd(f, &nd, 4);
        kf->mvInvDepths.resize(nd);
        rd(f, kf->mvInvDepths.data(), (size_t)nd * 4);
        nbr[i].resize(n_cov);
        rd(f, nbr[i].data(), (size_t)n_cov * 4);
        int32_t bad;
        rd(f, &bad, 4);
        kf->mbBad = bad != 0;
        kf->IncreaseMappingId();
        map.AddKeyFrame(kf);
    }
    std::vector<int32_t> probes((size_t)n_probe * 4);  // kf1, kf2, x, y
    rd(f, probes.data(), probes.size() * 4);
    fclose(f);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n_cov; j++) kfs[i]->mvpOrderedConnectedKeyFrames.push_back(kfs[nbr[i][j]].get());
    {   // keyframes "mapped" after ours: MappingIdDelay() needs more than 10 newer ones (KeyFrame.cc:789-794)
        KeyFrame dummy;
        for (int i = 0; i < extra_ids; i++) dummy.IncreaseMappingId();
    }

    ProbabilityMapping pm(&map);
    pm.SetCovisN(N);
    if (const char* e = getenv("SDM_SHIM_DEVICE_PLANES")) pm.SetProducePlanesOnDevice(e[0] == '1');
    pm.RequestFinish();
    pm.Run();  // = SemiDenseLoop() once, as in the reference's offline mode
    if (!pm.isFinished()) return 3;

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("out"); return 2; }
    for (int i = 0; i < n; i++) {
        KeyFrame* kf = kfs[i].get();
        int32_t flags[2] = {kf->semidense_flag_, kf->interKF_depth_flag_};
        fwrite(flags, 4, 2, o);
        for (int y = 0; y < H; y++) fwrite(kf->depth_map_.ptr<float>(y), 4, W, o);
        for (int y = 0; y < H; y++) fwrite(kf->depth_sigma_.ptr<float>(y), 4, W, o);
        for (int y = 0; y < H; y++) fwrite(kf->depth_map_checked_.ptr<float>(y), 4, W, o);
        for (int y = 0; y < H; y++) fwrite(kf->SemiDensePointSets_.ptr<float>(y), 4, (size_t)3 * W, o);
    }
    // single-method calls, argument order of the reference
    for (int p = 0; p < n_probe; p++) {
        KeyFrame *k1 = kfs[probes[4 * p]].get(), *k2 = kfs[probes[4 * p + 1]].get();
        const int x = probes[4 * p + 2], y = probes[4 * p + 3];
        float mind, maxd, umin = 0, umax = 0, bu = 0, bv = 0;
        pm.StereoSearchConstraints(k1, &mind, &maxd);
        pm.GetSearchRange(umin, umax, x, y, mind, maxd, k1, k2);
        ProbabilityMapping::depthHo dh;
        pm.EpipolarSearch(k1, k2, x, y, (float)k1->im_.at<uint8_t>(y, x), mind, maxd, &dh, Mat(), bu, bv,
                          k1->GradTheta.at<float>(y, x), 0.0f);
        float rec[8] = {mind, maxd, umin, umax, dh.depth, dh.sigma, dh.supported ? 1.0f : 0.0f, bu};
        fwrite(rec, 4, 8, o);
    }
    {   // InverseDepthHypothesisFusion on a hand-made set
        std::vector<ProbabilityMapping::depthHo> h(6);
        const float d[6] = {0.50f, 0.51f, 0.49f, 0.505f, 0.9f, 0.495f}, s[6] = {0.02f, 0.02f, 0.03f, 0.01f, 0.02f, 0.02f};
        for (int i = 0; i < 6; i++) { h[i].depth = d[i]; h[i].sigma = s[i]; h[i].supported = true; }
        ProbabilityMapping::depthHo out;
        pm.InverseDepthHypothesisFusion(h, out);
        float rec[3] = {out.depth, out.sigma, out.supported ? 1.0f : 0.0f};
        fwrite(rec, 4, 3, o);
    }
    {   // IntraKeyFrameDepthChecking on copies of keyframe 2's planes (cv::Mat& semantics: in place)
        Mat d = kfs[2]->depth_map_.clone(), s = kfs[2]->depth_sigma_.clone();
        pm.IntraKeyFrameDepthChecking(d, s, kfs[2]->GradImg);
        for (int y = 0; y < H; y++) fwrite(d.ptr<float>(y), 4, W, o);
        for (int y = 0; y < H; y++) fwrite(s.ptr<float>(y), 4, W, o);
    }
    {   // the exporters' point filter (SaveSemiDensePoints / DrawSemiDense), compacted on the device
        std::vector<sdm_point> pts;
        std::vector<uint64_t> counts;
        const size_t np = pm.ExportSemiDensePoints(0.02, pts, NULL, &counts);
        float hdr2[2] = {(float)np, (float)counts.size()};
        fwrite(hdr2, 4, 2, o);
        for (size_t i = 0; i < counts.size(); i++) { float c = (float)counts[i]; fwrite(&c, 4, 1, o); }
        fwrite(pts.data(), sizeof(sdm_point), np, o);
    }
    fclose(o);
    sdm_timing t = pm.LastTiming();
    printf("shim ok: pass1 scan %.3f ms, pass2 %.3f ms\n", t.pass1_scan_ms, t.pass2_ms);
    return 0;
}
