// linefit_driver.cc - C entry point that runs the REFERENCE'S OWN LineDetector::LineFit (text of
// /root/reference/src/LineDetector.cc:578-840, extracted at build time into oracle/_ref/linefit_part.cc and compiled against the
// stand-ins of this directory) over caller-supplied edge chains, the way LineDetector::LineFitting does (:884-900).
// TEST INFRASTRUCTURE ONLY: oracle/_ref/libref_linefit.so, used by oracle/make_linefit_golden.py.
#include <cstdint>
#include <vector>

#include "linefit_decl.h"

long unsigned int ORB_SLAM2::KeyFrame::nNextMappingId = 1;

extern "C" {

// chains: pixels (r, c) int32 pairs, chain k = pix[off[k] .. off[k+1]).  Returns the number of lines; out_seg [cap][4],
// out_xyz [cap][6], out_chain [cap] receive the rows of kf->mLinesSeg / kf->mLines3D in push_back order with their chain.
int ref_line_fitting(int W, int H, const float* checked, const float* sigma, const float* K4, const float* Tcw12, int n_chains,
                     const int32_t* off, const int32_t* rc, int cap, float* out_seg, float* out_xyz, int32_t* out_chain)
{
    ORB_SLAM2::KeyFrame kf;
    kf.fx = K4[0]; kf.fy = K4[1]; kf.cx = K4[2]; kf.cy = K4[3];
    cv::Mat T = cv::Mat::zeros(4, 4, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T.at<float>(r, c) = Tcw12[4 * r + c];
    T.at<float>(3, 3) = 1.0f;
    kf.SetPose(T);
    kf.depth_map_checked_ = cv::Mat(H, W, CV_32F);
    kf.depth_sigma_ = cv::Mat(H, W, CV_32F);
    memcpy(kf.depth_map_checked_.data, checked, (size_t)W * H * 4);
    memcpy(kf.depth_sigma_.data, sigma, (size_t)W * H * 4);
    LineDetector ld;
    int n = 0;
    for (int k = 0; k < n_chains; k++) {
        std::vector<Pixel> px((size_t)(off[k + 1] - off[k]));
        for (size_t i = 0; i < px.size(); i++) { px[i].r = rc[2 * (off[k] + i)]; px[i].c = rc[2 * (off[k] + i) + 1]; }
        ld.LineFit(px.data(), (int)px.size(), &kf);  // LineFitting's loop body (:893-895)
        for (; n < kf.mLines3D.rows; n++) {
            if (n >= cap) return -1;
            for (int q = 0; q < 4; q++) out_seg[4 * n + q] = kf.mLinesSeg.at<float>(n, q);
            for (int q = 0; q < 6; q++) out_xyz[6 * n + q] = kf.mLines3D.at<float>(n, q);
            out_chain[n] = k;
        }
    }
    return n;
}

}  // extern "C"
