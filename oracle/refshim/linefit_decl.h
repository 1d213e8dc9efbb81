// linefit_decl.h - pre-included in front of the text of LineDetector.cc:20-29 + :579-840 (extracted at build time into
// oracle/_ref/, never into the repository): the stand-in types and a LineDetector class that declares exactly the eight
// methods that text defines (include/LineDetector.h:76-84).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "KeyFrame.h"
#include "cvstub_linefit.h"

class LineDetector {
public:
    void ClosestPointOnLine(float a, float b, float c, int x, int y, float& cx, float& cy);
    int CountDepth(Pixel* pixelChain, int length, ORB_SLAM2::KeyFrame* kf);
    void LeastSquaresLineFit(Pixel* pixelChain, int initLength, float& u1, float& u2, float& u3, float& lineFitError);
    void LeastSquaresDepthFit(Pixel* pixelChain, int initLength, float la, float lb, float lc, float& u1, float& u2, float& depthFitError,
                              ORB_SLAM2::KeyFrame* kf);
    float ComputePointDistance2Line(float a, float b, float c, Pixel pixel);
    float ComputePointDepth2Line(float a, float b, float c, float alpha, float beta, Pixel* pixelChain, Pixel pixel, ORB_SLAM2::KeyFrame* kf);
    void LineFit(Pixel* pixelChain, int noPixels, ORB_SLAM2::KeyFrame* kf);
};
