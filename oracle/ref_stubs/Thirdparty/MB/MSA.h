// stands in for Thirdparty/MB/MSA.h: the dense MSA stereo solver is out of scope;
// frame::MB must still compile and, when Tracking::Track itself is driven (oracle/ref_g2o_harness.cc), return a
// disparity image: solve() hands back the image the harness deposited for this frame (the synthetic dense
// disparity of the test sequence), as float.  The other harness sets frame::dispimg directly and never calls MB.
#pragma once
#include "../../minicv.hpp"
inline cv::Mat &minicv_next_disparity() { static cv::Mat m; return m; }
class MSA {
public:
    cv::Mat solve(cv::Mat &, cv::Mat &, int, int, bool) { return minicv_next_disparity().clone(); }
};
