// stands in for Thirdparty/MB/MSA.h: the dense MSA stereo solver is out of scope;
// frame::MB must still compile.  The harness sets frame::dispimg directly instead of calling MB.
#pragma once
#include "../../minicv.hpp"
class MSA {
public:
    cv::Mat solve(cv::Mat &, cv::Mat &, int, int, bool) { return cv::Mat(); }
};
