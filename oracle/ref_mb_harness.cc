// ref_mb_harness.cc — TEST INFRASTRUCTURE.  The reference's OWN stereo path, the one Tracking::Track calls
// (src/Tracking.cc:226-228): frame::MB (src/frame.cc:82-91) -> Thirdparty/MB/MSA.cpp (minimum-spanning-tree cost
// aggregation) + ctmf.c, then computekeypoint_r and disp2Depth, all compiled unmodified (oracle/Makefile: `make ref_mb`).
// tests/test_oracle_stereo.py reads the dense disparity it produces at the keypoints and compares the sparse stereo
// oracle with it.  Nothing here is used by the product.
#include <frame.h>
#include <mappoint.h>

#include <cstring>
#include <vector>

extern "C" {

// L, R: 8-bit BGR images (MSA::init reads three bytes per pixel); out: h x w float disparities (frame::MB's return value)
int ref_mb_disparity(const unsigned char *L, const unsigned char *R, int w, int h, float *out)
{
    cv::Mat l(h, w, CV_8UC3), r(h, w, CV_8UC3), none, det(8, 8, CV_8U), K = cv::Mat::eye(3, 3, CV_32F);
    std::memcpy(l.data, L, (size_t)w * h * 3); std::memcpy(r.data, R, (size_t)w * h * 3);
    std::vector<std::vector<int> > boxes;
    double ts = 0.0;
    float bf = 1.f;
    frame *f = new frame(l, r, none, det, ts, K, bf, boxes);
    cv::Mat d = f->MB(f->leftimg, f->rightimg);                  // src/frame.cc:82-91
    if (d.rows != h || d.cols != w || d.type() != CV_32F) { delete f; return -1; }
    for (int y = 0; y < h; ++y) std::memcpy(out + (size_t)y * w, d.ptr<float>(y), sizeof(float) * (size_t)w);
    delete f;
    return 0;
}

}
