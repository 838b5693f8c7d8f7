// ref_g2o_harness.cc — TEST INFRASTRUCTURE.  Drives the reference's OWN pose optimisation: Optimizer::PoseOptimization
// (src/Optimizer.cc:15-86) with src/convert.cc and the vendored g2o (Thirdparty/g2o/g2o/{core,types,stuff}), all compiled
// unmodified where they lie under /root/reference (oracle/Makefile: `make ref_g2o`) against ref_stubs_g2o/minieigen.hpp
// (Eigen is not installed in this image) and ref_stubs/minicv.hpp.  tests/test_ref_pin_pose.py compares
// oracle/svo_pose_oracle.c with it.  Nothing here is used by the product.
#include <frame.h>
#include <mappoint.h>
#include <Optimizer.h>

#include <cstring>
#include <vector>

extern "C" {

// Tcw16: row-major 4x4 initial pose; n <= 500 keypoints (frame::N is fixed, src/frame.cc:54); xy: 2 n; has[i] != 0: the
// keypoint owns a map point at Xw[3 i ..]; K9: row-major intrinsics.  Returns PoseOptimization's return value (the number
// of correspondences) or -1; Tcw_out16 = the frame's pose after SetPose(optimised).
int ref_pose_optimize(const float *Tcw16, int n, const float *xy, const float *Xw, const unsigned char *has, const float *K9,
                      float *Tcw_out16)
{
    cv::Mat L(8, 8, CV_8U), R(8, 8, CV_8U), none, det(8, 8, CV_8U), K(3, 3, CV_32F);
    std::memset(L.data, 0, 64); std::memset(R.data, 0, 64); std::memset(det.data, 0, 64);
    for (int i = 0; i < 9; ++i) K.at<float>(i / 3, i % 3) = K9[i];
    std::vector<std::vector<int> > boxes;
    double ts = 0.0;
    float bf = 1.f;
    frame *f = new frame(L, R, none, det, ts, K, bf, boxes);          // src/frame.cc:36-64
    if (n > f->N) { delete f; return -1; }
    f->keypoints_l.resize((size_t)f->N);
    f->f_descriptor = cv::Mat(f->N, 32, CV_8U);
    std::memset(f->f_descriptor.data, 0, (size_t)f->N * 32);
    for (int i = 0; i < n; ++i) {
        f->keypoints_l[(size_t)i].pt.x = xy[2 * i]; f->keypoints_l[(size_t)i].pt.y = xy[2 * i + 1];
        if (!has[i]) continue;
        cv::Mat pos(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) pos.at<float>(k) = Xw[3 * i + k];
        f->MapPoints[(size_t)i] = new mappoint(pos, f, i);            // src/mappoint.cc:10-15
    }
    cv::Mat T(4, 4, CV_32F);
    std::memcpy(T.data, Tcw16, 64);
    f->SetPose(T);
    const int r = Optimizer::PoseOptimization(f);                     // src/Optimizer.cc:15-86
    std::memcpy(Tcw_out16, f->Tcw.data, 64);
    return r;
}

}
