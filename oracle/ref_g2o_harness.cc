// ref_g2o_harness.cc — TEST INFRASTRUCTURE.  Drives the reference's OWN pose optimisation: Optimizer::PoseOptimization
// (src/Optimizer.cc:15-86) with src/convert.cc and the vendored g2o (Thirdparty/g2o/g2o/{core,types,stuff}), all compiled
// unmodified where they lie under /root/reference (oracle/Makefile: `make ref_g2o`) against ref_stubs_g2o/minieigen.hpp
// (Eigen is not installed in this image) and ref_stubs/minicv.hpp.  tests/test_ref_pin_pose.py compares
// oracle/svo_pose_oracle.c with it.  Nothing here is used by the product.
#include <frame.h>
#include <mappoint.h>
#include <Optimizer.h>
#include <Tracking.h>
#include <view.h>
#include "Thirdparty/MB/MSA.h"

#include <cstring>
#include <fstream>
#include <vector>

// The viewer (src/view.cc: Pangolin / OpenGL drawing) is out of scope and not compiled; Tracking::SaveTrajectoryAndDraw
// calls it once per frame.
void View::DrawGraph(frame &, frame *) {}
void View::DrawMappoints(set<mappoint *> &, int) {}

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

// ---- Tracking::Track itself (src/Tracking.cc:180-252, compiled unmodified) -------------------------------------------
// settings: a YAML file with Camera.fx / fy / cx / cy / bf (Tracking::Tracking reads it, src/Tracking.cc:22-39).
void *ref_tracking_new(const char *settings)
{
    Tracking::frame_num = 0;                 // process-wide statics of the reference (src/Tracking.cc:18-19)
    Tracking::LocalMapPoints.clear();
    return new Tracking(std::string(settings));
}
void ref_tracking_free(void *t) { delete (Tracking *)t; }
// One call of Tracking::Track.  disp: the dense disparity image frame::MB's solver returns for this pair (the MSA solver
// itself is out of scope, ref_stubs/Thirdparty/MB/MSA.h); images 8-bit, ch channels.
void ref_tracking_track(void *t, const unsigned char *L, const unsigned char *R, int w, int h, int ch, const float *disp,
                        const float *K9, float bf, double ts, const int *boxes, int nboxes)
{
    cv::Mat l(h, w, CV_MAKETYPE(CV_8U, ch)), r(h, w, CV_MAKETYPE(CV_8U, ch)), none, K(3, 3, CV_32F), d(h, w, CV_32F);
    std::memcpy(l.data, L, (size_t)w * h * ch); std::memcpy(r.data, R, (size_t)w * h * ch);
    std::memcpy(d.data, disp, sizeof(float) * (size_t)w * h);
    for (int i = 0; i < 9; ++i) K.at<float>(i / 3, i % 3) = K9[i];
    cv::Mat det = l.clone();
    minicv_next_disparity() = d;
    std::vector<std::vector<int> > bx;
    for (int k = 0; k < nboxes; ++k) bx.push_back(std::vector<int>(boxes + 4 * k, boxes + 4 * k + 4));
    std::ofstream f("/dev/null"), f2("/dev/null");
    pangolin::OpenGlMatrix M;
    ((Tracking *)t)->Track(l, r, none, det, ts, K, bf, f, f2, M, bx);
}
void *ref_tracking_current(void *t) { return ((Tracking *)t)->currentframe; }
void *ref_tracking_last(void *t) { return &((Tracking *)t)->lastframe; }
void *ref_tracking_localmap() { return &Tracking::LocalMapPoints; }
int ref_tracking_frame_num() { return Tracking::frame_num; }
void ref_tracking_velocity(void *t, float *V16)
{
    const cv::Mat &V = ((Tracking *)t)->Velocity;
    if (!V.empty()) for (int i = 0; i < 16; ++i) V16[i] = V.at<float>(i / 4, i % 4);
}

}
