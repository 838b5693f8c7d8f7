// pnpmatch.h — drop-in for the reference's include/pnpmatch.h.  The Hamming work of
// poseEstimationPnP (src/pnpmatch.cc:61-199) and find_feature_matches (:253-300) runs on the
// device; findFundamentalMat stays with whatever OpenCV the integrator links (hook below); the PnP hook
// defaults to the device RANSAC (svo_pnp_ransac) and can be pointed back at cv::solvePnPRansac.
#pragma once
#include <functional>
#include <set>
#include <vector>
#include "frame.h"

class pnpmatch {
public:
    static cv::Mat Cur_Tcw;

    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);                      // :14-30
    static int poseEstimationPnP(frame *cframe, frame &lastframe, std::set<mappoint *> &localmappoints,
                                 cv::Mat &mVelocity, cv::Mat &K);                           // :33-251
    static void find_feature_matches(const cv::Mat &img_1, const cv::Mat &img_2,
                                     std::vector<cv::KeyPoint> &keypoints_1, std::vector<cv::KeyPoint> &keypoints_2,
                                     std::vector<cv::DMatch> &matches);                     // :253-300
    static int poseEstimation2D_2D(frame *CurrentFrame, frame &LastFrame, cv::Mat &K, cv::Mat &fundamental_matrix);  // :302-337

    // The two matching passes on their own (what poseEstimationPnP runs before the PnP solve).
    static int match_last_frame(frame *cur, frame &last, const cv::Mat &fundamental_matrix);        // pass 1
    static int match_local_map(frame *cur, std::set<mappoint *> &localmappoints);                   // pass 2

    // Hooks (unset: the step is skipped).  F from matched points (cv::findFundamentalMat, :336);
    // pose from 3D-2D pairs (cv::solvePnPRansac + Rodrigues, :227-247) returning a 4x4 CV_32F Tcl.
    // pnp_solver starts out as device_pnp: 100 samples, 8 px, like the reference's arguments.
    static std::function<cv::Mat(const std::vector<cv::Point2f> &, const std::vector<cv::Point2f> &)> fundamental_solver;
    static std::function<bool(const std::vector<cv::Mat> &pts3d, const std::vector<cv::Point2f> &pts2d, const cv::Mat &K,
                              cv::Mat &Tcl, int &inliers)> pnp_solver;
    static bool device_pnp(const std::vector<cv::Mat> &pts3d, const std::vector<cv::Point2f> &pts2d, const cv::Mat &K,
                           cv::Mat &Tcl, int &inliers);
    static std::vector<unsigned char> last_pnp_inliers;   // flags of the last device_pnp call, pts2d order

    // OPT-IN, changes results (north_star's "projection-guided" matching; off by default = the reference's brute-force
    // pass 2): when true, poseEstimationPnP predicts the pose as mVelocity * LastFrame.Tcw (the motion model the reference
    // computes in Tracking::GetVelocity, src/Tracking.cc:99-106, and leaves unused, src/pnpmatch.cc:53), projects the
    // local-map points with it on the device (svo_project_map) and pass 2 only looks inside each point's window
    // (half-size projection_th * scale[octave]).
    static bool use_projection;
    static float projection_th;
    static cv::Mat predicted_Tcw;                          // set by poseEstimationPnP, read by match_local_map
};
