// frame.h — drop-in for the reference's include/frame.h: same class name, same public data
// members, same method names; the work behind featuredetect / MB / disp2Depth runs on the B200
// through the C ABI (include/svo_b200.h).  Differences from the reference, all deliberate:
//   * nfeatures / nlevels / scaleFactor are constructor-time settings (frame::configure) instead
//     of the hard-wired ORB::create() defaults and N = 500 (src/frame.cc:54,77);
//   * MB() returns a SPARSE disparity image (values only at left keypoints) from the row-band
//     Hamming + SAD stage instead of running the dense MSA solver (src/frame.cc:82-91);
//   * nothing is leaked per frame (src/frame.cc:86).
#pragma once
#include <set>
#include <string>
#include <vector>
#include "cv_shim.h"
#include "mappoint.h"
#include "../../include/svo_b200.h"

class frame {
public:
    frame();
    frame(frame *other);
    frame(cv::Mat &imLeft, cv::Mat &imRight, cv::Mat &imdepth, cv::Mat &img_detect, double &timestamp,
          cv::Mat &K, float &bf, std::vector<std::vector<int>> &detection_box);

    // Extraction settings used by every frame constructed afterwards (the YAML's ORBextractor.*
    // keys, which the reference never reads: Stereo/KITTI00-02.yaml:38-51).
    static void configure(int nfeatures, int nlevels = 8, float scaleFactor = 1.2f, int fastThreshold = 20,
                          int max_map_rows = 8192, int device = 0, int distribution = SVO_DIST_RETAIN_BEST);
    // The context shared by all frames of one image size on this thread (created on demand).
    static svo_ctx *engine(int width, int height);
    static void shutdown();

    void SetPose(cv::Mat mTcw);
    void featuredetect(cv::Mat &img);                       // src/frame.cc:75-79
    cv::Mat MB(cv::Mat &left, cv::Mat &right);              // src/frame.cc:82-91 (sparse here)
    void computekeypoint_r();                               // src/frame.cc:122-138
    void disp2Depth(float bf);                              // src/frame.cc:140-164
    cv::Mat UnprojectStereo(const float &u, const float &v, const float &z);   // src/frame.cc:166-180
    void createmappoint(std::set<mappoint *> &localmap);    // src/frame.cc:182-238

    int N;
    double timestamp;
    long int id;
    cv::Mat leftimg, rightimg;
    cv::Mat dispimg, depthimg;
    cv::Mat detectimg;
    std::vector<cv::KeyPoint> keypoints_l;
    std::vector<cv::Point2f> keypoints_r;
    cv::Mat f_descriptor;
    std::vector<mappoint *> MapPoints;
    std::vector<float> match_score;
    std::vector<bool> inlier;
    bool have_detected;
    std::vector<unsigned char> status;
    std::vector<float> error;
    std::vector<std::vector<int>> offline_box;
    float width, height;

    cv::Mat K;
    float fx, fy, cx, cy, bf;
    cv::Mat Tcw, Twc, tcw, twc, Rcw, Rwc;

    // sparse-stereo results per left keypoint (what keypoints_r / depthimg are filled from)
    std::vector<float> u_right, kp_depth;

private:
    void resize_per_feature_arrays();
};

// 8-bit gray view of an image (BGR is converted with OpenCV's fixed-point weights).
cv::Mat svo_to_gray(const cv::Mat &img);
