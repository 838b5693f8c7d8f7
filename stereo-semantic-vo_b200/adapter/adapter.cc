// adapter.cc — host-side C++ mirror of the reference's frame / pnpmatch / mappoint classes on top
// of the C ABI (include/svo_b200.h).  No computation of the hot path happens here: pixels,
// descriptors and distances are produced by libsvo_b200.so; this file moves results into the
// public fields the rest of the reference (Tracking.cc, Optimizer.cc) reads.
#include "pnpmatch.h"
#include "Optimizer.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <tuple>

namespace {

struct Settings {
    int nfeatures = 500, nlevels = 8, fast = 20, max_rows = 8192, device = 0, distribution = SVO_DIST_RETAIN_BEST;
    float scale = 1.2f;
};
thread_local Settings g_set;
thread_local std::map<std::tuple<int, int>, svo_ctx *> g_engines;

[[noreturn]] void die(svo_ctx *c, const char *what, int rc)
{
    std::string msg = std::string(what) + " failed (" + std::to_string(rc) + "): " + (c ? svo_last_error(c) : "no context");
    throw std::runtime_error(msg);
}

svo_ctx *any_engine()
{
    if (g_engines.empty()) throw std::runtime_error("no engine yet: extract a frame first");
    return g_engines.begin()->second;
}

cv::Mat eye4()
{
    cv::Mat m(4, 4, CV_32F, 0.0);
    for (int i = 0; i < 4; ++i) m.at<float>(i, i) = 1.f;
    return m;
}

}  // namespace

// ------------------------------------------------------------------------------- mappoint
mappoint::mappoint(cv::Mat &pos, frame *pFrame, int id) : worldpos(pos), bad(false), observation_num(0), create_id(-1), octave(0)
{
    pFrame->f_descriptor.row(id).copyTo(m_descriptor);
    if (id >= 0 && id < (int)pFrame->keypoints_l.size()) octave = pFrame->keypoints_l[(size_t)id].octave;
}

void mappoint::AddObservation(frame *fm, size_t idx)
{
    if (observations.count(fm)) return;
    observations[fm] = (int)idx;
    ++observation_num;
}

// ---------------------------------------------------------------------------------- frame
void frame::configure(int nfeatures, int nlevels, float scaleFactor, int fastThreshold, int max_map_rows, int device,
                      int distribution)
{
    shutdown();
    g_set.nfeatures = nfeatures; g_set.nlevels = nlevels; g_set.scale = scaleFactor;
    g_set.fast = fastThreshold; g_set.max_rows = max_map_rows; g_set.device = device;
    g_set.distribution = distribution;   // SVO_DIST_OCTREE: north_star's opt-in quadtree distribution (not the reference's cv::ORB)
}

svo_ctx *frame::engine(int width, int height)
{
    auto key = std::make_tuple(width, height);
    auto it = g_engines.find(key);
    if (it != g_engines.end()) return it->second;
    svo_config cfg;
    svo_default_config(&cfg);
    cfg.device = g_set.device; cfg.width = width; cfg.height = height;
    cfg.nfeatures = g_set.nfeatures; cfg.nlevels = g_set.nlevels; cfg.scale_factor = g_set.scale;
    cfg.fast_threshold = g_set.fast; cfg.max_batch = 1; cfg.lanes = 1; cfg.max_rows = g_set.max_rows;
    cfg.distribution = g_set.distribution;
    cfg.max_channels = 3;   // colour KITTI frames (image_2/image_3) are converted on the device
    svo_ctx *ctx = nullptr;
    const int rc = svo_create(&cfg, &ctx);
    if (rc != SVO_OK) {
        std::string msg = ctx ? svo_last_error(ctx) : "svo_create";
        svo_destroy(ctx);
        throw std::runtime_error("svo_create failed (" + std::to_string(rc) + "): " + msg);
    }
    g_engines[key] = ctx;
    return ctx;
}

void frame::shutdown()
{
    for (auto &kv : g_engines) svo_destroy(kv.second);
    g_engines.clear();
}

cv::Mat svo_to_gray(const cv::Mat &img)
{
    if (img.channels() == 1) return img;
    cv::Mat g(img.rows, img.cols, CV_8UC1);
    for (int y = 0; y < img.rows; ++y) {
        const uint8_t *s = img.ptr(y);
        uint8_t *d = g.ptr(y);
        for (int x = 0; x < img.cols; ++x)   // OpenCV 4.x BGR2GRAY, 15-bit fixed point (same weights as the device path)
            d[x] = (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + 16384) >> 15);
    }
    return g;
}

namespace {
// cv::ORB on a gray or BGR image: colour input is converted on the device (svo_extract_bgr)
int extract_any(svo_ctx *ctx, int cam, const cv::Mat &img, svo_keypoint *kp, uint8_t *desc, int cap)
{
    if (img.channels() == 3)
        return svo_extract_bgr(ctx, cam, img.data, (int)img.step, img.cols, img.rows, kp, desc, cap);
    return svo_extract(ctx, cam, img.data, (int)img.step, img.cols, img.rows, kp, desc, cap);
}
}  // namespace

frame::frame() : N(0), timestamp(0), id(0), have_detected(false), width(0), height(0), fx(0), fy(0), cx(0), cy(0), bf(0) {}

frame::frame(frame *o)
    : N(o->N), timestamp(o->timestamp), id(o->id), leftimg(o->leftimg), rightimg(o->rightimg),
      dispimg(o->dispimg.clone()), depthimg(o->depthimg.clone()), detectimg(o->detectimg),
      keypoints_l(o->keypoints_l), keypoints_r(o->keypoints_r), f_descriptor(o->f_descriptor.clone()),
      MapPoints(o->MapPoints), match_score(o->match_score), inlier(o->inlier), have_detected(o->have_detected),
      status(o->status), error(o->error), offline_box(o->offline_box), width(o->width), height(o->height),
      K(o->K.clone()), fx(o->fx), fy(o->fy), cx(o->cx), cy(o->cy), bf(o->bf), u_right(o->u_right), kp_depth(o->kp_depth)
{
    if (!o->Tcw.empty()) SetPose(o->Tcw);
}

frame::frame(cv::Mat &imLeft, cv::Mat &imRight, cv::Mat &imdepth, cv::Mat &img_detect, double &time_stamp,
             cv::Mat &K_, float &mbf, std::vector<std::vector<int>> &detection_box)
    : N(g_set.nfeatures), timestamp(time_stamp), id(0), have_detected(false)
{
    (void)imdepth;
    K = K_;
    fx = K.at<float>(0, 0); fy = K.at<float>(1, 1); cx = K.at<float>(0, 2); cy = K.at<float>(1, 2);
    bf = mbf;
    leftimg = imLeft.clone(); rightimg = imRight.clone(); detectimg = img_detect.clone();
    width = (float)imLeft.cols; height = (float)imLeft.rows;
    resize_per_feature_arrays();
    depthimg = cv::Mat(imLeft.rows, imLeft.cols, CV_32F, -1.0);
    dispimg = cv::Mat(imLeft.rows, imLeft.cols, CV_32F, -1.0);
    offline_box = detection_box;
    SetPose(eye4());
}

void frame::resize_per_feature_arrays()
{
    MapPoints.assign((size_t)N, nullptr);
    inlier.assign((size_t)N, false);
    match_score.assign((size_t)N, -1.f);
}

void frame::SetPose(cv::Mat mTcw)
{
    Tcw = mTcw.clone();
    Rcw = cv::Mat(3, 3, CV_32F); Rwc = cv::Mat(3, 3, CV_32F); tcw = cv::Mat(3, 1, CV_32F); twc = cv::Mat(3, 1, CV_32F);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { Rcw.at<float>(r, c) = Tcw.at<float>(r, c); Rwc.at<float>(c, r) = Tcw.at<float>(r, c); }
        tcw.at<float>(r, 0) = Tcw.at<float>(r, 3);
    }
    for (int r = 0; r < 3; ++r) {
        float s = 0.f;
        for (int c = 0; c < 3; ++c) s += Rwc.at<float>(r, c) * tcw.at<float>(c, 0);
        twc.at<float>(r, 0) = -s;
    }
    Twc = eye4();
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Twc.at<float>(r, c) = Rwc.at<float>(r, c);
        Twc.at<float>(r, 3) = twc.at<float>(r, 0);
    }
}

void frame::featuredetect(cv::Mat &img)
{
    const cv::Mat &gray = img;
    svo_ctx *ctx = engine(gray.cols, gray.rows);
    const int cap = g_set.nfeatures * 2 + 1024;
    std::vector<svo_keypoint> kp((size_t)cap);
    cv::Mat desc(cap, 32, CV_8U);
    const int n = extract_any(ctx, SVO_CAM_LEFT, gray, kp.data(), desc.data, cap);
    if (n < 0) die(ctx, "svo_extract", n);
    const int m = n < cap ? n : cap;
    keypoints_l.resize((size_t)m);
    f_descriptor = cv::Mat(m, 32, CV_8U);
    for (int i = 0; i < m; ++i) {
        cv::KeyPoint &k = keypoints_l[(size_t)i];
        k.pt = cv::Point2f(kp[i].x, kp[i].y); k.size = kp[i].size; k.angle = kp[i].angle;
        k.response = kp[i].response; k.octave = kp[i].octave; k.class_id = -1;
        std::memcpy(f_descriptor.ptr(i), desc.ptr(i), 32);
    }
    N = m;   // the reference's fixed N = 500 becomes "the keypoints actually found"
    resize_per_feature_arrays();
}

cv::Mat frame::MB(cv::Mat &left, cv::Mat &right)
{
    // Sparse replacement of the dense MSA solve: extract the right image, run the row-band
    // Hamming + SAD stage, and scatter disparity / depth at the left keypoints' pixels.
    (void)left;
    const cv::Mat &gray = right;
    svo_ctx *ctx = engine(gray.cols, gray.rows);
    int n = extract_any(ctx, SVO_CAM_RIGHT, gray, nullptr, nullptr, 0);
    if (n < 0) die(ctx, "svo_extract(right)", n);
    const int cap = (int)keypoints_l.size();
    u_right.assign((size_t)cap, -1.f); kp_depth.assign((size_t)cap, -1.f);
    const float baseline = bf / fx;
    n = svo_stereo_sparse(ctx, bf, baseline, u_right.data(), kp_depth.data(), nullptr, nullptr, cap);
    if (n < 0) die(ctx, "svo_stereo_sparse", n);
    cv::Mat disp((int)height, (int)width, CV_32F, -1.0);
    for (int i = 0; i < cap; ++i)
        if (kp_depth[(size_t)i] > 0.f) {
            const cv::KeyPoint &k = keypoints_l[(size_t)i];
            disp.at<float>((int)k.pt.y, (int)k.pt.x) = k.pt.x - u_right[(size_t)i];
        }
    return disp;
}

void frame::computekeypoint_r()
{
    float rx = -1.f;
    keypoints_r.clear();
    keypoints_r.reserve(keypoints_l.size());
    for (size_t i = 0; i < keypoints_l.size(); ++i) {
        const float lx = keypoints_l[i].pt.x, ly = keypoints_l[i].pt.y;
        const float d = dispimg.at<float>((int)ly, (int)lx);
        if (d != -1.f) rx = lx - d;          // rx keeps its last value otherwise, like the reference
        keypoints_r.push_back(cv::Point2f(rx, ly));
    }
}

void frame::disp2Depth(float bf_)
{
    svo_ctx *ctx = engine((int)width, (int)height);
    cv::Mat depth(dispimg.rows, dispimg.cols, CV_32F);
    const int rc = svo_disp2depth(ctx, dispimg.ptr<float>(), depth.ptr<float>(), (size_t)dispimg.rows * dispimg.cols, bf_);
    if (rc != SVO_OK) die(ctx, "svo_disp2depth", rc);
    depthimg = depth;
}

cv::Mat frame::UnprojectStereo(const float &u, const float &v, const float &z)
{
    if (!(z > 0)) return cv::Mat();
    const float xc[3] = {(u - cx) * z * (1 / fx), (v - cy) * z * (1 / fy), z};
    // x3D = Rwc * x3Dc + twc as OpenCV evaluates it (one gemm with the addend, small-matrix path): the float products
    // summed left to right, then + twc — the order matters for the last bit (tests/test_ref_pin.py checks it on cv2.gemm)
    cv::Mat x3D(3, 1, CV_32F);
    for (int r = 0; r < 3; ++r) {
        float s = Rwc.at<float>(r, 0) * xc[0];
        s = s + Rwc.at<float>(r, 1) * xc[1];
        s = s + Rwc.at<float>(r, 2) * xc[2];
        x3D.at<float>(r, 0) = s + twc.at<float>(r, 0);
    }
    return x3D;
}

void frame::createmappoint(std::set<mappoint *> &localmap)
{
    const int n = (int)keypoints_l.size() < N ? (int)keypoints_l.size() : N;
    for (int i = 0; i < n; ++i) {
        if (MapPoints[(size_t)i]) continue;
        const float u = keypoints_l[(size_t)i].pt.x, v = keypoints_l[(size_t)i].pt.y;
        bool dynamic = false;
        for (const auto &b : offline_box)
            if (u > b[0] - 5 && u < b[1] + 5 && v > b[2] - 5 && v < b[3] + 5) { dynamic = true; break; }
        if (dynamic) continue;
        const float z = depthimg.at<float>((int)v, (int)u);
        if (z > 0) {
            cv::Mat x3D = UnprojectStereo(u, v, z);
            mappoint *mp = new mappoint(x3D, this, i);
            mp->AddObservation(this, (size_t)i);
            mp->create_id = (int)id;
            MapPoints[(size_t)i] = mp;
            localmap.insert(mp);
        }
    }
}

// ------------------------------------------------------------------------------- pnpmatch
cv::Mat pnpmatch::Cur_Tcw;
std::function<cv::Mat(const std::vector<cv::Point2f> &, const std::vector<cv::Point2f> &)> pnpmatch::fundamental_solver;
std::function<bool(const std::vector<cv::Mat> &, const std::vector<cv::Point2f> &, const cv::Mat &, cv::Mat &, int &)> pnpmatch::pnp_solver =
    pnpmatch::device_pnp;
std::vector<unsigned char> pnpmatch::last_pnp_inliers;
bool pnpmatch::use_projection = false;
float pnpmatch::projection_th = 7.f;
cv::Mat pnpmatch::predicted_Tcw;

// cv::solvePnPRansac(pts3d, pts2d, K, Mat(), rvec, tvec, false, 100, 8.0, 0.99, inliers) + Rodrigues + the 4x4
// Tcl of src/pnpmatch.cc:227-245, on the device.
bool pnpmatch::device_pnp(const std::vector<cv::Mat> &pts3d, const std::vector<cv::Point2f> &pts2d, const cv::Mat &K,
                          cv::Mat &Tcl, int &inliers)
{
    const int n = (int)pts2d.size();
    std::vector<float> p3((size_t)n * 3), p2((size_t)n * 2);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) p3[(size_t)i * 3 + k] = pts3d[(size_t)i].at<float>(k, 0);
        p2[(size_t)i * 2] = pts2d[(size_t)i].x; p2[(size_t)i * 2 + 1] = pts2d[(size_t)i].y;
    }
    svo_pose_problem pr{};
    pr.pts3d = p3.data(); pr.pts2d = p2.data(); pr.n = n;
    pr.fx = K.at<float>(0, 0); pr.fy = K.at<float>(1, 1); pr.cx = K.at<float>(0, 2); pr.cy = K.at<float>(1, 2);
    svo_pnp_result res{};
    last_pnp_inliers.assign((size_t)n, 0);
    svo_ctx *ctx = any_engine();
    const int rc = svo_pnp_ransac(ctx, &pr, 1, 100, 8.0f, 1u, 10, &res, last_pnp_inliers.data());
    if (rc < 0) die(ctx, "svo_pnp_ransac", rc);
    inliers = res.n_inliers;
    if (res.n_inliers == 0) return false;
    Tcl = eye4();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Tcl.at<float>(r, c) = (float)res.R[r * 3 + c]; Tcl.at<float>(r, 3) = (float)res.t[r]; }
    return true;
}

int pnpmatch::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    const uint64_t *pa = a.ptr<uint64_t>(), *pb = b.ptr<uint64_t>();
    int d = 0;
    for (int i = 0; i < 4; ++i) d += __builtin_popcountll(pa[i] ^ pb[i]);
    return d;
}

void pnpmatch::find_feature_matches(const cv::Mat &img_1, const cv::Mat &img_2, std::vector<cv::KeyPoint> &keypoints_1,
                                    std::vector<cv::KeyPoint> &keypoints_2, std::vector<cv::DMatch> &matches)
{
    // The reference re-runs ORB on both images here (4 redundant passes, src/pnpmatch.cc:268-273);
    // same results, so this does too but on the device, then BF-matches and filters.
    const cv::Mat &g1 = img_1, &g2 = img_2;
    svo_ctx *ctx = frame::engine(g1.cols, g1.rows);
    const int cap = g_set.nfeatures * 2 + 1024;
    std::vector<svo_keypoint> k1((size_t)cap), k2((size_t)cap);
    cv::Mat d1(cap, 32, CV_8U), d2(cap, 32, CV_8U);
    int n1 = extract_any(ctx, SVO_CAM_LEFT, g1, k1.data(), d1.data, cap);
    if (n1 < 0) die(ctx, "svo_extract", n1);
    int n2 = extract_any(ctx, SVO_CAM_RIGHT, g2, k2.data(), d2.data, cap);
    if (n2 < 0) die(ctx, "svo_extract", n2);
    auto fill = [](std::vector<cv::KeyPoint> &out, const std::vector<svo_keypoint> &in, int n) {
        out.resize((size_t)n);
        for (int i = 0; i < n; ++i) {
            out[(size_t)i].pt = cv::Point2f(in[(size_t)i].x, in[(size_t)i].y); out[(size_t)i].size = in[(size_t)i].size;
            out[(size_t)i].angle = in[(size_t)i].angle; out[(size_t)i].response = in[(size_t)i].response;
            out[(size_t)i].octave = in[(size_t)i].octave; out[(size_t)i].class_id = -1;
        }
    };
    fill(keypoints_1, k1, n1); fill(keypoints_2, k2, n2);
    std::vector<int32_t> idx((size_t)n1 + 1), dist((size_t)n1 + 1);
    std::vector<uint8_t> keep((size_t)n1 + 1);
    const int rc = svo_match_bf(ctx, d1.data, n1, d2.data, n2, idx.data(), dist.data(), keep.data());
    if (rc < 0) die(ctx, "svo_match_bf", rc);
    for (int i = 0; i < n1; ++i)
        if (keep[(size_t)i]) matches.push_back(cv::DMatch(i, idx[(size_t)i], (float)dist[(size_t)i]));
}

int pnpmatch::poseEstimation2D_2D(frame *cur, frame &last, cv::Mat &K, cv::Mat &F)
{
    (void)K;
    std::vector<cv::DMatch> matches;
    find_feature_matches(cur->leftimg, last.leftimg, cur->keypoints_l, last.keypoints_l, matches);
    std::vector<cv::Point2f> p1, p2;
    for (const cv::DMatch &m : matches) {
        const cv::Point2f c = cur->keypoints_l[(size_t)m.queryIdx].pt;
        bool dynamic = false;
        for (const auto &b : cur->offline_box)
            if (c.x > b[0] - 10 && c.x < b[1] + 10 && c.y > b[2] - 10 && c.y < b[3] + 10) { dynamic = true; break; }
        if (!dynamic) { p1.push_back(c); p2.push_back(last.keypoints_l[(size_t)m.trainIdx].pt); }
    }
    if (fundamental_solver) F = fundamental_solver(p1, p2);
    return (int)p1.size();
}

int pnpmatch::match_last_frame(frame *cur, frame &last, const cv::Mat &F)
{
    svo_ctx *ctx = frame::engine((int)cur->width, (int)cur->height);
    const int M = (int)last.keypoints_l.size() < last.N ? (int)last.keypoints_l.size() : last.N;
    const int Nc = (int)cur->keypoints_l.size();
    if (M == 0 || Nc == 0) return 0;
    // rows: one frozen descriptor per live map point of the last frame, in keypoint order
    cv::Mat rows(M, 32, CV_8U, 0.0);
    std::vector<uint8_t> live((size_t)M, 0), claimed((size_t)Nc, 0), took((size_t)M, 0), bad((size_t)M, 0);
    std::vector<float> row_xy((size_t)M * 2), cur_xy((size_t)Nc * 2);
    for (int i = 0; i < M; ++i) {
        mappoint *mp = last.MapPoints[(size_t)i];
        if (mp && !mp->bad) { live[(size_t)i] = 1; std::memcpy(rows.ptr(i), mp->m_descriptor.data, 32); }
        row_xy[2 * (size_t)i] = last.keypoints_l[(size_t)i].pt.x; row_xy[2 * (size_t)i + 1] = last.keypoints_l[(size_t)i].pt.y;
    }
    for (int j = 0; j < Nc; ++j) {
        claimed[(size_t)j] = cur->MapPoints[(size_t)j] ? 1 : 0;
        cur_xy[2 * (size_t)j] = cur->keypoints_l[(size_t)j].pt.x; cur_xy[2 * (size_t)j + 1] = cur->keypoints_l[(size_t)j].pt.y;
    }
    std::vector<int32_t> claim_row((size_t)Nc, -1), bi((size_t)M), bd((size_t)M), sd((size_t)M), boxes;
    for (const auto &b : cur->offline_box) for (int k = 0; k < 4; ++k) boxes.push_back(b[(size_t)k]);
    double Fd[9] = {0};
    svo_veto veto = {nullptr, 0, nullptr, nullptr, nullptr};
    // The reference always has an F here (src/pnpmatch.cc:36) and reads it for every would-be match inside a box
    // (:110-112); without one the "dynamic" test cannot run, and skipping it silently would change which map points
    // turn bad.  F comes from the fundamental_solver hook (cv::findFundamentalMat stays with the integrator's OpenCV).
    if (!boxes.empty() && F.empty())
        throw std::runtime_error("pnpmatch: offline_box is not empty but no fundamental matrix was produced "
                                 "(install pnpmatch::fundamental_solver, e.g. cv::findFundamentalMat(p1, p2, CV_FM_8POINT))");
    if (!F.empty() && (F.rows != 3 || F.cols != 3 || F.depth() != CV_64F))
        throw std::runtime_error("pnpmatch: the fundamental matrix must be 3x3 CV_64F (what cv::findFundamentalMat returns)");
    const bool use_veto = !boxes.empty() && !F.empty();
    if (use_veto) {
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Fd[3 * r + c] = reinterpret_cast<const double *>(F.ptr(r))[c];
        veto.boxes = boxes.data(); veto.n_boxes = (int)boxes.size() / 4; veto.F = Fd;
        veto.row_xy = row_xy.data(); veto.cur_xy = cur_xy.data();
    }
    const int rc = svo_match_greedy(ctx, rows.data, M, cur->f_descriptor.data, Nc, SVO_GREEDY_PASS1, live.data(), claimed.data(),
                                    claim_row.data(), 0, bi.data(), bd.data(), sd.data(), took.data(), nullptr, nullptr,
                                    use_veto ? &veto : nullptr, bad.data());
    if (rc < 0) die(ctx, "svo_match_greedy", rc);
    int matches = 0;
    if ((int)cur->match_score.size() < M) cur->match_score.resize((size_t)M, -1.f);
    for (int i = 0; i < M; ++i) {
        if (!live[(size_t)i]) continue;
        mappoint *mp = last.MapPoints[(size_t)i];
        cur->match_score[(size_t)i] = (float)sd[(size_t)i] / (float)bd[(size_t)i];     // src/pnpmatch.cc:99
        if (bad[(size_t)i]) mp->bad = true;
        else if (took[(size_t)i]) {
            cur->MapPoints[(size_t)bi[(size_t)i]] = mp;
            mp->AddObservation(cur, (size_t)bi[(size_t)i]);
            ++matches;
        }
    }
    return matches;
}

int pnpmatch::match_local_map(frame *cur, std::set<mappoint *> &localmappoints)
{
    svo_ctx *ctx = frame::engine((int)cur->width, (int)cur->height);
    const int Nc = (int)cur->keypoints_l.size();
    std::vector<mappoint *> order;           // the set's own iteration order (src/pnpmatch.cc:160)
    for (mappoint *mp : localmappoints)
        if (mp && !mp->bad && !mp->observations.count(cur)) order.push_back(mp);
    const int M = (int)order.size();
    if (M == 0 || Nc == 0) return 0;
    cv::Mat rows(M, 32, CV_8U);
    for (int i = 0; i < M; ++i) std::memcpy(rows.ptr(i), order[(size_t)i]->m_descriptor.data, 32);
    std::vector<uint8_t> claimed((size_t)Nc, 0), took((size_t)M, 0);
    for (int j = 0; j < Nc; ++j) claimed[(size_t)j] = cur->MapPoints[(size_t)j] ? 1 : 0;
    std::vector<int32_t> claim_row((size_t)Nc, -1);
    std::vector<float> win, cxy;
    if (use_projection && !predicted_Tcw.empty()) {   // opt-in: windows from the predicted pose, computed on the device
        std::vector<float> xyz((size_t)M * 3), T(16);
        std::vector<int32_t> oct((size_t)M);
        for (int i = 0; i < M; ++i) {
            for (int k = 0; k < 3; ++k) xyz[(size_t)i * 3 + k] = order[(size_t)i]->worldpos.at<float>(k, 0);
            oct[(size_t)i] = order[(size_t)i]->octave;
        }
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[(size_t)r * 4 + c] = predicted_Tcw.at<float>(r, c);
        win.resize((size_t)M * 3); cxy.resize((size_t)Nc * 2);
        const int pr = svo_project_map(ctx, xyz.data(), oct.data(), M, T.data(), cur->fx, cur->fy, cur->cx, cur->cy, projection_th, win.data());
        if (pr < 0) die(ctx, "svo_project_map", pr);
        for (int j = 0; j < Nc; ++j) { cxy[2 * (size_t)j] = cur->keypoints_l[(size_t)j].pt.x; cxy[2 * (size_t)j + 1] = cur->keypoints_l[(size_t)j].pt.y; }
    }
    const int rc = svo_match_greedy(ctx, rows.data, M, cur->f_descriptor.data, Nc, SVO_GREEDY_PASS2, nullptr, claimed.data(),
                                    claim_row.data(), 0, nullptr, nullptr, nullptr, took.data(), win.empty() ? nullptr : win.data(),
                                    win.empty() ? nullptr : cxy.data(), nullptr, nullptr);
    if (rc < 0) die(ctx, "svo_match_greedy", rc);
    int n = 0;
    for (int j = 0; j < Nc; ++j) {
        const int r = claim_row[(size_t)j];
        if (r < 0) continue;
        cur->MapPoints[(size_t)j] = order[(size_t)r];
        order[(size_t)r]->AddObservation(cur, (size_t)j);
        ++n;
    }
    return n;
}

int pnpmatch::poseEstimationPnP(frame *cur, frame &last, std::set<mappoint *> &localmappoints, cv::Mat &mVelocity, cv::Mat &K)
{
    // the reference never applies mVelocity (src/pnpmatch.cc:53 is commented out); with use_projection it predicts the pose
    predicted_Tcw = cv::Mat();
    if (use_projection && !mVelocity.empty() && !last.Tcw.empty()) {
        cv::Mat T(4, 4, CV_32F, 0.0);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += mVelocity.at<float>(r, k) * last.Tcw.at<float>(k, c);
            T.at<float>(r, c) = s;
        }
        predicted_Tcw = T;
    }
    cv::Mat F;
    poseEstimation2D_2D(cur, last, K, F);           // overwrites both frames' keypoints, as the reference does
    match_last_frame(cur, last, F);
    match_local_map(cur, localmappoints);
    std::vector<cv::Mat> pts3d;
    std::vector<cv::Point2f> pts2d;
    const int n = (int)cur->keypoints_l.size() < cur->N ? (int)cur->keypoints_l.size() : cur->N;
    for (int j = 0; j < n; ++j)
        if (mappoint *mp = cur->MapPoints[(size_t)j]) { pts2d.push_back(cur->keypoints_l[(size_t)j].pt); pts3d.push_back(mp->worldpos); }
    int inliers = 0;
    if (pnp_solver && !pts2d.empty()) {
        cv::Mat Tcl;
        if (pnp_solver(pts3d, pts2d, K, Tcl, inliers)) {
            cv::Mat T(4, 4, CV_32F, 0.0);
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
                float s = 0.f;
                for (int k = 0; k < 4; ++k) s += Tcl.at<float>(r, k) * cur->Tcw.at<float>(k, c);
                T.at<float>(r, c) = s;
            }
            cur->SetPose(T);
            Cur_Tcw = T;
        }
    }
    return (int)pts2d.size();
}

// ------------------------------------------------------------------------------- Optimizer
int Optimizer::PoseOptimization(frame *pFrame)
{
    std::vector<float> p3, p2;
    const int n = (int)pFrame->keypoints_l.size() < pFrame->N ? (int)pFrame->keypoints_l.size() : pFrame->N;
    for (int i = 0; i < n; ++i)
        if (mappoint *mp = pFrame->MapPoints[(size_t)i]) {                       // src/Optimizer.cc:42-71
            p2.push_back(pFrame->keypoints_l[(size_t)i].pt.x); p2.push_back(pFrame->keypoints_l[(size_t)i].pt.y);
            for (int k = 0; k < 3; ++k) p3.push_back(mp->worldpos.at<float>(k, 0));
        }
    svo_pose_problem pr{};
    pr.pts3d = p3.data(); pr.pts2d = p2.data(); pr.n = (int)(p2.size() / 2);
    pr.fx = pFrame->fx; pr.fy = pFrame->fy; pr.cx = pFrame->cx; pr.cy = pFrame->cy;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pr.Tcw[r * 4 + c] = pFrame->Tcw.at<float>(r, c);
    float T[16];
    svo_ctx *ctx = any_engine();
    const int rc = svo_pose_optimize(ctx, &pr, 1, 10, T, nullptr);
    if (rc < 0) die(ctx, "svo_pose_optimize", rc);
    cv::Mat pose(4, 4, CV_32F, 0.0);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pose.at<float>(r, c) = T[r * 4 + c];
    pFrame->SetPose(pose);                                                       // :82-84
    return pr.n;
}
