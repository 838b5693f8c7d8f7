// ref_harness.cc — TEST INFRASTRUCTURE.  C entry points that drive the reference's OWN classes (frame, mappoint,
// pnpmatch: /root/reference/src/{frame,mappoint,pnpmatch}.cc compiled unmodified against ref_stubs/minicv.hpp, see
// the `ref` target of oracle/Makefile) the way Tracking::Track / Tracklastframe drive them
// (src/Tracking.cc:184-250), so that tests can pin oracle/svo_oracle.c and the C++ drop-in adapter against what the
// reference's code itself computes.  Nothing here is used by the product.
#include <frame.h>
#include <mappoint.h>
#include <pnpmatch.h>

#include <map>
#include <set>
#include <vector>

namespace {
cv::Mat mat_from(const uint8_t *p, int w, int h, int ch)
{
    cv::Mat m(h, w, CV_MAKETYPE(CV_8U, ch));
    std::memcpy(m.data, p, (size_t)w * h * ch);
    return m;
}
cv::Mat K_from(const float *K9)
{
    cv::Mat K(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) K.at<float>(i / 3, i % 3) = K9[i];
    return K;
}
// a map point is named by (id of the frame that created it, keypoint index there): AddObservation(this, i) is the
// first observation it gets (src/frame.cc:229-231)
// When Tracking::Track itself runs (oracle/ref_g2o_harness.cc), createmappoint is called on Tracking::lastframe, ONE object
// that is re-assigned every frame: the creating observation of such a point is keyed by that object's address, whose id
// moves on (and AddObservation ignores every later observation from the same address, src/mappoint.cc:19-20).  The
// harness is told the address and reads the creating index from that entry.
frame *g_alias_frame = nullptr;
void name_of(mappoint *mp, int *create_id, int *idx)
{
    *create_id = mp->create_id; *idx = -1;
    for (auto &ob : mp->observations)
        if (ob.first->id == mp->create_id) { *idx = ob.second; return; }
    if (g_alias_frame) {
        auto it = mp->observations.find(g_alias_frame);
        if (it != mp->observations.end()) *idx = it->second;
    }
}
}  // namespace

extern "C" {

void ref_set_hooks(cv::minicv_orb_fn orb, cv::minicv_fund_fn fund, cv::minicv_pnp_fn pnp, cv::minicv_rodrigues_fn rod)
{
    cv::MiniCvHooks &h = cv::minicv_hooks();
    h.orb = orb; h.fund = fund; h.pnp = pnp; h.rodrigues = rod;
}

void ref_set_alias_frame(void *f) { g_alias_frame = (frame *)f; }

int ref_descriptor_distance(const uint8_t *a, const uint8_t *b)
{
    cv::Mat A(1, 32, CV_8U), B(1, 32, CV_8U);
    std::memcpy(A.data, a, 32); std::memcpy(B.data, b, 32);
    return pnpmatch::DescriptorDistance(A, B);        // src/pnpmatch.cc:14-30
}

// new frame(imLeft, imRight, imdepth, img_detect, timestamp, K, bf, detection_box)   (src/Tracking.cc:184, src/frame.cc:36-64)
void *ref_frame_new(const uint8_t *L, const uint8_t *R, int w, int h, int ch, const float *K9, float bf, const int *boxes,
                    int nboxes, double ts, long id)
{
    cv::Mat l = mat_from(L, w, h, ch), r = mat_from(R, w, h, ch), none, det = l.clone(), K = K_from(K9);
    std::vector<std::vector<int>> bx;
    for (int k = 0; k < nboxes; ++k) bx.push_back(std::vector<int>(boxes + 4 * k, boxes + 4 * k + 4));
    frame *f = new frame(l, r, none, det, ts, K, bf, bx);
    f->id = id;
    return f;
}
void *ref_frame_copy(void *f) { return new frame((frame *)f); }            // lastframe = frame(currentframe), Tracking.cc:237
void ref_frame_free(void *f) { delete (frame *)f; }
void ref_frame_featuredetect(void *f) { frame *F = (frame *)f; F->featuredetect(F->leftimg); }   // Tracking.cc:225
void ref_frame_set_disp(void *f, const float *disp)                                      // stands for dispimg = MB(..), :226
{
    frame *F = (frame *)f;
    cv::Mat d((int)F->height, (int)F->width, CV_32F);
    std::memcpy(d.data, disp, sizeof(float) * (size_t)d.rows * d.cols);
    F->dispimg = d;
}
void ref_frame_stereo(void *f) { frame *F = (frame *)f; F->computekeypoint_r(); F->disp2Depth(F->bf); }   // :227-228
void ref_frame_set_pose(void *f, const float *T16)
{
    cv::Mat T(4, 4, CV_32F);
    std::memcpy(T.data, T16, 64);
    ((frame *)f)->SetPose(T);
}
void ref_frame_unproject(void *f, float u, float v, float z, float *out3, int *ok)
{
    cv::Mat x = ((frame *)f)->UnprojectStereo(u, v, z);                                  // src/frame.cc:166-180
    *ok = !x.empty();
    if (*ok) for (int k = 0; k < 3; ++k) out3[k] = x.at<float>(k, 0);
}

void *ref_localmap_new() { return new std::set<mappoint *>(); }
int ref_localmap_size(void *lm) { return (int)((std::set<mappoint *> *)lm)->size(); }
// the set in ITS iteration order (pointer order: what pass 2 walks, src/pnpmatch.cc:160)
int ref_localmap_list(void *lm, int *create_id, int *idx, uint8_t *bad, float *pos3, uint8_t *desc32, int cap)
{
    int n = 0;
    for (mappoint *mp : *(std::set<mappoint *> *)lm) {
        if (n >= cap) break;
        name_of(mp, &create_id[n], &idx[n]);
        bad[n] = mp->bad ? 1 : 0;
        for (int k = 0; k < 3; ++k) pos3[3 * n + k] = mp->worldpos.at<float>(k, 0);
        std::memcpy(desc32 + 32 * (size_t)n, mp->m_descriptor.data, 32);
        ++n;
    }
    return n;
}
int ref_frame_createmappoint(void *f, void *lm)                                          // src/frame.cc:182-238, Tracking.cc:238
{
    std::set<mappoint *> &s = *(std::set<mappoint *> *)lm;
    const size_t before = s.size();
    ((frame *)f)->createmappoint(s);
    return (int)(s.size() - before);
}

// The local-map window of Tracking::Track (src/Tracking.cc:239-250), restated here because Tracking.cc itself needs
// Pangolin and is not compiled: when frame_num >= 4, every point with create_id <= frame_num - 4 leaves the set.
int ref_localmap_age(void *lm, int frame_num)
{
    std::set<mappoint *> &s = *(std::set<mappoint *> *)lm;
    int erased = 0;
    if (frame_num >= 4)
        for (auto it = s.begin(); it != s.end();) {
            if ((*it)->create_id <= frame_num - 4) { s.erase(it++); ++erased; }
            else ++it;
        }
    return erased;
}

// pnpmatch::poseEstimationPnP(currentframe, lastframe, localmappoints, mVelocity, K)   (src/Tracking.cc:114)
int ref_pose_estimation_pnp(void *cur, void *last, void *lm, const float *K9)
{
    cv::Mat K = K_from(K9), vel = cv::Mat::eye(4, 4, CV_32F);
    return pnpmatch::poseEstimationPnP((frame *)cur, *(frame *)last, *(std::set<mappoint *> *)lm, vel, K);
}

int ref_frame_counts(void *f, int *N, int *n_kp, int *n_kp_r)
{
    frame *F = (frame *)f;
    *N = F->N; *n_kp = (int)F->keypoints_l.size(); *n_kp_r = (int)F->keypoints_r.size();
    return F->f_descriptor.rows;
}
void ref_frame_get(void *f, float *kps6, uint8_t *desc, float *kp_r2, float *depth_at_kp, float *match_score, int *mp_create_id,
                   int *mp_idx, uint8_t *mp_bad, float *Tcw16)
{
    frame *F = (frame *)f;
    const int n = (int)F->keypoints_l.size();
    for (int i = 0; i < n; ++i) {
        const cv::KeyPoint &k = F->keypoints_l[i];
        if (kps6) { float *p = kps6 + 6 * (size_t)i; p[0] = k.pt.x; p[1] = k.pt.y; p[2] = k.size; p[3] = k.angle; p[4] = k.response; p[5] = (float)k.octave; }
        if (depth_at_kp && !F->depthimg.empty()) depth_at_kp[i] = F->depthimg.at<float>(k.pt.y, k.pt.x);   // as Tracking.cc:51 / frame.cc:194 read it
    }
    if (desc && F->f_descriptor.rows) std::memcpy(desc, F->f_descriptor.data, 32 * (size_t)F->f_descriptor.rows);
    if (kp_r2) for (size_t i = 0; i < F->keypoints_r.size(); ++i) { kp_r2[2 * i] = F->keypoints_r[i].x; kp_r2[2 * i + 1] = F->keypoints_r[i].y; }
    for (int i = 0; i < F->N; ++i) {
        if (match_score) match_score[i] = F->match_score[i];
        if (mp_create_id) {
            mappoint *mp = F->MapPoints[i];
            mp_create_id[i] = -1; mp_idx[i] = -1; mp_bad[i] = 0;
            if (mp) { name_of(mp, &mp_create_id[i], &mp_idx[i]); mp_bad[i] = mp->bad ? 1 : 0; }
        }
    }
    if (Tcw16 && !F->Tcw.empty()) std::memcpy(Tcw16, F->Tcw.data, 64);
}
void ref_frame_get_images(void *f, float *disp, float *depth)
{
    frame *F = (frame *)f;
    const size_t n = (size_t)F->height * (size_t)F->width;
    if (disp) std::memcpy(disp, F->dispimg.data, 4 * n);
    if (depth) std::memcpy(depth, F->depthimg.data, 4 * n);
}

}  // extern "C"
