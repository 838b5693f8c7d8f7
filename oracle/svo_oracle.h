/* svo_oracle.h — C interface of the CPU oracle (test infrastructure only;
 * see the header of svo_oracle.c for scope and reference citations). */
#ifndef SVO_ORACLE_H
#define SVO_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SVO_O_MAX_LEVELS 16

/* cv::KeyPoint minus class_id (always -1): src/frame.cc:78 fills vector<cv::KeyPoint> */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave;
} svo_o_keypoint;

typedef struct {
    int nlevels;
    int w[SVO_O_MAX_LEVELS], h[SVO_O_MAX_LEVELS];
    float scale[SVO_O_MAX_LEVELS];
    uint8_t *img[SVO_O_MAX_LEVELS];  /* un-blurred levels, stride == w */
    uint8_t *blur[SVO_O_MAX_LEVELS]; /* 7x7 sigma-2 blurred levels      */
} svo_o_pyramid;

void svo_o_geometry(int W, int H, int nlevels, float scale_factor, int nfeatures,
                    int *lw, int *lh, float *lscale, int *quota);
void svo_o_resize_coeffs(int src, int dst, int *ofs, int *w1);
void svo_o_resize(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride);
int svo_o_fast_score(const uint8_t *p, int stride);
int svo_o_fast_nms(const uint8_t *img, int w, int h, int stride, int threshold, int border,
                   int32_t *xs, int32_t *ys, int32_t *scores, int cap);
/* KeyPointsFilter::retainBest replay (retain_best.cpp): permutes resp/idx in place, returns kept count */
int svo_o_retain_best(float *resp, int32_t *idx, int n, int n_points);
void svo_o_introselect(float *resp, int32_t *idx, int n, int nth, int depth_limit);
void svo_o_harris(const uint8_t *img, int stride, const int32_t *xs, const int32_t *ys, int n, float *resp);
float svo_o_fast_atan2(float y, float x);
void svo_o_ic_angle(const uint8_t *img, int stride, const int32_t *xs, const int32_t *ys, int n, float *angle);
void svo_o_blur7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
void svo_o_brief(const uint8_t *blur, int stride, int cx, int cy, float angle_deg, uint8_t *desc);
int svo_o_orb(const uint8_t *gray, int W, int H, int stride, int nfeatures, float scale_factor,
              int nlevels, int fast_threshold, svo_o_keypoint *kps, uint8_t *desc, int cap,
              svo_o_pyramid *pyr_out);
int svo_o_orb_ex(const uint8_t *gray, int W, int H, int stride, int nfeatures, float scale_factor,
                 int nlevels, int fast_threshold, int distribution, svo_o_keypoint *kps, uint8_t *desc, int cap,
                 svo_o_pyramid *pyr_out);
void svo_o_pyramid_free(svo_o_pyramid *p);
/* opt-in quadtree keypoint distribution (svo_octree_oracle.c): points in raster order inside the rectangle
 * [x0,x1) x [y0,y1); writes the kept indices (node order) into out_idx[n] and returns their number */
#define SVO_O_OCT_MAXD 12
int svo_o_distribute_octree(const int32_t *xs, const int32_t *ys, const int32_t *score, int n,
                            int x0, int y0, int x1, int y1, int N, int32_t *out_idx);

int svo_o_hamming(const uint8_t *a, const uint8_t *b);
int svo_o_hamming_popcnt(const uint8_t *a, const uint8_t *b);
void svo_o_match_bf(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *idx, int32_t *dist, uint8_t *keep);
void svo_o_match_greedy(const uint8_t *rows, int M, const uint8_t *cur, int N, int mode,
                        const uint8_t *row_live, uint8_t *claimed, int32_t *claim_row, int row_base,
                        int32_t *best_idx, int32_t *best, int32_t *second, uint8_t *row_claimed,
                        const float *win_uvr, const float *cur_xy);
int svo_o_veto_dynamic(const int32_t *boxes, int n_boxes, const double *F, float lx, float ly, float cx, float cy);
void svo_o_match_greedy_veto(const uint8_t *rows, int M, const uint8_t *cur, int N, int mode,
                             const uint8_t *row_live, uint8_t *claimed, int32_t *claim_row, int row_base,
                             int32_t *best_idx, int32_t *best, int32_t *second, uint8_t *row_claimed,
                             const float *win_uvr, const float *cur_xy,
                             const int32_t *boxes, int n_boxes, const double *F, const float *row_xy,
                             const float *vcur_xy, uint8_t *row_bad);
void svo_o_project_map(const float *xyz, const int32_t *octave, int n, const float *Tcw, float fx, float fy, float cx, float cy,
                       int W, int H, float th, const float *lscale, int nlevels, float *uvr);
void svo_o_disp2depth(const float *disp, float *depth, size_t n, float bf);
int svo_o_stereo_sparse(const svo_o_keypoint *kl, const uint8_t *dl, int nl,
                        const svo_o_keypoint *kr, const uint8_t *dr, int nr,
                        const svo_o_pyramid *pl, const svo_o_pyramid *pr,
                        float bf, float b, float *u_right, float *depth, int32_t *match_r, int32_t *sad);
void svo_o_bgr2gray(const uint8_t *bgr, int w, int h, int sstride, uint8_t *gray, int dstride);
/* pose stage (svo_pose_oracle.c) */
int svo_o_pose_optimize(const float *Xw, const float *obs, int n, float fx, float fy, float cx, float cy,
                        const float *Tcw_in, float *Tcw_out, int iterations, double *stats);
int svo_o_pnp_ransac(const float *pts3d, const float *pts2d, int n, float fx, float fy, float cx, float cy,
                     int iterations, float reproj_err, uint32_t seed, int refine_iters,
                     double *R_out, double *t_out, uint8_t *inlier, int32_t *info);
#ifdef __cplusplus
}
#endif
#endif
