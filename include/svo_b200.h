/*
 * svo_b200.h — C ABI of libsvo_b200.so: the per-frame stereo front-end of
 * zssjh/stereo-semantic-vo as hand-written sm_100a CUDA.
 *
 * Plain pointers and sizes only; no cv:: / torch types.  The style follows the
 * reference's own precedent for a dlopen'ed C-ABI GPU plugin
 * (include/YOLOv3SE.h:61-64,208-232): opaque handle, caller-allocated result
 * arrays with a capacity, functions return a count or a negative status.
 *
 * Every entry point cites the reference interface it replaces (paths relative
 * to the reference root).  There is NO CPU fallback: every call fails with
 * SVO_E_CUDA when no sm_100 device / kernel image is available.
 *
 * Threading: one svo_ctx per host thread (not thread-safe), no process globals
 * (contrast src/Tracking.cc:19-20, src/pnpmatch.cc:13).
 */
#ifndef SVO_B200_H
#define SVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_OK 0
#define SVO_E_INVALID (-1)   /* bad argument                                   */
#define SVO_E_CUDA (-2)      /* CUDA error (see svo_last_error)                */
#define SVO_E_CAPACITY (-3)  /* an internal or caller capacity was exceeded    */
#define SVO_E_NOMEM (-4)

#define SVO_MAX_LEVELS 8
#define SVO_MAX_BOXES 256    /* offline YOLO boxes per frame (veto of pass 1) */
#define SVO_CAM_LEFT 0
#define SVO_CAM_RIGHT 1

/* greedy matcher modes (src/pnpmatch.cc) */
#define SVO_GREEDY_PASS1 0 /* :61-156  claim iff best < 15 (and not vetoed)                       */
#define SVO_GREEDY_PASS2 1 /* :160-199 claim iff best < 30 && (float)second/(float)best > 2       */

typedef struct svo_ctx svo_ctx;

/* cv::KeyPoint as filled by cv::ORB (src/frame.cc:78) minus class_id (always -1). */
typedef struct svo_keypoint {
    float x, y;     /* pt, level-0 pixel units                     */
    float size;     /* 31 * scale[octave]                          */
    float angle;    /* degrees, [0,360)                            */
    float response; /* Harris response                             */
    int32_t octave;
} svo_keypoint;

typedef struct svo_config {
    int device;          /* CUDA device ordinal                                                   */
    int width, height;   /* image size this context is built for                                  */
    int nfeatures;       /* ORBextractor.nFeatures (Stereo/KITTI00-02.yaml:38; frame.cc:77 uses 500) */
    int nlevels;         /* <= SVO_MAX_LEVELS                                                     */
    float scale_factor;  /* 1.2f                                                                  */
    int fast_threshold;  /* 20                                                                    */
    int max_batch;       /* stereo frames per svo_batch_submit (>= 1)                             */
    int lanes;           /* independent pipeline lanes (stream + buffers each), >= 1              */
    int max_rows;        /* capacity of one greedy row set / BF train set (e.g. 5000-row local map)*/
    void *stream;        /* optional cudaStream_t for lane 0 (NULL: the context creates its own)  */
    int max_channels;    /* 1 (default) or 3: 3 sizes the input staging for interleaved BGR images  */
    int distribution;    /* SVO_DIST_RETAIN_BEST (default, cv::ORB parity) or SVO_DIST_OCTREE (opt-in)  */
    int skip_match_score;/* 1: the batch path does not compute CurrentFrame->match_score (src/pnpmatch.cc:99:
                            p1_best_idx / p1_best / p1_second come back NULL).  Nothing in the reference reads
                            match_score (its only consumer, src/Optimizer.cc:56, is commented out); claims and
                            every other output are unchanged.  Default 0: computed, like the reference.     */
} svo_config;

/* Keypoint selection per pyramid level.
 * SVO_DIST_RETAIN_BEST: what the reference runs (src/frame.cc:77-78 -> cv::ORB): KeyPointsFilter::retainBest
 *   twice (2*quota by FAST score, quota by Harris response), output order included.  Bit-exact parity path.
 * SVO_DIST_OCTREE: north_star's "grid/octree distribution".  The reference has no such stage, so this is an
 *   opt-in, non-parity mode: the level's FAST corners go through ORB-SLAM2's DistributeOctTree scheme (quadtree
 *   split until the level quota is reached, best FAST score per node) with its pointer-order tie-breaks fixed
 *   (oracle/svo_octree_oracle.c is the definition); a level may return quota + 2 keypoints; `response` is still
 *   the Harris response.  Needs every level quota <= 4093 and a keypoint rectangle no wider than 16.5 x its height. */
#define SVO_DIST_RETAIN_BEST 0
#define SVO_DIST_OCTREE 1

void svo_default_config(svo_config *cfg);
const char *svo_version(void);

/* Replaces `new frame(...)`'s per-frame allocations (src/frame.cc:36-64) with one
 * up-front allocation of every device buffer, stream and graph. */
int svo_create(const svo_config *cfg, svo_ctx **out);
void svo_destroy(svo_ctx *ctx);
const char *svo_last_error(const svo_ctx *ctx);

/* Level geometry and per-level feature quotas (cv::ORB internals). */
int svo_get_geometry(const svo_ctx *ctx, int *lw, int *lh, float *lscale, int *quota);

/* ---------------------------------------------------------------------------
 * Synchronous single-call drop-ins (host buffers in, host buffers out).
 * ------------------------------------------------------------------------- */

/* frame::featuredetect (src/frame.cc:75-79) == cv::ORB::detectAndCompute.
 * gray: w x h 8-bit, `stride` bytes per row.  Writes up to `cap` keypoints (cv2's
 * exact output order) and cap x 32 descriptor bytes; returns the number of
 * keypoints found (may exceed nfeatures on response ties) or a negative status.
 * The pyramid/keypoints of `cam` stay resident for svo_stereo_sparse. */
int svo_extract(svo_ctx *ctx, int cam, const uint8_t *gray, int stride, int w, int h,
                svo_keypoint *kp_out, uint8_t *desc_out, int cap);

/* Same for an interleaved 8-bit BGR image (cv::imread of a colour KITTI frame, main.cpp:160-161):
 * cv::ORB converts colour input with cvtColor(BGR2GRAY) before anything else; the conversion runs on the
 * device while the image is repacked, gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15 (OpenCV 4.x's
 * 15-bit fixed point, equal to cv2.cvtColor on all 2^24 colours; OpenCV 3.2 used 14-bit weights that
 * differ on 0.3 % of colours).  Needs svo_config.max_channels = 3.  stride in bytes (>= 3*w). */
int svo_extract_bgr(svo_ctx *ctx, int cam, const uint8_t *bgr, int stride, int w, int h,
                    svo_keypoint *kp_out, uint8_t *desc_out, int cap);

/* north_star's ComputeStereoMatches stage: fills what frame::computekeypoint_r and
 * frame::disp2Depth deliver at keypoint pixels (src/frame.cc:122-164): u_right[i]
 * (keypoints_r[i].x) and depth[i] (depthimg at the keypoint), -1 where unmatched.
 * Row-band Hamming + 11x11 SAD + parabola, on the LEFT and
 * RIGHT images last given to svo_extract.  match_r / sad may be NULL.
 * Returns the number of left keypoints. */
int svo_stereo_sparse(svo_ctx *ctx, float bf, float baseline, float *u_right, float *depth,
                      int32_t *match_r, int32_t *sad, int cap);

/* cv::BFMatcher(NORM_HAMMING).match + the distance filter of find_feature_matches
 * (src/pnpmatch.cc:266,278,281-299).  q: nq x 32, t: nt x 32.  Per query: first
 * minimum train index, integer distance, keep = dist <= max(2*min_dist, 30). */
int svo_match_bf(svo_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                 int32_t *idx, int32_t *dist, uint8_t *keep);

/* The "dynamic" veto of pass 1 (src/pnpmatch.cc:103-137): a would-be match whose current
 * keypoint lies inside an offline YOLO box (+-10 px) and whose f64 epipolar distance under
 * F exceeds 0.1 marks the map point bad instead of claiming. */
typedef struct svo_veto {
    const int32_t *boxes; /* n_boxes x 4: left, right, top, bottom (main.cpp:59-97)        */
    int n_boxes;
    const double *F;      /* 3x3 row-major fundamental matrix (src/pnpmatch.cc:336)        */
    const float *row_xy;  /* 2 x M: LastFrame.keypoints_l[i].pt                            */
    const float *cur_xy;  /* 2 x N: CurrentFrame->keypoints_l[j].pt                        */
} svo_veto;

/* The sequential greedy scans of pnpmatch::poseEstimationPnP
 * (src/pnpmatch.cc:75-95 pass 1, :173-190 pass 2) with their accept rules.
 *   rows: M x 32 map-point descriptors in scan order; cur: N x 32 (f_descriptor)
 *   row_live[M] (NULL = all): 0 skips a row (no live map point, :66 / :165-172)
 *   claimed[N]  in/out: CurrentFrame->MapPoints[j] != NULL
 *   claim_row[N] in/out (may be NULL): row_base + i for the row that claimed j
 *   best_idx/best/second[M] (may be NULL): exact (bestIdx2, bestDist, secondBestDist)
 *   row_claimed[M] (may be NULL): 1 where the row claimed its best column
 *   win_uvr (3 x M) + cur_xy (2 x N): optional projection window |du|,|dv| <= r;
 *   NULL reproduces the reference (brute force over all columns).
 *   veto (pass 1 only, may be NULL) + row_bad[M] out: rows whose map point turns bad. */
int svo_match_greedy(svo_ctx *ctx, const uint8_t *rows, int M, const uint8_t *cur, int N, int mode,
                     const uint8_t *row_live, uint8_t *claimed, int32_t *claim_row, int row_base,
                     int32_t *best_idx, int32_t *best, int32_t *second, uint8_t *row_claimed,
                     const float *win_uvr, const float *cur_xy,
                     const svo_veto *veto, uint8_t *row_bad);

/* OPT-IN projection windows (north_star's "projection-guided" matching; the reference computes Velocity,
 * src/Tracking.cc:99-106, and never applies it, src/pnpmatch.cc:53-57): map point i at world position xyz[i] is
 * transformed with the predicted pose Tcw (row-major 4x4, Velocity * LastFrame.Tcw), projected with (fx, fy, cx, cy)
 * and gets the window (u, v, r = th * scale[octave[i]]) that svo_match_greedy / svo_frame_in.map_win_uvr consume;
 * points behind the camera or outside the image get r = -1 (matches nothing).  ORB-SLAM2's SearchByProjection form
 * (th = 7 for stereo); float32, every operation rounded separately in a fixed order.  octave may be NULL (level 0).
 * Host or device pointers.  Returns n or a negative status. */
int svo_project_map(svo_ctx *ctx, const float *xyz, const int32_t *octave, int n, const float *Tcw,
                    float fx, float fy, float cx, float cy, float th, float *uvr_out);

/* frame::disp2Depth (src/frame.cc:140-164): depth = bf/disp where disp != 0 else -1. */
int svo_disp2depth(svo_ctx *ctx, const float *disp, float *depth, size_t n, float bf);

/* ---------------------------------------------------------------------------
 * Pose stage (the step right after the matchers in
 * Tracking::Tracklastframe, src/Tracking.cc:108-121).  Batched: one CUDA block per problem.
 * ------------------------------------------------------------------------- */
typedef struct svo_pose_problem {
    const float *pts3d; /* n x 3: mapp->worldpos of CurrentFrame->MapPoints[j] (src/pnpmatch.cc:221, src/Optimizer.cc:62-65) */
    const float *pts2d; /* n x 2: CurrentFrame->keypoints_l[j].pt          (src/pnpmatch.cc:220, src/Optimizer.cc:47-48) */
    int n;              /* host or device pointers                                                     */
    float fx, fy, cx, cy;
    float Tcw[16];      /* svo_pose_optimize only: pFrame->Tcw, row-major 4x4 (cv::Mat CV_32F)          */
} svo_pose_problem;

typedef struct svo_pnp_result {
    double R[9];            /* cv::Rodrigues(rvec) (src/pnpmatch.cc:237-238), row-major               */
    double t[3];            /* tvec                                                                   */
    int32_t n_inliers;      /* inliers.rows (:229); 0 = no model                                      */
    int32_t best_iteration; /* sample that won, its P3P solution index, hypotheses scored             */
    int32_t best_solution;
    int32_t n_hypotheses;
} svo_pnp_result;

/* The role of cv::solvePnPRansac(pts3d, pts2d, K, Mat(), rvec, tvec, false, 100, 8.0, 0.99, inliers)
 * (src/pnpmatch.cc:227) as a data-parallel RANSAC: `iterations` 3-point samples (counter-based hash of
 * `seed`), P3P on each, every solution scored against all points with squared reprojection error
 * <= reproj_err^2, first maximum wins, Gauss-Newton refit (<= refine_iters steps) on its inliers.
 * NOT OpenCV's sampler/EPnP: results agree with cv2 statistically, not bit for bit (opt-in stage).
 * inliers: sum(n) flags, problems back to back (may be NULL).  Returns nproblems or a negative status. */
int svo_pnp_ransac(svo_ctx *ctx, const svo_pose_problem *problems, int nproblems, int iterations,
                   float reproj_err, uint32_t seed, int refine_iters, svo_pnp_result *results, uint8_t *inliers);

/* Optimizer::PoseOptimization (src/Optimizer.cc:15-86): g2o Levenberg-Marquardt (optimize(iterations),
 * reference: 10) on the frame pose over EdgeSE3ProjectXYZOnlyPose edges with Huber(sqrt(5.991)).
 * Tcw_out: 16 floats per problem (the pose SetPose receives); stats (may be NULL): per problem
 * {outer iterations run, final robust chi2}.  Returns nproblems or a negative status. */
int svo_pose_optimize(svo_ctx *ctx, const svo_pose_problem *problems, int nproblems, int iterations,
                      float *Tcw_out, double *stats);

/* ---------------------------------------------------------------------------
 * Batched, pipelined front-end: what Tracking::Track runs per stereo pair
 * (src/Tracking.cc:225-231): extract L+R, sparse stereo, BF match against the
 * previous frame, greedy pass 1 (previous frame's map points) and pass 2
 * (local map).  Inputs may be host (ideally pinned) or device pointers.
 * Results land in a context-owned pinned arena: one D2H copy per batch.
 * ------------------------------------------------------------------------- */
typedef struct svo_frame_in {
    const uint8_t *left, *right; /* width x height gray                                   */
    int stride;                  /* bytes per row                                         */
    float bf, baseline;          /* Camera.bf and bf/fx (Stereo/KITTI04-12.yaml:25)       */
    const uint8_t *prev_desc;    /* n_prev x 32: last frame's f_descriptor (BF train set  */
    int n_prev;                  /*   and pass-1 rows); NULL/0 skips both                 */
    const uint8_t *prev_live;    /* n_prev: 1 where LastFrame.MapPoints[i] is live; NULL=all */
    const uint8_t *map_desc;     /* n_map x 32 local-map descriptors in scan order        */
    int n_map;                   /*   NULL/0 skips pass 2                                 */
    const int32_t *map_prev_row; /* n_map (may be NULL): index of the pass-1 row holding the
                                    same map point, -1 if none; such rows are skipped in pass 2
                                    when pass 1 claimed them (observations.count, :167)   */
    int channels;                /* 0 or 1: gray; 3: interleaved BGR (both images), converted on
                                    the device like svo_extract_bgr                        */
    const float *map_win_uvr;    /* OPT-IN, changes results: 3 x n_map projection windows (u, v, r)
                                    for pass 2 — row i only sees current keypoints with |x-u| <= r and
                                    |y-v| <= r (the window of svo_match_greedy); the keypoints are
                                    binned into cells on the device and each row gathers its candidates
                                    instead of scanning all columns.  NULL (every frame of the batch or
                                    none) reproduces the reference's brute-force scan.          */
    /* Pass-1 "dynamic" veto (src/pnpmatch.cc:103-144), the batch form of svo_veto: a would-be match whose
       current keypoint lies inside an offline YOLO box (+-10 px) and whose f64 epipolar distance under F
       exceeds 0.1 claims nothing and marks the map point bad (p1_row_bad); pass 2 then skips the local-map
       rows linked to it (map_prev_row), as it skips bad points in the reference (:163).  The current keypoint
       positions are the extractor's own output.  Runs when boxes, F and prev_xy are all given.          */
    const int32_t *boxes;        /* n_boxes x 4: left, right, top, bottom (CurrentFrame->offline_box,
                                    main.cpp:59-97); at most SVO_MAX_BOXES                               */
    int n_boxes;
    const double *F;             /* 3x3 row-major fundamental matrix (findFundamentalMat, :336)         */
    const float *prev_xy;        /* 2 x n_prev: LastFrame.keypoints_l[i].pt                              */
    /* OPT-IN, changes results: the projection windows of pass 2 computed on the device (svo_project_map's
       arithmetic) instead of given as map_win_uvr: world positions and octaves of the local-map rows, the
       predicted pose and the intrinsics.  Used when map_xyz and Tcw_pred are given and map_win_uvr is NULL. */
    const float *map_xyz;        /* 3 x n_map                                                            */
    const int32_t *map_octave;   /* n_map (may be NULL: level 0)                                         */
    const float *Tcw_pred;       /* 16, row-major: Velocity * LastFrame.Tcw (src/Tracking.cc:99-106)     */
    float fx, fy, cx, cy, proj_th;   /* K and the window scale (ORB-SLAM2: 7 for stereo)                 */
    /* OPT-IN device-resident tracker state (svo_track_create): track_seq = 1 + the sequence this frame belongs
       to (0 = none).  prev_desc / n_prev / prev_live / map_desc / n_map / map_prev_row above are then IGNORED:
       the matchers read the sequence's state where it lies in HBM (pass-1 rows = the frozen descriptors of the
       map points the last frame's keypoints own, pass-2 rows = the local map) and the state is advanced on
       the device after the frame (createmappoint + the 4-frame window of src/Tracking.cc:237-250).  Frames of
       one sequence must be submitted in order, at most one per batch; frame_id is Tracking::frame_num.
       prev_xy is ignored too (the state holds the last frame's keypoints): the veto runs when boxes and F are
       given.  boxes also keep createmappoint from making points inside them (src/frame.cc:196-207).        */
    int track_seq;
    int frame_id;
} svo_frame_in;

typedef struct svo_frame_out {
    int32_t status;              /* SVO_OK or a negative status for this frame            */
    int32_t n_left, n_right;     /* keypoints found                                       */
    int32_t n_stereo;            /* left keypoints with depth > 0                         */
    const svo_keypoint *kp_left, *kp_right;
    const uint8_t *desc_left, *desc_right;       /* n x 32                                */
    const float *u_right, *depth;                /* n_left                                */
    const int32_t *bf_idx, *bf_dist;             /* n_left (query = current frame)        */
    const uint8_t *bf_keep;                      /* n_left                                */
    const int32_t *p1_best_idx, *p1_best, *p1_second; /* n_prev; match_score = second/best */
    const uint8_t *p1_row_claimed;               /* n_prev                                */
    const uint8_t *p1_row_bad;                   /* n_prev: 1 where the veto marked the row's map point bad (mp->bad, :141) */
    const uint8_t *p2_row_claimed;               /* n_map                                 */
    const int32_t *claim_row;                    /* n_left: row that claimed column j     */
                                                 /* (0..n_prev-1 pass 1, n_prev+i pass 2), -1 = free */
    int32_t n_prev, n_map;                       /* rows of the two passes (tracked frames: read from the state) */
    /* tracked frames only (else NULL): CurrentFrame->MapPoints[j] after the frame and createmappoint, named for the
       host: create_id of the owned point (-1 = none) and its position in the camera frame of the frame that created
       it (worldpos = Twc[create_id] * xyz; poses stay with the caller).  fx, fy, cx, cy of svo_frame_in must be set. */
    const int32_t *mp_create;                    /* n_left                                */
    const float *mp_xyz;                         /* n_left x 3                            */
} svo_frame_out;

/* Enqueue H2D + all kernels + one D2H for `n` frames on `lane`; returns at once. */
int svo_batch_submit(svo_ctx *ctx, int lane, const svo_frame_in *frames, int n);
/* Block until the lane's batch is complete. */
int svo_batch_wait(svo_ctx *ctx, int lane);
/* View of frame `i` of the lane's last batch (valid until the lane's next submit). */
int svo_batch_result(svo_ctx *ctx, int lane, int i, svo_frame_out *out);

/* ---------------------------------------------------------------------------
 * Device-resident tracker state (opt-in).  In the reference the previous frame's map points and the local map
 * are outputs of earlier frames (src/Tracking.cc:237-250: lastframe = frame(currentframe);
 * lastframe.createmappoint(LocalMapPoints); points with create_id <= frame_num - 4 are erased).  With a tracker
 * the batch path keeps them in HBM: a frame submitted with svo_frame_in.track_seq reads its pass-1 rows (frozen
 * map-point descriptors of the last frame's keypoints), their liveness, the local map and the pass-1 links from the
 * sequence's state, and advances the state on the device afterwards: every current keypoint keeps the point that
 * claimed it, or gets a new point when its stereo depth is > 0 and it lies outside every offline box grown by
 * 5 px (frame::createmappoint, src/frame.cc:182-238); points created `window` or more frames ago, and points the
 * pass-1 veto marked bad, leave the local map.  The BF matcher's train set is the last frame's own descriptors
 * (find_feature_matches re-extracts both images, src/pnpmatch.cc:253-300), also kept in the state.  Only images
 * cross PCIe per frame.
 * The pass-2 scan order is: surviving points in their previous order, then the new points in keypoint order (the
 * reference walks a std::set<mappoint*> in pointer order, i.e. an arbitrary one).
 * ------------------------------------------------------------------------- */
/* n_sequences independent states; map_capacity rows each (<= max_rows); window = 4 in the reference. */
int svo_track_create(svo_ctx *ctx, int n_sequences, int map_capacity, int window);
/* Empty state (no previous frame, empty map); optionally seeded with n_ballast map rows that never age out
 * (ballast: 32 bytes each; benchmarks use it to hold the local map at a fixed size).  Synchronous. */
int svo_track_reset(svo_ctx *ctx, int seq, const uint8_t *ballast, int n_ballast);
/* Test tap: the sequence's current state copied to caller buffers (any pointer may be NULL).  prev_* / last_desc:
 * capacity kp rows (svo_track_kp_capacity), map_*: map_capacity rows.  n_prev / n_map are filled in.  Synchronous. */
typedef struct svo_track_view {
    int32_t n_prev, n_map;
    uint8_t *last_desc;      /* n_prev x 32: the last frame's own descriptors (train set of find_feature_matches)     */
    uint8_t *prev_desc;      /* n_prev x 32: frozen m_descriptor of the point each keypoint of the last frame owns    */
    uint8_t *prev_live;      /* n_prev: 1 = LastFrame.MapPoints[i] != NULL                                            */
    int32_t *prev_map_row;   /* n_prev: the owned point's row in the local map, -1 = none / aged out of the map       */
    int32_t *prev_create;    /* n_prev: mappoint::create_id of the owned point, -1 = none                             */
    float *prev_xyz;         /* n_prev x 3: the owned point in the camera frame of its creating frame                  */
    float *prev_xy;          /* n_prev x 2: keypoints_l[i].pt                                                          */
    uint8_t *map_desc;       /* n_map x 32                                                                             */
    int32_t *map_create;     /* n_map: create_id (INT32_MAX = ballast)                                                 */
    int32_t *map_link;       /* n_map: keypoint of the last frame that owns the same point, -1 = none                  */
    float *map_xyz;          /* n_map x 3                                                                              */
    int32_t previous;        /* IN: 0 = the current state; 1 = the state the sequence's last frame read (the other
                                ping-pong copy: intact until the sequence's next frame is submitted)                  */
} svo_track_view;
int svo_track_state(svo_ctx *ctx, int seq, svo_track_view *view);
/* Rows of the per-keypoint arrays (the context's keypoint capacity per image). */
int svo_track_kp_capacity(const svo_ctx *ctx);

/* ---------------------------------------------------------------------------
 * Input staging (host side).  main.cpp:160-162 reads every KITTI frame with
 * cv::imread(path, CV_LOAD_IMAGE_UNCHANGED); these two calls decode the same PNG file image (bytes in memory) straight
 * into a caller buffer — ideally pinned (svo_alloc_pinned), so that svo_batch_submit's H2D copy reads it in place —
 * in cv::imread's memory order: 8-bit gray as is, 8-bit RGB / RGBA as interleaved BGR / BGRA (channels = 3 goes to
 * svo_extract_bgr / svo_frame_in.channels = 3, which convert on the device), 16-bit gray (the depth / disparity
 * PNGs) as native-endian uint16.  Host code over zlib: no context and no GPU needed; callers run one decode per core
 * beside the GPU lanes.  Interlaced and palette images are rejected (SVO_E_INVALID).
 * ------------------------------------------------------------------------- */
int svo_png_info(const uint8_t *file, size_t n, int *w, int *h, int *channels, int *bit_depth);
/* dst: h rows, dst_stride bytes apart (>= w * channels * bit_depth / 8), dst_cap bytes in all. */
int svo_png_decode(const uint8_t *file, size_t n, uint8_t *dst, size_t dst_stride, size_t dst_cap);

/* Pinned host / device memory helpers for callers that want zero staging. */
void *svo_alloc_pinned(svo_ctx *ctx, size_t bytes);
void svo_free_pinned(svo_ctx *ctx, void *p);
void *svo_alloc_device(svo_ctx *ctx, size_t bytes);
void svo_free_device(svo_ctx *ctx, void *p);
int svo_copy_to_device(svo_ctx *ctx, void *dst, const void *src, size_t bytes);

/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
long long svo_launch_count(const svo_ctx *ctx);
/* Device times (ms) of the lane's last batch, measured with CUDA events on the lane's own
 * stream (profiling must be on): ms[0] whole batch (H2D..D2H), [1] H2D, [2] pyramid,
 * [3] FAST, [4] first cull, [5] Harris, [6] second cull, [7] blur, [8] orient+BRIEF,
 * [9] stereo, [10] matching, [11] D2H; then two single kernels of the matching stage:
 * [12] k_pairs (fused BF + pass-1 distances), [13] k_shortlist of pass 2.  Writes min(n, 14) values. */
int svo_batch_stage_ms(svo_ctx *ctx, int lane, float *ms, int n);
/* Which results svo_batch_submit copies back, for the batches submitted from now on (default 0: every array at its
 * capacity, as one contiguous copy each).
 *   SVO_OUT_COMPACT   per frame only nfeatures + 32 rows of every per-keypoint and pass-1 array cross PCIe (strided
 *                     copies); a frame that holds more (response ties can push a level past its quota) gets the rest
 *                     fetched by svo_batch_wait.  Same results, ~15 % fewer bytes.
 *   SVO_OUT_NO_RIGHT  the right image's keypoints and descriptors stay on the device (kp_right / desc_right come back
 *                     NULL; n_right, u_right and depth are still filled).  The reference's frame keeps no right-image
 *                     features either: keypoints_r is keypoints_l shifted by the disparity (src/frame.cc:122-138).
 *   SVO_OUT_POSE_INPUTS  only what the steps after the matchers read — Tracklastframe's solvePnPRansac and
 *                     Optimizer::PoseOptimization (src/pnpmatch.cc:215-227, src/Optimizer.cc:40-70): the left keypoints,
 *                     depth, claim_row and (tracked frames) mp_create / mp_xyz.  Descriptors, u_right, the BF matches,
 *                     match_score and the row flags stay on the device (their pointers come back NULL); with the
 *                     tracker state in HBM the next frame does not need them from the host.  Implies SVO_OUT_NO_RIGHT. */
#define SVO_OUT_COMPACT 1
#define SVO_OUT_NO_RIGHT 2
#define SVO_OUT_POSE_INPUTS 4
int svo_set_outputs(svo_ctx *ctx, int flags);
/* Turn per-stage event recording on (1) or off (0, default). */
int svo_set_profiling(svo_ctx *ctx, int on);
void *svo_lane_stream(svo_ctx *ctx, int lane);

/* ---------------------------------------------------------------------------
 * Stage taps for the parity tests (device -> host copies of intermediate state
 * of the image last extracted on `cam`).
 * ------------------------------------------------------------------------- */
#define SVO_TAP_LEVEL 0      /* u8 level image, w*h bytes (tight)              */
#define SVO_TAP_BLUR 1       /* u8 blurred level                               */
#define SVO_TAP_FAST 2       /* int32 triples (x, y, score), raster order      */
#define SVO_TAP_SELECT1 3    /* int32 triples after the first retainBest       */
#define SVO_TAP_SELECT2 4    /* int32 triples (x, y, response bits) after the second retainBest */
/* Returns the number of elements written (bytes for images) or a negative status. */
long long svo_debug_tap(svo_ctx *ctx, int cam, int what, int level, void *out, size_t cap_bytes);

/* Runs the on-device retainBest replay (std::nth_element + std::partition order) on a
 * bare response array: idx_out receives the kept original indices in order; returns the
 * kept count.  depth_limit < 0 uses 2*floor(log2(n)) like libstdc++. */
int svo_debug_retain_best(svo_ctx *ctx, const float *resp, int n, int n_points, int depth_limit,
                          int32_t *idx_out);

/* Test tap of the tensor-core Hamming tiles the batch matchers run on (csrc/tcham.cu: tcgen05.mma.kind::i8 over
 * +-1-expanded descriptors): the full na x nb matrix of DescriptorDistance (src/pnpmatch.cc:14-30) values, row-major
 * into dist.  na, nb within the context's single-call capacities.  Returns na or a negative status. */
int svo_debug_hamming_matrix(svo_ctx *ctx, const uint8_t *a, int na, const uint8_t *b, int nb, int32_t *dist);
/* Developer tap: clock64 timeline of one CTA of each tensor-core matcher kernel of the last batch (context created
 * with SVO_B200_TC_PROF=1 in the environment): 1024 stamps = [mode][role][64]; tools/tc_timeline.py prints them. */
int svo_debug_tc_profile(svo_ctx *ctx, long long *stamps, int n);

#ifdef __cplusplus
}
#endif
#endif /* SVO_B200_H */
