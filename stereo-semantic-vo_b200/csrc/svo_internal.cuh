// svo_internal.cuh — shared layout of the device-resident front-end state.
//
// HBM layout (all sized once in svo_create, nothing allocated per frame):
//   one "image slot" per image in flight (2 per stereo frame), each holding
//     pyr   : the 8 un-blurred pyramid levels, rows padded to 16 B (level offsets 256-B aligned)
//     blur  : the 7x7 sigma-2 blurred levels, same layout
//     bands : per level, per 8-row band, the raster-ordered FAST corners (packed u32)
//     ckey/cval/lpos/rpos : per level candidate arrays the retainBest replay permutes
//     key2/val2           : Harris-rescored survivors of the first cull
//     kp/desc             : final keypoints (cv2 order) and 32-byte descriptors
// Everything a frame touches (~10 MB at 1241x376) stays L2-resident on B200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/svo_b200.h"

#define SVO_EDGE 31            // cv::ORB edgeThreshold
#define SVO_BLUR_ROWS 64       // rows per blur tile (describe.cu:k_blur; 6 halo rows are staged on top)
#define SVO_SHORT_CAP 128      // short-list entries per greedy row before the full-scan path
#define SVO_STRIDE_BGR (1 << 30)  // flag in a per-image stride word: the source is interleaved BGR
#define SVO_WIN_CELLS 4096     // most cells of the keypoint grid used by the windowed pass 2
#define SVO_TC_TILE_BYTES 32768  // one tensor-core operand image: 128 descriptors x 256 int8 (tcham.cu)
#define SVO_TC_STREAMS 2       // column ranges (streams) a tensor-core CTA splits a row's scan into (tcham.cu)
#ifndef SVO_TC_GATHER_ROWS
#define SVO_TC_GATHER_ROWS 1   // tensor-core pass 2 scans a gathered list of the live rows instead of masking the dead ones (0: the masked form)
#endif
#define SVO_TC_SEG (SVO_SHORT_CAP / SVO_TC_STREAMS)   // TC_SHORT: short-list slots of each range of a row
#define SVO_STATUS_OVERFLOW 1  // bit set in the per-image status word on a capacity overflow
#define SVO_STATUS_DEPTH 2     // introselect reached its depth limit (heap-select path ran)

struct LevelGeom {
    int w, h, pitch;   // pitch: bytes per row, multiple of 16
    unsigned pitch_magic;  // ceil(2^32 / pitch): __umulhi(v, pitch_magic) == v / pitch for v < 2^16
    int off;           // byte offset of the level inside a slot's pyramid buffer
    float scale;       // (float)pow(1.2, l)
    float inv_scale;   // 1.f / scale
    int quota;         // nfeaturesPerLevel[l]
    int x0, x1, y0, y1;  // keypoints live in [x0,x1) x [y0,y1) = [31, w-31) x [31, h-31)
    int nbands;        // ceil((y1-y0)/Geom.fast_band), 0 when the level is too small
    int band_cap;      // entries per band list
    int band_off;      // entry offset of this level's band lists inside a slot
    int bandcnt_off;   // offset of this level's band counters
    int cand_cap;      // candidate capacity (sum of band caps)
    int cand_off;      // entry offset of the candidate arrays
    int cap2;          // capacity after the first cull
    int off2;          // entry offset of key2/val2
    int tab_off;       // offset of this level's resize tables (x table then y table), level >= 1
    int rq_off;        // first entry of this level's quad table (resize_quads.h), level >= 1
    int rq_ok;         // 1: every quad's taps fit the 8-byte window, k_resize_q runs; 0: the per-byte k_resize
    int blur_tile_off; // first blur tile index of this level
    int blur_tiles_x;  // blur tiles per strip of 32 rows
    int blur_tq;       // pixel quads per blur tile (multiple of 4, <= 128)
    int oct_nini;      // octree mode: root nodes = round(width / height) of the keypoint rectangle
    float oct_hx;      // octree mode: root width (float), (x1 - x0) / oct_nini
};

struct Geom {
    int nlevels, W, H;
    int fast_threshold;
    int fast_band;       // output rows per FAST band (one CTA each): 16 for batch contexts, 8 for latency contexts and very wide images
    int pyr_bytes;       // per slot
    int band_total;      // entries per slot
    int bandcnt_total;   // counters per slot
    int cand_total;      // entries per slot
    int total2;          // entries per slot
    int kp_cap;          // final keypoints per image
    int blur_tiles;      // blur tiles per image
    int fast_bands;      // FAST bands per image (sum of nbands)
    int blur_margin;     // 1: k_blur<true> (reflected margins in the staged tile); 0 (SVO_B200_BLUR_MARGIN=0): per-lane reflect path
    int harris8;         // 1: k_harris4 (eight lanes per candidate); 0 (SVO_B200_HARRIS8=0): the warp-per-candidate k_harris
    int pyr_nbands;      // bands of the fused pyramid kernel (0: the geometry does not fit it)
    int pyr_fused_max;   // launches of at most this many images use the fused kernel, larger ones the per-level kernels
    int pyr_soff[SVO_MAX_LEVELS];   // shared-memory byte offset of each level's rows in that kernel
    int pyr_smem;        // its dynamic shared memory
    LevelGeom lv[SVO_MAX_LEVELS];
};

// One band of the fused pyramid kernel (pyramid.cu): per level the rows the CTA computes in shared memory
// ([clo, chi], a superset of what the next level needs from it) and the rows it owns, i.e. writes to HBM ([olo, ohi)).
struct PyrBand { short clo[SVO_MAX_LEVELS], chi[SVO_MAX_LEVELS], olo[SVO_MAX_LEVELS], ohi[SVO_MAX_LEVELS]; };

// Base pointers of the slot arrays (index = slot * per-slot size + offset).
struct Bufs {
    uint8_t *pyr, *blur;
    uint32_t *bands;
    int *bandcnt;
    float *ckey;
    uint32_t *cval, *lpos, *rpos;
    float *key2;
    uint32_t *val2;
    int *cnt1;    // [slot][8] candidates per level
    int *kept1;   // [slot][8] survivors of the first cull
    int *kept2;   // [slot][8] survivors of the second cull
    svo_keypoint *kp;
    uint8_t *desc;
    int *nkp;     // [slot]
    int *status;  // [slot]
    const uint32_t *rtab;  // resize tables: (ofs << 9) | w1
    const uint32_t *rqtab; // per output quad of every level >= 1: 8 words (resize_quads.h), 32-byte aligned
    const PyrBand *pyr_bands;   // [Geom.pyr_nbands] or NULL: the fused pyramid kernel is not used for this geometry
};

// packed candidate: x | y << 12 | score << 24
__host__ __device__ inline uint32_t pack_xy(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
__host__ __device__ inline int unpack_x(uint32_t v) { return (int)(v & 0xfffu); }
__host__ __device__ inline int unpack_y(uint32_t v) { return (int)((v >> 12) & 0xfffu); }
__host__ __device__ inline int unpack_s(uint32_t v) { return (int)(v >> 24); }

// Device-resident tracker state of one sequence (track.cu), one of its two ping-pong copies.
struct TrackState {
    // what Tracking keeps of the last frame (lastframe, src/Tracking.cc:237-238): one entry per keypoint
    uint8_t *last_desc;      // [kp_cap][32] the last frame's own f_descriptor (train set of find_feature_matches)
    uint8_t *prev_desc;      // [kp_cap][32] frozen m_descriptor of the map point the keypoint owns (pass-1 rows)
    uint8_t *prev_live;      // [kp_cap] 1 = owns a live map point
    int *prev_map_row;       // [kp_cap] that point's row in the local map, -1 = not in the map
    int *prev_create;        // [kp_cap] id of the frame that created the owned point, -1 = none
    float *prev_xyz;         // [kp_cap][3] the owned point in the camera frame of its creating frame (UnprojectStereo before Rwc / twc)
    float *prev_xy;          // [kp_cap][2] keypoints_l[i].pt (the veto's `last`)
    int *n_prev;             // [1]
    // LocalMapPoints in scan order
    uint8_t *map_desc;       // [map_cap][32] frozen descriptors
    int *map_create;         // [map_cap] id of the frame that created the point (INT_MAX: ballast, never ages out)
    int *map_link;           // [map_cap] keypoint of the last frame owning the same point, -1 = none (map_prev_row)
    float *map_xyz;          // [map_cap][3]
    int *n_map;              // [1]
};

// Per-frame input pointers of a batch (device table).  Inputs that already live in device memory are read in
// place; host inputs are gathered into the lane's landing zone with as few H2D copies as their layout allows.
struct FramePtrs {
    const uint8_t *left, *right;   // gray images (rows `stride` bytes apart)
    const uint8_t *prev;           // n_prev x 32, 16-byte aligned: pass-1 rows
    const uint8_t *last;           // n_prev x 32: train set of the BF matcher (== prev unless the frame is tracked)
    const uint8_t *prev_live;      // n_prev or NULL (all live)
    const uint8_t *map;            // n_map x 32, 16-byte aligned
    const int *map_prev_row;       // n_map or NULL
    const float *map_win;          // 3 x n_map (u, v, r) projection windows of pass 2, or NULL
    // pass-1 "dynamic" veto (src/pnpmatch.cc:103-144): offline YOLO boxes, fundamental matrix, last frame's keypoint positions
    const int *boxes;              // n_boxes x 4 (left, right, top, bottom) or NULL
    const double *F;               // 3x3 row-major, 8-byte aligned, or NULL
    const float *prev_xy;          // 2 x n_prev or NULL
    int n_boxes;
    // opt-in projection windows computed on the device (k_win_prepare): map_xyz != NULL
    const float *map_xyz;          // 3 x n_map
    const int *map_octave;         // n_map or NULL
    float Tcw[12];                 // first three rows of the predicted pose
    float fx, fy, cx, cy, proj_th;
    // device-resident tracker state (track.cu): the copy this frame reads and the copy its update writes; NULL = untracked
    const TrackState *trk_in, *trk_out;
    int frame_id;
};

// ---- stage launchers (each enqueues on `st` for images [slot0, slot0 + nimg)) ----
void launch_unpack(const Bufs &b, const Geom &g, int slot0, int nimg, const FramePtrs *fp,
                   const int *strides, cudaStream_t st, long long *launches);
void launch_pyramid(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
void launch_fast(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
void launch_select1(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
void launch_harris(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
void launch_select2(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
void launch_blur(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
void launch_describe(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
// opt-in quadtree distribution (octree.cu): replaces launch_select1 / launch_select2
void launch_octree(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
void launch_keep_all(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches);
int setup_octree_attributes();
int octree_max_quota();
int fast_smem_bytes(const Geom &g);
int setup_fast_attributes(const Geom &g);
int setup_describe();
int setup_select_attributes();
int setup_pyramid_attributes(const Geom &g);

// bare retainBest replay (debug/test entry)
void launch_retain_best_raw(float *key, uint32_t *val, int n, int n_points, int depth_limit,
                            uint32_t *lpos, uint32_t *rpos, int *kept_out, int *status, cudaStream_t st,
                            long long *launches);

// ---- stereo ----
struct StereoArgs {
    float *u_right, *depth;   // [frame][kp_cap]
    int *match_r, *sad;       // [frame][kp_cap]
    int *n_stereo;            // [frame]
    int stride;               // entries per frame in the arrays above
    const float *bf, *baseline;  // [frame]
    int *row_off;             // [frame][H + 1] CSR offsets of the per-row candidate lists (right keypoints)
    uint16_t *row_list;       // [frame][row_list_stride]
    int row_list_stride;      // kp_cap * (rows one right keypoint can cover)
};
void launch_stereo(const Bufs &b, const Geom &g, int slot0, int nframes, const StereoArgs &a, cudaStream_t st,
                   long long *launches);

// ---- matching ----
struct MatchSet {            // one descriptor set per frame, fixed stride
    const uint8_t *desc;     // [frame][stride_rows][32]
    const int *count;        // [frame * count_stride] (device) or NULL -> use fixed_count
    int count_stride;
    int stride_rows;         // rows per frame in the per-row/per-column side arrays
    int desc_stride;         // descriptor rows between consecutive frames' blocks
    int fixed_count;
    const uint8_t *const *tab;   // optional per-frame descriptor pointers (a field of FramePtrs[frame]); overrides desc
};
struct GreedyArgs {
    MatchSet rows, cols;
    int mode, row_base;
    const int *row_base_arr;      // [frame] or NULL: per-frame row base added to row_base
    const uint8_t *row_live;      // [frame][rows.stride_rows] or NULL
    const int *map_prev_row;      // [frame][rows.stride_rows] or NULL (pass 2)
    const FramePtrs *fp;          // batch path: row_live / map_prev_row come from fp[frame] (use_live / use_map_prev)
    int use_live, use_map_prev;
    // batch pass 2: work lists built by k_greedy_init (NULL: every row goes through k_shortlist), the u8
    // previous x current distance matrix k_pairs wrote (NULL: none) and the previous-frame set, so that rows
    // repeating a pass-1 row (map_prev_row) reuse its distances
    int *need_list, *reuse_list;  // [frame][rows.stride_rows]
    int *list_cnt;                // [frame][2]
    uint8_t *row_need;            // [frame][rows.stride_rows] or NULL: tensor-core pass 2 (tcham.cu) — k_greedy_init writes 1 for every row
                                  // that is scanned instead of building the work lists
    // tensor-core path: operand images (tcham.cu) of the row set, of the column set and of the gathered free columns
    uint8_t *img_rows, *img_cols, *img_free;
    size_t img_rows_stride, img_cols_stride, img_free_stride;
    int img_cols_ready;           // the column image was already written by an earlier launcher of this batch
    long long *tc_prof;           // optional in-kernel timeline buffer (TcArgs.prof)
    const uint8_t *dmat; size_t dmat_frame_stride; int dmat_pitch;
    MatchSet prev;
    const uint8_t *prev_row_claimed;  // [frame][prev stride] (pass 2, with map_prev_row)
    const uint8_t *prev_row_bad;      // [frame][prev stride] or NULL: pass-1 rows whose map point turned bad
    const int *prev_count;            // [frame] rows of the previous set (pass 2, with map_prev_row): larger links are ignored
    int prev_stride;
    uint8_t *claimed;             // [frame][cols.stride_rows] in/out
    int *claim_row;               // [frame][cols.stride_rows] in/out
    int *claim_time;              // [frame][cols.stride_rows] scratch (global row that claimed, INT_MAX = free, -1 = pre-claimed)
    int *best_idx, *best, *second;   // [frame][rows.stride_rows] or NULL
    uint8_t *row_claimed;         // [frame][rows.stride_rows]
    uint8_t *row_bad;             // [frame][rows.stride_rows] or NULL
    uint32_t *shortlist;          // [frame][rows.stride_rows][32]: entries 0..31
    uint32_t *shortlist_hi;       // [frame][rows.stride_rows][SVO_SHORT_CAP - 32]: the rest
    int *short_cnt;               // [frame][rows.stride_rows]
    int *res_rows, *res_off, *res_want, *res_perm;   // [frame][rows.stride_rows] resolver scratch (candidate rows, CSR offsets,
                                                      // decisions, length-sorted processing order)
    const float *win_uvr;         // [frame][rows.stride_rows][3] or NULL
    const float *cur_xy;          // [frame][cols.stride_rows][2] or NULL
    // batch pass 2 with projection windows (opt-in): k_win_prepare copies the windows and the current keypoints'
    // positions into win_out / cur_xy_out (= win_uvr / cur_xy) and bins the keypoints into square cells, then
    // k_shortlist_win gathers each row's candidates from the cells under its window
    int win_gather, cell_shift, ncx, ncy;
    int img_w, img_h, nlevels; float lscale[SVO_MAX_LEVELS];   // projection on the device (fp[f].map_xyz)
    int *cell_off;                // [frame][SVO_WIN_CELLS + 1]
    uint16_t *cell_list;          // [frame][cols.stride_rows]
    // batch pass 2 without windows: the ascending list of columns pass 1 left free (k_free_cols); k_shortlist scans only those
    uint16_t *free_col;           // [frame][cols.stride_rows] or NULL
    int *free_cnt;                // [frame]
    const svo_keypoint *kp; size_t kp_frame_stride;   // current (left) keypoints of frame f at kp + f * kp_frame_stride
    float *win_out, *cur_xy_out;
    // veto (pass 1).  Single call: the arrays below; batch (fp != NULL, use_veto): fp[frame].boxes / F / prev_xy and kp
    const int *boxes; int n_boxes; const double *F; const float *row_xy;
    int use_veto;
};
void launch_greedy(const GreedyArgs &a, int nframes, bool want_scores, cudaStream_t st, long long *launches,
                   cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);   // optional events around k_shortlist
void launch_prune_lists(const GreedyArgs &a, int nframes, cudaStream_t st, long long *launches);

struct BfArgs {
    MatchSet q, t;
    int *idx, *dist;      // [frame][q.stride_rows]
    uint8_t *keep;        // [frame][q.stride_rows]
    int *min_dist;        // [frame] scratch
};
void launch_bf(const BfArgs &a, int nframes, cudaStream_t st, long long *launches);
struct PairArgs {            // fused BF + pass-1 front of the batch path (match.cu: k_pairs / k_scores_m)
    GreedyArgs g;            // rows = previous frame, cols = current frame, pass-1 outputs
    uint8_t *dmat;           // [frame][rows][dmat_pitch] u8 distances, min(d, 255)
    size_t dmat_frame_stride;
    int dmat_pitch;          // bytes per matrix row, multiple of 16, >= column capacity
    uint32_t *bf_key;        // [frame][cols.stride_rows] (d << 16 | prev row) minima
    int T, lane_cols;        // filled by the launcher
    int skip_scores;         // 1: match_score (best_idx / best / second of every row) is not computed
    int use_tc;              // 1: tensor-core tiles (tcham.cu) instead of k_pairs / k_scores_m; dmat is not touched
    // tracked batches (track.cu): the BF train set is another descriptor set than the pass-1 rows (same counts): its
    // per-frame pointer table and operand images; tab == NULL: the row set serves both
    MatchSet bf_last; uint8_t *img_last; size_t img_last_stride;
};
void launch_pass1_fused(const PairArgs &p, const BfArgs &b, int nframes, cudaStream_t st, long long *launches,
                        cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr,   // optional events around k_pairs
                        cudaStream_t st_scores = nullptr, cudaEvent_t e_resolved = nullptr);   // k_scores_m on another stream
// ---- tensor-core Hamming tiles (tcham.cu): tcgen05.mma.kind::i8 over +-1-expanded descriptors ----
enum { TC_PAIRS = 0, TC_SCORES = 1, TC_SHORT = 2, TC_DUMP = 3 };
struct TcExpandArgs {            // descriptor set -> operand images ([frame][tile of 128 rows][SVO_TC_TILE_BYTES])
    MatchSet set;
    const uint16_t *index;       // optional ascending gather list ([frame][index_stride]) and its length per frame
    const int *index_cnt;
    int index_stride;
    const int *index32;          // the same as 32-bit entries in any order ([frame][index_stride]; length at index_cnt[frame * index_cnt_stride])
    int index_cnt_stride;        // 0 is read as 1
    uint8_t *img; size_t img_frame_stride;
};
struct TcArgs {
    MatchSet A, B;               // A rows = tile rows (one epilogue thread each), B rows = the columns they scan in ascending order
    const uint8_t *a_img, *b_img;            // their operand images
    size_t a_img_frame_stride, b_img_frame_stride;
    GreedyArgs g;                // per-row / per-column arrays of the pass the mode serves (rows.stride_rows / cols.stride_rows give the strides)
    int T;                       // TC_PAIRS: 15 (pass-1 candidates); TC_SHORT: 60
    uint32_t *bf_key;            // TC_PAIRS: [frame][cols.stride_rows] (d << 16 | first minimum row) per query
    const uint16_t *b_index;     // TC_SHORT: the ascending list of original column indices b_img was gathered with ([frame][b_index_stride])
    const int *b_index_cnt;      //           its length per frame
    int b_index_stride;
    const uint8_t *row_need;     // TC_SHORT: [frame][rows.stride_rows] 1 = the row is scanned
    const int *a_index;          // TC_SHORT: a_img was gathered with this list of scanned rows ([frame][rows.stride_rows], any order;
    const int *a_index_cnt;      //           length at a_index_cnt[2 * frame]); NULL = a_img holds every row in place
    int *dump; int dump_rows, dump_pitch;   // TC_DUMP: [frame][dump_rows][dump_pitch] dot products (256 - 2 d)
    long long *prof;             // optional clock64 timeline of CTA (0, 0): [mode][4 roles][64] (svo_debug_tc_profile)
};
void launch_tc_expand(const TcExpandArgs &e, int nframes, cudaStream_t st, long long *launches);
void launch_tc_expand2(const TcExpandArgs &e0, const TcExpandArgs &e1, int nframes, cudaStream_t st, long long *launches);
void launch_tc_hamming(const TcArgs &p, int mode, int nframes, cudaStream_t st, long long *launches);
int setup_tc_attributes();

// ---- device-resident tracker state (track.cu) ----
struct TrackUpdateArgs {
    const FramePtrs *fp;
    const int *nkp;              // [frame * 2] left keypoint counts of the batch's slots
    const svo_keypoint *kp; const uint8_t *desc;   // left image of frame f at kp + f * 2 * kp_cap
    const int *claim_row; const float *depth; int col_stride;
    const uint8_t *p1_row_bad; int row_stride;
    int kp_cap, map_cap, window;
    int *scratch;                // [frame][map_cap]
    int *mp_create; float *mp_xyz;   // [frame][col_stride], [frame][col_stride][3]: the point each current keypoint owns after the frame
};
void launch_track_load(const FramePtrs *fp, int *n_prev, int *n_map, int n, cudaStream_t st, long long *launches);
void launch_track_update(const TrackUpdateArgs &a, int n, cudaStream_t st, long long *launches);

// ---- pose stage (pose.cu) ----
struct PoseHdr { int off, n; float fx, fy, cx, cy; float Tcw[16]; };   // one problem: points [off, off + n) of the packed arrays
struct PoseArgs {
    const PoseHdr *hdr;
    const float *p3, *p2;      // packed n x 3 world points / n x 2 pixels of all problems
    double *hyp; float *hypf;  // [problem][iterations * 4][12] P3P solutions (R row-major, t), f64 and the f32 copy that is scored
    uint8_t *mask;             // packed inlier flags
    int *info;                 // [problem][4]: inliers, winning iteration, its solution index, hypotheses scored
    double *pose_out;          // [problem][12]: refined R (row-major), t
    int ransac_iterations; float thr2; uint32_t seed; int refine_iterations;
    int lm_iterations; float *Tcw_out; double *lm_stats;   // k_pose_lm: [problem][16] pose out, [problem][2] (iterations, chi2)
};
void launch_pnp_ransac(const PoseArgs &a, int nproblems, int max_n, cudaStream_t st, long long *launches);
void launch_pose_lm(const PoseArgs &a, int nproblems, cudaStream_t st, long long *launches);
int setup_pose();
int pose_max_iterations();
void launch_project(const float *xyz, const int *octave, int n, const float *Tcw12, float fx, float fy, float cx, float cy, int W, int H,
                    float th, const float *lscale, int nlevels, float *uvr, cudaStream_t st, long long *launches);
void launch_disp2depth(const float *disp, float *depth, size_t n, float bf, cudaStream_t st, long long *launches);
int setup_match_attributes();
int greedy_max_cols();   // most columns (current-frame keypoints) the greedy resolver supports
