// svo_api.cu — context, HBM layout, streams and the C ABI of libsvo_b200.so
// (include/svo_b200.h).  Host-side orchestration only; every computation is a kernel in
// pyramid.cu / fast.cu / select.cu / describe.cu / stereo.cu / match.cu.  No CPU fallback.
#include "svo_internal.cuh"
#include "resize_quads.h"

#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#define N_EVENTS 16

namespace {

struct FrameBufs {           // per-frame (stereo pair) device arrays
    int nframes, col_stride, row_stride;
    uint8_t *prev, *map;     // [f][row_stride][32]
    uint8_t *prev_live;      // [f][row_stride]
    int *map_prev;           // [f][row_stride]
    uint8_t *claimed;        // [f][col_stride]
    int *claim_row, *claim_time, *claim_time2;
    int *bf_idx, *bf_dist; uint8_t *bf_keep; int *min_dist;
    int *p1_best_idx, *p1_best, *p1_second; uint8_t *p1_row_claimed, *p1_row_bad;
    int *p2_best_idx, *p2_best, *p2_second; uint8_t *p2_row_claimed;
    uint32_t *shortlist, *shortlist_hi; int *short_cnt;
    int *res_rows, *res_off, *res_want, *res_perm;
    int *need_list, *reuse_list, *list_cnt; uint8_t *row_need;
    uint8_t *img_cur, *img_prev, *img_map, *img_free; size_t img_col_stride, img_row_stride;   // tensor-core operand images (tcham.cu)
    uint8_t *dmat; uint32_t *bf_key; int dmat_pitch; size_t dmat_frame_stride;   // fused pass-1 front (batch path)
    float *u_right, *depth; int *match_r, *sad, *n_stereo;
    int *row_off; uint16_t *row_list; int row_list_stride;
    int *params;             // [4][nframes]: n_prev, n_map, bf bits, baseline bits
    // sync-only extras
    uint8_t *cols;           // [col_stride][32] caller-provided column descriptors
    float *win, *cur_xy, *row_xy; int *boxes; double *F;
    // batch extras: keypoint grid of the windowed pass 2 (win / cur_xy are allocated for batch frames too)
    int *cell_off; uint16_t *cell_list;
    uint16_t *free_col; int *free_cnt;   // batch pass 2: columns pass 1 left free
};

struct HostArena {           // pinned mirror of one lane's outputs
    svo_keypoint *kp; uint8_t *desc; int *nkp, *status;
    float *u_right, *depth; int *n_stereo;
    int *bf_idx, *bf_dist; uint8_t *bf_keep;
    int *p1_best_idx, *p1_best, *p1_second; uint8_t *p1_row_claimed, *p1_row_bad, *p2_row_claimed;
    int *claim_row;
    int *np_out;             // [2][B] n_prev / n_map as the device used them
    int *mp_create; float *mp_xyz;   // tracked frames (allocated by svo_track_create)
    int *params;             // staging for the per-batch parameter upload
};

struct PoseBufs {            // pose stage: packed problems of one call (device) + pinned staging
    int cap_prob; size_t cap_pts;
    PoseHdr *d_hdr, *h_hdr;
    float *d_p3, *d_p2, *h_p3, *h_p2;
    uint8_t *d_mask;
    double *d_hyp; float *d_hypf;
    int *d_info, *h_info;
    double *d_pose, *h_pose, *d_stats, *h_stats;
    float *d_T, *h_T;
};

struct LaneGraph { int key; cudaGraphExec_t exec; long long launches; };

struct Lane {
    std::vector<LaneGraph> graphs;   // captured compute sequences, keyed by (n, stages)
    cudaStream_t st;
    bool own_stream;
    cudaStream_t side;         // second branch of the batch graph: blur beside FAST/selection, stereo + keypoint D2H beside matching
    cudaEvent_t fk[6];         // fork/join events of the two branches
    cudaEvent_t done;
    cudaEvent_t ev[N_EVENTS];
    int slot0, frame0, nframes;
    bool busy, veto, tracked;
    int out_flags, out_w, out_wp;   // of the batch in flight: output selection and the rows its result copies carried
    HostArena h;
    std::vector<svo_frame_in> in;
    uint8_t *d_stage;          // landing zone for the host inputs of a batch (images, descriptors, flags)
    size_t stage_cap;
    FramePtrs *d_fp, *h_fp;    // per-frame input pointers (device table + pinned staging)
    int *d_strides, *h_strides; // per image: row stride at its source, 0 = uploaded with a 2-D copy
};

}  // namespace

struct svo_ctx {
    svo_config cfg;
    Geom g;
    Bufs b;
    FrameBufs fb;            // batch frames
    FrameBufs sb;            // the single synchronous frame
    PoseBufs pb;
    std::vector<Lane> lanes;
    cudaStream_t sync_st;
    int sync_slot0;          // two slots: cam 0 / cam 1
    int nslots;
    std::vector<void *> dev_allocs, pinned_allocs;
    std::vector<uint32_t> rtab_host;
    std::vector<ResizeQuad> rq_host;
    long long launches;
    size_t stage_img_bytes;
    uint8_t *sync_stage;     // landing zone of svo_extract_bgr (one colour image)
    FramePtrs *sync_fp_d, *sync_fp_h;
    int *sync_str_d, *sync_str_h;
    bool profiling;
    int out_flags;           // svo_set_outputs
    int compact_rows;        // rows per frame of a compact result copy (nfeatures + 32; SVO_B200_COMPACT_ROWS overrides it for tests)
    // device-resident tracker states (svo_track_create)
    int trk_n, trk_cap, trk_window;
    TrackState *trk_d;                   // [seq][2] on the device
    std::vector<TrackState> trk_h;       // host mirror of the pointer values
    std::vector<int> trk_parity;         // which copy is current
    std::vector<int> trk_last_lane;      // lane whose batch last advanced the sequence, -1 = none
    int *trk_scratch;                    // [batch frames][map_cap]
    int *trk_mp_create; float *trk_mp_xyz;   // [batch frames][kp_cap], [..][3]: point owned by each current keypoint after the frame
    uint8_t *trk_img_last;               // [batch frames] operand images of the last frames' own descriptors (BF train set)
    long long *tc_prof;      // in-kernel timeline of the tensor-core matchers (SVO_B200_TC_PROF=1), else NULL
    bool use_tc;             // tensor-core Hamming tiles in the batch matchers (default; SVO_B200_TC=0 selects the SIMT kernels)
    bool sync_have[2];
    char err[512];
};

namespace {

int fail(svo_ctx *c, int code, const char *fmt, ...)
{
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(ctx, SVO_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
int dalloc(svo_ctx *ctx, T **p, size_t n)
{
    void *q = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, SVO_E_NOMEM, "cudaMalloc(%zu) -> %s", n * sizeof(T), cudaGetErrorString(e));
    ctx->dev_allocs.push_back(q);
    cudaMemset(q, 0, n * sizeof(T));   // defined contents: vector loads may touch (and ignore) entries past a list's end
    *p = (T *)q;
    return SVO_OK;
}
template <typename T>
int halloc(svo_ctx *ctx, T **p, size_t n)
{
    void *q = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMallocHost(&q, n * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, SVO_E_NOMEM, "cudaMallocHost(%zu) -> %s", n * sizeof(T), cudaGetErrorString(e));
    ctx->pinned_allocs.push_back(q);
    memset(q, 0, n * sizeof(T));
    *p = (T *)q;
    return SVO_OK;
}
#define TRY(x) do { int r_ = (x); if (r_ != SVO_OK) return r_; } while (0)

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// cv::ORB geometry: float scale = (float)pow((double)scaleFactor, l), sizes by
// cvRound(cols * (1 / scale)), quotas by the geometric series in float.
void orb_geometry(int W, int H, int nlevels, float scale_factor_f, int nfeatures, int *lw, int *lh, float *ls, int *quota)
{
    const double sf = (double)scale_factor_f;
    for (int l = 0; l < nlevels; ++l) {
        const float s = (float)pow(sf, (double)l);
        ls[l] = s;
        // cvRound(cols / scale) as the OpenCV build the parity oracle is pinned to (4.13.0) evaluates it: a float
        // multiplication by the reciprocal.  Differs from the quotient only where cols / scale is within an ulp of k + 0.5
        // (at 1.2: 140 of the dimensions below 4096, e.g. 249 -> 208; none of KITTI's)
        volatile float inv = 1.0f / s;
        lw[l] = (int)lrintf((float)W * inv);
        lh[l] = (int)lrintf((float)H * inv);
    }
    const float factor = (float)(1.0 / sf);
    float ndes = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        quota[l] = (int)lrintf(ndes);
        sum += quota[l];
        ndes *= factor;
    }
    quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

int build_geometry(svo_ctx *ctx)
{
    const svo_config &c = ctx->cfg;
    Geom &g = ctx->g;
    memset(&g, 0, sizeof(g));
    g.nlevels = c.nlevels; g.W = c.width; g.H = c.height; g.fast_threshold = c.fast_threshold;
    {   // FAST band height (fast.cu).  16 rows halve the halo work of 8-row bands and keep more warps per SM (k_fast: 340 ->
        // 312 us per 64 images), but a single frame is spread over half as many CTAs (single-frame p50 +9 us): batch contexts
        // get 16, latency contexts (max_batch < 8) 8.  Candidate positions are 16-bit offsets into the band's (band + 2)
        // staged rows, which limits 16-row bands to a level-0 pitch of 3640 bytes.  SVO_B200_FAST_BAND=8|16 overrides.
        const int pitch0 = align_up(c.width, 16);
        g.fast_band = (c.max_batch >= 8 && 18 * pitch0 <= 65535) ? 16 : 8;
        const char *e = getenv("SVO_B200_FAST_BAND");
        if (e && atoi(e) == 8) g.fast_band = 8;
        if (e && atoi(e) == 16 && 18 * pitch0 <= 65535) g.fast_band = 16;
    }
    { const char *e = getenv("SVO_B200_HARRIS8"); g.harris8 = !(e && e[0] == '0'); }
    { const char *e = getenv("SVO_B200_BLUR_MARGIN"); g.blur_margin = !(e && e[0] == '0'); }
    int lw[SVO_MAX_LEVELS], lh[SVO_MAX_LEVELS], quota[SVO_MAX_LEVELS];
    float ls[SVO_MAX_LEVELS];
    orb_geometry(c.width, c.height, c.nlevels, c.scale_factor, c.nfeatures, lw, lh, ls, quota);
    int off = 0, band_off = 0, bandcnt_off = 0, cand_off = 0, off2 = 0, tab_off = 0, tiles = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        LevelGeom &L = g.lv[l];
        L.w = lw[l]; L.h = lh[l]; L.pitch = align_up(lw[l], 16);
        L.pitch_magic = (unsigned)((0x100000000ull + L.pitch - 1) / L.pitch);
        if (L.w < 8 || L.h < 8) return fail(ctx, SVO_E_INVALID, "level %d is %dx%d: image too small for %d levels", l, L.w, L.h, g.nlevels);
        L.off = off; off += align_up(L.pitch * (L.h + 1), 256);
        L.scale = ls[l]; L.inv_scale = 1.f / ls[l]; L.quota = quota[l];
        L.x0 = SVO_EDGE; L.x1 = L.w - SVO_EDGE; L.y0 = SVO_EDGE; L.y1 = L.h - SVO_EDGE;
        if (L.x1 <= L.x0 || L.y1 <= L.y0) { L.x1 = L.x0; L.y1 = L.y0; L.nbands = 0; }
        else L.nbands = (L.y1 - L.y0 + g.fast_band - 1) / g.fast_band;
        if (L.nbands > 500) return fail(ctx, SVO_E_INVALID, "image too tall");
        {   // octree mode: nIni = round(width / height) roots of float width hX (ORB-SLAM2 DistributeOctTree)
            const int ow = L.x1 - L.x0, oh = L.y1 - L.y0;
            int ni = oh > 0 ? (int)roundf((float)ow / (float)oh) : 1;
            ni = ni < 1 ? 1 : (ni > 16 ? 16 : ni);
            L.oct_nini = ni; L.oct_hx = (float)ow / (float)ni;
        }
        L.band_cap = ((g.fast_band + 1) / 2) * ((L.x1 - L.x0 + 1) / 2) + 1;
        L.band_off = band_off; band_off += L.nbands * L.band_cap;
        L.bandcnt_off = bandcnt_off; bandcnt_off += L.nbands;
        L.cand_cap = L.nbands * L.band_cap + 4;
        L.cand_off = cand_off; cand_off += align_up(L.cand_cap, 4);
        L.cap2 = 4 * L.quota + 1024 < L.cand_cap ? 4 * L.quota + 1024 : L.cand_cap;
        L.off2 = off2; off2 += align_up(L.cap2, 4);
        L.tab_off = tab_off; if (l) tab_off += L.w + L.h;
        {   // blur tiles: SVO_BLUR_ROWS rows x blur_tq quads, the quads of a row split over the fewest tiles of <= 128 quads.
            // The even share is rounded up to whole warps (one lane per quad): every tile but the row's last one is then
            // made of full warps and the last one's idle warps exit at once — 94 % of the launched lanes hold a quad at
            // 1241x376 against 86 % for evenly sized tiles (104 quads = 3.25 warps).  SVO_B200_BLUR_PACK=0: even tiles.
            const int quads = (L.w + 3) / 4;
            L.blur_tiles_x = (quads + 127) / 128;
            const char *e = getenv("SVO_B200_BLUR_PACK");
            L.blur_tq = align_up((quads + L.blur_tiles_x - 1) / L.blur_tiles_x, (e && e[0] == '0') ? 4 : 32);
            L.blur_tile_off = tiles;                         // first blur CTA of the level
            tiles += L.blur_tiles_x * ((L.h + SVO_BLUR_ROWS - 1) / SVO_BLUR_ROWS);
        }
        g.fast_bands += L.nbands;
    }
    if (lw[0] >= 4096 || lh[0] >= 4096) return fail(ctx, SVO_E_INVALID, "images up to 4095x4095 are supported");
    g.pyr_bytes = off; g.band_total = band_off; g.bandcnt_total = bandcnt_off;
    g.cand_total = cand_off; g.total2 = off2; g.blur_tiles = tiles;
    g.kp_cap = align_up(c.nfeatures + c.nfeatures / 8 + 64, 64);
    ctx->rtab_host.assign((size_t)tab_off + 1, 0);
    int rq_off = 0;
    for (int l = 1; l < g.nlevels; ++l) { g.lv[l].rq_off = rq_off; rq_off += g.lv[l].pitch / 4; }
    ctx->rq_host.assign((size_t)rq_off + 1, ResizeQuad());
    const char *rq_env = getenv("SVO_B200_RESIZE_QUADS");       // =0: the per-byte kernel on every level
    for (int l = 1; l < g.nlevels; ++l) {
        resize_table(lw[l - 1], lw[l], ctx->rtab_host.data() + g.lv[l].tab_off);
        resize_table(lh[l - 1], lh[l], ctx->rtab_host.data() + g.lv[l].tab_off + lw[l]);
        g.lv[l].rq_ok = build_resize_quads(ctx->rtab_host.data() + g.lv[l].tab_off, lw[l], g.lv[l].pitch, ctx->rq_host.data() + g.lv[l].rq_off);
        if (rq_env && rq_env[0] == '0') g.lv[l].rq_ok = 0;
    }
    return SVO_OK;
}

// Bands of the fused pyramid kernel (pyramid.cu:k_pyramid).  Band j owns rows [j R, (j + 1) R) of the last level; going
// down the levels, the rows a band must compute are the rows the level above reads (from the same y tables the kernel
// uses: source row ofs and, when the second tap's weight is not zero, ofs + 1) plus the rows it owns, and a band owns at
// each level the rows from its own first needed row up to the next band's.  R is the largest band height whose shared
// memory fits `budget`; 0 bands = the geometry does not fit at all (very wide images) and the per-level kernels run.
int build_pyramid_bands(svo_ctx *ctx, std::vector<PyrBand> &bands, int budget)
{
    Geom &g = ctx->g;
    const int NL = g.nlevels;
    g.pyr_nbands = 0; g.pyr_smem = 0;
    bands.clear();
    if (NL < 2) return SVO_OK;
    const uint32_t *tab = ctx->rtab_host.data();
    auto yo = [&](int l, int y) { return (int)(tab[g.lv[l].tab_off + g.lv[l].w + y] >> 9); };
    auto yn = [&](int l, int y) { const uint32_t t = tab[g.lv[l].tab_off + g.lv[l].w + y]; return (int)(t >> 9) + ((t & 511u) ? 1 : 0); };
    const int hl = g.lv[NL - 1].h;
    for (int R = hl; R >= 1; --R) {
        const int nb = (hl + R - 1) / R;
        if (R > 1 && (long long)nb < 8) continue;          // keep at least a few bands per image
        std::vector<PyrBand> bs((size_t)nb);
        std::vector<int> lo((size_t)nb + 1);
        for (int j = 0; j < nb; ++j) {
            PyrBand &B = bs[(size_t)j];
            memset(&B, 0, sizeof(B));
            B.olo[NL - 1] = B.clo[NL - 1] = (short)(j * R);
            B.ohi[NL - 1] = (short)std::min((j + 1) * R, hl);
            B.chi[NL - 1] = (short)(B.ohi[NL - 1] - 1);
        }
        for (int l = NL - 2; l >= 0; --l) {
            for (int j = 0; j < nb; ++j) lo[(size_t)j] = j == 0 ? 0 : yo(l + 1, bs[(size_t)j].clo[l + 1]);
            lo[(size_t)nb] = g.lv[l].h;
            for (int j = 0; j < nb; ++j) {
                PyrBand &B = bs[(size_t)j];
                const int need_lo = yo(l + 1, B.clo[l + 1]), need_hi = yn(l + 1, B.chi[l + 1]);
                B.olo[l] = (short)lo[(size_t)j]; B.ohi[l] = (short)std::max(lo[(size_t)j + 1], lo[(size_t)j]);
                B.clo[l] = (short)std::min(need_lo, (int)B.olo[l]);
                B.chi[l] = (short)std::min(std::max(need_hi, B.ohi[l] - 1), g.lv[l].h - 1);
            }
        }
        int off = 0, soff[SVO_MAX_LEVELS];
        for (int l = 0; l < NL; ++l) {
            int rows = 0;
            for (int j = 0; j < nb; ++j) rows = std::max(rows, bs[(size_t)j].chi[l] - bs[(size_t)j].clo[l] + 1);
            soff[l] = off;
            off += align_up(rows * g.lv[l].pitch + 16, 128);   // + 16: the quad-table loads read whole words up to 11 bytes past a row's width
        }
        if (off > budget) continue;
        bands = bs;
        g.pyr_nbands = nb; g.pyr_smem = off;
        for (int l = 0; l < NL; ++l) g.pyr_soff[l] = soff[l];
        return SVO_OK;
    }
    return SVO_OK;
}

int alloc_frames(svo_ctx *ctx, FrameBufs &f, int nframes, int col_stride, int row_stride, bool sync_extras)
{
    f.nframes = nframes; f.col_stride = col_stride; f.row_stride = row_stride;
    const size_t F = nframes, C = (size_t)col_stride * F, R = (size_t)row_stride * F;
    TRY(dalloc(ctx, &f.prev, R * 32)); TRY(dalloc(ctx, &f.map, R * 32));
    TRY(dalloc(ctx, &f.prev_live, R)); TRY(dalloc(ctx, &f.map_prev, R));
    TRY(dalloc(ctx, &f.claimed, C)); TRY(dalloc(ctx, &f.claim_row, C)); TRY(dalloc(ctx, &f.claim_time, C)); TRY(dalloc(ctx, &f.claim_time2, C));
    TRY(dalloc(ctx, &f.bf_idx, C)); TRY(dalloc(ctx, &f.bf_dist, C)); TRY(dalloc(ctx, &f.bf_keep, C));
    TRY(dalloc(ctx, &f.min_dist, F));
    TRY(dalloc(ctx, &f.p1_best_idx, R)); TRY(dalloc(ctx, &f.p1_best, R)); TRY(dalloc(ctx, &f.p1_second, R));
    TRY(dalloc(ctx, &f.p1_row_claimed, R)); TRY(dalloc(ctx, &f.p1_row_bad, R));
    TRY(dalloc(ctx, &f.p2_best_idx, R)); TRY(dalloc(ctx, &f.p2_best, R)); TRY(dalloc(ctx, &f.p2_second, R));
    TRY(dalloc(ctx, &f.p2_row_claimed, R));
    TRY(dalloc(ctx, &f.shortlist, R * 32)); TRY(dalloc(ctx, &f.shortlist_hi, R * (SVO_SHORT_CAP - 32))); TRY(dalloc(ctx, &f.short_cnt, R));
    TRY(dalloc(ctx, &f.res_rows, R)); TRY(dalloc(ctx, &f.res_off, R)); TRY(dalloc(ctx, &f.res_want, R)); TRY(dalloc(ctx, &f.res_perm, R));
    TRY(dalloc(ctx, &f.need_list, R)); TRY(dalloc(ctx, &f.reuse_list, R)); TRY(dalloc(ctx, &f.list_cnt, 2 * F)); TRY(dalloc(ctx, &f.row_need, R));
    TRY(dalloc(ctx, &f.u_right, C)); TRY(dalloc(ctx, &f.depth, C)); TRY(dalloc(ctx, &f.match_r, C));
    TRY(dalloc(ctx, &f.sad, C)); TRY(dalloc(ctx, &f.n_stereo, F));
    {   // a right keypoint is a candidate for rows floor(y - r) .. ceil(y + r), r = 2 * scale[octave]
        int band = 2 * (int)ceilf(2.f * ctx->g.lv[ctx->g.nlevels - 1].scale) + 2;
        if (band > ctx->g.H) band = ctx->g.H;
        f.row_list_stride = ctx->g.kp_cap * band;
    }
    TRY(dalloc(ctx, &f.row_off, F * (ctx->g.H + 1))); TRY(dalloc(ctx, &f.row_list, F * f.row_list_stride));
    TRY(dalloc(ctx, &f.params, 4 * F));
    // operand images of the tensor-core matchers: whole tiles of 128 descriptors x 256 int8
    f.img_col_stride = (size_t)((col_stride + 127) / 128) * SVO_TC_TILE_BYTES;
    f.img_row_stride = (size_t)((row_stride + 127) / 128) * SVO_TC_TILE_BYTES;
    TRY(dalloc(ctx, &f.img_cur, F * f.img_col_stride)); TRY(dalloc(ctx, &f.img_map, F * f.img_row_stride));
    f.img_prev = f.img_free = nullptr;
    if (!sync_extras) { TRY(dalloc(ctx, &f.img_prev, F * f.img_row_stride)); TRY(dalloc(ctx, &f.img_free, F * f.img_col_stride)); }
    f.cols = nullptr; f.win = f.cur_xy = f.row_xy = nullptr; f.boxes = nullptr; f.F = nullptr;
    f.dmat = nullptr; f.bf_key = nullptr; f.dmat_pitch = 0; f.dmat_frame_stride = 0;
    f.cell_off = nullptr; f.cell_list = nullptr;
    f.free_col = nullptr; f.free_cnt = nullptr;
    if (!sync_extras) {
        TRY(dalloc(ctx, &f.free_col, C)); TRY(dalloc(ctx, &f.free_cnt, F));
        TRY(dalloc(ctx, &f.win, R * 3)); TRY(dalloc(ctx, &f.cur_xy, C * 2));
        TRY(dalloc(ctx, &f.cell_off, F * (SVO_WIN_CELLS + 1))); TRY(dalloc(ctx, &f.cell_list, C));
        // u8 distance matrix previous-frame rows x current-frame columns; k_scores_m gives every lane a
        // 16-byte aligned block of columns, so the pitch is 32 such blocks
        f.dmat_pitch = 32 * ((((col_stride + 31) / 32) + 15) & ~15);
        f.dmat_frame_stride = (size_t)col_stride * f.dmat_pitch;
        TRY(dalloc(ctx, &f.dmat, F * f.dmat_frame_stride));
        TRY(dalloc(ctx, &f.bf_key, C));
    }
    if (sync_extras) {
        TRY(dalloc(ctx, &f.cols, C * 32));
        TRY(dalloc(ctx, &f.win, R * 3)); TRY(dalloc(ctx, &f.cur_xy, C * 2)); TRY(dalloc(ctx, &f.row_xy, R * 2));
        TRY(dalloc(ctx, &f.boxes, 4 * 256)); TRY(dalloc(ctx, &f.F, 9));
    }
    return SVO_OK;
}

// enqueue the extraction kernels for images [slot0, slot0+nimg) on `st`
// side != nullptr: the blur (which only needs the pyramid) runs on `side` beside FAST -> cull -> Harris -> cull and
// joins before the descriptors (captured into the batch graph as a fork/join)
void enqueue_extract(svo_ctx *ctx, int slot0, int nimg, cudaStream_t st, cudaEvent_t *ev, cudaStream_t side = nullptr,
                     cudaEvent_t e_fork = nullptr, cudaEvent_t e_join = nullptr)
{
    const Bufs &b = ctx->b; const Geom &g = ctx->g;
    long long *n = &ctx->launches;
    cudaMemsetAsync(b.status + slot0, 0, sizeof(int) * nimg, st);
    if (ev) cudaEventRecord(ev[1], st);
    launch_pyramid(b, g, slot0, nimg, st, n);
    if (side) {
        cudaEventRecord(e_fork, st);
        cudaStreamWaitEvent(side, e_fork, 0);
        launch_blur(b, g, slot0, nimg, side, n);
        cudaEventRecord(e_join, side);
    }
    if (ev) cudaEventRecord(ev[2], st);
    launch_fast(b, g, slot0, nimg, st, n);
    if (ev) cudaEventRecord(ev[3], st);
    const bool octree = ctx->cfg.distribution == SVO_DIST_OCTREE;
    if (octree) launch_octree(b, g, slot0, nimg, st, n);
    else launch_select1(b, g, slot0, nimg, st, n);
    if (ev) cudaEventRecord(ev[4], st);
    launch_harris(b, g, slot0, nimg, st, n);
    if (ev) cudaEventRecord(ev[5], st);
    if (octree) launch_keep_all(b, g, slot0, nimg, st, n);
    else launch_select2(b, g, slot0, nimg, st, n);
    if (ev) cudaEventRecord(ev[6], st);
    if (side) cudaStreamWaitEvent(st, e_join, 0);
    else launch_blur(b, g, slot0, nimg, st, n);
    if (ev) cudaEventRecord(ev[7], st);
    launch_describe(b, g, slot0, nimg, st, n);
    if (ev) cudaEventRecord(ev[8], st);
}

int upload_image(svo_ctx *ctx, int slot, const uint8_t *img, int stride, cudaStream_t st)
{
    const Geom &g = ctx->g;
    uint8_t *dst = ctx->b.pyr + (size_t)slot * g.pyr_bytes + g.lv[0].off;
    CU(cudaMemcpy2DAsync(dst, g.lv[0].pitch, img, stride, g.W, g.H, cudaMemcpyDefault, st));
    return SVO_OK;
}

MatchSet make_set(const uint8_t *desc, const int *count, int count_stride, int stride_rows, int fixed, int desc_stride = -1)
{
    MatchSet s; s.desc = desc; s.count = count; s.count_stride = count_stride; s.stride_rows = stride_rows; s.fixed_count = fixed;
    s.tab = nullptr;
    s.desc_stride = desc_stride < 0 ? stride_rows : desc_stride;
    return s;
}

// true when kernels can read `p` in place (device or managed memory)
bool device_readable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

struct Seg { const uint8_t *src; size_t bytes; const void **field; bool need16; };

// Kernels, memsets and D2H copies of one batch on the lane's stream (capturable: no host-dependent arguments).
// phase 0: everything; 1: input repack + extraction only; 2: stereo, matching, tracker update and result copies only (the
// two halves of a TRACKED batch: the second waits for the previous batch to have advanced the tracker states)
// Result copies.  out_w rows of every per-keypoint array and out_wp rows of every pass-1 array leave per frame: the
// capacities by default, or (SVO_OUT_COMPACT) nfeatures + 32 rows as strided 2-D copies — svo_batch_wait fetches the
// rest of a frame that went over (response ties can push a level past its quota).  SVO_OUT_NO_RIGHT leaves the right
// image's keypoints and descriptors on the device (the reference's frame has none: src/frame.cc:122-138 keeps only
// keypoints_r = u_right).
cudaError_t copy_rows(void *dst, const void *src, size_t elem, size_t pitch_rows, size_t rows, size_t frames, cudaStream_t st)
{
    if (rows >= pitch_rows) return cudaMemcpyAsync(dst, src, elem * pitch_rows * frames, cudaMemcpyDeviceToHost, st);
    return cudaMemcpy2DAsync(dst, elem * pitch_rows, src, elem * pitch_rows, elem * rows, frames, cudaMemcpyDeviceToHost, st);
}

int enqueue_compute(svo_ctx *ctx, Lane &L, int n, bool any_prev, bool any_map, bool fused, bool windowed, bool veto, bool tracked,
                    int phase, cudaEvent_t *ev)
{
    const Geom &g = ctx->g;
    const Bufs &b = ctx->b;
    FrameBufs &fb = ctx->fb;
    cudaStream_t st = L.st;
    const int R = fb.row_stride, K = fb.col_stride;
    const int FT = fb.nframes;
    // Outside profiling runs the sequence has two branches (it is captured into a graph, so they are graph forks):
    // blur beside the keypoint selection, and stereo + the keypoint/descriptor/stereo D2H beside the matchers.
    const bool fork = ev == nullptr && L.side != nullptr;
    if (phase != 2) {
        launch_unpack(b, g, L.slot0, 2 * n, L.d_fp, L.d_strides, st, &ctx->launches);
        CU(cudaMemsetAsync(fb.claimed + (size_t)L.frame0 * K, 0, (size_t)n * K, st));
        CU(cudaMemsetAsync(fb.claim_row + (size_t)L.frame0 * K, 0xff, sizeof(int) * (size_t)n * K, st));
        // ---- extraction of 2n images
        enqueue_extract(ctx, L.slot0, 2 * n, st, ev, fork ? L.side : nullptr, L.fk[0], L.fk[1]);
        if (phase == 1) return SVO_OK;
    }
    if (tracked)   // the tracked frames' row counts come from their states
        launch_track_load(L.d_fp, fb.params + L.frame0, fb.params + FT + L.frame0, n, st, &ctx->launches);
    cudaStream_t ss = fork ? L.side : st;     // stream of the stereo branch
    if (fork) { CU(cudaEventRecord(L.fk[2], st)); CU(cudaStreamWaitEvent(ss, L.fk[2], 0)); }
    // ---- sparse stereo
    const int *d_nprev = fb.params + L.frame0, *d_nmap = fb.params + FT + L.frame0;
    StereoArgs sa;
    sa.u_right = fb.u_right + (size_t)L.frame0 * K; sa.depth = fb.depth + (size_t)L.frame0 * K;
    sa.match_r = fb.match_r + (size_t)L.frame0 * K; sa.sad = fb.sad + (size_t)L.frame0 * K;
    sa.n_stereo = fb.n_stereo + L.frame0; sa.stride = K;
    sa.row_off = fb.row_off + (size_t)L.frame0 * (g.H + 1); sa.row_list = fb.row_list + (size_t)L.frame0 * fb.row_list_stride;
    sa.row_list_stride = fb.row_list_stride;
    sa.bf = reinterpret_cast<const float *>(fb.params + 2 * (size_t)FT + L.frame0);
    sa.baseline = reinterpret_cast<const float *>(fb.params + 3 * (size_t)FT + L.frame0);
    launch_stereo(b, g, L.slot0, n, sa, ss, &ctx->launches);
    if (ev) cudaEventRecord(ev[9], st);
    HostArena &h = L.h;
    const size_t I = 2 * (size_t)n, KC = g.kp_cap;
    const size_t W = (size_t)L.out_w, WP = (size_t)L.out_wp;
    const bool no_right = (L.out_flags & SVO_OUT_NO_RIGHT) != 0;
    const bool pose_only = (L.out_flags & SVO_OUT_POSE_INPUTS) != 0;     // implies no_right (svo_set_outputs)
    auto copy_extract = [&](cudaStream_t cs) -> int {   // everything extraction and stereo produced
        CU(cudaMemcpyAsync(h.nkp, b.nkp + L.slot0, sizeof(int) * I, cudaMemcpyDeviceToHost, cs));
        if (no_right) {   // left images sit in the even slots
            CU(copy_rows(h.kp, b.kp + (size_t)L.slot0 * KC, sizeof(svo_keypoint), 2 * KC, W, n, cs));
            if (!pose_only) CU(copy_rows(h.desc, b.desc + (size_t)L.slot0 * KC * 32, 32, 2 * KC, W, n, cs));
        } else {
            CU(copy_rows(h.kp, b.kp + (size_t)L.slot0 * KC, sizeof(svo_keypoint), KC, W, I, cs));
            CU(copy_rows(h.desc, b.desc + (size_t)L.slot0 * KC * 32, 32, KC, W, I, cs));
        }
        if (!pose_only) CU(copy_rows(h.u_right, sa.u_right, sizeof(float), KC, W, n, cs));
        CU(copy_rows(h.depth, sa.depth, sizeof(float), KC, W, n, cs));
        CU(cudaMemcpyAsync(h.n_stereo, sa.n_stereo, sizeof(int) * n, cudaMemcpyDeviceToHost, cs));
        return SVO_OK;
    };
    if (fork) TRY(copy_extract(ss));   // leaves while the matchers run
    // ---- matching: BF (cur -> prev), greedy pass 1 (prev rows), greedy pass 2 (map rows)
    // left images sit in even slots: consecutive frames' descriptor blocks are 2*kp_cap rows apart
    const MatchSet cur = make_set(b.desc + (size_t)L.slot0 * g.kp_cap * 32, b.nkp + L.slot0, 2, g.kp_cap, 0, 2 * g.kp_cap);
    GreedyArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.cols = cur;
    ga.tc_prof = ctx->tc_prof;
    ga.claimed = fb.claimed + (size_t)L.frame0 * K; ga.claim_row = fb.claim_row + (size_t)L.frame0 * K;
    ga.claim_time = fb.claim_time + (size_t)L.frame0 * K;
    ga.shortlist = fb.shortlist + (size_t)L.frame0 * R * 32; ga.shortlist_hi = fb.shortlist_hi + (size_t)L.frame0 * R * (SVO_SHORT_CAP - 32);
    ga.short_cnt = fb.short_cnt + (size_t)L.frame0 * R;
    ga.res_rows = fb.res_rows + (size_t)L.frame0 * R; ga.res_off = fb.res_off + (size_t)L.frame0 * R;
    ga.res_want = fb.res_want + (size_t)L.frame0 * R; ga.res_perm = fb.res_perm + (size_t)L.frame0 * R;
    if (any_prev) {
        BfArgs ba;
        MatchSet prev_set = make_set(nullptr, d_nprev, 1, R, 0);
        prev_set.tab = reinterpret_cast<const uint8_t *const *>(&L.d_fp->prev);
        ba.q = cur; ba.t = prev_set;
        if (tracked) ba.t.tab = reinterpret_cast<const uint8_t *const *>(&L.d_fp->last);   // BF train set: the last frame's own descriptors
        ba.idx = fb.bf_idx + (size_t)L.frame0 * K; ba.dist = fb.bf_dist + (size_t)L.frame0 * K;
        ba.keep = fb.bf_keep + (size_t)L.frame0 * K; ba.min_dist = fb.min_dist + L.frame0;
        ga.rows = prev_set;
        ga.mode = SVO_GREEDY_PASS1; ga.row_base = 0; ga.row_base_arr = nullptr;
        ga.fp = L.d_fp; ga.use_live = 1; ga.use_map_prev = 0;
        ga.best_idx = fb.p1_best_idx + (size_t)L.frame0 * R; ga.best = fb.p1_best + (size_t)L.frame0 * R;
        ga.second = fb.p1_second + (size_t)L.frame0 * R;
        ga.row_claimed = fb.p1_row_claimed + (size_t)L.frame0 * R;
        // the "dynamic" veto: boxes / F / prev_xy of each frame sit behind the pointer table, the current keypoints'
        // positions are the extractor's output
        ga.row_bad = fb.p1_row_bad + (size_t)L.frame0 * R; ga.use_veto = veto ? 1 : 0;
        ga.kp = b.kp + (size_t)L.slot0 * g.kp_cap; ga.kp_frame_stride = 2 * (size_t)g.kp_cap;
        ga.img_rows = fb.img_prev + (size_t)L.frame0 * fb.img_row_stride; ga.img_rows_stride = fb.img_row_stride;
        ga.img_cols = fb.img_cur + (size_t)L.frame0 * fb.img_col_stride; ga.img_cols_stride = fb.img_col_stride;
        if (fused) {
            // BF (cur -> prev) and greedy pass 1 (prev rows over cur columns) share one distance matrix
            PairArgs pa;
            pa.g = ga;
            pa.dmat = fb.dmat + (size_t)L.frame0 * fb.dmat_frame_stride; pa.dmat_frame_stride = fb.dmat_frame_stride;
            pa.dmat_pitch = fb.dmat_pitch; pa.bf_key = fb.bf_key + (size_t)L.frame0 * K; pa.T = 0; pa.lane_cols = 0;
            pa.use_tc = ctx->use_tc ? 1 : 0; pa.skip_scores = ctx->cfg.skip_match_score ? 1 : 0;
            // match_score (k_scores_m) feeds nothing downstream: on the side branch it runs beside pass 2
            memset(&pa.bf_last, 0, sizeof(pa.bf_last)); pa.img_last = nullptr; pa.img_last_stride = 0;
            if (tracked && ctx->use_tc) {   // the BF train set (the last frame's own descriptors) is not the pass-1 row set (frozen m_descriptors)
                pa.bf_last.tab = reinterpret_cast<const uint8_t *const *>(&L.d_fp->last);
                pa.img_last = ctx->trk_img_last + (size_t)L.frame0 * fb.img_col_stride; pa.img_last_stride = fb.img_col_stride;
            }
            launch_pass1_fused(pa, ba, n, st, &ctx->launches, ev ? ev[12] : nullptr, ev ? ev[13] : nullptr,
                               fork ? ss : nullptr, L.fk[4]);
            if (tracked && !ctx->use_tc) launch_bf(ba, n, st, &ctx->launches);   // SIMT matchers: a separate BF scan over that set
        } else {
            launch_bf(ba, n, st, &ctx->launches);
            launch_greedy(ga, n, !ctx->cfg.skip_match_score, st, &ctx->launches);
        }
    }
    if (any_map) {
        ga.rows = make_set(nullptr, d_nmap, 1, R, 0);
        ga.rows.tab = reinterpret_cast<const uint8_t *const *>(&L.d_fp->map);
        ga.mode = SVO_GREEDY_PASS2; ga.row_base = 0; ga.row_base_arr = d_nprev;
        ga.claim_time = fb.claim_time2 + (size_t)L.frame0 * K;   // pass 1's claim times stay readable for k_scores_m
        ga.fp = L.d_fp; ga.use_live = 0; ga.use_map_prev = any_prev ? 1 : 0;
        ga.need_list = fb.need_list + (size_t)L.frame0 * R; ga.reuse_list = fb.reuse_list + (size_t)L.frame0 * R;
        ga.list_cnt = fb.list_cnt + 2 * (size_t)L.frame0;
        if (ctx->use_tc && !windowed) {   // tensor-core tiles scan every live row
            ga.row_need = fb.row_need + (size_t)L.frame0 * R;
            ga.img_rows = fb.img_map + (size_t)L.frame0 * fb.img_row_stride; ga.img_rows_stride = fb.img_row_stride;
            ga.img_cols = fb.img_cur + (size_t)L.frame0 * fb.img_col_stride; ga.img_cols_stride = fb.img_col_stride;
            ga.img_free = fb.img_free + (size_t)L.frame0 * fb.img_col_stride; ga.img_free_stride = fb.img_col_stride;
            ga.img_cols_ready = (any_prev && fused) ? 1 : 0;
        }
        if (fused && !ga.row_need) {
            ga.dmat = fb.dmat + (size_t)L.frame0 * fb.dmat_frame_stride; ga.dmat_frame_stride = fb.dmat_frame_stride;
            ga.dmat_pitch = fb.dmat_pitch;
            ga.prev = make_set(nullptr, d_nprev, 1, R, 0);
            ga.prev.tab = reinterpret_cast<const uint8_t *const *>(&L.d_fp->prev);
        }
        ga.prev_row_claimed = fb.p1_row_claimed + (size_t)L.frame0 * R; ga.prev_stride = R;
        ga.prev_row_bad = (any_prev && veto) ? fb.p1_row_bad + (size_t)L.frame0 * R : nullptr;
        ga.prev_count = d_nprev; ga.use_veto = 0;
        ga.best_idx = nullptr; ga.best = nullptr; ga.second = nullptr;
        ga.row_claimed = fb.p2_row_claimed + (size_t)L.frame0 * R; ga.row_bad = nullptr;
        if (windowed) {   // opt-in projection windows: gather from the keypoint grid instead of scanning every column
            int shift = 5;
            while (((g.W >> shift) + 1) * ((g.H >> shift) + 1) > SVO_WIN_CELLS) ++shift;
            ga.win_gather = 1; ga.cell_shift = shift; ga.ncx = (g.W >> shift) + 1; ga.ncy = (g.H >> shift) + 1;
            ga.img_w = g.W; ga.img_h = g.H; ga.nlevels = g.nlevels;
            for (int l = 0; l < SVO_MAX_LEVELS; ++l) ga.lscale[l] = l < g.nlevels ? g.lv[l].scale : 1.f;
            ga.cell_off = fb.cell_off + (size_t)L.frame0 * (SVO_WIN_CELLS + 1); ga.cell_list = fb.cell_list + (size_t)L.frame0 * K;
            ga.kp = b.kp + (size_t)L.slot0 * g.kp_cap; ga.kp_frame_stride = 2 * (size_t)g.kp_cap;
            ga.win_out = fb.win + (size_t)L.frame0 * R * 3; ga.cur_xy_out = fb.cur_xy + (size_t)L.frame0 * K * 2;
            ga.win_uvr = ga.win_out; ga.cur_xy = ga.cur_xy_out;
        }
        else if (any_prev) {   // pass 1 ran: its claims hide columns from every pass-2 row
            ga.free_col = fb.free_col + (size_t)L.frame0 * K; ga.free_cnt = fb.free_cnt + L.frame0;
        }
        launch_greedy(ga, n, false, st, &ctx->launches, ev ? ev[14] : nullptr, ev ? ev[15] : nullptr);
    }
    if (ev) cudaEventRecord(ev[10], st);
    // ---- D2H: one copy per output array for the whole batch
    if (fork) { CU(cudaEventRecord(L.fk[3], ss)); CU(cudaStreamWaitEvent(st, L.fk[3], 0)); }   // join the side branch
    if (tracked) {   // advance the tracker states (needs the claims and the stereo depths)
        TrackUpdateArgs ta;
        ta.fp = L.d_fp; ta.nkp = b.nkp + L.slot0; ta.kp = b.kp + (size_t)L.slot0 * g.kp_cap; ta.desc = b.desc + (size_t)L.slot0 * g.kp_cap * 32;
        ta.claim_row = fb.claim_row + (size_t)L.frame0 * K; ta.depth = fb.depth + (size_t)L.frame0 * K; ta.col_stride = K;
        ta.p1_row_bad = veto ? fb.p1_row_bad + (size_t)L.frame0 * R : nullptr; ta.row_stride = R;
        ta.kp_cap = g.kp_cap; ta.map_cap = ctx->trk_cap; ta.window = ctx->trk_window;
        ta.scratch = ctx->trk_scratch + (size_t)L.frame0 * ctx->trk_cap;
        ta.mp_create = ctx->trk_mp_create + (size_t)L.frame0 * K; ta.mp_xyz = ctx->trk_mp_xyz + (size_t)L.frame0 * K * 3;
        launch_track_update(ta, n, st, &ctx->launches);
        CU(copy_rows(h.mp_create, ta.mp_create, sizeof(int), KC, W, n, st));
        CU(copy_rows(h.mp_xyz, ta.mp_xyz, 3 * sizeof(float), KC, W, n, st));
    }
    CU(cudaMemcpyAsync(h.np_out, fb.params + L.frame0, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h.np_out + ctx->cfg.max_batch, fb.params + FT + L.frame0, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h.status, b.status + L.slot0, sizeof(int) * I, cudaMemcpyDeviceToHost, st));
    if (!fork) TRY(copy_extract(st));
    if (any_prev && !pose_only) {
        CU(copy_rows(h.bf_idx, fb.bf_idx + (size_t)L.frame0 * K, sizeof(int), KC, W, n, st));
        CU(copy_rows(h.bf_dist, fb.bf_dist + (size_t)L.frame0 * K, sizeof(int), KC, W, n, st));
        CU(copy_rows(h.bf_keep, fb.bf_keep + (size_t)L.frame0 * K, 1, KC, W, n, st));
        if (!ctx->cfg.skip_match_score) {
            CU(copy_rows(h.p1_best_idx, fb.p1_best_idx + (size_t)L.frame0 * R, sizeof(int), R, WP, n, st));
            CU(copy_rows(h.p1_best, fb.p1_best + (size_t)L.frame0 * R, sizeof(int), R, WP, n, st));
            CU(copy_rows(h.p1_second, fb.p1_second + (size_t)L.frame0 * R, sizeof(int), R, WP, n, st));
        }
        CU(copy_rows(h.p1_row_claimed, fb.p1_row_claimed + (size_t)L.frame0 * R, 1, R, WP, n, st));
        if (veto) CU(copy_rows(h.p1_row_bad, fb.p1_row_bad + (size_t)L.frame0 * R, 1, R, WP, n, st));
    }
    if (any_map && !pose_only) CU(cudaMemcpyAsync(h.p2_row_claimed, fb.p2_row_claimed + (size_t)L.frame0 * R, (size_t)n * R, cudaMemcpyDeviceToHost, st));
    CU(copy_rows(h.claim_row, fb.claim_row + (size_t)L.frame0 * K, sizeof(int), KC, W, n, st));
    return SVO_OK;
}


}  // namespace

// =========================================================================================
extern "C" {

void svo_default_config(svo_config *c)
{
    memset(c, 0, sizeof(*c));
    c->device = 0; c->width = 1241; c->height = 376;
    c->nfeatures = 2000; c->nlevels = 8; c->scale_factor = 1.2f; c->fast_threshold = 20;
    c->max_batch = 1; c->lanes = 1; c->max_rows = 5000; c->stream = nullptr; c->max_channels = 1; c->distribution = SVO_DIST_RETAIN_BEST;
}

const char *svo_version(void) { return "svo_b200 0.1 (sm_100a)"; }

const char *svo_last_error(const svo_ctx *ctx) { return ctx ? ctx->err : "null context"; }

void svo_destroy(svo_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    for (Lane &l : ctx->lanes) {
        for (LaneGraph &c : l.graphs) if (c.exec) cudaGraphExecDestroy(c.exec);
        if (l.done) cudaEventDestroy(l.done);
        for (int i = 0; i < N_EVENTS; ++i) if (l.ev[i]) cudaEventDestroy(l.ev[i]);
        if (l.own_stream && l.st) cudaStreamDestroy(l.st);
        if (l.side) cudaStreamDestroy(l.side);
        for (int i = 0; i < 6; ++i) if (l.fk[i]) cudaEventDestroy(l.fk[i]);
    }
    if (ctx->sync_st) cudaStreamDestroy(ctx->sync_st);
    for (void *p : ctx->dev_allocs) cudaFree(p);
    for (void *p : ctx->pinned_allocs) cudaFreeHost(p);
    delete ctx;
}

int svo_create(const svo_config *cfg, svo_ctx **out)
{
    if (!cfg || !out) return SVO_E_INVALID;
    *out = nullptr;
    svo_ctx *ctx = new svo_ctx();
    ctx->cfg = *cfg; ctx->launches = 0; ctx->profiling = false; ctx->err[0] = 0; ctx->out_flags = 0;
    { const char *e = getenv("SVO_B200_TC"); ctx->use_tc = !(e && e[0] == '0'); }
    { const char *e = getenv("SVO_B200_COMPACT_ROWS"); ctx->compact_rows = e ? atoi(e) : cfg->nfeatures + 32; if (ctx->compact_rows < 1) ctx->compact_rows = 1; }
    ctx->tc_prof = nullptr;
    ctx->trk_n = 0; ctx->trk_cap = 0; ctx->trk_window = 4; ctx->trk_d = nullptr; ctx->trk_scratch = nullptr;
    ctx->trk_mp_create = nullptr; ctx->trk_mp_xyz = nullptr; ctx->trk_img_last = nullptr;
    ctx->sync_st = nullptr; ctx->sync_have[0] = ctx->sync_have[1] = false;
    ctx->sync_stage = nullptr;
    if (ctx->cfg.max_channels == 0) ctx->cfg.max_channels = 1;
    *out = ctx;  // returned even on failure so the caller can read svo_last_error, then svo_destroy
    const svo_config &c = ctx->cfg;
    if (c.nlevels < 1 || c.nlevels > SVO_MAX_LEVELS || c.nfeatures < 1 || c.nfeatures > 60000 || c.max_batch < 1 ||
        c.lanes < 1 || c.max_rows < 1 || c.max_rows > 40000 || c.width < 64 || c.height < 64 || !(c.scale_factor > 1.f) ||
        (c.max_channels != 1 && c.max_channels != 3) || (c.distribution != SVO_DIST_RETAIN_BEST && c.distribution != SVO_DIST_OCTREE))
        return fail(ctx, SVO_E_INVALID, "invalid configuration");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, SVO_E_CUDA, "no CUDA device: %s (libsvo_b200 has no CPU fallback)", cudaGetErrorString(e));
    CU(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major != 10)
        return fail(ctx, SVO_E_CUDA, "device %s is sm_%d%d; this library carries sm_100a code only", prop.name, prop.major, prop.minor);
    TRY(build_geometry(ctx));
    const Geom &g = ctx->g;
    const int nbatch_frames = c.lanes * c.max_batch;
    ctx->nslots = 2 * nbatch_frames + 2;
    ctx->sync_slot0 = 2 * nbatch_frames;
    const size_t S = ctx->nslots;
    Bufs &b = ctx->b;
    TRY(dalloc(ctx, &b.pyr, S * g.pyr_bytes)); TRY(dalloc(ctx, &b.blur, S * g.pyr_bytes));
    TRY(dalloc(ctx, &b.bands, S * g.band_total)); TRY(dalloc(ctx, &b.bandcnt, S * g.bandcnt_total));
    TRY(dalloc(ctx, &b.ckey, S * g.cand_total)); TRY(dalloc(ctx, &b.cval, S * g.cand_total));
    TRY(dalloc(ctx, &b.lpos, S * g.cand_total)); TRY(dalloc(ctx, &b.rpos, S * g.cand_total));
    TRY(dalloc(ctx, &b.key2, S * g.total2)); TRY(dalloc(ctx, &b.val2, S * g.total2));
    TRY(dalloc(ctx, &b.cnt1, S * SVO_MAX_LEVELS)); TRY(dalloc(ctx, &b.kept1, S * SVO_MAX_LEVELS)); TRY(dalloc(ctx, &b.kept2, S * SVO_MAX_LEVELS));
    TRY(dalloc(ctx, &b.kp, S * g.kp_cap)); TRY(dalloc(ctx, &b.desc, S * g.kp_cap * 32));
    TRY(dalloc(ctx, &b.nkp, S)); TRY(dalloc(ctx, &b.status, S));
    uint32_t *rtab = nullptr;
    TRY(dalloc(ctx, &rtab, ctx->rtab_host.size()));
    CU(cudaMemcpy(rtab, ctx->rtab_host.data(), ctx->rtab_host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    b.rtab = rtab;
    {
        ResizeQuad *rq = nullptr;
        TRY(dalloc(ctx, &rq, ctx->rq_host.size()));
        CU(cudaMemcpy(rq, ctx->rq_host.data(), ctx->rq_host.size() * sizeof(ResizeQuad), cudaMemcpyHostToDevice));
        b.rqtab = reinterpret_cast<const uint32_t *>(rq);
    }
    {   // One-launch pyramid (pyramid.cu:k_pyramid).  Measured on B200 (profiles/r2_pyramid_fused_experiment.md): 128 us per
        // 64 images against 126 us for the seven per-level launches when run alone and a shorter single-frame critical path
        // (24 us instead of 50 us + six dependent launches), but 5 % LOWER batch throughput inside the three-lane pipeline,
        // where its 100 KB CTAs keep the other lanes' kernels off the SMs.  So the launch size decides: up to 4 stereo frames
        // (8 images) the fused kernel, above that the per-level kernels.  SVO_B200_PYRAMID_FUSED=1 / =0 forces one form.
        std::vector<PyrBand> bands;
        b.pyr_bands = nullptr;
        ctx->g.pyr_nbands = 0;
        const char *e = getenv("SVO_B200_PYRAMID_FUSED");
        ctx->g.pyr_fused_max = e ? (e[0] == '0' ? 0 : INT_MAX) : 8;
        if (ctx->g.pyr_fused_max > 0) {
            TRY(build_pyramid_bands(ctx, bands, 100 * 1024));
            if (ctx->g.pyr_nbands == 0) TRY(build_pyramid_bands(ctx, bands, 200 * 1024));
        }
        if (ctx->g.pyr_nbands) {
            PyrBand *d = nullptr;
            TRY(dalloc(ctx, &d, bands.size()));
            CU(cudaMemcpy(d, bands.data(), bands.size() * sizeof(PyrBand), cudaMemcpyHostToDevice));
            b.pyr_bands = d;
        } else ctx->g.pyr_nbands = 0;
    }
    CU(cudaMemset(b.pyr, 0, S * g.pyr_bytes)); CU(cudaMemset(b.blur, 0, S * g.pyr_bytes));
    CU(cudaMemset(b.nkp, 0, S * sizeof(int))); CU(cudaMemset(b.status, 0, S * sizeof(int)));
    CU(cudaMemset(b.kept1, 0, S * SVO_MAX_LEVELS * sizeof(int))); CU(cudaMemset(b.kept2, 0, S * SVO_MAX_LEVELS * sizeof(int)));
    if (g.kp_cap > greedy_max_cols())
        return fail(ctx, SVO_E_INVALID, "nfeatures %d exceeds the matcher's limit of %d keypoints per frame", c.nfeatures, greedy_max_cols());
    TRY(alloc_frames(ctx, ctx->fb, nbatch_frames, g.kp_cap, c.max_rows, false));
    const int sync_stride = g.kp_cap > c.max_rows ? g.kp_cap : c.max_rows;
    TRY(alloc_frames(ctx, ctx->sb, 1, sync_stride, sync_stride, true));
    {
        PoseBufs &p = ctx->pb;
        p.cap_prob = c.max_batch; p.cap_pts = (size_t)c.max_batch * g.kp_cap;
        const size_t NP = p.cap_prob, PTS = p.cap_pts, HY = NP * pose_max_iterations() * 4 * 12;
        TRY(dalloc(ctx, &p.d_hdr, NP)); TRY(halloc(ctx, &p.h_hdr, NP));
        TRY(dalloc(ctx, &p.d_p3, PTS * 3)); TRY(dalloc(ctx, &p.d_p2, PTS * 2));
        TRY(halloc(ctx, &p.h_p3, PTS * 3)); TRY(halloc(ctx, &p.h_p2, PTS * 2));
        TRY(dalloc(ctx, &p.d_mask, PTS));
        TRY(dalloc(ctx, &p.d_hyp, HY)); TRY(dalloc(ctx, &p.d_hypf, HY));
        TRY(dalloc(ctx, &p.d_info, NP * 4)); TRY(halloc(ctx, &p.h_info, NP * 4));
        TRY(dalloc(ctx, &p.d_pose, NP * 12)); TRY(halloc(ctx, &p.h_pose, NP * 12));
        TRY(dalloc(ctx, &p.d_stats, NP * 2)); TRY(halloc(ctx, &p.h_stats, NP * 2));
        TRY(dalloc(ctx, &p.d_T, NP * 16)); TRY(halloc(ctx, &p.h_T, NP * 16));
    }
    if (c.distribution == SVO_DIST_OCTREE) {
        for (int l = 0; l < g.nlevels; ++l) {
            if (g.lv[l].quota > octree_max_quota())
                return fail(ctx, SVO_E_CAPACITY, "octree distribution: level %d quota %d exceeds %d", l, g.lv[l].quota, octree_max_quota());
            if (g.lv[l].nbands && (float)(g.lv[l].x1 - g.lv[l].x0) / (float)(g.lv[l].y1 - g.lv[l].y0) >= 16.5f)
                return fail(ctx, SVO_E_INVALID, "octree distribution: level %d is wider than 16.5 x its height", l);
        }
    }
    if (setup_pyramid_attributes(g) != 0 || setup_fast_attributes(g) != 0 || setup_match_attributes() != 0 || setup_describe() != 0 || setup_select_attributes() != 0 ||
        setup_pose() != 0 || setup_octree_attributes() != 0)
        return fail(ctx, SVO_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (getenv("SVO_B200_TC_PROF")) TRY(dalloc(ctx, &ctx->tc_prof, 4 * 256));
    CU(cudaStreamCreateWithFlags(&ctx->sync_st, cudaStreamNonBlocking));
    ctx->stage_img_bytes = (size_t)g.H * ((size_t)g.W * c.max_channels + 256);
    if (c.max_channels == 3) {
        TRY(dalloc(ctx, &ctx->sync_stage, ctx->stage_img_bytes));
        TRY(dalloc(ctx, &ctx->sync_fp_d, 1)); TRY(halloc(ctx, &ctx->sync_fp_h, 1));
        TRY(dalloc(ctx, &ctx->sync_str_d, 2)); TRY(halloc(ctx, &ctx->sync_str_h, 2));
    }
    ctx->lanes.resize(c.lanes);
    for (int i = 0; i < c.lanes; ++i) {
        Lane &l = ctx->lanes[i];
        l.st = nullptr; l.own_stream = true; l.done = nullptr; l.side = nullptr; for (int k = 0; k < 6; ++k) l.fk[k] = nullptr; l.busy = false; l.nframes = 0;
        for (int k = 0; k < N_EVENTS; ++k) l.ev[k] = nullptr;
        l.slot0 = 2 * i * c.max_batch; l.frame0 = i * c.max_batch;
        if (i == 0 && c.stream) { l.st = (cudaStream_t)c.stream; l.own_stream = false; }
        else CU(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking));
        for (int k = 0; k < 6; ++k) CU(cudaEventCreateWithFlags(&l.fk[k], cudaEventDisableTiming));
        for (int k = 0; k < N_EVENTS; ++k) CU(cudaEventCreate(&l.ev[k]));
        const size_t B = c.max_batch, I = 2 * B, K = g.kp_cap, R = c.max_rows;
        HostArena &h = l.h;
        TRY(halloc(ctx, &h.kp, I * K)); TRY(halloc(ctx, &h.desc, I * K * 32));
        TRY(halloc(ctx, &h.nkp, I)); TRY(halloc(ctx, &h.status, I));
        TRY(halloc(ctx, &h.u_right, B * K)); TRY(halloc(ctx, &h.depth, B * K)); TRY(halloc(ctx, &h.n_stereo, B));
        TRY(halloc(ctx, &h.bf_idx, B * K)); TRY(halloc(ctx, &h.bf_dist, B * K)); TRY(halloc(ctx, &h.bf_keep, B * K));
        TRY(halloc(ctx, &h.p1_best_idx, B * R)); TRY(halloc(ctx, &h.p1_best, B * R)); TRY(halloc(ctx, &h.p1_second, B * R));
        TRY(halloc(ctx, &h.p1_row_claimed, B * R)); TRY(halloc(ctx, &h.p1_row_bad, B * R)); TRY(halloc(ctx, &h.p2_row_claimed, B * R));
        TRY(halloc(ctx, &h.claim_row, B * K));
        TRY(halloc(ctx, &h.params, 4 * B)); TRY(halloc(ctx, &h.np_out, 2 * B));
        h.mp_create = nullptr; h.mp_xyz = nullptr; l.tracked = false;
        l.stage_cap = I * ctx->stage_img_bytes + B * (R * 32 * 2 + R * 5 + R * 12 + R * 8 + R * 16 + SVO_MAX_BOXES * 16 + 72) + (12 * B + 8) * 512;
        TRY(dalloc(ctx, &l.d_stage, l.stage_cap));
        TRY(dalloc(ctx, &l.d_fp, B)); TRY(halloc(ctx, &l.h_fp, B));
        TRY(dalloc(ctx, &l.d_strides, I)); TRY(halloc(ctx, &l.h_strides, I));
    }
    CU(cudaDeviceSynchronize());
    return SVO_OK;
}

int svo_get_geometry(const svo_ctx *ctx, int *lw, int *lh, float *lscale, int *quota)
{
    if (!ctx) return SVO_E_INVALID;
    for (int l = 0; l < ctx->g.nlevels; ++l) {
        if (lw) lw[l] = ctx->g.lv[l].w;
        if (lh) lh[l] = ctx->g.lv[l].h;
        if (lscale) lscale[l] = ctx->g.lv[l].scale;
        if (quota) quota[l] = ctx->g.lv[l].quota;
    }
    return ctx->g.nlevels;
}

long long svo_launch_count(const svo_ctx *ctx) { return ctx ? ctx->launches : 0; }
int svo_set_outputs(svo_ctx *ctx, int flags)
{
    if (!ctx || (flags & ~(SVO_OUT_COMPACT | SVO_OUT_NO_RIGHT | SVO_OUT_POSE_INPUTS))) return fail(ctx, SVO_E_INVALID, "svo_set_outputs: unknown flag");
    if (flags & SVO_OUT_POSE_INPUTS) flags |= SVO_OUT_NO_RIGHT;
    ctx->out_flags = flags;
    return SVO_OK;
}
int svo_set_profiling(svo_ctx *ctx, int on) { if (!ctx) return SVO_E_INVALID; ctx->profiling = on != 0; return SVO_OK; }
void *svo_lane_stream(svo_ctx *ctx, int lane) { return (ctx && lane >= 0 && lane < (int)ctx->lanes.size()) ? (void *)ctx->lanes[lane].st : nullptr; }

void *svo_alloc_pinned(svo_ctx *ctx, size_t bytes)
{
    void *p = nullptr;
    if (!ctx || cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void svo_free_pinned(svo_ctx *ctx, void *p) { (void)ctx; if (p) cudaFreeHost(p); }
void *svo_alloc_device(svo_ctx *ctx, size_t bytes)
{
    void *p = nullptr;
    if (!ctx || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void svo_free_device(svo_ctx *ctx, void *p) { (void)ctx; if (p) cudaFree(p); }
int svo_copy_to_device(svo_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return SVO_E_INVALID;
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return SVO_OK;
}

// ------------------------------------------------------------------------------ sync API
namespace {
int extract_sync(svo_ctx *ctx, int cam, const uint8_t *img, int stride, int w, int h, int ch,
                 svo_keypoint *kp_out, uint8_t *desc_out, int cap);
}

int svo_extract(svo_ctx *ctx, int cam, const uint8_t *gray, int stride, int w, int h,
                svo_keypoint *kp_out, uint8_t *desc_out, int cap)
{
    return extract_sync(ctx, cam, gray, stride, w, h, 1, kp_out, desc_out, cap);
}

int svo_extract_bgr(svo_ctx *ctx, int cam, const uint8_t *bgr, int stride, int w, int h,
                    svo_keypoint *kp_out, uint8_t *desc_out, int cap)
{
    if (ctx && ctx->cfg.max_channels < 3) return fail(ctx, SVO_E_INVALID, "svo_extract_bgr: the context was created with max_channels = 1");
    return extract_sync(ctx, cam, bgr, stride, w, h, 3, kp_out, desc_out, cap);
}

namespace {
int extract_sync(svo_ctx *ctx, int cam, const uint8_t *gray, int stride, int w, int h, int ch,
                 svo_keypoint *kp_out, uint8_t *desc_out, int cap)
{
    if (!ctx || !gray || cam < 0 || cam > 1 || cap < 0) return fail(ctx, SVO_E_INVALID, "svo_extract: bad argument");
    const Geom &g = ctx->g;
    if (w != g.W || h != g.H || stride < w * ch || stride >= SVO_STRIDE_BGR)
        return fail(ctx, SVO_E_INVALID, "svo_extract: image is %dx%d (stride %d), context is %dx%d", w, h, stride, g.W, g.H);
    CU(cudaSetDevice(ctx->cfg.device));
    const int slot = ctx->sync_slot0 + cam;
    cudaStream_t st = ctx->sync_st;
    if (ch == 1) TRY(upload_image(ctx, slot, gray, stride, st));
    else {
        // colour: the rows land as they are (one contiguous copy, or read in place when already on the device) and
        // k_unpack converts while it repacks
        const size_t bytes = (size_t)stride * (h - 1) + (size_t)w * 3;
        const uint8_t *src = gray;
        if (!device_readable(gray)) {
            if (bytes > ctx->stage_img_bytes) return fail(ctx, SVO_E_CAPACITY, "svo_extract_bgr: rows %d bytes apart exceed the staging capacity", stride);
            CU(cudaMemcpyAsync(ctx->sync_stage, gray, bytes, cudaMemcpyHostToDevice, st));
            src = ctx->sync_stage;
        }
        memset(ctx->sync_fp_h, 0, sizeof(FramePtrs));
        ctx->sync_fp_h->left = src; ctx->sync_fp_h->right = src;
        ctx->sync_str_h[0] = ctx->sync_str_h[1] = stride | SVO_STRIDE_BGR;
        CU(cudaMemcpyAsync(ctx->sync_fp_d, ctx->sync_fp_h, sizeof(FramePtrs), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->sync_str_d, ctx->sync_str_h, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
        launch_unpack(ctx->b, g, slot, 1, ctx->sync_fp_d, ctx->sync_str_d, st, &ctx->launches);
    }
    enqueue_extract(ctx, slot, 1, st, nullptr);
    int hdr[2] = {0, 0};
    CU(cudaMemcpyAsync(&hdr[0], ctx->b.nkp + slot, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&hdr[1], ctx->b.status + slot, sizeof(int), cudaMemcpyDeviceToHost, st));
    // the outputs ride the same stream, sized by the caller's capacity (rows past the count are ignored): one sync
    const int mcap = cap < g.kp_cap ? cap : g.kp_cap;
    if (mcap > 0 && kp_out) CU(cudaMemcpyAsync(kp_out, ctx->b.kp + (size_t)slot * g.kp_cap, sizeof(svo_keypoint) * mcap, cudaMemcpyDeviceToHost, st));
    if (mcap > 0 && desc_out) CU(cudaMemcpyAsync(desc_out, ctx->b.desc + (size_t)slot * g.kp_cap * 32, 32 * (size_t)mcap, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (hdr[1] & SVO_STATUS_OVERFLOW) return fail(ctx, SVO_E_CAPACITY, "svo_extract: internal capacity exceeded (status %d, n %d)", hdr[1], hdr[0]);
    ctx->sync_have[cam] = true;
    return hdr[0];
}
}  // namespace

int svo_stereo_sparse(svo_ctx *ctx, float bf, float baseline, float *u_right, float *depth,
                      int32_t *match_r, int32_t *sad, int cap)
{
    if (!ctx || !u_right || !depth || cap < 0) return fail(ctx, SVO_E_INVALID, "svo_stereo_sparse: bad argument");
    if (!ctx->sync_have[0] || !ctx->sync_have[1]) return fail(ctx, SVO_E_INVALID, "svo_stereo_sparse: extract both cameras first");
    CU(cudaSetDevice(ctx->cfg.device));
    const Geom &g = ctx->g;
    FrameBufs &s = ctx->sb;
    cudaStream_t st = ctx->sync_st;
    float prm[2] = {bf, baseline};
    CU(cudaMemcpyAsync(s.params + 2, &prm[0], sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.params + 3, &prm[1], sizeof(float), cudaMemcpyHostToDevice, st));
    StereoArgs a;
    a.u_right = s.u_right; a.depth = s.depth; a.match_r = s.match_r; a.sad = s.sad; a.n_stereo = s.n_stereo;
    a.stride = s.col_stride;
    a.row_off = s.row_off; a.row_list = s.row_list; a.row_list_stride = s.row_list_stride;
    a.bf = reinterpret_cast<const float *>(s.params + 2); a.baseline = reinterpret_cast<const float *>(s.params + 3);
    launch_stereo(ctx->b, g, ctx->sync_slot0, 1, a, st, &ctx->launches);
    int n = 0;
    CU(cudaMemcpyAsync(&n, ctx->b.nkp + ctx->sync_slot0, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (n > g.kp_cap) n = g.kp_cap;
    const int m = n < cap ? n : cap;
    if (m > 0) {
        CU(cudaMemcpy(u_right, s.u_right, sizeof(float) * m, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(depth, s.depth, sizeof(float) * m, cudaMemcpyDeviceToHost));
        if (match_r) CU(cudaMemcpy(match_r, s.match_r, sizeof(int) * m, cudaMemcpyDeviceToHost));
        if (sad) CU(cudaMemcpy(sad, s.sad, sizeof(int) * m, cudaMemcpyDeviceToHost));
    }
    return n;
}

int svo_match_bf(svo_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                 int32_t *idx, int32_t *dist, uint8_t *keep)
{
    if (!ctx || nq < 0 || nt < 0 || (nq && !q) || (nt && !t)) return fail(ctx, SVO_E_INVALID, "svo_match_bf: bad argument");
    FrameBufs &s = ctx->sb;
    if (nq > s.col_stride || nt > s.row_stride) return fail(ctx, SVO_E_CAPACITY, "svo_match_bf: %d x %d exceeds capacity %d", nq, nt, s.col_stride);
    if (nq == 0) return 0;
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = ctx->sync_st;
    CU(cudaMemcpyAsync(s.cols, q, (size_t)nq * 32, cudaMemcpyDefault, st));
    if (nt) CU(cudaMemcpyAsync(s.prev, t, (size_t)nt * 32, cudaMemcpyDefault, st));
    BfArgs a;
    a.q = make_set(s.cols, nullptr, 0, s.col_stride, nq);
    a.t = make_set(s.prev, nullptr, 0, s.row_stride, nt);
    a.idx = s.bf_idx; a.dist = s.bf_dist; a.keep = s.bf_keep; a.min_dist = s.min_dist;
    launch_bf(a, 1, st, &ctx->launches);
    if (idx) CU(cudaMemcpyAsync(idx, s.bf_idx, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    if (dist) CU(cudaMemcpyAsync(dist, s.bf_dist, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    if (keep) CU(cudaMemcpyAsync(keep, s.bf_keep, (size_t)nq, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return nq;
}

int svo_match_greedy(svo_ctx *ctx, const uint8_t *rows, int M, const uint8_t *cur, int N, int mode,
                     const uint8_t *row_live, uint8_t *claimed, int32_t *claim_row, int row_base,
                     int32_t *best_idx, int32_t *best, int32_t *second, uint8_t *row_claimed,
                     const float *win_uvr, const float *cur_xy, const svo_veto *veto, uint8_t *row_bad)
{
    if (!ctx || M < 0 || N < 0 || (M && !rows) || (N && !cur) || !claimed || (mode != 0 && mode != 1) || (win_uvr && !cur_xy))
        return fail(ctx, SVO_E_INVALID, "svo_match_greedy: bad argument");
    FrameBufs &s = ctx->sb;
    if (M > s.row_stride || N > s.col_stride) return fail(ctx, SVO_E_CAPACITY, "svo_match_greedy: %d x %d exceeds capacity %d", M, N, s.col_stride);
    if (N > greedy_max_cols()) return fail(ctx, SVO_E_CAPACITY, "svo_match_greedy: at most %d columns", greedy_max_cols());
    if (veto && (veto->n_boxes > 256 || veto->n_boxes < 0)) return fail(ctx, SVO_E_CAPACITY, "svo_match_greedy: at most 256 boxes");
    if (M == 0) return 0;
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = ctx->sync_st;
    CU(cudaMemcpyAsync(s.map, rows, (size_t)M * 32, cudaMemcpyDefault, st));
    if (N) CU(cudaMemcpyAsync(s.cols, cur, (size_t)N * 32, cudaMemcpyDefault, st));
    if (N) CU(cudaMemcpyAsync(s.claimed, claimed, (size_t)N, cudaMemcpyDefault, st));
    if (N && claim_row) CU(cudaMemcpyAsync(s.claim_row, claim_row, sizeof(int) * N, cudaMemcpyDefault, st));
    if (row_live) CU(cudaMemcpyAsync(s.prev_live, row_live, (size_t)M, cudaMemcpyDefault, st));
    if (win_uvr) {
        CU(cudaMemcpyAsync(s.win, win_uvr, sizeof(float) * 3 * M, cudaMemcpyDefault, st));
    }
    if (cur_xy && N) CU(cudaMemcpyAsync(s.cur_xy, cur_xy, sizeof(float) * 2 * N, cudaMemcpyDefault, st));
    const bool use_veto = veto && veto->n_boxes > 0 && veto->F && veto->boxes && veto->row_xy && veto->cur_xy && mode == SVO_GREEDY_PASS1;
    if (use_veto) {
        CU(cudaMemcpyAsync(s.boxes, veto->boxes, sizeof(int) * 4 * veto->n_boxes, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(s.F, veto->F, sizeof(double) * 9, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(s.row_xy, veto->row_xy, sizeof(float) * 2 * M, cudaMemcpyDefault, st));
        if (N) CU(cudaMemcpyAsync(s.cur_xy, veto->cur_xy, sizeof(float) * 2 * N, cudaMemcpyDefault, st));
    }
    GreedyArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = make_set(s.map, nullptr, 0, s.row_stride, M);
    a.cols = make_set(s.cols, nullptr, 0, s.col_stride, N);
    a.mode = mode; a.row_base = row_base; a.row_base_arr = nullptr;
    a.row_live = row_live ? s.prev_live : nullptr;
    a.claimed = s.claimed; a.claim_row = s.claim_row; a.claim_time = s.claim_time;
    a.best_idx = s.p2_best_idx; a.best = s.p2_best; a.second = s.p2_second;
    a.row_claimed = s.p2_row_claimed; a.row_bad = s.p1_row_bad;
    a.shortlist = s.shortlist; a.shortlist_hi = s.shortlist_hi; a.short_cnt = s.short_cnt;
    a.res_rows = s.res_rows; a.res_off = s.res_off; a.res_want = s.res_want; a.res_perm = s.res_perm;
    a.win_uvr = win_uvr ? s.win : nullptr;
    a.cur_xy = (win_uvr || use_veto) ? s.cur_xy : nullptr;
    if (use_veto) { a.boxes = s.boxes; a.n_boxes = veto->n_boxes; a.F = s.F; a.row_xy = s.row_xy; }
    const bool want_scores = best_idx || best || second;
    launch_greedy(a, 1, want_scores, st, &ctx->launches);
    if (N) CU(cudaMemcpyAsync(claimed, s.claimed, (size_t)N, cudaMemcpyDeviceToHost, st));
    if (N && claim_row) CU(cudaMemcpyAsync(claim_row, s.claim_row, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
    if (best_idx) CU(cudaMemcpyAsync(best_idx, s.p2_best_idx, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
    if (best) CU(cudaMemcpyAsync(best, s.p2_best, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
    if (second) CU(cudaMemcpyAsync(second, s.p2_second, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
    if (row_claimed) CU(cudaMemcpyAsync(row_claimed, s.p2_row_claimed, (size_t)M, cudaMemcpyDeviceToHost, st));
    if (row_bad) CU(cudaMemcpyAsync(row_bad, s.p1_row_bad, (size_t)M, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return M;
}

namespace {
// Pack the problems of one pose call into the device arrays: host points go through the pinned staging
// (one H2D copy per array for the whole call), device points are copied in place.
int stage_pose_problems(svo_ctx *ctx, const char *who, const svo_pose_problem *pr, int np, int *max_n, size_t *total)
{
    PoseBufs &p = ctx->pb;
    if (np > p.cap_prob) return fail(ctx, SVO_E_CAPACITY, "%s: %d problems exceed max_batch %d", who, np, p.cap_prob);
    size_t off = 0; int mx = 0; bool any_host = false;
    for (int i = 0; i < np; ++i) {
        if (pr[i].n < 0 || (pr[i].n && (!pr[i].pts3d || !pr[i].pts2d))) return fail(ctx, SVO_E_INVALID, "%s: bad problem %d", who, i);
        if (off + pr[i].n > p.cap_pts) return fail(ctx, SVO_E_CAPACITY, "%s: more than %zu points in one call", who, p.cap_pts);
        PoseHdr &h = p.h_hdr[i];
        h.off = (int)off; h.n = pr[i].n; h.fx = pr[i].fx; h.fy = pr[i].fy; h.cx = pr[i].cx; h.cy = pr[i].cy;
        memcpy(h.Tcw, pr[i].Tcw, sizeof h.Tcw);
        off += pr[i].n; mx = std::max(mx, pr[i].n);
    }
    cudaStream_t st = ctx->sync_st;
    for (int i = 0; i < np; ++i) {
        const size_t o = p.h_hdr[i].off, n = pr[i].n;
        if (!n) continue;
        if (device_readable(pr[i].pts3d)) CU(cudaMemcpyAsync(p.d_p3 + 3 * o, pr[i].pts3d, 12 * n, cudaMemcpyDeviceToDevice, st));
        else { memcpy(p.h_p3 + 3 * o, pr[i].pts3d, 12 * n); any_host = true; }
        if (device_readable(pr[i].pts2d)) CU(cudaMemcpyAsync(p.d_p2 + 2 * o, pr[i].pts2d, 8 * n, cudaMemcpyDeviceToDevice, st));
        else memcpy(p.h_p2 + 2 * o, pr[i].pts2d, 8 * n), any_host = true;
    }
    if (any_host) {
        // host ranges of device-resident problems are left untouched on the device: copy per maximal host run
        for (int i = 0; i < np;) {
            if (!pr[i].n || device_readable(pr[i].pts3d)) { ++i; continue; }
            int j = i; size_t o = p.h_hdr[i].off, m = 0;
            while (j < np && (!pr[j].n || !device_readable(pr[j].pts3d))) { m += pr[j].n; ++j; }
            CU(cudaMemcpyAsync(p.d_p3 + 3 * o, p.h_p3 + 3 * o, 12 * m, cudaMemcpyHostToDevice, st));
            i = j;
        }
        for (int i = 0; i < np;) {
            if (!pr[i].n || device_readable(pr[i].pts2d)) { ++i; continue; }
            int j = i; size_t o = p.h_hdr[i].off, m = 0;
            while (j < np && (!pr[j].n || !device_readable(pr[j].pts2d))) { m += pr[j].n; ++j; }
            CU(cudaMemcpyAsync(p.d_p2 + 2 * o, p.h_p2 + 2 * o, 8 * m, cudaMemcpyHostToDevice, st));
            i = j;
        }
    }
    CU(cudaMemcpyAsync(p.d_hdr, p.h_hdr, sizeof(PoseHdr) * np, cudaMemcpyHostToDevice, st));
    *max_n = mx; *total = off;
    return SVO_OK;
}

PoseArgs pose_args(svo_ctx *ctx)
{
    PoseBufs &p = ctx->pb;
    PoseArgs a;
    a.hdr = p.d_hdr; a.p3 = p.d_p3; a.p2 = p.d_p2; a.hyp = p.d_hyp; a.hypf = p.d_hypf; a.mask = p.d_mask; a.info = p.d_info;
    a.pose_out = p.d_pose; a.ransac_iterations = 0; a.thr2 = 0; a.seed = 0; a.refine_iterations = 0;
    a.lm_iterations = 0; a.Tcw_out = p.d_T; a.lm_stats = p.d_stats;
    return a;
}
}  // namespace

int svo_pnp_ransac(svo_ctx *ctx, const svo_pose_problem *problems, int nproblems, int iterations,
                   float reproj_err, uint32_t seed, int refine_iters, svo_pnp_result *results, uint8_t *inliers)
{
    if (!ctx || nproblems < 0 || (nproblems && !problems) || !results || iterations < 1 || !(reproj_err > 0) || refine_iters < 0)
        return fail(ctx, SVO_E_INVALID, "svo_pnp_ransac: bad argument");
    if (iterations > pose_max_iterations()) return fail(ctx, SVO_E_CAPACITY, "svo_pnp_ransac: at most %d iterations", pose_max_iterations());
    if (nproblems == 0) return 0;
    CU(cudaSetDevice(ctx->cfg.device));
    int max_n = 0; size_t total = 0;
    TRY(stage_pose_problems(ctx, "svo_pnp_ransac", problems, nproblems, &max_n, &total));
    PoseBufs &p = ctx->pb;
    cudaStream_t st = ctx->sync_st;
    PoseArgs a = pose_args(ctx);
    a.ransac_iterations = iterations; a.thr2 = reproj_err * reproj_err; a.seed = seed; a.refine_iterations = refine_iters;
    launch_pnp_ransac(a, nproblems, max_n, st, &ctx->launches);
    CU(cudaMemcpyAsync(p.h_pose, p.d_pose, sizeof(double) * 12 * nproblems, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(p.h_info, p.d_info, sizeof(int) * 4 * nproblems, cudaMemcpyDeviceToHost, st));
    if (inliers && total) CU(cudaMemcpyAsync(inliers, p.d_mask, total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    for (int i = 0; i < nproblems; ++i) {
        svo_pnp_result &r = results[i];
        memcpy(r.R, p.h_pose + 12 * i, sizeof r.R); memcpy(r.t, p.h_pose + 12 * i + 9, sizeof r.t);
        r.n_inliers = p.h_info[4 * i]; r.best_iteration = p.h_info[4 * i + 1]; r.best_solution = p.h_info[4 * i + 2];
        r.n_hypotheses = p.h_info[4 * i + 3];
    }
    return nproblems;
}

int svo_pose_optimize(svo_ctx *ctx, const svo_pose_problem *problems, int nproblems, int iterations, float *Tcw_out, double *stats)
{
    if (!ctx || nproblems < 0 || (nproblems && !problems) || !Tcw_out || iterations < 0)
        return fail(ctx, SVO_E_INVALID, "svo_pose_optimize: bad argument");
    if (nproblems == 0) return 0;
    CU(cudaSetDevice(ctx->cfg.device));
    int max_n = 0; size_t total = 0;
    TRY(stage_pose_problems(ctx, "svo_pose_optimize", problems, nproblems, &max_n, &total));
    PoseBufs &p = ctx->pb;
    cudaStream_t st = ctx->sync_st;
    PoseArgs a = pose_args(ctx);
    a.lm_iterations = iterations;
    launch_pose_lm(a, nproblems, st, &ctx->launches);
    CU(cudaMemcpyAsync(p.h_T, p.d_T, sizeof(float) * 16 * nproblems, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(p.h_stats, p.d_stats, sizeof(double) * 2 * nproblems, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    memcpy(Tcw_out, p.h_T, sizeof(float) * 16 * nproblems);
    if (stats) memcpy(stats, p.h_stats, sizeof(double) * 2 * nproblems);
    return nproblems;
}

int svo_project_map(svo_ctx *ctx, const float *xyz, const int32_t *octave, int n, const float *Tcw,
                    float fx, float fy, float cx, float cy, float th, float *uvr_out)
{
    if (!ctx || n < 0 || (n && (!xyz || !uvr_out)) || !Tcw) return fail(ctx, SVO_E_INVALID, "svo_project_map: bad argument");
    if (!n) return 0;
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = ctx->sync_st;
    float T[12];
    CU(cudaMemcpy(T, Tcw, sizeof T, cudaMemcpyDefault));
    float *d_xyz = nullptr, *d_out = nullptr; int *d_oct = nullptr;
    CU(cudaMallocAsync((void **)&d_xyz, 3 * sizeof(float) * (size_t)n, st));
    CU(cudaMallocAsync((void **)&d_out, 3 * sizeof(float) * (size_t)n, st));
    CU(cudaMemcpyAsync(d_xyz, xyz, 3 * sizeof(float) * (size_t)n, cudaMemcpyDefault, st));
    if (octave) {
        CU(cudaMallocAsync((void **)&d_oct, sizeof(int) * (size_t)n, st));
        CU(cudaMemcpyAsync(d_oct, octave, sizeof(int) * (size_t)n, cudaMemcpyDefault, st));
    }
    float ls[SVO_MAX_LEVELS];
    for (int l = 0; l < ctx->g.nlevels; ++l) ls[l] = ctx->g.lv[l].scale;
    launch_project(d_xyz, d_oct, n, T, fx, fy, cx, cy, ctx->g.W, ctx->g.H, th, ls, ctx->g.nlevels, d_out, st, &ctx->launches);
    CU(cudaMemcpyAsync(uvr_out, d_out, 3 * sizeof(float) * (size_t)n, cudaMemcpyDefault, st));
    CU(cudaFreeAsync(d_xyz, st)); CU(cudaFreeAsync(d_out, st));
    if (d_oct) CU(cudaFreeAsync(d_oct, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return n;
}

int svo_disp2depth(svo_ctx *ctx, const float *disp, float *depth, size_t n, float bf)
{
    if (!ctx || (n && (!disp || !depth))) return fail(ctx, SVO_E_INVALID, "svo_disp2depth: bad argument");
    if (!n) return SVO_OK;
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = ctx->sync_st;
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMallocAsync((void **)&d_in, n * sizeof(float), st));
    CU(cudaMallocAsync((void **)&d_out, n * sizeof(float), st));
    CU(cudaMemcpyAsync(d_in, disp, n * sizeof(float), cudaMemcpyDefault, st));
    launch_disp2depth(d_in, d_out, n, bf, st, &ctx->launches);
    CU(cudaMemcpyAsync(depth, d_out, n * sizeof(float), cudaMemcpyDefault, st));
    CU(cudaFreeAsync(d_in, st));
    CU(cudaFreeAsync(d_out, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return SVO_OK;
}

// ----------------------------------------------------------------------------- device-resident tracker state
int svo_track_kp_capacity(const svo_ctx *ctx) { return ctx ? ctx->g.kp_cap : SVO_E_INVALID; }

int svo_track_create(svo_ctx *ctx, int n_sequences, int map_capacity, int window)
{
    if (!ctx || n_sequences < 1 || map_capacity < 1 || map_capacity > ctx->cfg.max_rows || window < 1)
        return fail(ctx, SVO_E_INVALID, "svo_track_create: bad argument (map_capacity must lie in 1..max_rows = %d)", ctx ? ctx->cfg.max_rows : 0);
    if (ctx->trk_n) return fail(ctx, SVO_E_INVALID, "svo_track_create: the context already has tracker states");
    CU(cudaSetDevice(ctx->cfg.device));
    const size_t S = 2 * (size_t)n_sequences, K = ctx->g.kp_cap, C = map_capacity, FT = ctx->fb.nframes;
    uint8_t *last_desc, *prev_desc, *prev_live, *map_desc;
    int *prev_map_row, *prev_create, *n_prev, *map_create, *map_link, *n_map;
    float *prev_xyz, *prev_xy, *map_xyz;
    TRY(dalloc(ctx, &last_desc, S * K * 32)); TRY(dalloc(ctx, &prev_desc, S * K * 32)); TRY(dalloc(ctx, &prev_live, S * K));
    TRY(dalloc(ctx, &prev_map_row, S * K)); TRY(dalloc(ctx, &prev_create, S * K)); TRY(dalloc(ctx, &prev_xyz, S * K * 3));
    TRY(dalloc(ctx, &prev_xy, S * K * 2)); TRY(dalloc(ctx, &n_prev, S));
    TRY(dalloc(ctx, &map_desc, S * C * 32)); TRY(dalloc(ctx, &map_create, S * C)); TRY(dalloc(ctx, &map_link, S * C));
    TRY(dalloc(ctx, &map_xyz, S * C * 3)); TRY(dalloc(ctx, &n_map, S));
    ctx->trk_h.resize(S);
    for (size_t i = 0; i < S; ++i) {
        TrackState &t = ctx->trk_h[i];
        t.last_desc = last_desc + i * K * 32; t.prev_desc = prev_desc + i * K * 32; t.prev_live = prev_live + i * K;
        t.prev_map_row = prev_map_row + i * K; t.prev_create = prev_create + i * K; t.prev_xyz = prev_xyz + i * K * 3;
        t.prev_xy = prev_xy + i * K * 2; t.n_prev = n_prev + i;
        t.map_desc = map_desc + i * C * 32; t.map_create = map_create + i * C; t.map_link = map_link + i * C;
        t.map_xyz = map_xyz + i * C * 3; t.n_map = n_map + i;
    }
    TRY(dalloc(ctx, &ctx->trk_d, S));
    CU(cudaMemcpy(ctx->trk_d, ctx->trk_h.data(), sizeof(TrackState) * S, cudaMemcpyHostToDevice));
    TRY(dalloc(ctx, &ctx->trk_scratch, FT * C));
    TRY(dalloc(ctx, &ctx->trk_mp_create, FT * K)); TRY(dalloc(ctx, &ctx->trk_mp_xyz, FT * K * 3));
    TRY(dalloc(ctx, &ctx->trk_img_last, FT * ctx->fb.img_col_stride));
    for (Lane &l : ctx->lanes) {
        const size_t B = ctx->cfg.max_batch;
        TRY(halloc(ctx, &l.h.mp_create, B * K)); TRY(halloc(ctx, &l.h.mp_xyz, B * K * 3));
    }
    ctx->trk_parity.assign(n_sequences, 0);
    ctx->trk_last_lane.assign(n_sequences, -1);
    ctx->trk_n = n_sequences; ctx->trk_cap = map_capacity; ctx->trk_window = window;
    return SVO_OK;
}

int svo_track_reset(svo_ctx *ctx, int seq, const uint8_t *ballast, int n_ballast)
{
    if (!ctx || seq < 0 || seq >= ctx->trk_n || n_ballast < 0 || n_ballast > ctx->trk_cap || (n_ballast && !ballast))
        return fail(ctx, SVO_E_INVALID, "svo_track_reset: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());                         // no batch may still be reading or advancing the state
    const TrackState &t = ctx->trk_h[2 * (size_t)seq + ctx->trk_parity[seq]];
    const int zero = 0;
    CU(cudaMemcpy(t.n_prev, &zero, sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t.n_map, &n_ballast, sizeof(int), cudaMemcpyHostToDevice));
    if (n_ballast) {
        std::vector<int> v((size_t)n_ballast, INT_MAX);
        CU(cudaMemcpy(t.map_desc, ballast, (size_t)n_ballast * 32, cudaMemcpyDefault));
        CU(cudaMemcpy(t.map_create, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
        CU(cudaMemset(t.map_link, 0xff, sizeof(int) * (size_t)n_ballast));
        CU(cudaMemset(t.map_xyz, 0, sizeof(float) * 3 * (size_t)n_ballast));
    }
    ctx->trk_last_lane[seq] = -1;
    return SVO_OK;
}

int svo_track_state(svo_ctx *ctx, int seq, svo_track_view *v)
{
    if (!ctx || !v || seq < 0 || seq >= ctx->trk_n) return fail(ctx, SVO_E_INVALID, "svo_track_state: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    const TrackState &t = ctx->trk_h[2 * (size_t)seq + (ctx->trk_parity[seq] ^ (v->previous ? 1 : 0))];
    CU(cudaMemcpy(&v->n_prev, t.n_prev, sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&v->n_map, t.n_map, sizeof(int), cudaMemcpyDeviceToHost));
    const size_t np = (size_t)std::min(v->n_prev, ctx->g.kp_cap), nm = (size_t)std::min(v->n_map, ctx->trk_cap);
    auto get = [&](void *dst, const void *src, size_t bytes) { return (dst && bytes) ? cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) : cudaSuccess; };
    CU(get(v->last_desc, t.last_desc, np * 32)); CU(get(v->prev_desc, t.prev_desc, np * 32)); CU(get(v->prev_live, t.prev_live, np));
    CU(get(v->prev_map_row, t.prev_map_row, np * 4)); CU(get(v->prev_create, t.prev_create, np * 4));
    CU(get(v->prev_xyz, t.prev_xyz, np * 12)); CU(get(v->prev_xy, t.prev_xy, np * 8));
    CU(get(v->map_desc, t.map_desc, nm * 32)); CU(get(v->map_create, t.map_create, nm * 4)); CU(get(v->map_link, t.map_link, nm * 4));
    CU(get(v->map_xyz, t.map_xyz, nm * 12));
    return SVO_OK;
}

// ----------------------------------------------------------------------------- batch API
int svo_batch_submit(svo_ctx *ctx, int lane_i, const svo_frame_in *frames, int n)
{
    if (!ctx || lane_i < 0 || lane_i >= (int)ctx->lanes.size() || !frames || n < 1 || n > ctx->cfg.max_batch)
        return fail(ctx, SVO_E_INVALID, "svo_batch_submit: bad argument");
    Lane &L = ctx->lanes[lane_i];
    if (L.busy) return fail(ctx, SVO_E_INVALID, "svo_batch_submit: lane %d still has a batch in flight", lane_i);
    CU(cudaSetDevice(ctx->cfg.device));
    const Geom &g = ctx->g;
    const Bufs &b = ctx->b;
    FrameBufs &fb = ctx->fb;
    cudaStream_t st = L.st;
    const int R = fb.row_stride, K = fb.col_stride, B = ctx->cfg.max_batch;
    cudaEvent_t *ev = ctx->profiling ? L.ev : nullptr;
    bool any_prev = false, any_map = false;
    int n_win = 0, n_mapped = 0;
    bool veto = false, tracked = false;
    for (int i = 0; i < n; ++i) {
        const svo_frame_in &f = frames[i];
        if (f.track_seq) {
            if (f.track_seq < 0 || f.track_seq > ctx->trk_n) return fail(ctx, SVO_E_INVALID, "svo_batch_submit: frame %d: track_seq %d of %d (svo_track_create)", i, f.track_seq, ctx->trk_n);
            for (int k = 0; k < i; ++k)
                if (frames[k].track_seq == f.track_seq) return fail(ctx, SVO_E_INVALID, "svo_batch_submit: sequence %d appears twice in one batch", f.track_seq - 1);
            tracked = true;
        }
        if (f.n_boxes < 0 || f.n_boxes > SVO_MAX_BOXES || (f.n_boxes && !f.boxes))
            return fail(ctx, SVO_E_INVALID, "svo_batch_submit: frame %d: n_boxes %d (at most %d)", i, f.n_boxes, SVO_MAX_BOXES);
        if (f.F && ((uintptr_t)f.F & 7)) return fail(ctx, SVO_E_INVALID, "svo_batch_submit: frame %d: F must be 8-byte aligned", i);
        veto |= f.n_boxes > 0 && f.F && (f.track_seq ? true : (f.n_prev > 0 && f.prev_xy != nullptr));
        if (f.n_map > 0 && !f.track_seq) { ++n_mapped; if (f.map_win_uvr || (f.map_xyz && f.Tcw_pred)) ++n_win; }
        const int ch = f.channels == 3 ? 3 : 1;
        if (!f.left || !f.right || f.stride < g.W * ch || f.stride >= SVO_STRIDE_BGR || !(f.baseline > 0.f) ||
            (!f.track_seq && (f.n_prev < 0 || f.n_map < 0 || f.n_prev > R || f.n_map > R || (f.n_prev && !f.prev_desc) || (f.n_map && !f.map_desc))) ||
            (f.channels != 0 && f.channels != 1 && f.channels != 3) || ch > ctx->cfg.max_channels)
            return fail(ctx, SVO_E_INVALID, "svo_batch_submit: frame %d has bad inputs%s", i,
                        ch > ctx->cfg.max_channels ? " (BGR input needs svo_config.max_channels = 3)" : "");
        any_prev |= f.n_prev > 0 || f.track_seq; any_map |= f.n_map > 0 || f.track_seq;
    }
    if (tracked && n_win) return fail(ctx, SVO_E_INVALID, "svo_batch_submit: projection windows and tracked frames cannot share a batch");
    if (n_win != 0 && n_win != n_mapped)
        return fail(ctx, SVO_E_INVALID, "svo_batch_submit: map_win_uvr (or map_xyz + Tcw_pred) must be given for every frame with a map, or for none");
    const bool windowed = n_win > 0;
    L.in.assign(frames, frames + n);
    L.nframes = n; L.veto = veto; L.tracked = tracked;
    L.out_flags = ctx->out_flags;
    L.out_w = g.kp_cap; L.out_wp = R;
    if (ctx->out_flags & SVO_OUT_COMPACT) {
        L.out_w = std::min(g.kp_cap, ctx->compact_rows);
        int mp = L.out_w;                                  // tracked frames: the last frame's keypoints are the pass-1 rows
        for (int i = 0; i < n; ++i) if (!frames[i].track_seq) mp = std::max(mp, frames[i].n_prev);
        L.out_wp = mp > L.out_w ? R : L.out_w;
    }
    if (ev) cudaEventRecord(ev[0], st);
    // ---- inputs: device-resident buffers are read in place; host buffers are gathered into the lane's
    // landing zone with one H2D copy per maximal run of adjacent source ranges (a strided 2-D copy of
    // 1241-byte rows, or 32 separate 466 kB copies, run at a fraction of PCIe speed); kernels reach every
    // input through the per-frame pointer table.
    int *hp = L.h.params;
    std::vector<Seg> segs;
    segs.reserve(6 * (size_t)n);
    auto place = [&](const void *p, size_t bytes, const void **field, bool need16) {
        *field = nullptr;
        if (!p || !bytes) return;
        if (device_readable(p) && (!need16 || ((uintptr_t)p & 15) == 0)) { *field = p; return; }
        segs.push_back(Seg{(const uint8_t *)p, bytes, field, need16});
    };
    for (int i = 0; i < n; ++i) {
        const svo_frame_in &f = frames[i];
        FramePtrs &P = L.h_fp[i];
        const int ch = f.channels == 3 ? 3 : 1;
        const size_t img_bytes = (size_t)f.stride * (g.H - 1) + (size_t)g.W * ch;
        const uint8_t *srcs[2] = {f.left, f.right};
        const void **fld[2] = {(const void **)&P.left, (const void **)&P.right};
        for (int s = 0; s < 2; ++s) {
            if (img_bytes <= ctx->stage_img_bytes || device_readable(srcs[s])) {
                place(srcs[s], img_bytes, fld[s], false);
                L.h_strides[2 * i + s] = f.stride | (ch == 3 ? SVO_STRIDE_BGR : 0);
            } else if (ch == 3) {
                return fail(ctx, SVO_E_CAPACITY, "svo_batch_submit: frame %d: BGR rows %d bytes apart exceed the staging capacity", i, f.stride);
            } else {   // oversized stride: 2-D copy straight into the level-0 slot
                *fld[s] = nullptr;
                TRY(upload_image(ctx, L.slot0 + 2 * i + s, srcs[s], f.stride, st));
                L.h_strides[2 * i + s] = 0;
            }
        }
        P.trk_in = P.trk_out = nullptr; P.frame_id = f.frame_id;
        P.fx = f.fx; P.fy = f.fy; P.cx = f.cx; P.cy = f.cy; P.proj_th = f.proj_th;
        if (f.track_seq) {
            // the sequence's state where it lies: this frame reads the current copy and its update writes the other one
            const int sq = f.track_seq - 1, cur = ctx->trk_parity[sq];
            const TrackState &t = ctx->trk_h[2 * (size_t)sq + cur];
            P.prev = t.prev_desc; P.last = t.last_desc; P.prev_live = t.prev_live; P.map = t.map_desc; P.map_prev_row = t.map_link;
            P.map_win = nullptr; P.map_xyz = nullptr; P.map_octave = nullptr;
            P.trk_in = ctx->trk_d + 2 * (size_t)sq + cur; P.trk_out = ctx->trk_d + 2 * (size_t)sq + (cur ^ 1);
            place(f.n_boxes ? f.boxes : nullptr, 4 * sizeof(int) * (size_t)f.n_boxes, (const void **)&P.boxes, false);   // createmappoint reads them too
            place((f.n_boxes && f.F) ? f.F : nullptr, 9 * sizeof(double), (const void **)&P.F, false);
            P.prev_xy = t.prev_xy;
            P.n_boxes = f.n_boxes;
            hp[i] = 0; hp[B + i] = 0;                     // k_track_load fills the counts in from the state
        } else {
            place(f.n_prev ? f.prev_desc : nullptr, (size_t)f.n_prev * 32, (const void **)&P.prev, true);
            place(f.n_prev ? f.prev_live : nullptr, (size_t)f.n_prev, (const void **)&P.prev_live, false);
            place(f.n_map ? f.map_desc : nullptr, (size_t)f.n_map * 32, (const void **)&P.map, true);
            place((f.n_map && f.n_prev) ? f.map_prev_row : nullptr, sizeof(int) * (size_t)f.n_map, (const void **)&P.map_prev_row, true);
            place((f.n_map && windowed) ? f.map_win_uvr : nullptr, 3 * sizeof(float) * (size_t)f.n_map, (const void **)&P.map_win, false);
            const bool proj = f.n_map && windowed && !f.map_win_uvr && f.map_xyz && f.Tcw_pred;
            place(proj ? f.map_xyz : nullptr, 3 * sizeof(float) * (size_t)f.n_map, (const void **)&P.map_xyz, false);
            place(proj ? f.map_octave : nullptr, sizeof(int) * (size_t)f.n_map, (const void **)&P.map_octave, false);
            if (proj) {
                if (device_readable(f.Tcw_pred)) return fail(ctx, SVO_E_INVALID, "svo_batch_submit: frame %d: Tcw_pred must be a host pointer", i);
                memcpy(P.Tcw, f.Tcw_pred, sizeof(P.Tcw));
            }
            const bool fv = f.n_prev > 0 && f.n_boxes > 0 && f.F && f.prev_xy;
            place(fv ? f.boxes : nullptr, 4 * sizeof(int) * (size_t)f.n_boxes, (const void **)&P.boxes, false);
            place(fv ? f.F : nullptr, 9 * sizeof(double), (const void **)&P.F, false);
            place(fv ? f.prev_xy : nullptr, 2 * sizeof(float) * (size_t)f.n_prev, (const void **)&P.prev_xy, false);
            P.n_boxes = fv ? f.n_boxes : 0;
            hp[i] = f.n_prev; hp[B + i] = f.n_map;
        }
        memcpy(&hp[2 * B + i], &f.bf, 4); memcpy(&hp[3 * B + i], &f.baseline, 4);
    }
    std::sort(segs.begin(), segs.end(), [](const Seg &x, const Seg &y) { return x.src < y.src; });
    size_t cursor = 0;
    for (size_t i = 0; i < segs.size();) {
        // a run: source ranges that touch or overlap (never bridges a gap: only the caller's bytes are read)
        const uint8_t *run_src = segs[i].src, *run_end = run_src + segs[i].bytes;
        size_t j = i + 1;
        const bool solo = segs[i].need16 && ((uintptr_t)run_src & 15);      // misaligned descriptors: realigned, alone
        while (!solo && j < segs.size() && segs[j].src <= run_end && !(segs[j].need16 && ((uintptr_t)segs[j].src & 15))) {
            run_end = std::max(run_end, segs[j].src + segs[j].bytes);
            ++j;
        }
        const size_t dst = ((cursor + 255) & ~(size_t)255) + (solo ? 0 : ((uintptr_t)run_src & 255));
        const size_t len = (size_t)(run_end - run_src);
        if (dst + len > L.stage_cap) return fail(ctx, SVO_E_CAPACITY, "svo_batch_submit: inputs exceed the staging capacity");
        CU(cudaMemcpyAsync(L.d_stage + dst, run_src, len, cudaMemcpyDefault, st));
        for (size_t k = i; k < j; ++k) *segs[k].field = L.d_stage + dst + (segs[k].src - run_src);
        cursor = dst + len;
        i = j;
    }
    for (int i = 0; i < n; ++i)
        if (!frames[i].track_seq) L.h_fp[i].last = L.h_fp[i].prev;      // one set serves the BF matcher and pass 1
    CU(cudaMemcpyAsync(L.d_fp, L.h_fp, sizeof(FramePtrs) * n, cudaMemcpyHostToDevice, st));
    // params live as [4][nframes_total] on the device; this lane owns columns frame0..frame0+B
    const int FT = fb.nframes;
    for (int k = 0; k < 4; ++k)
        CU(cudaMemcpyAsync(fb.params + (size_t)k * FT + L.frame0, hp + (size_t)k * B, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(L.d_strides, L.h_strides, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, st));
    bool fused = any_prev;
    for (int i = 0; i < n; ++i) fused = fused && frames[i].n_prev <= K;   // the distance matrix holds kp_cap rows
    // A tracked frame's matching must see the state its sequence's previous frame left: when that frame ran on another
    // lane, this lane waits for that lane's batch.  Extraction does not depend on the state, so the batch is split: its
    // first half (repack + extraction) is enqueued before the wait and overlaps the other lane's matching.
    std::vector<int> wait_lanes;
    for (int i = 0; i < n; ++i)
        if (frames[i].track_seq) {
            const int ll = ctx->trk_last_lane[frames[i].track_seq - 1];
            if (ll >= 0 && ll != lane_i && std::find(wait_lanes.begin(), wait_lanes.end(), ll) == wait_lanes.end()) wait_lanes.push_back(ll);
        }
    const bool split = !wait_lanes.empty();
    // ---- all kernels, memsets and result copies of the batch.  Their arguments depend only on (lane, n, which
    // stages run): every per-frame input is reached through device tables filled above.  Outside profiling runs
    // the sequence is therefore captured once into a CUDA graph and replayed (one launch instead of ~45 calls).
    auto run = [&](int phase) -> int {
        if (ev) return enqueue_compute(ctx, L, n, any_prev, any_map, fused, windowed, veto, tracked, phase, ev);
        const int key = n | (any_prev ? 1 << 16 : 0) | (any_map ? 1 << 17 : 0) | (fused ? 1 << 18 : 0) | (windowed ? 1 << 19 : 0) | (veto ? 1 << 20 : 0) |
                        (tracked ? 1 << 21 : 0) | (phase << 22) | (L.out_flags << 24) | (L.out_wp == R ? 1 << 28 : 0);
        LaneGraph *lg = nullptr;
        for (LaneGraph &c : L.graphs) if (c.key == key) lg = &c;
        if (!lg) {
            const long long before = ctx->launches;
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue_compute(ctx, L, n, any_prev, any_map, fused, windowed, veto, tracked, phase, nullptr);
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != SVO_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(ctx, SVO_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            LaneGraph ng; ng.key = key; ng.exec = nullptr; ng.launches = ctx->launches - before;
            ctx->launches = before;
            const cudaError_t ie = cudaGraphInstantiate(&ng.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return fail(ctx, SVO_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
            L.graphs.push_back(ng);
            lg = &L.graphs.back();
        }
        CU(cudaGraphLaunch(lg->exec, st));
        ctx->launches += lg->launches;
        return SVO_OK;
    };
    if (split) {
        TRY(run(1));
        for (int ll : wait_lanes) CU(cudaStreamWaitEvent(st, ctx->lanes[ll].done, 0));
        TRY(run(2));
    } else TRY(run(0));
    if (ev) cudaEventRecord(ev[11], st);
    CU(cudaEventRecord(L.done, st));
    CU(cudaGetLastError());
    for (int i = 0; i < n; ++i)
        if (frames[i].track_seq) { ctx->trk_parity[frames[i].track_seq - 1] ^= 1; ctx->trk_last_lane[frames[i].track_seq - 1] = lane_i; }
    L.busy = true;
    return SVO_OK;
}

int svo_batch_wait(svo_ctx *ctx, int lane_i)
{
    if (!ctx || lane_i < 0 || lane_i >= (int)ctx->lanes.size()) return fail(ctx, SVO_E_INVALID, "svo_batch_wait: bad lane");
    Lane &L = ctx->lanes[lane_i];
    if (!L.busy) return SVO_OK;
    CU(cudaEventSynchronize(L.done));
    L.busy = false;
    CU(cudaGetLastError());
    if (L.out_w < ctx->g.kp_cap || L.out_wp < ctx->cfg.max_rows) {
        // compact result copies: a frame with more keypoints (or pass-1 rows) than the copies carried gets the rest now
        const Geom &g = ctx->g; const Bufs &b = ctx->b; FrameBufs &fb = ctx->fb; HostArena &h = L.h;
        const size_t KC = g.kp_cap, R = ctx->cfg.max_rows, W = L.out_w, WP = L.out_wp;
        const bool no_right = (L.out_flags & SVO_OUT_NO_RIGHT) != 0, pose_only = (L.out_flags & SVO_OUT_POSE_INPUTS) != 0;
        bool any = false;
        auto tail = [&](void *hbase, const void *dbase, size_t elem, size_t row0, size_t rows) {
            any = true;
            return cudaMemcpyAsync((char *)hbase + elem * row0, (const char *)dbase + elem * row0, elem * rows, cudaMemcpyDeviceToHost, L.st);
        };
        for (int i = 0; i < L.nframes; ++i) {
            const size_t nl = std::min((size_t)std::max(h.nkp[2 * i], 0), KC), nr = std::min((size_t)std::max(h.nkp[2 * i + 1], 0), KC);
            const size_t fo = (size_t)L.frame0 + i, so = (size_t)L.slot0 + 2 * i;
            if (nl > W) {
                const size_t m = nl - W;
                CU(tail(h.kp + 2 * i * KC, b.kp + so * KC, sizeof(svo_keypoint), W, m));
                CU(tail(h.depth + i * KC, fb.depth + fo * KC, 4, W, m)); CU(tail(h.claim_row + i * KC, fb.claim_row + fo * KC, 4, W, m));
                if (!pose_only) {
                    CU(tail(h.desc + 2 * i * KC * 32, b.desc + so * KC * 32, 32, W, m)); CU(tail(h.u_right + i * KC, fb.u_right + fo * KC, 4, W, m));
                    CU(tail(h.bf_idx + i * KC, fb.bf_idx + fo * KC, 4, W, m)); CU(tail(h.bf_dist + i * KC, fb.bf_dist + fo * KC, 4, W, m));
                    CU(tail(h.bf_keep + i * KC, fb.bf_keep + fo * KC, 1, W, m));
                }
                if (L.tracked && h.mp_create) {
                    CU(tail(h.mp_create + i * KC, ctx->trk_mp_create + fo * KC, 4, W, m)); CU(tail(h.mp_xyz + i * KC * 3, ctx->trk_mp_xyz + fo * KC * 3, 12, W, m));
                }
            }
            if (!no_right && nr > W) {
                const size_t m = nr - W;
                CU(tail(h.kp + (2 * i + 1) * KC, b.kp + (so + 1) * KC, sizeof(svo_keypoint), W, m));
                CU(tail(h.desc + (2 * i + 1) * KC * 32, b.desc + (so + 1) * KC * 32, 32, W, m));
            }
            const size_t np = std::min((size_t)std::max(h.np_out[i], 0), R);
            if (np > WP && !pose_only) {
                const size_t m = np - WP;
                CU(tail(h.p1_best_idx + i * R, fb.p1_best_idx + fo * R, 4, WP, m)); CU(tail(h.p1_best + i * R, fb.p1_best + fo * R, 4, WP, m));
                CU(tail(h.p1_second + i * R, fb.p1_second + fo * R, 4, WP, m)); CU(tail(h.p1_row_claimed + i * R, fb.p1_row_claimed + fo * R, 1, WP, m));
                CU(tail(h.p1_row_bad + i * R, fb.p1_row_bad + fo * R, 1, WP, m));
            }
        }
        if (any) CU(cudaStreamSynchronize(L.st));
    }
    return SVO_OK;
}

int svo_batch_result(svo_ctx *ctx, int lane_i, int i, svo_frame_out *o)
{
    if (!ctx || !o || lane_i < 0 || lane_i >= (int)ctx->lanes.size()) return fail(ctx, SVO_E_INVALID, "svo_batch_result: bad argument");
    Lane &L = ctx->lanes[lane_i];
    if (L.busy) return fail(ctx, SVO_E_INVALID, "svo_batch_result: call svo_batch_wait first");
    if (i < 0 || i >= L.nframes) return fail(ctx, SVO_E_INVALID, "svo_batch_result: frame %d of %d", i, L.nframes);
    const size_t K = ctx->g.kp_cap, R = ctx->cfg.max_rows;
    const HostArena &h = L.h;
    const svo_frame_in &in = L.in[i];
    memset(o, 0, sizeof(*o));
    const int st = h.status[2 * i] | h.status[2 * i + 1];
    o->n_left = h.nkp[2 * i]; o->n_right = h.nkp[2 * i + 1];
    o->status = (st & SVO_STATUS_OVERFLOW) ? SVO_E_CAPACITY : SVO_OK;
    if (o->n_left > (int)K) { o->n_left = (int)K; o->status = SVO_E_CAPACITY; }
    if (o->n_right > (int)K) { o->n_right = (int)K; o->status = SVO_E_CAPACITY; }
    o->n_stereo = h.n_stereo[i];
    o->kp_left = h.kp + (2 * (size_t)i) * K; o->desc_left = h.desc + (2 * (size_t)i) * K * 32;
    if (!(L.out_flags & SVO_OUT_NO_RIGHT)) { o->kp_right = h.kp + (2 * (size_t)i + 1) * K; o->desc_right = h.desc + (2 * (size_t)i + 1) * K * 32; }
    const bool pose_only = (L.out_flags & SVO_OUT_POSE_INPUTS) != 0;
    o->u_right = h.u_right + (size_t)i * K; o->depth = h.depth + (size_t)i * K;
    o->n_prev = h.np_out[i]; o->n_map = h.np_out[ctx->cfg.max_batch + i];
    if (pose_only) { o->desc_left = nullptr; o->u_right = nullptr; }
    if ((in.n_prev || in.track_seq) && !pose_only) {
        o->bf_idx = h.bf_idx + (size_t)i * K; o->bf_dist = h.bf_dist + (size_t)i * K; o->bf_keep = h.bf_keep + (size_t)i * K;
        if (!ctx->cfg.skip_match_score) {
            o->p1_best_idx = h.p1_best_idx + (size_t)i * R; o->p1_best = h.p1_best + (size_t)i * R;
            o->p1_second = h.p1_second + (size_t)i * R;
        }
        o->p1_row_claimed = h.p1_row_claimed + (size_t)i * R;
        if (L.veto) o->p1_row_bad = h.p1_row_bad + (size_t)i * R;     // NULL: no frame of the batch ran the veto
    }
    if ((in.n_map || in.track_seq) && !pose_only) o->p2_row_claimed = h.p2_row_claimed + (size_t)i * R;
    if (in.track_seq) { o->mp_create = h.mp_create + (size_t)i * K; o->mp_xyz = h.mp_xyz + (size_t)i * K * 3; }
    o->claim_row = h.claim_row + (size_t)i * K;
    return SVO_OK;
}

int svo_batch_stage_ms(svo_ctx *ctx, int lane_i, float *ms, int n)
{
    if (!ctx || !ms || lane_i < 0 || lane_i >= (int)ctx->lanes.size()) return fail(ctx, SVO_E_INVALID, "svo_batch_stage_ms: bad argument");
    if (!ctx->profiling) return fail(ctx, SVO_E_INVALID, "svo_batch_stage_ms: profiling is off");
    Lane &L = ctx->lanes[lane_i];
    if (L.busy) return fail(ctx, SVO_E_INVALID, "svo_batch_stage_ms: call svo_batch_wait first");
    float v[14];
    CU(cudaEventElapsedTime(&v[0], L.ev[0], L.ev[11]));
    for (int k = 1; k < 12; ++k) CU(cudaEventElapsedTime(&v[k], L.ev[k - 1], L.ev[k]));
    // single kernels inside the matching stage (0 when the batch had no previous frame / no map)
    v[12] = v[13] = 0.f;
    if (cudaEventElapsedTime(&v[12], L.ev[12], L.ev[13]) != cudaSuccess) { cudaGetLastError(); v[12] = 0.f; }
    if (cudaEventElapsedTime(&v[13], L.ev[14], L.ev[15]) != cudaSuccess) { cudaGetLastError(); v[13] = 0.f; }
    for (int k = 0; k < n && k < 14; ++k) ms[k] = v[k];
    return n < 14 ? n : 14;
}

// ----------------------------------------------------------------------------- debug taps
long long svo_debug_tap(svo_ctx *ctx, int cam, int what, int level, void *out, size_t cap_bytes)
{
    if (!ctx || cam < 0 || cam > 1 || level < 0 || level >= ctx->g.nlevels || !out)
        return fail(ctx, SVO_E_INVALID, "svo_debug_tap: bad argument");
    if (!ctx->sync_have[cam]) return fail(ctx, SVO_E_INVALID, "svo_debug_tap: nothing extracted on cam %d", cam);
    CU(cudaSetDevice(ctx->cfg.device));
    const Geom &g = ctx->g; const LevelGeom &L = g.lv[level]; const Bufs &b = ctx->b;
    const int slot = ctx->sync_slot0 + cam;
    CU(cudaDeviceSynchronize());
    if (what == SVO_TAP_LEVEL || what == SVO_TAP_BLUR) {
        const size_t need = (size_t)L.w * L.h;
        if (cap_bytes < need) return fail(ctx, SVO_E_CAPACITY, "svo_debug_tap: need %zu bytes", need);
        const uint8_t *src = (what == SVO_TAP_LEVEL ? b.pyr : b.blur) + (size_t)slot * g.pyr_bytes + L.off;
        CU(cudaMemcpy2D(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
        return (long long)need;
    }
    int32_t *o = (int32_t *)out;
    const size_t cap = cap_bytes / (3 * sizeof(int32_t));
    if (what == SVO_TAP_FAST) {
        if (L.nbands == 0) return 0;
        std::vector<int> cnt(L.nbands);
        CU(cudaMemcpy(cnt.data(), b.bandcnt + (size_t)slot * g.bandcnt_total + L.bandcnt_off, sizeof(int) * L.nbands, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> band(L.band_cap);
        size_t n = 0;
        for (int bi = 0; bi < L.nbands; ++bi) {
            if (cnt[bi] > L.band_cap) return fail(ctx, SVO_E_CAPACITY, "band overflow");
            CU(cudaMemcpy(band.data(), b.bands + (size_t)slot * g.band_total + L.band_off + (size_t)bi * L.band_cap, sizeof(uint32_t) * cnt[bi], cudaMemcpyDeviceToHost));
            for (int i = 0; i < cnt[bi]; ++i, ++n)
                if (n < cap) { o[3 * n] = unpack_x(band[i]); o[3 * n + 1] = unpack_y(band[i]); o[3 * n + 2] = unpack_s(band[i]); }
        }
        return (long long)n;
    }
    if (what == SVO_TAP_SELECT1) {
        int n = 0;
        CU(cudaMemcpy(&n, b.kept1 + (size_t)slot * SVO_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<uint32_t> v(n > 0 ? n : 1);
        CU(cudaMemcpy(v.data(), b.cval + (size_t)slot * g.cand_total + L.cand_off, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n && (size_t)i < cap; ++i) { o[3 * i] = unpack_x(v[i]); o[3 * i + 1] = unpack_y(v[i]); o[3 * i + 2] = unpack_s(v[i]); }
        return n;
    }
    if (what == SVO_TAP_SELECT2) {
        int n = 0;
        CU(cudaMemcpy(&n, b.kept2 + (size_t)slot * SVO_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<uint32_t> v(n > 0 ? n : 1), k(n > 0 ? n : 1);
        CU(cudaMemcpy(v.data(), b.val2 + (size_t)slot * g.total2 + L.off2, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(k.data(), b.key2 + (size_t)slot * g.total2 + L.off2, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n && (size_t)i < cap; ++i) { o[3 * i] = unpack_x(v[i]); o[3 * i + 1] = unpack_y(v[i]); o[3 * i + 2] = (int32_t)k[i]; }
        return n;
    }
    return fail(ctx, SVO_E_INVALID, "svo_debug_tap: unknown tap %d", what);
}

int svo_debug_hamming_matrix(svo_ctx *ctx, const uint8_t *a, int na, const uint8_t *b, int nb, int32_t *dist)
{
    if (!ctx || na < 0 || nb < 0 || (na && !a) || (nb && !b) || !dist) return fail(ctx, SVO_E_INVALID, "svo_debug_hamming_matrix: bad argument");
    FrameBufs &s = ctx->sb;
    if (na > s.row_stride || nb > s.col_stride) return fail(ctx, SVO_E_CAPACITY, "svo_debug_hamming_matrix: %d x %d exceeds capacity %d", na, nb, s.col_stride);
    if (na == 0 || nb == 0) return 0;
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = ctx->sync_st;
    int *d_dump = nullptr;
    const int pitch = (nb + 31) & ~31;
    CU(cudaMalloc((void **)&d_dump, sizeof(int) * (size_t)na * pitch));
    CU(cudaMemsetAsync(d_dump, 0, sizeof(int) * (size_t)na * pitch, st));   // the pitch padding is copied back too
    CU(cudaMemcpyAsync(s.map, a, (size_t)na * 32, cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(s.cols, b, (size_t)nb * 32, cudaMemcpyDefault, st));
    TcArgs tc;
    memset(&tc, 0, sizeof(tc));
    tc.A = make_set(s.map, nullptr, 0, s.row_stride, na);
    tc.B = make_set(s.cols, nullptr, 0, s.col_stride, nb);
    tc.g.rows = tc.A; tc.g.cols = tc.B;
    TcExpandArgs ex;
    memset(&ex, 0, sizeof(ex));
    ex.set = tc.A; ex.img = s.img_map; ex.img_frame_stride = s.img_row_stride;
    launch_tc_expand(ex, 1, st, &ctx->launches);
    ex.set = tc.B; ex.img = s.img_cur; ex.img_frame_stride = s.img_col_stride;
    launch_tc_expand(ex, 1, st, &ctx->launches);
    tc.a_img = s.img_map; tc.a_img_frame_stride = s.img_row_stride; tc.b_img = s.img_cur; tc.b_img_frame_stride = s.img_col_stride;
    tc.dump = d_dump; tc.dump_rows = na; tc.dump_pitch = pitch;
    launch_tc_hamming(tc, TC_DUMP, 1, st, &ctx->launches);
    std::vector<int> h((size_t)na * pitch);
    CU(cudaMemcpyAsync(h.data(), d_dump, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    cudaFree(d_dump);
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) dist[(size_t)i * nb + j] = (256 - h[(size_t)i * pitch + j]) / 2;     // dot = 256 - 2 d
    return na;
}

int svo_debug_tc_profile(svo_ctx *ctx, long long *stamps, int n)
{
    if (!ctx || !stamps || n < 0) return fail(ctx, SVO_E_INVALID, "svo_debug_tc_profile: bad argument");
    if (!ctx->tc_prof) return fail(ctx, SVO_E_INVALID, "svo_debug_tc_profile: create the context with SVO_B200_TC_PROF=1 in the environment");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    const int m = n < 1024 ? n : 1024;
    CU(cudaMemcpy(stamps, ctx->tc_prof, sizeof(long long) * m, cudaMemcpyDeviceToHost));
    return m;
}

int svo_debug_retain_best(svo_ctx *ctx, const float *resp, int n, int n_points, int depth_limit, int32_t *idx_out)
{
    if (!ctx || n < 0 || (n && (!resp || !idx_out))) return fail(ctx, SVO_E_INVALID, "svo_debug_retain_best: bad argument");
    if (n == 0) return 0;
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = ctx->sync_st;
    float *key; uint32_t *val, *lp, *rp; int *kept;
    CU(cudaMalloc((void **)&key, sizeof(float) * n)); CU(cudaMalloc((void **)&val, sizeof(uint32_t) * n));
    CU(cudaMalloc((void **)&lp, sizeof(uint32_t) * n)); CU(cudaMalloc((void **)&rp, sizeof(uint32_t) * n));
    CU(cudaMalloc((void **)&kept, 2 * sizeof(int)));
    std::vector<uint32_t> iota(n);
    for (int i = 0; i < n; ++i) iota[i] = (uint32_t)i;
    CU(cudaMemcpyAsync(key, resp, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(val, iota.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(kept, 0, 2 * sizeof(int), st));
    launch_retain_best_raw(key, val, n, n_points, depth_limit, lp, rp, kept, kept + 1, st, &ctx->launches);
    int k = 0;
    CU(cudaMemcpyAsync(&k, kept, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (k > 0) CU(cudaMemcpy(idx_out, val, sizeof(uint32_t) * k, cudaMemcpyDeviceToHost));
    cudaFree(key); cudaFree(val); cudaFree(lp); cudaFree(rp); cudaFree(kept);
    return k;
}

}  // extern "C"
