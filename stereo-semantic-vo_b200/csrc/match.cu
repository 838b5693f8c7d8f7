// match.cu — 256-bit Hamming matching: BFMatcher 1-NN (src/pnpmatch.cc:266,278-299), the two
// sequential greedy scans of pnpmatch::poseEstimationPnP (src/pnpmatch.cc:75-95 + :99-153,
// :173-197) and frame::disp2Depth (src/frame.cc:140-164).  SURVEY.md Appendix B.
//
// Distance = popc(xor) over two uint4 halves of a 32-byte descriptor row (pnpmatch.cc:14-30).
// Column descriptors are staged in shared memory in tiles of 992 rows, 16-byte units swizzled
// so that both lane-strided and lane-blocked 128-bit reads are bank-conflict free; the row
// descriptor lives in registers.  Minima are reduced with redux.sync on (dist << 16 | col)
// keys, which yields "first minimum wins" for free.
//
// The greedy scans are sequentially dependent (a claim hides a column from every later
// row).  Exact parallel form:
//   1. k_shortlist : all row x column distances; per row the ascending-column list of
//                    columns with d < T (T = 15 pass 1, 60 pass 2 — beyond T a column can
//                    neither be claimed nor break the ratio test).
//   2. k_resolve   : one warp per frame walks the non-empty rows in order and applies the
//                    reference's accept rule against the live claim set (rows whose list
//                    overflowed are re-scanned exhaustively), recording claim times.
//   3. k_scores    : exact (bestIdx, bestDist, secondBestDist) per row, in parallel, by
//                    replaying each row against the columns not claimed before it.
#include "svo_internal.cuh"
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_pipeline.h>

#define COL_TILE 992          // 31 columns per lane
#define COLS_PER_LANE 31
#define M_THREADS 256
#define M_WARPS (M_THREADS / 32)

__device__ __forceinline__ int unit_of(int j, int h) { return 2 * j + (h ^ ((j >> 2) & 1)); }

__device__ __forceinline__ int set_count(const MatchSet &s, int f) { return s.count ? min(s.count[(size_t)f * s.count_stride], s.stride_rows) : s.fixed_count; }
__device__ __forceinline__ const uint8_t *set_desc(const MatchSet &s, int f)
{
    if (s.tab) return *reinterpret_cast<const uint8_t *const *>(reinterpret_cast<const char *>(s.tab) + (size_t)f * sizeof(FramePtrs));
    return s.desc + (size_t)f * s.desc_stride * 32;
}
// per-frame row_live / map_prev_row arrays (NULL = all live / no link)
__device__ __forceinline__ const uint8_t *live_of(const GreedyArgs &a, int f)
{
    if (a.fp) return a.use_live ? a.fp[f].prev_live : nullptr;
    return a.row_live ? a.row_live + (size_t)f * a.rows.stride_rows : nullptr;
}
__device__ __forceinline__ const int *map_prev_of(const GreedyArgs &a, int f)
{
    if (a.fp) return a.use_map_prev ? a.fp[f].map_prev_row : nullptr;
    return a.map_prev_row ? a.map_prev_row + (size_t)f * a.rows.stride_rows : nullptr;
}

// Pass 2 skips a local-map row whose map point is the one pass-1 row `pr` owns when pass 1 matched it
// (observations.count(CurrentFrame), src/pnpmatch.cc:165) or marked it bad (mp_local->bad, :163, set at :141).
__device__ __forceinline__ bool prev_row_done(const GreedyArgs &a, int f, int pr)
{
    const size_t o = (size_t)f * a.prev_stride + pr;
    return a.prev_row_claimed[o] || (a.prev_row_bad && a.prev_row_bad[o]);
}

// stage columns [c0, c0+nc) of a descriptor set into the swizzled shared tile
__device__ __forceinline__ void load_tile(uint4 *tile, const uint8_t *desc, int c0, int nc)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(desc) + (size_t)c0 * 2;
    for (int i = threadIdx.x; i < nc * 2; i += blockDim.x) {
        const int j = i >> 1, h = i & 1;
        tile[unit_of(j, h)] = src[i];
    }
}

struct Row { uint4 a, b; };
__device__ __forceinline__ Row load_row(const uint8_t *desc, int r)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(desc) + (size_t)r * 2;
    Row R; R.a = p[0]; R.b = p[1];
    return R;
}
// popcount of 8 words with one level of carry-save adders: 6 POPC (XU pipe) + 4 extra LOP3 (ALU)
// instead of 8 POPC.  The XU pipe is the bottleneck of these kernels (84 % busy with the plain
// form); the full 4-POPC adder tree moves the bottleneck to the ALU pipe and is no faster.
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }
__device__ __forceinline__ int popc8(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5,
                                     uint32_t x6, uint32_t x7)
{
    const uint32_t s1 = x0 ^ x1 ^ x2, c1 = maj3(x0, x1, x2);
    const uint32_t s2 = x3 ^ x4 ^ x5, c2 = maj3(x3, x4, x5);
    return __popc(s1) + __popc(s2) + __popc(x6) + __popc(x7) + 2 * (__popc(c1) + __popc(c2));
}
// Three carry-save adders: 5 POPC + 6 extra LOP3.  For kernels whose ALU pipe has head-room (k_shortlist) this
// is the balance point of the two pipes (XU 5/16 clk per pair against ~20/64 on the ALU).
__device__ __forceinline__ int popc8_5(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5,
                                       uint32_t x6, uint32_t x7)
{
    const uint32_t s1 = x0 ^ x1 ^ x2, c1 = maj3(x0, x1, x2);
    const uint32_t s2 = x3 ^ x4 ^ x5, c2 = maj3(x3, x4, x5);
    const uint32_t s3 = s1 ^ s2 ^ x6, c3 = maj3(s1, s2, x6);
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}
__device__ __forceinline__ int ham(const Row &R, const uint4 *tile, int j)
{
    const uint4 x = tile[unit_of(j, 0)], y = tile[unit_of(j, 1)];
    return popc8(R.a.x ^ x.x, R.a.y ^ x.y, R.a.z ^ x.z, R.a.w ^ x.w, R.b.x ^ y.x, R.b.y ^ y.y, R.b.z ^ y.z, R.b.w ^ y.w);
}
__device__ __forceinline__ int ham_global(const Row &R, const uint8_t *desc, int j)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(desc) + (size_t)j * 2;
    const uint4 x = p[0], y = p[1];
    return popc8(R.a.x ^ x.x, R.a.y ^ x.y, R.a.z ^ x.z, R.a.w ^ x.w, R.b.x ^ y.x, R.b.y ^ y.y, R.b.z ^ y.z, R.b.w ^ y.w);
}

__device__ __forceinline__ bool in_window(const float *win, const float *cxy, int j)
{
    if (!win) return true;
    const float du = cxy[2 * j] - win[0], dv = cxy[2 * j + 1] - win[1], r = win[2];
    return !(du < -r || du > r || dv < -r || dv > r);
}

// ---------------------------------------------------------------------------------------
// BFMatcher: per query the first minimum over the train set
// ---------------------------------------------------------------------------------------
#define BF_ROWS_PER_WARP 4

__global__ void __launch_bounds__(M_THREADS) k_bf(BfArgs a)
{
    __shared__ uint4 tile[COL_TILE * 2];
    const int f = blockIdx.y;
    const int nq = set_count(a.q, f), nt = set_count(a.t, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = (blockIdx.x * M_WARPS + warp) * BF_ROWS_PER_WARP;
    if (blockIdx.x * M_WARPS * BF_ROWS_PER_WARP >= nq) return;
    const uint8_t *qd = set_desc(a.q, f), *td = set_desc(a.t, f);
    Row R[BF_ROWS_PER_WARP];
    uint32_t best[BF_ROWS_PER_WARP];
#pragma unroll
    for (int k = 0; k < BF_ROWS_PER_WARP; ++k) {
        best[k] = 0xffffffffu;
        R[k] = load_row(qd, min(r0 + k, nq - 1));
    }
    for (int c0 = 0; c0 < nt; c0 += COL_TILE) {
        const int nc = min(COL_TILE, nt - c0);
        __syncthreads();
        load_tile(tile, td, c0, nc);
        __syncthreads();
        for (int j = lane; j < nc; j += 32) {
#pragma unroll
            for (int k = 0; k < BF_ROWS_PER_WARP; ++k) {
                const uint32_t key = ((uint32_t)ham(R[k], tile, j) << 16) | (uint32_t)(c0 + j);
                best[k] = min(best[k], key);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < BF_ROWS_PER_WARP; ++k) {
        const uint32_t m = __reduce_min_sync(0xffffffffu, best[k]);
        const int r = r0 + k;
        if (lane == 0 && r < nq) {
            const size_t o = (size_t)f * a.q.stride_rows + r;
            if (nt > 0) {
                a.idx[o] = (int)(m & 0xffffu); a.dist[o] = (int)(m >> 16);
                atomicMin(a.min_dist + f, (int)(m >> 16));
            } else { a.idx[o] = -1; a.dist[o] = -1; }
        }
    }
}

// keep = dist <= max(2*min_dist, 30)   (src/pnpmatch.cc:281-299)
__global__ void k_bf_keep(BfArgs a)
{
    const int f = blockIdx.y;
    const int nq = set_count(a.q, f), nt = set_count(a.t, f);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const size_t o = (size_t)f * a.q.stride_rows + i;
    const double thr = fmax(2.0 * (double)a.min_dist[f], 30.0);
    a.keep[o] = (nt > 0 && (double)a.dist[o] <= thr) ? 1 : 0;
}

__global__ void k_fill_int(int *p, int n, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

void launch_bf(const BfArgs &a, int nframes, cudaStream_t st, long long *launches)
{
    k_fill_int<<<(nframes + 255) / 256, 256, 0, st>>>(a.min_dist, nframes, 10000);
    const int maxq = a.q.count ? a.q.stride_rows : a.q.fixed_count;
    if (maxq > 0) {
        dim3 grid((maxq + M_WARPS * BF_ROWS_PER_WARP - 1) / (M_WARPS * BF_ROWS_PER_WARP), nframes);
        k_bf<<<grid, M_THREADS, 0, st>>>(a);
        dim3 g2((maxq + 255) / 256, nframes);
        k_bf_keep<<<g2, 256, 0, st>>>(a);
        *launches += 2;
    }
    ++*launches;
}

// ---------------------------------------------------------------------------------------
// greedy step 0: claim_time from the incoming claim mask; per-row outputs cleared
// ---------------------------------------------------------------------------------------
__global__ void k_greedy_init(GreedyArgs a, int reset_time)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = set_count(a.cols, f), M = set_count(a.rows, f);
    if (reset_time && i < N) {
        const size_t o = (size_t)f * a.cols.stride_rows + i;
        a.claim_time[o] = a.claimed[o] ? INT_MIN : INT_MAX;
    }
    if (i < M) {
        const size_t ro = (size_t)f * a.rows.stride_rows, o = ro + i;
        a.row_claimed[o] = 0;
        if (a.row_bad) a.row_bad[o] = 0;
        a.short_cnt[o] = 0;
        if (a.best_idx) { a.best_idx[o] = -1; a.best[o] = 256; a.second[o] = 256; }
        if (a.need_list || a.row_need) {
            // Work lists of the batch path's pass 2.  A row that is the map point of pass-1 row `pr`
            // (src/pnpmatch.cc:167: the same mappoint object, hence the same frozen m_descriptor in both passes) is
            // dropped when pass 1 matched it; otherwise, if the two descriptors really are equal, its distances
            // already sit in row `pr` of the matrix k_pairs wrote (k_reuse thresholds that row); every other
            // live row goes to k_shortlist.  The lists are unordered: rows are independent until k_resolve.
            const uint8_t *rl = live_of(a, f);
            const int *mpr = map_prev_of(a, f);
            bool live = !rl || rl[i];
            int reuse = -1;
            if (live && mpr) {
                const int pr = mpr[i];
                if (pr >= 0 && pr < a.prev_count[f]) {   // a link past the previous set is no link
                    if (prev_row_done(a, f, pr)) live = false;
                    else if (a.dmat && !a.win_gather) {
                        const Row P = load_row(set_desc(a.prev, f), pr), Q = load_row(set_desc(a.rows, f), i);
                        if (P.a.x == Q.a.x && P.a.y == Q.a.y && P.a.z == Q.a.z && P.a.w == Q.a.w &&
                            P.b.x == Q.b.x && P.b.y == Q.b.y && P.b.z == Q.b.z && P.b.w == Q.b.w) reuse = pr;
                    }
                }
            }
            if (a.row_need) {                                      // tensor-core pass 2: every live row is scanned
                a.row_need[o] = live ? 1 : 0;
                if (SVO_TC_GATHER_ROWS && live) a.need_list[ro + atomicAdd(a.list_cnt + 2 * f, 1)] = i;   // ... as a gathered tile row (any order)
            } else if (live) {
                if (reuse >= 0) a.reuse_list[ro + atomicAdd(a.list_cnt + 2 * f + 1, 1)] = i | (reuse << 16);
                else a.need_list[ro + atomicAdd(a.list_cnt + 2 * f, 1)] = i;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// batch pass 2: the columns pass 1 left free, ascending.  The reference's scan skips a claimed column before it
// computes anything (src/pnpmatch.cc:176 `if (CurrentFrame->MapPoints[j]) continue`), and every claim made before
// pass 2 starts is final, so those columns are invisible to all pass-2 rows: k_shortlist stages and scans only
// the free ones (a third of the columns are claimed by pass 1 on the bench sequence).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_free_cols(GreedyArgs a)
{
    __shared__ int wsum[32];
    __shared__ int base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = set_count(a.cols, f);
    const uint8_t *claimed = a.claimed + (size_t)f * a.cols.stride_rows;
    uint16_t *out = a.free_col + (size_t)f * a.cols.stride_rows;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < N; c0 += 1024) {
        const int c = c0 + tid;
        const bool fr = c < N && !claimed[c];
        const uint32_t m = __ballot_sync(0xffffffffu, fr);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int v = wsum[lane], inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const int off = base + __shfl_sync(0xffffffffu, inc - v, warp);
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (fr) out[off + __popc(m & ((1u << lane) - 1u))] = (uint16_t)c;
        __syncthreads();
        if (tid == 0) base += tot;
        __syncthreads();
    }
    if (tid == 0) a.free_cnt[f] = base;
}

// stage the columns idx[c0 .. c0+nc) of a descriptor set (and their indices) into the swizzled shared tile
__device__ __forceinline__ void load_tile_indexed(uint4 *tile, uint16_t *tcol, const uint8_t *desc, const uint16_t *idx, int c0, int nc)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(desc);
    for (int i = threadIdx.x; i < nc * 2; i += blockDim.x) {
        const int j = i >> 1, h = i & 1;
        const int c = idx[c0 + j];
        tile[unit_of(j, h)] = src[(size_t)c * 2 + h];
        if (h == 0) tcol[j] = (uint16_t)c;
    }
}

// ---------------------------------------------------------------------------------------
// greedy step 1: short lists (ascending column) of columns with d < T
// ---------------------------------------------------------------------------------------
#define SL_ROWS_PER_WARP 4

// Short lists live in two arrays so the common case (<= 32 entries) touches one 128-byte line per
// row: entries 0..31 in shortlist[row][32], entries 32..CAP-1 in shortlist_hi[row][CAP-32].
__device__ __forceinline__ uint32_t *short_slot(const GreedyArgs &a, size_t row, int pos)
{
    return pos < 32 ? a.shortlist + row * 32 + pos : a.shortlist_hi + row * (SVO_SHORT_CAP - 32) + (pos - 32);
}

// Pass 2 prunes each finished list to what can influence the row's decision: a row claims only a column
// with d < 30, and the ratio test second > 2 * best can only be broken by an entry with d <= 2 * best
// <= 2 * dmax, dmax = the largest d < 30 in the list.  Entries above that bound (the bulk of a T = 60
// list) are dropped, and a list without any d < 30 entry is emptied.  Overflowed lists are left alone
// (they are incomplete; the resolver re-scans those rows exhaustively).  Whole warp; returns the new count.
__device__ __forceinline__ int prune_list(const GreedyArgs &a, size_t row, int c, int lane)
{
    if (c == 0 || c > SVO_SHORT_CAP) return c;   // warp-uniform
    uint32_t e[SVO_SHORT_CAP / 32];
    int dmax = -1;
#pragma unroll
    for (int t = 0; t < SVO_SHORT_CAP / 32; ++t) {
        e[t] = lane + 32 * t < c ? *short_slot(a, row, lane + 32 * t) : 0xffffffffu;
        const int d = (int)(e[t] >> 16);
        if (d < 30) dmax = max(dmax, d);
    }
    dmax = __reduce_max_sync(0xffffffffu, dmax);
    __syncwarp();
    int n = 0;
    if (dmax >= 0) {
#pragma unroll
        for (int t = 0; t < SVO_SHORT_CAP / 32; ++t) {
            const bool keep = e[t] != 0xffffffffu && (int)(e[t] >> 16) <= 2 * dmax;
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) *short_slot(a, row, n + __popc(m & ((1u << lane) - 1u))) = e[t];
            n += __popc(m);
        }
    }
    return n;
}

// Persistent form: the work of a launch is the list of (frame, group of 48 rows) items, frame major, and every
// CTA takes a contiguous, equal share of it, so the frame's columns are staged ONCE per CTA (twice when its share
// straddles a frame boundary) instead of once per 32 rows, and the column loop runs without barriers.  When a
// frame has more columns than the tile holds, the tile is re-staged per item as before.
#define SLN_THREADS 384
#define SLN_WARPS (SLN_THREADS / 32)
#define SLN_GROUP (SLN_WARPS * SL_ROWS_PER_WARP)

template <bool WIN, bool FREE>
__global__ void __launch_bounds__(SLN_THREADS, 2) k_shortlist(GreedyArgs a, int T, int tile_cap, int nframes)
{
    extern __shared__ __align__(16) uint8_t sl_dyn[];
    uint4 *tile = reinterpret_cast<uint4 *>(sl_dyn);
    uint16_t *tcol = reinterpret_cast<uint16_t *>(sl_dyn + (size_t)tile_cap * 32);
    int *pref = reinterpret_cast<int *>(sl_dyn + (size_t)tile_cap * 34);      // [nframes + 1] first work item of a frame
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < nframes; base += 32) {
            const int f = base + lane;
            int gcount = 0;
            if (f < nframes) {
                const int Mf = set_count(a.rows, f);
                const int nr = a.need_list ? min(a.list_cnt[2 * f], Mf) : Mf;
                gcount = (nr + SLN_GROUP - 1) / SLN_GROUP;
            }
            int inc = gcount;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (f < nframes) pref[f] = run + inc - gcount;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) pref[nframes] = run;
    }
    __syncthreads();
    const int W = pref[nframes];
    const int w0 = (int)((long long)blockIdx.x * W / gridDim.x), w1 = (int)((long long)(blockIdx.x + 1) * W / gridDim.x);
    int f = 0, cur_f = -1, N = 0;
    bool single = false;
    const uint8_t *cd = nullptr;
    const uint16_t *fcol = nullptr;
    for (int w = w0; w < w1; ++w) {
        while (pref[f + 1] <= w) ++f;
        if (f != cur_f) {
            cur_f = f;
            // FREE: the columns are the frame's free-column list (k_free_cols), entries carry the original index
            N = FREE ? a.free_cnt[f] : set_count(a.cols, f);
            cd = set_desc(a.cols, f);
            fcol = FREE ? a.free_col + (size_t)f * a.cols.stride_rows : nullptr;
            single = N <= tile_cap;
            if (single) {
                __syncthreads();
                if (FREE) load_tile_indexed(tile, tcol, cd, fcol, 0, N);
                else load_tile(tile, cd, 0, N);
                __syncthreads();
            }
        }
        const int grp = w - pref[f];
        const int M = set_count(a.rows, f);
        const size_t ro = (size_t)f * a.rows.stride_rows;
        // rows to process: all of them, or (batch pass 2) the work list k_greedy_init compacted
        const int *list = a.need_list ? a.need_list + ro : nullptr;
        const int nrows = list ? min(a.list_cnt[2 * f], M) : M;
        const int r0 = (grp * SLN_WARPS + warp) * SL_ROWS_PER_WARP;
        const uint8_t *rd = set_desc(a.rows, f);
        const float *cxy = WIN ? a.cur_xy + (size_t)f * a.cols.stride_rows * 2 : nullptr;
        const uint8_t *rl = list ? nullptr : live_of(a, f);   // the work list holds live rows only
        Row R[SL_ROWS_PER_WARP];
        int cnt[SL_ROWS_PER_WARP], rid[SL_ROWS_PER_WARP];
        bool live[SL_ROWS_PER_WARP];
        float wu[SL_ROWS_PER_WARP], wv[SL_ROWS_PER_WARP], wr[SL_ROWS_PER_WARP];
#pragma unroll
        for (int k = 0; k < SL_ROWS_PER_WARP; ++k) {
            const int i = r0 + k;
            const int r = i < nrows ? (list ? list[i] : i) : (list ? list[nrows - 1] : M - 1);
            rid[k] = r;
            cnt[k] = 0;
            live[k] = i < nrows && (!rl || rl[r]);
            R[k] = load_row(rd, r);
            wu[k] = wv[k] = wr[k] = 0.f;
            if (WIN && i < nrows) { wu[k] = a.win_uvr[(ro + r) * 3]; wv[k] = a.win_uvr[(ro + r) * 3 + 1]; wr[k] = a.win_uvr[(ro + r) * 3 + 2]; }
        }
        for (int c0 = 0; c0 < N; c0 += tile_cap) {
            const int nc = min(tile_cap, N - c0);
            if (!single) {
                __syncthreads();
                if (FREE) load_tile_indexed(tile, tcol, cd, fcol, c0, nc);
                else load_tile(tile, cd, c0, nc);
                __syncthreads();
            }
            for (int jb = 0; jb < nc; jb += 32) {
                const int j = jb + lane;
                const bool inb = j < nc;
                const int jj = inb ? j : 0;
                const uint4 x = tile[unit_of(jj, 0)], y = tile[unit_of(jj, 1)];
                float cu = 0.f, cv = 0.f;
                if (WIN && inb) { cu = cxy[2 * (c0 + j)]; cv = cxy[2 * (c0 + j) + 1]; }
                int d[SL_ROWS_PER_WARP];
#pragma unroll
                for (int k = 0; k < SL_ROWS_PER_WARP; ++k)
                    d[k] = popc8_5(R[k].a.x ^ x.x, R[k].a.y ^ x.y, R[k].a.z ^ x.z, R[k].a.w ^ x.w,
                                 R[k].b.x ^ y.x, R[k].b.y ^ y.y, R[k].b.z ^ y.z, R[k].b.w ^ y.w);
                // one vote for the 4 rows: most 32-column steps hold no entry below T at all
                const int dmin = min(min(d[0], d[1]), min(d[2], d[3]));
                if (!__any_sync(0xffffffffu, inb && dmin < T)) continue;
#pragma unroll
                for (int k = 0; k < SL_ROWS_PER_WARP; ++k) {
                    if (!live[k]) continue;   // warp-uniform
                    bool hit = inb && d[k] < T;
                    if (WIN) {
                        const float du = cu - wu[k], dv = cv - wv[k];
                        hit = hit && !(du < -wr[k] || du > wr[k] || dv < -wr[k] || dv > wr[k]);
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, hit);
                    if (m) {
                        if (hit) {
                            const int pos = cnt[k] + __popc(m & ((1u << lane) - 1u));
                            if (pos < SVO_SHORT_CAP)
                                *short_slot(a, ro + rid[k], pos) = ((uint32_t)d[k] << 16) | (FREE ? (uint32_t)tcol[j] : (uint32_t)(c0 + j));
                        }
                        cnt[k] += __popc(m);
                    }
                }
            }
        }
        if (a.mode == SVO_GREEDY_PASS2) {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < SL_ROWS_PER_WARP; ++k)
                if (live[k]) cnt[k] = prune_list(a, ro + rid[k], cnt[k], lane);   // warp-uniform
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < SL_ROWS_PER_WARP; ++k)
                if (r0 + k < nrows) a.short_cnt[ro + rid[k]] = live[k] ? cnt[k] : 0;
        }
    }
}

// Projection window of one map point under the predicted pose (opt-in; ORB-SLAM2 SearchByProjection form; every
// float32 operation rounded separately, in a fixed order, so that the CPU restatement used by the tests gives the same bits).
__device__ __forceinline__ void project_point(const float *T, float fx, float fy, float cx, float cy, int W, int H, float th,
                                              const float *lscale, int nlevels, float X, float Y, float Z, int octave, float *uvr)
{
    const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], X), __fmul_rn(T[1], Y)), __fmul_rn(T[2], Z)), T[3]);
    const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], X), __fmul_rn(T[5], Y)), __fmul_rn(T[6], Z)), T[7]);
    const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], X), __fmul_rn(T[9], Y)), __fmul_rn(T[10], Z)), T[11]);
    float u = 0.f, v = 0.f, r = -1.f;
    if (zc > 0.f) {
        const float invz = __fdiv_rn(1.0f, zc);
        const float pu = __fadd_rn(__fmul_rn(__fmul_rn(fx, xc), invz), cx);
        const float pv = __fadd_rn(__fmul_rn(__fmul_rn(fy, yc), invz), cy);
        if (pu >= 0.f && pu < (float)W && pv >= 0.f && pv < (float)H) {
            const int o = min(max(octave, 0), nlevels - 1);
            u = pu; v = pv; r = __fmul_rn(th, lscale[o]);
        }
    }
    uvr[0] = u; uvr[1] = v; uvr[2] = r;
}

struct ProjectArgs { const float *xyz; const int *octave; int n; float T[12]; float fx, fy, cx, cy, th; int W, H, nlevels; float lscale[SVO_MAX_LEVELS]; float *uvr; };
__global__ void k_project(ProjectArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    project_point(a.T, a.fx, a.fy, a.cx, a.cy, a.W, a.H, a.th, a.lscale, a.nlevels, a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2],
                  a.octave ? a.octave[i] : 0, a.uvr + 3 * (size_t)i);
}
void launch_project(const float *xyz, const int *octave, int n, const float *Tcw12, float fx, float fy, float cx, float cy, int W, int H,
                    float th, const float *lscale, int nlevels, float *uvr, cudaStream_t st, long long *launches)
{
    if (n <= 0) return;
    ProjectArgs a;
    a.xyz = xyz; a.octave = octave; a.n = n; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.th = th; a.W = W; a.H = H; a.nlevels = nlevels; a.uvr = uvr;
    for (int k = 0; k < 12; ++k) a.T[k] = Tcw12[k];
    for (int k = 0; k < SVO_MAX_LEVELS; ++k) a.lscale[k] = k < nlevels ? lscale[k] : 1.f;
    k_project<<<(n + 255) / 256, 256, 0, st>>>(a);
    ++*launches;
}

// ---------------------------------------------------------------------------------------
// Batch pass 2 with projection windows (opt-in; the reference scans every column, src/pnpmatch.cc:173-190).
// k_win_prepare, one CTA per frame: the frame's windows and the current keypoints' positions go into the
// [frame][stride] arrays the resolver reads, and the keypoints are binned into square cells (CSR).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_win_prepare(GreedyArgs a)
{
    __shared__ int cell[SVO_WIN_CELLS + 1];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int N = set_count(a.cols, f), M = set_count(a.rows, f);
    const size_t ro = (size_t)f * a.rows.stride_rows, co = (size_t)f * a.cols.stride_rows;
    const int ncell = a.ncx * a.ncy;
    const svo_keypoint *kp = a.kp + (size_t)f * a.kp_frame_stride;
    const float *win = a.fp[f].map_win;
    if (win) { for (int i = tid; i < 3 * M; i += 256) a.win_out[ro * 3 + i] = win[i]; }
    else {   // windows projected here from the map points' positions and the predicted pose
        const FramePtrs &P = a.fp[f];
        for (int i = tid; i < M; i += 256)
            project_point(P.Tcw, P.fx, P.fy, P.cx, P.cy, a.img_w, a.img_h, P.proj_th, a.lscale, a.nlevels, P.map_xyz[3 * i], P.map_xyz[3 * i + 1],
                          P.map_xyz[3 * i + 2], P.map_octave ? P.map_octave[i] : 0, a.win_out + (ro + i) * 3);
    }
    for (int c = tid; c <= ncell; c += 256) cell[c] = 0;
    __syncthreads();
    auto cell_of = [&](float x, float y) {
        const int cx = min(max((int)x, 0) >> a.cell_shift, a.ncx - 1), cy = min(max((int)y, 0) >> a.cell_shift, a.ncy - 1);
        return cy * a.ncx + cx;
    };
    for (int j = tid; j < N; j += 256) {
        const float x = kp[j].x, y = kp[j].y;
        a.cur_xy_out[(co + j) * 2] = x; a.cur_xy_out[(co + j) * 2 + 1] = y;
        atomicAdd(&cell[cell_of(x, y)], 1);
    }
    __syncthreads();
    if (tid < 32) {                                      // exclusive scan of the cell counts
        int carry = 0;
        for (int b0 = 0; b0 <= ncell; b0 += 32) {
            const int c = b0 + tid <= ncell ? cell[b0 + tid] : 0;
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, d);
                if (tid >= d) inc += v;
            }
            if (b0 + tid <= ncell) cell[b0 + tid] = carry + inc - c;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    int *off = a.cell_off + (size_t)f * (SVO_WIN_CELLS + 1);
    for (int c = tid; c <= ncell; c += 256) off[c] = cell[c];
    __syncthreads();
    uint16_t *list = a.cell_list + co;
    for (int j = tid; j < N; j += 256) list[atomicAdd(&cell[cell_of(kp[j].x, kp[j].y)], 1)] = (uint16_t)j;   // order inside a cell is free
}

// One warp per live row: the cells under the row's window are contiguous per cell row in the CSR; lanes take 32
// candidates at a time, test the window exactly as in_window() does, and only then touch the descriptor.  Hits
// (d < T) are collected in shared memory and written in ascending column order (the resolver's invariant).
__global__ void __launch_bounds__(M_THREADS) k_shortlist_win(GreedyArgs a, int T)
{
    __shared__ uint32_t hits[M_WARPS][SVO_SHORT_CAP];
    const int f = blockIdx.y;
    const int M = set_count(a.rows, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ro = (size_t)f * a.rows.stride_rows, co = (size_t)f * a.cols.stride_rows;
    const int *list = a.need_list + ro;
    const int nrows = min(a.list_cnt[2 * f], M);
    const uint8_t *rd = set_desc(a.rows, f), *cd = set_desc(a.cols, f);
    const float *cxy = a.cur_xy + co * 2;
    const int *coff = a.cell_off + (size_t)f * (SVO_WIN_CELLS + 1);
    const uint16_t *clist = a.cell_list + co;
    const uint32_t lt = (1u << lane) - 1u;
    const float fw = (float)(a.ncx << a.cell_shift), fh = (float)(a.ncy << a.cell_shift);
    for (int i = blockIdx.x * M_WARPS + warp; i < nrows; i += gridDim.x * M_WARPS) {
        const int r = list[i];
        const Row R = load_row(rd, r);
        const float wu = a.win_uvr[(ro + r) * 3], wv = a.win_uvr[(ro + r) * 3 + 1], wr = a.win_uvr[(ro + r) * 3 + 2];
        // conservative cell range: anything that is not provably outside (NaN bounds included) is scanned
        const float xl = wu - wr, xh = wu + wr, yl = wv - wr, yh = wv + wr;
        const int cx0 = xl >= 0.f ? (xl < fw ? (int)xl >> a.cell_shift : a.ncx) : 0;
        const int cx1 = xh < fw ? (xh >= 0.f ? (int)xh >> a.cell_shift : -1) : a.ncx - 1;
        const int cy0 = yl >= 0.f ? (yl < fh ? (int)yl >> a.cell_shift : a.ncy) : 0;
        const int cy1 = yh < fh ? (yh >= 0.f ? (int)yh >> a.cell_shift : -1) : a.ncy - 1;
        int cnt = 0;
        if (cx0 <= cx1)
            for (int cy = cy0; cy <= cy1; ++cy) {
                const int b0 = coff[cy * a.ncx + cx0], b1 = coff[cy * a.ncx + cx1 + 1];
                for (int base = b0; base < b1; base += 32) {
                    const int idx = base + lane;
                    bool hit = false; int d = 0, col = 0;
                    if (idx < b1) {
                        col = clist[idx];
                        const float du = cxy[2 * col] - wu, dv = cxy[2 * col + 1] - wv;
                        if (!(du < -wr || du > wr || dv < -wr || dv > wr)) { d = ham_global(R, cd, col); hit = d < T; }
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int pos = cnt + __popc(m & lt);
                        if (pos < SVO_SHORT_CAP) hits[warp][pos] = ((uint32_t)d << 16) | (uint32_t)col;
                    }
                    cnt += __popc(m);
                }
            }
        __syncwarp();
        const int n = min(cnt, SVO_SHORT_CAP);
        for (int t = lane; t < n; t += 32) {             // rank by column (columns are distinct)
            const uint32_t e = hits[warp][t];
            int rank = 0;
            for (int u = 0; u < n; ++u) rank += (hits[warp][u] & 0xffffu) < (e & 0xffffu);
            *short_slot(a, ro + r, rank) = e;
        }
        __syncwarp();
        if (a.mode == SVO_GREEDY_PASS2) cnt = prune_list(a, ro + r, cnt, lane);
        if (lane == 0) a.short_cnt[ro + r] = cnt;
        __syncwarp();
    }
}

// Batch pass 2, rows whose distances k_pairs already computed (reuse_list): threshold the row of the u8 matrix,
// 4 columns per lane and step, 8 steps in flight (the matrix is not cache resident).  One warp per row.
__global__ void __launch_bounds__(M_THREADS) k_reuse(GreedyArgs a, int T)
{
    const int f = blockIdx.y;
    const int N = set_count(a.cols, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ro = (size_t)f * a.rows.stride_rows;
    const int nre = a.list_cnt[2 * f + 1];
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = blockIdx.x * M_WARPS + warp; i < nre; i += gridDim.x * M_WARPS) {
    const int e = a.reuse_list[ro + i], r = e & 0xffff, pr = e >> 16;
    const uint8_t *drow = a.dmat + (size_t)f * a.dmat_frame_stride + (size_t)pr * a.dmat_pitch;
    int cnt = 0;
    for (int cb = 0; cb < N; cb += 8 * 128) {
        uint32_t w8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int col = cb + 128 * u + 4 * lane;
            w8[u] = col < N ? __ldg(reinterpret_cast<const uint32_t *>(drow + col)) : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int col = cb + 128 * u + 4 * lane;
            uint32_t w = w8[u];
            if (col + 4 > N && col < N) w |= 0xffffffffu << (8 * (N - col));          // columns past N hold no distances
            const bool any = ((w - (uint32_t)T * 0x01010101u) & ~w & 0x80808080u) != 0;   // some byte < T (T <= 128)
            if (!__any_sync(0xffffffffu, any)) continue;
            int before = 0, tot = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t m = __ballot_sync(0xffffffffu, (int)((w >> (8 * q)) & 0xffu) < T);
                before += __popc(m & lt); tot += __popc(m);
            }
            int pos = cnt + before;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int d = (int)((w >> (8 * q)) & 0xffu);
                if (d < T) {
                    if (pos < SVO_SHORT_CAP) *short_slot(a, ro + r, pos) = ((uint32_t)d << 16) | (uint32_t)(col + q);
                    ++pos;
                }
            }
            cnt += tot;
        }
    }
    __syncwarp();
    cnt = prune_list(a, ro + r, cnt, lane);
    if (lane == 0) a.short_cnt[ro + r] = cnt;
    }
}

// ---------------------------------------------------------------------------------------
// greedy step 2: sequential resolution, one CTA per frame (warp 0 walks the rows)
// ---------------------------------------------------------------------------------------
extern __shared__ __align__(16) uint8_t resolve_smem[];

// true when pass 1 of frame f runs the "dynamic" test at all
__device__ __forceinline__ bool veto_active(const GreedyArgs &a, int f)
{
    if (a.mode != SVO_GREEDY_PASS1) return false;
    if (a.fp) return a.use_veto && a.fp[f].n_boxes > 0 && a.fp[f].boxes && a.fp[f].F && a.fp[f].prev_xy;
    return a.n_boxes > 0 && a.F;
}

__device__ bool veto_dynamic(const GreedyArgs &a, int f, int row, int col)
{
    // src/pnpmatch.cc:103-122.  Single call: boxes / F / row_xy / cur_xy arrays of the one frame; batch: the frame's
    // entries of the pointer table, current keypoint positions straight from the extractor's output.
    const int *boxes; const double *F; int nb; float cx, cy, lx, ly;
    if (a.fp) {
        const FramePtrs &P = a.fp[f];
        boxes = P.boxes; nb = P.n_boxes; F = P.F;
        lx = P.prev_xy[2 * row]; ly = P.prev_xy[2 * row + 1];
        const svo_keypoint &k = a.kp[(size_t)f * a.kp_frame_stride + col];
        cx = k.x; cy = k.y;
    } else {
        boxes = a.boxes + (size_t)f * a.n_boxes * 4; nb = a.n_boxes; F = a.F + (size_t)f * 9;
        const float *cxy = a.cur_xy + ((size_t)f * a.cols.stride_rows + col) * 2;
        const float *lxy = a.row_xy + ((size_t)f * a.rows.stride_rows + row) * 2;
        cx = cxy[0]; cy = cxy[1]; lx = lxy[0]; ly = lxy[1];
    }
    for (int k = 0; k < nb; ++k) {
        const int *bx = boxes + 4 * k;
        const int left = bx[0], right = bx[1], top = bx[2], bottom = bx[3];
        if (cx > left - 10 && cx < right + 10 && cy > top - 10 && cy < bottom + 10) {
            const double A = __dadd_rn(__dadd_rn(__dmul_rn(F[0], lx), __dmul_rn(F[1], ly)), F[2]);
            const double B = __dadd_rn(__dadd_rn(__dmul_rn(F[3], lx), __dmul_rn(F[4], ly)), F[5]);
            const double C = __dadd_rn(__dadd_rn(__dmul_rn(F[6], lx), __dmul_rn(F[7], ly)), F[8]);
            const double num = fabs(__dadd_rn(__dadd_rn(__dmul_rn(A, cx), __dmul_rn(B, cy)), C));
            const double den = __dsqrt_rn(__dadd_rn(__dmul_rn(A, A), __dmul_rn(B, B)));
            if (__ddiv_rn(num, den) > 0.1) return true;
        }
    }
    return false;
}

#define RES_THREADS 1024
#define RES_WARPS (RES_THREADS / 32)
#define RES_FREE 0x7fffffff
#define RES_SC 8     // compaction: list sizes cached in registers for rows up to 8 * 1024
#define RES_RC 6     // candidate rows per thread whose state stays in registers (6 * 1024 rows per frame)

// The greedy scan is sequential in the reference (a claim hides the column from every later row), but
// its result is the unique fixed point of a fully parallel map.  Let want[k] be the column the k-th
// candidate row claims (or none) and ct[c] = min{k : want[k] == c} the "claim time" of column c.  One
// sweep recomputes every row's decision against the columns with ct[c] >= k (not claimed by an earlier
// row) and rebuilds ct from the new decisions.  After sweep s the first s rows hold their sequential
// decisions (induction on k), so the iteration reaches the sequential answer after at most M sweeps;
// on real data dependency chains are short (measured: 3 sweeps for pass 1, 5 for pass 2 on the bench
// sequence).  A sweep costs one pass over the short-list entries (staged once in shared memory, CSR)
// with one thread per row.  One CTA per frame.
//
// Rows whose list overflowed SVO_SHORT_CAP are re-scanned exhaustively (a warp per row) in every sweep.
__global__ void __launch_bounds__(RES_THREADS) k_resolve(GreedyArgs a, int max_cols, int ent_cap, int unsorted)
{
    const int f = blockIdx.x;
    const int M = set_count(a.rows, f), N = set_count(a.cols, f);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // shared: [ct0: max_cols ints][ct1: max_cols ints][ent: ent_cap u32][base: max_cols bytes]
    int *ct0 = reinterpret_cast<int *>(resolve_smem);
    int *ct1 = ct0 + max_cols;
    uint32_t *ent = reinterpret_cast<uint32_t *>(ct1 + max_cols);
    uint8_t *pre = reinterpret_cast<uint8_t *>(ent + ent_cap);    // 1 = claimed before this pass
    __shared__ int wrow[RES_WARPS], went[RES_WARPS];
    __shared__ int s_total, s_novf;
    __shared__ int s_hist[SVO_SHORT_CAP + 2];
    const size_t ro = (size_t)f * a.rows.stride_rows, co = (size_t)f * a.cols.stride_rows;
    int *rows_ne = a.res_rows + ro;    // (row | min(cnt, CAP+1) << 16) of the rows that can claim, ascending
    int *roff = a.res_off + ro;        // CSR offset of the row's entries
    int *want = a.res_want + ro;       // column claimed (>= 0), -1 none, -2 vetoed ("bad")
    for (int j = tid; j < N; j += RES_THREADS) {
        const uint8_t c = a.claimed[co + j];
        pre[j] = c; ct0[j] = c ? -1 : RES_FREE; ct1[j] = c ? -1 : RES_FREE;
    }
    if (tid == 0) s_novf = 0;

    const int *mpr = (a.need_list || a.row_need) ? nullptr : map_prev_of(a, f);   // with work lists, rows pass 1 matched keep a zero count
    auto row_size = [&](int r) -> int {   // 0 = cannot claim, else min(cnt, CAP+1)
        if (r >= M) return 0;
        const int c = a.short_cnt[ro + r];
        if (c == 0) return 0;
        if (mpr) {
            const int pr = mpr[r];
            if (pr >= 0 && pr < a.prev_count[f] && prev_row_done(a, f, pr)) return 0;
        }
        return min(c, SVO_SHORT_CAP + 1);
    };
    // ---- ordered compaction of the candidate rows + CSR offsets of their entries
    const int seg = (((M + RES_WARPS - 1) / RES_WARPS) + 31) & ~31;
    const int beg = warp * seg, end = min(beg + seg, M);
    int cr = 0, ce = 0, novf = 0;
    int sc[RES_SC];                       // list sizes of this thread's first RES_SC rows (second pass reuses them)
#pragma unroll
    for (int u = 0; u < RES_SC; ++u) sc[u] = 0;
    for (int base = beg, u = 0; base < end; base += 32, ++u) {
        const int s = base + lane < end ? row_size(base + lane) : 0;
#pragma unroll
        for (int v = 0; v < RES_SC; ++v) if (v == u) sc[v] = s;
        cr += __popc(__ballot_sync(0xffffffffu, s > 0));
        ce += __reduce_add_sync(0xffffffffu, s > SVO_SHORT_CAP ? 0 : s);
        novf += __popc(__ballot_sync(0xffffffffu, s > SVO_SHORT_CAP));
    }
    if (lane == 0) { wrow[warp] = cr; went[warp] = ce; if (novf) atomicAdd(&s_novf, novf); }
    __syncthreads();
    int orow = 0, oent = 0, tot = 0;
    for (int w = 0; w < RES_WARPS; ++w) { if (w < warp) { orow += wrow[w]; oent += went[w]; } tot += wrow[w]; }
    for (int base = beg, u = 0; base < end; base += 32, ++u) {
        int s = 0;
        if (u < RES_SC) {
#pragma unroll
            for (int v = 0; v < RES_SC; ++v) if (v == u) s = sc[v];
        } else s = base + lane < end ? row_size(base + lane) : 0;
        const uint32_t m = __ballot_sync(0xffffffffu, s > 0);
        const int e = s > SVO_SHORT_CAP ? 0 : s;
        int inc = e;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        if (s > 0) {
            const int k = orow + __popc(m & ((1u << lane) - 1u));
            rows_ne[k] = (base + lane) | (s << 16);
            roff[k] = oent + inc - e;
            want[k] = -1;
        }
        orow += __popc(m);
        oent += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (tid == 0) s_total = tot;
    __syncthreads();
    const int total = s_total;
    const bool any_ovf = s_novf > 0;
    // ---- processing order: candidate rows sorted by list length (counting sort), so the 32 rows a warp walks
    // in lock step have equal trip counts.  The order of evaluation inside a sweep is free.
    int *perm = a.res_perm + ro;
    for (int i = tid; i < SVO_SHORT_CAP + 2; i += RES_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (int k = tid; k < total; k += RES_THREADS) atomicAdd(&s_hist[rows_ne[k] >> 16], 1);
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int b0 = 0; b0 < SVO_SHORT_CAP + 2; b0 += 32) {
            const int c = b0 + lane < SVO_SHORT_CAP + 2 ? s_hist[b0 + lane] : 0;
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += v;
            }
            if (b0 + lane < SVO_SHORT_CAP + 2) s_hist[b0 + lane] = carry + inc - c;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    for (int k = tid; k < total; k += RES_THREADS) perm[atomicAdd(&s_hist[rows_ne[k] >> 16], 1)] = k;
    __syncthreads();
    // ---- each thread owns the sorted positions pos(j) = j * 1024 + (tid, or 1023 - tid on odd j: the snake keeps
    // the per-warp work even although the positions are length-sorted).  The rows' metadata and decisions of the
    // first RES_RC rounds live in registers, so a sweep touches shared memory only.
    const uint8_t *rd = set_desc(a.rows, f), *cd = set_desc(a.cols, f);
    const float *cxy = a.cur_xy ? a.cur_xy + co * 2 : nullptr;
    const int rbase = a.row_base + (a.row_base_arr ? a.row_base_arr[f] : 0);
    const bool pass1 = a.mode == SVO_GREEDY_PASS1;
    const bool use_veto = veto_active(a, f);
    auto pos_of = [&](int j) -> int { return j * RES_THREADS + ((j & 1) ? RES_THREADS - 1 - tid : tid); };
    const int rounds = (total + RES_THREADS - 1) / RES_THREADS;
    int mk[RES_RC], mrs[RES_RC], moff[RES_RC], mw[RES_RC];   // k, row | s << 16, CSR offset, decision
#pragma unroll
    for (int j = 0; j < RES_RC; ++j) {
        mk[j] = -1; mrs[j] = 0; moff[j] = 0; mw[j] = -1;
        const int i = pos_of(j);
        if (i < total) { mk[j] = perm[i]; mrs[j] = rows_ne[mk[j]]; moff[j] = roff[mk[j]]; }
    }
    // stage the short lists in shared memory (rows beyond ent_cap stay in global memory)
    auto stage_row = [&](int r, int sz, int off) {
        if (sz > SVO_SHORT_CAP) return;
        if (off + sz > ent_cap) {
            // stays in global memory; k_pairs appends in arrival order, the sweeps need ascending columns
            if (unsorted)
                for (int i = 1; i < sz; ++i) {
                    const uint32_t e = *short_slot(a, ro + r, i);
                    int j = i - 1;
                    for (; j >= 0 && (*short_slot(a, ro + r, j) & 0xffffu) > (e & 0xffffu); --j) *short_slot(a, ro + r, j + 1) = *short_slot(a, ro + r, j);
                    *short_slot(a, ro + r, j + 1) = e;
                }
            return;
        }
        const uint4 *lo = reinterpret_cast<const uint4 *>(a.shortlist + (ro + r) * 32);
        const uint4 *hi = reinterpret_cast<const uint4 *>(a.shortlist_hi + (ro + r) * (SVO_SHORT_CAP - 32));
        for (int j = 0; j < sz; j += 4) {
            const uint4 v = j < 32 ? lo[j >> 2] : hi[(j - 32) >> 2];
            ent[off + j] = v.x;
            if (j + 1 < sz) ent[off + j + 1] = v.y;
            if (j + 2 < sz) ent[off + j + 2] = v.z;
            if (j + 3 < sz) ent[off + j + 3] = v.w;
        }
        if (unsorted)
            for (int i = 1; i < sz; ++i) {
                const uint32_t e = ent[off + i];
                int j = i - 1;
                for (; j >= 0 && (ent[off + j] & 0xffffu) > (e & 0xffffu); --j) ent[off + j + 1] = ent[off + j];
                ent[off + j + 1] = e;
            }
    };
#pragma unroll
    for (int j = 0; j < RES_RC; ++j)
        if (mk[j] >= 0) stage_row(mrs[j] & 0xffff, mrs[j] >> 16, moff[j]);
    for (int j = RES_RC; j < rounds; ++j) {
        const int i = pos_of(j);
        if (i < total) { const int k = perm[i]; stage_row(rows_ne[k] & 0xffff, rows_ne[k] >> 16, roff[k]); }
    }
    __syncthreads();
    int *ctc = ct0, *ctn = ct1;
    // one row's decision against the claim times of the previous sweep
    auto decide_row = [&](int k, int r, int sz, int off) -> int {
        int bd = 256, sd = 256, bi = -1;
        if (off + sz <= ent_cap) {
            for (int j = 0; j < sz; ++j) {
                const uint32_t e = ent[off + j];
                const int col = (int)(e & 0xffffu), d = (int)(e >> 16);
                if (ctc[col] >= k && d < bd) { sd = bd; bd = d; bi = col; }
            }
        } else {
            for (int j = 0; j < sz; ++j) {
                const uint32_t e = *short_slot(a, ro + r, j);
                const int col = (int)(e & 0xffffu), d = (int)(e >> 16);
                if (ctc[col] >= k && d < bd) { sd = bd; bd = d; bi = col; }
            }
        }
        int w = (bi >= 0 && (pass1 ? bd < 15 : (bd < 30 && sd > 2 * bd))) ? bi : -1;
        if (w >= 0 && use_veto && veto_dynamic(a, f, r, w)) w = -2;
        if (w >= 0) atomicMin(&ctn[w], k);
        return w;
    };
    // ---- sweeps
    for (;;) {
        int changed = 0;
#pragma unroll
        for (int j = 0; j < RES_RC; ++j) {
            if (mk[j] < 0 || (mrs[j] >> 16) > SVO_SHORT_CAP) continue;
            const int w = decide_row(mk[j], mrs[j] & 0xffff, mrs[j] >> 16, moff[j]);
            if (w != mw[j]) { mw[j] = w; changed = 1; }
        }
        for (int j = RES_RC; j < rounds; ++j) {     // more than RES_RC * 1024 candidate rows: metadata from global memory
            const int i = pos_of(j);
            if (i >= total) continue;
            const int k = perm[i], pk = rows_ne[k];
            if ((pk >> 16) > SVO_SHORT_CAP) continue;
            const int w = decide_row(k, pk & 0xffff, pk >> 16, roff[k]);
            if (w != want[k]) { want[k] = w; changed = 1; }
        }
        if (any_ovf) {   // rows with an unknown list: exhaustive scan, one warp per row
            for (int k = warp; k < total; k += RES_WARPS) {
                const int pk = rows_ne[k], r = pk & 0xffff, s = pk >> 16;
                if (s <= SVO_SHORT_CAP) continue;
                const Row R = load_row(rd, r);
                const float *win = a.win_uvr ? a.win_uvr + (ro + r) * 3 : nullptr;
                uint32_t key = 0xffffffffu;
                for (int j = lane; j < N; j += 32)
                    if (ctc[j] >= k && in_window(win, cxy, j)) key = min(key, ((uint32_t)ham_global(R, cd, j) << 16) | (uint32_t)j);
                const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
                int w = -1;
                if (kmin != 0xffffffffu) {
                    const int bd = (int)(kmin >> 16), bi = (int)(kmin & 0xffffu);
                    uint32_t sm = 256u;
                    for (int j = lane; j < bi; j += 32)
                        if (ctc[j] >= k && in_window(win, cxy, j)) sm = min(sm, (uint32_t)ham_global(R, cd, j));
                    const int sd = (int)__reduce_min_sync(0xffffffffu, sm);
                    if (pass1 ? bd < 15 : (bd < 30 && sd > 2 * bd)) w = bi;
                }
                if (lane == 0) {
                    if (w >= 0 && use_veto && veto_dynamic(a, f, r, w)) w = -2;
                    if (w >= 0) atomicMin(&ctn[w], k);
                    if (w != want[k]) { want[k] = w; changed = 1; }
                }
            }
        }
        if (!__syncthreads_or(changed)) break;
        int *t = ctc; ctc = ctn; ctn = t;
        for (int j = tid; j < N; j += RES_THREADS) ctn[j] = pre[j] ? -1 : RES_FREE;
        __syncthreads();
    }
    // ---- fixed point reached: publish the claims
#pragma unroll
    for (int j = 0; j < RES_RC; ++j)
        if (mk[j] >= 0 && (mrs[j] >> 16) <= SVO_SHORT_CAP) want[mk[j]] = mw[j];
    __syncthreads();
    for (int k = tid; k < total; k += RES_THREADS) {
        const int w = want[k], r = rows_ne[k] & 0xffff;
        if (w >= 0) {
            a.claimed[co + w] = 1;
            if (a.claim_row) a.claim_row[co + w] = rbase + r;
            a.claim_time[co + w] = rbase + r;
            a.row_claimed[ro + r] = 1;
        } else if (w == -2 && a.row_bad) a.row_bad[ro + r] = 1;
    }
}

// ---------------------------------------------------------------------------------------
// greedy step 3: exact (bestIdx, bestDist, secondBestDist) per row.  Lane l owns the
// contiguous columns [31 l, 31 l + 31) of each tile so "second = running best before the
// final update" composes across lanes and tiles.
// ---------------------------------------------------------------------------------------
template <bool WIN>
__global__ void __launch_bounds__(M_THREADS) k_scores(GreedyArgs a)
{
    __shared__ uint4 tile[COL_TILE * 2];
    __shared__ int s_time[COL_TILE];
    const int f = blockIdx.y;
    const int M = set_count(a.rows, f), N = set_count(a.cols, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x * M_WARPS >= M) return;
    const int r = blockIdx.x * M_WARPS + warp;
    const size_t ro = (size_t)f * a.rows.stride_rows, co = (size_t)f * a.cols.stride_rows;
    const uint8_t *rd = set_desc(a.rows, f), *cd = set_desc(a.cols, f);
    const float *cxy = a.cur_xy ? a.cur_xy + co * 2 : nullptr;
    const uint8_t *rl = live_of(a, f);
    const int *mpr = map_prev_of(a, f);
    bool live = r < M && (!rl || rl[r]);
    if (live && mpr) {
        const int pr = mpr[r];
        if (pr >= 0 && pr < a.prev_count[f] && prev_row_done(a, f, pr)) live = false;
    }
    const Row R = load_row(rd, min(r, M - 1));
    const float *win = (a.win_uvr && r < M) ? a.win_uvr + (ro + r) * 3 : nullptr;
    const int g = a.row_base + (a.row_base_arr ? a.row_base_arr[f] : 0) + r;
    int bd = 256, sd = 256, bi = -1;
    for (int c0 = 0; c0 < N; c0 += COL_TILE) {
        const int nc = min(COL_TILE, N - c0);
        __syncthreads();
        load_tile(tile, cd, c0, nc);
        for (int j = threadIdx.x; j < nc; j += M_THREADS) s_time[j] = a.claim_time[co + c0 + j];
        __syncthreads();
        if (!live) continue;
        int lb = 256, ls = 256, li = -1;
        const int jb = lane * COLS_PER_LANE;
#pragma unroll 1
        for (int k = 0; k < COLS_PER_LANE; ++k) {
            const int j = jb + k;
            if (j < nc && s_time[j] >= g && (!WIN || in_window(win, cxy, c0 + j))) {
                const int d = ham(R, tile, j);
                if (d < lb) { ls = lb; lb = d; li = c0 + j; }
            }
        }
        const uint32_t key = li >= 0 ? (((uint32_t)lb << 16) | (uint32_t)(li - c0)) : 0xffffffffu;
        const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
        if (kmin == 0xffffffffu) continue;
        const int tb = (int)(kmin >> 16), ti = c0 + (int)(kmin & 0xffffu);
        const int lstar = (ti - c0) / COLS_PER_LANE;
        const uint32_t before = __reduce_min_sync(0xffffffffu, lane < lstar ? (uint32_t)lb : 256u);
        const int ls_star = __shfl_sync(0xffffffffu, ls, lstar);
        const int tsec = min((int)before, ls_star);
        if (tb < bd) { sd = min(bd, tsec); bd = tb; bi = ti; }
    }
    if (live && lane == 0) { a.best_idx[ro + r] = bi; a.best[ro + r] = bd; a.second[ro + r] = sd; }
}

// ---------------------------------------------------------------------------------------
// Fused pass-1 front: the previous-frame x current-frame distance matrix is needed three times per
// frame -- BFMatcher (cur -> prev 1-NN, pnpmatch.cc:266-278), the pass-1 short lists (prev rows over
// cur columns, d < 15) and the exact (bestIdx, best, second) of every pass-1 row (match_score,
// pnpmatch.cc:99).  k_pairs computes every distance ONCE: a CTA owns a tile of 256 current-frame
// columns (8 per lane, lane-blocked) in swizzled shared memory and streams its share of the rows
// through registers, 4 rows per warp step.  Per step a lane holds 4 x 8 distances: they update the
// lane's 8 running column minima (BF), are tested against the pass-1 threshold (hits are rare and are
// appended with an atomic; the resolver sorts each list by column) and are packed into bytes and
// stored as 256 contiguous bytes per row into the frame's u8 distance matrix (min(d, 255)), which
// k_scores_m later scans instead of recomputing the popcounts.
// ---------------------------------------------------------------------------------------
#define PT_COLS 256
#define PT_ROWS 4

__device__ __forceinline__ int pt_unit(int j, int h) { const int u = 2 * j + h; return u ^ ((u >> 4) & 7); }

__global__ void __launch_bounds__(M_THREADS, 3) k_pairs(PairArgs p)
{
    __shared__ uint4 tile[PT_COLS * 2];
    __shared__ uint32_t cm[PT_COLS];
    const GreedyArgs &a = p.g;
    const int f = blockIdx.z;
    const int M = set_count(a.rows, f), N = set_count(a.cols, f);
    const int c0 = blockIdx.x * PT_COLS;
    if (c0 >= N || M <= 0) return;
    const int nc = min(PT_COLS, N - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *rd = set_desc(a.rows, f), *cd = set_desc(a.cols, f);
    const size_t ro = (size_t)f * a.rows.stride_rows;
    const uint8_t *rl = live_of(a, f);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(cd) + (size_t)c0 * 2;
        for (int i = tid; i < PT_COLS * 2; i += M_THREADS)
            tile[i ^ ((i >> 4) & 7)] = i < nc * 2 ? src[i] : make_uint4(0, 0, 0, 0);
        for (int i = tid; i < PT_COLS; i += M_THREADS) cm[i] = 0xffffffffu;
    }
    __syncthreads();
    uint32_t colmin[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) colmin[t] = 0xffffffffu;
    const int groups = (M + PT_ROWS - 1) / PT_ROWS;
    const int per = (groups + gridDim.y - 1) / gridDim.y;
    const int q0 = blockIdx.y * per, q1 = min(q0 + per, groups);
    uint8_t *dm = p.dmat + (size_t)f * p.dmat_frame_stride;
    const bool store_ok = c0 + 8 * lane + 8 <= p.dmat_pitch;
    const int T = p.T;
    for (int q = q0 + warp; q < q1; q += M_WARPS) {
        const int r0 = q * PT_ROWS;
        Row R[PT_ROWS];
        bool live[PT_ROWS];
#pragma unroll
        for (int k = 0; k < PT_ROWS; ++k) {
            const int r = min(r0 + k, M - 1);
            R[k] = load_row(rd, r);
            live[k] = r0 + k < M && (!rl || rl[r]);
        }
        uint32_t lo[PT_ROWS], hi[PT_ROWS];
#pragma unroll
        for (int k = 0; k < PT_ROWS; ++k) lo[k] = hi[k] = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int j = 8 * lane + t;
            const uint4 x = tile[pt_unit(j, 0)], y = tile[pt_unit(j, 1)];
            int d[PT_ROWS];
            uint32_t key[PT_ROWS];
#pragma unroll
            for (int k = 0; k < PT_ROWS; ++k) {
                d[k] = popc8(R[k].a.x ^ x.x, R[k].a.y ^ x.y, R[k].a.z ^ x.z, R[k].a.w ^ x.w,
                             R[k].b.x ^ y.x, R[k].b.y ^ y.y, R[k].b.z ^ y.z, R[k].b.w ^ y.w);
                // rows past M are copies of row M-1 with a larger index: they can never win the (d, row) minimum
                key[k] = (uint32_t)d[k] * 65536u + (uint32_t)(r0 + k);
                const uint32_t byte = (uint32_t)min(d[k], 255);
                if (t < 4) lo[k] = byte * (1u << (8 * t)) + lo[k]; else hi[k] = byte * (1u << (8 * (t - 4))) + hi[k];
            }
            const uint32_t kk = min(min(key[0], key[1]), min(key[2], key[3]));
            colmin[t] = min(colmin[t], kk);
            if (kk < (uint32_t)T << 16) {   // rare: some row of this lane's column is below the pass-1 threshold
#pragma unroll
                for (int k = 0; k < PT_ROWS; ++k)
                    if (d[k] < T && live[k] && j < nc) {
                        const int pos = atomicAdd(a.short_cnt + ro + r0 + k, 1);
                        if (pos < SVO_SHORT_CAP) *short_slot(a, ro + r0 + k, pos) = ((uint32_t)d[k] << 16) | (uint32_t)(c0 + j);
                    }
            }
        }
        if (store_ok) {
#pragma unroll
            for (int k = 0; k < PT_ROWS; ++k)
                if (r0 + k < M)
                    *reinterpret_cast<uint2 *>(dm + (size_t)(r0 + k) * p.dmat_pitch + c0 + 8 * lane) = make_uint2(lo[k], hi[k]);
        }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) atomicMin(&cm[8 * lane + t], colmin[t]);
    __syncthreads();
    for (int i = tid; i < nc; i += M_THREADS) atomicMin(p.bf_key + (size_t)f * a.cols.stride_rows + c0 + i, cm[i]);
}

// BF epilogue: key -> (trainIdx, distance), min_dist over the frame's queries, keep = d <= max(2*min, 30)
__global__ void __launch_bounds__(256) k_bf_finish(PairArgs p, BfArgs b)
{
    __shared__ int wmin[8];
    const int f = blockIdx.y;
    const int nq = set_count(b.q, f), nt = set_count(b.t, f);
    const uint32_t *key = p.bf_key + (size_t)f * b.q.stride_rows;
    int m = 10000;
    for (int i = threadIdx.x; i < nq; i += 256) m = min(m, (int)(key[i] >> 16));
    m = __reduce_min_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = m;
    __syncthreads();
    m = wmin[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = min(m, wmin[w]);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nq) return;
    const size_t o = (size_t)f * b.q.stride_rows + i;
    if (nt > 0) {
        const uint32_t k = key[i];
        const int d = (int)(k >> 16);
        b.idx[o] = (int)(k & 0xffffu); b.dist[o] = d;
        b.keep[o] = (double)d <= fmax(2.0 * (double)m, 30.0) ? 1 : 0;
    } else { b.idx[o] = -1; b.dist[o] = -1; b.keep[o] = 0; }
}

// Exact (bestIdx, bestDist, secondBestDist) of every live pass-1 row from the u8 distance matrix: lane l
// owns the contiguous columns [l W, (l+1) W), scans them in ascending order with the reference's running
// update, and the 32 partial results compose exactly (see k_scores).  A byte of 255 stands for d >= 255
// and is recomputed from the descriptors.
#define SM_ROWS_PER_WARP 4
__global__ void __launch_bounds__(M_THREADS) k_scores_m(PairArgs p)
{
    extern __shared__ int sm_time[];     // claim times, lane-block padded: column j at (j / W) * (W + 1) + j % W
    const GreedyArgs &a = p.g;
    const int f = blockIdx.y;
    const int M = set_count(a.rows, f), N = set_count(a.cols, f);
    if (blockIdx.x * M_WARPS * SM_ROWS_PER_WARP >= M) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = p.lane_cols;           // multiple of 16
    const size_t ro = (size_t)f * a.rows.stride_rows, co = (size_t)f * a.cols.stride_rows;
    for (int j = threadIdx.x; j < 32 * W; j += M_THREADS) {
        const int l = j / W;
        sm_time[l * (W + 1) + (j - l * W)] = j < N ? a.claim_time[co + j] : INT_MIN;
    }
    __syncthreads();
    const uint8_t *rd = set_desc(a.rows, f), *cd = set_desc(a.cols, f);
    const uint8_t *dm = p.dmat + (size_t)f * p.dmat_frame_stride;
    const int rb = a.row_base + (a.row_base_arr ? a.row_base_arr[f] : 0);
    const int *tl = sm_time + lane * (W + 1);
    const uint8_t *rl = live_of(a, f);
    for (int k = 0; k < SM_ROWS_PER_WARP; ++k) {
        const int r = (blockIdx.x * M_WARPS + warp) * SM_ROWS_PER_WARP + k;
        if (r >= M) break;
        if (rl && !rl[r]) continue;
        const int g = rb + r;
        const uint4 *src = reinterpret_cast<const uint4 *>(dm + (size_t)r * p.dmat_pitch + lane * W);
        int lb = 256, ls = 256, li = -1;
        for (int c = 0; c < W; c += 16) {
            const uint4 v = src[c >> 4];
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                int d = (int)((w4[i >> 2] >> (8 * (i & 3))) & 0xffu);
                if (tl[c + i] >= g && d < lb) {
                    const int j = lane * W + c + i;
                    if (d == 255) d = ham_global(load_row(rd, r), cd, j);
                    if (d < lb) { ls = lb; lb = d; li = j; }
                }
            }
        }
        const uint32_t key = li >= 0 ? (((uint32_t)lb << 16) | (uint32_t)li) : 0xffffffffu;
        const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
        int bd = 256, sd = 256, bi = -1;
        if (kmin != 0xffffffffu) {
            bd = (int)(kmin >> 16); bi = (int)(kmin & 0xffffu);
            const int lstar = bi / W;
            const uint32_t before = __reduce_min_sync(0xffffffffu, lane < lstar ? (uint32_t)lb : 256u);
            const int ls_star = __shfl_sync(0xffffffffu, ls, lstar);
            sd = min((int)before, ls_star);
        }
        if (lane == 0) { a.best_idx[ro + r] = bi; a.best[ro + r] = bd; a.second[ro + r] = sd; }
    }
}

// Tensor-core pass 2: TC_SHORT leaves every row's hits (d < 60) in SVO_TC_STREAMS segments of SVO_TC_SEG slots, one per
// contiguous column range, each ascending, with the segment lengths packed into short_cnt (one byte each).  One warp
// per row: join the segments into one ascending list and prune it exactly as k_shortlist prunes its own
// (see prune_list): entries that cannot influence the row's decision go, a list without any d < 30 entry is emptied.
// A segment that overflowed makes the row an "unknown list" row (count > SVO_SHORT_CAP), which the resolver re-scans
// exhaustively.
__global__ void __launch_bounds__(M_THREADS) k_prune_lists(GreedyArgs a)
{
    const int f = blockIdx.y;
    const int M = set_count(a.rows, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ro = (size_t)f * a.rows.stride_rows;
    const int r = blockIdx.x * M_WARPS + warp;
    if (r >= M) return;
    const uint32_t packed = (uint32_t)a.short_cnt[ro + r];
    if (packed == 0) return;                              // warp-uniform
    constexpr int PER = SVO_TC_SEG / 32, NE = SVO_TC_STREAMS * PER;
    uint32_t e[NE];
    int dmax = -1;
    bool ovf = false;
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int q = u / PER, o = (u % PER) * 32 + lane;
        const int c = (int)((packed >> (8 * q)) & 0xffu);
        ovf |= c > SVO_TC_SEG;
        e[u] = o < min(c, SVO_TC_SEG) ? *short_slot(a, ro + r, q * SVO_TC_SEG + o) : 0xffffffffu;
        const int d = (int)(e[u] >> 16);
        if (d < 30) dmax = max(dmax, d);
    }
    if (ovf) { if (lane == 0) a.short_cnt[ro + r] = SVO_SHORT_CAP + 1; return; }
    dmax = __reduce_max_sync(0xffffffffu, dmax);
    __syncwarp();
    int n = 0;
    if (dmax >= 0) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
            const bool keep = e[u] != 0xffffffffu && (int)(e[u] >> 16) <= 2 * dmax;
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) *short_slot(a, ro + r, n + __popc(m & ((1u << lane) - 1u))) = e[u];
            n += __popc(m);
        }
    }
    if (lane == 0) a.short_cnt[ro + r] = n;
}

void launch_prune_lists(const GreedyArgs &a, int nframes, cudaStream_t st, long long *launches)
{
    const int maxM = a.rows.count ? a.rows.stride_rows : a.rows.fixed_count;
    if (maxM <= 0 || nframes <= 0) return;
    const int gx = (maxM + M_WARPS - 1) / M_WARPS;
    k_prune_lists<<<dim3(gx, nframes), M_THREADS, 0, st>>>(a);
    ++*launches;
}

static int g_resolve_smem_limit = 48 * 1024;
static int g_shortlist_smem_limit = 48 * 1024;   // dynamic shared memory of one k_shortlist CTA (two per SM)
static int g_num_sms = 148;

int setup_match_attributes()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_num_sms = sms;
    g_shortlist_smem_limit = 110 * 1024;
    if (cudaFuncSetAttribute(k_shortlist<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_shortlist_smem_limit) != cudaSuccess ||
        cudaFuncSetAttribute(k_shortlist<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_shortlist_smem_limit) != cudaSuccess ||
        cudaFuncSetAttribute(k_shortlist<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_shortlist_smem_limit) != cudaSuccess)
        return 1;
    g_resolve_smem_limit = 200 * 1024;
    if (cudaFuncSetAttribute(k_scores_m, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) return 1;
    if (setup_tc_attributes() != 0) return 1;
    return (int)cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, g_resolve_smem_limit);
}

// k_resolve keeps 9 bytes per column in shared memory next to the staged short lists
int greedy_max_cols() { return (200 * 1024 - 16 * 1024) / 9; }

// Short-list entries k_resolve stages in shared memory (the rest is read from global memory).  Pruned pass-2 lists
// hold a few entries per row, so 10 k entries cover a 5000-row map; asking for the whole 200 KB instead would keep
// every other lane's CTAs off the SM while a resolver runs (measured: +0.8 % batch throughput with the small carve-out).
#define RES_ENT_TARGET 10240
static int resolve_ent_cap(int colsA)
{
    const int fit = (g_resolve_smem_limit - colsA * 9 - 64) / 4;
    return fit < RES_ENT_TARGET ? fit : RES_ENT_TARGET;
}

void launch_greedy(const GreedyArgs &a, int nframes, bool want_scores, cudaStream_t st, long long *launches,
                   cudaEvent_t ev0, cudaEvent_t ev1)
{
    const int maxM = a.rows.count ? a.rows.stride_rows : a.rows.fixed_count;
    const int maxN = a.cols.count ? a.cols.stride_rows : a.cols.fixed_count;
    if (maxM <= 0 || nframes <= 0) return;
    const int mx = maxM > maxN ? maxM : maxN;
    dim3 gi((mx + 255) / 256, nframes);
    if (a.need_list && (!a.row_need || SVO_TC_GATHER_ROWS)) cudaMemsetAsync(a.list_cnt, 0, sizeof(int) * 2 * nframes, st);
    if (a.win_gather) { k_win_prepare<<<nframes, 256, 0, st>>>(a); ++*launches; }
    k_greedy_init<<<gi, 256, 0, st>>>(a, 1);
    if (a.free_col && !a.win_gather && !a.win_uvr) { k_free_cols<<<nframes, 1024, 0, st>>>(a); ++*launches; }
    const int T = a.mode == SVO_GREEDY_PASS1 ? 15 : 60;
    // k_shortlist: two persistent CTAs per SM share the (frame, row group) items; the tile holds a whole frame's columns
    // when they fit (34 bytes per column)
    int tile_cap = (maxN + 31) & ~31;
    const int tile_max = ((g_shortlist_smem_limit - 4 * (nframes + 1)) / 34) & ~31;
    if (tile_cap > tile_max) tile_cap = tile_max;
    if (tile_cap < 32) tile_cap = 32;
    const size_t sl_smem = (size_t)tile_cap * 34 + sizeof(int) * (size_t)(nframes + 1);
    const int items = nframes * ((maxM + SLN_GROUP - 1) / SLN_GROUP);
    const int gs = items < 2 * g_num_sms ? items : 2 * g_num_sms;
    if (ev0) cudaEventRecord(ev0, st);
    if (a.row_need) {
        // tensor-core tiles: rows = local map (tile rows), columns = the free columns in ascending order
        TcArgs tc;
        TcExpandArgs ex;
        memset(&tc, 0, sizeof(tc)); memset(&ex, 0, sizeof(ex));
        ex.set = a.rows; ex.img = a.img_rows; ex.img_frame_stride = a.img_rows_stride;
        tc.A = a.rows; tc.B = a.cols; tc.g = a; tc.T = T; tc.row_need = a.row_need; tc.prof = a.tc_prof;
        tc.a_img = a.img_rows; tc.a_img_frame_stride = a.img_rows_stride;
        if (SVO_TC_GATHER_ROWS && a.need_list) {
            // only the rows pass 2 scans become tile rows (a quarter of a 5000-row map is matched by pass 1 on the bench
            // sequence): k_greedy_init listed them, the expansion gathers them, TC_SHORT maps a tile row back to its map row
            ex.index32 = a.need_list; ex.index_cnt = a.list_cnt; ex.index_cnt_stride = 2; ex.index_stride = a.rows.stride_rows;
            tc.a_index = a.need_list; tc.a_index_cnt = a.list_cnt;
        }
        TcExpandArgs ex1 = ex;
        ex1.index32 = nullptr; ex1.index_cnt = nullptr; ex1.index_cnt_stride = 0;
        if (a.free_col) {   // the columns pass 1 left free, gathered in ascending order
            ex1.set = a.cols; ex1.index = a.free_col; ex1.index_cnt = a.free_cnt; ex1.index_stride = a.cols.stride_rows;
            ex1.img = a.img_free; ex1.img_frame_stride = a.img_free_stride;
            launch_tc_expand2(ex, ex1, nframes, st, launches);
            tc.b_index = a.free_col; tc.b_index_cnt = a.free_cnt; tc.b_index_stride = a.cols.stride_rows;
            tc.b_img = a.img_free; tc.b_img_frame_stride = a.img_free_stride;
        } else {
            if (!a.img_cols_ready) {
                ex1.set = a.cols; ex1.img = a.img_cols; ex1.img_frame_stride = a.img_cols_stride;
                launch_tc_expand2(ex, ex1, nframes, st, launches);
            } else launch_tc_expand(ex, nframes, st, launches);
            tc.b_img = a.img_cols; tc.b_img_frame_stride = a.img_cols_stride;
        }
        launch_tc_hamming(tc, TC_SHORT, nframes, st, launches);
        launch_prune_lists(a, nframes, st, launches);
        --*launches;   // the common tail below counts three launches
    } else if (a.win_gather) {
        const int gx = (maxM + M_WARPS - 1) / M_WARPS;
        k_shortlist_win<<<dim3(gx < 160 ? gx : 160, nframes), M_THREADS, 0, st>>>(a, T);
    } else if (a.win_uvr) k_shortlist<true, false><<<gs, SLN_THREADS, sl_smem, st>>>(a, T, tile_cap, nframes);
    else if (a.free_col) k_shortlist<false, true><<<gs, SLN_THREADS, sl_smem, st>>>(a, T, tile_cap, nframes);
    else k_shortlist<false, false><<<gs, SLN_THREADS, sl_smem, st>>>(a, T, tile_cap, nframes);
    if (ev1) cudaEventRecord(ev1, st);
    if (a.need_list && a.dmat && !a.win_gather && !a.row_need) {
        const int gx = (maxM + M_WARPS - 1) / M_WARPS;
        k_reuse<<<dim3(gx < 128 ? gx : 128, nframes), M_THREADS, 0, st>>>(a, T);   // warps stride over the frame's reuse list
        ++*launches;
    }
    // shared memory: two claim-time arrays + pre-claimed bytes + as many short-list entries as fit
    const int colsA = (maxN + 3) & ~3;
    const int ent_cap = resolve_ent_cap(colsA);
    const size_t smem = (size_t)colsA * 9 + (size_t)ent_cap * 4;
    k_resolve<<<nframes, RES_THREADS, smem, st>>>(a, colsA, ent_cap, 0);
    *launches += 3;
    if (want_scores && a.best_idx) {
        dim3 gf((maxM + M_WARPS - 1) / M_WARPS, nframes);
        if (a.win_uvr) k_scores<true><<<gf, M_THREADS, 0, st>>>(a);
        else k_scores<false><<<gf, M_THREADS, 0, st>>>(a);
        ++*launches;
    }
}

// BF + greedy pass 1 of a batch with every distance computed once (k_pairs).
void launch_pass1_fused(const PairArgs &p0, const BfArgs &b, int nframes, cudaStream_t st, long long *launches,
                        cudaEvent_t ev0, cudaEvent_t ev1, cudaStream_t st_scores, cudaEvent_t e_resolved)
{
    PairArgs p = p0;
    const GreedyArgs &a = p.g;
    const int maxM = a.rows.count ? a.rows.stride_rows : a.rows.fixed_count;
    const int maxN = a.cols.count ? a.cols.stride_rows : a.cols.fixed_count;
    if (maxM <= 0 || maxN <= 0 || nframes <= 0) return;
    const int mx = maxM > maxN ? maxM : maxN;
    dim3 gi((mx + 255) / 256, nframes);
    k_greedy_init<<<gi, 256, 0, st>>>(a, 1);
    cudaMemsetAsync(p.bf_key, 0xff, sizeof(uint32_t) * (size_t)nframes * a.cols.stride_rows, st);
    p.T = 15;
    const int tiles = (maxN + PT_COLS - 1) / PT_COLS;
    int splits = (148 * 6 + tiles * nframes - 1) / (tiles * nframes);   // about two waves of 3 CTAs per SM
    const int max_splits = (maxM + 8 * PT_ROWS - 1) / (8 * PT_ROWS);
    splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
    if (ev0) cudaEventRecord(ev0, st);
    TcArgs tc;
    if (p.use_tc) {
        // queries (current frame) are the tile rows, the previous frame's rows stream past them in ascending order
        memset(&tc, 0, sizeof(tc));
        TcExpandArgs ex;
        memset(&ex, 0, sizeof(ex));
        ex.set = a.cols; ex.img = a.img_cols; ex.img_frame_stride = a.img_cols_stride;
        TcExpandArgs ex1 = ex;
        ex1.set = a.rows; ex1.img = a.img_rows; ex1.img_frame_stride = a.img_rows_stride;
        launch_tc_expand2(ex, ex1, nframes, st, launches);
        tc.A = a.cols; tc.B = a.rows; tc.g = a; tc.T = p.T; tc.bf_key = p.bf_key; tc.prof = a.tc_prof;
        tc.a_img = a.img_cols; tc.a_img_frame_stride = a.img_cols_stride;
        tc.b_img = a.img_rows; tc.b_img_frame_stride = a.img_rows_stride;
        launch_tc_hamming(tc, TC_PAIRS, nframes, st, launches);
        --*launches;
        if (p.bf_last.tab) {   // tracked frames: the per-query minima once more, over the last frame's own descriptors (no candidates: T = 0)
            MatchSet last = a.rows;
            last.tab = p.bf_last.tab;
            ex.set = last; ex.img = p.img_last; ex.img_frame_stride = p.img_last_stride;
            launch_tc_expand(ex, nframes, st, launches);
            TcArgs tb = tc;
            tb.B = last; tb.b_img = p.img_last; tb.b_img_frame_stride = p.img_last_stride; tb.T = 0;
            launch_tc_hamming(tb, TC_PAIRS, nframes, st, launches);
        }
    } else k_pairs<<<dim3(tiles, splits, nframes), M_THREADS, 0, st>>>(p);
    if (ev1) cudaEventRecord(ev1, st);
    k_bf_finish<<<dim3((maxN + 255) / 256, nframes), 256, 0, st>>>(p, b);
    const int colsA = (maxN + 3) & ~3;
    const int ent_cap = resolve_ent_cap(colsA);
    k_resolve<<<nframes, RES_THREADS, (size_t)colsA * 9 + (size_t)ent_cap * 4, st>>>(a, colsA, ent_cap, 1);
    p.lane_cols = (((maxN + 31) / 32) + 15) & ~15;
    const size_t sm = (size_t)32 * (p.lane_cols + 1) * sizeof(int);
    cudaStream_t sq = st;
    if (st_scores) {   // the scores only leave the device: another branch of the graph, beside pass 2
        cudaEventRecord(e_resolved, st);
        cudaStreamWaitEvent(st_scores, e_resolved, 0);
        sq = st_scores;
    }
    if (p.skip_scores) { *launches += 4; return; }
    if (p.use_tc) {
        tc.A = a.rows; tc.B = a.cols;      // previous-frame rows are the tile rows, current columns in ascending order
        tc.a_img = a.img_rows; tc.a_img_frame_stride = a.img_rows_stride;
        tc.b_img = a.img_cols; tc.b_img_frame_stride = a.img_cols_stride;
        launch_tc_hamming(tc, TC_SCORES, nframes, sq, launches);
        --*launches;
    } else k_scores_m<<<dim3((maxM + M_WARPS * SM_ROWS_PER_WARP - 1) / (M_WARPS * SM_ROWS_PER_WARP), nframes), M_THREADS, sm, sq>>>(p);
    *launches += 5;
}

// ---------------------------------------------------------------------------------------
// frame::disp2Depth (src/frame.cc:140-164)
// ---------------------------------------------------------------------------------------
__global__ void k_disp2depth(const float *__restrict__ disp, float *__restrict__ depth, size_t n, float bf)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = disp[i];
        depth[i] = d != 0.f ? __fdiv_rn(bf, d) : -1.f;
    }
}

void launch_disp2depth(const float *disp, float *depth, size_t n, float bf, cudaStream_t st, long long *launches)
{
    if (!n) return;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_disp2depth<<<(unsigned)blocks, 256, 0, st>>>(disp, depth, n, bf);
    ++*launches;
}
