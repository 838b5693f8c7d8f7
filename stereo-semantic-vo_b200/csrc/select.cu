// select.cu — cv::KeyPointsFilter::retainBest on the device, ORDER included.
//
// cv::ORB culls each level twice (2*quota by FAST score, quota by Harris response) with
//   std::nth_element(begin, begin+n-1, end, response-greater);
//   amb = v[n-1].response;  std::partition(begin+n, end, response >= amb);
// and never sorts, so the output order is whatever libstdc++'s introselect and partition
// leave — and the reference's matchers are order-sensitive
// (src/pnpmatch.cc:75-95 first-minimum tie-break, greedy claims).  This file replays both
// algorithms exactly, as a data-parallel program in one thread block per (image, level):
//
//   Hoare's unguarded partition is deterministic: the k-th element (from the left) at which
//   the left scan stops is swapped with the k-th element (from the right) at which the
//   right scan stops, for as long as the former lies left of the latter.  So one round is:
//   flag stoppers -> block-wide prefix sums -> positions of the k-th stoppers -> count the
//   crossing point m -> m independent swaps.  std::partition has the same two-pointer shape.
//   The median-of-3 pivot move, the <=3-element insertion sort and the (never observed)
//   depth-limit heap-select fallback are run by one thread, instruction for instruction.
//
// tools/model_retain_best.py is the executable design model of this file; the oracle
// (oracle/retain_best.cpp) calls the real std:: algorithms.
#include "svo_internal.cuh"
#include <stdlib.h>

#define SEL_THREADS 1024
#define SEL_WARPS (SEL_THREADS / 32)

struct SelSmem {
    int wl[SEL_WARPS], wr[SEL_WARPS];
    int m;
};

// Element stores.  The replay only needs "key of element i" and moves of whole elements, so the two culls use
// different layouts: the first cull's key IS the FAST score packed in the candidate's top byte (no key array at
// all: 8 bytes of shared memory per candidate with 16-bit position scratch), the second cull carries separate
// float Harris responses.
struct Item { float k; uint32_t v; };
struct KeyVal {
    float *key; uint32_t *val;
    __device__ __forceinline__ float k(int i) const { return key[i]; }
    __device__ __forceinline__ Item get(int i) const { Item t; t.k = key[i]; t.v = val[i]; return t; }
    __device__ __forceinline__ void put(int i, Item t) const { key[i] = t.k; val[i] = t.v; }
    __device__ __forceinline__ void copy(int dst, int src) const { key[dst] = key[src]; val[dst] = val[src]; }
    __device__ __forceinline__ void swap(int i, int j) const { const Item t = get(i); copy(i, j); put(j, t); }
};
struct ScoreInVal {
    uint32_t *val;
    __device__ __forceinline__ float k(int i) const { return (float)(val[i] >> 24); }
    __device__ __forceinline__ Item get(int i) const { Item t; t.v = val[i]; t.k = (float)(t.v >> 24); return t; }
    __device__ __forceinline__ void put(int i, Item t) const { val[i] = t.v; }
    __device__ __forceinline__ void copy(int dst, int src) const { val[dst] = val[src]; }
    __device__ __forceinline__ void swap(int i, int j) const { const uint32_t t = val[i]; val[i] = val[j]; val[j] = t; }
};

// Two-pointer swap round over [lo, hi).  MODE 0: Hoare step against `pivot` (left scan stops
// on !(x > pivot), right scan on !(pivot > x)).  MODE 1: std::partition with x >= pivot (left
// stops on false, right on true).  Returns through nl/nr the stopper totals and `cut`.
// P: position scratch type (uint16_t when the arrays live in shared memory, n < 65536).
template <int MODE, class A, class P>
__device__ void two_pointer_round(const A &e, int lo, int hi, float pivot, P *lpos, P *rasc, SelSmem &sh, int &nl, int &nr, int &cut)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len = hi - lo;
    const int seg = (((len + SEL_WARPS - 1) / SEL_WARPS) + 31) & ~31;
    const int beg = lo + warp * seg, end = min(beg + seg, hi);
    auto flags = [&](int i, bool &fl, bool &fr) {
        if (i < end) {
            const float x = e.k(i);
            if (MODE == 0) { fl = !(x > pivot); fr = !(pivot > x); }
            else { fr = x >= pivot; fl = !fr; }
        } else { fl = false; fr = false; }
    };
    int cl = 0, cr = 0;
    for (int base = beg; base < end; base += 32) {
        bool fl, fr;
        flags(base + lane, fl, fr);
        cl += __popc(__ballot_sync(0xffffffffu, fl));
        cr += __popc(__ballot_sync(0xffffffffu, fr));
    }
    if (lane == 0) { sh.wl[warp] = cl; sh.wr[warp] = cr; }
    if (tid == 0) sh.m = 0;
    __syncthreads();
    int ol = 0, orr = 0; nl = 0; nr = 0;
    {   // every warp scans the 32 per-warp counts itself
        int a = sh.wl[lane], c = sh.wr[lane];
        int ia = a, ic = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, ia, d), tc = __shfl_up_sync(0xffffffffu, ic, d);
            if (lane >= d) { ia += ta; ic += tc; }
        }
        nl = __shfl_sync(0xffffffffu, ia, 31); nr = __shfl_sync(0xffffffffu, ic, 31);
        ol = __shfl_sync(0xffffffffu, ia - a, warp); orr = __shfl_sync(0xffffffffu, ic - c, warp);
    }
    for (int base = beg; base < end; base += 32) {
        bool fl, fr;
        flags(base + lane, fl, fr);
        const uint32_t ml = __ballot_sync(0xffffffffu, fl), mr = __ballot_sync(0xffffffffu, fr);
        const uint32_t lt = (1u << lane) - 1u;
        if (fl) lpos[ol + __popc(ml & lt)] = (P)(base + lane);
        if (fr) rasc[orr + __popc(mr & lt)] = (P)(base + lane);
        ol += __popc(ml); orr += __popc(mr);
    }
    __syncthreads();
    // m = #{k : lpos[k] < rpos[k]},  rpos[k] = rasc[nr-1-k]
    const int kmax = min(nl, nr);
    int c = 0;
    for (int k = tid; k < kmax; k += SEL_THREADS) c += lpos[k] < rasc[nr - 1 - k];
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) atomicAdd(&sh.m, c);
    __syncthreads();
    const int m = sh.m;
    const int c1 = m < nl ? (int)lpos[m] : 0x7fffffff;
    const int c2 = m >= 1 ? (int)rasc[nr - m] : 0x7fffffff;
    cut = min(c1, c2);
    for (int k = tid; k < m; k += SEL_THREADS) e.swap((int)lpos[k], (int)rasc[nr - 1 - k]);
    __syncthreads();
}

// libstdc++ __adjust_heap + __push_heap with comp = greater (min-heap on response)
template <class A>
__device__ void adjust_heap(const A &e, int first, int hole, int len, Item v)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (e.k(first + child) > e.k(first + child - 1)) --child;
        e.copy(first + hole, first + child);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        e.copy(first + hole, first + child - 1);
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && e.k(first + parent) > v.k) {
        e.copy(first + hole, first + parent);
        hole = parent;
        parent = (hole - 1) / 2;
    }
    e.put(first + hole, v);
}

template <class A>
__device__ void heap_select(const A &e, int first, int middle, int last)
{
    const int len = middle - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            adjust_heap(e, first, parent, len, e.get(first + parent));
            if (parent == 0) break;
            --parent;
        }
    }
    for (int i = middle; i < last; ++i)
        if (e.k(i) > e.k(first)) {
            const Item v = e.get(i);
            e.copy(i, first);
            adjust_heap(e, first, 0, len, v);
        }
}

template <class A>
__device__ void move_median_to_first(const A &e, int r, int a, int b, int c)
{
    const float ka = e.k(a), kb = e.k(b), kc = e.k(c);
    if (ka > kb) {
        if (kb > kc) e.swap(r, b);
        else if (ka > kc) e.swap(r, c);
        else e.swap(r, a);
    } else if (ka > kc) e.swap(r, a);
    else if (kb > kc) e.swap(r, c);
    else e.swap(r, b);
}

// Whole retainBest; every thread of the block must call it; returns the kept count.
template <class A, class P>
__device__ int block_retain_best(const A &e, int n, int n_points, int depth_limit, P *lpos, P *rasc, SelSmem &sh, int *status)
{
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    const int nth = n_points - 1;
    int first = 0, last = n;
    if (depth_limit < 0) depth_limit = 2 * (31 - __clz(n));
    bool done = false;
    while (last - first > 3) {
        if (depth_limit == 0) {
            if (threadIdx.x == 0) {
                heap_select(e, first, nth + 1, last);
                e.swap(first, nth);
                if (status) atomicOr(status, SVO_STATUS_DEPTH);
            }
            __syncthreads();
            done = true;
            break;
        }
        --depth_limit;
        if (threadIdx.x == 0) move_median_to_first(e, first, first + 1, first + (last - first) / 2, last - 1);
        __syncthreads();
        const float pivot = e.k(first);
        int nl, nr, cut;
        two_pointer_round<0>(e, first + 1, last, pivot, lpos, rasc, sh, nl, nr, cut);
        if (cut <= nth) first = cut; else last = cut;
    }
    if (!done) {
        if (threadIdx.x == 0) {  // __insertion_sort on <= 3 elements
            for (int i = first + 1; i < last; ++i) {
                const Item v = e.get(i);
                int j = i;
                if (v.k > e.k(first)) {
                    while (j > first) { e.copy(j, j - 1); --j; }
                } else {
                    while (v.k > e.k(j - 1)) { e.copy(j, j - 1); --j; }
                }
                e.put(j, v);
            }
        }
        __syncthreads();
    }
    const float amb = e.k(n_points - 1);
    int nl, nr, cut;
    two_pointer_round<1>(e, n_points, n, amb, lpos, rasc, sh, nl, nr, cut);
    return n_points + nr;
}

// ---- first cull: gather the level's band lists (raster order) and keep 2*quota by FAST score
extern __shared__ __align__(16) uint32_t sel_dyn[];

// shared layout of the first cull: [val: cap u32][lpos: cap u16][rasc: cap u16] = 8 bytes per candidate
__global__ void __launch_bounds__(SEL_THREADS) k_select1(Bufs b, Geom g, int slot0, int smem_cap)
{
    __shared__ SelSmem sh;
    __shared__ int band_base[512];
    const int l = blockIdx.x, slot = slot0 + blockIdx.y;
    const LevelGeom &L = g.lv[l];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int *cnt1 = b.cnt1 + (size_t)slot * SVO_MAX_LEVELS;
    int *kept1 = b.kept1 + (size_t)slot * SVO_MAX_LEVELS;
    if (L.nbands == 0) {
        if (tid == 0) { cnt1[l] = 0; kept1[l] = 0; }
        return;
    }
    const int *bc = b.bandcnt + (size_t)slot * g.bandcnt_total + L.bandcnt_off;
    if (warp == 0) {  // exclusive scan of the band counts
        int run = 0;
        for (int base = 0; base < L.nbands; base += 32) {
            const int i = base + lane;
            const int c = i < L.nbands ? bc[i] : 0;
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (i < L.nbands) band_base[i] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) band_base[L.nbands] = run;
    }
    __syncthreads();
    const int n = band_base[L.nbands];
    uint32_t *gval = b.cval + (size_t)slot * g.cand_total + L.cand_off;
    // the replay is a chain of short dependent phases: with the arrays in shared memory a phase costs a
    // shared-memory round trip instead of an L2 one (falls back to global memory when n exceeds the carve-out)
    const bool in_smem = n <= smem_cap;
    uint32_t *val = in_smem ? sel_dyn : gval;
    const uint32_t *bands = b.bands + (size_t)slot * g.band_total + L.band_off;
    for (int bi = warp; bi < L.nbands; bi += SEL_WARPS) {
        const int c = bc[bi], o = band_base[bi];
        const uint32_t *src = bands + (size_t)bi * L.band_cap;
        for (int i = lane; i < c; i += 32) val[o + i] = src[i];
    }
    __syncthreads();
    const ScoreInVal e{val};
    int kept;
    if (in_smem) {
        uint16_t *lpos = reinterpret_cast<uint16_t *>(sel_dyn + smem_cap), *rasc = lpos + smem_cap;
        kept = block_retain_best(e, n, 2 * L.quota, -1, lpos, rasc, sh, b.status + slot);
    } else {
        kept = block_retain_best(e, n, 2 * L.quota, -1, b.lpos + (size_t)slot * g.cand_total + L.cand_off,
                                 b.rpos + (size_t)slot * g.cand_total + L.cand_off, sh, b.status + slot);
    }
    if (tid == 0) { cnt1[l] = n; kept1[l] = kept; }
    if (in_smem) {
        __syncthreads();
        for (int i = tid; i < kept; i += SEL_THREADS) gval[i] = val[i];
    }
}

// ---- second cull: keep quota by Harris response.  shared: [key: cap f32][val: cap u32][lpos: cap u16][rasc: cap u16]
__global__ void __launch_bounds__(SEL_THREADS) k_select2(Bufs b, Geom g, int slot0, int smem_cap)
{
    __shared__ SelSmem sh;
    const int l = blockIdx.x, slot = slot0 + blockIdx.y;
    const LevelGeom &L = g.lv[l];
    const int n = min(b.kept1[(size_t)slot * SVO_MAX_LEVELS + l], L.cap2);
    float *gkey = b.key2 + (size_t)slot * g.total2 + L.off2;
    uint32_t *gval = b.val2 + (size_t)slot * g.total2 + L.off2;
    const bool in_smem = n <= smem_cap;
    int kept;
    if (in_smem) {
        const KeyVal e{reinterpret_cast<float *>(sel_dyn), sel_dyn + smem_cap};
        uint16_t *lpos = reinterpret_cast<uint16_t *>(sel_dyn + 2 * smem_cap), *rasc = lpos + smem_cap;
        for (int i = threadIdx.x; i < n; i += SEL_THREADS) { e.key[i] = gkey[i]; e.val[i] = gval[i]; }
        __syncthreads();
        kept = block_retain_best(e, n, L.quota, -1, lpos, rasc, sh, b.status + slot);
        __syncthreads();
        for (int i = threadIdx.x; i < kept; i += SEL_THREADS) { gkey[i] = e.key[i]; gval[i] = e.val[i]; }
    } else {
        const KeyVal e{gkey, gval};
        kept = block_retain_best(e, n, L.quota, -1, b.lpos + (size_t)slot * g.cand_total + L.cand_off,
                                 b.rpos + (size_t)slot * g.cand_total + L.cand_off, sh, b.status + slot);
    }
    if (threadIdx.x == 0) b.kept2[(size_t)slot * SVO_MAX_LEVELS + l] = kept;
}

__global__ void __launch_bounds__(SEL_THREADS) k_retain_best_raw(float *key, uint32_t *val, int n, int n_points,
                                                                 int depth_limit, uint32_t *lpos, uint32_t *rasc,
                                                                 int *kept_out, int *status)
{
    __shared__ SelSmem sh;
    const KeyVal e{key, val};
    const int kept = block_retain_best(e, n, n_points, depth_limit, lpos, rasc, sh, status);
    if (threadIdx.x == 0) *kept_out = kept;
}

#define SEL_SMEM_MAX 12000   // candidates held in shared memory: 8 B each in the first cull (94 KB), 12 B in the second

static int sel_cap1(const Geom &g)
{
    int mx = 0;
    for (int l = 0; l < g.nlevels; ++l) mx = g.lv[l].cand_cap > mx ? g.lv[l].cand_cap : mx;
    return mx < SEL_SMEM_MAX ? mx : SEL_SMEM_MAX;
}
static int sel_cap2(const Geom &g)
{
    int mx = 0;
    for (int l = 0; l < g.nlevels; ++l) mx = g.lv[l].cap2 > mx ? g.lv[l].cap2 : mx;
    return mx < SEL_SMEM_MAX ? mx : SEL_SMEM_MAX;
}

int setup_select_attributes()
{
    if (cudaFuncSetAttribute(k_select1, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SEL_SMEM_MAX) != cudaSuccess) return 1;
    return (int)cudaFuncSetAttribute(k_select2, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * SEL_SMEM_MAX);
}

void launch_select1(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    dim3 grid(g.nlevels, nimg);
    const int cap = sel_cap1(g);
    k_select1<<<grid, SEL_THREADS, (size_t)8 * cap, st>>>(b, g, slot0, cap);
    ++*launches;
}

void launch_select2(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    dim3 grid(g.nlevels, nimg);
    const int cap = sel_cap2(g);
    k_select2<<<grid, SEL_THREADS, (size_t)12 * cap, st>>>(b, g, slot0, cap);
    ++*launches;
}

void launch_retain_best_raw(float *key, uint32_t *val, int n, int n_points, int depth_limit, uint32_t *lpos,
                            uint32_t *rpos, int *kept_out, int *status, cudaStream_t st, long long *launches)
{
    k_retain_best_raw<<<1, SEL_THREADS, 0, st>>>(key, val, n, n_points, depth_limit, lpos, rpos, kept_out, status);
    ++*launches;
}
