// octree.cu — opt-in quadtree ("octree") keypoint distribution (svo_config.distribution = SVO_DIST_OCTREE).
//
// north_star names "FAST keypoints with grid/octree distribution"; the reference has no such stage
// (src/frame.cc:75-79 runs cv::ORB, whose selection is KeyPointsFilter::retainBest — select.cu), so this
// mode is NOT on the parity path: it follows the published algorithm of ORB-SLAM2's
// ORBextractor::DistributeOctTree with its two accidental orders fixed, and is
// defined by oracle/svo_octree_oracle.c, against which it is exact.
//
// The CPU algorithm splits boxes and re-buckets point lists.  Here a point's whole descent is computed once:
// its path code = root index, then one 2-bit child digit per level down to depth 12 (boxes halve with ceil,
// like ExtractorNode::DivideNode).  After sorting the points by code, every node of every possible tree is a
// RUN of the array, so the tree state is just one depth per point:
//   point i starts a node        <=>  i == 0 or shared(i) < depth[i]     (shared = digits it has in common with i-1)
//   a node holds a single point  <=>  it starts a node and so does i+1
//   splitting a node             <=>  depth += 1 on its run; it gains one node per point with shared == old depth
// A full round is therefore one data-parallel pass and a block-wide sum; the one-at-a-time phase (most
// populated node first until the count reaches N) is a sort of the candidate nodes, a prefix sum of their gains
// and a cut.  One thread block per (level, image); tools/model_octree.py is the executable model of this file.
#include "svo_internal.cuh"

#define OCT_THREADS 1024
#define OCT_WARPS (OCT_THREADS / 32)
#define OCT_MAXD 12
#define OCT_SMEM_CAP 12288     // points held in shared memory (beyond: the level's global scratch arrays)
#define OCT_NODE_CAP 4096      // nodes; the level quota + 3 must fit

struct OctSh {
    int red[2][OCT_WARPS];
    int wcnt[OCT_WARPS];
    int band_base[513];
    int ncand, cut, newnn;
};

extern __shared__ __align__(16) uint32_t oct_dyn[];

// path code of a point at rectangle-relative (xr, yr): root << 24 | digit d at bits 23-2d..22-2d
__device__ __forceinline__ uint32_t oct_code(int xr, int yr, const LevelGeom &L)
{
    int r = (int)__fdiv_rn((float)xr, L.oct_hx);
    r = min(max(r, 0), L.oct_nini - 1);
    int ulx = (int)__fmul_rn(L.oct_hx, (float)r), urx = (int)__fmul_rn(L.oct_hx, (float)(r + 1));
    int uly = 0, bry = L.y1 - L.y0;
    uint32_t code = (uint32_t)r << 24;
#pragma unroll
    for (int d = 0; d < OCT_MAXD; ++d) {
        const int sx = ulx + ((urx - ulx + 1) >> 1), sy = uly + ((bry - uly + 1) >> 1);
        uint32_t q = 0;
        if (xr < sx) urx = sx; else { ulx = sx; q |= 1u; }
        if (yr < sy) bry = sy; else { uly = sy; q |= 2u; }
        code |= q << (22 - 2 * d);
    }
    return code;
}

// digits two consecutive sorted codes share: -1 = different roots, OCT_MAXD = equal codes
__device__ __forceinline__ int oct_shared(uint32_t a, uint32_t b)
{
    const uint32_t x = a ^ b;
    if (x >> 24) return -1;
    if (x == 0) return OCT_MAXD;
    return (23 - (31 - __clz(x))) >> 1;
}

__device__ __forceinline__ void oct_sum2(int &a, int &b, OctSh &sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = __reduce_add_sync(0xffffffffu, a); b = __reduce_add_sync(0xffffffffu, b);
    __syncthreads();
    if (lane == 0) { sh.red[0][warp] = a; sh.red[1][warp] = b; }
    __syncthreads();
    a = __reduce_add_sync(0xffffffffu, sh.red[0][lane]); b = __reduce_add_sync(0xffffffffu, sh.red[1][lane]);
}

// Ascending bitonic sort of n (hi, lo) pairs, any n: every merge compares in the same direction (first step
// mirrored), so the virtual +inf padding behind index n never moves and needs no storage.
__device__ void oct_sort(uint32_t *hi, uint32_t *lo, int n)
{
    int npad = 2;
    while (npad < n) npad <<= 1;
    const int pairs = npad >> 1;
    auto cmpx = [&](int i, int p) {
        const uint32_t hi_i = hi[i], hi_p = hi[p], lo_i = lo[i], lo_p = lo[p];
        if (hi_i > hi_p || (hi_i == hi_p && lo_i > lo_p)) { hi[i] = hi_p; hi[p] = hi_i; lo[i] = lo_p; lo[p] = lo_i; }
    };
    for (int k = 2, lk = 1; k <= npad; k <<= 1, ++lk) {
        const int hk = k >> 1;
        for (int t = threadIdx.x; t < pairs; t += OCT_THREADS) {
            const int blk = t >> (lk - 1), off = t & (hk - 1);
            const int i = (blk << lk) + off, p = (blk << lk) + k - 1 - off;
            if (p < n) cmpx(i, p);
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < pairs; t += OCT_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i + j;
                if (p < n) cmpx(i, p);
            }
            __syncthreads();
        }
    }
}

// ordered compaction of the node starts into ns[0..nn), ns[nn] = n; returns nn (the same in every thread)
__device__ int oct_starts(const int8_t *sh8, const uint8_t *dep, int n, int *ns, OctSh &sh)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = (((n + OCT_WARPS - 1) / OCT_WARPS) + 31) & ~31;
    const int beg = warp * seg, end = min(beg + seg, n);
    int c = 0;
    for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        const bool st = i < end && (i == 0 || (int)sh8[i] < (int)dep[i]);
        c += __popc(__ballot_sync(0xffffffffu, st));
    }
    __syncthreads();
    if (lane == 0) sh.wcnt[warp] = c;
    __syncthreads();
    int v = sh.wcnt[lane], inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    const int nn = __shfl_sync(0xffffffffu, inc, 31);
    int o = __shfl_sync(0xffffffffu, inc - v, warp);
    for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        const bool st = i < end && (i == 0 || (int)sh8[i] < (int)dep[i]);
        const uint32_t m = __ballot_sync(0xffffffffu, st);
        const int pos = o + __popc(m & ((1u << lane) - 1u));
        if (st && pos < OCT_NODE_CAP) ns[pos] = i;
        o += __popc(m);
    }
    if (tid == 0) ns[min(nn, OCT_NODE_CAP)] = n;
    __syncthreads();
    return nn;
}

__global__ void __launch_bounds__(OCT_THREADS) k_octree(Bufs b, Geom g, int slot0)
{
    __shared__ OctSh sh;
    const int l = blockIdx.x, slot = slot0 + blockIdx.y;
    const LevelGeom &L = g.lv[l];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int *cnt1 = b.cnt1 + (size_t)slot * SVO_MAX_LEVELS;
    int *kept1 = b.kept1 + (size_t)slot * SVO_MAX_LEVELS;
    if (L.nbands == 0) {
        if (tid == 0) { cnt1[l] = 0; kept1[l] = 0; }
        return;
    }
    // ---- gather the level's band lists (raster order) -------------------------------------------------------
    const int *bc = b.bandcnt + (size_t)slot * g.bandcnt_total + L.bandcnt_off;
    if (warp == 0) {
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
            if (i < L.nbands) sh.band_base[i] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) sh.band_base[L.nbands] = run;
    }
    __syncthreads();
    const int n = sh.band_base[L.nbands];
    float *gkey = b.ckey + (size_t)slot * g.cand_total + L.cand_off;
    uint32_t *gval = b.cval + (size_t)slot * g.cand_total + L.cand_off;
    if (n == 0) {
        if (tid == 0) { cnt1[l] = 0; kept1[l] = 0; }
        return;
    }
    const bool in_smem = n <= OCT_SMEM_CAP;
    const int nal = (n + 15) & ~15;
    uint32_t *hi = in_smem ? oct_dyn : b.lpos + (size_t)slot * g.cand_total + L.cand_off;
    uint32_t *lo = in_smem ? oct_dyn + OCT_SMEM_CAP : b.rpos + (size_t)slot * g.cand_total + L.cand_off;
    uint8_t *bytes = in_smem ? reinterpret_cast<uint8_t *>(oct_dyn + 2 * OCT_SMEM_CAP) : reinterpret_cast<uint8_t *>(gkey);
    const int bstride = in_smem ? OCT_SMEM_CAP : nal;       // 3 * nal <= 4 * cand_cap: n <= cand_cap - 4
    int8_t *sh8 = reinterpret_cast<int8_t *>(bytes);
    uint8_t *dep = bytes + bstride, *dep2 = bytes + 2 * bstride;
    int *ns = reinterpret_cast<int *>(oct_dyn + 2 * OCT_SMEM_CAP + 3 * (OCT_SMEM_CAP / 4));
    uint32_t *chi = reinterpret_cast<uint32_t *>(ns + OCT_NODE_CAP + 4), *clo = chi + OCT_NODE_CAP;
    const uint32_t *bands = b.bands + (size_t)slot * g.band_total + L.band_off;
    for (int bi = warp; bi < L.nbands; bi += OCT_WARPS) {
        const int c = bc[bi], o = sh.band_base[bi];
        const uint32_t *src = bands + (size_t)bi * L.band_cap;
        for (int i = lane; i < c; i += 32) {
            const uint32_t e = src[i];
            hi[o + i] = oct_code(unpack_x(e) - L.x0, unpack_y(e) - L.y0, L);
            lo[o + i] = e;
        }
    }
    __syncthreads();
    oct_sort(hi, lo, n);
    for (int i = tid; i < n; i += OCT_THREADS) {
        sh8[i] = (int8_t)(i == 0 ? -1 : oct_shared(hi[i - 1], hi[i]));
        dep[i] = 0;
    }
    __syncthreads();
    const int N = L.quota;
    // nodes at the start = non-empty roots
    int nn = 0, ne = 0;
    for (int i = tid; i < n; i += OCT_THREADS) nn += i == 0 || sh8[i] < 0;
    oct_sum2(nn, ne, sh);
    // ---- rounds ----------------------------------------------------------------------------------------------
    bool finish = false;
    while (!finish) {
        const int prev = nn;
        nn = 0; ne = 0;
        for (int i = tid; i < n; i += OCT_THREADS) {
            const int d0 = dep[i];
            const int d1 = i + 1 < n ? dep[i + 1] : 0, d2 = i + 2 < n ? dep[i + 2] : 0;
            const bool s0 = i == 0 || (int)sh8[i] < d0;
            const bool s1 = i + 1 >= n || (int)sh8[i + 1] < d1;
            const bool s2 = i + 2 >= n || (int)sh8[i + 2] < d2;
            const int e0 = d0 + ((!(s0 && s1) && d0 < OCT_MAXD) ? 1 : 0);
            const int e1 = d1 + ((!(s1 && s2) && d1 < OCT_MAXD) ? 1 : 0);
            const bool t0 = i == 0 || (int)sh8[i] < e0;
            const bool t1 = i + 1 >= n || (int)sh8[i + 1] < e1;
            dep2[i] = (uint8_t)e0;
            nn += t0;
            ne += t0 && !t1 && e0 < OCT_MAXD;
        }
        oct_sum2(nn, ne, sh);          // (its barriers also order the dep2 writes before the next reads)
        { uint8_t *t = dep; dep = dep2; dep2 = t; }
        if (nn >= N || nn == prev) break;
        if (nn + 3 * ne <= N) continue;
        // ---- one node at a time, most populated first -------------------------------------------------------
        while (!finish) {
            const int prev2 = nn;
            if (tid == 0) { sh.ncand = 0; sh.cut = 0x7fffffff; sh.newnn = nn; }
            oct_starts(sh8, dep, n, ns, sh);
            const int nk = min(nn, OCT_NODE_CAP);
            for (int k = warp; k < nk; k += OCT_WARPS) {
                const int beg = ns[k], end = ns[k + 1];
                const int d = dep[beg];
                uint32_t key_hi = 0xffffffffu, gain = 0;
                if (end - beg > 1 && d < OCT_MAXD) {
                    int c = 0;
                    for (int i = beg + 1 + lane; i < end; i += 32) c += (int)sh8[i] == d;
                    gain = (uint32_t)__reduce_add_sync(0xffffffffu, c);
                    key_hi = 0xffffffffu - (uint32_t)(end - beg);
                    if (lane == 0) atomicAdd(&sh.ncand, 1);
                }
                if (lane == 0) { chi[k] = key_hi; clo[k] = ((uint32_t)k << 2) | gain; }
            }
            __syncthreads();
            oct_sort(chi, clo, nk);
            const int ncand = sh.ncand;
            // inclusive prefix of the gains in sorted order, 4 candidates per thread
            int inc4[4], s = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = 4 * tid + q;
                s += c < ncand ? (int)(clo[c] & 3u) : 0;
                inc4[q] = s;
            }
            int winc = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            __syncthreads();
            if (lane == 31) sh.wcnt[warp] = winc;
            __syncthreads();
            int base = winc - s;
            for (int w = 0; w < warp; ++w) base += sh.wcnt[w];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = 4 * tid + q;
                if (c < ncand && nn + base + inc4[q] >= N) { atomicMin(&sh.cut, c); break; }
            }
            __syncthreads();
            const int cut = min(sh.cut, ncand - 1);     // every candidate when the count never reaches N
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * tid + q == cut) sh.newnn = nn + base + inc4[q];
            for (int c = warp; c <= cut; c += OCT_WARPS) {
                const int k = (int)(clo[c] >> 2);
                for (int i = ns[k] + lane; i < ns[k + 1]; i += 32) dep[i] = (uint8_t)(dep[i] + 1);
            }
            __syncthreads();
            nn = sh.newnn;
            if (nn >= N || nn == prev2) finish = true;
            __syncthreads();
        }
    }
    // ---- each node keeps its best-scoring point (ties: first in raster order); nodes leave in Z order --------
    const int total = oct_starts(sh8, dep, n, ns, sh);
    const int kept = min(total, min(OCT_NODE_CAP, L.cap2));
    if (tid == 0) {
        cnt1[l] = n; kept1[l] = kept;
        if (total > kept) atomicOr(b.status + slot, SVO_STATUS_OVERFLOW);
    }
    for (int k = warp; k < kept; k += OCT_WARPS) {
        uint32_t best = 0;
        for (int i = ns[k] + lane; i < ns[k + 1]; i += 32) {
            const uint32_t e = lo[i];
            best = max(best, (e & 0xff000000u) | (0x00ffffffu - (e & 0x00ffffffu)));
        }
        best = __reduce_max_sync(0xffffffffu, best);
        if (lane == 0) {
            // in the global-scratch case gkey aliases sh8/dep, which nobody reads any more (oct_starts ended with a barrier)
            gval[k] = (best & 0xff000000u) | (0x00ffffffu - (best & 0x00ffffffu));
            gkey[k] = (float)(best >> 24);
        }
    }
}

// second cull of the octree mode: everything the distribution kept stays, in node order
__global__ void k_keep_all(Bufs b, Geom g, int slot0)
{
    const int slot = slot0 + blockIdx.x, l = threadIdx.x;
    if (l < g.nlevels)
        b.kept2[(size_t)slot * SVO_MAX_LEVELS + l] = min(b.kept1[(size_t)slot * SVO_MAX_LEVELS + l], g.lv[l].cap2);
}

static size_t oct_smem_bytes()
{
    return sizeof(uint32_t) * (2 * OCT_SMEM_CAP + 3 * (OCT_SMEM_CAP / 4)) + sizeof(int) * (OCT_NODE_CAP + 4) +
           sizeof(uint32_t) * 2 * OCT_NODE_CAP;
}

int setup_octree_attributes()
{
    return (int)cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oct_smem_bytes());
}

int octree_max_quota() { return OCT_NODE_CAP - 3; }

void launch_octree(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    dim3 grid(g.nlevels, nimg);
    k_octree<<<grid, OCT_THREADS, oct_smem_bytes(), st>>>(b, g, slot0);
    ++*launches;
}

void launch_keep_all(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    k_keep_all<<<nimg, 32, 0, st>>>(b, g, slot0);
    ++*launches;
}
