// fast.cu — FAST-9/16 (threshold t, 3x3 strict non-max suppression) on every pyramid
// level, restricted to cv::ORB's 31-px keypoint box, emitted in raster order.
// SURVEY.md A.3; runs inside cv::ORB::detectAndCompute (src/frame.cc:75-79).
//
// One CTA per (image, level, band of Geom.fast_band output rows).  The band's pixel rows
// (+4 rows of halo on each side) are staged into shared memory with 128-bit loads.  A cheap
// necessary test runs on every pixel and compacts the survivors into a shared candidate list;
// the exact score then runs on densely packed candidates (no intra-warp divergence), corners
// are suppressed against their 8 neighbours into a bitmap, and the bitmap is emitted in raster
// order so the downstream retainBest replay sees exactly the order cv::FAST produces.
// Algorithmic bytes: each level pixel read once (+ (8/fast_band) halo re-read),
// 4 B written per corner.
#include "svo_internal.cuh"
#include <cuda/barrier>

#define FAST_THREADS 512

// ---- stage A: necessary condition on 4 horizontally adjacent pixels at once ----
// Any 9-arc of the 16-ring holds at least one pixel of every antipodal pair, so a "centre brighter"
// corner needs min(p[k], p[k+8]) < v - t for every k, and a "centre darker" corner max(p[k], p[k+8]) > v + t
// for every k.  Four of the eight pairs are tested (N/S, E/W and the two diagonals):
//     bright  <=>  max over pairs of min(pair) + t < v          dark  <=>  min over pairs of max(pair) > v + t
// The pixels are widened to u16x2 (PRMT) so the min/max trees run on the packed DPX instructions
// (VIMNMX3.U16x2, two pixels per instruction) and the two final compares are packed subtractions whose
// bit 15 cannot borrow across halves.  (An earlier all-byte SWAR version built every compare from LOP3/IADD —
// sm_100a emulates the byte-wise video compares — and cost 2.2x the instructions; ncu: 167 -> 76 per quad.)
#define SW_H 0x80808080u
#define U16_H 0x80008000u

// two pixels (one u16x2 half of the quad): bit 15 of a half is set when that pixel passes
__device__ __forceinline__ uint32_t fast_quick2(uint32_t v, uint32_t n0, uint32_t n8, uint32_t n4, uint32_t n12,
                                                uint32_t n2, uint32_t n10, uint32_t n6, uint32_t n14, uint32_t t1)
{
    const uint32_t mn = __vmaxu2(__vimax3_u16x2(__vminu2(n0, n8), __vminu2(n4, n12), __vminu2(n2, n10)), __vminu2(n6, n14));
    const uint32_t mx = __vminu2(__vimin3_u16x2(__vmaxu2(n0, n8), __vmaxu2(n4, n12), __vmaxu2(n2, n10)), __vmaxu2(n6, n14));
    // halves stay below 2^15, so (a | H) - b keeps bit 15 exactly when a >= b, independently per half
    const uint32_t bright = (v | U16_H) - (mn + t1);      // v >= maxmin + t + 1
    const uint32_t dark = (mx | U16_H) - (v + t1);        // minmax >= v + t + 1
    return (bright | dark) & U16_H;
}
// q: address of the quad's first pixel in the staged rows (4-byte aligned); returns bit 7 flags; t1 = (t + 1) * 0x10001
__device__ __forceinline__ uint32_t fast_quick4(const uint8_t *q, int sp, uint32_t t1)
{
    const uint32_t *c = reinterpret_cast<const uint32_t *>(q);
    const int sw = sp >> 2;
    const uint32_t v = c[0];
    const uint32_t r0 = c[3 * sw], r8 = c[-3 * sw];                                     // ring 0 / 8
    const uint32_t wl = c[-1], wr = c[1];
    const uint32_t r4 = __byte_perm(v, wr, 0x6543), r12 = __byte_perm(wl, v, 0x4321);   // ring 4 (x+3) / 12 (x-3)
    const uint32_t al = c[2 * sw - 1], a0 = c[2 * sw], ar = c[2 * sw + 1];
    const uint32_t bl = c[-2 * sw - 1], b0 = c[-2 * sw], br = c[-2 * sw + 1];
    const uint32_t r2 = __byte_perm(a0, ar, 0x5432), r10 = __byte_perm(bl, b0, 0x5432); // ring 2 (+2,+2) / 10 (-2,-2)
    const uint32_t r6 = __byte_perm(b0, br, 0x5432), r14 = __byte_perm(al, a0, 0x5432); // ring 6 (+2,-2) / 14 (-2,+2)
#define LO2(w) __byte_perm(w, 0, 0x4140)
#define HI2(w) __byte_perm(w, 0, 0x4342)
    const uint32_t f01 = fast_quick2(LO2(v), LO2(r0), LO2(r8), LO2(r4), LO2(r12), LO2(r2), LO2(r10), LO2(r6), LO2(r14), t1);
    const uint32_t f23 = fast_quick2(HI2(v), HI2(r0), HI2(r8), HI2(r4), HI2(r12), HI2(r2), HI2(r10), HI2(r6), HI2(r14), t1);
#undef LO2
#undef HI2
    return __byte_perm(f01, f23, 0x7531) & SW_H;     // bit 15 of each half -> bit 7 of the pixel's byte
}

// Exact score: max over the 16 arcs of 9 contiguous ring pixels of min(v - p) (centre brighter) and of
// min(p - v) (centre darker), minus 1; 0 unless >= t.  With r the ring values, min over an arc of (v - r) is
// v - max over the arc of r, so both polarities come from the arcs' maxima and minima of r itself:
//     bright = v - min over arcs (max over arc r)          dark = max over arcs (min over arc r) - v
// branch-free on packed u16x2 values with the DPX min3/max3 instructions (VIMNMX3.U16x2): register X[k] holds ring
// pixels k (low half) and k + 8 (high half), so one instruction advances two arcs.  Sliding 9-extremum = 3-extremum of
// 3-extrema (windows of 3, then offsets 0/3/6); the eight results per polarity are reduced as a tree.
__device__ __forceinline__ uint32_t swap16(uint32_t x) { return __byte_perm(x, 0, 0x1032); }
__device__ __forceinline__ int fast_score(const uint8_t *p, int sp, int t)
{
    const int v = p[0];
    uint32_t X[16];
#define FAST_E(k, lo, hi) X[k] = (uint32_t)p[lo] + ((uint32_t)p[hi] << 16)
    FAST_E(0, 3 * sp, -3 * sp);          FAST_E(1, 3 * sp + 1, -3 * sp - 1);
    FAST_E(2, 2 * sp + 2, -2 * sp - 2);  FAST_E(3, sp + 3, -sp - 3);
    FAST_E(4, 3, -3);                    FAST_E(5, -sp + 3, sp - 3);
    FAST_E(6, -2 * sp + 2, 2 * sp - 2);  FAST_E(7, -3 * sp + 1, 3 * sp - 1);
#undef FAST_E
#pragma unroll
    for (int k = 0; k < 8; ++k) X[k + 8] = swap16(X[k]);
    uint32_t a[14], b[14];                  // maxima / minima of ring pixels k, k + 1, k + 2
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = __vimax3_u16x2(X[k], X[k + 1], X[k + 2]);
        b[k] = __vimin3_u16x2(X[k], X[k + 1], X[k + 2]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { a[k + 8] = swap16(a[k]); b[k + 8] = swap16(b[k]); }
    uint32_t A[8], B[8];                    // arc k (low half) and arc k + 8 (high half): maximum / minimum of its 9 pixels
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        A[k] = __vimax3_u16x2(a[k], a[k + 3], a[k + 6]);
        B[k] = __vimin3_u16x2(b[k], b[k + 3], b[k + 6]);
    }
    const uint32_t lo2 = __vimin3_u16x2(__vimin3_u16x2(A[0], A[1], A[2]), __vimin3_u16x2(A[3], A[4], A[5]), __vminu2(A[6], A[7]));
    const uint32_t hi2 = __vimax3_u16x2(__vimax3_u16x2(B[0], B[1], B[2]), __vimax3_u16x2(B[3], B[4], B[5]), __vmaxu2(B[6], B[7]));
    const int bright = v - min((int)(lo2 & 0xffffu), (int)(lo2 >> 16));   // max arc-min of (v - p)
    const int dark = max((int)(hi2 & 0xffffu), (int)(hi2 >> 16)) - v;     // max arc-min of (p - v)
    const int s = max(bright, dark) - 1;
    return s >= t ? s : 0;
}

extern __shared__ __align__(128) uint8_t fast_smem[];

// shared-memory carve-up for a level of pitch sp (bytes)
__host__ __device__ inline int fast_off_sc(int sp, int band) { return (band + 8) * sp; }
__host__ __device__ inline int fast_off_cand(int sp, int band) { return fast_off_sc(sp, band) + (band + 2) * sp; }
__host__ __device__ inline int fast_off_mask(int sp, int band)
{
    return fast_off_cand(sp, band) + 2 * ((band + 2) * (sp + 128) + (FAST_THREADS / 32) * 128);   // every warp's share is rounded up to a chunk
}
__host__ __device__ inline int fast_mask_words(int sp, int band) { return band * ((sp + 31) / 32); }

__global__ void __launch_bounds__(FAST_THREADS) k_fast(Bufs b, Geom g, int slot0)
{
    int band = blockIdx.x, l = 0;
    while (band >= g.lv[l].nbands) { band -= g.lv[l].nbands; ++l; }
    const LevelGeom &L = g.lv[l];
    const int slot = slot0 + blockIdx.y;
    const int sp = L.pitch;
    const int FB = g.fast_band;
    const int yb = L.y0 + band * FB;
    const int ye = min(yb + FB, L.y1);
    const int nrow = ye - yb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = FAST_THREADS / 32;
    const int t = g.fast_threshold;
    uint8_t *pix = fast_smem;                                                    // rows yb-4 .. ye+3
    uint8_t *sc = fast_smem + fast_off_sc(sp, FB);                                   // rows yb-1 .. ye
    uint16_t *cand = reinterpret_cast<uint16_t *>(fast_smem + fast_off_cand(sp, FB)); // candidate positions in sc
    uint32_t *mask = reinterpret_cast<uint32_t *>(fast_smem + fast_off_mask(sp, FB)); // kept-corner bitmap, row major
    const int wv = L.x1 - L.x0;
    const int wpr = (wv + 31) >> 5;   // bitmap words per row
    __shared__ int wsum[FAST_THREADS / 32];

    {   // The band's pixel rows are contiguous in HBM (16-byte pitched rows): one TMA bulk copy
        // (cp.async.bulk, completes on an mbarrier) stages them while the threads clear the score tile and the
        // bitmap.
        __shared__ cuda::barrier<cuda::thread_scope_block> bar;
        const uint8_t *src = b.pyr + (size_t)slot * g.pyr_bytes + L.off + (size_t)(yb - 4) * sp;
        const uint32_t bytes = (uint32_t)((nrow + 8) * sp);
        if (tid == 0) {
            init(&bar, 1);
            cuda::device::experimental::fence_proxy_async_shared_cta();
        }
        __syncthreads();
        cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
        if (tid == 0) {
            cuda::device::memcpy_async_tx(pix, src, cuda::aligned_size_t<16>(bytes), bar);
            tok = cuda::device::barrier_arrive_tx(bar, 1, bytes);
        }
        uint4 *z = reinterpret_cast<uint4 *>(sc);
        const int z16 = ((nrow + 2) * sp) >> 4;
        for (int i = tid; i < z16; i += FAST_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < nrow * wpr; i += FAST_THREADS) mask[i] = 0;
        if (tid == 0) bar.wait(std::move(tok));
    }
    __syncthreads();
    // A. SWAR necessary test on every pixel of rows yb-1..ye, columns x0-1..x1, four pixels per thread, in
    //    128-pixel chunks dealt round-robin to the warps; each warp appends survivors to its OWN candidate
    //    list (no atomics, no block barrier), then
    // B. exact-scores its own candidates, densely packed over the lanes.
    const int xq0 = (L.x0 - 1) & ~3;
    const int nq = (L.x1 + 1 - xq0 + 3) >> 2;
    const int cpr = (nq + 31) >> 5;
    const int nchunks = (nrow + 2) * cpr;
    const int wcap = ((nchunks + nwarps - 1) / nwarps) << 7;
    uint16_t *mine = cand + warp * wcap;
    const uint32_t t4 = (uint32_t)(t + 1) * 0x10001u;     // fast_quick4's packed threshold
    int nmine = 0;
    {
        // Survivors are compacted once per 8 chunks: a lane keeps the 4-bit hit masks of its quads in one register (acc) and
        // the chunks' first positions go to a small per-warp table; a flush is one warp prefix sum of the hit counts plus a
        // store per hit (compacting every chunk with three ballots cost 70 instructions per chunk against 105 for the test).
        __shared__ uint16_t cbase[FAST_THREADS / 32][8];
        uint16_t *wb = cbase[warp];
        const int l4 = lane << 2;
        uint32_t acc = 0;
        int jn = 0;
        auto flush = [&]() {
            __syncwarp();
            const int hc = __popc(acc);
            int inc = hc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += v;
            }
            int pos = nmine + inc - hc;
            nmine += __shfl_sync(0xffffffffu, inc, 31);
            uint32_t bits = acc;
            while (bits) {
                const int bi = __ffs((int)bits) - 1;
                bits &= bits - 1;
                mine[pos++] = (uint16_t)(wb[bi >> 2] + l4 + (bi & 3));
            }
            __syncwarp();
            acc = 0; jn = 0;
        };
        int r = 0, ch = warp;
        while (ch >= cpr) { ch -= cpr; ++r; }
        for (int c = warp; c < nchunks; c += nwarps) {
            const int qi = (ch << 5) + lane;
            const int x = xq0 + (qi << 2);
            uint32_t pass = 0;
            // The first and last quad of a row reach up to 3 pixels outside [x0 - 1, x1]; those pixels may become
            // candidates and get a score, which nothing reads: stage C only suppresses and emits x0 <= x < x1, whose
            // neighbours lie inside [x0 - 1, x1] (masking them here cost 12 instructions on every quad).
            if (qi < nq) pass = fast_quick4(pix + (r + 3) * sp + x, sp, t4);
            const uint32_t nib = (((pass >> 7) & 0x01010101u) * 0x10204080u) >> 28;   // 4-bit hit mask of the quad
            acc |= nib << (jn << 2);
            if (lane == 0) wb[jn] = (uint16_t)(r * sp + xq0 + (ch << 7));             // position of the chunk's first pixel
            ch += nwarps;
            while (ch >= cpr) { ch -= cpr; ++r; }
            if (++jn == 8) flush();
        }
        if (jn) flush();
    }
    __syncwarp();
    for (int i = lane; i < nmine; i += 32) {
        const int pos = mine[i];
        const int s = fast_score(pix + pos + 3 * sp, sp, t);
        sc[pos] = (uint8_t)s;
        if (s == 0) mine[i] = 0xffffu;   // not a corner
    }
    __syncthreads();
    // C. 3x3 strict non-max suppression of the corners in the output rows -> bitmap
    for (int i = lane; i < nmine; i += 32) {
        const int pos = mine[i];
        if (pos == 0xffff) continue;
        const int r = __umulhi((uint32_t)pos, L.pitch_magic), x = pos - r * sp;
        if (r < 1 || r > nrow || x < L.x0 || x >= L.x1) continue;
        const uint8_t *q = sc + pos;
        const int s = q[0];
        // branch-free: the largest of the 8 neighbours (a short-circuit chain diverges on every comparison)
        const int m = __vimax3_s32(__vimax3_s32(q[-sp - 1], q[-sp], q[-sp + 1]), __vimax3_s32(q[-1], q[1], q[sp - 1]),
                                   max((int)q[sp], (int)q[sp + 1]));
        if (s > m) atomicOr(&mask[(r - 1) * wpr + ((x - L.x0) >> 5)], 1u << ((x - L.x0) & 31));
    }
    __syncthreads();
    // D. raster-ordered emission from the bitmap (words are in raster order)
    const int nwords = nrow * wpr;
    const int per = (nwords + FAST_THREADS - 1) / FAST_THREADS;
    const int w0 = tid * per, w1 = min(w0 + per, nwords);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(mask[w]);
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int off = inc - cnt, tot = 0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) {
        const int c = wsum[w];
        if (w < warp) off += c;
        tot += c;
    }
    uint32_t *out = b.bands + (size_t)slot * g.band_total + L.band_off + (size_t)band * L.band_cap;
    for (int w = w0; w < w1; ++w) {
        uint32_t m = mask[w];
        const int r = w / wpr, xw = L.x0 + ((w - r * wpr) << 5);
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int x = xw + bit;
            out[off++] = pack_xy(x, yb + r, sc[(r + 1) * sp + x]);
        }
    }
    if (tid == 0) b.bandcnt[(size_t)slot * g.bandcnt_total + L.bandcnt_off + band] = tot;
}

int fast_smem_bytes(const Geom &g)
{
    int mx = 0;
    for (int l = 0; l < g.nlevels; ++l) mx = g.lv[l].pitch > mx ? g.lv[l].pitch : mx;
    return fast_off_mask(mx, g.fast_band) + 4 * fast_mask_words(mx, g.fast_band) + 16;
}

void launch_fast(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    if (g.fast_bands == 0) return;
    dim3 grid(g.fast_bands, nimg);
    k_fast<<<grid, FAST_THREADS, fast_smem_bytes(g), st>>>(b, g, slot0);
    ++*launches;
}

// The attribute belongs to the kernel, not to a context: it only ever grows, so a context created for a small
// image never lowers the limit under one that is already serving a larger geometry.
int setup_fast_attributes(const Geom &g)
{
    static int granted[64] = {0};   // per device (function attributes are per device)
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    const int need = fast_smem_bytes(g);
    if (need <= granted[dev]) return 0;
    const cudaError_t e = cudaFuncSetAttribute(k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    if (e == cudaSuccess) granted[dev] = need;
    return (int)e;
}
