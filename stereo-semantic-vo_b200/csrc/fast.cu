// fast.cu — FAST-9/16 (threshold t, 3x3 strict non-max suppression) on every pyramid
// level, restricted to cv::ORB's 31-px keypoint box, emitted in raster order.
// SURVEY.md A.3; runs inside cv::ORB::detectAndCompute (src/frame.cc:75-79).
//
// One CTA per (image, level, band of SVO_FAST_BAND output rows).  The band's pixel rows
// (+4 rows of halo on each side) are staged into shared memory with 128-bit loads.  A cheap
// necessary test runs on every pixel and compacts the survivors into a shared candidate list;
// the exact score then runs on densely packed candidates (no intra-warp divergence), corners
// are suppressed against their 8 neighbours into a bitmap, and the bitmap is emitted in raster
// order so the downstream retainBest replay sees exactly the order cv::FAST produces.
// Algorithmic bytes: each level pixel read once (+ (8/SVO_FAST_BAND) halo re-read),
// 4 B written per corner.
#include "svo_internal.cuh"

#define FAST_THREADS 256

// Necessary condition for a FAST-9 corner, cheap first: any 9-arc of the 16-ring holds at
// least one pixel of every antipodal pair, so "brighter" needs d[k] > t or d[k+8] > t for every
// k (and likewise "darker").  Four pairs (8 ring pixels) are tested here; survivors go to the
// full score.  Returns 1 when the pixel may still be a corner.
__device__ __forceinline__ int fast_quick(const uint8_t *p, int sp, int t)
{
    const int v = p[0];
    const int d0 = v - p[3 * sp], d8 = v - p[-3 * sp], d4 = v - p[3], d12 = v - p[-3];
    bool pb = (d0 > t || d8 > t) && (d4 > t || d12 > t);
    bool pd = (d0 < -t || d8 < -t) && (d4 < -t || d12 < -t);
    if (!pb && !pd) return 0;
    const int d2 = v - p[2 * sp + 2], d10 = v - p[-2 * sp - 2], d6 = v - p[-2 * sp + 2], d14 = v - p[2 * sp - 2];
    pb = pb && (d2 > t || d10 > t) && (d6 > t || d14 > t);
    pd = pd && (d2 < -t || d10 < -t) && (d6 < -t || d14 < -t);
    return (pb || pd) ? 1 : 0;
}

// Exact score: max over the 16 arcs of 9 contiguous ring pixels of min(v - p) (brighter) and of
// min(p - v) (darker), minus 1; 0 unless >= t.  Sliding 9-minimum as min3 of min3 (VIMNMX3).
__device__ __forceinline__ int fast_score(const uint8_t *p, int sp, int t)
{
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * sp];       d[1] = v - p[3 * sp + 1];   d[2] = v - p[2 * sp + 2];   d[3] = v - p[sp + 3];
    d[4] = v - p[3];            d[5] = v - p[-sp + 3];      d[6] = v - p[-2 * sp + 2];  d[7] = v - p[-3 * sp + 1];
    d[8] = v - p[-3 * sp];      d[9] = v - p[-3 * sp - 1];  d[10] = v - p[-2 * sp - 2]; d[11] = v - p[-sp - 3];
    d[12] = v - p[-3];          d[13] = v - p[sp - 3];      d[14] = v - p[2 * sp - 2];  d[15] = v - p[3 * sp - 1];
    bool pb = true, pd = true;   // all 8 antipodal pairs: exact-score only the polarity that can still win
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        pb = pb && (d[k] > t || d[k + 8] > t);
        pd = pd && (d[k] < -t || d[k + 8] < -t);
    }
    int best = 0;
    if (pb) {  // max over arcs of min d (centre brighter)
        int a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = __vimin3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        int m = -256;
#pragma unroll
        for (int k = 0; k < 16; ++k) m = max(m, __vimin3_s32(a[k], a[(k + 3) & 15], a[(k + 6) & 15]));
        best = m;
    }
    if (pd) {  // max over arcs of min -d (centre darker)
        int a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = __vimax3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        int m = 256;
#pragma unroll
        for (int k = 0; k < 16; ++k) m = min(m, __vimax3_s32(a[k], a[(k + 3) & 15], a[(k + 6) & 15]));
        best = max(best, -m);
    }
    const int s = best - 1;
    return s >= t ? s : 0;
}

extern __shared__ __align__(16) uint8_t fast_smem[];

// shared-memory carve-up for a level of pitch sp (bytes)
__host__ __device__ inline int fast_off_sc(int sp) { return (SVO_FAST_BAND + 8) * sp; }
__host__ __device__ inline int fast_off_cand(int sp) { return fast_off_sc(sp) + (SVO_FAST_BAND + 2) * sp; }
__host__ __device__ inline int fast_off_mask(int sp) { return fast_off_cand(sp) + 2 * (SVO_FAST_BAND + 2) * sp + 512; }
__host__ __device__ inline int fast_mask_words(int sp) { return SVO_FAST_BAND * ((sp + 31) / 32); }

__global__ void __launch_bounds__(FAST_THREADS) k_fast(Bufs b, Geom g, int slot0)
{
    int band = blockIdx.x, l = 0;
    while (band >= g.lv[l].nbands) { band -= g.lv[l].nbands; ++l; }
    const LevelGeom &L = g.lv[l];
    const int slot = slot0 + blockIdx.y;
    const int sp = L.pitch;
    const int yb = L.y0 + band * SVO_FAST_BAND;
    const int ye = min(yb + SVO_FAST_BAND, L.y1);
    const int nrow = ye - yb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = FAST_THREADS / 32;
    const int t = g.fast_threshold;
    uint8_t *pix = fast_smem;                                                    // rows yb-4 .. ye+3
    uint8_t *sc = fast_smem + fast_off_sc(sp);                                   // rows yb-1 .. ye
    uint16_t *cand = reinterpret_cast<uint16_t *>(fast_smem + fast_off_cand(sp)); // candidate positions in sc
    uint32_t *mask = reinterpret_cast<uint32_t *>(fast_smem + fast_off_mask(sp)); // kept-corner bitmap, row major
    const int wv = L.x1 - L.x0;
    const int wpr = (wv + 31) >> 5;   // bitmap words per row
    __shared__ int wsum[FAST_THREADS / 32];

    {   // stage pixel rows with 128-bit loads; clear the score tile and the bitmap
        const uint4 *src = reinterpret_cast<const uint4 *>(b.pyr + (size_t)slot * g.pyr_bytes + L.off + (size_t)(yb - 4) * sp);
        uint4 *dst = reinterpret_cast<uint4 *>(pix);
        const int n16 = ((nrow + 8) * sp) >> 4;
        for (int i = tid; i < n16; i += FAST_THREADS) dst[i] = src[i];
        uint4 *z = reinterpret_cast<uint4 *>(sc);
        const int z16 = ((nrow + 2) * sp) >> 4;
        for (int i = tid; i < z16; i += FAST_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < nrow * wpr; i += FAST_THREADS) mask[i] = 0;
    }
    __syncthreads();
    // A. cheap necessary test on every pixel of rows yb-1..ye, columns x0-1..x1, in 32-pixel chunks
    //    dealt round-robin to the warps; each warp appends survivors to its OWN candidate list (no
    //    atomics, no block barrier) and then
    // B. exact-scores its own candidates, densely packed over the lanes.
    const int sw = L.x1 - L.x0 + 2;
    const int cpr = (sw + 31) >> 5;
    const int nchunks = (nrow + 2) * cpr;
    const int wcap = ((nchunks + nwarps - 1) / nwarps) << 5;
    uint16_t *mine = cand + warp * wcap;
    int nmine = 0;
    for (int c = warp; c < nchunks; c += nwarps) {
        const int r = c / cpr, x = L.x0 - 1 + ((c - r * cpr) << 5) + lane;
        const bool ok = x < L.x1 + 1 && fast_quick(pix + (r + 3) * sp + x, sp, t);
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (ok) mine[nmine + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(r * sp + x);
        nmine += __popc(m);
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
        if (s > q[-1] && s > q[1] && s > q[-sp - 1] && s > q[-sp] && s > q[-sp + 1] &&
            s > q[sp - 1] && s > q[sp] && s > q[sp + 1])
            atomicOr(&mask[(r - 1) * wpr + ((x - L.x0) >> 5)], 1u << ((x - L.x0) & 31));
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
    return fast_off_mask(mx) + 4 * fast_mask_words(mx) + 16;
}

void launch_fast(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    if (g.fast_bands == 0) return;
    dim3 grid(g.fast_bands, nimg);
    k_fast<<<grid, FAST_THREADS, fast_smem_bytes(g), st>>>(b, g, slot0);
    ++*launches;
}

int setup_fast_attributes(const Geom &g)
{
    return (int)cudaFuncSetAttribute(k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, fast_smem_bytes(g));
}
