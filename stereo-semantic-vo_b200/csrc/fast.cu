// fast.cu — FAST-9/16 (threshold t, 3x3 strict non-max suppression) on every pyramid
// level, restricted to cv::ORB's 31-px keypoint box, emitted in raster order.
// SURVEY.md A.3; runs inside cv::ORB::detectAndCompute (src/frame.cc:75-79).
//
// One CTA per (image, level, band of SVO_FAST_BAND output rows).  The band's pixel rows
// (+4 rows of halo on each side) are staged into shared memory with 128-bit loads; the
// score of each pixel is computed from the shared tile, suppressed against its 8
// neighbours, and the survivors are compacted in raster order with warp ballots so the
// downstream retainBest replay sees exactly the order cv::FAST produces.
// Algorithmic bytes: each level pixel read once (+ (8/SVO_FAST_BAND) halo re-read),
// 4 B written per corner.
#include "svo_internal.cuh"

#define FAST_THREADS 256

__device__ __forceinline__ int fast_score(const uint8_t *p, int sp, int t)
{
    const int v = p[0];
    const int d0 = v - p[3 * sp], d8 = v - p[-3 * sp], d4 = v - p[3], d12 = v - p[-3];
    // any 9-arc of the 16-ring holds one pixel of every antipodal pair
    const bool pb = (d0 > t || d8 > t) && (d4 > t || d12 > t);
    const bool pd = (d0 < -t || d8 < -t) && (d4 < -t || d12 < -t);
    if (!pb && !pd) return 0;
    int d[16];
    d[0] = d0; d[4] = d4; d[8] = d8; d[12] = d12;
    d[1] = v - p[3 * sp + 1];  d[2] = v - p[2 * sp + 2];   d[3] = v - p[sp + 3];
    d[5] = v - p[-sp + 3];     d[6] = v - p[-2 * sp + 2];  d[7] = v - p[-3 * sp + 1];
    d[9] = v - p[-3 * sp - 1]; d[10] = v - p[-2 * sp - 2]; d[11] = v - p[-sp - 3];
    d[13] = v - p[sp - 3];     d[14] = v - p[2 * sp - 2];  d[15] = v - p[3 * sp - 1];
    int best = 0;
    if (pb) {  // max over arcs of min d (centre brighter)
        int a[16], c[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) c[k] = min(a[k], a[(k + 2) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = min(c[k], c[(k + 4) & 15]);
        int m = -256;
#pragma unroll
        for (int k = 0; k < 16; ++k) m = max(m, min(a[k], d[(k + 8) & 15]));
        best = m;
    }
    if (pd) {  // max over arcs of min -d (centre darker)
        int a[16], c[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = max(d[k], d[(k + 1) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) c[k] = max(a[k], a[(k + 2) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = max(c[k], c[(k + 4) & 15]);
        int m = 256;
#pragma unroll
        for (int k = 0; k < 16; ++k) m = min(m, max(a[k], d[(k + 8) & 15]));
        best = max(best, -m);
    }
    const int s = best - 1;
    return s >= t ? s : 0;
}

extern __shared__ __align__(16) uint8_t fast_smem[];

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
    uint8_t *pix = fast_smem;                               // rows yb-4 .. ye+3
    uint8_t *sc = fast_smem + (SVO_FAST_BAND + 8) * sp;     // rows yb-1 .. ye

    {   // stage pixel rows
        const uint4 *src = reinterpret_cast<const uint4 *>(b.pyr + (size_t)slot * g.pyr_bytes + L.off + (size_t)(yb - 4) * sp);
        uint4 *dst = reinterpret_cast<uint4 *>(pix);
        const int n16 = ((nrow + 8) * sp) >> 4;
        for (int i = tid; i < n16; i += FAST_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    {   // scores for rows yb-1..ye, columns x0-1..x1
        const int sw = L.x1 - L.x0 + 2;
        const int n = (nrow + 2) * sw;
        for (int i = tid; i < n; i += FAST_THREADS) {
            const int r = i / sw, x = L.x0 - 1 + (i - r * sw);
            sc[r * sp + x] = (uint8_t)fast_score(pix + (r + 3) * sp + x, sp, g.fast_threshold);
        }
    }
    __syncthreads();

    // non-max suppression + raster-ordered compaction
    const int wv = L.x1 - L.x0;
    const int total = nrow * wv;
    const int nwarps = FAST_THREADS / 32;
    const int seg = (((total + nwarps - 1) / nwarps) + 31) & ~31;
    const int beg = warp * seg, end = min(beg + seg, total);
    __shared__ int wcnt[FAST_THREADS / 32];

    auto kept = [&](int i, uint32_t &packed) -> bool {
        if (i >= end) return false;
        const int r = i / wv, x = L.x0 + (i - r * wv);
        const uint8_t *q = sc + (r + 1) * sp + x;
        const int s = q[0];
        if (!s) return false;
        const bool k = s > q[-1] && s > q[1] && s > q[-sp - 1] && s > q[-sp] && s > q[-sp + 1] &&
                       s > q[sp - 1] && s > q[sp] && s > q[sp + 1];
        packed = pack_xy(x, yb + r, s);
        return k;
    };

    int cnt = 0;
    for (int base = beg; base < end; base += 32) {
        uint32_t pk;
        cnt += __popc(__ballot_sync(0xffffffffu, kept(base + lane, pk)));
    }
    if (lane == 0) wcnt[warp] = cnt;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) {
        const int c = wcnt[w];
        if (w < warp) off += c;
        tot += c;
    }
    uint32_t *out = b.bands + (size_t)slot * g.band_total + L.band_off + (size_t)band * L.band_cap;
    for (int base = beg; base < end; base += 32) {
        uint32_t pk = 0;
        const bool k = kept(base + lane, pk);
        const uint32_t m = __ballot_sync(0xffffffffu, k);
        if (k) out[off + __popc(m & ((1u << lane) - 1u))] = pk;
        off += __popc(m);
    }
    if (tid == 0) b.bandcnt[(size_t)slot * g.bandcnt_total + L.bandcnt_off + band] = tot;
}

int fast_smem_bytes(const Geom &g)
{
    int mx = 0;
    for (int l = 0; l < g.nlevels; ++l) mx = g.lv[l].pitch > mx ? g.lv[l].pitch : mx;
    return (2 * SVO_FAST_BAND + 10) * mx;
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
