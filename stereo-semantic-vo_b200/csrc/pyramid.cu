// pyramid.cu — cv::ORB's image pyramid: level l = resize(level l-1, INTER_LINEAR_EXACT)
// (8-bit fixed point: Q8 horizontal, Q8 vertical, round at Q16).  SURVEY.md A.2;
// called from frame::featuredetect (src/frame.cc:75-79) via cv::ORB.
//
// HBM-bound byte work: per output quad one thread reads two source rows through
// per-level coefficient tables ((ofs << 9) | w1, built on the host in double exactly as
// OpenCV does) and stores one aligned 32-bit word.  Algorithmic bytes per level:
// px(l-1) read + px(l) written.
#include "svo_internal.cuh"


// One thread produces a 4-pixel wide, `rs` rows tall column strip (8 rows on the large levels, 4 on the small ones,
// which need the parallelism more).  The horizontal Q8 pass of a source row is
// shared by the two output rows that straddle it: walking down the strip, the lower source row of one output row
// is usually the upper source row of the next (scale 1.2: 1.2 horizontal passes per output row instead of 2),
// and the four x-table entries are decoded once per strip.  Same integer arithmetic as the per-pixel form.
__global__ void __launch_bounds__(256) k_resize(Bufs b, Geom g, int l, int slot0, int rs)
{
    const LevelGeom &d = g.lv[l];
    const LevelGeom &s = g.lv[l - 1];
    const int qpr = d.pitch >> 2;  // quads per (padded) row
    const int strips = (d.h + rs - 1) / rs;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= qpr * strips) return;
    const int st = q / qpr, x = (q - st * qpr) << 2;
    const int y0 = st * rs, y1 = min(y0 + rs, d.h);
    const size_t base = (size_t)(slot0 + blockIdx.y) * g.pyr_bytes;
    const uint8_t *src = b.pyr + base + s.off;
    uint8_t *dst = b.pyr + base + d.off;
    const uint32_t *xt = b.rtab + d.tab_off;
    const int sp = s.pitch;
    int i0[4], i1[4];
    uint32_t wx0[4], wx1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t tx = x + k < d.w ? xt[x + k] : 0u;
        i0[k] = tx >> 9; wx1[k] = tx & 511u; wx0[k] = 256u - wx1[k];
        i1[k] = wx1[k] ? i0[k] + 1 : i0[k];
        if (x + k >= d.w) { wx0[k] = 0; wx1[k] = 0; }          // padding columns hold 0
    }
    auto hrow = [&](int yy, uint32_t (&h)[4]) {
        const uint8_t *r = src + (size_t)yy * sp;
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = r[i0[k]] * wx0[k] + r[i1[k]] * wx1[k];
    };
    uint32_t ha[4], hb[4];
    int ya = -1, ybr = -1;                                     // source rows currently held in ha / hb
    for (int y = y0; y < y1; ++y) {
        const uint32_t ty = xt[d.w + y];
        const int yo = ty >> 9;
        const uint32_t wy1 = ty & 511u, wy0 = 256u - wy1;
        const int yn = wy1 ? yo + 1 : yo;
        if (yo == ybr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ha[k] = hb[k];
        } else if (yo != ya) hrow(yo, ha);
        ya = yo;
        if (yn == yo) {
#pragma unroll
            for (int k = 0; k < 4; ++k) hb[k] = ha[k];
        } else hrow(yn, hb);
        ybr = yn;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) out |= ((ha[k] * wy0 + hb[k] * wy1 + 32768u) >> 16) << (8 * k);
        *reinterpret_cast<uint32_t *>(dst + (size_t)y * d.pitch + x) = out;
    }
}

// The same strips from the quad table (resize_quads.h): the eight horizontal taps of an output quad lie within 8
// consecutive source bytes, so a source row costs three aligned word loads, two funnel shifts (an 8-byte window that
// starts at the first tap), two byte permutes (the taps of columns 0,1 and 2,3) and four two-way dot products (dp2a:
// (256 - w1, w1) as two u16 times two bytes) instead of eight byte loads with 64-bit address arithmetic and eight
// multiply-adds: 146 -> ~45 warp instructions per output quad row.  The two source rows of an output row live in two
// register sets whose roles alternate from row to row (the lower row of one output row is usually the upper row of the
// next), tracked by the source row each set holds, so nothing is copied.  Identical integer arithmetic.
__device__ __forceinline__ void resize_hrow(const uint8_t *srcq, int sp, int yy, uint32_t sh, uint32_t sel01, uint32_t sel23,
                                            const uint32_t (&W)[4], uint32_t (&h)[4])
{
    const uint32_t *r = reinterpret_cast<const uint32_t *>(srcq + (size_t)yy * sp);
    const uint32_t w0 = r[0], w1 = r[1], w2 = r[2];     // up to 11 bytes past the row's width: inside the padded row or the next one
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    const uint32_t p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);
    h[0] = __dp2a_lo(W[0], p01, 0u); h[1] = __dp2a_hi(W[1], p01, 0u);
    h[2] = __dp2a_lo(W[2], p23, 0u); h[3] = __dp2a_hi(W[3], p23, 0u);
}
__device__ __forceinline__ uint32_t resize_vpack(const uint32_t (&u)[4], const uint32_t (&v)[4], uint32_t wy0, uint32_t wy1)
{
    // each sum is below 2^24, so the rounded Q16 result is its byte 2
    const uint32_t s0 = u[0] * wy0 + v[0] * wy1 + 32768u, s1 = u[1] * wy0 + v[1] * wy1 + 32768u;
    const uint32_t s2 = u[2] * wy0 + v[2] * wy1 + 32768u, s3 = u[3] * wy0 + v[3] * wy1 + 32768u;
    return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}

// one output row y of a strip: upper source row in U (holding source row ru), lower in V (holding rv); `yt` = the level's
// y table, `srcq` = the strip's first aligned source word in source row 0, `dstq` = the strip's quad in output row 0
#define RESIZE_ROW(y, U, ru, V, rv)                                                              \
    if ((y) < y1) {                                                                              \
        const uint32_t ty = __ldg(yt + (y));                                                     \
        const int yo = (int)(ty >> 9);                                                           \
        const uint32_t wy1 = ty & 511u, wy0 = 256u - wy1;                                        \
        if (ru != yo) { resize_hrow(srcq, sp, yo, sh, sel01, sel23, W, U); ru = yo; }            \
        if (wy1 && rv != yo + 1) { resize_hrow(srcq, sp, yo + 1, sh, sel01, sel23, W, V); rv = yo + 1; }   \
        *reinterpret_cast<uint32_t *>(dstq + (size_t)(y) * dp) = resize_vpack(U, V, wy0, wy1);   \
    }

__global__ void __launch_bounds__(256) k_resize_q(Bufs b, Geom g, int l, int slot0, int rs)
{
    const LevelGeom &d = g.lv[l];
    const LevelGeom &s = g.lv[l - 1];
    const int qpr = d.pitch >> 2;  // quads per (padded) row
    const int strips = (d.h + rs - 1) / rs;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= qpr * strips) return;
    const int st = q / qpr, xq = q - st * qpr;
    const int y0 = st * rs, y1 = min(y0 + rs, d.h);
    const size_t base = (size_t)(slot0 + blockIdx.y) * g.pyr_bytes;
    const uint4 *qe = reinterpret_cast<const uint4 *>(b.rqtab) + 2 * (size_t)(d.rq_off + xq);
    const uint4 e0 = __ldg(qe);
    const uint2 e1 = __ldg(reinterpret_cast<const uint2 *>(qe + 1));
    const uint32_t sh = e0.x >> 16, sel01 = e0.y & 0xffffu, sel23 = e0.y >> 16;
    const uint32_t W[4] = {e0.z, e0.w, e1.x, e1.y};
    const uint8_t *srcq = b.pyr + base + s.off + (e0.x & 0xffffu);
    uint8_t *dstq = b.pyr + base + d.off + 4 * xq;
    const uint32_t *yt = b.rtab + d.tab_off + d.w;
    const int sp = s.pitch, dp = d.pitch;
    uint32_t ha[4], hb[4];
    int ra = -1, rb = -1;                                       // source rows held in ha / hb
    for (int y = y0; y < y1; y += 2) {
        RESIZE_ROW(y, ha, ra, hb, rb)
        RESIZE_ROW(y + 1, hb, rb, ha, ra)
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// All seven resized levels in ONE launch. A CTA owns a band of rows of the last level and everything below it that the
// band depends on: the level-0 rows arrive with one TMA bulk copy (the pitched rows of a band are contiguous in HBM),
// then level 1 is computed from them in shared memory, level 2 from level 1, ... — every level is read from shared
// memory, never from HBM — and each level's rows this band OWNS are stored with 16-byte writes while the next level is
// being computed.  Bands overlap by the few rows of halo the 2-tap vertical filter needs (the host precomputes, per
// band and level, the computed and the owned row range from the same coefficient tables the arithmetic uses), so
// neighbouring CTAs recompute those rows instead of synchronising.  Same Q8 x Q8 integer arithmetic as k_resize.
// Replaces 7 dependent launches (126 us of device time per 64 images, and 0.5 ms of queueing inside a pipeline lane).
// ---------------------------------------------------------------------------------------------------------------------
#define PYR_THREADS 512
#define PYR_STRIP 8
extern __shared__ __align__(128) uint8_t pyr_smem[];

__global__ void __launch_bounds__(PYR_THREADS) k_pyramid(Bufs b, Geom g, int slot0)
{
    __shared__ PyrBand bd;
    if (threadIdx.x == 0) bd = b.pyr_bands[blockIdx.x];
    const size_t base = (size_t)(slot0 + blockIdx.y) * g.pyr_bytes;
    const int tid = threadIdx.x;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    {   // level 0 rows [clo, chi]: one bulk copy
        const LevelGeom &L = g.lv[0];
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const uint32_t bytes = (uint32_t)(bd.chi[0] - bd.clo[0] + 1) * (uint32_t)L.pitch;
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                             "r"((uint32_t)__cvta_generic_to_shared(pyr_smem + g.pyr_soff[0])),
                         "l"(b.pyr + base + L.off + (size_t)bd.clo[0] * L.pitch), "r"(bytes), "r"(bar_a)
                         : "memory");
        }
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "W_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@p bra D_%=;\n\t"
            "bra W_%=;\n\t"
            "D_%=:\n\t"
            "}\n" ::"r"(bar_a) : "memory");
    }
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom &d = g.lv[l];
        const LevelGeom &s = g.lv[l - 1];
        const uint8_t *src = pyr_smem + g.pyr_soff[l - 1];      // rows [clo[l-1], chi[l-1]] of level l-1, pitch s.pitch
        uint8_t *dst = pyr_smem + g.pyr_soff[l];
        const uint32_t *xt = b.rtab + d.tab_off;
        const int qpr = d.pitch >> 2, clo = bd.clo[l], nrows = bd.chi[l] - clo + 1, slo = bd.clo[l - 1], sp = s.pitch;
        // k_resize's column strips, over shared memory: a thread owns 4 output columns and walks PYR_STRIP rows down, sharing
        // the horizontal pass of a source row between the two output rows that straddle it
        const int strips = (nrows + PYR_STRIP - 1) / PYR_STRIP;
        const uint4 *qt = reinterpret_cast<const uint4 *>(b.rqtab) + 2 * (size_t)d.rq_off;
        const uint32_t *yt = xt + d.w;
        const int dp = d.pitch;
        if (d.rq_ok) {
            // the quad-table strips of k_resize_q over shared memory (source rows are addressed relative to the tile's first row)
            for (int it = tid; it < qpr * strips; it += PYR_THREADS) {
                const int st = it / qpr, xq = it - st * qpr;
                const int y0 = clo + st * PYR_STRIP, y1 = min(y0 + PYR_STRIP, clo + nrows);
                const uint4 e0 = __ldg(qt + 2 * xq);
                const uint2 e1 = __ldg(reinterpret_cast<const uint2 *>(qt + 2 * xq + 1));
                const uint32_t sh = e0.x >> 16, sel01 = e0.y & 0xffffu, sel23 = e0.y >> 16;
                const uint32_t W[4] = {e0.z, e0.w, e1.x, e1.y};
                const uint8_t *srcq = src + (int)(e0.x & 0xffffu) - slo * sp;
                uint8_t *dstq = dst + 4 * xq - clo * dp;
                uint32_t ha[4], hb[4];
                int ra = -1, rb = -1;
                for (int y = y0; y < y1; y += 2) {
                    RESIZE_ROW(y, ha, ra, hb, rb)
                    RESIZE_ROW(y + 1, hb, rb, ha, ra)
                }
            }
        } else
        for (int it = tid; it < qpr * strips; it += PYR_THREADS) {
            const int st = it / qpr, x = (it - st * qpr) << 2;
            const int y0 = clo + st * PYR_STRIP, y1 = min(y0 + PYR_STRIP, clo + nrows);
            int i0[4], i1[4];
            uint32_t wx0[4], wx1[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t tx = x + k < d.w ? __ldg(xt + x + k) : 0u;
                i0[k] = tx >> 9; wx1[k] = tx & 511u; wx0[k] = 256u - wx1[k];
                i1[k] = wx1[k] ? i0[k] + 1 : i0[k];
                if (x + k >= d.w) { wx0[k] = 0; wx1[k] = 0; }          // padding columns hold 0
            }
            auto hrow = [&](int yy, uint32_t (&h)[4]) {
                const uint8_t *r = src + (size_t)(yy - slo) * sp;
#pragma unroll
                for (int k = 0; k < 4; ++k) h[k] = r[i0[k]] * wx0[k] + r[i1[k]] * wx1[k];
            };
            uint32_t ha[4], hb[4];
            int ya = -1, ybr = -1;
            for (int y = y0; y < y1; ++y) {
                const uint32_t ty = __ldg(xt + d.w + y);
                const int yo = ty >> 9;
                const uint32_t wy1 = ty & 511u, wy0 = 256u - wy1;
                const int yn = wy1 ? yo + 1 : yo;
                if (yo == ybr) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ha[k] = hb[k];
                } else if (yo != ya) hrow(yo, ha);
                ya = yo;
                if (yn == yo) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) hb[k] = ha[k];
                } else hrow(yn, hb);
                ybr = yn;
                uint32_t out = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) out |= ((ha[k] * wy0 + hb[k] * wy1 + 32768u) >> 16) << (8 * k);
                *reinterpret_cast<uint32_t *>(dst + (size_t)(y - clo) * d.pitch + x) = out;
            }
        }
        __syncthreads();
        // the rows this band owns go to HBM (16-byte stores, whole pitched rows) while the next level is computed from the same tile
        const int v16 = d.pitch >> 4, own = bd.ohi[l] - bd.olo[l];
        uint8_t *gdst = b.pyr + base + d.off + (size_t)bd.olo[l] * d.pitch;
        const uint8_t *sown = dst + (size_t)(bd.olo[l] - clo) * d.pitch;
        for (int it = tid; it < v16 * own; it += PYR_THREADS)
            reinterpret_cast<uint4 *>(gdst)[it] = reinterpret_cast<const uint4 *>(sown)[it];
    }
}

static int g_pyr_smem_set = 0;
int setup_pyramid_attributes(const Geom &g)
{
    // per kernel, not per context: a context for a smaller geometry must not lower the limit under a larger one
    if (!g.pyr_nbands || g.pyr_smem <= g_pyr_smem_set) return 0;
    if (cudaFuncSetAttribute(k_pyramid, cudaFuncAttributeMaxDynamicSharedMemorySize, g.pyr_smem) != cudaSuccess) return 1;
    g_pyr_smem_set = g.pyr_smem;
    return 0;
}

// Level 0 is read where it lies in device memory (the caller's device buffer, or the lane's landing zone
// for host inputs; rows `stride` bytes apart, any alignment) and repacked into the 16-byte-pitched level-0
// layout.  One thread per 16 output bytes.  Interleaved BGR sources (SVO_STRIDE_BGR set in the stride word) are
// converted on the way with OpenCV's 15-bit fixed-point weights (cv::ORB's cvtColor(BGR2GRAY) of colour input).
__global__ void __launch_bounds__(256) k_unpack(Bufs b, Geom g, int slot0, const FramePtrs *__restrict__ fp,
                                                const int *__restrict__ strides)
{
    const LevelGeom &L = g.lv[0];
    const int per_row = L.pitch >> 4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_row * L.h) return;
    const int y = i / per_row, x = (i - y * per_row) << 4;
    const int sraw = strides[blockIdx.y];
    if (sraw <= 0) return;                        // this image was uploaded with a 2-D copy
    const int stride = sraw & (SVO_STRIDE_BGR - 1);
    const FramePtrs &F = fp[blockIdx.y >> 1];
    const uint8_t *img = (blockIdx.y & 1) ? F.right : F.left;
    uint32_t w[4] = {0, 0, 0, 0};
    if (sraw & SVO_STRIDE_BGR) {
        const uint8_t *src = img + (size_t)y * stride + 3 * x;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (x + k < L.w) {
                const uint32_t v = (src[3 * k] * 3735u + src[3 * k + 1] * 19235u + src[3 * k + 2] * 9798u + 16384u) >> 15;
                w[k >> 2] |= v << (8 * (k & 3));
            }
    } else {
        const uint8_t *src = img + (size_t)y * stride + x;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (x + k < L.w) w[k >> 2] |= (uint32_t)src[k] << (8 * (k & 3));
    }
    *reinterpret_cast<uint4 *>(b.pyr + (size_t)(slot0 + blockIdx.y) * g.pyr_bytes + L.off + (size_t)y * L.pitch + x) =
        make_uint4(w[0], w[1], w[2], w[3]);
}

void launch_unpack(const Bufs &b, const Geom &g, int slot0, int nimg, const FramePtrs *fp,
                   const int *strides, cudaStream_t st, long long *launches)
{
    const int n = (g.lv[0].pitch >> 4) * g.lv[0].h;
    dim3 grid((n + 255) / 256, nimg);
    k_unpack<<<grid, 256, 0, st>>>(b, g, slot0, fp, strides);
    ++*launches;
}

void launch_pyramid(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    if (g.pyr_nbands && b.pyr_bands && nimg <= g.pyr_fused_max) {
        k_pyramid<<<dim3(g.pyr_nbands, nimg), PYR_THREADS, g.pyr_smem, st>>>(b, g, slot0);
        ++*launches;
        return;
    }
    for (int l = 1; l < g.nlevels; ++l) {
        const int rs = (size_t)g.lv[l].w * g.lv[l].h * nimg >= (size_t)8 << 20 ? 8 : 4;   // fewer rows per thread when the launch is small
        const int quads = (g.lv[l].pitch >> 2) * ((g.lv[l].h + rs - 1) / rs);
        dim3 grid((quads + 255) / 256, nimg);
        if (g.lv[l].rq_ok) k_resize_q<<<grid, 256, 0, st>>>(b, g, l, slot0, rs);
        else k_resize<<<grid, 256, 0, st>>>(b, g, l, slot0, rs);
        ++*launches;
    }
}
