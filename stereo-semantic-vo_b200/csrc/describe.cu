// describe.cu — Harris response, intensity-centroid angle, 7x7 Gaussian blur and rBRIEF-256,
// as cv::ORB computes them inside frame::featuredetect (src/frame.cc:75-79).
// SURVEY.md A.5 (Harris, IC angle, fastAtan2), A.7 (blur), A.8 (rBRIEF).
//
// Float bit-exactness: every float operation below is written with the _rn intrinsics so
// nvcc cannot contract a multiply-add into an FMA; OpenCV's portable path rounds each
// operation separately and that is what the oracle is pinned to.
#include "svo_internal.cuh"
#include <float.h>
#include <cuda/barrier>

// Warps per CTA of k_harris / k_harris4 / k_describe.  Four, not eight: a 4-warp k_describe CTA (5.9 KB of shared memory)
// fits beside the two 109-KB k_fast CTAs another lane keeps on an SM, an 8-warp one (11.8 KB) does not: 26.86 k -> 27.11 k
// frames/s in the four-lane pipeline (profiles/r2_experiments.md).
#ifndef DESC_WARPS
#define DESC_WARPS 4
#endif

__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
// cv::getGaussianKernel(7, 2, CV_32F), exact bits
__constant__ uint32_t c_gauss[7] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u,
                                    0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u};

// ---------------------------------------------------------------------------------------
// Harris response of the survivors of the first cull: one warp per candidate, the 49 block
// positions spread over the lanes, int32 sums reduced with redux.sync (order-free, exact).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DESC_WARPS * 32) k_harris(Bufs b, Geom g, int slot0)
{
    // per warp: the candidate's 9x9 patch as 9 rows of three aligned words (the 9 columns x-4..x+4 always lie
    // inside three words), fetched with 27 word loads instead of 8 byte loads for each of the 49 positions
    __shared__ uint32_t patch[DESC_WARPS][9 * 3 + 1];
    const int l = blockIdx.y, slot = slot0 + blockIdx.z;
    const LevelGeom &L = g.lv[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int n = b.kept1[(size_t)slot * SVO_MAX_LEVELS + l];
    if (n > L.cap2) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(b.status + slot, SVO_STATUS_OVERFLOW);
        n = L.cap2;
    }
    const uint8_t *img = b.pyr + (size_t)slot * g.pyr_bytes + L.off;
    const uint32_t *cval = b.cval + (size_t)slot * g.cand_total + L.cand_off;
    float *key2 = b.key2 + (size_t)slot * g.total2 + L.off2;
    uint32_t *val2 = b.val2 + (size_t)slot * g.total2 + L.off2;
    const int sp = L.pitch;
    const uint8_t *pb = reinterpret_cast<const uint8_t *>(patch[warp]);
    for (int c = blockIdx.x * DESC_WARPS + warp; c < n; c += gridDim.x * DESC_WARPS) {
        const uint32_t e = cval[c];
        const int x = unpack_x(e), y = unpack_y(e);
        const int xw = (x - 4) & ~3, xo = (x - 4) & 3;          // first word column, byte offset of column x-4 in it
        __syncwarp();
        if (lane < 27) {
            const int r = lane / 3, w = lane - 3 * r;
            patch[warp][lane] = *reinterpret_cast<const uint32_t *>(img + (size_t)(y - 4 + r) * sp + xw + 4 * w);
        }
        __syncwarp();
        int a = 0, bb = 0, cc = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = lane + 32 * r;
            if (k < 49) {
                const int dy = k / 7, dx = k % 7;                // block position: rows dy..dy+2, columns dx..dx+2 of the patch
                const uint8_t *p = pb + (dy + 1) * 12 + xo + dx + 1;
                const int p00 = p[-13], p01 = p[-12], p02 = p[-11];
                const int p10 = p[-1], p12 = p[1];
                const int p20 = p[11], p21 = p[12], p22 = p[13];
                const int Ix = (p12 - p10) * 2 + (p02 - p00) + (p22 - p20);
                const int Iy = (p21 - p01) * 2 + (p20 - p00) + (p22 - p02);
                a += Ix * Ix; bb += Iy * Iy; cc += Ix * Iy;
            }
        }
        a = __reduce_add_sync(0xffffffffu, a);
        bb = __reduce_add_sync(0xffffffffu, bb);
        cc = __reduce_add_sync(0xffffffffu, cc);
        if (lane == 0) {
            const float scale = 1.f / ((1 << 2) * 7 * 255.f);
            const float s2 = __fmul_rn(scale, scale);
            const float s4 = __fmul_rn(__fmul_rn(s2, scale), scale);
            const float fa = (float)a, fb = (float)bb, fc = (float)cc;
            const float t1 = __fmul_rn(fa, fb);
            const float t2 = __fmul_rn(fc, fc);
            const float s = __fadd_rn(fa, fb);
            const float t3 = __fmul_rn(__fmul_rn(0.04f, s), s);
            const float v = __fsub_rn(__fsub_rn(t1, t2), t3);
            key2[c] = __fmul_rn(v, s4);
            val2[c] = e;
        }
    }
}

// The same response with EIGHT lanes per candidate (four candidates per warp): lane r of a group owns block row r
// (7 of 8 lanes busy) and walks its 7 positions from three patch rows held in registers.  With the column sums
// s[k] = p0[k] + 2 p1[k] + p2[k] and differences v[k] = p2[k] - p0[k] of its three rows,
//     Ix(dx) = s[dx + 2] - s[dx]            Iy(dx) = v[dx] + 2 v[dx + 1] + v[dx + 2]
// (the 3x3 Sobel pair written out in k_harris, regrouped; integers, so exact).  The warp-per-candidate form spends
// its instructions on per-candidate overhead (two passes over 49 positions with 32 lanes, three full-warp reductions):
// 248 warp instructions per candidate against ~60 here.  The float tail is the one of k_harris.
__global__ void __launch_bounds__(DESC_WARPS * 32) k_harris4(Bufs b, Geom g, int slot0)
{
    __shared__ uint32_t patch[DESC_WARPS][4][28];    // per candidate 9 rows of three aligned words
    const int l = blockIdx.y, slot = slot0 + blockIdx.z;
    const LevelGeom &L = g.lv[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int n = b.kept1[(size_t)slot * SVO_MAX_LEVELS + l];
    if (n > L.cap2) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(b.status + slot, SVO_STATUS_OVERFLOW);
        n = L.cap2;
    }
    const uint8_t *img = b.pyr + (size_t)slot * g.pyr_bytes + L.off;
    const uint32_t *cval = b.cval + (size_t)slot * g.cand_total + L.cand_off;
    float *key2 = b.key2 + (size_t)slot * g.total2 + L.off2;
    uint32_t *val2 = b.val2 + (size_t)slot * g.total2 + L.off2;
    const int sp = L.pitch;
    const int grp = lane >> 3, r = lane & 7;
    for (int c0 = (blockIdx.x * DESC_WARPS + warp) * 4; c0 < n; c0 += gridDim.x * DESC_WARPS * 4) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {                // 4 x 27 words, lane after lane
            const int t = lane + 32 * i;
            const int cg = (t * 19) >> 9, w = t - 27 * cg;       // t / 27 for t < 128
            if (t < 108 && c0 + cg < n) {
                const uint32_t e = cval[c0 + cg];
                const int x = unpack_x(e), y = unpack_y(e);
                const int row = (w * 11) >> 5, col = w - 3 * row; // w / 3 for w < 27
                patch[warp][cg][w] = *reinterpret_cast<const uint32_t *>(img + (size_t)(y - 4 + row) * sp + ((x - 4) & ~3) + 4 * col);
            }
        }
        __syncwarp();
        const int c = c0 + grp;
        const bool have = c < n;
        const uint32_t e = have ? cval[c] : 0u;
        int a = 0, bb = 0, cc = 0;
        if (have && r < 7) {
            const uint32_t sh = 8u * (uint32_t)((unpack_x(e) - 4) & 3);      // byte offset of column x - 4 in the first word
            int p[3][9];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint32_t *pw = patch[warp][grp] + 3 * (r + j);
                const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2];
                const uint32_t lo = __funnelshift_r(w0, w1, sh), mid = __funnelshift_r(w1, w2, sh), hi = w2 >> sh;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    p[j][k] = (int)__byte_perm(lo, 0u, 0x4440u + k);       // byte k, zero-extended: one PRMT
                    p[j][4 + k] = (int)__byte_perm(mid, 0u, 0x4440u + k);
                }
                p[j][8] = (int)(hi & 0xffu);
            }
            int s[9], v[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { s[k] = p[0][k] + 2 * p[1][k] + p[2][k]; v[k] = p[2][k] - p[0][k]; }
#pragma unroll
            for (int dx = 0; dx < 7; ++dx) {
                const int Ix = s[dx + 2] - s[dx];
                const int Iy = v[dx] + 2 * v[dx + 1] + v[dx + 2];
                a += Ix * Ix; bb += Iy * Iy; cc += Ix * Iy;
            }
        }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {            // the 8 lanes of a group
            a += __shfl_xor_sync(0xffffffffu, a, d);
            bb += __shfl_xor_sync(0xffffffffu, bb, d);
            cc += __shfl_xor_sync(0xffffffffu, cc, d);
        }
        if (have && r == 0) {
            const float scale = 1.f / ((1 << 2) * 7 * 255.f);
            const float s2 = __fmul_rn(scale, scale);
            const float s4 = __fmul_rn(__fmul_rn(s2, scale), scale);
            const float fa = (float)a, fb = (float)bb, fc = (float)cc;
            const float t1 = __fmul_rn(fa, fb);
            const float t2 = __fmul_rn(fc, fc);
            const float sm = __fadd_rn(fa, fb);
            const float t3 = __fmul_rn(__fmul_rn(0.04f, sm), sm);
            const float val = __fsub_rn(__fsub_rn(t1, t2), t3);
            key2[c] = __fmul_rn(val, s4);
            val2[c] = e;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Blur: u8 -> f32 row pass (taps left to right) -> f32 column pass (centre, then symmetric
// pairs) -> rint -> u8, reflect-101 borders; every float op separately rounded.
// One thread owns a 4-pixel-wide, BLUR_ROWS-tall strip: per row it reads the 12 bytes around
// its quad as three aligned words (neighbouring lanes overlap, so these are L1 hits), turns
// them into floats with a PRMT + FADD magic-number trick (keeps the XU pipe free), forms the 4
// horizontal sums and pushes them into a 7-row register window from which the vertical pass
// is taken.  No shared memory, no barriers; stores are one aligned word per row.
// ---------------------------------------------------------------------------------------
#define BLUR_ROWS SVO_BLUR_ROWS
#define BLUR_THREADS 128

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// bytes {b, 00, 00, 4B} = 2^23 + b as a float; subtracting 2^23 is exact
__device__ __forceinline__ float byte_magic(uint32_t word, int sel)
{
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440 | sel));
}

// Products and tap-free sums are packed (FMUL2 / FADD2, two pixels per instruction; every lane of a packed op is
// rounded separately, exactly like the scalar _rn ops).  An add that consumes a product stays SCALAR: ptxas
// (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, which would change bits.
__device__ __forceinline__ float2 add_prod(float2 acc, float2 prod)
{
    return make_float2(__fadd_rn(acc.x, prod.x), __fadd_rn(acc.y, prod.y));
}
// The same sum as ONE packed instruction: prod * 1 + acc.  The multiplication by one is exact, so the fused result is
// the separately rounded sum bit for bit, and because the multiplicand comes from constant memory (a value the compiler
// cannot see) neither NVVM nor ptxas can turn it back into the mul + add pair that ptxas would contract.
__constant__ float c_one[2] = {1.f, 1.f};
template <bool PACKED>
__device__ __forceinline__ float2 add_prod_p(float2 acc, float2 prod, float2 one)
{
    return PACKED ? __ffma2_rn(prod, one, acc) : add_prod(acc, prod);
}
#define BLUR_RB (4 * 128 + 32)   // staged bytes per tile row: 128 quads + a 16-byte aligned halo on each side

// MARGIN (the current form; <false> is kept for SVO_B200_BLUR_MARGIN=0): sums of products as packed FFMA2 (add_prod_p), and
// the reflect-101 columns left of pixel 0 and right of the last pixel are written into the staged tile once
// (3 bytes per row and side), so every quad — also the first and the last of a row — takes the word-load path; without
// it the warps that hold an edge quad execute both paths on every row (two of the ten warps of a level-0 row).
template <bool MARGIN>
__global__ void __launch_bounds__(BLUR_THREADS) k_blur(Bufs b, Geom g, int slot0)
{
    // The tile's source rows (BLUR_ROWS + 6 halo rows, reflect-101 at the image top/bottom folded into the row choice)
    // are staged in shared memory by TMA bulk copies, one per row, all completing on one mbarrier: the kernel
    // was bound by the latency of its global loads (ncu: long-scoreboard stalls at 28 % occupancy), shared
    // memory removes that from the per-row loop.
    __shared__ __align__(128) uint8_t tile[(BLUR_ROWS + 6) * BLUR_RB];
    __shared__ cuda::barrier<cuda::thread_scope_block> bar;
    int blk = blockIdx.x, l = 0;
    while (l + 1 < g.nlevels && blk >= g.lv[l + 1].blur_tile_off) ++l;
    const int Lw = g.lv[l].w, Lh = g.lv[l].h, sp = g.lv[l].pitch, Loff = g.lv[l].off;
    const int tiles_x = g.lv[l].blur_tiles_x, tq = g.lv[l].blur_tq;
    blk -= g.lv[l].blur_tile_off;
    const int strip = blk / tiles_x, tx = blk - strip * tiles_x;
    const int y0 = strip * BLUR_ROWS;
    const int rows = min(BLUR_ROWS, Lh - y0);
    const int nst = rows + 6;
    const int xa = 4 * tq * tx;                                   // first pixel of the tile (multiple of 16)
    const int xs = max(xa - 16, 0), xe = min(xa + 4 * tq + 16, sp);
    const uint32_t rowbytes = (uint32_t)(xe - xs);                // multiple of 16
    const int slot = slot0 + blockIdx.y;
    const uint8_t *__restrict__ img = b.pyr + (size_t)slot * g.pyr_bytes + Loff;
    uint8_t *__restrict__ out = b.blur + (size_t)slot * g.pyr_bytes + Loff;
    const int tid = threadIdx.x;
    if (tid == 0) {
        init(&bar, BLUR_THREADS);
        cuda::device::experimental::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
    const int xv = MARGIN ? xa - 16 : xs;          // pixel held by byte 0 of a staged row (MARGIN: also for the first tile, where it is -16)
    if (tid < 32) {
        for (int r = tid; r < nst; r += 32)
            cuda::device::memcpy_async_tx(tile + r * BLUR_RB + (xs - xv), img + (size_t)reflect101(y0 + r - 3, Lh) * sp + xs,
                                          cuda::aligned_size_t<16>(rowbytes), bar);
    }
    if (tid == 0) tok = cuda::device::barrier_arrive_tx(bar, 1, rowbytes * (uint32_t)nst);
    else tok = bar.arrive();
    float2 gk[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) gk[k] = make_float2(__uint_as_float(c_gauss[k]), __uint_as_float(c_gauss[k]));
    const float2 two23 = make_float2(-8388608.f, -8388608.f), rnd = make_float2(12582912.f, 12582912.f);
    const float2 one = make_float2(c_one[0], c_one[1]);
    const int x0 = xa + 4 * tid;
    const bool active = tid < tq && x0 < Lw;
    const bool interior = MARGIN || (x0 >= 4 && x0 + 6 < Lw);   // bytes x0-3 .. x0+6 all inside the row (or its margins)
    const int lx = x0 - xv;
    float2 win[7][2];                               // horizontal sums of the last 7 rows: pixels (0,1) and (2,3)
#pragma unroll
    for (int k = 0; k < 7; ++k) win[k][0] = win[k][1] = make_float2(0.f, 0.f);
    bar.wait(std::move(tok));
    if (MARGIN) {
        // reflect-101: pixel -k = pixel k, pixel Lw - 1 + k = pixel Lw - 1 - k (k = 1..3), one staged row per thread.  The
        // right margin is written by every tile whose quads can reach it (the row's end lies at most 8 pixels past the
        // tile's quads); what lies beyond only feeds the padding columns of the output.
        if (tid < nst) {
            uint8_t *row = tile + tid * BLUR_RB;
            if (xa == 0) { row[15] = row[17]; row[14] = row[18]; row[13] = row[19]; }
            if (Lw > xs + 4 && Lw <= xa + 4 * tq + 8) {
                const int e = Lw - xv;
                row[e] = row[e - 2]; row[e + 1] = row[e - 3]; row[e + 2] = row[e - 4];
            }
        }
        __syncthreads();
    }
    if (!active) return;
    for (int r0 = 0; r0 < nst; r0 += 7) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int r = r0 + k;                       // window row r <-> image row y0 + r - 3
            if (r < nst) {
                const uint8_t *row = tile + r * BLUR_RB;
                float m[10];                            // 2^23 + pixel x0-3+j
                if (interior) {
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(row + lx - 4);
                    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                    m[0] = byte_magic(w0, 1); m[1] = byte_magic(w0, 2); m[2] = byte_magic(w0, 3);
                    m[3] = byte_magic(w1, 0); m[4] = byte_magic(w1, 1); m[5] = byte_magic(w1, 2); m[6] = byte_magic(w1, 3);
                    m[7] = byte_magic(w2, 0); m[8] = byte_magic(w2, 1); m[9] = byte_magic(w2, 2);
                } else {
#pragma unroll
                    for (int j = 0; j < 10; ++j) m[j] = __uint_as_float(0x4B000000u | row[reflect101(x0 - 3 + j, Lw) - xv]);
                }
                float2 P[9];                            // P[t] = pixels (x0-3+t, x0-2+t)
#pragma unroll
                for (int t = 0; t < 9; ++t) P[t] = __fadd2_rn(make_float2(m[t], m[t + 1]), two23);
                float2 a01 = __fmul2_rn(gk[0], P[0]), a23 = __fmul2_rn(gk[0], P[2]);
#pragma unroll
                for (int t = 1; t < 7; ++t) {
                    a01 = add_prod_p<MARGIN>(a01, __fmul2_rn(gk[t], P[t]), one);
                    a23 = add_prod_p<MARGIN>(a23, __fmul2_rn(gk[t], P[t + 2]), one);
                }
                win[k][0] = a01; win[k][1] = a23;
                if (r >= 6) {                           // rows r-6 .. r are in the window: emit image row y0 + r - 6
                    float2 o[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float2 acc = __fmul2_rn(gk[3], win[(k + 4) % 7][j]);
                        acc = add_prod_p<MARGIN>(acc, __fmul2_rn(gk[4], __fadd2_rn(win[(k + 5) % 7][j], win[(k + 3) % 7][j])), one);
                        acc = add_prod_p<MARGIN>(acc, __fmul2_rn(gk[5], __fadd2_rn(win[(k + 6) % 7][j], win[(k + 2) % 7][j])), one);
                        acc = add_prod_p<MARGIN>(acc, __fmul2_rn(gk[6], __fadd2_rn(win[k][j], win[(k + 1) % 7][j])), one);
                        // round to nearest even by adding 1.5 * 2^23: the low mantissa byte is the pixel
                        // (0 <= acc <= 255 * (sum of taps) < 255.5, so no clamp is needed)
                        o[j] = __fadd2_rn(acc, rnd);
                    }
                    const uint32_t p01 = __byte_perm(__float_as_uint(o[0].x), __float_as_uint(o[0].y), 0x0040);
                    const uint32_t p23 = __byte_perm(__float_as_uint(o[1].x), __float_as_uint(o[1].y), 0x0040);
                    *reinterpret_cast<uint32_t *>(out + (size_t)(y0 + r - 6) * sp + x0) = __byte_perm(p01, p23, 0x5410);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// IC angle (un-blurred level) + rBRIEF (blurred level) + final keypoint record.
// One warp per surviving keypoint; lane i computes descriptor byte i.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k);
    const float p5 = __fmul_rn(0.1555786518463281f, k), p7 = __fmul_rn(-0.04432655554792128f, k);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__device__ __forceinline__ int rint_small(float v) { return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4B400000; }

__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe(Bufs b, Geom g, int slot0, const int *__restrict__ pattern)
{
    const int slot = slot0 + blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int *kept2 = b.kept2 + (size_t)slot * SVO_MAX_LEVELS;
    // output index -> (level, index within level): levels are concatenated in order
    const int oi = blockIdx.x * DESC_WARPS + warp;
    int l = 0, base = 0, total = 0;
    for (int i = 0; i < g.nlevels; ++i) {
        const int k = kept2[i];
        if (oi >= total + k) { l = i + 1; base = total + k; }
        total += k;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        b.nkp[slot] = total;
        if (total > g.kp_cap) atomicOr(b.status + slot, SVO_STATUS_OVERFLOW);
    }
    if (oi >= total || oi >= g.kp_cap) return;
    const LevelGeom &L = g.lv[l];
    const int c = oi - base;
    const uint8_t *img = b.pyr + (size_t)slot * g.pyr_bytes + L.off;
    const uint8_t *blr = b.blur + (size_t)slot * g.pyr_bytes + L.off;
    const int sp = L.pitch;
    const uint32_t e = b.val2[(size_t)slot * g.total2 + L.off2 + c];
    const int x = unpack_x(e), y = unpack_y(e);
    // lane i owns descriptor byte i: its 8 tests (x0,y0,x1,y1 as int8) are loop-invariant
    int pat[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pat[j] = pattern[8 * lane + j];
    // intensity centroid over the radius-15 disc: lane <-> column u, rows |v| <= umax[|u|]
    // (the disc is symmetric, so the row limit of column u is the column limit of row |u|)
    const uint8_t *ctr = img + (size_t)y * sp + x;
    const int u = lane - 15;
    const int vm = lane < 31 ? c_umax[u < 0 ? -u : u] : -1;
    // rows +v and -v are taken together (as cv::ORB's IC_Angle does): sum += a + b, m01 += v * (a - b)
    int sum = 0, m01 = 0;
    if (0 <= vm) sum = ctr[u];
#pragma unroll
    for (int v = 1; v <= 15; ++v) {
        if (v <= vm) {
            const int a = ctr[v * sp + u], c = ctr[-v * sp + u];
            sum += a + c;
            m01 += v * (a - c);
        }
    }
    const int m10 = __reduce_add_sync(0xffffffffu, u * sum);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    const float px = l ? __fmul_rn((float)x, L.scale) : (float)x;
    const float py = l ? __fmul_rn((float)y, L.scale) : (float)y;
    if (lane == 0) {
        svo_keypoint k;
        k.x = px; k.y = py; k.size = __fmul_rn(31.f, L.scale); k.angle = angle;
        k.response = b.key2[(size_t)slot * g.total2 + L.off2 + c]; k.octave = l;
        b.kp[(size_t)slot * g.kp_cap + oi] = k;
    }
    // rBRIEF on the blurred level around (cvRound(pt.x/scale), cvRound(pt.y/scale))
    const int cx = __float2int_rn(__fmul_rn(px, L.inv_scale)), cy = __float2int_rn(__fmul_rn(py, L.inv_scale));
    const float ang = __fmul_rn(angle, (float)(3.1415926535897932384626433832795 / 180.f));
    float ca = 0.f, sa = 0.f;
    if (lane == 0) {
        double ds, dc;
        sincos((double)ang, &ds, &dc);
        ca = (float)dc; sa = (float)ds;
    }
    ca = __shfl_sync(0xffffffffu, ca, 0); sa = __shfl_sync(0xffffffffu, sa, 0);
    // The 512 samples of a keypoint fall in the 37 x 37 window around (cx, cy) (the pattern's radius is 18.4, so a
    // rotated and rounded coordinate stays within +-18).  The window is staged in shared memory with word loads laid
    // out along the rows (12 warp loads touching ~75 sectors) and the scattered samples are byte reads from shared
    // memory, instead of 16 global byte gathers per lane that each touch up to 32 sectors (ncu: the gathers kept
    // L1TEX 65 % busy and the warps on the long scoreboard).
    __shared__ uint32_t patch[DESC_WARPS][37 * 10];
    const int a0 = (cx - 18) & ~3;                      // first staged column, word aligned (the level base is 256-B aligned)
    {
        const uint8_t *prow = blr + (size_t)(cy - 18) * sp + a0;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int t = lane + 32 * k;
            if (t < 370) {
                const int row = (t * 205) >> 11, w = t - row * 10;       // t / 10 for t < 384
                patch[warp][t] = *reinterpret_cast<const uint32_t *>(prow + (size_t)row * sp + 4 * w);
            }
        }
    }
    __syncwarp();
    const uint8_t *bc = reinterpret_cast<const uint8_t *>(patch[warp]) + 18 * 40 + (cx - a0);
    int byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float x0 = (float)(int8_t)(pat[j] & 0xff), y0 = (float)(int8_t)((pat[j] >> 8) & 0xff);
        const float x1 = (float)(int8_t)((pat[j] >> 16) & 0xff), y1 = (float)(int8_t)((pat[j] >> 24) & 0xff);
        // cvRound without the XU pipe: |v| < 19, so v + 1.5 * 2^23 rounds to the nearest integer (ties to even, like
        // cvRound / F2I.RN) and the integer sits in the low mantissa bits
        const int ix0 = rint_small(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sa)));
        const int iy0 = rint_small(__fadd_rn(__fmul_rn(x0, sa), __fmul_rn(y0, ca)));
        const int ix1 = rint_small(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sa)));
        const int iy1 = rint_small(__fadd_rn(__fmul_rn(x1, sa), __fmul_rn(y1, ca)));
        const int v0 = bc[iy0 * 40 + ix0], v1 = bc[iy1 * 40 + ix1];
        byte |= (v0 < v1) << j;
    }
    b.desc[((size_t)slot * g.kp_cap + oi) * 32 + lane] = (uint8_t)byte;
}

void launch_harris(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    int mx = 1;
    for (int l = 0; l < g.nlevels; ++l) mx = g.lv[l].quota * 2 + 64 > mx ? g.lv[l].quota * 2 + 64 : mx;
    if (g.harris8) {
        dim3 grid((mx + 4 * DESC_WARPS - 1) / (4 * DESC_WARPS), g.nlevels, nimg);
        k_harris4<<<grid, DESC_WARPS * 32, 0, st>>>(b, g, slot0);
    } else {
        dim3 grid((mx + DESC_WARPS - 1) / DESC_WARPS, g.nlevels, nimg);
        k_harris<<<grid, DESC_WARPS * 32, 0, st>>>(b, g, slot0);
    }
    ++*launches;
}

void launch_blur(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    dim3 grid(g.blur_tiles, nimg);
    if (g.blur_margin) k_blur<true><<<grid, BLUR_THREADS, 0, st>>>(b, g, slot0);
    else k_blur<false><<<grid, BLUR_THREADS, 0, st>>>(b, g, slot0);
    ++*launches;
}

static int *g_pattern_dev[64] = {nullptr};

// The rBRIEF pattern as 256 packed int32 (x0 | y0<<8 | x1<<16 | y1<<24) in global memory of the current device.
static const int *pattern_on_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    if (!g_pattern_dev[dev]) {
        static const int8_t host_pat[1024] = {
#include "../../include/svo_orb_pattern.inc"
        };
        int packed[256];
        for (int t = 0; t < 256; ++t)
            packed[t] = (int)((uint32_t)(uint8_t)host_pat[4 * t] | ((uint32_t)(uint8_t)host_pat[4 * t + 1] << 8) |
                              ((uint32_t)(uint8_t)host_pat[4 * t + 2] << 16) | ((uint32_t)(uint8_t)host_pat[4 * t + 3] << 24));
        int *p = nullptr;
        if (cudaMalloc((void **)&p, sizeof(packed)) != cudaSuccess) return nullptr;
        cudaMemcpy(p, packed, sizeof(packed), cudaMemcpyHostToDevice);
        g_pattern_dev[dev] = p;
    }
    return g_pattern_dev[dev];
}

int setup_describe() { return pattern_on_device() ? 0 : 1; }

void launch_describe(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    dim3 grid((g.kp_cap + DESC_WARPS - 1) / DESC_WARPS, nimg);
    k_describe<<<grid, DESC_WARPS * 32, 0, st>>>(b, g, slot0, pattern_on_device());
    ++*launches;
}
