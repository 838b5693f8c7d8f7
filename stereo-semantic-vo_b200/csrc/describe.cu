// describe.cu — Harris response, intensity-centroid angle, 7x7 Gaussian blur and rBRIEF-256,
// as cv::ORB computes them inside frame::featuredetect (src/frame.cc:75-79).
// SURVEY.md A.5 (Harris, IC angle, fastAtan2), A.7 (blur), A.8 (rBRIEF).
//
// Float bit-exactness: every float operation below is written with the _rn intrinsics so
// nvcc cannot contract a multiply-add into an FMA; OpenCV's portable path rounds each
// operation separately and that is what the oracle is pinned to.
#include "svo_internal.cuh"
#include <float.h>

#define DESC_WARPS 8

__constant__ int8_t c_pattern[256 * 4] = {
#include "../../include/svo_orb_pattern.inc"
};
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
    for (int c = blockIdx.x * DESC_WARPS + warp; c < n; c += gridDim.x * DESC_WARPS) {
        const uint32_t e = cval[c];
        const int x = unpack_x(e), y = unpack_y(e);
        int a = 0, bb = 0, cc = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = lane + 32 * r;
            if (k < 49) {
                const int dy = k / 7 - 3, dx = k % 7 - 3;
                const uint8_t *p = img + (size_t)(y + dy) * sp + (x + dx);
                const int p00 = p[-sp - 1], p01 = p[-sp], p02 = p[-sp + 1];
                const int p10 = p[-1], p12 = p[1];
                const int p20 = p[sp - 1], p21 = p[sp], p22 = p[sp + 1];
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

// ---------------------------------------------------------------------------------------
// Blur: u8 -> f32 row pass (taps left to right) -> f32 column pass (centre, then symmetric
// pairs) -> rint -> u8, reflect-101 borders.  One CTA per 128x16 tile, staged in shared memory.
// ---------------------------------------------------------------------------------------
#define BLUR_TW 128
#define BLUR_TH 16

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(256) k_blur(Bufs b, Geom g, int slot0)
{
    __shared__ uint8_t s_in[BLUR_TH + 6][BLUR_TW + 8];
    __shared__ float s_row[BLUR_TH + 6][BLUR_TW];
    int tile = blockIdx.x, l = 0;
    while (l + 1 < g.nlevels && tile >= g.lv[l + 1].blur_tile_off) ++l;
    const LevelGeom &L = g.lv[l];
    tile -= L.blur_tile_off;
    const int ty = tile / L.blur_tiles_x, tx = tile - ty * L.blur_tiles_x;
    const int x0 = tx * BLUR_TW, y0 = ty * BLUR_TH;
    const int slot = slot0 + blockIdx.y;
    const uint8_t *img = b.pyr + (size_t)slot * g.pyr_bytes + L.off;
    uint8_t *out = b.blur + (size_t)slot * g.pyr_bytes + L.off;
    const int tid = threadIdx.x;
    for (int i = tid; i < (BLUR_TH + 6) * (BLUR_TW + 6); i += 256) {
        const int r = i / (BLUR_TW + 6), c = i - r * (BLUR_TW + 6);
        const int yy = reflect101(y0 - 3 + r, L.h), xx = reflect101(x0 - 3 + c, L.w);
        s_in[r][c] = img[(size_t)yy * L.pitch + xx];
    }
    __syncthreads();
    float gk[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) gk[k] = __uint_as_float(c_gauss[k]);
    for (int i = tid; i < (BLUR_TH + 6) * BLUR_TW; i += 256) {
        const int r = i / BLUR_TW, c = i - r * BLUR_TW;
        float acc = __fmul_rn(gk[0], (float)s_in[r][c]);
#pragma unroll
        for (int k = 1; k < 7; ++k) acc = __fadd_rn(acc, __fmul_rn(gk[k], (float)s_in[r][c + k]));
        s_row[r][c] = acc;
    }
    __syncthreads();
    for (int i = tid; i < BLUR_TH * BLUR_TW; i += 256) {
        const int r = i / BLUR_TW, c = i - r * BLUR_TW;
        const int x = x0 + c, y = y0 + r;
        if (x >= L.w || y >= L.h) continue;
        float acc = __fmul_rn(gk[3], s_row[r + 3][c]);
#pragma unroll
        for (int k = 1; k <= 3; ++k)
            acc = __fadd_rn(acc, __fmul_rn(gk[3 + k], __fadd_rn(s_row[r + 3 + k][c], s_row[r + 3 - k][c])));
        int v = __float2int_rn(acc);
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        out[(size_t)y * L.pitch + x] = (uint8_t)v;
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

__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe(Bufs b, Geom g, int slot0)
{
    __shared__ int8_t s_pat[4][8][32];  // [coord][bit j][lane]: conflict-free per-lane reads
    const int l = blockIdx.y, slot = slot0 + blockIdx.z;
    const LevelGeom &L = g.lv[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += DESC_WARPS * 32) {
        const int test = i >> 2, cidx = i & 3;       // test = 8*byte + bit
        s_pat[cidx][test & 7][test >> 3] = c_pattern[i];
    }
    __syncthreads();
    const int *kept2 = b.kept2 + (size_t)slot * SVO_MAX_LEVELS;
    int base = 0, total = 0;
    for (int i = 0; i < g.nlevels; ++i) {
        const int k = kept2[i];
        if (i < l) base += k;
        total += k;
    }
    if (blockIdx.x == 0 && l == 0 && threadIdx.x == 0) {
        b.nkp[slot] = total;
        if (total > g.kp_cap) atomicOr(b.status + slot, SVO_STATUS_OVERFLOW);
    }
    const int n = kept2[l];
    const uint8_t *img = b.pyr + (size_t)slot * g.pyr_bytes + L.off;
    const uint8_t *blr = b.blur + (size_t)slot * g.pyr_bytes + L.off;
    const float *key2 = b.key2 + (size_t)slot * g.total2 + L.off2;
    const uint32_t *val2 = b.val2 + (size_t)slot * g.total2 + L.off2;
    svo_keypoint *kp = b.kp + (size_t)slot * g.kp_cap;
    uint8_t *desc = b.desc + (size_t)slot * g.kp_cap * 32;
    const int sp = L.pitch;
    for (int c = blockIdx.x * DESC_WARPS + warp; c < n; c += gridDim.x * DESC_WARPS) {
        const int oi = base + c;
        if (oi >= g.kp_cap) continue;
        const uint32_t e = val2[c];
        const int x = unpack_x(e), y = unpack_y(e);
        // intensity centroid over the radius-15 disc
        const uint8_t *ctr = img + (size_t)y * sp + x;
        int m10 = 0, m01 = 0;
        const int u = lane - 15;
        if (lane < 31) {
#pragma unroll 1
            for (int v = -15; v <= 15; ++v) {
                const int av = v < 0 ? -v : v;
                if (u >= -c_umax[av] && u <= c_umax[av]) {
                    const int val = ctr[v * sp + u];
                    m10 += u * val;
                    m01 += v * val;
                }
            }
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        // keypoint record
        const float px = l ? __fmul_rn((float)x, L.scale) : (float)x;
        const float py = l ? __fmul_rn((float)y, L.scale) : (float)y;
        if (lane == 0) {
            svo_keypoint k;
            k.x = px; k.y = py; k.size = __fmul_rn(31.f, L.scale); k.angle = angle;
            k.response = key2[c]; k.octave = l;
            kp[oi] = k;
        }
        // rBRIEF on the blurred level around (cvRound(pt.x/scale), cvRound(pt.y/scale))
        const int cx = __float2int_rn(__fmul_rn(px, L.inv_scale)), cy = __float2int_rn(__fmul_rn(py, L.inv_scale));
        const float ang = __fmul_rn(angle, (float)(3.1415926535897932384626433832795 / 180.f));
        float ca = 0.f, sa = 0.f;
        if (lane == 0) { ca = (float)cos((double)ang); sa = (float)sin((double)ang); }
        ca = __shfl_sync(0xffffffffu, ca, 0); sa = __shfl_sync(0xffffffffu, sa, 0);
        const uint8_t *bc = blr + (size_t)cy * sp + cx;
        int byte = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x0 = (float)s_pat[0][j][lane], y0 = (float)s_pat[1][j][lane];
            const float x1 = (float)s_pat[2][j][lane], y1 = (float)s_pat[3][j][lane];
            const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sa)));
            const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sa), __fmul_rn(y0, ca)));
            const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sa)));
            const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sa), __fmul_rn(y1, ca)));
            const int v0 = bc[iy0 * sp + ix0], v1 = bc[iy1 * sp + ix1];
            byte |= (v0 < v1) << j;
        }
        desc[(size_t)oi * 32 + lane] = (uint8_t)byte;
    }
}

void launch_harris(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    int mx = 1;
    for (int l = 0; l < g.nlevels; ++l) mx = g.lv[l].quota * 2 + 64 > mx ? g.lv[l].quota * 2 + 64 : mx;
    dim3 grid((mx + DESC_WARPS - 1) / DESC_WARPS, g.nlevels, nimg);
    k_harris<<<grid, DESC_WARPS * 32, 0, st>>>(b, g, slot0);
    ++*launches;
}

void launch_blur(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    dim3 grid(g.blur_tiles, nimg);
    k_blur<<<grid, 256, 0, st>>>(b, g, slot0);
    ++*launches;
}

void launch_describe(const Bufs &b, const Geom &g, int slot0, int nimg, cudaStream_t st, long long *launches)
{
    int mx = 1;
    for (int l = 0; l < g.nlevels; ++l) mx = g.lv[l].quota + 32 > mx ? g.lv[l].quota + 32 : mx;
    dim3 grid((mx + DESC_WARPS - 1) / DESC_WARPS, g.nlevels, nimg);
    k_describe<<<grid, DESC_WARPS * 32, 0, st>>>(b, g, slot0);
    ++*launches;
}
