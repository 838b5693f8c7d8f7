// pyramid.cu — cv::ORB's image pyramid: level l = resize(level l-1, INTER_LINEAR_EXACT)
// (8-bit fixed point: Q8 horizontal, Q8 vertical, round at Q16).  SURVEY.md A.2;
// called from frame::featuredetect (src/frame.cc:75-79) via cv::ORB.
//
// HBM-bound byte work: per output quad one thread reads two source rows through
// per-level coefficient tables ((ofs << 9) | w1, built on the host in double exactly as
// OpenCV does) and stores one aligned 32-bit word.  Algorithmic bytes per level:
// px(l-1) read + px(l) written.
#include "svo_internal.cuh"

__global__ void __launch_bounds__(256) k_resize(Bufs b, Geom g, int l, int slot0)
{
    const LevelGeom &d = g.lv[l];
    const LevelGeom &s = g.lv[l - 1];
    const int qpr = d.pitch >> 2;  // quads per (padded) row
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= qpr * d.h) return;
    const int y = q / qpr, x = (q - y * qpr) << 2;
    const size_t base = (size_t)(slot0 + blockIdx.y) * g.pyr_bytes;
    const uint8_t *src = b.pyr + base + s.off;
    uint8_t *dst = b.pyr + base + d.off;
    const uint32_t *xt = b.rtab + d.tab_off;
    const uint32_t ty = xt[d.w + y];
    const int yo = ty >> 9;
    const uint32_t wy1 = ty & 511u, wy0 = 256u - wy1;
    const uint8_t *r0 = src + (size_t)yo * s.pitch;
    const uint8_t *r1 = wy1 ? r0 + s.pitch : r0;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x + k;
        if (xx < d.w) {
            const uint32_t tx = xt[xx];
            const int i0 = tx >> 9;
            const uint32_t wx1 = tx & 511u, wx0 = 256u - wx1;
            const int i1 = wx1 ? i0 + 1 : i0;
            const uint32_t h0 = r0[i0] * wx0 + r0[i1] * wx1;
            const uint32_t h1 = r1[i0] * wx0 + r1[i1] * wx1;
            const uint32_t v = (h0 * wy0 + h1 * wy1 + 32768u) >> 16;
            out |= v << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t *>(dst + (size_t)y * d.pitch + x) = out;
}

// Level 0 is read where it lies in device memory (the caller's device buffer, or the lane's landing zone
// for host inputs; rows `stride` bytes apart, any alignment) and repacked into the 16-byte-pitched level-0
// layout.  One thread per 16 output bytes.
__global__ void __launch_bounds__(256) k_unpack(Bufs b, Geom g, int slot0, const FramePtrs *__restrict__ fp,
                                                const int *__restrict__ strides)
{
    const LevelGeom &L = g.lv[0];
    const int per_row = L.pitch >> 4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_row * L.h) return;
    const int y = i / per_row, x = (i - y * per_row) << 4;
    const int stride = strides[blockIdx.y];
    if (stride <= 0) return;                      // this image was uploaded with a 2-D copy
    const FramePtrs &F = fp[blockIdx.y >> 1];
    const uint8_t *src = ((blockIdx.y & 1) ? F.right : F.left) + (size_t)y * stride + x;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (x + k < L.w) w[k >> 2] |= (uint32_t)src[k] << (8 * (k & 3));
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
    for (int l = 1; l < g.nlevels; ++l) {
        const int quads = (g.lv[l].pitch >> 2) * g.lv[l].h;
        dim3 grid((quads + 255) / 256, nimg);
        k_resize<<<grid, 256, 0, st>>>(b, g, l, slot0);
        ++*launches;
    }
}
