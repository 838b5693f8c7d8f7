// CPU model of pyramid.cu:k_resize_q — the SAME host table code (csrc/resize_quads.h) driven through plain-C++
// restatements of the three device intrinsics the kernel uses (__funnelshift_r, __byte_perm, __dp2a_lo/_hi), checked
// against the per-byte form of cv::resize(INTER_LINEAR_EXACT) that k_resize computes.
// Test infrastructure: built and run by tests/test_resize_quads.py; exports resize_model() for ctypes.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../stereo-semantic-vo_b200/csrc/resize_quads.h"

static inline uint32_t funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh & 31u));
}
static inline uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (s >> (4 * i)) & 0xfu;
        uint32_t b = (uint32_t)(v >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) b = (b & 0x80u) ? 0xffu : 0u;      // PTX prmt default mode: msb of the nibble replicates the sign
        r |= b << (8 * i);
    }
    return r;
}
static inline uint32_t dp2a(uint32_t a, uint32_t b, uint32_t c, int hi)
{
    const uint32_t b0 = (b >> (hi ? 16 : 0)) & 0xffu, b1 = (b >> (hi ? 24 : 8)) & 0xffu;
    return c + (a & 0xffffu) * b0 + (a >> 16) * b1;
}

// src: sh x spitch bytes (+ one slack row), level sw x sh -> dst dw x dh with pitch dpitch, both ways.
// Returns -1 when the quad table does not apply (rq_ok == 0), else the number of differing bytes between the two forms
// (padding columns included: both must hold 0).  dst_out (dh x dpitch) receives the quad-table result.
extern "C" int resize_model(const uint8_t *src, int sw, int sh, int spitch, int dw, int dh, int dpitch, uint8_t *dst_out, int rs)
{
    std::vector<uint32_t> tab((size_t)dw + dh);
    resize_table(sw, dw, tab.data());
    resize_table(sh, dh, tab.data() + dw);
    std::vector<ResizeQuad> rq((size_t)dpitch / 4);
    if (!build_resize_quads(tab.data(), dw, dpitch, rq.data())) return -1;
    std::vector<uint8_t> ref((size_t)dh * dpitch, 0);
    for (int y = 0; y < dh; ++y) {
        const uint32_t ty = tab[(size_t)dw + y];
        const int yo = (int)(ty >> 9), wy1 = (int)(ty & 511u), wy0 = 256 - wy1, yn = wy1 ? yo + 1 : yo;
        for (int x = 0; x < dw; ++x) {
            const uint32_t tx = tab[x];
            const int i0 = (int)(tx >> 9), wx1 = (int)(tx & 511u), wx0 = 256 - wx1, i1 = wx1 ? i0 + 1 : i0;
            const uint32_t ha = src[(size_t)yo * spitch + i0] * wx0 + src[(size_t)yo * spitch + i1] * wx1;
            const uint32_t hb = src[(size_t)yn * spitch + i0] * wx0 + src[(size_t)yn * spitch + i1] * wx1;
            ref[(size_t)y * dpitch + x] = (uint8_t)((ha * wy0 + hb * wy1 + 32768u) >> 16);
        }
    }
    // the kernel, thread by thread: strips of rs rows, two register sets with alternating roles
    const int qpr = dpitch / 4, strips = (dh + rs - 1) / rs;
    memset(dst_out, 0xee, (size_t)dh * dpitch);
    for (int q = 0; q < qpr * strips; ++q) {
        const int st = q / qpr, xq = q - st * qpr;
        const int y0 = st * rs, y1 = y0 + rs < dh ? y0 + rs : dh;
        const ResizeQuad &e = rq[xq];
        const uint32_t shf = e.base_shift >> 16, sel01 = e.sel & 0xffffu, sel23 = e.sel >> 16;
        const uint8_t *srcq = src + (e.base_shift & 0xffffu);
        uint32_t h[2][4];
        int held[2] = {-1, -1};
        memset(h, 0xcd, sizeof(h));
        auto hrow = [&](int yy, uint32_t *o) {
            uint32_t w[3];
            memcpy(w, srcq + (size_t)yy * spitch, 12);
            const uint32_t lo = funnelshift_r(w[0], w[1], shf), hi = funnelshift_r(w[1], w[2], shf);
            const uint32_t p01 = byte_perm(lo, hi, sel01), p23 = byte_perm(lo, hi, sel23);
            o[0] = dp2a(e.w[0], p01, 0, 0); o[1] = dp2a(e.w[1], p01, 0, 1);
            o[2] = dp2a(e.w[2], p23, 0, 0); o[3] = dp2a(e.w[3], p23, 0, 1);
        };
        for (int y = y0; y < y1; ++y) {
            const int U = (y - y0) & 1, V = U ^ 1;
            const uint32_t ty = tab[(size_t)dw + y];
            const int yo = (int)(ty >> 9);
            const uint32_t wy1 = ty & 511u, wy0 = 256u - wy1;
            if (held[U] != yo) { hrow(yo, h[U]); held[U] = yo; }
            if (wy1 && held[V] != yo + 1) { hrow(yo + 1, h[V]); held[V] = yo + 1; }
            uint32_t s[4];
            for (int k = 0; k < 4; ++k) s[k] = h[U][k] * wy0 + h[V][k] * wy1 + 32768u;
            const uint32_t out = byte_perm(byte_perm(s[0], s[1], 0x0062), byte_perm(s[2], s[3], 0x0062), 0x5410);
            memcpy(dst_out + (size_t)y * dpitch + 4 * xq, &out, 4);
        }
    }
    int diff = 0;
    for (size_t i = 0; i < ref.size(); ++i) diff += ref[i] != dst_out[i];
    return diff;
}
