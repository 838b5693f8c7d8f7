// resize_quads.h — host-side table of the per-level resize kernel (pyramid.cu:k_resize), plain C++.
//
// cv::resize(INTER_LINEAR_EXACT) reads, for output column d, the source bytes ofs[d] and ofs[d] + 1 with the Q8 weights
// 256 - w1[d] and w1[d] (the per-column table (ofs << 9) | w1 is built by resize_table below).
// Four adjacent output columns of a pyramid level (scale <= 2) read source bytes that lie within 8 consecutive bytes,
// so a thread that produces an output quad fetches three aligned words, shifts them into an 8-byte window that starts
// at ofs[first column] and picks its eight taps with two byte permutes; a two-way dot product (dp2a: two u16 weights
// times two bytes) then forms each column's Q8 sum.  One entry of this table describes one output quad:
//     word 0   byte offset of the first aligned word in the source row | (8 * (ofs[0] & 3)) << 16   (funnel shift)
//     word 1   permute selector of columns 0, 1 (tap0[0], tap1[0], tap0[1], tap1[1]) | selector of columns 2, 3 << 16
//     word 2-5 per column (256 - w1) | w1 << 16; 0 for the padding columns past the level's width
//     word 6-7 unused (32-byte entries)
// Integer arithmetic identical to the per-byte form: h = p[ofs] * (256 - w1) + p[ofs + 1] * w1.
#pragma once
#include <math.h>
#include <stdint.h>

// INTER_LINEAR_EXACT coefficients: packed (ofs << 9) | w1, w1 in Q8
static inline void resize_table(int src, int dst, uint32_t *tab)
{
    const double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; ++d) {
        const double fv = scale * ((double)d + 0.5) - 0.5;
        int iv = (int)floor(fv);
        int w1 = 0;
        if (iv >= 0 && src > 1) {
            if (iv < src - 1) w1 = (int)lrint((fv - (double)iv) * 256.0);
            else iv = src - 1;
        } else iv = 0;
        tab[d] = ((uint32_t)iv << 9) | (uint32_t)w1;
    }
}

struct ResizeQuad { uint32_t base_shift, sel, w[4], pad[2]; };

// xtab: the level's column table ((ofs << 9) | w1), dw columns; dpitch: padded row bytes of the level (multiple of 4);
// out: dpitch / 4 entries.  Returns 0 when some quad's taps do not fit the 8-byte window (the caller then keeps the
// per-byte kernel for this level), 1 otherwise.
static inline int build_resize_quads(const uint32_t *xtab, int dw, int dpitch, ResizeQuad *out)
{
    int ok = 1;
    for (int q = 0; q < dpitch / 4; ++q) {
        ResizeQuad e = {0u, 0u, {0u, 0u, 0u, 0u}, {0u, 0u}};
        const int x = 4 * q;
        if (x < dw) {
            const int first = (int)(xtab[x] >> 9);
            uint32_t sel[4] = {0u, 0u, 0u, 0u};
            for (int k = 0; k < 4 && x + k < dw; ++k) {
                const uint32_t t = xtab[x + k];
                const int i0 = (int)(t >> 9), w1 = (int)(t & 511u);
                const int i1 = w1 ? i0 + 1 : i0;                 // a zero weight never reads past the row (ofs = src - 1 there)
                const int o0 = i0 - first, o1 = i1 - first;
                if (o0 < 0 || o1 > 7) { ok = 0; continue; }
                sel[k] = (uint32_t)o0 | ((uint32_t)o1 << 4);
                e.w[k] = (uint32_t)(256 - w1) | ((uint32_t)w1 << 16);
            }
            e.base_shift = (uint32_t)(first & ~3) | ((uint32_t)(8 * (first & 3)) << 16);
            e.sel = sel[0] | (sel[1] << 8) | (sel[2] << 16) | (sel[3] << 24);
        }
        out[q] = e;
    }
    return ok;
}
