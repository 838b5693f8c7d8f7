/*
 * svo_oracle.c — CPU restatement of the per-frame stereo front-end of
 * zssjh/stereo-semantic-vo.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (libsvo_b200.so) never
 * does and has no CPU fallback.
 *
 * What it restates (reference file:line, relative to /root/reference):
 *   - frame::featuredetect            src/frame.cc:75-79   -> cv::ORB::detectAndCompute.
 *     The arithmetic lives in OpenCV features2d (un-vendored; the author linked
 *     3.2.0, Thirdparty/MB/build/CMakeCache.txt:515).  The runnable stand-in in
 *     this image is opencv-python-headless 4.13.0; the algorithm below follows
 *     its published ORB and is pinned bit-for-bit against
 *     cv2.ORB_create(...).detectAndCompute run with cv2.setUseOptimized(False)
 *     (OpenCV's portable scalar code path; the SIMD path contracts the blur into
 *     FMAs on a CPU-dependent subset of columns and is not a stable definition)
 *     by tests/test_oracle_vs_cv2.py and the fixtures in tests/golden/.
 *   - pnpmatch::DescriptorDistance    src/pnpmatch.cc:14-30
 *   - BFMatcher + distance filter     src/pnpmatch.cc:266-299
 *   - greedy scans of poseEstimationPnP  src/pnpmatch.cc:75-95 (pass 1, accept
 *     :99-101,139-153) and :173-197 (pass 2)
 *   - frame::computekeypoint_r / disp2Depth   src/frame.cc:122-164
 *   - sparse stereo + SAD refinement: NOT in the reference (it runs dense MSA,
 *     src/Tracking.cc:226).  north_star asks for ORB-SLAM2-lineage
 *     ComputeStereoMatches; this file DEFINES it.
 *     PARITY UNPINNED for that stage: no reference code or vectors exist.
 *   - the pose stage (svo_pose_oracle.c) and the opt-in quadtree keypoint distribution
 *     (svo_octree_oracle.c) live in their own files; each header states what pins it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no -march=native so
 * every float op is separately rounded, which is what reproduces cv2's bits).
 * retainBest's ordering uses libstdc++ std::nth_element and lives in
 * retain_best.cpp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include "svo_oracle.h"

static const int8_t k_pattern[256 * 4] = {
#include "../include/svo_orb_pattern.inc"
};

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

/* ------------------------------------------------------------------ */
/* A.1 geometry: level sizes, scales and per-level quotas              */
/* ------------------------------------------------------------------ */
void svo_o_geometry(int W, int H, int nlevels, float scale_factor_f, int nfeatures,
                    int *lw, int *lh, float *lscale, int *quota)
{
    double scale_factor = (double)scale_factor_f; /* ORB stores the float ctor arg as double */
    for (int l = 0; l < nlevels; ++l) {
        float s = (float)pow(scale_factor, (double)l);
        lscale[l] = s;
        /* cvRound(image.cols / scale): the OpenCV build this oracle is pinned to (cv2 4.13.0) evaluates the quotient as a
         * float multiplication by the reciprocal of the scale, which differs from the true quotient only when cols / scale
         * falls within a float ulp of k + 0.5 — at 1.2 every level of 140 of the dimensions below 4096 (249 -> 208, not
         * 207); found by tools/fuzz_orb_cv2.py, rule fitted on 34 probed sizes, none of KITTI's is among them. */
        volatile float inv = 1.0f / s;
        lw[l] = cv_round_f((float)W * inv);
        lh[l] = cv_round_f((float)H * inv);
    }
    float factor = (float)(1.0 / scale_factor);
    float ndes = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        quota[l] = cv_round_f(ndes);
        sum += quota[l];
        ndes *= factor;
    }
    quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

/* ------------------------------------------------------------------ */
/* A.2 INTER_LINEAR_EXACT resize (8-bit, 1 channel), Q8 x Q8 -> Q16    */
/* ------------------------------------------------------------------ */
void svo_o_resize_coeffs(int src, int dst, int *ofs, int *w1)
{
    /* OpenCV (resize.cpp, interpolationLinear) forms inv_scale = dst / src first and divides one by it: two roundings, not
     * src / dst.  It matters only where a coefficient is an exact tie ((fv - iv) * 256 = k + 0.5 in exact arithmetic:
     * v2(dst) - v2(src) = 8), e.g. 3993 -> 3328, the one transition among the pyramid chains of every image dimension up
     * to 4095 at 1.2 where the two scales give different taps (found by sweeping tests/host_models' resize model). */
    volatile double inv_scale = (double)dst / (double)src;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        double fv = scale * ((double)d + 0.5) - 0.5;
        int iv = (int)floor(fv);
        if (iv >= 0 && src > 1) {
            if (iv < src - 1) {
                ofs[d] = iv;
                w1[d] = cv_round_d((fv - (double)iv) * 256.0);
            } else {
                ofs[d] = src - 1;
                w1[d] = 0;
            }
        } else {
            ofs[d] = 0;
            w1[d] = 0;
        }
    }
}

void svo_o_resize(const uint8_t *src, int sw, int sh, int sstride,
                  uint8_t *dst, int dw, int dh, int dstride)
{
    int *xo = (int *)malloc(sizeof(int) * (size_t)dw * 2);
    int *xw = xo + dw;
    int *yo = (int *)malloc(sizeof(int) * (size_t)dh * 2);
    int *yw = yo + dh;
    svo_o_resize_coeffs(sw, dw, xo, xw);
    svo_o_resize_coeffs(sh, dh, yo, yw);
    for (int y = 0; y < dh; ++y) {
        const uint8_t *r0 = src + (size_t)yo[y] * sstride;
        const uint8_t *r1 = src + (size_t)(yw[y] ? yo[y] + 1 : yo[y]) * sstride;
        uint32_t wy1 = (uint32_t)yw[y], wy0 = 256u - wy1;
        for (int x = 0; x < dw; ++x) {
            int i0 = xo[x], i1 = xw[x] ? i0 + 1 : i0;
            uint32_t wx1 = (uint32_t)xw[x], wx0 = 256u - wx1;
            uint32_t h0 = r0[i0] * wx0 + r0[i1] * wx1;
            uint32_t h1 = r1[i0] * wx0 + r1[i1] * wx1;
            uint32_t v = h0 * wy0 + h1 * wy1;
            dst[(size_t)y * dstride + x] = (uint8_t)((v + 32768u) >> 16);
        }
    }
    free(xo);
    free(yo);
}

/* ------------------------------------------------------------------ */
/* A.3 FAST-9/16 score, 3x3 strict NMS, raster order, border cull      */
/* ------------------------------------------------------------------ */
static const int k_circle[16][2] = {
    {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
    {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

int svo_o_fast_score(const uint8_t *p, int stride)
{
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; ++k)
        d[k] = v - p[k_circle[k][1] * stride + k_circle[k][0]];
    for (int k = 16; k < 25; ++k)
        d[k] = d[k - 16];
    int best_b = -255, best_d = -255;
    for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; ++k) {
            if (d[s + k] < mn) mn = d[s + k];
            if (d[s + k] > mx) mx = d[s + k];
        }
        if (mn > best_b) best_b = mn;       /* centre brighter than the whole arc by mn */
        if (-mx > best_d) best_d = -mx;     /* centre darker than the whole arc by -mx  */
    }
    return (best_b > best_d ? best_b : best_d) - 1;
}

/* score map: score if corner (score >= threshold) else 0; FAST skips a 3-px frame */
static void fast_score_map(const uint8_t *img, int w, int h, int stride, int threshold, uint8_t *sc)
{
    memset(sc, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = svo_o_fast_score(img + (size_t)y * stride + x, stride);
            if (s >= threshold) sc[(size_t)y * w + x] = (uint8_t)s;
        }
}

int svo_o_fast_nms(const uint8_t *img, int w, int h, int stride, int threshold, int border,
                   int32_t *xs, int32_t *ys, int32_t *scores, int cap)
{
    uint8_t *sc = (uint8_t *)malloc((size_t)w * h);
    fast_score_map(img, w, h, stride, threshold, sc);
    int n = 0;
    /* OpenCV's FAST never emits the outermost scored ring row h-4..: rows are
     * emitted for y in [3, h-4]; all of that lies inside `border` (>= 4). */
    int b = border < 4 ? 4 : border;
    for (int y = b; y < h - b; ++y)
        for (int x = b; x < w - b; ++x) {
            int s = sc[(size_t)y * w + x];
            if (!s) continue;
            const uint8_t *q = sc + (size_t)y * w + x;
            if (s > q[-1] && s > q[1] && s > q[-w - 1] && s > q[-w] && s > q[-w + 1] &&
                s > q[w - 1] && s > q[w] && s > q[w + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; scores[n] = s; }
                ++n;
            }
        }
    free(sc);
    return n;
}

/* ------------------------------------------------------------------ */
/* A.5 Harris response (7x7 block, k = 0.04) and IC angle              */
/* ------------------------------------------------------------------ */
void svo_o_harris(const uint8_t *img, int stride, const int32_t *xs, const int32_t *ys, int n, float *resp)
{
    const int bs = 7, r = bs / 2;
    const float harris_k = 0.04f;
    float scale = 1.f / ((1 << 2) * bs * 255.f);
    float scale_sq_sq = scale * scale * scale * scale;
    for (int i = 0; i < n; ++i) {
        const uint8_t *p0 = img + (size_t)(ys[i] - r) * stride + (xs[i] - r);
        int a = 0, b = 0, c = 0;
        for (int dy = 0; dy < bs; ++dy)
            for (int dx = 0; dx < bs; ++dx) {
                const uint8_t *p = p0 + dy * stride + dx;
                int Ix = (p[1] - p[-1]) * 2 + (p[-stride + 1] - p[-stride - 1]) + (p[stride + 1] - p[stride - 1]);
                int Iy = (p[stride] - p[-stride]) * 2 + (p[stride - 1] - p[-stride - 1]) + (p[stride + 1] - p[-stride + 1]);
                a += Ix * Ix;
                b += Iy * Iy;
                c += Ix * Iy;
            }
        float fa = (float)a, fb = (float)b, fc = (float)c;
        float t1 = fa * fb;
        float t2 = fc * fc;
        float s = fa + fb;
        float t3 = harris_k * s;
        t3 = t3 * s;
        float v = t1 - t2;
        v = v - t3;
        resp[i] = v * scale_sq_sq;
    }
}

static const int k_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

float svo_o_fast_atan2(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

void svo_o_ic_angle(const uint8_t *img, int stride, const int32_t *xs, const int32_t *ys, int n, float *angle)
{
    for (int i = 0; i < n; ++i) {
        const uint8_t *c = img + (size_t)ys[i] * stride + xs[i];
        int m01 = 0, m10 = 0;
        for (int u = -15; u <= 15; ++u) m10 += u * c[u];
        for (int v = 1; v <= 15; ++v) {
            int vs = 0, d = k_umax[v];
            for (int u = -d; u <= d; ++u) {
                int vp = c[u + v * stride], vm = c[u - v * stride];
                vs += vp - vm;
                m10 += u * (vp + vm);
            }
            m01 += v * vs;
        }
        angle[i] = svo_o_fast_atan2((float)m01, (float)m10);
    }
}

/* ------------------------------------------------------------------ */
/* A.7 blur: u8 -> f32 row pass (taps left to right), f32 column pass  */
/* (centre first, then symmetric pairs), rint -> u8.  Every mul/add    */
/* separately rounded.  Border: reflect-101.                           */
/* ------------------------------------------------------------------ */
static const uint32_t k_gauss_bits[7] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u,
                                         0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u};

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

void svo_o_blur7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride)
{
    float g[7];
    memcpy(g, k_gauss_bits, sizeof(g));
    float *rows = (float *)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (size_t)y * sstride;
        float *r = rows + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc;
            if (x >= 3 && x < w - 3) {
                acc = g[0] * (float)s[x - 3];
                for (int k = 1; k < 7; ++k) {
                    float t = g[k] * (float)s[x - 3 + k];
                    acc = acc + t;
                }
            } else {
                acc = g[0] * (float)s[reflect101(x - 3, w)];
                for (int k = 1; k < 7; ++k) {
                    float t = g[k] * (float)s[reflect101(x - 3 + k, w)];
                    acc = acc + t;
                }
            }
            r[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const float *c0 = rows + (size_t)y * w;
        const float *rp[4], *rm[4];
        for (int k = 1; k <= 3; ++k) {
            rp[k] = rows + (size_t)reflect101(y + k, h) * w;
            rm[k] = rows + (size_t)reflect101(y - k, h) * w;
        }
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            float acc = g[3] * c0[x];
            for (int k = 1; k <= 3; ++k) {
                float pr = rp[k][x] + rm[k][x];
                float t = g[3 + k] * pr;
                acc = acc + t;
            }
            int v = cv_round_f(acc);
            d[x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(rows);
}

/* ------------------------------------------------------------------ */
/* A.8 rBRIEF-256                                                      */
/* ------------------------------------------------------------------ */
void svo_o_brief(const uint8_t *blur, int stride, int cx, int cy, float angle_deg, uint8_t *desc)
{
    float ang = angle_deg * (float)(3.1415926535897932384626433832795 / 180.f);
    float a = (float)cos((double)ang), b = (float)sin((double)ang);
    const uint8_t *c = blur + (size_t)cy * stride + cx;
    for (int i = 0; i < 32; ++i) {
        int byte = 0;
        for (int j = 0; j < 8; ++j) {
            const int8_t *pp = k_pattern + (size_t)(8 * i + j) * 4;
            float x0 = (float)pp[0] * a, t0 = (float)pp[1] * b;
            float y0 = (float)pp[0] * b, u0 = (float)pp[1] * a;
            float x1 = (float)pp[2] * a, t1 = (float)pp[3] * b;
            float y1 = (float)pp[2] * b, u1 = (float)pp[3] * a;
            int ix0 = cv_round_f(x0 - t0), iy0 = cv_round_f(y0 + u0);
            int ix1 = cv_round_f(x1 - t1), iy1 = cv_round_f(y1 + u1);
            int v0 = c[iy0 * stride + ix0], v1 = c[iy1 * stride + ix1];
            byte |= (v0 < v1) << j;
        }
        desc[i] = (uint8_t)byte;
    }
}

/* ------------------------------------------------------------------ */
/* whole extractor: cv::ORB::detectAndCompute (src/frame.cc:75-79)     */
/* ------------------------------------------------------------------ */
typedef struct {
    int w, h, stride;
    uint8_t *img, *blur;
    float scale;
} level_t;

int svo_o_orb(const uint8_t *gray, int W, int H, int stride, int nfeatures, float scale_factor,
              int nlevels, int fast_threshold, svo_o_keypoint *kps, uint8_t *desc, int cap,
              svo_o_pyramid *pyr_out)
{
    return svo_o_orb_ex(gray, W, H, stride, nfeatures, scale_factor, nlevels, fast_threshold, 0, kps, desc, cap, pyr_out);
}

/* distribution 0: cv::ORB's two retainBest culls (the reference's behaviour).
 * distribution 1 (opt-in, parity unpinned): the level's FAST corners go through the quadtree distribution of
 * svo_octree_oracle.c with N = the level quota; the survivors keep node (Z) order, get their Harris response
 * (reported, not used for selection) and the rest of the pipeline is unchanged. */
int svo_o_orb_ex(const uint8_t *gray, int W, int H, int stride, int nfeatures, float scale_factor,
                 int nlevels, int fast_threshold, int distribution, svo_o_keypoint *kps, uint8_t *desc, int cap,
                 svo_o_pyramid *pyr_out)
{
    const int edge = 31;
    int lw[SVO_O_MAX_LEVELS], lh[SVO_O_MAX_LEVELS], quota[SVO_O_MAX_LEVELS];
    float ls[SVO_O_MAX_LEVELS];
    level_t L[SVO_O_MAX_LEVELS];
    if (nlevels > SVO_O_MAX_LEVELS) return -1;
    svo_o_geometry(W, H, nlevels, scale_factor, nfeatures, lw, lh, ls, quota);
    for (int l = 0; l < nlevels; ++l) {
        L[l].w = lw[l]; L[l].h = lh[l]; L[l].stride = lw[l]; L[l].scale = ls[l];
        L[l].img = (uint8_t *)malloc((size_t)lw[l] * lh[l]);
        L[l].blur = (uint8_t *)malloc((size_t)lw[l] * lh[l]);
        if (l == 0)
            for (int y = 0; y < H; ++y) memcpy(L[0].img + (size_t)y * W, gray + (size_t)y * stride, (size_t)W);
        else
            svo_o_resize(L[l - 1].img, lw[l - 1], lh[l - 1], lw[l - 1], L[l].img, lw[l], lh[l], lw[l]);
    }
    int total = 0;
    for (int l = 0; l < nlevels; ++l) {
        int w = L[l].w, h = L[l].h;
        if (w <= 2 * edge || h <= 2 * edge) continue;
        int capl = ((w + 1) / 2) * ((h + 1) / 2);
        int32_t *xs = (int32_t *)malloc(sizeof(int32_t) * (size_t)capl * 4);
        int32_t *ys = xs + capl, *sc = ys + capl, *idx = sc + capl;
        float *resp = (float *)malloc(sizeof(float) * (size_t)capl);
        int n = svo_o_fast_nms(L[l].img, w, h, L[l].stride, fast_threshold, edge, xs, ys, sc, capl);
        for (int i = 0; i < n; ++i) { idx[i] = i; resp[i] = (float)sc[i]; }
        if (distribution == 1) n = svo_o_distribute_octree(xs, ys, sc, n, edge, edge, w - edge, h - edge, quota[l], idx);
        else n = svo_o_retain_best(resp, idx, n, 2 * quota[l]);
        /* Harris on the survivors, in their post-retainBest order */
        int32_t *hx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1) * 3);
        int32_t *hy = hx + n + 1, *hidx = hy + n + 1;
        for (int i = 0; i < n; ++i) { hx[i] = xs[idx[i]]; hy[i] = ys[idx[i]]; hidx[i] = i; }
        svo_o_harris(L[l].img, L[l].stride, hx, hy, n, resp);
        int m = distribution == 1 ? n : svo_o_retain_best(resp, hidx, n, quota[l]);
        float *ang = (float *)malloc(sizeof(float) * (size_t)(m + 1));
        int32_t *fx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m + 1) * 2);
        int32_t *fy = fx + m + 1;
        for (int i = 0; i < m; ++i) { fx[i] = hx[hidx[i]]; fy[i] = hy[hidx[i]]; }
        svo_o_ic_angle(L[l].img, L[l].stride, fx, fy, m, ang);
        svo_o_blur7(L[l].img, w, h, L[l].stride, L[l].blur, L[l].stride);
        for (int i = 0; i < m; ++i) {
            if (total < cap) {
                svo_o_keypoint *k = &kps[total];
                float s = L[l].scale;
                /* OpenCV multiplies pt by the layer scale for every level but 0 */
                k->x = l ? (float)fx[i] * s : (float)fx[i];
                k->y = l ? (float)fy[i] * s : (float)fy[i];
                k->size = 31.f * s;
                k->angle = ang[i];
                k->response = resp[i];
                k->octave = l;
                float inv = 1.f / s;
                int cx = cv_round_f(k->x * inv), cy = cv_round_f(k->y * inv);
                svo_o_brief(L[l].blur, L[l].stride, cx, cy, ang[i], desc + (size_t)total * 32);
            }
            ++total;
        }
        free(xs); free(resp); free(hx); free(ang); free(fx);
    }
    if (pyr_out) {
        pyr_out->nlevels = nlevels;
        for (int l = 0; l < nlevels; ++l) {
            pyr_out->w[l] = L[l].w; pyr_out->h[l] = L[l].h; pyr_out->scale[l] = L[l].scale;
            pyr_out->img[l] = L[l].img; pyr_out->blur[l] = L[l].blur;
        }
    } else {
        for (int l = 0; l < nlevels; ++l) { free(L[l].img); free(L[l].blur); }
    }
    return total;
}

void svo_o_pyramid_free(svo_o_pyramid *p)
{
    for (int l = 0; l < p->nlevels; ++l) { free(p->img[l]); free(p->blur[l]); p->img[l] = p->blur[l] = 0; }
    p->nlevels = 0;
}

#include "svo_ham.h"

/* ------------------------------------------------------------------ */
/* Appendix C sparse stereo + SAD refinement (defined here)            */
/* ------------------------------------------------------------------ */
typedef struct { int dist; int idx; } dist_idx_t;
static int cmp_dist_idx(const void *a, const void *b)
{
    const dist_idx_t *x = (const dist_idx_t *)a, *y = (const dist_idx_t *)b;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int svo_o_stereo_sparse(const svo_o_keypoint *kl, const uint8_t *dl, int nl,
                        const svo_o_keypoint *kr, const uint8_t *dr, int nr,
                        const svo_o_pyramid *pl, const svo_o_pyramid *pr,
                        float bf, float b, float *u_right, float *depth, int32_t *match_r, int32_t *sad)
{
    const int TH_HIGH = 100, TH_LOW = 50;
    const int th_orb = (TH_HIGH + TH_LOW) / 2;
    const float minD = 0.f, maxD = bf / b;
    const int rows = pl->h[0];
    dist_idx_t *acc = (dist_idx_t *)malloc(sizeof(dist_idx_t) * (size_t)(nl + 1));
    int nacc = 0;
    for (int i = 0; i < nl; ++i) { u_right[i] = -1.f; depth[i] = -1.f; if (match_r) match_r[i] = -1; if (sad) sad[i] = -1; }
    /* per-row table of right keypoints (ORB-SLAM2's vRowIndices): iR is listed, in ascending iR, in every image
     * row of [floor(y - r), ceil(y + r)], r = 2 * scale[octave] */
    int *row_off = (int *)calloc((size_t)rows + 1, sizeof(int)), *cursor = (int *)calloc((size_t)rows + 1, sizeof(int));
    int *row_list = NULL;
    for (int pass = 0; pass < 2; ++pass) {
        for (int iR = 0; iR < nr; ++iR) {
            float r = 2.0f * pr->scale[kr[iR].octave];
            int maxr = (int)ceilf(kr[iR].y + r), minr = (int)floorf(kr[iR].y - r);
            if (minr < 0) minr = 0;
            if (maxr > rows - 1) maxr = rows - 1;
            for (int y = minr; y <= maxr; ++y) {
                if (pass == 0) cursor[y]++;
                else row_list[cursor[y]++] = iR;
            }
        }
        if (pass == 0) {
            for (int y = 0; y < rows; ++y) { row_off[y + 1] = row_off[y] + cursor[y]; cursor[y] = row_off[y]; }
            row_list = (int *)malloc(sizeof(int) * (size_t)(row_off[rows] + 1));
        }
    }
    free(cursor);
    for (int iL = 0; iL < nl; ++iL) {
        const svo_o_keypoint *kpL = &kl[iL];
        int levelL = kpL->octave;
        float vL = kpL->y, uL = kpL->x;
        int row = (int)vL;
        if (row < 0 || row >= rows) continue;
        float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH, bestIdxR = 0;
        /* candidates: right keypoints whose row band [floor(y-r), ceil(y+r)] holds `row`,
         * visited in ascending iR (the order the per-row table is filled in) */
        for (int c = row_off[row]; c < row_off[row + 1]; ++c) {
            const int iR = row_list[c];
            const svo_o_keypoint *kpR = &kr[iR];
            if (kpR->octave < levelL - 1 || kpR->octave > levelL + 1) continue;
            float uR = kpR->x;
            if (uR >= minU && uR <= maxU) {
                int d = ham256(dl + 32 * (size_t)iL, dr + 32 * (size_t)iR);
                if (d < bestDist) { bestDist = d; bestIdxR = iR; }
            }
        }
        if (bestDist >= th_orb) continue;
        float uR0 = kr[bestIdxR].x;
        float sf = 1.f / pl->scale[levelL];
        int suL = (int)roundf(kpL->x * sf), svL = (int)roundf(kpL->y * sf), suR0 = (int)roundf(uR0 * sf);
        const int w = 5, Lr = 5;
        int lwid = pl->w[levelL], lhei = pl->h[levelL], rwid = pr->w[levelL];
        if (svL - w < 0 || svL + w >= lhei || suL - w < 0 || suL + w >= lwid) continue;
        int iniu = suR0 + Lr - w, endu = suR0 + Lr + w + 1;
        if (iniu - Lr - w < 0 || endu >= rwid) continue;
        const uint8_t *IL = pl->img[levelL], *IR = pr->img[levelL];
        int cL = IL[(size_t)svL * lwid + suL];
        int dists[11];
        int bestSad = 1 << 30, bestInc = 0;
        for (int inc = -Lr; inc <= Lr; ++inc) {
            int cR = IR[(size_t)svL * rwid + suR0 + inc];
            int s = 0;
            for (int dy = -w; dy <= w; ++dy)
                for (int dx = -w; dx <= w; ++dx) {
                    int a = IL[(size_t)(svL + dy) * lwid + suL + dx] - cL;
                    int c = IR[(size_t)(svL + dy) * rwid + suR0 + inc + dx] - cR;
                    s += abs(a - c);
                }
            if (s < bestSad) { bestSad = s; bestInc = inc; }
            dists[Lr + inc] = s;
        }
        if (bestInc == -Lr || bestInc == Lr) continue;
        float d1 = (float)dists[Lr + bestInc - 1], d2 = (float)dists[Lr + bestInc], d3 = (float)dists[Lr + bestInc + 1];
        float num = d1 - d3;
        float den = d1 + d3;
        float two_d2 = 2.0f * d2;
        den = den - two_d2;
        den = 2.0f * den;
        float deltaR = num / den;
        if (!(deltaR >= -1.f && deltaR <= 1.f)) continue;
        float pos = (float)suR0 + (float)bestInc;
        pos = pos + deltaR;
        float bestuR = pl->scale[levelL] * pos;
        float disparity = uL - bestuR;
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) { disparity = 0.01f; bestuR = uL - 0.01f; }
            depth[iL] = bf / disparity;
            u_right[iL] = bestuR;
            if (match_r) match_r[iL] = bestIdxR;
            if (sad) sad[iL] = bestSad;
            acc[nacc].dist = bestSad; acc[nacc].idx = iL; ++nacc;
        }
    }
    if (nacc) {
        qsort(acc, (size_t)nacc, sizeof(dist_idx_t), cmp_dist_idx);
        float median = (float)acc[nacc / 2].dist;
        float th = 1.5f * 1.4f * median;
        for (int i = nacc - 1; i >= 0; --i) {
            if ((float)acc[i].dist < th) break;
            u_right[acc[i].idx] = -1.f;
            depth[acc[i].idx] = -1.f;
        }
    }
    free(acc); free(row_off); free(row_list);
    return nacc;
}

/* cv::cvtColor(BGR2GRAY) on 8-bit input, the first thing cv::ORB::detectAndCompute does with a colour image
 * (the reference passes whatever cv::imread(.., CV_LOAD_IMAGE_UNCHANGED) returned, main.cpp:160-161, to
 * frame::featuredetect, src/frame.cc:75-79).  OpenCV 4.x fixed point: 15-bit weights B 3735, G 19235, R 9798,
 * round to nearest.  Pinned against cv2.cvtColor on all 2^24 colours (tests/test_oracle_vs_cv2.py).
 * (OpenCV 3.2 used the 14-bit weights 1868/9617/4899, which differ on 0.3 % of colours.) */
void svo_o_bgr2gray(const uint8_t *bgr, int w, int h, int sstride, uint8_t *gray, int dstride)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t *p = bgr + (size_t)y * sstride + 3 * x;
            gray[(size_t)y * dstride + x] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
        }
}
