/*
 * svo_matchers.c — CPU ORACLE, matcher half (TEST INFRASTRUCTURE; see the header of svo_oracle.c for scope).
 * The integer matchers of the path: DescriptorDistance, BFMatcher + distance filter, the two greedy scans of
 * pnpmatch::poseEstimationPnP with the pass-1 "dynamic" veto, and frame::disp2Depth.  Pinned to the reference's own
 * compiled code by tests/test_ref_pin.py (oracle/_ref).  Its own translation unit so that bench.py's CPU arm can
 * also build it on the running host with -O3 -march=native (oracle.native_matchers()), as the reference's
 * CMakeLists.txt:10-11 would.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "svo_oracle.h"
#include "svo_ham.h"

/* ------------------------------------------------------------------ */
/* B.1 Hamming distance  (src/pnpmatch.cc:14-30)                       */
/* ------------------------------------------------------------------ */
int svo_o_hamming(const uint8_t *a, const uint8_t *b)
{
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4);
        memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

int svo_o_hamming_popcnt(const uint8_t *a, const uint8_t *b) { return ham256(a, b); }

/* B.2 BFMatcher(NORM_HAMMING).match: per query, first minimum over train
 * (src/pnpmatch.cc:266,278); then the keep filter of :281-299. */
void svo_o_match_bf(const uint8_t *q, int nq, const uint8_t *t, int nt,
                    int32_t *idx, int32_t *dist, uint8_t *keep)
{
    int min_dist = 10000;
    for (int i = 0; i < nq; ++i) {
        int best = 1 << 30, bi = -1;
        for (int j = 0; j < nt; ++j) {
            int d = ham256(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < best) { best = d; bi = j; }
        }
        idx[i] = bi;
        dist[i] = nt ? best : -1;
        if (nt && best < min_dist) min_dist = best;
    }
    if (keep) {
        double thr = 2.0 * (double)min_dist > 30.0 ? 2.0 * (double)min_dist : 30.0;
        for (int i = 0; i < nq; ++i) keep[i] = (nt && (double)dist[i] <= thr) ? 1 : 0;
    }
}

/* The "dynamic" test of pass 1 (src/pnpmatch.cc:103-122): the would-be match (last-frame keypoint `last`,
 * current keypoint `cur`) is dynamic when `cur` lies inside some offline YOLO box grown by 10 px AND its f64
 * distance to the line F * (last.x, last.y, 1) exceeds 0.1.  float/int compares and the double arithmetic are
 * written in the reference's operand order; every product and sum is rounded separately (no FMA). */
int svo_o_veto_dynamic(const int32_t *boxes, int n_boxes, const double *F, float lx, float ly, float cx, float cy)
{
    for (int k = 0; k < n_boxes; ++k) {
        int left = boxes[4 * k], right = boxes[4 * k + 1], top = boxes[4 * k + 2], bottom = boxes[4 * k + 3];
        if (cx > left - 10 && cx < right + 10 && cy > top - 10 && cy < bottom + 10) {
            double A = F[0] * lx + F[1] * ly + F[2];
            double B = F[3] * lx + F[4] * ly + F[5];
            double C = F[6] * lx + F[7] * ly + F[8];
            double dd = fabs(A * cx + B * cy + C) / sqrt(A * A + B * B);
            if (dd > 0.1) return 1;
        }
    }
    return 0;
}

/* B.3 greedy scan of poseEstimationPnP.
 *   mode 0 = pass 1 (src/pnpmatch.cc:61-156): claim iff best < 15 and the would-be match is not
 *            "dynamic" (:103-122, svo_o_veto_dynamic above); a dynamic match marks the row's map point
 *            bad instead (row_bad[i] = 1, :139-144) and claims nothing;
 *   mode 1 = pass 2 (:160-199): claim iff best < 30 && (float)second/(float)best > 2.
 * row_live[i] == 0 skips the row (no live map point / already observing).
 * claimed[j] != 0 marks a taken column (CurrentFrame->MapPoints[j] != NULL);
 * on a claim claimed[j] is set and claim_row[j] = i + row_base.
 * win_*: optional projection window (u, v, radius per row; cur keypoint x/y per
 * column); NULL => reference behaviour (brute force over all columns).
 * veto (pass 1 only; n_boxes == 0 or F == NULL => none): boxes n_boxes x 4 (left, right, top, bottom),
 * F 3x3 row-major f64, row_xy 2 x M (LastFrame.keypoints_l[i].pt), vcur_xy 2 x N (CurrentFrame->keypoints_l[j].pt).
 */
void svo_o_match_greedy_veto(const uint8_t *rows, int M, const uint8_t *cur, int N, int mode,
                             const uint8_t *row_live, uint8_t *claimed, int32_t *claim_row, int row_base,
                             int32_t *best_idx, int32_t *best, int32_t *second, uint8_t *row_claimed,
                             const float *win_uvr, const float *cur_xy,
                             const int32_t *boxes, int n_boxes, const double *F, const float *row_xy,
                             const float *vcur_xy, uint8_t *row_bad)
{
    const int use_veto = mode == 0 && n_boxes > 0 && boxes && F && row_xy && vcur_xy;
    for (int i = 0; i < M; ++i) {
        best_idx[i] = -1; best[i] = 256; second[i] = 256;
        if (row_claimed) row_claimed[i] = 0;
        if (row_bad) row_bad[i] = 0;
        if (row_live && !row_live[i]) continue;
        int bd = 256, sd = 256, bi = -1;
        for (int j = 0; j < N; ++j) {
            if (claimed[j]) continue;
            if (win_uvr) {
                float du = cur_xy[2 * j] - win_uvr[3 * i], dv = cur_xy[2 * j + 1] - win_uvr[3 * i + 1];
                float r = win_uvr[3 * i + 2];
                if (du < -r || du > r || dv < -r || dv > r) continue;
            }
            int d = ham256(rows + 32 * (size_t)i, cur + 32 * (size_t)j);
            if (d < bd) { sd = bd; bd = d; bi = j; }
        }
        best_idx[i] = bi; best[i] = bd; second[i] = sd;
        int take;
        if (mode == 0) take = bd < 15;
        else take = bd < 30 && (float)sd / (float)bd > 2;
        if (take && bi >= 0) {
            if (use_veto && svo_o_veto_dynamic(boxes, n_boxes, F, row_xy[2 * i], row_xy[2 * i + 1],
                                               vcur_xy[2 * bi], vcur_xy[2 * bi + 1])) {
                if (row_bad) row_bad[i] = 1;      /* mp->bad = true; continue (:139-144) */
                continue;
            }
            claimed[bi] = 1;
            if (claim_row) claim_row[bi] = i + row_base;
            if (row_claimed) row_claimed[i] = 1;
        }
    }
}

void svo_o_match_greedy(const uint8_t *rows, int M, const uint8_t *cur, int N, int mode,
                        const uint8_t *row_live, uint8_t *claimed, int32_t *claim_row, int row_base,
                        int32_t *best_idx, int32_t *best, int32_t *second, uint8_t *row_claimed,
                        const float *win_uvr, const float *cur_xy)
{
    svo_o_match_greedy_veto(rows, M, cur, N, mode, row_live, claimed, claim_row, row_base, best_idx, best, second,
                            row_claimed, win_uvr, cur_xy, NULL, 0, NULL, NULL, NULL, NULL);
}

/* frame::disp2Depth (src/frame.cc:140-164) */
void svo_o_disp2depth(const float *disp, float *depth, size_t n, float bf)
{
    for (size_t i = 0; i < n; ++i) depth[i] = disp[i] != 0.f ? bf / disp[i] : -1.f;
}


/* Projection windows for the opt-in "projection-guided" pass 2 (north_star).  The
 * reference computes the motion model (Velocity = Tcw * LastTwc, src/Tracking.cc:99-106) but never applies it
 * (src/pnpmatch.cc:53 is commented out, Rcw/tcw are read and unused, :56-57), so this stage follows ORB-SLAM2's
 * SearchByProjection, the code the reference descends from: a map point X is transformed with the predicted pose
 * Tcw (row-major 4x4: Velocity * LastFrame.Tcw), projected with the pinhole intrinsics, and gets a square search
 * window of half-size th * scale[octave] around (u, v).  Points behind the camera or outside the image get r = -1
 * (a window nothing falls into).  float32 throughout, every operation rounded separately, in exactly this order.
 * DEFINED HERE (parity unpinned: there is no reference code to compare with). */
void svo_o_project_map(const float *xyz, const int32_t *octave, int n, const float *Tcw, float fx, float fy, float cx, float cy,
                       int W, int H, float th, const float *lscale, int nlevels, float *uvr)
{
    for (int i = 0; i < n; ++i) {
        const float X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        float xc = Tcw[0] * X; xc = xc + Tcw[1] * Y; xc = xc + Tcw[2] * Z; xc = xc + Tcw[3];
        float yc = Tcw[4] * X; yc = yc + Tcw[5] * Y; yc = yc + Tcw[6] * Z; yc = yc + Tcw[7];
        float zc = Tcw[8] * X; zc = zc + Tcw[9] * Y; zc = zc + Tcw[10] * Z; zc = zc + Tcw[11];
        float u = 0.f, v = 0.f, r = -1.f;
        if (zc > 0.f) {
            const float invz = 1.0f / zc;
            float pu = fx * xc; pu = pu * invz; pu = pu + cx;
            float pv = fy * yc; pv = pv * invz; pv = pv + cy;
            if (pu >= 0.f && pu < (float)W && pv >= 0.f && pv < (float)H) {
                int o = octave ? octave[i] : 0;
                if (o < 0) o = 0;
                if (o > nlevels - 1) o = nlevels - 1;
                u = pu; v = pv; r = th * lscale[o];
            }
        }
        uvr[3 * i] = u; uvr[3 * i + 1] = v; uvr[3 * i + 2] = r;
    }
}
