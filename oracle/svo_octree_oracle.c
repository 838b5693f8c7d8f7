/*
 * svo_octree_oracle.c — CPU definition of the opt-in quadtree ("octree") keypoint distribution.
 * TEST INFRASTRUCTURE ONLY (same rules as svo_oracle.c).
 *
 * north_star names "FAST keypoints with grid/octree distribution"; the reference has none
 * (src/frame.cc:75-79 runs cv::ORB, whose selection is KeyPointsFilter::retainBest — SURVEY.md
 * section 0 and 8f rank 3).  This mode follows the published algorithm of ORB-SLAM2's
 * ORBextractor::DistributeOctTree (the lineage north_star's vocabulary comes from; not in
 * /root/reference, not installed here): PARITY UNPINNED, this file is the definition.
 *
 *   - the level's keypoint rectangle is cut into nIni = round(width / height) root nodes;
 *   - rounds: every node holding more than one point is split into four (ceil-half boxes), empty
 *     children are dropped, until the node count reaches N or stops changing;
 *   - when one more full round could overshoot (nodes + 3 * expandable > N) nodes are split one
 *     at a time, most populated first, until the count reaches N;
 *   - each final node keeps its best-scoring point.
 * ORB-SLAM2 leaves two things to chance (its careful phase sorts (size, node pointer) pairs and its
 * output follows std::list push_front order); they are fixed here so the result is a function of
 * the input:
 *   - nodes are kept in Z order (root index, then child digit per level: 0 = top-left, 1 = top-right,
 *     2 = bottom-left, 3 = bottom-right); the careful phase takes candidates by size descending and
 *     Z order ascending; the output lists the final nodes in Z order;
 *   - score ties inside a node go to the first point in raster order (y, then x);
 *   - a node is never split below depth SVO_O_OCT_MAXD (12 levels: boxes of 4095 px reach 1 px).
 * This file keeps explicit node boxes and point lists; the device kernel (csrc/octree.cu) sorts the
 * points by their full-depth path code instead and works on runs of the sorted array.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "svo_oracle.h"

typedef struct {
    int ulx, urx, uly, bry; /* box [ulx, urx) x [uly, bry), rectangle-relative */
    int depth;
    int *pts;               /* indices into the caller's arrays */
    int n;
    int split;              /* marked for splitting in the current pass */
} onode;

static int split_node(const onode *p, const int32_t *xs, const int32_t *ys, int x0, int y0, onode *out)
{
    /* ExtractorNode::DivideNode: halfX = ceil((UR.x - UL.x) / 2), halfY = ceil((BR.y - UL.y) / 2) */
    const int sx = p->ulx + (p->urx - p->ulx + 1) / 2, sy = p->uly + (p->bry - p->uly + 1) / 2;
    onode c[4];
    for (int q = 0; q < 4; ++q) {
        c[q].ulx = (q & 1) ? sx : p->ulx; c[q].urx = (q & 1) ? p->urx : sx;
        c[q].uly = (q & 2) ? sy : p->uly; c[q].bry = (q & 2) ? p->bry : sy;
        c[q].depth = p->depth + 1; c[q].n = 0; c[q].split = 0;
        c[q].pts = (int *)malloc(sizeof(int) * (size_t)p->n);
    }
    for (int i = 0; i < p->n; ++i) {
        const int id = p->pts[i];
        const int q = ((xs[id] - x0) < sx ? 0 : 1) | ((ys[id] - y0) < sy ? 0 : 2);
        c[q].pts[c[q].n++] = id;
    }
    int m = 0;
    for (int q = 0; q < 4; ++q) {
        if (c[q].n) out[m++] = c[q];
        else free(c[q].pts);
    }
    return m;
}

typedef struct { int cnt, pos; } ocand;
static int cand_cmp(const void *a, const void *b)
{
    const ocand *x = (const ocand *)a, *y = (const ocand *)b;
    if (x->cnt != y->cnt) return x->cnt > y->cnt ? -1 : 1; /* most populated first */
    return x->pos < y->pos ? -1 : (x->pos > y->pos);      /* then Z order          */
}

/* replace every node marked `split` by its non-empty children, in place order (keeps Z order) */
static onode *apply_splits(onode *nodes, int *nn, const int32_t *xs, const int32_t *ys, int x0, int y0)
{
    onode *out = (onode *)malloc(sizeof(onode) * ((size_t)*nn * 4 + 4));
    int m = 0;
    for (int k = 0; k < *nn; ++k) {
        if (!nodes[k].split) { out[m++] = nodes[k]; continue; }
        m += split_node(&nodes[k], xs, ys, x0, y0, out + m);
        free(nodes[k].pts);
    }
    free(nodes);
    *nn = m;
    return out;
}

static int expandable(const onode *p) { return p->n > 1 && p->depth < SVO_O_OCT_MAXD; }

int svo_o_distribute_octree(const int32_t *xs, const int32_t *ys, const int32_t *score, int n,
                            int x0, int y0, int x1, int y1, int N, int32_t *out_idx)
{
    const int width = x1 - x0, height = y1 - y0;
    if (n <= 0 || width <= 0 || height <= 0) return 0;
    int nIni = (int)roundf((float)width / (float)height);
    if (nIni < 1) nIni = 1;
    if (nIni > 16) nIni = 16;
    const float hX = (float)width / (float)nIni;
    onode *nodes = (onode *)malloc(sizeof(onode) * (size_t)nIni);
    for (int i = 0; i < nIni; ++i) {
        nodes[i].ulx = (int)(hX * (float)i); nodes[i].urx = (int)(hX * (float)(i + 1));
        nodes[i].uly = 0; nodes[i].bry = height; nodes[i].depth = 0; nodes[i].n = 0; nodes[i].split = 0;
        nodes[i].pts = (int *)malloc(sizeof(int) * (size_t)n);
    }
    for (int i = 0; i < n; ++i) {   /* points arrive in raster order and keep it inside every node */
        int r = (int)((float)(xs[i] - x0) / hX);
        if (r > nIni - 1) r = nIni - 1;
        if (r < 0) r = 0;
        nodes[r].pts[nodes[r].n++] = i;
    }
    int nn = 0;
    for (int i = 0; i < nIni; ++i) {
        if (nodes[i].n) nodes[nn++] = nodes[i];
        else free(nodes[i].pts);
    }
    int finish = 0;
    while (!finish) {
        const int prev = nn;
        for (int k = 0; k < nn; ++k) nodes[k].split = expandable(&nodes[k]);
        nodes = apply_splits(nodes, &nn, xs, ys, x0, y0);
        int n_expand = 0;
        for (int k = 0; k < nn; ++k) n_expand += expandable(&nodes[k]);
        if (nn >= N || nn == prev) finish = 1;
        else if (nn + 3 * n_expand > N) {
            while (!finish) {
                const int prev2 = nn;
                ocand *c = (ocand *)malloc(sizeof(ocand) * (size_t)(nn + 1));
                int nc = 0;
                for (int k = 0; k < nn; ++k)
                    if (expandable(&nodes[k])) { c[nc].cnt = nodes[k].n; c[nc].pos = k; ++nc; }
                qsort(c, (size_t)nc, sizeof(ocand), cand_cmp);
                int total = nn;
                for (int j = 0; j < nc; ++j) {
                    onode tmp[4];
                    const int m = split_node(&nodes[c[j].pos], xs, ys, x0, y0, tmp);
                    for (int q = 0; q < m; ++q) free(tmp[q].pts);
                    nodes[c[j].pos].split = 1;
                    total += m - 1;
                    if (total >= N) break;
                }
                free(c);
                nodes = apply_splits(nodes, &nn, xs, ys, x0, y0);
                if (nn >= N || nn == prev2) finish = 1;
            }
        }
    }
    for (int k = 0; k < nn; ++k) {
        int best = nodes[k].pts[0];
        for (int i = 1; i < nodes[k].n; ++i) {
            const int id = nodes[k].pts[i];
            if (score[id] > score[best] ||
                (score[id] == score[best] && (ys[id] < ys[best] || (ys[id] == ys[best] && xs[id] < xs[best]))))
                best = id;
        }
        out_idx[k] = best;
        free(nodes[k].pts);
    }
    free(nodes);
    return nn;
}
