// stereo.cu — sparse left/right matching with SAD sub-pixel refinement (north_star's
// ComputeStereoMatches stage).  The reference itself runs a dense
// MSA solve here (src/Tracking.cc:226-228) and reads the disparity back at keypoint pixels
// (frame::computekeypoint_r / disp2Depth, src/frame.cc:122-164); this stage delivers the
// same per-keypoint fields (u_right, depth) from the sparse formulation north_star names.
// Its semantics are DEFINED by oracle/svo_oracle.c:svo_o_stereo_sparse (parity unpinned
// against the reference: no such code or vectors exist there).
//
// k_stereo_rows first bins the right keypoints by image row (every right keypoint is a candidate for
// the rows [floor(y - r), ceil(y + r)], r = 2 * scale[octave]; CSR lists built with shared-memory
// counters, one CTA per frame).  Then one warp per left keypoint: lanes stride the candidates of the
// keypoint's row, test the octave / disparity-range predicate, and reduce (dist << 20 | iR) minima so
// the first minimum wins (the key makes the result independent of the list order); the same warp then evaluates the 11 SAD windows (121 px each, lanes
// over pixels, integer sums via redux.sync) and lane 0 fits the parabola.
// A second kernel applies the 1.5*1.4*median SAD cut per frame.
#include "svo_internal.cuh"

#define ST_WARPS 8

__global__ void __launch_bounds__(1024) k_stereo_rows(Bufs b, Geom g, int slot0, StereoArgs a)
{
    extern __shared__ int row_cnt[];       // [H + 1] counts, then exclusive offsets, then fill cursors
    __shared__ int wsum[32];
    const int f = blockIdx.x, sr = slot0 + 2 * f + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nr = min(b.nkp[sr], g.kp_cap), H = g.H;
    const svo_keypoint *kr = b.kp + (size_t)sr * g.kp_cap;
    int *off = a.row_off + (size_t)f * (H + 1);
    uint16_t *list = a.row_list + (size_t)f * a.row_list_stride;
    for (int i = tid; i <= H; i += 1024) row_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < nr; i += 1024) {
        const svo_keypoint q = kr[i];
        const float r = __fmul_rn(2.0f, g.lv[q.octave].scale);
        const int maxr = min((int)ceilf(__fadd_rn(q.y, r)), H - 1), minr = max((int)floorf(__fsub_rn(q.y, r)), 0);
        for (int y = minr; y <= maxr; ++y) atomicAdd(&row_cnt[y], 1);
    }
    __syncthreads();
    // exclusive scan over H + 1 counters (each thread owns a contiguous chunk)
    const int per = (H + 1 + 1023) / 1024;
    const int i0 = min(tid * per, H + 1), i1 = min(i0 + per, H + 1);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += row_cnt[i];
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int base = inc - sum;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    for (int i = i0; i < i1; ++i) { const int c = row_cnt[i]; row_cnt[i] = base; off[i] = base; base += c; }
    __syncthreads();
    for (int i = tid; i < nr; i += 1024) {
        const svo_keypoint q = kr[i];
        const float r = __fmul_rn(2.0f, g.lv[q.octave].scale);
        const int maxr = min((int)ceilf(__fadd_rn(q.y, r)), H - 1), minr = max((int)floorf(__fsub_rn(q.y, r)), 0);
        for (int y = minr; y <= maxr; ++y) {
            const int pos = atomicAdd(&row_cnt[y], 1);
            if (pos < a.row_list_stride) list[pos] = (uint16_t)i;
        }
    }
}

__global__ void __launch_bounds__(ST_WARPS * 32) k_stereo(Bufs b, Geom g, int slot0, StereoArgs a)
{
    const int f = blockIdx.y;
    const int sl = slot0 + 2 * f, sr = sl + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nl = min(b.nkp[sl], g.kp_cap);
    const svo_keypoint *kl = b.kp + (size_t)sl * g.kp_cap, *kr = b.kp + (size_t)sr * g.kp_cap;
    const uint4 *dl = reinterpret_cast<const uint4 *>(b.desc + (size_t)sl * g.kp_cap * 32);
    const uint4 *dr = reinterpret_cast<const uint4 *>(b.desc + (size_t)sr * g.kp_cap * 32);
    const float bf = a.bf[f], base = a.baseline[f];
    const float minD = 0.f, maxD = __fdiv_rn(bf, base);
    const int rows = g.H;
    const int *row_off = a.row_off + (size_t)f * (rows + 1);
    const uint16_t *row_list = a.row_list + (size_t)f * a.row_list_stride;
    for (int iL = blockIdx.x * ST_WARPS + warp; iL < nl; iL += gridDim.x * ST_WARPS) {
        const size_t o = (size_t)f * a.stride + iL;
        const svo_keypoint kp = kl[iL];
        float uR_out = -1.f, depth_out = -1.f;
        int match_out = -1, sad_out = -1;
        const int levelL = kp.octave;
        const float vL = kp.y, uL = kp.x;
        const int row = (int)vL;
        const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
        bool go = row >= 0 && row < rows && !(maxU < 0);
        uint32_t key = (100u << 20);  // TH_HIGH, bestIdxR = 0
        if (go) {
            const uint4 a0 = dl[2 * iL], a1 = dl[2 * iL + 1];
            const int c0 = row_off[row], c1 = min(row_off[row + 1], a.row_list_stride);
            for (int c = c0 + lane; c < c1; c += 32) {
                const int iR = row_list[c];               // right keypoints whose row band holds `row`
                const float qx = kr[iR].x;
                const int qo = kr[iR].octave;
                if (qo < levelL - 1 || qo > levelL + 1) continue;
                if (qx >= minU && qx <= maxU) {
                    const uint4 x = dr[2 * iR], y = dr[2 * iR + 1];
                    const int d = __popc(a0.x ^ x.x) + __popc(a0.y ^ x.y) + __popc(a0.z ^ x.z) + __popc(a0.w ^ x.w) +
                                  __popc(a1.x ^ y.x) + __popc(a1.y ^ y.y) + __popc(a1.z ^ y.z) + __popc(a1.w ^ y.w);
                    if (d < 100) key = min(key, ((uint32_t)d << 20) | (uint32_t)iR);
                }
            }
        }
        key = __reduce_min_sync(0xffffffffu, key);
        const int bestDist = (int)(key >> 20), bestIdxR = (int)(key & 0xfffffu);
        go = go && bestDist < 75;  // (TH_HIGH + TH_LOW) / 2
        if (go) {
            const LevelGeom &L = g.lv[levelL];
            const float uR0 = kr[bestIdxR].x;
            const float sf = L.inv_scale;
            const int suL = (int)roundf(__fmul_rn(kp.x, sf)), svL = (int)roundf(__fmul_rn(kp.y, sf));
            const int suR0 = (int)roundf(__fmul_rn(uR0, sf));
            const int w = 5, Lr = 5;
            const bool fits = !(svL - w < 0 || svL + w >= L.h || suL - w < 0 || suL + w >= L.w) &&
                              !(suR0 - Lr - w < 0 || suR0 + Lr + w + 1 >= L.w);
            if (fits) {
                const uint8_t *IL = b.pyr + (size_t)sl * g.pyr_bytes + L.off;
                const uint8_t *IR = b.pyr + (size_t)sr * g.pyr_bytes + L.off;
                const int sp = L.pitch;
                const int cL = IL[(size_t)svL * sp + suL];
                int cR[11], s[11];
#pragma unroll
                for (int k = 0; k < 11; ++k) { cR[k] = IR[(size_t)svL * sp + suR0 + k - Lr]; s[k] = 0; }
                for (int p = lane; p < 121; p += 32) {
                    const int dy = p / 11 - w, dx = p % 11 - w;
                    const int av = IL[(size_t)(svL + dy) * sp + suL + dx] - cL;
                    const uint8_t *rp = IR + (size_t)(svL + dy) * sp + suR0 + dx - Lr;
#pragma unroll
                    for (int k = 0; k < 11; ++k) s[k] += abs(av - ((int)rp[k] - cR[k]));
                }
                int bestSad = 1 << 30, bestInc = 0;
#pragma unroll
                for (int k = 0; k < 11; ++k) {
                    s[k] = __reduce_add_sync(0xffffffffu, s[k]);
                    if (s[k] < bestSad) { bestSad = s[k]; bestInc = k - Lr; }
                }
                if (bestInc != -Lr && bestInc != Lr) {
                    float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int k = 1; k < 10; ++k)
                        if (k - Lr == bestInc) { d1 = (float)s[k - 1]; d2 = (float)s[k]; d3 = (float)s[k + 1]; }
                    const float num = __fsub_rn(d1, d3);
                    float den = __fadd_rn(d1, d3);
                    den = __fsub_rn(den, __fmul_rn(2.0f, d2));
                    den = __fmul_rn(2.0f, den);
                    const float deltaR = __fdiv_rn(num, den);
                    if (deltaR >= -1.f && deltaR <= 1.f) {
                        float pos = __fadd_rn((float)suR0, (float)bestInc);
                        pos = __fadd_rn(pos, deltaR);
                        float bestuR = __fmul_rn(L.scale, pos);
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= minD && disparity < maxD) {
                            if (disparity <= 0) { disparity = 0.01f; bestuR = __fsub_rn(uL, 0.01f); }
                            depth_out = __fdiv_rn(bf, disparity);
                            uR_out = bestuR;
                            match_out = bestIdxR;
                            sad_out = bestSad;
                        }
                    }
                }
            }
        }
        if (lane == 0) {
            a.u_right[o] = uR_out; a.depth[o] = depth_out;
            a.match_r[o] = match_out; a.sad[o] = sad_out;
        }
    }
}

// median SAD cut: th = 1.5f*1.4f*sorted_sad[n/2]; invalidate sad >= th.  One CTA per frame;
// the k-th smallest SAD (16-bit range) is found with a two-level 256-bin histogram.
__global__ void __launch_bounds__(1024) k_stereo_median(Bufs b, Geom g, int slot0, StereoArgs a)
{
    __shared__ int hist[256];
    __shared__ int s_n, s_bin, s_rank, s_med;
    const int f = blockIdx.x, sl = slot0 + 2 * f;
    const int nl = min(b.nkp[sl], g.kp_cap);
    const int tid = threadIdx.x;
    const size_t o = (size_t)f * a.stride;
    if (tid < 256) hist[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();
    int c = 0;
    for (int i = tid; i < nl; i += 1024) {
        const int s = a.sad[o + i];
        if (s >= 0) { atomicAdd(&hist[(s >> 8) & 255], 1); ++c; }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((tid & 31) == 0 && c) atomicAdd(&s_n, c);
    __syncthreads();
    const int n = s_n;
    if (tid == 0) a.n_stereo[f] = n;
    if (n == 0) return;
    if (tid == 0) {
        int k = n / 2, bin = 0;
        while (k >= hist[bin]) { k -= hist[bin]; ++bin; }
        s_bin = bin; s_rank = k;
    }
    __syncthreads();
    const int bin = s_bin;
    __syncthreads();
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nl; i += 1024) {
        const int s = a.sad[o + i];
        if (s >= 0 && ((s >> 8) & 255) == bin) atomicAdd(&hist[s & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int k = s_rank, lo = 0;
        while (k >= hist[lo]) { k -= hist[lo]; ++lo; }
        s_med = (bin << 8) | lo;
    }
    __syncthreads();
    const float th = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_med);
    int removed = 0;
    for (int i = tid; i < nl; i += 1024) {
        const int s = a.sad[o + i];
        if (s >= 0 && !((float)s < th)) { a.u_right[o + i] = -1.f; a.depth[o + i] = -1.f; ++removed; }
    }
    removed = __reduce_add_sync(0xffffffffu, removed);
    if ((tid & 31) == 0 && removed) atomicSub(a.n_stereo + f, removed);
}

void launch_stereo(const Bufs &b, const Geom &g, int slot0, int nframes, const StereoArgs &a, cudaStream_t st,
                   long long *launches)
{
    int quota = 0;
    for (int l = 0; l < g.nlevels; ++l) quota += g.lv[l].quota;
    dim3 grid((quota + 64 + ST_WARPS - 1) / ST_WARPS, nframes);
    k_stereo_rows<<<nframes, 1024, (size_t)(g.H + 1) * sizeof(int), st>>>(b, g, slot0, a);
    k_stereo<<<grid, ST_WARPS * 32, 0, st>>>(b, g, slot0, a);
    k_stereo_median<<<nframes, 1024, 0, st>>>(b, g, slot0, a);
    *launches += 3;
}
