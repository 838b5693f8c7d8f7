// track.cu — device-resident tracker state (opt-in): what Tracking::Track carries from one frame to the next stays in
// HBM instead of travelling to the host and back every frame.
//
// In the reference the previous frame's map points and the local map are products of earlier frames
// (src/Tracking.cc:237-250): `lastframe = frame(currentframe); lastframe.createmappoint(LocalMapPoints);` then every
// point with `create_id <= frame_num - 4` is erased.  A frame's inputs to pnpmatch::poseEstimationPnP are therefore
//   pass-1 rows   LastFrame.MapPoints[i]->m_descriptor for every keypoint i that owns a live map point — the descriptor
//                 FROZEN at the point's creation (src/mappoint.cc:12), not the last frame's own descriptor
//   pass-2 rows   the m_descriptor of every point of LocalMapPoints (last 4 frames), skipping points already matched in
//                 pass 1 (observations.count, src/pnpmatch.cc:165) or bad (:163)
// and after tracking, CurrentFrame->MapPoints[j] is the point that claimed column j (pass 1: src/pnpmatch.cc:151,
// pass 2: :195) or, from createmappoint, a NEW point when the keypoint has depth > 0 and lies outside every offline
// box grown by 5 px (src/frame.cc:182-238).
//
// Per sequence the state is (ping-pong: the update reads one copy and writes the other):
//   last_desc[K][32], prev_xy[K][2]                              last frame's own descriptors (train set of the BF matcher,
//                                                                src/pnpmatch.cc:253-300) and keypoint positions (the veto's `last`)
//   prev_desc[K][32], prev_live[K], prev_map_row[K], n_prev      last frame's keypoints: frozen descriptor of the owned
//                                                                point, 1 if it owns a live point, that point's map row
//   prev_create[K], prev_xyz[K][3]                               the owned point's name for the host: id of its creating frame and
//                                                                its position in that frame's camera coordinates (UnprojectStereo,
//                                                                src/frame.cc:166-180, before Rwc / twc: poses stay on the host)
//   map_desc[C][32], map_create[C], map_link[C], map_xyz[C][3], n_map   local map in scan order: frozen descriptor, id of the
//                                                                creating frame, owning keypoint of the last frame (-1), position
// The pass-2 scan order is "survivors in their previous order, then the new points in keypoint order" — the reference
// iterates a std::set<mappoint*> in pointer order, which is arbitrary, so any fixed order is as faithful as another.
#include "svo_internal.cuh"
#include <limits.h>

#define TRK_THREADS 1024

// counts of the tracked frames of a batch: state -> the batch's parameter block (n_prev, n_map)
__global__ void k_track_load(const FramePtrs *__restrict__ fp, int *n_prev, int *n_map, int n)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n || !fp[f].trk_in) return;
    const TrackState &s = *fp[f].trk_in;
    n_prev[f] = *s.n_prev; n_map[f] = *s.n_map;
}

// exclusive prefix sum of one value per thread over the block; *total receives the block sum (valid after the call)
__device__ __forceinline__ int block_scan_excl(int v, int *wsum, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();                       // wsum may still be read from the previous call
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = wsum[lane];
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        wsum[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return wsum[warp] + inc - v;
}

// One CTA per tracked frame, after both passes and the stereo stage of the batch.
__global__ void __launch_bounds__(TRK_THREADS) k_track_update(TrackUpdateArgs a)
{
    const int f = blockIdx.x;
    const FramePtrs &P = a.fp[f];
    if (!P.trk_in) return;                                             // whole CTA
    const TrackState si = *P.trk_in, so = *P.trk_out;
    __shared__ int wsum[32];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int N = min(a.nkp[(size_t)f * 2], a.kp_cap);                 // current (left) keypoints
    const int np = min(*si.n_prev, a.kp_cap), nm = min(*si.n_map, a.map_cap);
    const int *claim_row = a.claim_row + (size_t)f * a.col_stride;
    const float *depth = a.depth + (size_t)f * a.col_stride;
    const uint8_t *row_bad = (a.p1_row_bad && P.n_boxes > 0) ? a.p1_row_bad + (size_t)f * a.row_stride : nullptr;
    const svo_keypoint *kp = a.kp + (size_t)f * 2 * a.kp_cap;
    const uint8_t *cur_desc = a.desc + (size_t)f * 2 * a.kp_cap * 32;
    int *remap = a.scratch + (size_t)f * a.map_cap;                    // old map row -> new row, -1 = gone
    const int frame_id = P.frame_id, window = a.window;
    // ---- which old rows survive: not aged out (src/Tracking.cc:239-250: erased when create_id <= frame_num - 4) and not
    // marked bad by this frame's pass-1 veto (src/pnpmatch.cc:141: a bad point is never used again)
    for (int r = tid; r < nm; r += TRK_THREADS) remap[r] = si.map_create[r] > frame_id - window ? 0 : -1;
    __syncthreads();
    if (row_bad)
        for (int i = tid; i < np; i += TRK_THREADS)
            if (row_bad[i]) { const int r = si.prev_map_row[i]; if (r >= 0 && r < nm) remap[r] = -1; }
    __syncthreads();
    int base = 0;
    for (int r0 = 0; r0 < nm; r0 += TRK_THREADS) {
        const int r = r0 + tid;
        const int keep = (r < nm && remap[r] == 0) ? 1 : 0;
        const int pos = block_scan_excl(keep, wsum, &s_total);
        if (keep) {
            const int nr = base + pos;                                 // survivors never exceed the capacity they came from
            remap[r] = nr;
            const uint4 *src = reinterpret_cast<const uint4 *>(si.map_desc) + (size_t)r * 2;
            uint4 *dst = reinterpret_cast<uint4 *>(so.map_desc) + (size_t)nr * 2;
            dst[0] = src[0]; dst[1] = src[1];
            so.map_create[nr] = si.map_create[r]; so.map_link[nr] = -1;
            so.map_xyz[3 * nr] = si.map_xyz[3 * r]; so.map_xyz[3 * nr + 1] = si.map_xyz[3 * r + 1]; so.map_xyz[3 * nr + 2] = si.map_xyz[3 * r + 2];
        }
        base += s_total;
        __syncthreads();
    }
    const int nsurv = base;
    __syncthreads();
    // ---- columns: the point each current keypoint owns from now on
    int nnew = 0;
    for (int j0 = 0; j0 < N; j0 += TRK_THREADS) {
        const int j = j0 + tid;
        const int cr = j < N ? claim_row[j] : -1;
        const bool by_p1 = cr >= 0 && cr < np, by_p2 = cr >= np && cr - np < nm;
        int create = 0;
        if (j < N && !by_p1 && !by_p2) {
            bool make = depth[j] > 0.f;                                // createmappoint: z > 0 ...
            const float u = kp[j].x, v = kp[j].y;
            for (int k = 0; make && k < P.n_boxes; ++k) {              // ... and outside every box grown by 5 px (src/frame.cc:196-207)
                const int *bx = P.boxes + 4 * k;
                if (u > bx[0] - 5 && u < bx[1] + 5 && v > bx[2] - 5 && v < bx[3] + 5) make = false;
            }
            create = make ? 1 : 0;
        }
        const int pos = block_scan_excl(create, wsum, &s_total);
        if (j < N) {
            int owner = -1;
            if (by_p1) { const int r = si.prev_map_row[cr]; owner = (r >= 0 && r < nm) ? remap[r] : -1; }
            else if (by_p2) owner = remap[cr - np];
            else if (create) { owner = nsurv + nnew + pos; if (owner >= a.map_cap) owner = -1; }   // map full: the point lives in the frame only
            const bool owns = by_p1 || by_p2 || create;
            so.prev_live[j] = owns ? 1 : 0;                            // owns a live point (one aged out of the map still lives in the frame)
            so.prev_map_row[j] = owner;
            const uint4 *cd = reinterpret_cast<const uint4 *>(cur_desc) + (size_t)j * 2;
            const uint4 c0 = cd[0], c1 = cd[1];
            uint4 *ld = reinterpret_cast<uint4 *>(so.last_desc) + (size_t)j * 2;
            ld[0] = c0; ld[1] = c1;
            so.prev_xy[2 * j] = kp[j].x; so.prev_xy[2 * j + 1] = kp[j].y;
            uint4 d0 = c0, d1 = c1;                                    // frozen descriptor of the owned point (src/mappoint.cc:12)
            int cid = -1; float px = 0.f, py = 0.f, pz = 0.f;
            if (by_p1) {
                const uint4 *src = reinterpret_cast<const uint4 *>(si.prev_desc) + (size_t)cr * 2;
                d0 = src[0]; d1 = src[1];
                cid = si.prev_create[cr]; px = si.prev_xyz[3 * cr]; py = si.prev_xyz[3 * cr + 1]; pz = si.prev_xyz[3 * cr + 2];
            } else if (by_p2) {
                const int r = cr - np;
                const uint4 *src = reinterpret_cast<const uint4 *>(si.map_desc) + (size_t)r * 2;
                d0 = src[0]; d1 = src[1];
                cid = si.map_create[r]; px = si.map_xyz[3 * r]; py = si.map_xyz[3 * r + 1]; pz = si.map_xyz[3 * r + 2];
            } else if (create) {
                // UnprojectStereo's camera-frame point (src/frame.cc:171-173): x = (u - cx) * z * (1 / fx)
                const float z = depth[j];
                cid = frame_id; pz = z;
                if (P.fx != 0.f && P.fy != 0.f) {
                    px = __fmul_rn(__fmul_rn(__fsub_rn(kp[j].x, P.cx), z), __fdiv_rn(1.f, P.fx));
                    py = __fmul_rn(__fmul_rn(__fsub_rn(kp[j].y, P.cy), z), __fdiv_rn(1.f, P.fy));
                }
            }
            uint4 *dst = reinterpret_cast<uint4 *>(so.prev_desc) + (size_t)j * 2;
            dst[0] = d0; dst[1] = d1;
            so.prev_create[j] = cid;
            so.prev_xyz[3 * j] = px; so.prev_xyz[3 * j + 1] = py; so.prev_xyz[3 * j + 2] = pz;
            if (a.mp_create) {
                a.mp_create[(size_t)f * a.col_stride + j] = cid;
                float *o = a.mp_xyz + ((size_t)f * a.col_stride + j) * 3;
                o[0] = px; o[1] = py; o[2] = pz;
            }
            if (owner >= 0) {
                so.map_link[owner] = j;                                // next frame: pass-1 row j and this map row are the same point
                if (create) {
                    uint4 *md = reinterpret_cast<uint4 *>(so.map_desc) + (size_t)owner * 2;
                    md[0] = d0; md[1] = d1;
                    so.map_create[owner] = frame_id;
                    so.map_xyz[3 * owner] = px; so.map_xyz[3 * owner + 1] = py; so.map_xyz[3 * owner + 2] = pz;
                }
            }
        }
        nnew += s_total;
        __syncthreads();
    }
    if (tid == 0) { *so.n_prev = N; *so.n_map = min(nsurv + nnew, a.map_cap); }
}

void launch_track_load(const FramePtrs *fp, int *n_prev, int *n_map, int n, cudaStream_t st, long long *launches)
{
    k_track_load<<<(n + 127) / 128, 128, 0, st>>>(fp, n_prev, n_map, n);
    ++*launches;
}

void launch_track_update(const TrackUpdateArgs &a, int n, cudaStream_t st, long long *launches)
{
    k_track_update<<<n, TRK_THREADS, 0, st>>>(a);
    ++*launches;
}
