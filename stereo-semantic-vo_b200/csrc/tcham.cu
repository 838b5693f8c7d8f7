// tcham.cu — the all-pairs 256-bit Hamming distances of the batch matchers on the 5th-generation tensor cores.
//
// A 256-bit descriptor becomes 256 int8 values, +1 for a clear bit and -1 for a set bit; the dot product of two such
// rows is (#equal bits) - (#different bits) = 256 - 2 d, so one tcgen05.mma.kind::i8 tile (128 x 128 x K = 256, s32
// accumulators in tensor memory) holds 16384 exact distances that the SIMT kernels of match.cu spend ~40 instructions
// each on (8 XOR, carry-save adders, 5-6 POPC on the XU pipe, adds).  What is left for the CUDA cores is the expansion
// of the operands into shared memory (done on the fly: the packed descriptors are all that ever lives in HBM) and an
// epilogue of about one instruction per distance.  Integer arithmetic throughout: results are bit-exact.
//
// One CTA owns 128 "A" rows (tile rows = TMEM lanes, one epilogue thread per row) of one frame and streams the frame's
// "B" rows through two shared-memory stages and two accumulator buffers:
//     warps 4-7  expand the A tile once, then every B tile (128 descriptors -> 128 x 256 int8, K-major, 128-byte swizzle:
//                the canonical UMMA layout), fence.proxy.async, arrive on full[stage]
//     warp 8     one thread issues 8 x tcgen05.mma (K = 32 bytes each) per tile and commits to empty[stage] / tfull[buf]
//     warps 0-3  tcgen05.ld their 32 lanes x 32 columns at a time and run the mode's epilogue; arrive on tempty[buf]
// A thread meets its row's columns in ascending order, which is exactly the order of the reference's scans
// (src/pnpmatch.cc:79-95, :177-190), so "first minimum" and "second = best before the last update" need no
// cross-lane composition.
//
// Modes (what the epilogue does with dot = 256 - 2 d):
//   TC_PAIRS   A = current frame (BFMatcher queries), B = previous frame.  Per query the first minimum over the train
//              rows (cv::BFMatcher, src/pnpmatch.cc:266,278) and, for pass 1, every (row, column) with d < 15 appended
//              to the row's short list (src/pnpmatch.cc:101).  Replaces k_pairs and the u8 distance matrix.
//   TC_SCORES  A = previous frame rows, B = current columns + their pass-1 claim times: the exact
//              (bestIdx2, bestDist, secondBestDist) of every live row as the sequential scan saw them
//              (match_score, src/pnpmatch.cc:99).  Replaces k_scores_m.
//   TC_SHORT   A = local-map rows, B = the columns pass 1 left free (k_free_cols): the ascending list of columns with
//              d < 60 per live row (pass 2, src/pnpmatch.cc:160-199).  Replaces k_shortlist / k_reuse in the batch path.
//   TC_DUMP    every dot product to global memory (bring-up / test tap).
#include "svo_internal.cuh"
#include <limits.h>

#define TC_M 128
#define TC_N 128
#define TC_THREADS 288
#define TC_OPERAND_BYTES (128 * 256)     // one expanded operand tile: two K atoms of 128 rows x 128 bytes
#define TC_ATOM_BYTES (128 * 128)
#define TC_TMEM_COLS 256                 // two accumulator buffers of TC_N columns
#define TC_SMEM_BYTES (3 * TC_OPERAND_BYTES + 1024 + 2048)   // A + two B stages, alignment slack, side arrays + barriers

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused by swizzled K-major layouts; 1 like CUTLASS) in [16,30), stride byte offset
// (8 rows x 128 B = 1024) >> 4 in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2) at [4,6), A and B = signed 8-bit (1) at [7,10) and
// [10,13), both K-major (0) at 15 / 16, N >> 3 at [17,23), M >> 4 at [24,29); dense, no saturation, no negation.
#define TC_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24))

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes (this warp's quarter) x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 4 descriptor bits -> 4 int8: +1 (0x01) for a clear bit, -1 (0xFF) for a set bit.  The multiply copies bit k to
// bit 8k (the four shifted copies of the nibble occupy disjoint bit ranges, so nothing carries).
__device__ __forceinline__ uint32_t expand4(uint32_t nib)
{
    const uint32_t x = (nib * 0x00204081u) & 0x01010101u;
    return (x * 0xFFu) | 0x01010101u;
}
// one descriptor (8 words) -> its 256-byte row of an operand tile.  Row r of a tile lies in 8-row groups of 1024 bytes;
// inside a group the 16-byte chunk index is XORed with the row (Swizzle<3,4,3>): the canonical K-major SWIZZLE_128B
// layout tcgen05.mma reads.  Chunks 0-7 (bits 0-127) go to K atom 0, chunks 8-15 to K atom 1.
__device__ __forceinline__ void expand_row(uint8_t *tile, int r, const uint4 &lo, const uint4 &hi, bool valid)
{
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint8_t *base = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const uint32_t h = (w[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
        uint4 o;
        if (valid) {
            o.x = expand4(h & 15u); o.y = expand4((h >> 4) & 15u); o.z = expand4((h >> 8) & 15u); o.w = expand4(h >> 12);
        } else o = make_uint4(0, 0, 0, 0);      // rows past the end of a set: zero dot products (masked by the epilogues)
        *reinterpret_cast<uint4 *>(base + (c >> 3) * TC_ATOM_BYTES + (((c & 7) ^ (r & 7)) << 4)) = o;
    }
}

__device__ __forceinline__ int tc_count(const MatchSet &s, int f) { return s.count ? min(s.count[(size_t)f * s.count_stride], s.stride_rows) : s.fixed_count; }
__device__ __forceinline__ const uint8_t *tc_desc(const MatchSet &s, int f)
{
    if (s.tab) return *reinterpret_cast<const uint8_t *const *>(reinterpret_cast<const char *>(s.tab) + (size_t)f * sizeof(FramePtrs));
    return s.desc + (size_t)f * s.desc_stride * 32;
}
__device__ __forceinline__ int popc256(const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d)
{
    return __popc(a.x ^ c.x) + __popc(a.y ^ c.y) + __popc(a.z ^ c.z) + __popc(a.w ^ c.w) + __popc(b.x ^ d.x) + __popc(b.y ^ d.y) +
           __popc(b.z ^ d.z) + __popc(b.w ^ d.w);
}
__device__ __forceinline__ uint32_t *tc_short_slot(const GreedyArgs &a, size_t row, int pos)
{
    return pos < 32 ? a.shortlist + row * 32 + pos : a.shortlist_hi + row * (SVO_SHORT_CAP - 32) + (pos - 32);
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_hamming(TcArgs p)
{
    extern __shared__ uint8_t tc_smem_raw[];
    const int f = blockIdx.y;
    const int a0 = blockIdx.x * TC_M;
    const int nA = tc_count(p.A, f);
    int nB = tc_count(p.B, f);
    if (MODE == TC_SHORT && p.b_index) nB = min(p.b_index_cnt[f], nB);
    if (a0 >= nA || nB <= 0) return;                                   // whole CTA, before anything is allocated
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operand tiles need 1024-byte alignment (the swizzle pattern repeats every 8 rows x 128 bytes)
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *tileA = smem, *tileB = smem + TC_OPERAND_BYTES;           // tileB: two stages
    int *side = reinterpret_cast<int *>(smem + 3 * TC_OPERAND_BYTES);  // [2][TC_N] per-column side data (claim time / original column)
    uint64_t *bars = reinterpret_cast<uint64_t *>(side + 2 * TC_N);    // full[2], empty[2], tfull[2], tempty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    const uint32_t bar0 = smem_u32(bars);
#define BAR_FULL(s) (bar0 + 8u * (s))
#define BAR_EMPTY(s) (bar0 + 16u + 8u * (s))
#define BAR_TFULL(b) (bar0 + 32u + 8u * (b))
#define BAR_TEMPTY(b) (bar0 + 48u + 8u * (b))
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(BAR_FULL(s), 128); mbar_init(BAR_EMPTY(s), 1); mbar_init(BAR_TFULL(s), 1); mbar_init(BAR_TEMPTY(s), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {   // one warp allocates the accumulator columns and owns their release
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ntiles = (nB + TC_N - 1) / TC_N;
    const GreedyArgs &g = p.g;

    if (warp >= 4 && warp < 8) {
        // ================= producers: expand descriptors into the operand tiles =================
        const int r = tid - 128;                                       // tile row handled by this thread
        const uint8_t *ad = tc_desc(p.A, f), *bd = tc_desc(p.B, f);
        {
            const int ar = a0 + r;
            const bool ok = ar < nA;
            const uint4 *src = reinterpret_cast<const uint4 *>(ad) + (size_t)(ok ? ar : 0) * 2;
            expand_row(tileA, r, src[0], src[1], ok);
        }
        const uint16_t *bidx = (MODE == TC_SHORT && p.b_index) ? p.b_index + (size_t)f * p.b_index_stride : nullptr;
        const int *ctime = MODE == TC_SCORES ? g.claim_time + (size_t)f * g.cols.stride_rows : nullptr;
        for (int t = 0; t < ntiles; ++t) {
            const int s = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbar_wait(BAR_EMPTY(s), ph ^ 1u);                          // the MMAs that read this stage are done
            mbar_wait(BAR_TEMPTY(s), ph ^ 1u);                         // and the epilogue no longer reads its side data
            const int j = t * TC_N + r;
            const bool ok = j < nB;
            const int col = ok ? (bidx ? (int)bidx[j] : j) : 0;
            const uint4 *src = reinterpret_cast<const uint4 *>(bd) + (size_t)col * 2;
            expand_row(tileB + s * TC_OPERAND_BYTES, r, src[0], src[1], ok);
            if (MODE == TC_SCORES) side[s * TC_N + r] = ok ? ctime[j] : INT_MIN;
            if (MODE == TC_SHORT) side[s * TC_N + r] = col;
            fence_proxy_async();                                       // generic-proxy writes -> visible to the tensor core's reads
            mbar_arrive(BAR_FULL(s));
        }
    } else if (warp == 8) {
        // ================= MMA issuer: one thread =================
        if (lane == 0) {
            const uint32_t aaddr = smem_u32(tileA), baddr = smem_u32(tileB);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t & 1;
                const uint32_t ph = (uint32_t)(t >> 1) & 1u;
                mbar_wait(BAR_TEMPTY(s), ph ^ 1u);                     // accumulator buffer drained by the epilogue
                mbar_wait(BAR_FULL(s), ph);                            // operands expanded
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {                          // K = 256 int8 = 8 x 32 bytes; 4 steps per 128-byte atom
                    const uint32_t off = (uint32_t)(k >> 2) * TC_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                    umma_i8(tmem_base + (uint32_t)s * TC_N, umma_desc(aaddr + off), umma_desc(baddr + (uint32_t)s * TC_OPERAND_BYTES + off), k > 0);
                }
                umma_commit(BAR_EMPTY(s));                             // arrives when the MMAs above have read their operands
                umma_commit(BAR_TFULL(s));                             // ... and written the accumulators
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: thread = one A row, columns in ascending order =================
        const int row = a0 + tid;                                      // tid 0..127 = TMEM lane
        const bool row_ok = row < nA;
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        const size_t ro = (size_t)f * g.rows.stride_rows, co = (size_t)f * g.cols.stride_rows;
        // ---- per-mode state
        int best = INT_MIN, best_chunk = -1;                           // TC_PAIRS: running maximum dot and the 32-column chunk it first appeared in
        int bdot = -256, sdot = -256, bidx = -1, grow = 0;             // TC_SCORES
        bool live = row_ok;
        int cnt = 0;                                                   // TC_SHORT
        const uint8_t *rl = nullptr;
        if (MODE == TC_PAIRS) rl = g.fp ? (g.use_live ? g.fp[f].prev_live : nullptr) : (g.row_live ? g.row_live + ro : nullptr);
        if (MODE == TC_SCORES) {
            const uint8_t *l2 = g.fp ? (g.use_live ? g.fp[f].prev_live : nullptr) : (g.row_live ? g.row_live + ro : nullptr);
            live = row_ok && (!l2 || l2[row]);
            grow = g.row_base + (g.row_base_arr ? g.row_base_arr[f] : 0) + row;
        }
        if (MODE == TC_SHORT) live = row_ok && p.row_need[ro + row];
        const int thr_dot = 256 - 2 * p.T;                             // d < T  <=>  dot > 256 - 2 T
        for (int t = 0; t < ntiles; ++t) {
            const int s = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            if (MODE == TC_SCORES || MODE == TC_SHORT) mbar_wait(BAR_FULL(s), ph);   // the side data written by the producers
            mbar_wait(BAR_TFULL(s), ph);
            tc_fence_after();
            const int *sd = side + s * TC_N;
#pragma unroll 1
            for (int c = 0; c < TC_N / 32; ++c) {
                const int j0 = t * TC_N + c * 32;
                if (j0 >= nB) break;                                   // warp-uniform
                int v[32];
                tmem_ld32(lane_base + (uint32_t)(s * TC_N + c * 32), v);
                const int nv = min(32, nB - j0);                       // valid columns of this chunk (32 except at the very end)
                if (MODE == TC_DUMP) {
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nv) p.dump[((size_t)f * p.dump_rows + row) * p.dump_pitch + j0 + i] = v[i];
                    }
                } else if (MODE == TC_PAIRS) {
                    if (nv < 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (i >= nv) v[i] = INT_MIN;
                    }
                    int m = v[0];
#pragma unroll
                    for (int i = 1; i < 32; ++i) m = max(m, v[i]);
                    if (row_ok && m > best) { best = m; best_chunk = t * (TC_N / 32) + c; }   // strict: the first chunk keeps a tie
                    if (__any_sync(0xffffffffu, row_ok && m > thr_dot)) {       // rare: a pass-1 candidate (d < 15) in this chunk
#pragma unroll                                                                   // (unrolled: v[] must stay in registers)
                        for (int i = 0; i < 32; ++i) {
                            const int pr = j0 + i;                              // previous-frame row; columns past nB hold INT_MIN
                            if (row_ok && v[i] > thr_dot && (!rl || rl[pr])) {
                                const int d = (256 - v[i]) >> 1;
                                const int pos = atomicAdd(g.short_cnt + ro + pr, 1);
                                if (pos < SVO_SHORT_CAP) *tc_short_slot(g, ro + pr, pos) = ((uint32_t)d << 16) | (uint32_t)row;
                            }
                        }
                    }
                } else if (MODE == TC_SCORES) {
                    // cheap reject: no column of the chunk beats the running best, whatever its claim time
                    int m = v[0];
#pragma unroll
                    for (int i = 1; i < 32; ++i) m = max(m, v[i]);
                    if (__any_sync(0xffffffffu, live && m > bdot)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const bool take = live && sd[c * 32 + i] >= grow && v[i] > bdot;   // unclaimed when this row scans, strictly better
                            if (take) { sdot = bdot; bdot = v[i]; bidx = j0 + i; }
                        }
                    }
                } else if (MODE == TC_SHORT) {
                    if (nv < 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (i >= nv) v[i] = INT_MIN;
                    }
                    int m = v[0];
#pragma unroll
                    for (int i = 1; i < 32; ++i) m = max(m, v[i]);
                    if (__any_sync(0xffffffffu, live && m > thr_dot)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (live && v[i] > thr_dot) {
                                if (cnt < SVO_SHORT_CAP) *tc_short_slot(g, ro + row, cnt) = ((uint32_t)((256 - v[i]) >> 1) << 16) | (uint32_t)sd[c * 32 + i];
                                ++cnt;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(BAR_TEMPTY(s));
        }
        // ---- per-row results
        if (MODE == TC_PAIRS && row_ok) {
            // the exact first minimum inside the winning chunk: 32 candidate rows, recomputed from the descriptors
            const uint8_t *ad = tc_desc(p.A, f), *bd = tc_desc(p.B, f);
            const uint4 *q = reinterpret_cast<const uint4 *>(ad) + (size_t)row * 2;
            const uint4 qa = q[0], qb = q[1];
            const int want = (256 - best) >> 1;
            int first = best_chunk * 32;
            const int end = min(first + 32, nB);
            for (; first < end; ++first) {
                const uint4 *tr = reinterpret_cast<const uint4 *>(bd) + (size_t)first * 2;
                if (popc256(qa, qb, tr[0], tr[1]) == want) break;
            }
            p.bf_key[co + row] = ((uint32_t)want << 16) | (uint32_t)first;
        }
        if (MODE == TC_SCORES && live) {
            g.best_idx[ro + row] = bidx; g.best[ro + row] = (256 - bdot) >> 1; g.second[ro + row] = (256 - sdot) >> 1;
        }
        if (MODE == TC_SHORT && row_ok) g.short_cnt[ro + row] = live ? cnt : 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS) : "memory");
    }
#undef BAR_FULL
#undef BAR_EMPTY
#undef BAR_TFULL
#undef BAR_TEMPTY
}

}  // namespace

int setup_tc_attributes()
{
    if (cudaFuncSetAttribute(k_tc_hamming<TC_PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tc_hamming<TC_SCORES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tc_hamming<TC_SHORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tc_hamming<TC_DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess) return 1;
    return 0;
}

void launch_tc_hamming(const TcArgs &p, int mode, int nframes, cudaStream_t st, long long *launches)
{
    const int maxA = p.A.count ? p.A.stride_rows : p.A.fixed_count;
    if (maxA <= 0 || nframes <= 0) return;
    dim3 grid((maxA + TC_M - 1) / TC_M, nframes);
    switch (mode) {
    case TC_PAIRS: k_tc_hamming<TC_PAIRS><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(p); break;
    case TC_SCORES: k_tc_hamming<TC_SCORES><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(p); break;
    case TC_SHORT: k_tc_hamming<TC_SHORT><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(p); break;
    default: k_tc_hamming<TC_DUMP><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(p); break;
    }
    ++*launches;
}
