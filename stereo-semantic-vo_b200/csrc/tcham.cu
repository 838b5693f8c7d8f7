// tcham.cu — the all-pairs 256-bit Hamming distances of the batch matchers on the 5th-generation tensor cores.
//
// A 256-bit descriptor becomes 256 int8 values, +1 for a clear bit and -1 for a set bit; the dot product of two such
// rows is (#equal bits) - (#different bits) = 256 - 2 d, so one tcgen05.mma.kind::i8 tile (128 x 128 x K = 256, s32
// accumulators in tensor memory) holds 16384 exact distances that the SIMT kernels of match.cu spend ~40 instructions
// each on (8 XOR, carry-save adders, 5-6 POPC on the XU pipe, adds).  What is left for the CUDA cores is an epilogue of
// about one instruction per distance.  Integer arithmetic throughout: results are bit-exact.
//
// Operands.  k_tc_expand writes every descriptor set once per batch as "operand images": per frame, per group of 128
// descriptors, a 32 KB block that IS the shared-memory image tcgen05.mma reads (two K atoms of 128 rows x 128 bytes,
// K-major, 16-byte chunks XOR-swizzled by the row: the canonical SWIZZLE_128B layout).  A tile is therefore one
// contiguous bulk copy (cp.async.bulk -> mbarrier complete_tx), no tensor map, no per-tile expansion work.
//
// One CTA owns 128 "A" rows (tile rows = TMEM lanes) of one frame and runs TC_STREAMS (2) independent streams over the
// frame's "B" tiles: stream q owns the q-th contiguous share of the tiles, one shared-memory stage, one 128-column
// accumulator buffer and one epilogue warpgroup; two such CTAs fit an SM (111 KB of shared memory and 256 of the 512
// TMEM columns each):
//     warp 8      one thread issues the bulk copies (A once, then every stream's tiles) and waits on empty[stream]
//     warp 9      one thread issues 8 x tcgen05.mma (K = 32 bytes each) per tile, round-robin over the streams, and
//                 commits to empty[stream] / tfull[stream]; the warp also allocates and frees the TMEM columns
//     warps 0-7   warp w serves stream w / 4 and TMEM lane quarter w % 4: tcgen05.ld 32 lanes x 32 columns at a time,
//                 the mode's epilogue, arrive on tempty[stream]
// The tensor pipe needs ~512 cycles per tile; the epilogues need several thousand issue slots per tile, so the kernel is
// bound by them: sixteen epilogue warps per SM (four per scheduler) keep the issue slots busy, where a single warpgroup
// left them idle behind instruction latencies (measured: 3x).  Measured in the three-lane pipeline (gpurun_out/
// bench_r2{r,s,t}*.json): one CTA of four streams per SM (197 KB: nothing of another lane fits beside it) 24.3-24.4 k
// frames/s, two CTAs of two streams 24.8-24.9 k, three CTAs of one stream 24.6-24.7 k.  Inside a stream a thread meets
// its row's columns in ascending order, which is the order of the reference's scans (src/pnpmatch.cc:79-95, :177-190);
// the contiguous ranges of a row compose in stream order exactly like the lane blocks of match.cu's k_scores.
//
// Modes (what the epilogue does with dot = 256 - 2 d):
//   TC_PAIRS   A = current frame (BFMatcher queries), B = previous frame.  Per query the first minimum over the train
//              rows (cv::BFMatcher, src/pnpmatch.cc:266,278) and, for pass 1, every (row, column) with d < 15 appended
//              to the row's short list (src/pnpmatch.cc:101).  Replaces k_pairs and the u8 distance matrix.
//   TC_SCORES  A = previous frame rows, B = current columns + their pass-1 claim times: the exact
//              (bestIdx2, bestDist, secondBestDist) of every live row as the sequential scan saw them
//              (match_score, src/pnpmatch.cc:99).  Replaces k_scores_m.
//   TC_SHORT   A = local-map rows, B = the columns pass 1 left free (k_free_cols), gathered in ascending order by
//              k_tc_expand: the ascending list of columns with d < 60 per live row (pass 2, src/pnpmatch.cc:160-199).
//              Replaces k_shortlist / k_reuse in the batch path.
//   TC_DUMP    every dot product to global memory (bring-up / test tap).
#include "svo_internal.cuh"
#include <limits.h>

#define TC_M 128
#define TC_N 128
#define TC_STREAMS SVO_TC_STREAMS        // independent (stage, accumulator buffer, epilogue warpgroup) pipelines per CTA
#define TC_EPI_WARPS (4 * TC_STREAMS)
#define TC_THREADS (32 * (TC_EPI_WARPS + 2))
#define TC_ATOM_BYTES (128 * 128)
#define TC_TMEM_COLS (TC_STREAMS * TC_N) // 512: all of the SM's tensor memory
// A tile + one B stage per stream, alignment slack, per-warp claim-time staging, per-row combine buffers, barriers
#define TC_WARP_SCRATCH 320              // ints per epilogue warp: claim times of a tile (TC_SCORES) / a chunk's distance bytes [8][32] + columns [32]
#define TC_HITCAP 128                    // TC_PAIRS: pass-1 candidates buffered per CTA before they go to the rows' lists
#define TC_SMEM_BYTES ((1 + TC_STREAMS) * SVO_TC_TILE_BYTES + 1024 + TC_EPI_WARPS * TC_WARP_SCRATCH * 4 + TC_STREAMS * TC_M * 12 + TC_HITCAP * 8 + 256)

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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// global -> shared bulk copy (TMA, no tensor map) completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
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

// store that the compiler cannot turn into a branch: the epilogues are straight-line code (a warp-uniform branch per
// element costs ~100 cycles of dependent latency, measured; predicated instructions cost an issue slot)
__device__ __forceinline__ void st_global_if(uint32_t *ptr, uint32_t val, bool pred)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "@p st.global.u32 [%0], %1;\n\t"
        "}\n" ::"l"(ptr), "r"(val), "r"((uint32_t)pred) : "memory");
}
// maximum of the 32 accumulators of a chunk as a tree of three-input maxima (VIMNMX3): 16 instructions, depth 4
__device__ __forceinline__ int max32(const int (&v)[32])
{
    int a[12];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = __vimax3_s32(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    a[10] = v[30]; a[11] = v[31];
    return max(__vimax3_s32(__vimax3_s32(a[0], a[1], a[2]), __vimax3_s32(a[3], a[4], a[5]), __vimax3_s32(a[6], a[7], a[8])),
               __vimax3_s32(a[9], a[10], a[11]));
}

// The low byte of every accumulator of a chunk, 4 per word, word k of lane l at [k][l] (conflict-free): enough to
// recover the distance of a HIT, whose dot product lies in (136, 256] — d = ((256 - byte) & 255) >> 1 — while letting
// the few hits be picked by a runtime index (registers cannot be indexed dynamically).
__device__ __forceinline__ void spill_low_bytes(int *scratch, int lane, const int (&v)[32])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t lo = __byte_perm((uint32_t)v[4 * k], (uint32_t)v[4 * k + 1], 0x0040);
        const uint32_t hi = __byte_perm((uint32_t)v[4 * k + 2], (uint32_t)v[4 * k + 3], 0x0040);
        scratch[k * 32 + lane] = (int)__byte_perm(lo, hi, 0x5410);
    }
}
__device__ __forceinline__ uint32_t hit_distance(const int *scratch, int lane, int e)
{
    const uint32_t b = ((uint32_t)scratch[(e >> 2) * 32 + lane] >> (8 * (e & 3))) & 0xffu;
    return ((256u - b) & 0xffu) >> 1;
}

// 4 descriptor bits -> 4 int8: +1 (0x01) for a clear bit, -1 (0xFF) for a set bit.  The multiply copies bit k to
// bit 8k (the four shifted copies of the nibble occupy disjoint bit ranges, so nothing carries).
__device__ __forceinline__ uint32_t expand4(uint32_t nib)
{
    const uint32_t x = (nib * 0x00204081u) & 0x01010101u;
    return (x * 0xFFu) | 0x01010101u;
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

// ---------------------------------------------------------------------------------------------------------------------
// Operand images.  Thread = one 16-byte chunk (16 descriptor bits) of one descriptor; the 8 chunks of a 128-byte image
// row are written by 8 consecutive threads (one full line).  Row r of an image tile lies in 8-row groups of 1024 bytes;
// inside a group the chunk index is XORed with the row (Swizzle<3,4,3>).  Chunks 0-7 (bits 0-127) are K atom 0, chunks
// 8-15 K atom 1.  Rows between the set's count and the end of its last tile are written as zeros (dot product 0; the
// epilogues mask them); `index` gathers rows (the free columns of pass 2, ascending).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tc_expand(TcExpandArgs e0, TcExpandArgs e1)
{
    const TcExpandArgs &e = blockIdx.z ? e1 : e0;      // up to two descriptor sets per launch
    const int f = blockIdx.y;
    int n = tc_count(e.set, f);
    if (e.index || e.index32) n = min(n, e.index_cnt[(size_t)f * (e.index_cnt_stride ? e.index_cnt_stride : 1)]);
    const int ntile_rows = ((n + TC_M - 1) / TC_M) * TC_M;
    const uint8_t *src = tc_desc(e.set, f);
    const uint16_t *idx = e.index ? e.index + (size_t)f * e.index_stride : nullptr;
    const int *idx32 = e.index32 ? e.index32 + (size_t)f * e.index_stride : nullptr;
    uint8_t *img = e.img + (size_t)f * e.img_frame_stride;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < ntile_rows * 16; i += gridDim.x * 256) {
        const int r = i >> 4, c = i & 15;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (r < n) {
            const int sr = idx ? (int)idx[r] : (idx32 ? idx32[r] : r);
            const uint32_t h = reinterpret_cast<const uint16_t *>(src + (size_t)sr * 32)[c];
            o.x = expand4(h & 15u); o.y = expand4((h >> 4) & 15u); o.z = expand4((h >> 8) & 15u); o.w = expand4(h >> 12);
        }
        const int tr = r & (TC_M - 1);
        uint8_t *tile = img + (size_t)(r / TC_M) * SVO_TC_TILE_BYTES;
        *reinterpret_cast<uint4 *>(tile + (c >> 3) * TC_ATOM_BYTES + (tr >> 3) * 1024 + (tr & 7) * 128 + (((c & 7) ^ (tr & 7)) << 4)) = o;
    }
}

// Across the CTA's streams a row's columns are split into contiguous ranges, stream 0 first, so per-row results
// compose in stream order exactly like the lane blocks of match.cu's k_scores.
// Optional in-kernel timeline (TcArgs.prof != NULL, CTA (0, 0) only): clock64 stamps per role, read back by
// svo_debug_tc_profile.  Layout per mode: [role][64] with role 0 = producer, 1 = MMA issuer, 2 = epilogue warp 0.
#define TC_STAMP(role, slot) do { if (prof && (slot) < 64) prof[(role) * 64 + (slot)] = clock64(); } while (0)

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 2) k_tc_hamming(TcArgs p)
{
    extern __shared__ uint8_t tc_smem_raw[];
    const int f = blockIdx.y;
    const int a0 = blockIdx.x * TC_M;
    int nA = tc_count(p.A, f);
    if (MODE == TC_SHORT && p.a_index) nA = min(nA, p.a_index_cnt[2 * f]);     // the scanned rows only, gathered by k_tc_expand
    int nB = tc_count(p.B, f);
    if (MODE == TC_SHORT && p.b_index) nB = min(p.b_index_cnt[f], nB);
    if (a0 >= nA || nB <= 0) return;                                   // whole CTA, before anything is allocated
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long *prof = (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? p.prof + MODE * 256 : nullptr;
    if (tid == 0) TC_STAMP(3, 0);
    // operand tiles need 1024-byte alignment (the swizzle pattern repeats every 8 rows x 128 bytes)
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *tileA = smem, *tileB = smem + SVO_TC_TILE_BYTES;          // tileB: one stage per stream
    int *ctw = reinterpret_cast<int *>(smem + (1 + TC_STREAMS) * SVO_TC_TILE_BYTES);   // [epilogue warp][TC_N] claim times of the warp's current tile (TC_SCORES)
    int *comb = ctw + TC_EPI_WARPS * TC_WARP_SCRATCH;                  // [stream][TC_M][3] per-row partial results
    uint2 *hits = reinterpret_cast<uint2 *>(comb + TC_STREAMS * TC_M * 3);             // TC_PAIRS: (previous-frame row, d << 16 | column) candidates
    uint64_t *bars = reinterpret_cast<uint64_t *>(hits + TC_HITCAP);                   // afull, full[], empty[], tfull[], tempty[]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 1 + 4 * TC_STREAMS);
    int *hit_cnt = reinterpret_cast<int *>(tmem_slot + 1);
    const uint32_t bar0 = smem_u32(bars);
#define BAR_AFULL (bar0)
#define BAR_FULL(q) (bar0 + 8u + 8u * (q))
#define BAR_EMPTY(q) (bar0 + 8u + 8u * (TC_STREAMS + (q)))
#define BAR_TFULL(q) (bar0 + 8u + 8u * (2 * TC_STREAMS + (q)))
#define BAR_TEMPTY(q) (bar0 + 8u + 8u * (3 * TC_STREAMS + (q)))
    if (tid == 0) {
        *hit_cnt = 0;
        mbar_init(BAR_AFULL, 1);
        for (int q = 0; q < TC_STREAMS; ++q) { mbar_init(BAR_FULL(q), 1); mbar_init(BAR_EMPTY(q), 1); mbar_init(BAR_TFULL(q), 1); mbar_init(BAR_TEMPTY(q), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_EPI_WARPS + 1) {   // one warp allocates the accumulator columns and owns their release
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) TC_STAMP(3, 1);
    const int ntiles = (nB + TC_N - 1) / TC_N;
    const int per = (ntiles + TC_STREAMS - 1) / TC_STREAMS;            // tiles per stream: stream q owns tiles [q per, min((q + 1) per, ntiles))
    const GreedyArgs &g = p.g;

    if (warp == TC_EPI_WARPS) {
        // ================= producer: one thread, bulk copies of ready-made operand images =================
        if (lane == 0) {
            const uint8_t *aimg = p.a_img + (size_t)f * p.a_img_frame_stride + (size_t)blockIdx.x * SVO_TC_TILE_BYTES;
            const uint8_t *bimg = p.b_img + (size_t)f * p.b_img_frame_stride;
            const uint32_t baddr = smem_u32(tileB);
            mbar_expect_tx(BAR_AFULL, SVO_TC_TILE_BYTES);
            bulk_g2s(smem_u32(tileA), aimg, SVO_TC_TILE_BYTES, BAR_AFULL);
            int ps = 0;
            for (int i = 0; i < per; ++i)
                for (int q = 0; q < TC_STREAMS; ++q) {
                    const int t = q * per + i;
                    if (t >= ntiles) continue;
                    mbar_wait(BAR_EMPTY(q), ((uint32_t)i & 1u) ^ 1u);         // the MMAs that read this stage are done
                    mbar_expect_tx(BAR_FULL(q), SVO_TC_TILE_BYTES);
                    bulk_g2s(baddr + (uint32_t)q * SVO_TC_TILE_BYTES, bimg + (size_t)t * SVO_TC_TILE_BYTES, SVO_TC_TILE_BYTES, BAR_FULL(q));
                    TC_STAMP(0, ps); ++ps;
                }
        }
        __syncwarp();
    } else if (warp == TC_EPI_WARPS + 1) {
        // ================= MMA issuer: one thread, round-robin over the streams =================
        if (lane == 0) {
            const uint32_t aaddr = smem_u32(tileA), baddr = smem_u32(tileB);
            mbar_wait(BAR_AFULL, 0);
            int ms = 0;
            for (int i = 0; i < per; ++i)
                for (int q = 0; q < TC_STREAMS; ++q) {
                    if (q * per + i >= ntiles) continue;
                    const uint32_t ph = (uint32_t)i & 1u;
                    mbar_wait(BAR_TEMPTY(q), ph ^ 1u);                 // accumulator buffer drained by the stream's epilogue
                    TC_STAMP(1, ms); ++ms;
                    mbar_wait(BAR_FULL(q), ph);                        // operands landed
                    TC_STAMP(1, ms); ++ms;
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {                      // K = 256 int8 = 8 x 32 bytes; 4 steps per 128-byte atom
                        const uint32_t off = (uint32_t)(k >> 2) * TC_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                        umma_i8(tmem_base + (uint32_t)q * TC_N, umma_desc(aaddr + off), umma_desc(baddr + (uint32_t)q * SVO_TC_TILE_BYTES + off), k > 0);
                    }
                    umma_commit(BAR_EMPTY(q));                         // arrives when the MMAs above have read their operands
                    umma_commit(BAR_TFULL(q));                         // ... and written the accumulators
                }
        }
        __syncwarp();
    } else {
        // ================= epilogue: thread = one A row of one stream, columns in ascending order =================
        const int q = warp >> 2;                                       // stream; warp & 3 = TMEM lane quarter (== warp % 4)
        const int trow = (warp & 3) * 32 + lane;
        const bool row_ok = a0 + trow < nA;
        int row = a0 + trow;
        if (MODE == TC_SHORT && p.a_index) row = row_ok ? p.a_index[(size_t)f * g.rows.stride_rows + row] : 0;   // tile row -> map row
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)q * TC_N;
        const size_t ro = (size_t)f * g.rows.stride_rows, co = (size_t)f * g.cols.stride_rows;
        const int t0 = q * per, nq = max(0, min(per, ntiles - t0));
        // ---- per-mode state
        int best = INT_MIN, best_col = INT_MAX;                        // TC_PAIRS: running maximum dot and the column it first appeared in
        int bdot = -256, sdot = -256, bidx = -1, grow = 0;             // TC_SCORES
        bool live = row_ok;
        int cnt = 0;                                                   // TC_SHORT
        const uint8_t *rl = nullptr;
        const int *ctime = nullptr;
        int *ct = ctw + warp * TC_WARP_SCRATCH;
        if (MODE == TC_PAIRS) rl = g.fp ? (g.use_live ? g.fp[f].prev_live : nullptr) : (g.row_live ? g.row_live + ro : nullptr);
        if (MODE == TC_SCORES) {
            const uint8_t *l2 = g.fp ? (g.use_live ? g.fp[f].prev_live : nullptr) : (g.row_live ? g.row_live + ro : nullptr);
            live = row_ok && (!l2 || l2[row]);
            grow = g.row_base + (g.row_base_arr ? g.row_base_arr[f] : 0) + row;
            ctime = g.claim_time + co;
        }
        const uint16_t *bcol = nullptr;
        if (MODE == TC_SHORT) {
            live = row_ok && (p.a_index || p.row_need[ro + row]);
            bcol = p.b_index ? p.b_index + (size_t)f * p.b_index_stride : nullptr;
        }
        if (MODE == TC_SCORES && !live) grow = INT_MAX;                // a dead row takes nothing
        const int thr_dot = 256 - 2 * p.T;                             // d < T  <=>  dot > 256 - 2 T
        for (int i = 0; i < nq; ++i) {
            const int t = t0 + i;
            if (MODE == TC_SCORES) {                                   // this warp's copy of the tile's claim times (broadcast reads below)
                __syncwarp();
#pragma unroll
                for (int u = 0; u < TC_N / 32; ++u) {
                    const int j = t * TC_N + u * 32 + lane;
                    ct[u * 32 + lane] = j < nB ? ctime[j] : INT_MIN;   // INT_MIN: never visible to a row
                }
                __syncwarp();
            }
            if (warp == 0) TC_STAMP(2, 6 * i);
            mbar_wait(BAR_TFULL(q), (uint32_t)i & 1u);
            tc_fence_after();
            if (warp == 0) TC_STAMP(2, 6 * i + 1);
#pragma unroll 1
            for (int c = 0; c < TC_N / 32; ++c) {
                const int j0 = t * TC_N + c * 32;
                if (j0 >= nB) break;                                   // warp-uniform
                int v[32];
                tmem_ld32(lane_base + (uint32_t)(c * 32), v);
                const int nv = min(32, nB - j0);                       // valid columns of this chunk (32 except at the very end)
                if (warp == 0) TC_STAMP(2, 6 * i + 2 + c);
                if (MODE == TC_DUMP) {
                    if (row_ok) {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (e < nv) p.dump[((size_t)f * p.dump_rows + row) * p.dump_pitch + j0 + e] = v[e];
                    }
                } else if (MODE == TC_PAIRS) {
                    if (nv < 32) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) if (e >= nv) v[e] = -4096;
                    }
                    // maximum and its FIRST position in one tree: key = dot * 32 + (31 - e), |dot| <= 256
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = v[e] * 32 + (31 - e);
                    const int mk = max32(v);
                    const int m = mk >> 5;
                    if (row_ok && m > best) { best = m; best_col = j0 + 31 - (mk & 31); }   // strict: the first chunk keeps a tie
                    if (__any_sync(0xffffffffu, row_ok && m > thr_dot)) {       // a pass-1 candidate (d < 15) somewhere in this chunk (1 in 4)
                        uint32_t mask = 0;                                      // straight-line hit mask, then only the few set bits are visited
#pragma unroll
                        for (int e = 0; e < 32; ++e) mask |= ((v[e] >> 5) > thr_dot ? 1u : 0u) << e;
                        if (!row_ok) mask = 0;
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] >>= 5;                // back to dot products for the byte spill
                        __syncwarp();
                        spill_low_bytes(ct, lane, v);
                        __syncwarp();
                        while (mask) {
                            const int e = __ffs((int)mask) - 1;
                            mask &= mask - 1;
                            const int pr = j0 + e;                              // previous-frame row
                            if (!rl || rl[pr]) {
                                const uint32_t ent = (hit_distance(ct, lane, e) << 16) | (uint32_t)row;
                                const int hp = atomicAdd(hit_cnt, 1);           // shared-memory counter: the global lists are updated once, at the end
                                if (hp < TC_HITCAP) hits[hp] = make_uint2((uint32_t)pr, ent);
                                else {
                                    const int pos = atomicAdd(g.short_cnt + ro + pr, 1);
                                    if (pos < SVO_SHORT_CAP) *tc_short_slot(g, ro + pr, pos) = ent;
                                }
                            }
                        }
                    }
                } else if (MODE == TC_SCORES) {
                    // the reference's running update (src/pnpmatch.cc:89-94) over the columns still unclaimed when this
                    // row scans, as straight-line selects
                    const int4 *c4 = reinterpret_cast<const int4 *>(ct + c * 32);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int4 tm = c4[k];
                        const int tt[4] = {tm.x, tm.y, tm.z, tm.w};
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int x = v[4 * k + b];
                            const bool take = (x > bdot) & (tt[b] >= grow);
                            sdot = take ? bdot : sdot; bidx = take ? j0 + 4 * k + b : bidx; bdot = take ? x : bdot;
                        }
                    }
                } else if (MODE == TC_SHORT) {
                    // every column with d < 60 goes to the row's segment of this stream, in scan (= ascending) order: a
                    // straight-line hit mask (2 instructions per column), then only the set bits are visited; their distance
                    // comes from the spilled low bytes, their column from the chunk's slice of the gathered column list.
                    // (Deciding per chunk on its maximum first — 16 instructions, then a vote — and building the mask only
                    // when some row has a hit lowers the instruction count but not the time: 26.6 k against 26.8 k frames/s,
                    // the kernel alone 128 against 116 us; the vote puts the TMEM load's latency on every chunk's critical
                    // path.  profiles/r2_experiments.md)
                    int mycol = j0 + lane;                             // the column behind position j0 + lane of the gathered list
                    if (bcol && j0 + lane < nB) mycol = bcol[j0 + lane];
                    uint32_t mask = 0;
#pragma unroll
                    for (int e = 0; e < 32; ++e) mask |= (v[e] > thr_dot ? 1u : 0u) << e;
                    if (nv < 32) mask &= (1u << nv) - 1u;
                    if (!live) mask = 0;
                    __syncwarp();
                    spill_low_bytes(ct, lane, v);
                    ct[256 + lane] = mycol;
                    __syncwarp();
                    while (mask) {
                        const int e = __ffs((int)mask) - 1;
                        mask &= mask - 1;
                        if (cnt < SVO_TC_SEG) *tc_short_slot(g, ro + row, q * SVO_TC_SEG + cnt) = (hit_distance(ct, lane, e) << 16) | (uint32_t)ct[256 + e];
                        ++cnt;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(BAR_TEMPTY(q));
        }
        // ---- this stream's share of the row
        int *cb = comb + (q * TC_M + trow) * 3;
        if (MODE == TC_PAIRS) { cb[0] = best; cb[1] = best_col; }
        if (MODE == TC_SCORES) { cb[0] = bdot; cb[1] = sdot; cb[2] = bidx; }
        if (MODE == TC_SHORT) cb[0] = min(cnt, 255);
    }
    if (tid == 0) TC_STAMP(3, 2);
    tc_fence_before();
    __syncthreads();
    if (tid == 0) TC_STAMP(3, 3);
    if (warp == TC_EPI_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS) : "memory");
    }
    if (MODE == TC_PAIRS) {   // the buffered pass-1 candidates go to their rows' lists (arrival order; the resolver sorts each list)
        const size_t ro = (size_t)f * g.rows.stride_rows;
        const int nh = min(*hit_cnt, TC_HITCAP);
        for (int h = tid; h < nh; h += TC_THREADS) {
            const uint2 e = hits[h];
            const int pos = atomicAdd(g.short_cnt + ro + e.x, 1);
            if (pos < SVO_SHORT_CAP) *tc_short_slot(g, ro + e.x, pos) = e.y;
        }
    }
    // ================= per-row results: the streams' contiguous column ranges composed in order =================
    if (tid < TC_M && a0 + tid < nA) {
        const int row = a0 + tid;
        const size_t ro = (size_t)f * g.rows.stride_rows, co = (size_t)f * g.cols.stride_rows;
        if (MODE == TC_PAIRS) {
            int best = INT_MIN, best_col = INT_MAX;
            for (int q = 0; q < TC_STREAMS; ++q) {
                const int *cb = comb + (q * TC_M + tid) * 3;
                if (cb[0] > best) { best = cb[0]; best_col = cb[1]; }          // strict: an earlier range keeps a tie
            }
            p.bf_key[co + row] = ((uint32_t)((256 - best) >> 1) << 16) | (uint32_t)best_col;
        }
        if (MODE == TC_SCORES) {
            const uint8_t *l2 = g.fp ? (g.use_live ? g.fp[f].prev_live : nullptr) : (g.row_live ? g.row_live + ro : nullptr);
            if (!l2 || l2[row]) {
                int B = -256, S = -256, I = -1;
                for (int q = 0; q < TC_STREAMS; ++q) {
                    const int *cb = comb + (q * TC_M + tid) * 3;
                    // the range's last record beats the running best: "second" is whatever was best just before it
                    if (cb[0] > B) { S = max(B, cb[1]); B = cb[0]; I = cb[2]; }
                }
                g.best_idx[ro + row] = I; g.best[ro + row] = (256 - B) >> 1; g.second[ro + row] = (256 - S) >> 1;
            }
        }
        if (MODE == TC_SHORT) {
            // per-stream segment lengths, one byte each (k_merge_prune_lists joins the segments, ascending by construction)
            const int mrow = p.a_index ? p.a_index[ro + row] : row;              // rows outside a gathered list keep the 0 of k_greedy_init
            uint32_t packed = 0;
            if (p.a_index || p.row_need[ro + row])
                for (int q = 0; q < TC_STREAMS; ++q) packed |= (uint32_t)comb[(q * TC_M + tid) * 3] << (8 * q);
            g.short_cnt[ro + mrow] = (int)packed;
        }
    }
    if (tid == 0) TC_STAMP(3, 4);
#undef BAR_AFULL
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

void launch_tc_expand(const TcExpandArgs &e, int nframes, cudaStream_t st, long long *launches)
{
    const int maxn = e.set.count ? e.set.stride_rows : e.set.fixed_count;
    if (maxn <= 0 || nframes <= 0) return;
    int gx = (maxn * 16 + 255) / 256;
    if (gx > 64) gx = 64;
    k_tc_expand<<<dim3(gx, nframes, 1), 256, 0, st>>>(e, e);
    ++*launches;
}

// two sets in one launch (the operand images of both sides of a matcher)
void launch_tc_expand2(const TcExpandArgs &e0, const TcExpandArgs &e1, int nframes, cudaStream_t st, long long *launches)
{
    const int m0 = e0.set.count ? e0.set.stride_rows : e0.set.fixed_count, m1 = e1.set.count ? e1.set.stride_rows : e1.set.fixed_count;
    if (m0 <= 0 || nframes <= 0) { launch_tc_expand(e1, nframes, st, launches); return; }
    if (m1 <= 0) { launch_tc_expand(e0, nframes, st, launches); return; }
    int gx = ((m0 > m1 ? m0 : m1) * 16 + 255) / 256;
    if (gx > 64) gx = 64;
    k_tc_expand<<<dim3(gx, nframes, 2), 256, 0, st>>>(e0, e1);
    ++*launches;
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
