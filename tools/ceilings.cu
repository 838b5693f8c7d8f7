// ceilings.cu — measures, on the box it runs on, the two SM-side ceilings the roofline of this path is quoted against
// ("measured peaks from a copy kernel / popc microbench on the same box"):
//   * warp-instruction issue rate (independent integer LOP3/IADD chains, 8 per thread, enough warps to fill every scheduler)
//   * POPC rate (the XU pipe: what bounded the SIMT Hamming kernels of round 1)
//   * MUFU-free integer min/max (VIMNMX) rate, the instruction the FAST and tensor-core epilogues lean on
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ceilings tools/ceilings.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_issue(unsigned *out, unsigned seed)
{
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // 8 independent dependency chains, 2 instructions each per round: 64 per iteration
            a0 = (a0 ^ a1) + 1u; a1 = (a1 ^ a2) + 3u; a2 = (a2 ^ a3) + 5u; a3 = (a3 ^ a4) + 7u;
            a4 = (a4 ^ a5) + 9u; a5 = (a5 ^ a6) + 11u; a6 = (a6 ^ a7) + 13u; a7 = (a7 ^ a0) + 15u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

__global__ void k_popc(unsigned *out, unsigned seed)
{
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // 4 POPC + 4 adds + 4 xors per round
            s0 += __popc(a0); s1 += __popc(a1); s2 += __popc(a2); s3 += __popc(a3);
            a0 ^= s1; a1 ^= s2; a2 ^= s3; a3 ^= s0;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 ^ s1 ^ s2 ^ s3;
}

__global__ void k_minmax(int *out, int seed)
{
    int a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = max(a0, a1 - 1); a1 = min(a1, a2 + 3); a2 = max(a2, a3 - 5); a3 = min(a3, a4 + 7);
            a4 = max(a4, a5 - 9); a5 = min(a5, a6 + 11); a6 = max(a6, a7 - 13); a7 = min(a7, a0 + 15);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

template <typename K, typename T>
static double run(K kern, T *buf, int blocks, int threads)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) kern<<<blocks, threads>>>(buf, w);
    double best = 1e30;
    for (int r = 0; r < 10; ++r) {
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(buf, r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const int sms = p.multiProcessorCount, threads = 1024, blocks = sms * 2;   // 64 warps per SM: every scheduler has 16
    unsigned *buf;
    cudaMalloc(&buf, sizeof(unsigned) * (size_t)blocks * threads);
    const double warps = (double)blocks * threads / 32;
    const double t_issue = run(k_issue, buf, blocks, threads);
    const double t_popc = run(k_popc, buf, blocks, threads);
    const double t_mm = run(k_minmax, (int *)buf, blocks, threads);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // instruction counts per thread: k_issue 64 per iteration (+3 loop), k_popc 8*(4 popc + 8 alu), k_minmax 8*16
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d, "
           "\"issue_warp_inst_per_s\": %.4g, \"issue_per_sm_per_clk_at_max_clock\": %.3f, "
           "\"popc_per_s\": %.4g, \"popc_per_sm_per_clk_at_max_clock\": %.3f, "
           "\"minmax_add_warp_inst_per_s\": %.4g, "
           "\"how\": \"best of 10 launches, %d blocks x %d threads, %d iterations; issue: 8 independent xor+add chains; "
           "popc: 4 POPC + 8 ALU per round (reports thread-level POPC/s)\"}\n",
           p.name, sms, clk, warps * ITERS * 64.0 / t_issue, warps * ITERS * 64.0 / t_issue / sms / (clk * 1e3),
           warps * 32 * ITERS * 32.0 / t_popc, warps * 32 * ITERS * 32.0 / t_popc / sms / (clk * 1e3),
           warps * ITERS * 128.0 / t_mm, blocks, threads, ITERS);
    return 0;
}
