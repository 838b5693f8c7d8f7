// hostio_probe2.cu — does it matter HOW results leave the GPU?  One process drives every GPU of the box at once:
// pinned H2D copies (copy engine) of the bench's per-step image block (30 MB) together with a 7-MB D2H per step done
// (a) by the copy engine (cudaMemcpyAsync), or (b) by a kernel storing into mapped pinned host memory (16-byte stores).
// Prints aggregate GB/s per direction for: H2D alone, D2H alone (both ways), and the two mixes.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/hostio_probe2.cu -o tools/hostio_probe2 && tools/hostio_probe2
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_store_host(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main(int argc, char **argv)
{
    int ng = 0;
    CK(cudaGetDeviceCount(&ng));
    if (argc > 1) ng = atoi(argv[1]) < ng ? atoi(argv[1]) : ng;
    const size_t NI = 30u << 20, NO = 7u << 20;
    const int REP = 12;
    std::vector<uint8_t *> hin(ng), hout(ng), din(ng), dout(ng), hout_dev(ng);
    std::vector<cudaStream_t> s1(ng), s2(ng);
    for (int g = 0; g < ng; ++g) {
        CK(cudaSetDevice(g));
        CK(cudaHostAlloc((void **)&hin[g], NI, cudaHostAllocPortable));
        CK(cudaHostAlloc((void **)&hout[g], NO, cudaHostAllocPortable | cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void **)&hout_dev[g], hout[g], 0));
        CK(cudaMalloc((void **)&din[g], NI)); CK(cudaMalloc((void **)&dout[g], NO));
        CK(cudaMemset(dout[g], 1, NO));
        CK(cudaStreamCreateWithFlags(&s1[g], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2[g], cudaStreamNonBlocking));
    }
    auto run = [&](bool h2d, int d2h /*0 none, 1 copy engine, 2 kernel*/) -> double {
        double best = 1e9;
        for (int it = 0; it < 4; ++it) {
            for (int g = 0; g < ng; ++g) { cudaSetDevice(g); cudaDeviceSynchronize(); }
            auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < REP; ++r)
                for (int g = 0; g < ng; ++g) {
                    cudaSetDevice(g);
                    if (h2d) cudaMemcpyAsync(din[g], hin[g], NI, cudaMemcpyHostToDevice, s1[g]);
                    if (d2h == 1) cudaMemcpyAsync(hout[g], dout[g], NO, cudaMemcpyDeviceToHost, s2[g]);
                    if (d2h == 2) k_store_host<<<148, 256, 0, s2[g]>>>((const uint4 *)dout[g], (uint4 *)hout_dev[g], NO / 16);
                }
            for (int g = 0; g < ng; ++g) { cudaSetDevice(g); cudaDeviceSynchronize(); }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            best = dt < best ? dt : best;
        }
        return best;
    };
    const double GI = (double)NI * REP * ng / 1e9, GO = (double)NO * REP * ng / 1e9;
    double t;
    printf("{\"gpus\": %d, \"h2d_bytes\": %zu, \"d2h_bytes\": %zu", ng, NI, NO);
    t = run(true, 0); printf(", \"h2d_alone_gbs\": %.1f", GI / t);
    t = run(false, 1); printf(", \"d2h_copy_engine_alone_gbs\": %.1f", GO / t);
    t = run(false, 2); printf(", \"d2h_kernel_stores_alone_gbs\": %.1f", GO / t);
    t = run(true, 1); printf(", \"mix_copy_engine\": {\"h2d_gbs\": %.1f, \"d2h_gbs\": %.1f, \"steps_per_s_per_gpu\": %.0f}", GI / t, GO / t, REP / t);
    t = run(true, 2); printf(", \"mix_kernel_stores\": {\"h2d_gbs\": %.1f, \"d2h_gbs\": %.1f, \"steps_per_s_per_gpu\": %.0f}", GI / t, GO / t, REP / t);
    printf("}\n");
    return 0;
}
