// Microbenchmark: throughput of the single-block BLAKE3 compression under different B3_SCHED pipe schedules
// (see csrc/blake3.cuh).  Every variant must produce the same digests.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/b3_bench tools/b3_bench.cu && tools/b3_bench
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../streaming-zero-knowledge-proofs_b200/csrc/blake3.cuh"

typedef uint32_t u32;
constexpr int CHAIN = 64;

template <u32 SCHED>
__global__ void __launch_bounds__(256) chain_kernel(u32* __restrict__ out, u32 seed) {
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    u32 m[16], d[8];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = seed * 0x9E3779B1u + tid * 16 + i;
#pragma unroll 1
    for (int it = 0; it < CHAIN; it++) {
        b3::hash_block<SCHED>(m, 64, d);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            m[i] = d[i];
            m[8 + i] ^= d[7 - i];
        }
    }
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= d[i];
    out[tid] = x;
}

template <u32 SCHED>
float run(u32* d_out, int blocks, std::vector<u32>& host) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    chain_kernel<SCHED><<<blocks, 256>>>(d_out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) chain_kernel<SCHED><<<blocks, 256>>>(d_out, 1);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    host.resize((size_t)blocks * 256);
    cudaMemcpy(host.data(), d_out, host.size() * 4, cudaMemcpyDeviceToHost);
    return ms / 5;
}

#define CASE(S)                                                                                              \
    {                                                                                                        \
        std::vector<u32> h;                                                                                  \
        const float ms = run<S>(d_out, blocks, h);                                                           \
        if (ref.empty()) ref = h;                                                                            \
        const bool same = h == ref;                                                                          \
        int wide = 0;                                                                                        \
        for (int r = 0; r < 7; r++) wide += 4 * __builtin_popcount(((S) >> (4 * r)) & 15);                   \
        printf("SCHED %07x  wide rotations %3d  %.3f ms  %.3e compressions/s  %s\n", (unsigned)(S), wide, ms, \
               (double)blocks * 256 * CHAIN / (ms * 1e-3), same ? "same" : "MISMATCH");                      \
    }

int main() {
    const int blocks = 148 * 6 * 8;
    u32* d_out;
    cudaMalloc(&d_out, (size_t)blocks * 256 * 4);
    std::vector<u32> ref;
    CASE(0x0000000u)
    CASE(0x0202020u)
    CASE(0x0202000u)
    CASE(0x0002020u)
    CASE(0x0200020u)
    CASE(0x2020202u)
    CASE(0x0808080u)
    CASE(0x8080808u)
    CASE(0x0208020u)
    CASE(0x0802080u)
    CASE(0x0202022u)
    CASE(0x2202020u)
    CASE(0x0222020u)
    CASE(0x0202220u)
    CASE(0x0101010u)
    CASE(0x0404040u)
    CASE(0x0204020u)
    CASE(0x0201020u)
    CASE(0x0a02020u)
    CASE(0x020a020u)
    CASE(0x0202028u)
    CASE(0x0602020u)
    CASE(0x0302020u)
    CASE(0x0202060u)
    CASE(0x2000002u)
    CASE(0x0200200u)
    CASE(0x0020020u)
    CASE(0x0220220u)
    CASE(0x0000000u)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
