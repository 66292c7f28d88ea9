// Standalone device-arithmetic self-test: gl::lazy::{add,sub,mul} against host canonical arithmetic on edge and
// random operands.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/arith_test tools/arith_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../streaming-zero-knowledge-proofs_b200/csrc/gl.cuh"

typedef unsigned long long ull;
__global__ void k(const gl::u64* a, const gl::u64* b, int n, gl::u64* o_add, gl::u64* o_sub, gl::u64* o_mul) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    o_add[i] = gl::lazy::canon(gl::lazy::add(a[i], b[i]));
    o_sub[i] = gl::lazy::canon(gl::lazy::sub(a[i], b[i]));
    o_mul[i] = gl::lazy::canon(gl::lazy::mul(a[i], b[i]));
}
static gl::u64 rnd() {
    gl::u64 x = 0;
    for (int i = 0; i < 5; i++) x = (x << 15) ^ (gl::u64)rand();
    return x;
}
int main() {
    std::vector<gl::u64> edge = {0, 1, 2, gl::P - 1, gl::P, gl::P + 1, 0xffffffffULL, 0x100000000ULL, 0xffffffff00000000ULL,
                                 0xfffffffeffffffffULL, ~0ULL, ~0ULL - 1, 0x8000000000000000ULL, 0xffffffffULL << 31};
    std::vector<gl::u64> a, b;
    for (auto x : edge)
        for (auto y : edge) {
            a.push_back(x);
            b.push_back(y);
        }
    for (int i = 0; i < 1 << 20; i++) {
        gl::u64 x = rnd(), y = rnd();
        if (i & 1) x |= 0xffffffff00000000ULL;
        if (i & 2) y |= 0xffffffff00000000ULL;
        a.push_back(x);
        b.push_back(y);
    }
    int n = (int)a.size();
    gl::u64 *da, *db, *d1, *d2, *d3;
    cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&d1, n * 8); cudaMalloc(&d2, n * 8); cudaMalloc(&d3, n * 8);
    cudaMemcpy(da, a.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(da, db, n, d1, d2, d3);
    std::vector<gl::u64> o1(n), o2(n), o3(n);
    cudaMemcpy(o1.data(), d1, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(o2.data(), d2, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(o3.data(), d3, n * 8, cudaMemcpyDeviceToHost);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 2; }
    int bad[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) {
        gl::u64 x = a[i] % gl::P, y = b[i] % gl::P;
        gl::u64 e1 = gl::add(x, y), e2 = gl::sub(x, y), e3 = gl::mul(x, y);
        if (o1[i] != e1 && bad[0]++ < 5) printf("add  a=%016llx b=%016llx got=%016llx exp=%016llx\n", (ull)a[i], (ull)b[i], (ull)o1[i], (ull)e1);
        if (o2[i] != e2 && bad[1]++ < 5) printf("sub  a=%016llx b=%016llx got=%016llx exp=%016llx\n", (ull)a[i], (ull)b[i], (ull)o2[i], (ull)e2);
        if (o3[i] != e3 && bad[2]++ < 5) printf("mul  a=%016llx b=%016llx got=%016llx exp=%016llx\n", (ull)a[i], (ull)b[i], (ull)o3[i], (ull)e3);
    }
    printf("n=%d bad add=%d sub=%d mul=%d\n", n, bad[0], bad[1], bad[2]);
    return (bad[0] || bad[1] || bad[2]) ? 1 : 0;
}
