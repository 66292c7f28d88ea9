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
    o_mul[i] = gl::lazy::canon(gl::lazy::mul_v1(a[i], b[i]));
}

// ---- second-generation primitives (mulc / add1 / sub1 / canon2 / mul_pow2 / dft_pow2) ----
__global__ void k2(const gl::u64* a, const gl::u64* b, int n, unsigned* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const gl::u32 eps = gl::lazy::k_eps32;
    const gl::u64 x = a[i], y = b[i];
    const gl::u64 xc = x >= gl::P ? x - gl::P : x, yc = y >= gl::P ? y - gl::P : y;
    if (gl::lazy::mulc(x, y, eps) != gl::mul(xc, yc)) atomicAdd(&bad[0], 1u);
    if (gl::lazy::canon2(x) != xc) atomicAdd(&bad[1], 1u);
    // one operand <= p (also try exactly p), the other arbitrary
    const gl::u64 yp = (i & 7) == 0 ? gl::P : yc;
    if (gl::lazy::canon2(gl::lazy::add1(x, yp)) != gl::add(xc, yc % gl::P * ((i & 7) != 0))) atomicAdd(&bad[2], 1u);
    if (gl::lazy::canon2(gl::lazy::add1(yp, x)) != gl::add(xc, yc % gl::P * ((i & 7) != 0))) atomicAdd(&bad[2], 1u);
    if (gl::lazy::canon2(gl::lazy::sub1(x, yp)) != gl::sub(xc, yc % gl::P * ((i & 7) != 0))) atomicAdd(&bad[3], 1u);
    gl::u64 pw = 1;
#pragma unroll
    for (int e = 1; e < 96; e++) {
        pw = gl::add(pw, pw);
        if (gl::lazy::mul_pow2(x, e, eps) != gl::mul(xc, pw)) atomicAdd(&bad[4], 1u);
    }
    if (gl::lazy::red3(x, (gl::u32)y, (gl::u32)(y >> 32), eps) !=
        gl::sub(gl::add(xc, gl::mul((gl::u32)y, gl::EPS)), (gl::u32)(y >> 32))) atomicAdd(&bad[5], 1u);
}
template <int K, bool INV>
__global__ void k3(const gl::u64* a, int n, unsigned* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((i + 1) * (1 << K) > n) return;
    const gl::u32 eps = gl::lazy::k_eps32;
    gl::u64 v[1 << K], in[1 << K];
#pragma unroll
    for (int t = 0; t < (1 << K); t++) {
        gl::u64 x = a[i * (1 << K) + t];
        x = x >= gl::P ? x - gl::P : x;
        if ((i & 3) == 0 && x == 0) x = gl::P;  // the transform accepts p itself as an input representative
        in[t] = v[t] = x;
    }
    gl::lazy::dft_pow2<K, INV>(v, eps);
    gl::u64 w = gl::root_2exp(K);
    if (INV) w = gl::inv(w);
    for (int k = 0; k < (1 << K); k++) {
        gl::u64 acc = 0, wk = gl::pow(w, k), x = 1;
        for (int t = 0; t < (1 << K); t++) {
            acc = gl::add(acc, gl::mul(in[t] % gl::P, x));
            x = gl::mul(x, wk);
        }
        gl::u64 got = 0;
#pragma unroll
        for (int t = 0; t < (1 << K); t++)
            if (t == gl::lazy::brev_bits(k, K)) got = v[t];
        if (gl::lazy::canon2(got) != acc) atomicAdd(&bad[6 + (INV ? 1 : 0)], 1u);
    }
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
    unsigned* dbad;
    cudaMalloc(&dbad, 8 * 4);
    cudaMemset(dbad, 0, 8 * 4);
    k2<<<(n + 255) / 256, 256>>>(da, db, n, dbad);
    const int nd = 1 << 16;  // elements fed to the register DFTs (edge x edge pairs first, then random)
    k3<5, false><<<nd / 32 / 64, 64>>>(da, nd, dbad);
    k3<5, true><<<nd / 32 / 64, 64>>>(da, nd, dbad);
    k3<4, false><<<nd / 16 / 64, 64>>>(db, nd, dbad);
    k3<3, true><<<nd / 8 / 64, 64>>>(db, nd, dbad);
    k3<2, false><<<nd / 4 / 64, 64>>>(da, nd, dbad);
    k3<1, true><<<nd / 2 / 64, 64>>>(da, nd, dbad);
    unsigned hb[8];
    cudaMemcpy(hb, dbad, 32, cudaMemcpyDeviceToHost);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error (gen 2): %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
    printf("gen2 bad mulc=%u canon2=%u add1=%u sub1=%u mul_pow2=%u red3=%u dft=%u idft=%u\n", hb[0], hb[1], hb[2], hb[3], hb[4], hb[5], hb[6], hb[7]);
    for (int i = 0; i < 8; i++)
        if (hb[i]) return 1;
    return (bad[0] || bad[1] || bad[2]) ? 1 : 0;
}
