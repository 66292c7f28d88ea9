// Goldilocks field arithmetic for device and host (p = 2^64 - 2^32 + 1).
// Replaces the reference's `Fp64<P>` ops (crates/sezkp-ffts/src/lib.rs:57-133): same mathematical
// functions, canonical residues in [0,p) at every interface, but reduction uses the special form
// 2^64 ≡ 2^32-1, 2^96 ≡ -1 instead of the reference's `u128 % p`.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

namespace gl {

typedef uint64_t u64;
typedef uint32_t u32;

constexpr u64 P = 0xffffffff00000001ULL;
constexpr u64 EPS = 0xffffffffULL;  // 2^32 - 1 = 2^64 mod p

GL_HD u64 add(u64 a, u64 b) {  // canonical in -> canonical out
    u64 s = a + b;
    return (s < a || s >= P) ? s - P : s;
}
GL_HD u64 sub(u64 a, u64 b) {
    u64 d = a - b;
    return (a < b) ? d + P : d;
}
GL_HD u64 neg(u64 a) { return a ? P - a : 0; }

GL_HD void mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 pr = (unsigned __int128)a * b;
    lo = (u64)pr;
    hi = (u64)(pr >> 64);
#endif
}
// 128-bit -> canonical residue. x = lo + 2^64*(hl + 2^32*hh) ≡ lo - hh + hl*(2^32-1).
GL_HD u64 reduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= EPS;  // +p
    u64 t1 = hl * EPS;
    u64 t2 = t0 + t1;
    if (t2 < t1) t2 += EPS;  // 2^64 ≡ EPS
    return t2 >= P ? t2 - P : t2;
}
GL_HD u64 mul(u64 a, u64 b) {
    u64 lo, hi;
    mul_wide(a, b, lo, hi);
    return reduce128(lo, hi);
}
GL_HD u64 sqr(u64 a) { return mul(a, a); }
GL_HD u64 pow(u64 base, u64 e) {
    u64 acc = 1;
    while (e) {
        if (e & 1) acc = mul(acc, base);
        base = sqr(base);
        e >>= 1;
    }
    return acc;
}
GL_HD u64 inv(u64 a) { return pow(a, P - 2); }
GL_HD u64 from_u64(u64 x) { return x >= P ? x - P : x; }  // x < 2^64 < 2p
GL_HD u64 from_i64(int64_t x) { return x >= 0 ? from_u64((u64)x) : P - (u64)(-(x + 1)) - 1; }  // |x| < 2^63 < p
// primitive 2^k-th root of unity, reference convention g = 7 (sezkp-ffts/src/lib.rs:237-242)
GL_HD u64 root_2exp(unsigned k) { return pow(7, (P - 1) >> k); }

}  // namespace gl
