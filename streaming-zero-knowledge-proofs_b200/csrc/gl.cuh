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

// ---------------------------------------------------------------------------------------------
// Device-only "lazy" arithmetic for kernel inner loops: operands and results are arbitrary 64-bit
// representatives (in [0, 2^64), not necessarily < p); gl::lazy::canon() brings a value back to the
// canonical residue before it is stored or hashed.  Carry chains are written in PTX so that the
// 2^64 ≡ 2^32-1 corrections cost one masked add instead of compare + select sequences.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
namespace gl {
namespace lazy {

__device__ __forceinline__ u64 canon(u64 x) { return x >= P ? x - P : x; }

// a + b for arbitrary representatives (two masked corrections: the first can wrap again only when both
// operands lie in the top 2^32 of the range).
__device__ __forceinline__ u64 add(u64 a, u64 b) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, m;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        "add.cc.u32 a0, a0, b0;\n\t"
        "addc.cc.u32 a1, a1, b1;\n\t"
        "addc.u32 m, 0, 0;\n\t"         // carry (0/1); NB: `subc` after an add chain sees the inverted flag in hardware
        "neg.s32 m, m;\n\t"             // EPS * carry
        "add.cc.u32 a0, a0, m;\n\t"
        "addc.cc.u32 a1, a1, 0;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "neg.s32 m, m;\n\t"
        "add.cc.u32 a0, a0, m;\n\t"
        "addc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0, a1};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b));
    return r;
}
// a - b for arbitrary representatives.
__device__ __forceinline__ u64 sub(u64 a, u64 b) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, m;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        "sub.cc.u32 a0, a0, b0;\n\t"
        "subc.cc.u32 a1, a1, b1;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // m = borrow ? 0xffffffff : 0
        "sub.cc.u32 a0, a0, m;\n\t"
        "subc.cc.u32 a1, a1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 a0, a0, m;\n\t"
        "subc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0, a1};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b));
    return r;
}
// 128-bit value (w0..w3, little-endian 32-bit words) -> 64-bit representative:
//   x ≡ (w0 + 2^32 w1) - w3 + w2*(2^32 - 1)
__device__ __forceinline__ u64 reduce_words(u32 w0, u32 w1, u32 w2, u32 w3) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0, x1, y0, y1, m;\n\t"
        "sub.cc.u32 x0, %1, %4;\n\t"      // (w1:w0) - w3
        "subc.cc.u32 x1, %2, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 x0, x0, m;\n\t"       // borrow: -EPS (cannot borrow again)
        "subc.u32 x1, x1, 0;\n\t"
        "sub.cc.u32 y0, 0, %3;\n\t"       // w2*(2^32-1) = (w2 << 32) - w2
        "subc.u32 y1, %3, 0;\n\t"
        "add.cc.u32 x0, x0, y0;\n\t"
        "addc.cc.u32 x1, x1, y1;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "neg.s32 m, m;\n\t"
        "add.cc.u32 x0, x0, m;\n\t"       // carry: +EPS (cannot carry again: y <= 2^64 - 2^33 + 1)
        "addc.u32 x1, x1, 0;\n\t"
        "mov.b64 %0, {x0, x1};\n\t"
        "}"
        : "=l"(r)
        : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
    return r;
}
// a * b for arbitrary representatives: four 32x32->64 multiply-adds (FMA pipe) + special-form reduction.
__device__ __forceinline__ u64 mul(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 p00 = (u64)a0 * b0;
    const u64 p01 = (u64)a0 * b1 + (p00 >> 32);          // <= (2^32-1)^2 + 2^32 - 1 < 2^64
    const u64 p10 = (u64)a1 * b0 + (u32)p01;             // word 1 in the low half
    const u64 p11 = (u64)a1 * b1 + (p01 >> 32) + (p10 >> 32);  // <= 2^64 - 1
    return reduce_words((u32)p00, (u32)p10, (u32)p11, (u32)(p11 >> 32));
}

}  // namespace lazy
}  // namespace gl
#endif
