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
// a * b for arbitrary representatives (first-generation form: explicit partial products + reduce_words, 28 SASS
// instructions).  Kept for tools/arith_test.cu; `mul` below is what the kernels call.
__device__ __forceinline__ u64 mul_v1(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 p00 = (u64)a0 * b0;
    const u64 p01 = (u64)a0 * b1 + (p00 >> 32);          // <= (2^32-1)^2 + 2^32 - 1 < 2^64
    const u64 p10 = (u64)a1 * b0 + (u32)p01;             // word 1 in the low half
    const u64 p11 = (u64)a1 * b1 + (p01 >> 32) + (p10 >> 32);  // <= 2^64 - 1
    return reduce_words((u32)p00, (u32)p10, (u32)p11, (u32)(p11 >> 32));
}
__device__ __forceinline__ u64 mulc(u64 a, u64 b, u32 eps);
static __device__ __constant__ u32 k_eps32 = 0xffffffffu;  // multiplier from the constant bank: keeps ptxas from
                                                           // strength-reducing w*EPS into an ALU-pipe shift/sub pair
// a * b for arbitrary representatives -> canonical residue (a valid representative for add/sub above): the
// second-generation product below, 19 SASS instructions.
__device__ __forceinline__ u64 mul(u64 a, u64 b) { return mulc(a, b, k_eps32); }

// ---------------------------------------------------------------------------------------------
// Second-generation primitives for the NTT (csrc/ntt.cu).  The first set above keeps every value an
// arbitrary 64-bit representative, which costs two masked corrections per add/sub.  Here products are
// reduced to the canonical residue (< p) and sums / differences are single-correction:
//   add1(a, b): exact when at least one operand is <= p      (a + b - 2^64 + EPS cannot wrap again)
//   sub1(a, b): exact when the subtrahend b is <= p
// A decimation-in-time butterfly (a + w*b, a - w*b) always has the freshly reduced product as one
// operand, so `a` may stay lazy through all stages.  ptxas turns the mul.lo/mul.hi pair into four
// IMAD.WIDE with carry predicates (7 instructions for the 128-bit product).
// ---------------------------------------------------------------------------------------------
// A + w*EPS - B (mod p): canonical result (< p) for ANY 64-bit A and 32-bit w, B.
//   U = A - B (borrow k);  V = U + (w + 1 - k)*EPS (carry C) = T + EPS with T = A - B + w*EPS + k*p in [0, 2p)
//   C ? V - 2^64 (= T - p) : V - EPS (= T)
__device__ __forceinline__ u64 red3(u64 A, u32 w, u32 B, u32 eps) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0,x1,m,y0,y1;\n\t"
        ".reg .u64 Y;\n\t"
        "mov.b64 {x0,x1}, %1;\n\t"
        "sub.cc.u32 x0, x0, %3;\n\t"
        "subc.cc.u32 x1, x1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"        // borrow ? ~0 : 0
        "not.b32 m, m;\n\t"            // (1 - k)*EPS
        "mov.b64 Y, {m, 0};\n\t"
        "mad.wide.u32 Y, %2, %4, Y;\n\t"
        "mov.b64 {y0,y1}, Y;\n\t"
        "add.cc.u32 x0, x0, y0;\n\t"
        "addc.cc.u32 x1, x1, y1;\n\t"
        "addc.u32 m, 0xffffffff, 0;\n\t"  // carry ? 0 : ~0
        "sub.cc.u32 x0, x0, m;\n\t"
        "subc.u32 x1, x1, 0;\n\t"
        "mov.b64 %0, {x0,x1};\n\t"
        "}"
        : "=l"(r)
        : "l"(A), "r"(w), "r"(B), "r"(eps));
    return r;
}
// A + w*EPS (mod p), canonical, any 64-bit A.
__device__ __forceinline__ u64 red2(u64 A, u32 w, u32 eps) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0,x1,m,y0,y1;\n\t"
        ".reg .u64 Y;\n\t"
        "mov.b64 {x0,x1}, %1;\n\t"
        "mad.wide.u32 Y, %2, %3, 0xffffffff;\n\t"
        "mov.b64 {y0,y1}, Y;\n\t"
        "add.cc.u32 x0, x0, y0;\n\t"
        "addc.cc.u32 x1, x1, y1;\n\t"
        "addc.u32 m, 0xffffffff, 0;\n\t"
        "sub.cc.u32 x0, x0, m;\n\t"
        "subc.u32 x1, x1, 0;\n\t"
        "mov.b64 %0, {x0,x1};\n\t"
        "}"
        : "=l"(r)
        : "l"(A), "r"(w), "r"(eps));
    return r;
}
// w*EPS - S (mod p) for S < 2^63, canonical (w*EPS <= 2^64 - 2^33 + 1 < p, so one conditional +p suffices).
__device__ __forceinline__ u64 red_neg(u32 w, u64 S, u32 eps) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0,x1,s0,s1,m;\n\t"
        ".reg .u64 Y;\n\t"
        "mov.b64 {s0,s1}, %2;\n\t"
        "mul.wide.u32 Y, %1, %3;\n\t"
        "mov.b64 {x0,x1}, Y;\n\t"
        "sub.cc.u32 x0, x0, s0;\n\t"
        "subc.cc.u32 x1, x1, s1;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 x0, x0, m;\n\t"
        "subc.u32 x1, x1, 0;\n\t"
        "mov.b64 %0, {x0,x1};\n\t"
        "}"
        : "=l"(r)
        : "r"(w), "l"(S), "r"(eps));
    return r;
}
// a * b -> canonical residue, any 64-bit representatives in.
__device__ __forceinline__ u64 mulc(u64 a, u64 b, u32 eps) {
    u64 lo, hi;
    asm("mul.lo.u64 %0, %2, %3;\n\tmul.hi.u64 %1, %2, %3;" : "=l"(lo), "=l"(hi) : "l"(a), "l"(b));
    return red3(lo, (u32)hi, (u32)(hi >> 32), eps);
}
__device__ __forceinline__ u64 add1(u64 a, u64 b) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 a0,a1,b0,b1,m;\n\t"
        "mov.b64 {a0,a1}, %1;\n\t"
        "mov.b64 {b0,b1}, %2;\n\t"
        "add.cc.u32 a0, a0, b0;\n\t"
        "addc.cc.u32 a1, a1, b1;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 a0, a0, m;\n\t"       // + carry*EPS = - carry + (carry << 32)
        "subc.u32 m, m, 0;\n\t"
        "add.u32 a1, a1, m;\n\t"
        "mov.b64 %0, {a0,a1};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub1(u64 a, u64 b) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 a0,a1,b0,b1,m;\n\t"
        "mov.b64 {a0,a1}, %1;\n\t"
        "mov.b64 {b0,b1}, %2;\n\t"
        "sub.cc.u32 a0, a0, b0;\n\t"
        "subc.cc.u32 a1, a1, b1;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 a0, a0, m;\n\t"
        "subc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0,a1};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b));
    return r;
}
// any 64-bit representative -> canonical residue
__device__ __forceinline__ u64 canon2(u64 x) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0,x1,t0,t1,c;\n\t"
        ".reg .pred q;\n\t"
        "mov.b64 {x0,x1}, %1;\n\t"
        "add.cc.u32 t0, x0, 0xffffffff;\n\t"
        "addc.cc.u32 t1, x1, 0;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "setp.ne.u32 q, c, 0;\n\t"
        "selp.u32 x0, t0, x0, q;\n\t"
        "selp.u32 x1, t1, x1, q;\n\t"
        "mov.b64 %0, {x0,x1};\n\t"
        "}"
        : "=l"(r)
        : "l"(x));
    return r;
}
// x * 2^e (mod p), 0 < e < 96, canonical result, any 64-bit x; e is a literal after unrolling.
//   x << (e & 31) = (y2, y1, y0);  word offset q = e >> 5 and 2^64 = EPS, 2^96 = -1, 2^128 = -2^32:
//   q = 0: (y1:y0) + y2*EPS      q = 1: (y0:0) + y1*EPS - y2      q = 2: y0*EPS - (y2:y1)
__device__ __forceinline__ u64 mul_pow2(u64 x, int e, u32 eps) {
    const int q = e >> 5, r = e & 31;
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32);
    u32 y0, y1, y2;
    if (r == 0) {
        y0 = x0;
        y1 = x1;
        y2 = 0;
    } else {
        y0 = x0 << r;
        y1 = __funnelshift_l(x0, x1, r);
        y2 = x1 >> (32 - r);
    }
    if (q == 0) return red2(((u64)y1 << 32) | y0, y2, eps);
    if (q == 1) return red3((u64)y0 << 32, y1, y2, eps);
    return red_neg(y0, ((u64)y2 << 32) | y1, eps);
}
constexpr __host__ __device__ int brev_bits(int x, int k) {
    int r = 0;
    for (int i = 0; i < k; i++) r |= ((x >> i) & 1) << (k - 1 - i);
    return r;
}
// 2^K-point DFT (K <= 5) in registers with the reference's root: w_32 = 7^((p-1)/32) = 2^78, so every twiddle of a
// transform of up to 32 points is a power of two (2^192 = 1, 2^96 = -1) — shifts instead of multiplications.
// INV uses the inverse root.  Radix-2 decimation in time; in: v[t] <= p in natural order; out: X[k] in
// v[brev_bits(k, K)] as arbitrary 64-bit representatives.
template <int K, bool INV>
__device__ __forceinline__ void dft_pow2(u64 (&v)[1 << K], u32 eps) {
#pragma unroll
    for (int s = 0; s < K; s++) {
        const int half = 1 << s;
#pragma unroll
        for (int i = 0; i < (1 << K); i++) {
            if (i & half) continue;
            const int j = i & (half - 1);
            int E = (78 * (16 >> s) * j) % 192;  // w_{2^(s+1)}^j = 2^E
            if (INV) E = (192 - E) % 192;
            const int pa = brev_bits(i, K), pb = brev_bits(i + half, K);
            const u64 a = v[pa], b = v[pb];
            const int e = E % 96;
            u64 m;
            if (e == 0) m = (s == 0) ? b : canon2(b);
            else m = mul_pow2(b, e, eps);
            if (E < 96) {
                v[pa] = add1(a, m);
                v[pb] = sub1(a, m);
            } else {
                v[pa] = sub1(a, m);
                v[pb] = add1(a, m);
            }
        }
    }
}

}  // namespace lazy
}  // namespace gl
#endif
