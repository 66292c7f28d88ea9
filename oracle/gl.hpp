// ORACLE — TEST INFRASTRUCTURE ONLY. Not linked into, imported by, or shipped with the product
// (libsezkp_cuda.so). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// arm may use anything under oracle/.
//
// Goldilocks field, restating the reference's `Fp64<P>` semantics
// (crates/sezkp-ffts/src/lib.rs:34-133, GOLDILOCKS :229, primitive root :237-242):
// canonical residues in [0,p), every operation fully reduced, mul = (u128 product) % p.
#pragma once
#include <cstdint>
#include <cstring>

namespace oracle {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i64 = int64_t;
using u128 = unsigned __int128;
using i128 = __int128;

constexpr u64 GL_P = 0xffffffff00000001ULL;  // lib.rs:229

// add_raw lib.rs:57-62
inline u64 gl_add(u64 a, u64 b) {
    u128 s = (u128)a + (u128)b;
    if (s >= (u128)GL_P) s -= (u128)GL_P;
    return (u64)s;
}
// sub_raw lib.rs:66-73
inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : (u64)((u128)a + (u128)GL_P - (u128)b); }
// mul_raw lib.rs:78-81 — deliberately the reference's u128 % p (this is what the CPU baseline times)
inline u64 gl_mul(u64 a, u64 b) { return (u64)(((u128)a * (u128)b) % (u128)GL_P); }
// pow lib.rs:86-97
inline u64 gl_pow(u64 base, u64 e) {
    u64 acc = 1;
    while (e > 0) {
        if (e & 1) acc = gl_mul(acc, base);
        base = gl_mul(base, base);
        e >>= 1;
    }
    return acc;
}
// inv lib.rs:102-104 (Fermat)
inline u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }
// from_i64 lib.rs:109-111 (rem_euclid)
inline u64 gl_from_i64(i64 x) {
    i128 r = (i128)x % (i128)GL_P;
    if (r < 0) r += (i128)GL_P;
    return (u64)r;
}
// from_u64 lib.rs:116-118
inline u64 gl_from_u64(u64 x) { return x % GL_P; }
// neg lib.rs:130-132
inline u64 gl_neg(u64 a) { return a == 0 ? 0 : GL_P - a; }
// goldilocks_primitive_root_2exp lib.rs:237-242 : 7^((p-1)>>k)
inline u64 gl_root_2exp(unsigned k) { return gl_pow(7, (GL_P - 1) >> k); }

inline void gl_to_le(u64 v, u8 out[8]) { std::memcpy(out, &v, 8); }  // x86 is little-endian
inline u64 gl_from_le(const u8 in[8]) { u64 v; std::memcpy(&v, in, 8); return v; }

}  // namespace oracle
