// Single-compression BLAKE3 for the device.  Every hash on the commitment path is one compression
// with cv = IV, counter = 0, flags = CHUNK_START|CHUNK_END|ROOT (0x0B):
//   * unlabeled leaf  BLAKE3(le8)                       (reference v1/merkle.rs:150-159, fri_stream.rs:37-41)
//   * labeled leaf    BLAKE3("col_leaf"||u32 len||label||le8)   (v1/merkle.rs:132-146)
//   * parent          BLAKE3(left||right), a plain 64-byte message (v1/merkle.rs:57-61,
//                     fri_stream.rs:45-50, sezkp-merkle/src/lib.rs:123-128)
#pragma once
#include <cstdint>

namespace b3 {

typedef uint32_t u32;
typedef uint64_t u64;

#define B3_IV0 0x6A09E667u
#define B3_IV1 0xBB67AE85u
#define B3_IV2 0x3C6EF372u
#define B3_IV3 0xA54FF53Au
#define B3_IV4 0x510E527Fu
#define B3_IV5 0x9B05688Cu
#define B3_IV6 0x1F83D9ABu
#define B3_IV7 0x5BE0CD19u
#define B3_FLAGS_ONE_BLOCK 0x0Bu

__device__ __forceinline__ u32 rotr16(u32 x) { return __byte_perm(x, x, 0x1032); }
__device__ __forceinline__ u32 rotr8(u32 x) { return __byte_perm(x, x, 0x0321); }
__device__ __forceinline__ u32 rotr12(u32 x) { return __funnelshift_r(x, x, 12); }
__device__ __forceinline__ u32 rotr7(u32 x) { return __funnelshift_r(x, x, 7); }

// Pipe balance.  XOR and rotate can only issue on the ALU pipe (LOP3 / SHF / PRMT, 64 lanes/clk/SM), which is the
// binding unit of every hashing kernel here; the FMA pipe (IMAD) is otherwise idle.  Two rewrites move work across:
//  (1) the additions are multiply-adds by a run-time 1 (IMAD): 448 ALU + 336 FMA issue slots per compression instead
//      of 560 + 112;
//  (2) a rotation of the b word can be a widening multiply by a run-time power of two (IMAD.WIDE: the high half is
//      x >> r, the low half x << (32-r)); the two halves are never OR-ed together — the following addition takes
//      both (one more IMAD) and the following XOR is a three-input LOP3 anyway — so each such rotation trades one
//      ALU slot for two FMA slots.  B3_SCHED selects which rotations do this: one hex digit per round, bit 0 / 1 =
//      the 12- / 7-bit rotation of the column step, bit 2 / 3 = of the diagonal step.
__device__ __constant__ u32 k_one = 1;
__device__ __constant__ u32 k_pow20 = 1u << 20;  // rotr 12
__device__ __constant__ u32 k_pow25 = 1u << 25;  // rotr 7
#ifndef B3_ADD_ON_FMA
#define B3_ADD_ON_FMA 1
#endif
#ifndef B3_SCHED
#define B3_SCHED 0x0202000u
#endif
#if B3_ADD_ON_FMA
#define B3_ADD(x, y) ((y) * one + (x))
#else
#define B3_ADD(x, y) ((x) + (y))
#endif
struct Consts {
    u32 one, p20, p25;
};
// b word = b | bl when SI (split in); on return split iff W7.
template <bool SI, bool W12, bool W7>
__device__ __forceinline__ void g_fn(u32& a, u32& b, u32& bl, u32& c, u32& d, u32 mx, u32 my, const Consts& k) {
    const u32 one = k.one;
    (void)one;
    a = B3_ADD(B3_ADD(a, b), mx);
    if (SI) a = B3_ADD(a, bl);
    d = rotr16(d ^ a);
    c = B3_ADD(c, d);
    u32 t = SI ? (b ^ bl ^ c) : (b ^ c);
    u32 h, l = 0;
    if (W12) {
        const u64 w = (u64)t * k.p20;
        h = (u32)(w >> 32);
        l = (u32)w;
    } else h = rotr12(t);
    a = B3_ADD(B3_ADD(a, h), my);
    if (W12) a = B3_ADD(a, l);
    d = rotr8(d ^ a);
    c = B3_ADD(c, d);
    t = W12 ? (h ^ l ^ c) : (h ^ c);
    if (W7) {
        const u64 w = (u64)t * k.p25;
        b = (u32)(w >> 32);
        bl = (u32)w;
    } else b = rotr7(t);
}

// One round; message words are read through BLAKE3's fixed permutation (indices are literals at every call site).
// SI: the b words enter split (previous round's diagonal step used the widening rotation); M: this round's digit.
#define B3_ROUND(SI, M, m, i0, i1, i2, i3, i4, i5, i6, i7, i8, i9, i10, i11, i12, i13, i14, i15)                  \
    g_fn<SI, ((M) & 1) != 0, ((M) & 2) != 0>(s0, s4, l4, s8, s12, m[i0], m[i1], k);                                \
    g_fn<SI, ((M) & 1) != 0, ((M) & 2) != 0>(s1, s5, l5, s9, s13, m[i2], m[i3], k);                                \
    g_fn<SI, ((M) & 1) != 0, ((M) & 2) != 0>(s2, s6, l6, s10, s14, m[i4], m[i5], k);                               \
    g_fn<SI, ((M) & 1) != 0, ((M) & 2) != 0>(s3, s7, l7, s11, s15, m[i6], m[i7], k);                               \
    g_fn<((M) & 2) != 0, ((M) & 4) != 0, ((M) & 8) != 0>(s0, s5, l5, s10, s15, m[i8], m[i9], k);                   \
    g_fn<((M) & 2) != 0, ((M) & 4) != 0, ((M) & 8) != 0>(s1, s6, l6, s11, s12, m[i10], m[i11], k);                 \
    g_fn<((M) & 2) != 0, ((M) & 4) != 0, ((M) & 8) != 0>(s2, s7, l7, s8, s13, m[i12], m[i13], k);                  \
    g_fn<((M) & 2) != 0, ((M) & 4) != 0, ((M) & 8) != 0>(s3, s4, l4, s9, s14, m[i14], m[i15], k);

// Compile-time G with zero message words: the three column steps of an unlabeled leaf's first round that do not touch
// the value (columns 1-3: state words are IV / counter / block_len / flags, message words m2..m7 are zero) are constants.
struct GConst {
    u32 a, b, c, d;
};
constexpr u32 c_rotr(u32 x, int r) { return (x >> r) | (x << (32 - r)); }
constexpr GConst g_const(u32 a, u32 b, u32 c, u32 d) {
    a = a + b;
    d = c_rotr(d ^ a, 16);
    c = c + d;
    b = c_rotr(b ^ c, 12);
    a = a + b;
    d = c_rotr(d ^ a, 8);
    c = c + d;
    b = c_rotr(b ^ c, 7);
    return GConst{a, b, c, d};
}

// G with message words on the host / at compile time (label templates precompute the value-independent part of round 0).
constexpr void g_plain(u32& a, u32& b, u32& c, u32& d, u32 mx, u32 my) {
    a = a + b + mx;
    d = c_rotr(d ^ a, 16);
    c = c + d;
    b = c_rotr(b ^ c, 12);
    a = a + b + my;
    d = c_rotr(d ^ a, 8);
    c = c + d;
    b = c_rotr(b ^ c, 7);
}

// One-block hash: out[8] = BLAKE3(message of block_len bytes held zero-padded in m[16]).  MODE selects how much of round 0
// is already known:
//   0  nothing;
//   1  unlabeled 8-byte leaf (m[2..15] = 0, block_len = 8): round 0 runs one column step instead of four, the other three
//      are folded at compile time (the run-time-1 multiplier of B3_ADD would otherwise hide them from constant propagation);
//   2  labeled leaf: column step 0 (message words m0, m1 = "col_leaf", state words IV) comes precomputed in pre[0..3];
//   3  labeled leaf whose value starts at word >= 4: column step 1 (m2 = label length, m3 = label bytes) in pre[4..7] too.
template <u32 SCHED = B3_SCHED, int MODE = 0>
__device__ __forceinline__ void hash_block(const u32 (&m)[16], u32 block_len, u32 (&out)[8], const u32* pre = nullptr) {
    constexpr bool LEAF8 = MODE == 1;
    u32 s0 = B3_IV0, s1 = B3_IV1, s2 = B3_IV2, s3 = B3_IV3, s4 = B3_IV4, s5 = B3_IV5, s6 = B3_IV6, s7 = B3_IV7;
    u32 s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3, s12 = 0, s13 = 0, s14 = block_len, s15 = B3_FLAGS_ONE_BLOCK;
    u32 l4 = 0, l5 = 0, l6 = 0, l7 = 0;
    const Consts k = {k_one, k_pow20, k_pow25};
    constexpr u32 M0 = SCHED & 15, M1 = (SCHED >> 4) & 15, M2 = (SCHED >> 8) & 15, M3 = (SCHED >> 12) & 15, M4 = (SCHED >> 16) & 15,
                  M5 = (SCHED >> 20) & 15, M6 = (SCHED >> 24) & 15;
    if (LEAF8 && (M0 & 3) == 0) {
        constexpr GConst c1 = g_const(B3_IV1, B3_IV5, B3_IV1, 0), c2 = g_const(B3_IV2, B3_IV6, B3_IV2, 8),
                         c3 = g_const(B3_IV3, B3_IV7, B3_IV3, B3_FLAGS_ONE_BLOCK);
        g_fn<false, false, false>(s0, s4, l4, s8, s12, m[0], m[1], k);
        s1 = c1.a; s5 = c1.b; s9 = c1.c; s13 = c1.d;
        s2 = c2.a; s6 = c2.b; s10 = c2.c; s14 = c2.d;
        s3 = c3.a; s7 = c3.b; s11 = c3.c; s15 = c3.d;
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s0, s5, l5, s10, s15, m[8], m[9], k);
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s1, s6, l6, s11, s12, m[10], m[11], k);
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s2, s7, l7, s8, s13, m[12], m[13], k);
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s3, s4, l4, s9, s14, m[14], m[15], k);
    } else if (MODE >= 2 && (M0 & 3) == 0) {
        s0 = pre[0]; s4 = pre[1]; s8 = pre[2]; s12 = pre[3];
        if (MODE == 3) {
            s1 = pre[4]; s5 = pre[5]; s9 = pre[6]; s13 = pre[7];
        } else {
            g_fn<false, false, false>(s1, s5, l5, s9, s13, m[2], m[3], k);
        }
        g_fn<false, false, false>(s2, s6, l6, s10, s14, m[4], m[5], k);
        g_fn<false, false, false>(s3, s7, l7, s11, s15, m[6], m[7], k);
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s0, s5, l5, s10, s15, m[8], m[9], k);
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s1, s6, l6, s11, s12, m[10], m[11], k);
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s2, s7, l7, s8, s13, m[12], m[13], k);
        g_fn<false, (M0 & 4) != 0, (M0 & 8) != 0>(s3, s4, l4, s9, s14, m[14], m[15], k);
    } else {
        B3_ROUND(false, M0, m, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    }
    B3_ROUND((M0 & 8) != 0, M1, m, 2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    B3_ROUND((M1 & 8) != 0, M2, m, 3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    B3_ROUND((M2 & 8) != 0, M3, m, 10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    B3_ROUND((M3 & 8) != 0, M4, m, 12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    B3_ROUND((M4 & 8) != 0, M5, m, 9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    B3_ROUND((M5 & 8) != 0, M6, m, 11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    if ((M6 & 8) != 0) {  // b words still split: fold the halves into the output XOR
        s4 ^= l4; s5 ^= l5; s6 ^= l6; s7 ^= l7;
    }
    out[0] = s0 ^ s8;  out[1] = s1 ^ s9;  out[2] = s2 ^ s10; out[3] = s3 ^ s11;
    out[4] = s4 ^ s12; out[5] = s5 ^ s13; out[6] = s6 ^ s14; out[7] = s7 ^ s15;
}

// Unlabeled leaf of one canonical field element.
__device__ __forceinline__ void leaf(u64 v, u32 (&out)[8]) {
    u32 m[16] = {(u32)v, (u32)(v >> 32), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    hash_block<B3_SCHED, 1>(m, 8, out);
}
// Parent of two digests.
__device__ __forceinline__ void parent(const u32 (&l)[8], const u32 (&r)[8], u32 (&out)[8]) {
    u32 m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        m[i] = l[i];
        m[8 + i] = r[i];
    }
    hash_block(m, 64, out);
}

// Labeled-leaf message template: the 12+L prefix bytes laid out in 16 words; the 8 value bytes
// are inserted at byte offset `off` = 12+L (unaligned in general).
struct LabelTemplate {
    u32 words[16];
    u32 off;        // byte offset of the value
    u32 block_len;  // 20 + L
    u32 pre[8];     // round 0 after the value-independent column steps: {s0,s4,s8,s12} after G(m0,m1) and {s1,s5,s9,s13}
                    // after G(m2,m3); the second quadruple is only used when the value starts at word >= 4 (off >= 16)
};
// fills t.pre from t.words (host side, when the template is built)
inline void label_template_precompute(LabelTemplate& t) {
    u32 a = B3_IV0, b = B3_IV4, c = B3_IV0, d = 0;
    g_plain(a, b, c, d, t.words[0], t.words[1]);
    t.pre[0] = a; t.pre[1] = b; t.pre[2] = c; t.pre[3] = d;
    a = B3_IV1; b = B3_IV5; c = B3_IV1; d = 0;
    g_plain(a, b, c, d, t.words[2], t.words[3]);
    t.pre[4] = a; t.pre[5] = b; t.pre[6] = c; t.pre[7] = d;
}
// W = word index of the first value byte (static so that the 16 message words stay in registers).
template <int W>
__device__ __forceinline__ void leaf_labeled_w(const LabelTemplate& t, u64 v, u32 (&out)[8]) {
    u32 m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = t.words[i];
    const u32 sh = (t.off & 3) * 8;
    const u32 lo = (u32)v, hi = (u32)(v >> 32);
    m[W] |= lo << sh;
    if (W + 1 < 16) m[W + 1] |= __funnelshift_l(lo, hi, sh);
    if (W + 2 < 16) m[W + 2] |= __funnelshift_l(hi, 0u, sh);
    hash_block<B3_SCHED, (W >= 4 ? 3 : 2)>(m, t.block_len, out, t.pre);
}
// Runs BODY with a compile-time W matching the (block-uniform) template; BODY uses LEAF(v, out).
#define B3_DISPATCH_LABELED(t, BODY)                                       \
    switch ((t).off >> 2) {                                                \
        case 3: { constexpr int B3W = 3; BODY } break;                     \
        case 4: { constexpr int B3W = 4; BODY } break;                     \
        case 5: { constexpr int B3W = 5; BODY } break;                     \
        case 6: { constexpr int B3W = 6; BODY } break;                     \
        case 7: { constexpr int B3W = 7; BODY } break;                     \
        case 8: { constexpr int B3W = 8; BODY } break;                     \
        case 9: { constexpr int B3W = 9; BODY } break;                     \
        case 10: { constexpr int B3W = 10; BODY } break;                   \
        case 11: { constexpr int B3W = 11; BODY } break;                   \
        case 12: { constexpr int B3W = 12; BODY } break;                   \
        case 13: { constexpr int B3W = 13; BODY } break;                   \
        default: { constexpr int B3W = 14; BODY } break;                   \
    }

}  // namespace b3
