// BLAKE3 leaf hashing and chunked Merkle commitments (sm_100a).
// Replaces hash_field_leaves{,_labeled} (reference v1/merkle.rs:132-159), MerkleTree::from_leaves / open
// (v1/merkle.rs:46-108), OnDemandOpenings::{build_roots, open} (v1/openings.rs:306-497) and
// StreamingLayerBuilder (v1/fri_stream.rs:55-122).  For a power-of-two number of leaves the chunked
// tree (2^10-leaf chunk trees under an outer tree of chunk roots) is the plain binary tree, so one
// kernel reduces a whole chunk in shared memory and only chunk roots and the levels above are stored.
#include <algorithm>
#include <cstring>

#include "hash.cuh"
#include "gl.cuh"

namespace {

constexpr int HASH_THREADS = 256;
constexpr int MAX_CL = 10;

// Shared-memory digest array in word-major layout: word w of digest i at s[w*pitch + i]; a thread reads
// (left.w, right.w) of pair i as one 64-bit load at index 2i.
__device__ __forceinline__ void pair_parent(const u32* s, int pitch, int i, u32 (&out)[8]) {
    u32 l[8], r[8];
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint2 p = *reinterpret_cast<const uint2*>(&s[w * pitch + 2 * i]);
        l[w] = p.x;
        r[w] = p.y;
    }
    b3::parent(l, r, out);
}
__device__ __forceinline__ void put_digest(u32* s, int pitch, int i, const u32 (&d)[8], u32* gdst) {
#pragma unroll
    for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
    if (gdst) {
        uint4* d4 = reinterpret_cast<uint4*>(gdst);
        d4[0] = make_uint4(d[0], d[1], d[2], d[3]);
        d4[1] = make_uint4(d[4], d[5], d[6], d[7]);
    }
}
// Reduces `count` (power of two, <= 1024) digests held in s to one (left in slot 0).  Optionally stores every
// produced level into the retained upper-tree array (`levels_base`, level l0+1.. of a column with n_ch_total
// chunk roots, this group being number `grp` at level l0) and/or records the sibling path of leaf `path_idx`.
__device__ __forceinline__ void reduce_levels_smem(u32* s, int pitch, int count, u32* levels_base, u64 grp, int l0,
                                                   u64 n_ch_total, u32* path_out, int path_idx, int stop = 1) {
    int lvl = 0;
    while (count > stop) {
        if (path_out != nullptr && threadIdx.x < 8) {
            const int sib = (path_idx >> lvl) ^ 1;
            path_out[lvl * 8 + threadIdx.x] = s[threadIdx.x * pitch + sib];
        }
        const int half = count >> 1;
        const int i0 = threadIdx.x, i1 = threadIdx.x + HASH_THREADS;  // half <= 512 = 2*HASH_THREADS
        u32 d0[8], d1[8];
        if (i0 < half) pair_parent(s, pitch, i0, d0);
        if (i1 < half) pair_parent(s, pitch, i1, d1);
        __syncthreads();
        u32* gl = nullptr;
        if (levels_base) {
            const int l = l0 + lvl + 1;
            gl = levels_base + ((2 * n_ch_total - ((2 * n_ch_total) >> l)) + grp * (u64)half) * 8;
        }
        if (i0 < half) put_digest(s, pitch, i0, d0, gl ? gl + (u64)i0 * 8 : nullptr);
        if (i1 < half) put_digest(s, pitch, i1, d1, gl ? gl + (u64)i1 * 8 : nullptr);
        __syncthreads();
        count = half;
        lvl++;
    }
}

// Leaf hashing of a chunk into the word-major digest array (block-uniform label template or unlabeled).
#define HASH_LEAVES_INTO(s, pitch, leaves, VALUE_EXPR, templates_nonnull, t)          \
    if (templates_nonnull) {                                                          \
        B3_DISPATCH_LABELED(t, {                                                      \
            for (int i = threadIdx.x; i < (leaves); i += HASH_THREADS) {              \
                u32 d[8];                                                             \
                b3::leaf_labeled_w<B3W>(t, (VALUE_EXPR), d);                          \
                _Pragma("unroll") for (int w = 0; w < 8; w++) s[w * (pitch) + i] = d[w]; \
            }                                                                         \
        })                                                                            \
    } else {                                                                          \
        for (int i = threadIdx.x; i < (leaves); i += HASH_THREADS) {                  \
            u32 d[8];                                                                 \
            b3::leaf((VALUE_EXPR), d);                                                \
            _Pragma("unroll") for (int w = 0; w < 8; w++) s[w * (pitch) + i] = d[w];  \
        }                                                                             \
    }

// values -> leaf hashes -> chunk root.  grid (n_ch, cols).  Writes chunk roots into level 0 of `upper`.
// FOLD: the values are produced on the fly as the FRI fold of the previous layer,
//   y'[i] = y[i] + beta*y[i+half]  (reference v1/prover.rs:204-238), written to `values` and hashed in one pass.
template <bool FOLD>
__global__ void __launch_bounds__(HASH_THREADS, 5) chunk_commit_kernel(u64* __restrict__ values, u64 n, u64 col_stride, int cl, int sub,
                                                                    const b3::LabelTemplate* __restrict__ templates,
                                                                    u32* __restrict__ upper, u64 n_ch,
                                                                    const u64* __restrict__ fold_src, u64 beta, u64 chunk0) {
    // one CTA reduces 2^(cl+sub) leaves = 2^sub chunks of 2^cl leaves and writes their 2^sub roots (sub = 0: one chunk)
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    const u64 chunk = chunk0 + ((u64)blockIdx.x << sub);
    const int col = blockIdx.y;
    const int leaves = 1 << (cl + sub);
    u64* v = values + (u64)col * col_stride + (chunk << cl);
    b3::LabelTemplate t;
    if (templates) t = templates[col];
    if (FOLD) {
        const u64* lo = fold_src + (chunk << cl);
        for (int i = threadIdx.x; i < leaves; i += HASH_THREADS) {
            const u64 y = gl::add(lo[i], gl::mul(beta, lo[i + n]));
            v[i] = y;
            u32 d[8];
            b3::leaf(y, d);
#pragma unroll
            for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
        }
    } else {
        HASH_LEAVES_INTO(s, pitch, leaves, v[i], templates != nullptr, t)
    }
    __syncthreads();
    const int roots = 1 << sub;
    reduce_levels_smem(s, pitch, leaves, nullptr, 0, 0, 0, nullptr, 0, roots);
    u32* dst = upper + ((u64)col * (2 * n_ch - 1) + chunk) * 8;
    for (int i = threadIdx.x; i < 8 * roots; i += HASH_THREADS) dst[i] = s[(i & 7) * pitch + (i >> 3)];
}

// Unlabeled chunk commit of several single-column commitments in one launch (the small FRI layers, whose values are
// folded beforehand: hashing a layer does not feed the next fold, so all of them can be hashed side by side).
// Job j owns CTAs [cta0[j], cta0[j+1]); one CTA per chunk of 2^cl leaves.
__global__ void __launch_bounds__(HASH_THREADS, 5) chunk_commit_multi_kernel(const CommitJobs jobs) {
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    int j = 0;
    while (j + 1 < jobs.n && blockIdx.x >= jobs.cta0[j + 1]) j++;
    const CommitJob jb = jobs.j[j];
    const u64 chunk = blockIdx.x - jobs.cta0[j];
    const int leaves = 1 << jb.cl;
    const u64* v = jb.values + (chunk << jb.cl);
    for (int i = threadIdx.x; i < leaves; i += HASH_THREADS) {
        u32 d[8];
        b3::leaf(v[i], d);
#pragma unroll
        for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
    }
    __syncthreads();
    reduce_levels_smem(s, pitch, leaves, nullptr, 0, 0, 0, nullptr, 0);
    if (threadIdx.x < 8) jb.upper[chunk * 8 + threadIdx.x] = s[threadIdx.x * pitch];
}

/* ------------------------------------------------------------------------------------------ */
/* value-aware chunk commit: identical leaves / identical sibling pairs are hashed once         */
/* ------------------------------------------------------------------------------------------ */
// Trace columns take few distinct values per 1024-row chunk (flags, moves in {-1,0,1}, 4-bit symbols, block
// constants, a slowly moving head).  BLAKE3 is a function: equal inputs give equal digests, so the kernel assigns
// every node a small id per distinct digest, hashes each distinct leaf value and each distinct (left id, right id)
// pair once, and falls back to the plain reduction as soon as ids stop repeating.  Outputs are bit-identical to
// the plain kernel; only redundant compressions are skipped.
constexpr int DD_PA = (1 << MAX_CL) + 2;       // pitch of the digest array (1024 entries); tables live in its two halves
constexpr int DD_HALF = 1 << (MAX_CL - 1);     // 512 entries per half
constexpr int DD_PAIR_CAP = 2048;              // dedup a level while D*D <= cap
struct DedupSmem {
    u32 A[8 * DD_PA];
    u32 bitmap[DD_PAIR_CAP / 32];
    u32 wprefix[DD_PAIR_CAP / 32];
    unsigned short idA[1 << MAX_CL];
    unsigned short idB[1 << (MAX_CL - 1)];
    unsigned short list[DD_PAIR_CAP];
    u64 red[2 * (HASH_THREADS / 32)];
    u32 total;
};

// exclusive prefix of popcounts over `words` bitmap words (<= 64); leaves the total in sm.total
__device__ __forceinline__ void bitmap_prefix(DedupSmem& sm, int words) {
    if (threadIdx.x < 32) {
        const int per = (words + 31) >> 5;  // <= 2
        u32 cnt[2], sum = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = threadIdx.x * per + j;
            cnt[j] = (j < per && w < words) ? __popc(sm.bitmap[w]) : 0;
            sum += cnt[j];
        }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += y;
        }
        u32 run = incl - sum;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = threadIdx.x * per + j;
            if (j < per && w < words) sm.wprefix[w] = run;
            run += cnt[j];
        }
        if (threadIdx.x == 31) sm.total = incl;
    }
}
__device__ __forceinline__ u32 bitmap_rank(const DedupSmem& sm, u32 key) {
    return sm.wprefix[key >> 5] + __popc(sm.bitmap[key >> 5] & ((1u << (key & 31)) - 1u));
}
// Set bit `key` for every valid lane.  Returns true for exactly one lane per key that was not yet set ("owner"), which
// later writes the key into the compacted list — no scan over the key space.  Small key spaces (one bitmap word) are
// aggregated into one atomic per warp (all lanes of a low-entropy chunk hit the same word; unaggregated atomics
// would serialise); larger spaces use one atomic per lane.  Whole warp must call.
__device__ __forceinline__ bool bitmap_set_owner(u32* bitmap, u32 key, bool valid, bool single_word) {
    const u32 bit = 1u << (key & 31);
    if (single_word) {
        const u32 bits = __reduce_or_sync(0xffffffffu, valid ? bit : 0u);
        u32 old = 0;
        if ((threadIdx.x & 31) == 0 && bits) old = atomicOr(&bitmap[0], bits);
        old = __shfl_sync(0xffffffffu, old, 0);
        // first lane (lowest id) carrying a newly set bit owns it
        const unsigned same = __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
        return valid && !(old & bit) && ((int)(threadIdx.x & 31) == __ffs(same) - 1);
    }
    if (!valid) return false;
    const u32 old = atomicOr(&bitmap[key >> 5], bit);
    return !(old & bit);
}

__device__ unsigned long long g_memo_stats[4];  // debug counters: hits, misses, claimed, chains (SEZKP_MEMO_STATS builds only)
#ifdef SEZKP_MEMO_STATS
#define MEMO_STAT(i) atomicAdd(&g_memo_stats[i], 1ULL)
#else
#define MEMO_STAT(i)
#endif
constexpr int MEMO_SLOTS = 16384;  // power of two
constexpr int MEMO_WORDS = 18;    // state, height, x[8], root[8]
__device__ __forceinline__ u32 ld_acquire_u32(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(u32* p, u32 v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(HASH_THREADS, 5) chunk_commit_dedup_kernel(const u64* __restrict__ values, u64 n, u64 col_stride, int cl,
                                                                          const b3::LabelTemplate* __restrict__ templates,
                                                                          u32* __restrict__ upper, u64 n_ch, u32* memo, u64 chunk0) {
    extern __shared__ __align__(16) unsigned char dd_raw[];
    DedupSmem& sm = *reinterpret_cast<DedupSmem*>(dd_raw);
    const u64 chunk = chunk0 + blockIdx.x;
    const int col = blockIdx.y;
    const int leaves = 1 << cl;
    const u64* v = values + (u64)col * col_stride + (chunk << cl);
    const int tid = threadIdx.x;
    b3::LabelTemplate t;
    if (templates) t = templates[col];
    u32* out_root = upper + ((u64)col * (2 * n_ch - 1) + chunk) * 8;

    // ---- keys: value + 2^31 (mod p) so that small negative residues sit next to small positive ones ----
    constexpr u64 HALF = 1ULL << 31;
    u64 key[4];
    u64 mn = ~0ULL, mx = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = tid + k * HASH_THREADS;
        key[k] = 0;
        if (i < leaves) {
            key[k] = gl::add(v[i], HALF);
            mn = key[k] < mn ? key[k] : mn;
            mx = key[k] > mx ? key[k] : mx;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((tid & 31) == 0) {
        sm.red[tid >> 5] = mn;
        sm.red[HASH_THREADS / 32 + (tid >> 5)] = mx;
    }
    if (tid < DD_PAIR_CAP / 32) sm.bitmap[tid] = 0;
    __syncthreads();
    mn = sm.red[0];
    mx = sm.red[HASH_THREADS / 32];
#pragma unroll
    for (int w = 1; w < HASH_THREADS / 32; w++) {
        mn = sm.red[w] < mn ? sm.red[w] : mn;
        mx = sm.red[HASH_THREADS / 32 + w] > mx ? sm.red[HASH_THREADS / 32 + w] : mx;
    }
    int D = 0;
    bool dedup = (mx - mn) < (u64)(1 << MAX_CL);
    bool own[4] = {false, false, false, false};
    if (dedup) {  // ids of the distinct values
        const bool one_word = (mx - mn) < 32;
#pragma unroll
        for (int k = 0; k < 4; k++) own[k] = bitmap_set_owner(sm.bitmap, (u32)(key[k] - mn), tid + k * HASH_THREADS < leaves, one_word);
        __syncthreads();
        bitmap_prefix(sm, ((int)(mx - mn) + 32) >> 5);
        __syncthreads();
        D = (int)sm.total;
        dedup = D <= DD_HALF;
    }
    if (!dedup) {  // values too spread out / too many distinct values: plain path
        HASH_LEAVES_INTO(sm.A, DD_PA, leaves, v[i], templates != nullptr, t)
        __syncthreads();
        reduce_levels_smem(sm.A, DD_PA, leaves, nullptr, 0, 0, 0, nullptr, 0);
        if (tid < 8) out_root[tid] = sm.A[tid * DD_PA];
        return;
    }

    // ---- level 0: one leaf hash per distinct value (table in the low half of A) ----
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = tid + k * HASH_THREADS;
        if (i < leaves) {
            const u32 kk = (u32)(key[k] - mn), rk = bitmap_rank(sm, kk);
            sm.idA[i] = (unsigned short)rk;
            if (own[k]) sm.list[rk] = (unsigned short)kk;
        }
    }
    __syncthreads();
    HASH_LEAVES_INTO(sm.A, DD_PA, D, gl::sub(mn + (u64)sm.list[i], HALF), templates != nullptr, t)
    __syncthreads();

    // ---- levels: dedup while the pair space is small, then one indirect level into the other half, then plain ----
    u32* Tcur = sm.A;
    u32* Tnext = sm.A + DD_HALF;
    unsigned short* idcur = sm.idA;
    unsigned short* idnext = sm.idB;
    int nodes = leaves;
    while (nodes > 1) {
        const int half = nodes >> 1;
        const int i0 = tid, i1 = tid + HASH_THREADS;
        if (D == 1) {
            // every node of this level carries the same digest x: the rest of the tree is the chain x -> H(x, x).
            // Block-constant columns repeat the same x in most chunks, so finished chains are shared through a small
            // direct-mapped memo in global memory keyed by (x, height); a miss just computes the chain.
            if (tid == 0) {
                u32 d[8], e[8];
#pragma unroll
                for (int w = 0; w < 8; w++) d[w] = Tcur[w * DD_PA];
                const u32 height = (u32)nodes;
                u32* ent = memo + (size_t)((d[0] ^ (d[1] * 0x9E3779B1u) ^ height) & (MEMO_SLOTS - 1)) * MEMO_WORDS;
                bool hit = false;
                if (ld_acquire_u32(ent) == 2u && __ldcg(ent + 1) == height) {
                    hit = true;
#pragma unroll
                    for (int w = 0; w < 8; w++) hit = hit && (__ldcg(ent + 2 + w) == d[w]);
                    if (hit) {
#pragma unroll
                        for (int w = 0; w < 8; w++) out_root[w] = __ldcg(ent + 10 + w);
                    }
                }
                if (!hit) {
                    u32 x[8];
#pragma unroll
                    for (int w = 0; w < 8; w++) x[w] = d[w];
                    for (int m2 = nodes; m2 > 1; m2 >>= 1) {
                        b3::parent(d, d, e);
#pragma unroll
                        for (int w = 0; w < 8; w++) d[w] = e[w];
                    }
#pragma unroll
                    for (int w = 0; w < 8; w++) out_root[w] = d[w];
                    if (atomicCAS(ent, 0u, 1u) == 0u) {  // claim an empty slot; busy or occupied slots are left alone
                        ent[1] = height;
#pragma unroll
                        for (int w = 0; w < 8; w++) {
                            ent[2 + w] = x[w];
                            ent[10 + w] = d[w];
                        }
                        __threadfence();
                        st_release_u32(ent, 2u);
                    }
                }
            }
            return;
        }
        if (D * D <= DD_PAIR_CAP) {
            const int space = D * D;
            for (int w = tid; w < ((space + 31) >> 5); w += HASH_THREADS) sm.bitmap[w] = 0;
            __syncthreads();
            const u32 k0 = (i0 < half) ? (u32)idcur[2 * i0] * D + idcur[2 * i0 + 1] : 0u;
            const u32 k1 = (i1 < half) ? (u32)idcur[2 * i1] * D + idcur[2 * i1 + 1] : 0u;
            const bool own0 = bitmap_set_owner(sm.bitmap, k0, i0 < half, space <= 32);
            const bool own1 = bitmap_set_owner(sm.bitmap, k1, i1 < half, space <= 32);
            __syncthreads();
            bitmap_prefix(sm, (space + 31) >> 5);
            __syncthreads();
            const int Dn = (int)sm.total;
            if (i0 < half) {
                const u32 rk = bitmap_rank(sm, k0);
                idnext[i0] = (unsigned short)rk;
                if (own0) sm.list[rk] = (unsigned short)k0;
            }
            if (i1 < half) {
                const u32 rk = bitmap_rank(sm, k1);
                idnext[i1] = (unsigned short)rk;
                if (own1) sm.list[rk] = (unsigned short)k1;
            }
            __syncthreads();
            for (int r = tid; r < Dn; r += HASH_THREADS) {
                const int kk = sm.list[r], a = kk / D, b = kk - a * D;
                u32 l[8], rr[8], d[8];
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    l[w] = Tcur[w * DD_PA + a];
                    rr[w] = Tcur[w * DD_PA + b];
                }
                b3::parent(l, rr, d);
#pragma unroll
                for (int w = 0; w < 8; w++) Tnext[w * DD_PA + r] = d[w];
            }
            __syncthreads();
            D = Dn;
            nodes = half;
            u32* tp = Tcur; Tcur = Tnext; Tnext = tp;
            unsigned short* ip = idcur; idcur = idnext; idnext = ip;
        } else {
            // too many distinct digests: compute the next level through the ids into the other half, then plain
            u32 d0[8], d1[8];
            auto indirect_parent = [&](int i, u32 (&d)[8]) {
                const int a = idcur[2 * i], b = idcur[2 * i + 1];
                u32 l[8], rr[8];
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    l[w] = Tcur[w * DD_PA + a];
                    rr[w] = Tcur[w * DD_PA + b];
                }
                b3::parent(l, rr, d);
            };
            if (i0 < half) indirect_parent(i0, d0);
            if (i1 < half) indirect_parent(i1, d1);
            if (i0 < half) put_digest(Tnext, DD_PA, i0, d0, nullptr);
            if (i1 < half) put_digest(Tnext, DD_PA, i1, d1, nullptr);
            __syncthreads();
            reduce_levels_smem(Tnext, DD_PA, half, nullptr, 0, 0, 0, nullptr, 0);
            if (tid < 8) out_root[tid] = Tnext[tid * DD_PA];
            return;
        }
    }
    if (tid < 8) out_root[tid] = Tcur[tid * DD_PA];  // single node left: id 0
}

/* ------------------------------------------------------------------------------------------ */
/* value-aware chunk commit, 128-thread variant: ~32 KB of shared memory per CTA so that 7 chunks   */
/* (instead of 4) are in flight per SM — the low-entropy phases are latency-bound chains of         */
/* dependent compressions, so concurrency across chunks is what buys throughput.                    */
/* ------------------------------------------------------------------------------------------ */
constexpr int DT = 128;
constexpr int D2_PX = 512 + 2, D2_PY = 256 + 2;
struct Dedup128Smem {
    u32 X[8 * D2_PX];  // 512 digests
    u32 Y[8 * D2_PY];  // 256 digests
    u32 bitmap[DD_PAIR_CAP / 32];
    u32 wprefix[DD_PAIR_CAP / 32];
    unsigned short idA[1 << MAX_CL];
    unsigned short idB[1 << (MAX_CL - 1)];
    unsigned short list[DD_PAIR_CAP];
    u64 red[2 * (DT / 32)];
    u32 total;
};
__device__ __forceinline__ void bitmap_prefix128(Dedup128Smem& sm, int words) {
    if (threadIdx.x < 32) {
        const int per = (words + 31) >> 5;  // <= 2
        u32 cnt[2], sum = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = threadIdx.x * per + j;
            cnt[j] = (j < per && w < words) ? __popc(sm.bitmap[w]) : 0;
            sum += cnt[j];
        }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += y;
        }
        u32 run = incl - sum;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = threadIdx.x * per + j;
            if (j < per && w < words) sm.wprefix[w] = run;
            run += cnt[j];
        }
        if (threadIdx.x == 31) sm.total = incl;
    }
}
__device__ __forceinline__ u32 bitmap_rank128(const Dedup128Smem& sm, u32 key) {
    return sm.wprefix[key >> 5] + __popc(sm.bitmap[key >> 5] & ((1u << (key & 31)) - 1u));
}
__device__ __forceinline__ void load_pair(const u32* T, int pitch, int a, int b, u32 (&l)[8], u32 (&r)[8]) {
#pragma unroll
    for (int w = 0; w < 8; w++) {
        l[w] = T[w * pitch + a];
        r[w] = T[w * pitch + b];
    }
}
__device__ __forceinline__ void store_digest(u32* T, int pitch, int i, const u32 (&d)[8]) {
#pragma unroll
    for (int w = 0; w < 8; w++) T[w * pitch + i] = d[w];
}

__global__ void __launch_bounds__(DT, 7) chunk_commit_dedup128_kernel(const u64* __restrict__ values, u64 n, u64 col_stride, int cl,
                                                                      const b3::LabelTemplate* __restrict__ templates,
                                                                      u32* __restrict__ upper, u64 n_ch, u32* memo, u64 chunk0,
                                                                      const int* __restrict__ col_list, const u32* __restrict__ work) {
    extern __shared__ __align__(16) unsigned char dd_raw[];
    Dedup128Smem& sm = *reinterpret_cast<Dedup128Smem*>(dd_raw);
    // three addressing modes: dense grid (chunk, column); grid over a column list; or a work list of (column, chunk)
    // pairs — work[0] = count, pairs from work[2] — used to redo single chunks that a tabled kernel could not serve
    u64 chunk = chunk0 + blockIdx.x;
    int col = col_list ? col_list[blockIdx.y] : (int)blockIdx.y;
    if (work) {
        if (blockIdx.x >= work[0]) return;
        col = (int)work[2 + 2 * blockIdx.x];
        chunk = work[3 + 2 * blockIdx.x];
    }
    const int leaves = 1 << cl;
    const u64* v = values + (u64)col * col_stride + (chunk << cl);
    const int tid = threadIdx.x;
    b3::LabelTemplate t;
    if (templates) t = templates[col];
    u32* out_root = upper + ((u64)col * (2 * n_ch - 1) + chunk) * 8;

    constexpr u64 HALF = 1ULL << 31;
    constexpr int LPT = (1 << MAX_CL) / DT;  // leaves per thread (8)
    u64 mn = ~0ULL, mx = 0;
#pragma unroll
    for (int k = 0; k < LPT; k++) {
        const int i = tid + k * DT;
        if (i < leaves) {
            const u64 kk = gl::add(v[i], HALF);
            mn = kk < mn ? kk : mn;
            mx = kk > mx ? kk : mx;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((tid & 31) == 0) {
        sm.red[tid >> 5] = mn;
        sm.red[DT / 32 + (tid >> 5)] = mx;
    }
    if (tid < DD_PAIR_CAP / 32) sm.bitmap[tid] = 0;
    __syncthreads();
    mn = sm.red[0];
    mx = sm.red[DT / 32];
#pragma unroll
    for (int w = 1; w < DT / 32; w++) {
        mn = sm.red[w] < mn ? sm.red[w] : mn;
        mx = sm.red[DT / 32 + w] > mx ? sm.red[DT / 32 + w] : mx;
    }
    int D = 0;
    bool dedup = (mx - mn) < (u64)(1 << MAX_CL);
    u32 ownmask = 0;
    if (dedup) {
        const bool one_word = (mx - mn) < 32;
#pragma unroll
        for (int k = 0; k < LPT; k++) {
            const int i = tid + k * DT;
            const u32 kk = i < leaves ? (u32)(gl::add(v[i], HALF) - mn) : 0u;
            if (bitmap_set_owner(sm.bitmap, kk, i < leaves, one_word)) ownmask |= 1u << k;
        }
        __syncthreads();
        bitmap_prefix128(sm, ((int)(mx - mn) + 32) >> 5);
        __syncthreads();
        D = (int)sm.total;
        dedup = D <= 256;
    }
    u32 *Tcur, *Tnext;
    int Pcur, Pnext, nodes;
    bool identity;
    unsigned short* idcur = sm.idA;
    unsigned short* idnext = sm.idB;
    if (!dedup) {
        // plain path in quarters of 256 leaves: leaf digests into Y, their 128 parents into X (level 1, <= 512 nodes)
        const int per = leaves < 256 ? leaves : 256;
        for (int base = 0; base < leaves; base += per) {
            if (templates) {
                B3_DISPATCH_LABELED(t, {
                    for (int i = tid; i < per; i += DT) {
                        u32 d[8];
                        b3::leaf_labeled_w<B3W>(t, v[base + i], d);
                        store_digest(sm.Y, D2_PY, i, d);
                    }
                })
            } else {
                for (int i = tid; i < per; i += DT) {
                    u32 d[8];
                    b3::leaf(v[base + i], d);
                    store_digest(sm.Y, D2_PY, i, d);
                }
            }
            __syncthreads();
            if (leaves == 1) break;
            for (int i = tid; i < (per >> 1); i += DT) {
                u32 l[8], r[8], d[8];
                load_pair(sm.Y, D2_PY, 2 * i, 2 * i + 1, l, r);
                b3::parent(l, r, d);
                store_digest(sm.X, D2_PX, (base >> 1) + i, d);
            }
            __syncthreads();
        }
        if (leaves == 1) {  // a single leaf is the root itself
            if (tid < 8) out_root[tid] = sm.Y[tid * D2_PY];
            return;
        }
        Tcur = sm.X; Pcur = D2_PX; Tnext = sm.Y; Pnext = D2_PY;
        nodes = leaves >> 1;
        identity = true;
    } else {
#pragma unroll
        for (int k = 0; k < LPT; k++) {
            const int i = tid + k * DT;
            if (i < leaves) {
                const u32 kk = (u32)(gl::add(v[i], HALF) - mn), rk = bitmap_rank128(sm, kk);
                sm.idA[i] = (unsigned short)rk;
                if (ownmask & (1u << k)) sm.list[rk] = (unsigned short)kk;
            }
        }
        __syncthreads();
        // one leaf hash per distinct value -> table in Y
        if (templates) {
            B3_DISPATCH_LABELED(t, {
                for (int r = tid; r < D; r += DT) {
                    u32 d[8];
                    b3::leaf_labeled_w<B3W>(t, gl::sub(mn + (u64)sm.list[r], HALF), d);
                    store_digest(sm.Y, D2_PY, r, d);
                }
            })
        } else {
            for (int r = tid; r < D; r += DT) {
                u32 d[8];
                b3::leaf(gl::sub(mn + (u64)sm.list[r], HALF), d);
                store_digest(sm.Y, D2_PY, r, d);
            }
        }
        __syncthreads();
        Tcur = sm.Y; Pcur = D2_PY; Tnext = sm.X; Pnext = D2_PX;
        nodes = leaves;
        identity = false;
    }
    while (nodes > 1) {
        if (identity) {  // plain level: parents of entries 2i, 2i+1 into the other buffer
            const int half = nodes >> 1;
            for (int i = tid; i < half; i += DT) {
                u32 l[8], r[8], d[8];
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const uint2 pr = *reinterpret_cast<const uint2*>(&Tcur[w * Pcur + 2 * i]);
                    l[w] = pr.x;
                    r[w] = pr.y;
                }
                b3::parent(l, r, d);
                store_digest(Tnext, Pnext, i, d);
            }
            __syncthreads();
            nodes = half;
            u32* tp = Tcur; Tcur = Tnext; Tnext = tp;
            const int pp = Pcur; Pcur = Pnext; Pnext = pp;
            continue;
        }
        // ---- uniform runs: while every node pairs with an identical sibling, level l+1 has the same ids and
        // digest'[id] = H(digest[id], digest[id]); k such levels are k-step chains per distinct id (memoised across
        // chunks by (x, k)) — block-constant columns and all-equal subtrees collapse here without per-level machinery
        int k = 0;
        while (nodes > 1) {
            const int half = nodes >> 1;
            int ok = 1;
            for (int i = tid; i < half; i += DT) ok &= (idcur[2 * i] == idcur[2 * i + 1]);
            if (!__syncthreads_and(ok)) break;
            for (int i = tid; i < half; i += DT) idnext[i] = idcur[2 * i];
            __syncthreads();
            unsigned short* ip = idcur; idcur = idnext; idnext = ip;
            nodes = half;
            k++;
        }
        if (k > 0) {
            for (int r = tid; r < D; r += DT) {
                u32 d[8], e[8], x[8];
#pragma unroll
                for (int w = 0; w < 8; w++) x[w] = d[w] = Tcur[w * Pcur + r];
                u32* ent = memo + (size_t)((d[0] ^ (d[1] * 0x9E3779B1u) ^ (u32)k) & (MEMO_SLOTS - 1)) * MEMO_WORDS;
                // slot states: 0 empty, 1 being computed by a resident CTA, 2 published.  The first arrival claims the
                // slot and computes; concurrent arrivals (a whole wave of chunks of the same column starts together)
                // wait for the publication instead of redoing the chain — the owner never waits on anyone, so this
                // cannot deadlock.  A published or awaited slot that turns out to hold another key is simply a miss.
                u32 st = ld_acquire_u32(ent);
                bool owner = false;
                if (st == 0u) {
                    owner = atomicCAS(ent, 0u, 1u) == 0u;
                    st = owner ? 0u : 1u;
                }
                if (st == 1u) {
                    // bounded: the slot may hold another key, or sit in state 1 for ever after a faulted launch — after a
                    // few polls (a chain of k <= 10 compressions takes ~2 us) it is treated as a miss and computed locally
                    for (int spin = 0; spin < 64 && (st = ld_acquire_u32(ent)) != 2u; spin++) __nanosleep(200);
                }
                bool hit = false;
                if (st == 2u && __ldcg(ent + 1) == (u32)k) {  // fields via L2: L1 is not coherent
                    hit = true;
#pragma unroll
                    for (int w = 0; w < 8; w++) hit = hit && (__ldcg(ent + 2 + w) == d[w]);
                    if (hit) {
#pragma unroll
                        for (int w = 0; w < 8; w++) d[w] = __ldcg(ent + 10 + w);
                    }
                }
                MEMO_STAT(hit ? 0 : 1);
                if (!hit) {
                    for (int j = 0; j < k; j++) {
                        b3::parent(d, d, e);
#pragma unroll
                        for (int w = 0; w < 8; w++) d[w] = e[w];
                    }
                    if (owner) {
                        MEMO_STAT(2);
                        ent[1] = (u32)k;
#pragma unroll
                        for (int w = 0; w < 8; w++) {
                            ent[2 + w] = x[w];
                            ent[10 + w] = d[w];
                        }
                        __threadfence();
                        st_release_u32(ent, 2u);
                    }
                }
                store_digest(Tcur, Pcur, r, d);  // in place: entry r is read and written by this thread only
            }
            __syncthreads();
            if (nodes == 1) break;
        }
        const int half = nodes >> 1;
        if (D * D <= DD_PAIR_CAP) {  // dedup level: one compression per distinct (left id, right id)
            const int space = D * D;
            const bool light = space <= 32;  // one bitmap word: ranks by popcount, no block-wide prefix
            for (int w = tid; w < ((space + 31) >> 5); w += DT) sm.bitmap[w] = 0;
            __syncthreads();
            u32 keyv[4];
            u32 own = 0;
#pragma unroll
            for (int kq = 0; kq < 4; kq++) {
                const int i = tid + kq * DT;
                keyv[kq] = (i < half) ? (u32)idcur[2 * i] * D + idcur[2 * i + 1] : 0u;
                if (bitmap_set_owner(sm.bitmap, keyv[kq], i < half, light)) own |= 1u << kq;
            }
            __syncthreads();
            int Dn;
            if (light) {
                const u32 word = sm.bitmap[0];
                Dn = __popc(word);
#pragma unroll
                for (int kq = 0; kq < 4; kq++) {
                    const int i = tid + kq * DT;
                    if (i < half) {
                        const u32 rk = __popc(word & ((1u << keyv[kq]) - 1u));
                        idnext[i] = (unsigned short)rk;
                        if (own & (1u << kq)) sm.list[rk] = (unsigned short)keyv[kq];
                    }
                }
            } else {
                bitmap_prefix128(sm, (space + 31) >> 5);
                __syncthreads();
                Dn = (int)sm.total;
#pragma unroll
                for (int kq = 0; kq < 4; kq++) {
                    const int i = tid + kq * DT;
                    if (i < half) {
                        const u32 rk = bitmap_rank128(sm, keyv[kq]);
                        idnext[i] = (unsigned short)rk;
                        if (own & (1u << kq)) sm.list[rk] = (unsigned short)keyv[kq];
                    }
                }
            }
            __syncthreads();
            for (int r = tid; r < Dn; r += DT) {
                const int kk = sm.list[r], a = kk / D, b = kk - a * D;
                u32 l[8], rr[8], d[8];
                load_pair(Tcur, Pcur, a, b, l, rr);
                b3::parent(l, rr, d);
                store_digest(Tnext, Pnext, r, d);
            }
            __syncthreads();
            D = Dn;
            unsigned short* ip = idcur; idcur = idnext; idnext = ip;
        } else {  // too many distinct digests: next level through the ids, plain afterwards
            for (int i = tid; i < half; i += DT) {
                u32 l[8], rr[8], d[8];
                load_pair(Tcur, Pcur, idcur[2 * i], idcur[2 * i + 1], l, rr);
                b3::parent(l, rr, d);
                store_digest(Tnext, Pnext, i, d);
            }
            __syncthreads();
            identity = true;
        }
        nodes = half;
        u32* tp = Tcur; Tcur = Tnext; Tnext = tp;
        const int pp = Pcur; Pcur = Pnext; Pnext = pp;
    }
    if (tid < 8) out_root[tid] = Tcur[tid * Pcur];
}

/* ------------------------------------------------------------------------------------------ */
/* structured columns: subtree tables                                                            */
/* ------------------------------------------------------------------------------------------ */
// BLAKE3 is a function, so the digest of an aligned subtree of g = 2^k leaves depends only on its g values.  Trace
// columns are highly structured, and for three structures the set of subtrees that can occur is small enough to
// hash ONCE per column into a table (<= 65536 entries, L2-resident) before the chunks are reduced:
//   ALPHA  every value lies in a range of A consecutive residues (flags, moves in {-1,0,1}, 4-bit symbols):
//          A^g subtrees; g = 16 (A = 2), 8 (A <= 4), 4 (A <= 16) or 2 (A <= 256)
//   WALK   consecutive rows differ by -1, 0 or +1 inside aligned groups of 4 (a head position): 27*A subtrees, g = 4
//   CONST  aligned runs of g = 2^k equal values (block constants: window length, head offsets): A subtrees, k <= 10
// A chunk kernel then reads the values (8 B/row, the only HBM traffic), forms each group's table index, gathers the
// 1024/g level-k digests and hashes only the levels above: 1 - 1/g of the compressions of a tree are never executed.
// The classification comes from a sample of the column and is only a performance hint: every value is checked
// against the table's assumptions while it is read, and any chunk holding a value that breaks them is redone by
// the generic value-aware kernel through a work list, so roots never depend on the classification being right.
enum { TAB_NONE = 0, TAB_ALPHA = 1, TAB_WALK = 2, TAB_CONST = 3 };
struct ColTab {
    u64 min_key;       // key = value + 2^31 (mod p); the table covers keys min_key .. min_key + A - 1
    u32 A;
    u32 cls;           // TAB_*
    u32 logg;          // log2 of the leaves under one table entry
    u32 pad;
    u32* lv[4];        // ALPHA: tables of levels 1..4 (lv[logg-1] is the one gathered from); others: lv[0]
};
struct ColStats {
    u64 min_key, max_key;
    u32 kconst;        // all sampled aligned runs of 2^kconst rows are constant (10 = whole chunks)
    u32 walk_ok;       // sampled rows move by at most one inside aligned groups of 4
};
constexpr u64 KEY_HALF = 1ULL << 31;
constexpr int VIO_CAP = 16384;  // (column, chunk) pairs a tabled pass may hand to the generic kernel

// Sampled column statistics: up to 64 whole chunks spread evenly over the rows that are already final.
// grid (columns), 256 threads, 4 consecutive rows per thread and chunk.
__global__ void __launch_bounds__(256) column_stats_kernel(const u64* __restrict__ values, u64 col_stride, u64 avail_chunks,
                                                           ColStats* __restrict__ out) {
    const u64* v = values + (u64)blockIdx.x * col_stride;
    const u64 segs = avail_chunks < 64 ? avail_chunks : 64;
    u64 mn = ~0ULL, mx = 0;
    u32 kc = MAX_CL, walk = 1;
    const int i0 = 4 * threadIdx.x;
    for (u64 sgi = 0; sgi < segs; sgi++) {
        const u64* c = v + (((avail_chunks * sgi) / segs) << MAX_CL);
        u64 prev = i0 ? gl::add(c[i0 - 1], KEY_HALF) : 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = i0 + j;
            const u64 a = gl::add(c[i], KEY_HALF);
            mn = a < mn ? a : mn;
            mx = a > mx ? a : mx;
            if (i > 0 && a != prev) {
                const u32 tz = (u32)__ffs(i) - 1;
                kc = tz < kc ? tz : kc;
                if (j > 0) {
                    const u64 d = gl::sub(a, prev);
                    if (d != 1 && d != gl::P - 1) walk = 0;
                }
            }
            prev = a;
        }
    }
    __shared__ u64 smn[256], smx[256];
    __shared__ u32 skc[256], swk[256];
    const int t = threadIdx.x;
    smn[t] = mn; smx[t] = mx; skc[t] = kc; swk[t] = walk;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) {
            smn[t] = smn[t + o] < smn[t] ? smn[t + o] : smn[t];
            smx[t] = smx[t + o] > smx[t] ? smx[t + o] : smx[t];
            skc[t] = skc[t + o] < skc[t] ? skc[t + o] : skc[t];
            swk[t] &= swk[t + o];
        }
        __syncthreads();
    }
    if (t == 0) out[blockIdx.x] = ColStats{smn[0], smx[0], skc[0], swk[0]};
}

__device__ __forceinline__ void tab_leaf(const b3::LabelTemplate* tpl, u64 v, u32 (&d)[8]) {
    if (tpl) {
        const b3::LabelTemplate t = *tpl;
        B3_DISPATCH_LABELED(t, { b3::leaf_labeled_w<B3W>(t, v, d); })
    } else b3::leaf(v, d);
}
__device__ __forceinline__ void st_digest_g(u32* dst, const u32 (&d)[8]) {
    uint4* o = reinterpret_cast<uint4*>(dst);
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
}
__device__ __forceinline__ void ld_digest_g(const u32* src, u32 (&d)[8]) {
    const uint4* e = reinterpret_cast<const uint4*>(src);
    const uint4 e0 = __ldg(e), e1 = __ldg(e + 1);
    d[0] = e0.x; d[1] = e0.y; d[2] = e0.z; d[3] = e0.w;
    d[4] = e1.x; d[5] = e1.y; d[6] = e1.z; d[7] = e1.w;
}

// First table of every tabled column (grid.y indexes `list`): ALPHA lv[0][a0 + A*a1] = H(leaf a0, leaf a1);
// WALK lv[0][a0 + A*(d1 + 3*d2 + 9*d3)] = the 4-leaf subtree of a0, a0+d1-1, ...; CONST lv[0][a] = 2^logg equal leaves.
__global__ void __launch_bounds__(128) tab_build_first_kernel(const ColTab* __restrict__ tabs, const int* __restrict__ list,
                                                              const b3::LabelTemplate* __restrict__ templates) {
    const int col = list[blockIdx.y];
    const ColTab tb = tabs[col];
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const b3::LabelTemplate* tpl = templates ? templates + col : nullptr;
    auto val = [&](u32 a) { return gl::sub(gl::add(tb.min_key, (u64)a), KEY_HALF); };  // key offset -> field element
    u32 d[8];
    if (tb.cls == TAB_ALPHA) {
        if (i >= tb.A * tb.A) return;
        u32 l[8], r[8];
        tab_leaf(tpl, val(i % tb.A), l);
        tab_leaf(tpl, val(i / tb.A), r);
        b3::parent(l, r, d);
    } else if (tb.cls == TAB_WALK) {
        if (i >= 27 * tb.A) return;
        const u32 a0 = i % tb.A, code = i / tb.A;
        // offsets relative to a0 lie in [-3, 3]; keys outside the table range are still field elements
        const u64 k0 = gl::add(tb.min_key, (u64)a0);
        const u64 k1 = gl::sub(gl::add(k0, (u64)(code % 3)), 1), k2 = gl::sub(gl::add(k1, (u64)((code / 3) % 3)), 1),
                  k3 = gl::sub(gl::add(k2, (u64)(code / 9)), 1);
        u32 l[8], r[8], p0[8], p1[8];
        tab_leaf(tpl, gl::sub(k0, KEY_HALF), l);
        tab_leaf(tpl, gl::sub(k1, KEY_HALF), r);
        b3::parent(l, r, p0);
        tab_leaf(tpl, gl::sub(k2, KEY_HALF), l);
        tab_leaf(tpl, gl::sub(k3, KEY_HALF), r);
        b3::parent(l, r, p1);
        b3::parent(p0, p1, d);
    } else {
        if (i >= tb.A) return;
        u32 e[8];
        tab_leaf(tpl, val(i), d);
        for (u32 j = 0; j < tb.logg; j++) {
            b3::parent(d, d, e);
#pragma unroll
            for (int w = 0; w < 8; w++) d[w] = e[w];
        }
    }
    st_digest_g(tb.lv[0] + (size_t)i * 8, d);
}
// ALPHA level s+1 from level s: lv[s][i] = H(lv[s-1][i % cnt], lv[s-1][i / cnt]), cnt = A^(2^s) entries below.
__global__ void __launch_bounds__(128) tab_build_up_kernel(const ColTab* __restrict__ tabs, const int* __restrict__ list, int s) {
    const ColTab tb = tabs[list[blockIdx.y]];
    if (tb.cls != TAB_ALPHA || (int)tb.logg <= s) return;
    u32 cnt = tb.A;
    for (int j = 0; j < s; j++) cnt *= cnt;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (u64)cnt * cnt) return;
    u32 l[8], r[8], d[8];
    ld_digest_g(tb.lv[s - 1] + (size_t)(i % cnt) * 8, l);
    ld_digest_g(tb.lv[s - 1] + (size_t)(i / cnt) * 8, r);
    b3::parent(l, r, d);
    st_digest_g(tb.lv[s] + (size_t)i * 8, d);
}

// Chunk roots of tabled columns.  One CTA gathers the table entries of 2^(LOGG-1) <= 8 chunks (at most 512 nodes)
// and reduces them to chunk roots.  Each warp walks over contiguous 256-row steps (8 consecutive rows per lane, two
// full sectors), so a group of up to 256 rows is keyed with warp shuffles and longer runs are carried across steps.
// grid (ceil(chunks / tab_chunks_per_cta(LOGG)), columns of this LOGG), 128 threads.
constexpr int tab_chunks_per_cta(int logg) { return logg <= 4 ? (1 << (logg - 1)) : 8; }
template <int LOGG>
__global__ void __launch_bounds__(DT, 8) chunk_commit_tabled_kernel(const u64* __restrict__ values, u64 col_stride,
                                                                    const ColTab* __restrict__ tabs, const int* __restrict__ col_list,
                                                                    u32* __restrict__ upper, u64 n_ch, u64 chunk0, u64 chunk1,
                                                                    u32* __restrict__ vio) {
    __shared__ __align__(16) u32 X[8 * D2_PX];
    __shared__ __align__(16) u32 Y[8 * D2_PY];
    constexpr int CH = tab_chunks_per_cta(LOGG);                 // chunks per CTA
    constexpr int STEPS_PER_WARP = CH;                           // CH*4 steps of 256 rows over 4 warps
    constexpr int NODES = (CH << MAX_CL) >> LOGG;                // table entries gathered per CTA (<= 512)
    constexpr int NPS = LOGG <= 8 ? (256 >> LOGG) : 0;           // nodes per step (0: a node spans several steps)
    const int col = col_list[blockIdx.y];
    const ColTab tb = tabs[col];
    const u32* table = tb.lv[tb.cls == TAB_ALPHA ? LOGG - 1 : 0];
    const u64 cta_chunk0 = chunk0 + (u64)blockIdx.x * CH;
    const u64* v = values + (u64)col * col_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 A = tb.A;

    u32 run_key = 0;   // LOGG > 8: key and state of the run being assembled
    bool run_bad = false;
    for (int st = 0; st < STEPS_PER_WARP; st++) {
        const int step = warp * STEPS_PER_WARP + st;             // 256-row step inside the CTA's rows
        const u64 row = (cta_chunk0 << MAX_CL) + ((u64)step << 8);
        const u64 chunk = row >> MAX_CL;
        if (chunk >= chunk1) break;                              // warp-uniform
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(v + row + 8 * lane);
        const ulonglong2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
        const u64 raw[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
        u32 a[8];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const u64 k = gl::sub(gl::add(raw[j], KEY_HALF), tb.min_key);
            bad |= k >= (u64)A;
            a[j] = (u32)k;
        }
        u32 idx[4] = {0, 0, 0, 0};   // table indices of this lane's nodes (LOGG = 1: 4, 2: 2, >= 3: at most one)
        bool lead = true;            // this lane gathers idx[0]
        if (tb.cls == TAB_ALPHA) {
            if (LOGG == 1) {
#pragma unroll
                for (int m = 0; m < 4; m++) idx[m] = a[2 * m] + A * a[2 * m + 1];
            } else if (LOGG == 2) {
#pragma unroll
                for (int m = 0; m < 2; m++) idx[m] = a[4 * m] + A * (a[4 * m + 1] + A * (a[4 * m + 2] + A * a[4 * m + 3]));
            } else {
                u32 h = 0;
#pragma unroll
                for (int j = 7; j >= 0; j--) h = h * A + a[j];
                if (LOGG == 4) {   // A = 2: two lanes per node
                    const u32 o = __shfl_xor_sync(0xffffffffu, h, 1);
                    bad |= (bool)__shfl_xor_sync(0xffffffffu, (int)bad, 1);
                    h += o << 8;
                    lead = (lane & 1) == 0;
                }
                idx[0] = h;
            }
        } else if (tb.cls == TAB_WALK) {  // LOGG == 2
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const u32 d1 = a[4 * m + 1] - a[4 * m] + 1, d2 = a[4 * m + 2] - a[4 * m + 1] + 1, d3 = a[4 * m + 3] - a[4 * m + 2] + 1;
                bad |= (d1 > 2) | (d2 > 2) | (d3 > 2);
                idx[m] = a[4 * m] + A * (d1 + 3 * d2 + 9 * d3);
            }
        } else {  // TAB_CONST: runs of 2^LOGG equal values
            if (LOGG == 1) {
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    bad |= a[2 * m] != a[2 * m + 1];
                    idx[m] = a[2 * m];
                }
            } else if (LOGG == 2) {
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    bad |= (a[4 * m] != a[4 * m + 1]) | (a[4 * m] != a[4 * m + 2]) | (a[4 * m] != a[4 * m + 3]);
                    idx[m] = a[4 * m];
                }
            } else {
#pragma unroll
                for (int j = 1; j < 8; j++) bad |= a[j] != a[0];
                constexpr int SEG = LOGG >= 8 ? 32 : (1 << (LOGG >= 3 ? LOGG - 3 : 0));   // lanes per run inside a step
                const int leader = lane & ~(SEG - 1);
                bad |= a[0] != __shfl_sync(0xffffffffu, a[0], leader);
                const u32 bm = __ballot_sync(0xffffffffu, bad);
                const u32 segmask = (SEG == 32 ? 0xffffffffu : ((1u << SEG) - 1u)) << leader;
                bad = (bm & segmask) != 0;
                lead = lane == leader;
                idx[0] = a[0];
                if (LOGG > 8) {   // the run spans 2^(LOGG-8) steps of this warp
                    constexpr int SPN = 1 << (LOGG > 8 ? LOGG - 8 : 0);
                    if ((st & (SPN - 1)) == 0) {
                        run_key = a[0];
                        run_bad = false;
                    }
                    run_bad |= bad | (a[0] != run_key);
                    bad = run_bad;
                    lead = lead && ((st & (SPN - 1)) == SPN - 1);
                }
            }
        }
        if (__any_sync(0xffffffffu, bad)) {
            // hand the chunk to the generic kernel (one entry per warp and step is enough; duplicates are harmless)
            if (lane == 0) {
                const u32 slot = atomicAdd(&vio[0], 1u);
                if (slot < VIO_CAP) {
                    vio[2 + 2 * slot] = (u32)col;
                    vio[3 + 2 * slot] = (u32)chunk;
                }
            }
#pragma unroll
            for (int m = 0; m < 4; m++) idx[m] = 0;
        }
        // gather
        if (LOGG == 1) {
#pragma unroll
            for (int m = 0; m < 4; m++) {
                u32 d[8];
                ld_digest_g(table + (size_t)idx[m] * 8, d);
                store_digest(X, D2_PX, step * 128 + lane * 4 + m, d);
            }
        } else if (LOGG == 2) {
#pragma unroll
            for (int m = 0; m < 2; m++) {
                u32 d[8];
                ld_digest_g(table + (size_t)idx[m] * 8, d);
                store_digest(X, D2_PX, step * 64 + lane * 2 + m, d);
            }
        } else if (lead) {
            u32 d[8];
            ld_digest_g(table + (size_t)idx[0] * 8, d);
            int node;
            if (LOGG <= 8) node = step * NPS + (lane >> (LOGG - 3 > 0 ? LOGG - 3 : 0));
            else node = step >> (LOGG - 8);
            store_digest(X, D2_PX, node, d);
        }
    }
    __syncthreads();
    // NODES nodes of level LOGG -> chunk roots (level 10): 10 - LOGG plain levels, ping-pong X -> Y -> X ...
    u32* Tcur = X;
    u32* Tnext = Y;
    int Pcur = D2_PX, Pnext = D2_PY;
    int nodes = NODES;
#pragma unroll 1
    for (int lvl = LOGG; lvl < MAX_CL; lvl++) {
        const int half = nodes >> 1;
        for (int i = threadIdx.x; i < half; i += DT) {
            u32 l[8], r[8], d[8];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint2 pr = *reinterpret_cast<const uint2*>(&Tcur[w * Pcur + 2 * i]);
                l[w] = pr.x;
                r[w] = pr.y;
            }
            b3::parent(l, r, d);
            store_digest(Tnext, Pnext, i, d);
        }
        __syncthreads();
        nodes = half;
        u32* tp = Tcur; Tcur = Tnext; Tnext = tp;
        const int pp = Pcur; Pcur = Pnext; Pnext = pp;
    }
    // nodes == CH chunk roots
    u32* out = upper + ((u64)col * (2 * n_ch - 1) + cta_chunk0) * 8;
    const u64 valid = chunk1 > cta_chunk0 ? (chunk1 - cta_chunk0 < (u64)CH ? chunk1 - cta_chunk0 : (u64)CH) : 0;
    for (int i = threadIdx.x; i < (int)valid * 8; i += DT) out[i] = Tcur[(i & 7) * Pcur + (i >> 3)];
}

// digests at level l0 of `upper` -> reduce groups of 2^k -> levels l0+1..l0+k stored.  grid (count>>k, cols)
__global__ void __launch_bounds__(HASH_THREADS, 5) upper_reduce_kernel(u32* __restrict__ upper, u64 n_ch, int l0, int k) {
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    const u64 grp = blockIdx.x;
    const int col = blockIdx.y;
    u32* base = upper + (u64)col * (2 * n_ch - 1) * 8;
    const u32* src = base + ((2 * n_ch - ((2 * n_ch) >> l0)) + (grp << k)) * 8;
    const int cnt = 1 << k;
    for (int i = threadIdx.x; i < cnt * 8; i += HASH_THREADS) s[(i & 7) * pitch + (i >> 3)] = src[i];
    __syncthreads();
    reduce_levels_smem(s, pitch, cnt, base, grp, l0, n_ch, nullptr, 0);
}

// The same for several single-column commitments in one launch (the FRI layers: their upper levels are off the
// critical path of the fold chain, so they are reduced together at the end instead of 2-3 latency-bound launches per
// layer).  Job j owns CTAs [cta0[j], cta0[j+1]); a job that reaches its root also writes it to root_out.
__global__ void __launch_bounds__(HASH_THREADS, 5) upper_reduce_multi_kernel(const UpperJobs jobs) {
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    int j = 0;
    while (j + 1 < jobs.n && blockIdx.x >= jobs.cta0[j + 1]) j++;
    const UpperJob jb = jobs.j[j];
    const u64 grp = jb.grp0 + (blockIdx.x - jobs.cta0[j]);
    const u32* src = jb.upper + ((2 * jb.n_ch - ((2 * jb.n_ch) >> jb.l0)) + (grp << jb.k)) * 8;
    const int cnt = 1 << jb.k;
    for (int i = threadIdx.x; i < cnt * 8; i += HASH_THREADS) s[(i & 7) * pitch + (i >> 3)] = src[i];
    __syncthreads();
    reduce_levels_smem(s, pitch, cnt, jb.upper, grp, jb.l0, jb.n_ch, nullptr, 0);
    if (jb.root_out && (jb.n_ch >> (jb.l0 + jb.k)) == 1 && threadIdx.x < 8) jb.root_out[threadIdx.x] = s[threadIdx.x * pitch];
}

// One CTA per opening: rebuild the chunk, record the in-chunk sibling path and chunk root, gather the upper path.
// Requests carry their own commitment pointers so that openings into many commitments (all FRI layers, all
// columns) go out in ONE launch.
__global__ void __launch_bounds__(HASH_THREADS, 5) open_kernel(const OpenReq* __restrict__ reqs, u64* __restrict__ out_values,
                                                            u32* __restrict__ out_chunk_roots, u32* __restrict__ out_paths) {
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    const u64 q = blockIdx.x;
    const OpenReq rq = reqs[q];
    const int cl = (int)rq.cl;
    const u64 chunk = rq.row >> cl;
    const int idx_in = (int)(rq.row & ((1ULL << cl) - 1));
    const int leaves = 1 << cl;
    const u64* v = rq.values + (chunk << cl);
    b3::LabelTemplate t;
    if (rq.tpl) t = *rq.tpl;
    HASH_LEAVES_INTO(s, pitch, leaves, v[i], rq.tpl != nullptr, t)
    if (threadIdx.x == 0) out_values[q] = v[idx_in];
    __syncthreads();
    u32* path = out_paths + (u64)rq.out_off * 8;
    reduce_levels_smem(s, pitch, leaves, nullptr, 0, 0, 0, path, idx_in);
    if (threadIdx.x < 8) out_chunk_roots[q * 8 + threadIdx.x] = s[threadIdx.x * pitch];
    const u64 n_ch = rq.n_ch;
    for (int i = threadIdx.x; i < (int)rq.depth_out * 8; i += HASH_THREADS) {
        const int l = i >> 3, w = i & 7;
        const u64 sib = (chunk >> l) ^ 1;
        path[(u64)cl * 8 + i] = rq.upper[((2 * n_ch - ((2 * n_ch) >> l)) + sib) * 8 + w];
    }
}

__global__ void leaf_hash_kernel(const u64* __restrict__ vals, size_t n, const b3::LabelTemplate* __restrict__ tpl,
                                 u32* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 d[8];
    if (tpl) {
        b3::LabelTemplate t = *tpl;
        B3_DISPATCH_LABELED(t, { b3::leaf_labeled_w<B3W>(t, vals[i], d); })
    } else b3::leaf(vals[i], d);
    uint4* o = reinterpret_cast<uint4*>(out + i * 8);
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

// One Merkle level with odd promotion: out[i] = H(in[2i], in[2i+1]) or in[2i] when 2i+1 == n.
__global__ void merkle_level_kernel(const u32* __restrict__ in, size_t n, u32* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_out = (n + 1) / 2;
    if (i >= n_out) return;
    const uint4* p = reinterpret_cast<const uint4*>(in + 16 * i);
    uint4 a0 = p[0], a1 = p[1];
    u32 d[8];
    if (2 * i + 1 < n) {
        uint4 b0 = p[2], b1 = p[3];
        u32 l[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        u32 r[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        b3::parent(l, r, d);
    } else {
        d[0] = a0.x; d[1] = a0.y; d[2] = a0.z; d[3] = a0.w;
        d[4] = a1.x; d[5] = a1.y; d[6] = a1.z; d[7] = a1.w;
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 8);
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

}  // namespace

b3::LabelTemplate make_label_template(const char* label) {
    const size_t L = std::strlen(label);
    REQUIRE(L <= 44, "column label longer than 44 bytes (labeled leaf must fit one BLAKE3 block)");
    u8 bytes[64] = {0};
    std::memcpy(bytes, "col_leaf", 8);  // reference v1/params.rs:58 DS_COL_LEAF
    const u32 l32 = (u32)L;
    std::memcpy(bytes + 8, &l32, 4);
    std::memcpy(bytes + 12, label, L);
    b3::LabelTemplate t;
    std::memcpy(t.words, bytes, 64);
    t.off = (u32)(12 + L);
    t.block_len = (u32)(20 + L);
    b3::label_template_precompute(t);
    return t;
}

extern "C" void sezkp_debug_memo_stats(unsigned long long out[4], int reset) {
    cudaMemcpyFromSymbol(out, g_memo_stats, sizeof(unsigned long long) * 4);
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_memo_stats, z, sizeof z);
    }
}

// debug: columns served from subtree tables / chunks handed back to the generic kernel in the last value-aware commit
extern "C" void sezkp_debug_tab_stats(sezkp_ctx* ctx, unsigned long long out[2]) {
    out[0] = (unsigned long long)ctx->tab_columns;
    out[1] = ctx->tab_redone_chunks;
}

void Commit::release(sezkp_ctx* ctx) {
    if (owns_values && values) ctx->pool.free((void*)values);
    ctx->pool.free(upper);
    ctx->pool.free(templates);
    ctx->pool.free(tab_dev);
    tab_dev = nullptr;
    tab_classified = false;
    values = nullptr;
    upper = nullptr;
    templates = nullptr;
}

// ---- host side of the subtree tables --------------------------------------------------------------------------
namespace {
struct TabLayout {
    size_t off_tabs, off_stats, off_lists, off_vio, bytes;
    explicit TabLayout(int cols) {
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        off_tabs = 0;
        off_stats = al(off_tabs + sizeof(ColTab) * cols);
        off_lists = al(off_stats + sizeof(ColStats) * cols);
        off_vio = al(off_lists + sizeof(int) * cols);
        bytes = al(off_vio + sizeof(u32) * (2 + 2 * VIO_CAP));
    }
};
template <int LOGG>
void launch_tabled(sezkp_ctx* ctx, const Commit& cm, const ColTab* tabs, const int* list, int count, u64 chunk0, u64 chunk1, u32* vio) {
    constexpr u64 CH = tab_chunks_per_cta(LOGG);
    dim3 grid((unsigned)((chunk1 - chunk0 + CH - 1) / CH), (unsigned)count);
    chunk_commit_tabled_kernel<LOGG><<<grid, DT, 0, ctx->stream>>>(cm.values, cm.col_stride, tabs, list, cm.upper, cm.n_ch, chunk0, chunk1, vio);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

// Sample the columns, pick a table class per column, build the tables.  Rows of chunks [0, avail_chunks) are final.
void tab_classify(sezkp_ctx* ctx, Commit& cm, u64 avail_chunks) {
    const int cols = cm.cols;
    const TabLayout lay(cols);
    u8* base = (u8*)ctx->pool.alloc(lay.bytes);
    cm.tab_dev = base;
    ColTab* tabs_dev = (ColTab*)(base + lay.off_tabs);
    ColStats* stats_dev = (ColStats*)(base + lay.off_stats);
    int* lists_dev = (int*)(base + lay.off_lists);
    u32* vio = (u32*)(base + lay.off_vio);
    column_stats_kernel<<<cols, 256, 0, ctx->stream>>>(cm.values, cm.col_stride, avail_chunks, stats_dev);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    std::vector<ColStats> st(cols);
    CUDA_CHECK(cudaMemcpyAsync(st.data(), stats_dev, sizeof(ColStats) * cols, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(vio, 0, 8, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));

    std::vector<ColTab> tabs(cols);
    std::vector<size_t> entries(cols, 0);
    cm.tab_logg.assign(cols, 0);
    size_t total_entries = 0;
    for (int c = 0; c < cols; c++) {
        ColTab& tb = tabs[c];
        std::memset(&tb, 0, sizeof tb);
        const ColStats& s = st[c];
        if (s.max_key < s.min_key || s.max_key - s.min_key >= 16384) continue;
        const u32 A0 = (u32)(s.max_key - s.min_key) + 1;
        const int la = A0 <= 2 ? 4 : A0 <= 4 ? 3 : A0 <= 16 ? 2 : A0 <= 256 ? 1 : 0;
        const int lw = (s.walk_ok && A0 <= 680) ? 2 : 0;
        const int lc = (int)s.kconst;
        int logg = la, cls = TAB_ALPHA;
        if (lw > logg) logg = lw, cls = TAB_WALK;
        if (lc > logg) logg = lc, cls = TAB_CONST;
        if (logg == 0) continue;
        // ranges come from a sample: pad them where the table stays small (values outside still only cost a redo)
        u32 A = A0;
        if (cls == TAB_ALPHA) {
            if (logg == 4) A = 2;
            else if (logg == 1) A = std::min<u32>(256, 2 * A0 + 8);
        } else if (cls == TAB_WALK) A = std::min<u32>(2048, 3 * A0 + 64);
        else A = std::min<u32>(65536, 3 * A0 + 64);
        u64 lo = (A - A0) / 2;
        if (lo > s.min_key) lo = s.min_key;
        tb.min_key = s.min_key - lo;
        tb.A = A;
        tb.cls = (u32)cls;
        tb.logg = (u32)logg;
        size_t e = 0;
        if (cls == TAB_ALPHA) {
            size_t cnt = A;
            for (int l = 0; l < logg; l++) {
                cnt *= cnt;
                e += cnt;
            }
        } else e = cls == TAB_WALK ? 27 * (size_t)A : (size_t)A;
        entries[c] = e;
        total_entries += e;
        cm.tab_logg[c] = logg;
    }
    u32* tables = (u32*)ctx->scratch[11].ensure(total_entries * 32 + 256);
    size_t off = 0, max_first = 0, max_up[4] = {0, 0, 0, 0};
    std::vector<int> order;  // tabled columns sorted by logg, then the generic ones
    for (int l = 1; l <= MAX_CL; l++)
        for (int c = 0; c < cols; c++)
            if (cm.tab_logg[c] == l) order.push_back(c);
    const int n_tab = (int)order.size();
    for (int c = 0; c < cols; c++)
        if (cm.tab_logg[c] == 0) order.push_back(c);
    for (int c = 0; c < cols; c++) {
        ColTab& tb = tabs[c];
        if (!tb.cls) continue;
        if (tb.cls == TAB_ALPHA) {
            size_t cnt = tb.A;
            for (u32 l = 0; l < tb.logg; l++) {
                cnt *= cnt;
                tb.lv[l] = tables + off * 8;
                off += cnt;
                if (l == 0) max_first = std::max(max_first, cnt);
                else max_up[l] = std::max(max_up[l], cnt);
            }
        } else {
            tb.lv[0] = tables + off * 8;
            off += entries[c];
            max_first = std::max(max_first, entries[c]);
        }
    }
    CUDA_CHECK(cudaMemcpyAsync(tabs_dev, tabs.data(), sizeof(ColTab) * cols, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(lists_dev, order.data(), sizeof(int) * cols, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    // The tables are pure functions of (class, key range, label): a context that proves trace after trace of the same
    // machine finds them already built (the key holds the table parameters with their addresses and the label templates).
    std::vector<u8> key(sizeof(ColTab) * cols + sizeof(b3::LabelTemplate) * cm.templates_host.size() + sizeof(int) * cols);
    std::memcpy(key.data(), tabs.data(), sizeof(ColTab) * cols);
    std::memcpy(key.data() + sizeof(ColTab) * cols, order.data(), sizeof(int) * cols);
    if (!cm.templates_host.empty())
        std::memcpy(key.data() + (sizeof(ColTab) + sizeof(int)) * cols, cm.templates_host.data(), sizeof(b3::LabelTemplate) * cm.templates_host.size());
    const bool cached = n_tab && ctx->tab_cache_enabled && key == ctx->tab_cache_key;
    if (n_tab && !cached) {
        ctx->tab_cache_key.clear();  // invalid while the kernels below rewrite the tables
        tab_build_first_kernel<<<dim3((unsigned)((max_first + 127) / 128), (unsigned)n_tab), 128, 0, ctx->stream>>>(tabs_dev, lists_dev, cm.templates);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
        for (int l = 1; l < 4; l++) {
            if (!max_up[l]) continue;
            tab_build_up_kernel<<<dim3((unsigned)((max_up[l] + 127) / 128), (unsigned)n_tab), 128, 0, ctx->stream>>>(tabs_dev, lists_dev, l);
            CUDA_CHECK(cudaGetLastError());
            ctx->launches++;
        }
        ctx->tab_cache_key = std::move(key);
    }
    cm.tab_classified = true;
    ctx->tab_columns = n_tab;
}

void tab_commit_chunks(sezkp_ctx* ctx, Commit& cm, u64 chunk0, u64 chunk1, u32* memo) {
    if (!cm.tab_classified) tab_classify(ctx, cm, chunk1);
    const TabLayout lay(cm.cols);
    u8* base = (u8*)cm.tab_dev;
    const ColTab* tabs = (const ColTab*)(base + lay.off_tabs);
    const int* lists = (const int*)(base + lay.off_lists);
    u32* vio = (u32*)(base + lay.off_vio);
    int pos = 0;
    for (int l = 1; l <= MAX_CL; l++) {
        int count = 0;
        for (int c = 0; c < cm.cols; c++) count += cm.tab_logg[c] == l;
        if (!count) continue;
        switch (l) {
            case 1: launch_tabled<1>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 2: launch_tabled<2>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 3: launch_tabled<3>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 4: launch_tabled<4>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 5: launch_tabled<5>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 6: launch_tabled<6>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 7: launch_tabled<7>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 8: launch_tabled<8>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            case 9: launch_tabled<9>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
            default: launch_tabled<10>(ctx, cm, tabs, lists + pos, count, chunk0, chunk1, vio); break;
        }
        pos += count;
    }
    if (pos < cm.cols) {
        dim3 grid((unsigned)(chunk1 - chunk0), (unsigned)(cm.cols - pos));
        chunk_commit_dedup128_kernel<<<grid, DT, sizeof(Dedup128Smem), ctx->stream>>>(cm.values, cm.n, cm.col_stride, cm.cl, cm.templates, cm.upper,
                                                                                      cm.n_ch, memo, chunk0, lists + pos, nullptr);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    }
    cm.tab_chunks_done = chunk1;
}

// Chunks that broke their column's table assumptions are redone by the generic kernel (same outputs).
void tab_commit_finish(sezkp_ctx* ctx, Commit& cm, u32* memo) {
    const TabLayout lay(cm.cols);
    u8* base = (u8*)cm.tab_dev;
    const int* lists = (const int*)(base + lay.off_lists);
    u32* vio = (u32*)(base + lay.off_vio);
    u32 count = 0;
    CUDA_CHECK(cudaMemcpyAsync(&count, vio, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->tab_redone_chunks = count;
    if (count == 0) return;
    if (count <= VIO_CAP) {
        chunk_commit_dedup128_kernel<<<count, DT, sizeof(Dedup128Smem), ctx->stream>>>(cm.values, cm.n, cm.col_stride, cm.cl, cm.templates, cm.upper,
                                                                                       cm.n_ch, memo, 0, nullptr, vio);
    } else {  // the sample misjudged the columns: redo every tabled column
        int n_tab = 0;
        for (int c = 0; c < cm.cols; c++) n_tab += cm.tab_logg[c] != 0;
        dim3 grid((unsigned)cm.tab_chunks_done, (unsigned)n_tab);
        chunk_commit_dedup128_kernel<<<grid, DT, sizeof(Dedup128Smem), ctx->stream>>>(cm.values, cm.n, cm.col_stride, cm.cl, cm.templates, cm.upper,
                                                                                      cm.n_ch, memo, 0, lists, nullptr);
    }
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}
}  // namespace

// commit_begin / commit_chunks / commit_finish: the three stages of commit_build, exposed so that a caller can hash
// ranges of chunks as their rows become available (pipelined H2D in the prover) and build the upper levels at the end.
void commit_begin(sezkp_ctx* ctx, Commit& cm, const u64* values_dev, u64 n, int cols, int chunk_log2, const char* const* labels,
                  const CommitOpts& opt) {
    REQUIRE(n >= 1 && (n & (n - 1)) == 0, "column length %llu is not a power of two", (unsigned long long)n);
    REQUIRE(cols >= 1 && cols <= 65535, "column count %d out of range", cols);
    REQUIRE(chunk_log2 >= 0 && chunk_log2 <= MAX_CL, "chunk_log2 %d out of range (0..10)", chunk_log2);
    const int ln = ilog2(n);
    cm.values = values_dev;
    cm.n = n;
    cm.col_stride = opt.col_stride ? opt.col_stride : n;
    cm.cols = cols;
    cm.cl = ln < chunk_log2 ? ln : chunk_log2;
    cm.cta_cl = cm.cl;
    if (opt.cta_log2 > cm.cl) cm.cta_cl = std::min(std::min(opt.cta_log2, MAX_CL), ln);
    cm.n_ch = n >> cm.cl;
    REQUIRE(cm.n_ch < (1ULL << 31), "too many chunks per column");
    if (labels) {
        std::vector<b3::LabelTemplate> h(cols);
        for (int c = 0; c < cols; c++) {
            REQUIRE(labels[c] != nullptr, "label %d is NULL", c);
            h[c] = make_label_template(labels[c]);
        }
        cm.templates = (b3::LabelTemplate*)ctx->pool.alloc(sizeof(b3::LabelTemplate) * cols);
        CUDA_CHECK(cudaMemcpyAsync(cm.templates, h.data(), sizeof(b3::LabelTemplate) * cols, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // h goes out of scope
        cm.templates_host = h;
    }
    cm.upper = (u32*)ctx->pool.alloc((size_t)cols * (2 * cm.n_ch - 1) * 32);
    if (opt.dedup && ctx->dedup_enabled && !opt.fold_src) {
        size_t& configured = ctx->func_smem[(const void*)chunk_commit_dedup128_kernel];  // per context: the attribute is per device
        if (!configured) {
            CUDA_CHECK(cudaFuncSetAttribute(chunk_commit_dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DedupSmem)));
            CUDA_CHECK(cudaFuncSetAttribute(chunk_commit_dedup128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Dedup128Smem)));
            configured = 1;
        }
        u32* memo = (u32*)ctx->scratch[8].ensure((size_t)MEMO_SLOTS * MEMO_WORDS * 4);
        CUDA_CHECK(cudaMemsetAsync(memo, 0, (size_t)MEMO_SLOTS * MEMO_WORDS * 4, ctx->stream));
    }
}

void commit_chunks(sezkp_ctx* ctx, Commit& cm, u64 chunk0, u64 chunk1, const CommitOpts& opt) {
    if (chunk1 <= chunk0) return;
    REQUIRE(chunk1 <= cm.n_ch, "internal: chunk range out of bounds");
    const int sub = cm.cta_cl - cm.cl;  // chunks per CTA of the plain kernel = 2^sub
    REQUIRE(((chunk0 | chunk1) & ((1ULL << sub) - 1)) == 0, "internal: chunk range not aligned to the CTA granularity");
    dim3 grid((unsigned)(chunk1 - chunk0), (unsigned)cm.cols);
    u64* vals = (u64*)cm.values;
    if (opt.fold_src) {
        REQUIRE(cm.cols == 1 && !cm.templates, "internal: fused fold needs a single unlabeled column");
        grid.x >>= sub;
        chunk_commit_kernel<true><<<grid, HASH_THREADS, 0, ctx->stream>>>(vals, cm.n, cm.col_stride, cm.cl, sub, nullptr, cm.upper, cm.n_ch,
                                                                           opt.fold_src, opt.fold_beta, chunk0);
    } else if (opt.dedup && ctx->dedup_enabled) {
        REQUIRE(sub == 0, "internal: the value-aware kernels reduce exactly one chunk per CTA");
        u32* memo = (u32*)ctx->scratch[8].p;
        if (ctx->dedup_variant == 2 && ctx->tabled_enabled && cm.cl == MAX_CL) {
            tab_commit_chunks(ctx, cm, chunk0, chunk1, memo);
            return;
        }
        if (ctx->dedup_variant == 2)
            chunk_commit_dedup128_kernel<<<grid, DT, sizeof(Dedup128Smem), ctx->stream>>>(cm.values, cm.n, cm.col_stride, cm.cl, cm.templates,
                                                                                          cm.upper, cm.n_ch, memo, chunk0, nullptr, nullptr);
        else
            chunk_commit_dedup_kernel<<<grid, HASH_THREADS, sizeof(DedupSmem), ctx->stream>>>(cm.values, cm.n, cm.col_stride, cm.cl,
                                                                                               cm.templates, cm.upper, cm.n_ch, memo, chunk0);
    } else {
        grid.x >>= sub;
        chunk_commit_kernel<false><<<grid, HASH_THREADS, 0, ctx->stream>>>(vals, cm.n, cm.col_stride, cm.cl, sub, cm.templates, cm.upper, cm.n_ch,
                                                                            nullptr, 0, chunk0);
    }
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void commit_finish(sezkp_ctx* ctx, Commit& cm, const CommitOpts& opt) {
    if (cm.tab_classified) tab_commit_finish(ctx, cm, (u32*)ctx->scratch[8].p);
    const int depth = ilog2(cm.n_ch);
    int l0 = 0;
    while (l0 < depth) {
        const int k = (depth - l0) < MAX_CL ? (depth - l0) : MAX_CL;
        dim3 g((unsigned)(cm.n_ch >> (l0 + k)), (unsigned)cm.cols);
        upper_reduce_kernel<<<g, HASH_THREADS, 0, ctx->stream>>>(cm.upper, cm.n_ch, l0, k);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
        l0 += k;
    }
    const u8* root0 = (const u8*)cm.upper + (2 * cm.n_ch - 2) * 32;
    const size_t col_pitch = (2 * cm.n_ch - 1) * 32;
    if (opt.roots_dev) CUDA_CHECK(cudaMemcpy2DAsync(opt.roots_dev, 32, root0, col_pitch, 32, cm.cols, cudaMemcpyDeviceToDevice, ctx->stream));
    if (opt.roots_host) {
        CUDA_CHECK(cudaMemcpy2DAsync(opt.roots_host, 32, root0, col_pitch, 32, cm.cols, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
}

// commit_chunks for a batch of single-column unlabeled commitments (begin done, values final): one launch.
void commit_chunks_multi(sezkp_ctx* ctx, Commit* cms, int count) {
    for (int base = 0; base < count; base += UPPER_MAX_JOBS) {
        const int m = std::min(UPPER_MAX_JOBS, count - base);
        CommitJobs jobs;
        jobs.n = m;
        u32 ctas = 0;
        for (int i = 0; i < m; i++) {
            const Commit& cm = cms[base + i];
            REQUIRE(cm.cols == 1 && !cm.templates && cm.cta_cl == cm.cl, "internal: batched chunk commit needs single unlabeled columns");
            jobs.j[i].values = cm.values;
            jobs.j[i].upper = cm.upper;
            jobs.j[i].cl = cm.cl;
            jobs.cta0[i] = ctas;
            ctas += (u32)cm.n_ch;
        }
        chunk_commit_multi_kernel<<<ctas, HASH_THREADS, 0, ctx->stream>>>(jobs);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    }
}

// commit_finish for a batch of single-column commitments: all upper levels in at most three launches, roots to
// roots_dev[32 * index] (device).
void commit_finish_multi(sezkp_ctx* ctx, Commit* cms, int count, u8* roots_dev) {
    for (int base = 0; base < count; base += UPPER_MAX_JOBS) {
        const int m = std::min(UPPER_MAX_JOBS, count - base);
        for (int l0 = 0;; l0 += MAX_CL) {
            UpperJobs jobs;
            jobs.n = 0;
            u32 ctas = 0;
            for (int i = 0; i < m; i++) {
                const Commit& cm = cms[base + i];
                REQUIRE(cm.cols == 1, "internal: batched finish needs single-column commitments");
                const int depth = ilog2(cm.n_ch);
                if (l0 > depth || (l0 == depth && l0 > 0)) continue;  // done in an earlier round (depth 0: copy the root once)
                UpperJob& jb = jobs.j[jobs.n];
                jb.upper = cm.upper;
                jb.n_ch = cm.n_ch;
                jb.l0 = l0;
                jb.k = std::min(depth - l0, MAX_CL);
                jb.root_out = roots_dev ? (u32*)(roots_dev + 32 * (size_t)(base + i)) : nullptr;
                jb.grp0 = 0;
                jobs.cta0[jobs.n] = ctas;
                ctas += (u32)(cm.n_ch >> (l0 + jb.k));
                jobs.n++;
            }
            if (jobs.n == 0) break;
            upper_reduce_multi_kernel<<<ctas, HASH_THREADS, 0, ctx->stream>>>(jobs);
            CUDA_CHECK(cudaGetLastError());
            ctx->launches++;
        }
    }
}

void commit_finish_multi_range(sezkp_ctx* ctx, Commit* const* cms, int count, int rank, int world) {
    REQUIRE(count <= UPPER_MAX_JOBS && world >= 1 && (world & (world - 1)) == 0 && rank >= 0 && rank < world, "internal: bad range finish");
    const int wl = ilog2((u64)world);
    for (int l0 = 0;; l0 += MAX_CL) {
        UpperJobs jobs;
        jobs.n = 0;
        u32 ctas = 0;
        for (int i = 0; i < count; i++) {
            const Commit& cm = *cms[i];
            REQUIRE(cm.cols == 1 && cm.n_ch % (u64)world == 0, "internal: range finish needs single columns and world | n_ch");
            const int depth = ilog2(cm.n_ch) - wl;  // levels inside the own subtree
            if (l0 >= depth) continue;
            UpperJob& jb = jobs.j[jobs.n];
            jb.upper = cm.upper;
            jb.n_ch = cm.n_ch;
            jb.l0 = l0;
            jb.k = std::min(depth - l0, MAX_CL);
            jb.root_out = nullptr;
            const u64 groups = (cm.n_ch / (u64)world) >> (l0 + jb.k);  // of 2^k nodes at level l0, inside the own range
            jb.grp0 = groups * (u64)rank;
            jobs.cta0[jobs.n] = ctas;
            ctas += (u32)groups;
            jobs.n++;
        }
        if (jobs.n == 0) break;
        upper_reduce_multi_kernel<<<ctas, HASH_THREADS, 0, ctx->stream>>>(jobs);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    }
}
void commit_finish_multi_top(sezkp_ctx* ctx, Commit* const* cms, int count, int world, u8* const* roots_dev) {
    REQUIRE(count <= UPPER_MAX_JOBS && world >= 1 && (world & (world - 1)) == 0, "internal: bad top finish");
    const int wl = ilog2((u64)world);
    UpperJobs jobs;
    jobs.n = count;
    for (int i = 0; i < count; i++) {
        const Commit& cm = *cms[i];
        REQUIRE(cm.cols == 1 && cm.n_ch % (u64)world == 0 && wl <= MAX_CL, "internal: top finish needs single columns and world | n_ch");
        UpperJob& jb = jobs.j[i];
        jb.upper = cm.upper;
        jb.n_ch = cm.n_ch;
        jb.l0 = ilog2(cm.n_ch) - wl;
        jb.k = wl;  // one group of `world` nodes -> the root
        jb.root_out = roots_dev && roots_dev[i] ? (u32*)roots_dev[i] : nullptr;
        jb.grp0 = 0;
        jobs.cta0[i] = (u32)i;
    }
    upper_reduce_multi_kernel<<<count, HASH_THREADS, 0, ctx->stream>>>(jobs);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void commit_build(sezkp_ctx* ctx, Commit& cm, const u64* values_dev, u64 n, int cols, int chunk_log2,
                  const char* const* labels, const CommitOpts& opt) {
    commit_begin(ctx, cm, values_dev, n, cols, chunk_log2, labels, opt);
    commit_chunks(ctx, cm, 0, cm.n_ch, opt);
    commit_finish(ctx, cm, opt);
}

OpenReq make_open_req(const Commit& cm, u32 col, u64 row, u32 out_off) {
    REQUIRE(col < (u32)cm.cols, "opening: column %u out of range", col);
    REQUIRE(row < cm.n, "opening: row %llu out of range", (unsigned long long)row);
    OpenReq r;
    r.values = cm.values + (u64)col * cm.col_stride;
    r.upper = cm.upper + (u64)col * (2 * cm.n_ch - 1) * 8;
    r.tpl = cm.templates ? cm.templates + col : nullptr;
    r.n_ch = cm.n_ch;
    r.row = row;
    r.cl = (u32)cm.cl;
    r.depth_out = (u32)ilog2(cm.n_ch);
    r.out_off = out_off;
    r.pad = 0;
    return r;
}

// One launch for any mix of openings.  Results land in the context's pinned staging buffer: per request a value, a chunk
// root and, at out_off (32 B units), cl + depth_out path digests.  The returned pointers stay valid until the next
// open_batch_staged on this context.
void open_batch_staged(sezkp_ctx* ctx, const std::vector<OpenReq>& reqs, size_t path_digests, u64** values, u8** chunk_roots, u8** paths) {
    const size_t k = reqs.size();
    const size_t b_req = (k * sizeof(OpenReq) + 63) & ~(size_t)63, b_val = (k * 8 + 63) & ~(size_t)63, b_cr = k * 32, b_path = path_digests * 32;
    u8* d = (u8*)ctx->scratch[7].ensure(b_req + b_val + b_cr + b_path + 64);
    u8* h = (u8*)ctx->pinned[0].ensure(b_req + b_val + b_cr + b_path + 64);
    *values = (u64*)(h + b_req);
    *chunk_roots = h + b_req + b_val;
    *paths = h + b_req + b_val + b_cr;
    if (k == 0) return;
    std::memcpy(h, reqs.data(), k * sizeof(OpenReq));
    CUDA_CHECK(cudaMemcpyAsync(d, h, b_req, cudaMemcpyHostToDevice, ctx->stream));
    open_kernel<<<(unsigned)k, HASH_THREADS, 0, ctx->stream>>>((const OpenReq*)d, (u64*)(d + b_req), (u32*)(d + b_req + b_val),
                                                               (u32*)(d + b_req + b_val + b_cr));
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    CUDA_CHECK(cudaMemcpyAsync(h + b_req, d + b_req, b_val + b_cr + b_path, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
void open_batch(sezkp_ctx* ctx, const std::vector<OpenReq>& reqs, size_t path_digests, u64* values, u8* chunk_roots, u8* paths_host) {
    if (reqs.empty()) return;
    u64* v;
    u8 *cr, *pa;
    open_batch_staged(ctx, reqs, path_digests, &v, &cr, &pa);
    std::memcpy(values, v, reqs.size() * 8);
    std::memcpy(chunk_roots, cr, reqs.size() * 32);
    if (path_digests) std::memcpy(paths_host, pa, path_digests * 32);
}

void commit_open(sezkp_ctx* ctx, const Commit& cm, const u32* col_idx, const u64* row_idx, size_t k, u64* values,
                 u8* chunk_roots, u8* path_in, u8* path_to) {
    if (k == 0) return;
    const int din = cm.cl, dout = ilog2(cm.n_ch), depth = din + dout;
    std::vector<OpenReq> reqs(k);
    for (size_t i = 0; i < k; i++) reqs[i] = make_open_req(cm, col_idx[i], row_idx[i], (u32)(i * depth));
    std::vector<u8> paths(k * (size_t)depth * 32 + 32);
    open_batch(ctx, reqs, k * (size_t)depth, values, chunk_roots, paths.data());
    for (size_t i = 0; i < k; i++) {
        if (din) std::memcpy(path_in + i * (size_t)din * 32, &paths[i * (size_t)depth * 32], (size_t)din * 32);
        if (dout) std::memcpy(path_to + i * (size_t)dout * 32, &paths[(i * (size_t)depth + din) * 32], (size_t)dout * 32);
    }
}

void leaf_hash_device(sezkp_ctx* ctx, const u64* vals_dev, size_t n, const char* label, u32* out_dev) {
    if (n == 0) return;
    b3::LabelTemplate* d_t = nullptr;
    if (label) {
        b3::LabelTemplate t = make_label_template(label);
        d_t = (b3::LabelTemplate*)ctx->scratch[9].ensure(sizeof t);
        CUDA_CHECK(cudaMemcpyAsync(d_t, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    leaf_hash_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(vals_dev, n, d_t, out_dev);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void merkle_root_device(sezkp_ctx* ctx, u32* level, u32* tmp, size_t n, u8* root_host) {
    u32 *a = level, *b = tmp;
    while (n > 1) {
        const size_t n_out = (n + 1) / 2;
        merkle_level_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, ctx->stream>>>(a, n, b);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
        u32* t = a;
        a = b;
        b = t;
        n = n_out;
    }
    CUDA_CHECK(cudaMemcpyAsync(root_host, a, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

/* ------------------------------------------------------------------------------------------ */
/* batched path verification (a15: MerkleTree::verify v1/merkle.rs:111-126, verify_chunked_open :243-280) */
/* ------------------------------------------------------------------------------------------ */
namespace {
struct VerifyArgs {
    const u32* col_roots;              // [c][8]
    const b3::LabelTemplate* tpl;      // [c] or null (unlabeled leaves)
    const u32* col_idx;                // [k] or null (column 0)
    const u64* values;                 // [k]
    const u64* idx_in;                 // [k]
    const u64* idx_out;                // [k] (unused when dout == 0)
    const u32* chunk_roots;            // [k][8] or null: no intermediate check
    const u32* path_in;                // [k][din][8]
    const u32* path_to;                // [k][dout][8]
    int din, dout;
    size_t k;
    u8* ok;                            // [k]
};
__device__ __forceinline__ void walk_up(u32 (&cur)[8], u64 idx, const u32* __restrict__ sibs, int depth) {
    for (int l = 0; l < depth; l++) {
        u32 s[8], o[8];
#pragma unroll
        for (int w = 0; w < 8; w++) s[w] = sibs[l * 8 + w];
        if ((idx & 1) == 0) b3::parent(cur, s, o);
        else b3::parent(s, cur, o);
#pragma unroll
        for (int w = 0; w < 8; w++) cur[w] = o[w];
        idx >>= 1;
    }
}
__global__ void __launch_bounds__(128) verify_paths_kernel(const VerifyArgs a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.k) return;
    const u32 col = a.col_idx ? a.col_idx[i] : 0;
    const u64 v = a.values[i];
    u32 cur[8];
    if (a.tpl) {
        const b3::LabelTemplate t = a.tpl[col];
        B3_DISPATCH_LABELED(t, { b3::leaf_labeled_w<B3W>(t, v, cur); })
    } else b3::leaf(v, cur);
    walk_up(cur, a.idx_in[i], a.path_in + i * (size_t)a.din * 8, a.din);
    bool good = v < gl::P;  // a non-canonical value can never be an honest opening
    if (a.chunk_roots) {
#pragma unroll
        for (int w = 0; w < 8; w++) good = good && cur[w] == a.chunk_roots[i * 8 + w];
    }
    if (a.dout) walk_up(cur, a.idx_out[i], a.path_to + i * (size_t)a.dout * 8, a.dout);
#pragma unroll
    for (int w = 0; w < 8; w++) good = good && cur[w] == a.col_roots[(size_t)col * 8 + w];
    a.ok[i] = good ? 1 : 0;
}
}  // namespace

void verify_paths_device(sezkp_ctx* ctx, const u8* col_roots, const char* const* labels, int c, const u32* col_idx, const u64* values,
                         const u64* idx_in, const u64* idx_out, const u8* chunk_roots, const u8* path_in, int din, const u8* path_to,
                         int dout, size_t k, u8* ok_host) {
    if (k == 0) return;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_roots = 0, o_tpl = al(o_roots + (size_t)c * 32), o_col = al(o_tpl + (labels ? sizeof(b3::LabelTemplate) * c : 0)),
                 o_val = al(o_col + (col_idx ? k * 4 : 0)), o_ii = al(o_val + k * 8), o_io = al(o_ii + k * 8),
                 o_cr = al(o_io + (dout ? k * 8 : 0)), o_pi = al(o_cr + (chunk_roots ? k * 32 : 0)),
                 o_pt = al(o_pi + k * (size_t)din * 32), o_ok = al(o_pt + k * (size_t)dout * 32), total = al(o_ok + k);
    u8* h = (u8*)ctx->pinned[1].ensure(total);
    u8* d = (u8*)ctx->scratch[7].ensure(total);
    std::memcpy(h + o_roots, col_roots, (size_t)c * 32);
    if (labels)
        for (int j = 0; j < c; j++) {
            REQUIRE(labels[j] != nullptr, "label %d is NULL", j);
            const b3::LabelTemplate t = make_label_template(labels[j]);
            std::memcpy(h + o_tpl + sizeof t * j, &t, sizeof t);
        }
    if (col_idx) {
        for (size_t i = 0; i < k; i++) REQUIRE(col_idx[i] < (u32)c, "opening %zu: column index out of range", i);
        std::memcpy(h + o_col, col_idx, k * 4);
    }
    std::memcpy(h + o_val, values, k * 8);
    std::memcpy(h + o_ii, idx_in, k * 8);
    if (dout) std::memcpy(h + o_io, idx_out, k * 8);
    if (chunk_roots) std::memcpy(h + o_cr, chunk_roots, k * 32);
    if (din) std::memcpy(h + o_pi, path_in, k * (size_t)din * 32);
    if (dout) std::memcpy(h + o_pt, path_to, k * (size_t)dout * 32);
    CUDA_CHECK(cudaMemcpyAsync(d, h, o_ok, cudaMemcpyHostToDevice, ctx->stream));
    VerifyArgs a;
    a.col_roots = (const u32*)(d + o_roots);
    a.tpl = labels ? (const b3::LabelTemplate*)(d + o_tpl) : nullptr;
    a.col_idx = col_idx ? (const u32*)(d + o_col) : nullptr;
    a.values = (const u64*)(d + o_val);
    a.idx_in = (const u64*)(d + o_ii);
    a.idx_out = (const u64*)(d + o_io);
    a.chunk_roots = chunk_roots ? (const u32*)(d + o_cr) : nullptr;
    a.path_in = (const u32*)(d + o_pi);
    a.path_to = (const u32*)(d + o_pt);
    a.din = din;
    a.dout = dout;
    a.k = k;
    a.ok = d + o_ok;
    verify_paths_kernel<<<(unsigned)((k + 127) / 128), 128, 0, ctx->stream>>>(a);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    CUDA_CHECK(cudaMemcpyAsync(h + o_ok, d + o_ok, k, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(ok_host, h + o_ok, k);
}
