// STARK v1 device stages other than NTT and hashing, plus the host orchestration of prove_v1.
//   * column expansion from the compact trace     (reference v1/columns.rs:252-365, v1/openings.rs:193-273)
//   * AIR composition + boundary + ZK mask         (v1/air.rs:49-136, v1/masking.rs:86-103, v1/prover.rs:142-158)
//   * DEEP quotient with batched inversion         (v1/lde.rs:76-93)
//   * FRI fold                                     (v1/prover.rs:204-238)
//   * prove_v1 schedule, openings, bincode         (v1/prover.rs:61-462, v1/proof.rs, sezkp-stark/src/lib.rs:131)
#include <chrono>
#include <cstring>

#include "gl.cuh"
#include "group.cuh"
#include "hash.cuh"
#include "ntt.cuh"
#include "stark.cuh"

namespace {

/* ------------------------------------------------------------------------------------------ */
/* column expansion                                                                            */
/* ------------------------------------------------------------------------------------------ */
// One warp per (block, tape): running head position (post-move, relative to 0 at block entry).  Lanes take
// consecutive rows, so the prefix sum is a warp scan plus a carry and the stores are coalesced.
// f.col_mod > 1 (one proof sharded over several GPUs): only the columns c with c % col_mod == col_rem are produced for every
// row — the columns this rank commits — plus ALL columns for the rows [full_lo, full_hi) and the row `halo`: the row slice
// this rank composes (row i of the composition reads rows i and i + 1 mod n).
__global__ void __launch_bounds__(256) head_scan_kernel(DeviceTrace t, u64* __restrict__ cols, u64 blk0, u64 blk1, const ExpandFilter f) {
    const u64 id = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    if (id >= (blk1 - blk0) * t.tau) return;
    const u64 k = blk0 + id / t.tau;
    const u32 r = (u32)(id % t.tau);
    if (f.col_mod > 1 && (3 + 3 * t.tau + r) % f.col_mod != f.col_rem) {  // not an own column: only blocks that touch the row slice
        const u64 s0 = t.block_start[k], s1 = s0 + t.block_len[k];
        if (!((s0 < f.full_hi && s1 > f.full_lo) || (f.halo >= s0 && f.halo < s1))) return;
    }
    const u64 start = t.block_start[k], len = t.block_len[k];
    u64* head = cols + (3 + 3ULL * t.tau + r) * t.n_rows + start;  // group order: mv, wflag, wsym, head, ...
    const signed char* mv = (const signed char*)(t.ops ? (const int8_t*)t.ops : t.mv) + start * t.tau + r;
    const bool packed = t.ops != nullptr;
    int64_t carry = 0;
#pragma unroll 4
    for (u64 j0 = 0; j0 < len; j0 += 32) {
        const u64 j = j0 + lane;
        int v = j < len ? (packed ? (int)(((u8)mv[j * t.tau]) & 3) - 1 : (int)mv[j * t.tau]) : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if ((int)lane >= o) v += u;
        }
        if (j < len) head[j] = gl::from_i64(carry + v);
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
}
// One thread per row: everything except head.
__global__ void expand_rows_kernel(DeviceTrace t, u64* __restrict__ cols, u64 row0, u64 row1, const ExpandFilter f) {
    const u64 i = row0 + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row1) return;
    const bool all_cols = f.col_mod <= 1 || (i >= f.full_lo && i < f.full_hi) || i == f.halo;
    auto own = [&](u32 c) { return all_cols || c % f.col_mod == f.col_rem; };
    // block containing row i: largest k with block_start[k] <= i (blocks of length 0 are skipped by the search)
    u64 lo = 0, hi = t.n_blocks;
    while (hi - lo > 1) {
        const u64 mid = (lo + hi) >> 1;
        if (t.block_start[mid] <= i) lo = mid;
        else hi = mid;
    }
    const u64 k = lo;
    const u64 start = t.block_start[k], len = t.block_len[k];
    const u64 n = t.n_rows;
    const u32 tau = t.tau;
    if (own(0)) cols[0 * n + i] = gl::from_i64((int64_t)t.input_mv[i]);
    if (own(1)) cols[1 * n + i] = (i == start) ? 1 : 0;
    if (own(2)) cols[2 * n + i] = (i + 1 == start + len) ? 1 : 0;
    for (u32 r = 0; r < tau; r++) {
        const u64 p = i * tau + r;
        int wf, mvv;
        u64 sym;
        if (t.ops) {
            const u32 o = t.ops[p];
            mvv = (int)(o & 3) - 1;
            wf = (o >> 2) & 1;
            sym = o >> 3;
        } else {
            wf = t.write_flag[p] ? 1 : 0;
            mvv = (int)t.mv[p];
            sym = (u64)t.write_sym[p];
        }
        if (own(3 + 0 * tau + r)) cols[(3 + 0ULL * tau + r) * n + i] = gl::from_i64((int64_t)mvv);
        if (own(3 + 1 * tau + r)) cols[(3 + 1ULL * tau + r) * n + i] = (u64)wf;
        if (own(3 + 2 * tau + r)) cols[(3 + 2ULL * tau + r) * n + i] = wf ? sym : 0;
        const int64_t diff = t.win_right[k * tau + r] - t.win_left[k * tau + r];
        const u64 ad = diff < 0 ? (u64)0 - (u64)diff : (u64)diff;
        if (own(3 + 4 * tau + r)) cols[(3 + 4ULL * tau + r) * n + i] = gl::from_u64(ad + 1);
        if (own(3 + 5 * tau + r)) cols[(3 + 5ULL * tau + r) * n + i] = (u64)t.head_in_off[k * tau + r];
        if (own(3 + 6 * tau + r)) cols[(3 + 6ULL * tau + r) * n + i] = (u64)t.head_out_off[k * tau + r];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* composition                                                                                 */
/* ------------------------------------------------------------------------------------------ */
struct ComposeParams {
    u64 a[8];        // alphas as drawn (mapping of v1/prover.rs:86-98 applied in the kernel)
    u64 mask[8];     // ascending mask coefficients
    int mask_deg;
    u64 w_hi[64];    // w_n^(j << lo_bits): with w_lo (device table) w_n^i = w_hi[i >> lo_bits] * w_lo[i & (2^lo_bits - 1)]
    int lo_bits;
};
// One thread per row.  The composition is linear in the alphas, so the per-tape constraint values are summed per
// constraint kind first and the alphas (and the row-level selectors is_first / is_last / 1 - is_last) are applied
// once per row: 6 multiplications per tape + 13 per row instead of 19 per tape + a 22-bit power per row.  All sums are
// lazy representatives (gl::lazy); only values whose BITS matter (range terms) are canonical.
__global__ void __launch_bounds__(128) compose_kernel(const u64* __restrict__ cols, u64 n, u32 tau, ComposeParams cp,
                                                      const u64* __restrict__ w_lo, u64* __restrict__ out, u64 row0, u64 row1) {
    namespace L = gl::lazy;
    const u64 i = row0 + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row1) return;
    const u64 ip1 = (i + 1 == n) ? 0 : i + 1;
    const u64 is_first = cols[1 * n + i], is_last = cols[2 * n + i];
    u64 s_bool = 0, s_mv = 0, s_hu = 0, s_hr = 0, s_sr = 0, s_bf = 0, s_bl = 0;
    for (u32 r = 0; r < tau; r++) {
        const u64 mv = cols[(3 + 0ULL * tau + r) * n + i], mv_next = cols[(3 + 0ULL * tau + r) * n + ip1];
        const u64 flg = cols[(3 + 1ULL * tau + r) * n + i];
        const u64 sym = cols[(3 + 2ULL * tau + r) * n + i];
        const u64 head = cols[(3 + 3ULL * tau + r) * n + i], head_next = cols[(3 + 3ULL * tau + r) * n + ip1];
        const u64 wlen = cols[(3 + 4ULL * tau + r) * n + i];
        const u64 in_off = cols[(3 + 5ULL * tau + r) * n + i], out_off = cols[(3 + 6ULL * tau + r) * n + i];
        // C1, C2, C3 (v1/air.rs:64-72).  Selector-like factors (flags, is_first, is_last) are 0 or 1 on every well-formed
        // trace: then the product is a select; any other value takes the general multiplication (same result).
        if (flg > 1) s_bool = L::add(s_bool, L::mul(flg, gl::sub(flg, 1)));
        if (mv > 1 && mv != gl::P - 1) s_mv = L::add(s_mv, L::mul(L::mul(mv, gl::sub(mv, 1)), gl::add(mv, 1)));
        s_hu = L::add(s_hu, L::sub(L::sub(head_next, head), mv_next));
        // Bit columns are the honest decompositions of the canonical residues (v1/columns.rs:324-342), so every
        // b*(b-1) term is zero and the reconstructed sums are the low bits of the residue (v1/air.rs:74-112).
        const u64 slack = gl::sub(gl::sub(wlen, 1), head);
        // head range (alpha[4]); slack range (alpha[6]); symbol range shares alpha[0] with the booleanity term
        const u64 hr = head - (head & 0xFFFFULL), sr = slack - (slack & 0xFFFFULL), yr = sym - (sym & 0xFULL);
        if (flg == 1) {
            s_hr = L::add(s_hr, hr);
            s_sr = L::add(s_sr, sr);
            s_bool = L::add(s_bool, yr);
        } else if (flg != 0) {
            s_hr = L::add(s_hr, L::mul(flg, hr));
            s_sr = L::add(s_sr, L::mul(flg, sr));
            s_bool = L::add(s_bool, L::mul(flg, yr));
        }
        // boundary (v1/air.rs:116-136)
        s_bf = L::add(s_bf, L::sub(L::sub(head, mv), in_off));
        s_bl = L::add(s_bl, L::sub(head, out_off));
    }
    // alphas: bool = sym range = a[0], mv = a[1], head update = both boundaries = a[2], head range = a[4], slack = a[6]
    auto sel = [](u64 f, u64 x) { return f == 0 ? 0 : (f == 1 ? x : gl::lazy::mul(f, x)); };
    const u64 t2 = L::add(L::add(sel(gl::sub(1, is_last), s_hu), sel(is_first, s_bf)), sel(is_last, s_bl));
    u64 acc = L::mul(cp.a[0], s_bool);
    acc = L::add(acc, L::mul(cp.a[1], s_mv));
    acc = L::add(acc, L::mul(cp.a[2], t2));
    acc = L::add(acc, L::mul(cp.a[4], s_hr));
    acc = L::add(acc, L::mul(cp.a[6], s_sr));
    // mask R(w^i), Horner over ascending coefficients (v1/masking.rs:86-103)
    const u64 x = L::mul(cp.w_hi[i >> cp.lo_bits], w_lo[i & ((1ULL << cp.lo_bits) - 1)]);
    u64 m = 0;
    for (int j = cp.mask_deg - 1; j >= 0; j--) m = L::add(L::mul(m, x), cp.mask[j]);
    out[i] = L::canon(L::add(acc, m));
}

/* ------------------------------------------------------------------------------------------ */
/* DEEP quotient: y[i] *= (shift*w^i - z)^-1, one field inversion per 4096 elements              */
/* ------------------------------------------------------------------------------------------ */
// Montgomery batch inversion at CTA scope: every thread multiplies up its 16 denominators, warp shuffles and one
// shared-memory step give each thread the product of all OTHER threads' denominators, one thread inverts the CTA
// total (Fermat, 127 multiplications — the reference pays that per element, v1/lde.rs:84), and the per-element
// inverses are peeled off backwards.  6.3 multiplications per element; inverses are unique, so the canonical
// results equal the reference's.
constexpr int DEEP_PER_THREAD = 16;
constexpr int DEEP_THREADS = 128;
struct DeepParams {
    u64 shift, z;
    u64 w_cta;        // w^(DEEP_THREADS * DEEP_PER_THREAD)
    u64 w_k[DEEP_PER_THREAD];       // w^(32 * k)
    u64 w_warp[DEEP_THREADS / 32];  // w^(32 * DEEP_PER_THREAD * j)
    u64 w_lane[32];   // w^l
};
__global__ void __launch_bounds__(DEEP_THREADS, 4) deep_kernel(u64* __restrict__ y, u64 N, const DeepParams dp, const u64* __restrict__ x0_table) {
    namespace L = gl::lazy;
    __shared__ u64 s_inv, s_wtot[DEEP_THREADS / 32];
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 cta_x0 = x0_table[blockIdx.x];  // shift * w^(first index of this CTA)
    // thread handles i = i0 + 32*k, k < DEEP_PER_THREAD (coalesced).  Everything below is arranged as shallow trees
    // of independent multiplications (groups of 4) rather than one running product: the kernel is latency-bound.
    const u64 i0 = (u64)blockIdx.x * (DEEP_THREADS * DEEP_PER_THREAD) + warp * (32 * DEEP_PER_THREAD) + lane;
    const u64 x0 = L::mul(L::mul(cta_x0, dp.w_warp[warp]), dp.w_lane[lane]);
    u64 den[DEEP_PER_THREAD], yv[DEEP_PER_THREAD];
#pragma unroll
    for (int k = 0; k < DEEP_PER_THREAD; k++) {
        const bool ok = i0 + 32ULL * k < N;
        den[k] = ok ? L::sub(k ? L::mul(x0, dp.w_k[k]) : x0, dp.z) : 1;
        yv[k] = ok ? y[i0 + 32ULL * k] : 0;
    }
    u64 ab[4], cd[4], g[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        ab[j] = L::mul(den[4 * j], den[4 * j + 1]);
        cd[j] = L::mul(den[4 * j + 2], den[4 * j + 3]);
        g[j] = L::mul(ab[j], cd[j]);
    }
    const u64 g01 = L::mul(g[0], g[1]), g23 = L::mul(g[2], g[3]);
    const u64 run = L::mul(g01, g23);
    // products of the other lanes of the warp: exclusive prefix * exclusive suffix
    u64 pf = run, sf = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 u = __shfl_up_sync(0xffffffffu, pf, o), d = __shfl_down_sync(0xffffffffu, sf, o);
        if ((int)lane >= o) pf = L::mul(pf, u);
        if ((int)lane + o < 32) sf = L::mul(sf, d);
    }
    if (lane == 31) s_wtot[warp] = pf;  // warp total
    u64 others = 1;
    {
        const u64 pe = __shfl_up_sync(0xffffffffu, pf, 1), se = __shfl_down_sync(0xffffffffu, sf, 1);
        if (lane > 0) others = pe;
        if (lane < 31) others = L::mul(others, se);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < DEEP_THREADS / 32; j++)
        if (j != (int)warp) others = L::mul(others, s_wtot[j]);
    if (tid == 0) {  // CTA total = others * run of thread 0; Fermat inverse, square and multiply chains interleaved
        const u64 total = L::mul(others, run);
        u64 acc = 1, base = total;
        for (u64 e = gl::P - 2; e; e >>= 1) {
            if (e & 1) acc = L::mul(acc, base);
            base = L::mul(base, base);
        }
        s_inv = acc;
    }
    __syncthreads();
    const u64 inv = L::mul(s_inv, others);  // 1 / run
    const u64 ig[4] = {L::mul(inv, L::mul(g[1], g23)), L::mul(inv, L::mul(g[0], g23)), L::mul(inv, L::mul(g01, g[3])),
                       L::mul(inv, L::mul(g01, g[2]))};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const u64 iab = L::mul(ig[j], cd[j]), icd = L::mul(ig[j], ab[j]);
        const u64 di[4] = {L::mul(iab, den[4 * j + 1]), L::mul(iab, den[4 * j]), L::mul(icd, den[4 * j + 3]), L::mul(icd, den[4 * j + 2])};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const u64 i = i0 + 32ULL * (4 * j + t);
            if (i < N) y[i] = L::canon(L::mul(yv[4 * j + t], di[t]));
        }
    }
}

// The same computation split in three launches so that no CTA waits for a serial inversion: in deep_kernel one thread
// runs the 64-deep squaring chain of the Fermat inverse while the other 127 threads of the CTA sit at the barrier (ncu:
// ALU pipe 39 %, the chain is more than half of a CTA's lifetime).  Here
//   deep_forward_kernel  computes, per thread, the product of all OTHER threads' denominators of its CTA (`others`) and,
//                        per CTA, the product of all its denominators (`total`);
//   deep_invert_kernel   inverts all CTA totals at once — one Fermat chain per LANE instead of one per CTA;
//   deep_apply_kernel    recomputes the thread's denominators and their product tree (cheaper than storing them) and peels
//                        the element inverses off inv(total) * others.
// 9.1 instead of 6.3 multiplications per element, but all of them at full occupancy.  Results are the unique inverses, so
// the outputs are bit-identical to deep_kernel's.
__device__ __forceinline__ void deep_thread_products(const DeepParams& dp, const u64* __restrict__ x0_table, u64 N, u64& i0,
                                                     u64 (&den)[DEEP_PER_THREAD], u64 (&ab)[4], u64 (&cd)[4], u64 (&g)[4], u64& g01,
                                                     u64& g23, u64& run) {
    namespace L = gl::lazy;
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 cta_x0 = x0_table[blockIdx.x];
    i0 = (u64)blockIdx.x * (DEEP_THREADS * DEEP_PER_THREAD) + warp * (32 * DEEP_PER_THREAD) + lane;
    const u64 x0 = L::mul(L::mul(cta_x0, dp.w_warp[warp]), dp.w_lane[lane]);
#pragma unroll
    for (int k = 0; k < DEEP_PER_THREAD; k++) {
        const bool ok = i0 + 32ULL * k < N;
        den[k] = ok ? L::sub(k ? L::mul(x0, dp.w_k[k]) : x0, dp.z) : 1;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        ab[j] = L::mul(den[4 * j], den[4 * j + 1]);
        cd[j] = L::mul(den[4 * j + 2], den[4 * j + 3]);
        g[j] = L::mul(ab[j], cd[j]);
    }
    g01 = L::mul(g[0], g[1]);
    g23 = L::mul(g[2], g[3]);
    run = L::mul(g01, g23);
}
__global__ void __launch_bounds__(DEEP_THREADS, 6) deep_forward_kernel(u64 N, const DeepParams dp, const u64* __restrict__ x0_table,
                                                                       u64* __restrict__ others_out, u64* __restrict__ totals) {
    namespace L = gl::lazy;
    __shared__ u64 s_wtot[DEEP_THREADS / 32];
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 i0, den[DEEP_PER_THREAD], ab[4], cd[4], g[4], g01, g23, run;
    deep_thread_products(dp, x0_table, N, i0, den, ab, cd, g, g01, g23, run);
    u64 pf = run, sf = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 u = __shfl_up_sync(0xffffffffu, pf, o), d = __shfl_down_sync(0xffffffffu, sf, o);
        if ((int)lane >= o) pf = L::mul(pf, u);
        if ((int)lane + o < 32) sf = L::mul(sf, d);
    }
    if (lane == 31) s_wtot[warp] = pf;
    u64 others = 1;
    {
        const u64 pe = __shfl_up_sync(0xffffffffu, pf, 1), se = __shfl_down_sync(0xffffffffu, sf, 1);
        if (lane > 0) others = pe;
        if (lane < 31) others = L::mul(others, se);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < DEEP_THREADS / 32; j++)
        if (j != (int)warp) others = L::mul(others, s_wtot[j]);
    others_out[(u64)blockIdx.x * DEEP_THREADS + tid] = others;
    if (tid == 0) totals[blockIdx.x] = L::mul(others, run);
}
__global__ void __launch_bounds__(128) deep_invert_kernel(u64* __restrict__ totals, u64 count) {
    namespace L = gl::lazy;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 acc = 1, base = totals[i];
    for (u64 e = gl::P - 2; e; e >>= 1) {
        if (e & 1) acc = L::mul(acc, base);
        base = L::mul(base, base);
    }
    totals[i] = acc;
}
__global__ void __launch_bounds__(DEEP_THREADS, 5) deep_apply_kernel(u64* __restrict__ y, u64 N, const DeepParams dp,
                                                                     const u64* __restrict__ x0_table, const u64* __restrict__ others_in,
                                                                     const u64* __restrict__ inv_totals) {
    namespace L = gl::lazy;
    u64 i0, den[DEEP_PER_THREAD], ab[4], cd[4], g[4], g01, g23, run;
    u64 yv[DEEP_PER_THREAD];
    {
        const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const u64 j0 = (u64)blockIdx.x * (DEEP_THREADS * DEEP_PER_THREAD) + warp * (32 * DEEP_PER_THREAD) + lane;
#pragma unroll
        for (int k = 0; k < DEEP_PER_THREAD; k++) yv[k] = j0 + 32ULL * k < N ? y[j0 + 32ULL * k] : 0;
    }
    const u64 others = others_in[(u64)blockIdx.x * DEEP_THREADS + threadIdx.x];
    const u64 inv_total = inv_totals[blockIdx.x];
    deep_thread_products(dp, x0_table, N, i0, den, ab, cd, g, g01, g23, run);
    const u64 inv = L::mul(inv_total, others);  // 1 / run
    const u64 ig[4] = {L::mul(inv, L::mul(g[1], g23)), L::mul(inv, L::mul(g[0], g23)), L::mul(inv, L::mul(g01, g[3])),
                       L::mul(inv, L::mul(g01, g[2]))};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const u64 iab = L::mul(ig[j], cd[j]), icd = L::mul(ig[j], ab[j]);
        const u64 di[4] = {L::mul(iab, den[4 * j + 1]), L::mul(iab, den[4 * j]), L::mul(icd, den[4 * j + 3]), L::mul(icd, den[4 * j + 2])};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const u64 i = i0 + 32ULL * (4 * j + t);
            if (i < N) y[i] = L::canon(L::mul(yv[4 * j + t], di[t]));
        }
    }
}

__global__ void fri_fold_kernel(const u64* __restrict__ in, u64 half, u64 beta, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    out[i] = gl::add(in[i], gl::mul(beta, in[i + half]));
}
// ---- coset-resident FRI layers (one proof over the GPUs of a context group, see fri_commit_device) ----
// The extension domain is the union of the 8 cosets g_j<w_n>; rank r evaluated cosets [r*per, (r+1)*per) and keeps every
// large FRI layer in that coset-major form: local[q][i] = layer[(r*per + q) + 8*i].  The fold pairs (k, k + len/2) have the
// same residue mod 8, so folding needs no communication: it is the same fold on every local coset array.
__global__ void __launch_bounds__(256) fri_fold_cosets_kernel(const u64* __restrict__ in, u64 m, u64 beta, u64* __restrict__ out) {
    const u64 half = m >> 1;  // m = elements per coset of the input layer
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const u64* src = in + (u64)blockIdx.y * m;
    out[(u64)blockIdx.y * half + i] = gl::add(src[i], gl::mul(beta, src[i + half]));
}
// Hashing works on contiguous index ranges.  This kernel builds, for up to FRI_GATHER_MAX_JOBS layers at once, the natural-
// order values of a range [8*i0, 8*(i0 + cnt)) from the eight coset arrays, reading the slices of the seven remote cosets
// straight out of the peers' HBM (peer loads over NVLink, 2 KiB contiguous per coset and CTA) and writing 16 KiB of
// consecutive outputs per CTA through a shared-memory transpose.  One launch replaces world*per peer copies + an interleave
// pass per layer; only 1/world of every layer crosses NVLink per rank (the all-gather it replaces moved all of it).
constexpr int FRI_GATHER_MAX_JOBS = 14, FRI_GATHER_MAX_WORLD = 8;  // log_N <= 32: at most 13 layers down to the first small one
struct FriGatherJob {
    u64 src_off;  // element offset of this layer's local array in every rank's buffer
    u64 m;        // elements per coset (layer length / 8)
    u64 i0, cnt;  // coset-index range; cnt is a multiple of 256
    u64* dst;     // natural-order output: dst[8*(i - i0) + j]
    u32 blk0;     // first CTA of this job
};
struct FriGatherJobs {
    const u64* peer[FRI_GATHER_MAX_WORLD];
    int per, n;
    FriGatherJob j[FRI_GATHER_MAX_JOBS];
};
__global__ void __launch_bounds__(256) fri_gather_ranges_kernel(const FriGatherJobs jobs) {
    __shared__ u64 s[8][257];
    int k = 0;
    while (k + 1 < jobs.n && blockIdx.x >= jobs.j[k + 1].blk0) k++;
    const FriGatherJob& jb = jobs.j[k];
    const u64 ib = (u64)(blockIdx.x - jb.blk0) * 256;
    const int t = threadIdx.x;
#pragma unroll
    for (int c = 0; c < 8; c++) {  // coset c = local coset c % per of rank c / per
        const u64* src = jobs.peer[c / jobs.per] + jb.src_off + (u64)(c % jobs.per) * jb.m + jb.i0 + ib;
        s[c][t] = src[t];
    }
    __syncthreads();
    u64* dst = jb.dst + ib * 8;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int o = t + 256 * w;
        dst[o] = s[o & 7][o >> 3];
    }
}
// All folds from a layer of 2*len0 values down to the single final value in one CTA (the layers are contiguous in
// `buf`: layer of 2*len0 at buf, its fold of len0 behind it, and so on).  betas[k] folds the k-th of these layers.
constexpr int FRI_TAIL_MAX_LAYERS = 16;
struct TailBetas { u64 b[FRI_TAIL_MAX_LAYERS]; };
__global__ void __launch_bounds__(1024) fri_tail_fold_kernel(u64* __restrict__ buf, u64 len0, int layers, const TailBetas betas) {
    u64* in = buf;
    u64 len = len0;
    for (int k = 0; k < layers; k++) {
        u64* out = in + 2 * len;
        const u64 beta = betas.b[k];
        for (u64 i = threadIdx.x; i < len; i += blockDim.x) out[i] = gl::add(in[i], gl::mul(beta, in[i + len]));
        __syncthreads();  // writes of this CTA are visible to its own threads after the barrier
        in = out;
        len >>= 1;
    }
}

// Scatter of a packed all-gather of FRI subtree roots: recv is [world][total_words]; inside a rank's block, layer i occupies
// own_words[i] words at off_words[i] and belongs at upper[i] + rank * own_words[i] (equal chunk ranges per rank).
struct GatherScatter {
    static constexpr int MAX = 12;
    u32* upper[MAX];
    u32 own_words[MAX], off_words[MAX];
    u32 total_words;
    int n, world;
    const u32* recv;
};
__global__ void __launch_bounds__(256) gather_scatter_kernel(const GatherScatter gs) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte quad per thread (digests are 32 B: everything is aligned)
    const u64 quads_per_rank = gs.total_words / 4;
    if (q >= quads_per_rank * gs.world) return;
    const u32 r = (u32)(q / quads_per_rank), w = (u32)(q % quads_per_rank) * 4;
    int i = 0;
    while (i + 1 < gs.n && w >= gs.off_words[i + 1]) i++;
    const uint4 v = *(const uint4*)(gs.recv + (u64)r * gs.total_words + w);
    *(uint4*)(gs.upper[i] + (u64)r * gs.own_words[i] + (w - gs.off_words[i])) = v;
}

inline unsigned blocks_for(u64 n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

/* ------------------------------------------------------------------------------------------ */
/* device trace                                                                                */
/* ------------------------------------------------------------------------------------------ */
void validate_trace(const sezkp_trace_desc* d) {
    REQUIRE(d != nullptr, "trace descriptor is NULL");
    REQUIRE((d->flags & ~SEZKP_TRACE_PACKED_OPS) == 0, "trace descriptor: unknown flags 0x%x", d->flags);
    REQUIRE(d->tau >= 1 && d->tau <= 4096, "tau %u out of range", d->tau);
    REQUIRE(d->n_blocks >= 1, "empty trace");
    REQUIRE(d->n_rows >= 2 && (d->n_rows & (d->n_rows - 1)) == 0,
            "n_rows = %llu: the STARK v1 path needs a power-of-two trace length >= 2 (reference v1/lde.rs:51)",
            (unsigned long long)d->n_rows);
    REQUIRE(d->n_rows <= (1ULL << 29), "n_rows too large (LDE domain limited to 2^32)");
    REQUIRE(d->block_len && d->win_left && d->win_right && d->head_in_off && d->head_out_off && d->input_mv && d->mv &&
                ((d->flags & SEZKP_TRACE_PACKED_OPS) || (d->write_flag && d->write_sym)),
            "trace descriptor has NULL arrays");
    u64 sum = 0;
    for (u64 k = 0; k < d->n_blocks; k++) {
        REQUIRE(d->block_len[k] >= 1, "block %llu is empty", (unsigned long long)k);
        sum += d->block_len[k];
    }
    REQUIRE(sum == d->n_rows, "n_rows != sum(block_len)");
}

// One packed allocation, 16-byte aligned sections: per-block metadata first, then the row arrays.
void DeviceTraceOwner::layout(const sezkp_trace_desc* d) {
    const u64 nb = d->n_blocks, n = d->n_rows, tau = d->tau;
    size_t off = 0;
    auto sect = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 15) & ~(size_t)15;
        return o;
    };
    o_start = sect(nb * 8); o_len = sect(nb * 8); o_wl = sect(nb * tau * 8); o_wr = sect(nb * tau * 8);
    packed = (d->flags & SEZKP_TRACE_PACKED_OPS) != 0;
    o_io = sect(nb * tau * 4); o_oo = sect(nb * tau * 4); o_imv = sect(n); o_mv = sect(n * tau);
    o_wf = sect(packed ? 0 : n * tau); o_ws = sect(packed ? 0 : n * tau * 2);
    total = off;
}
void DeviceTraceOwner::upload_meta(sezkp_ctx* ctx, const sezkp_trace_desc* d) {
    const u64 nb = d->n_blocks, n = d->n_rows, tau = d->tau;
    layout(d);
    std::vector<u64> starts(nb);
    u64 acc = 0;
    for (u64 k = 0; k < nb; k++) {
        starts[k] = acc;
        acc += d->block_len[k];
    }
    u8* base = (u8*)buf.ensure(total);
    auto put = [&](size_t o, const void* src, size_t bytes) {
        CUDA_CHECK(cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    };
    put(o_start, starts.data(), nb * 8);
    put(o_len, d->block_len, nb * 8);
    put(o_wl, d->win_left, nb * tau * 8);
    put(o_wr, d->win_right, nb * tau * 8);
    put(o_io, d->head_in_off, nb * tau * 4);
    put(o_oo, d->head_out_off, nb * tau * 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // `starts` is a local
    t.tau = (u32)tau;
    t.n_blocks = nb;
    t.n_rows = n;
    t.block_start = (const u64*)(base + o_start);
    t.block_len = (const u64*)(base + o_len);
    t.win_left = (const int64_t*)(base + o_wl);
    t.win_right = (const int64_t*)(base + o_wr);
    t.head_in_off = (const u32*)(base + o_io);
    t.head_out_off = (const u32*)(base + o_oo);
    t.input_mv = (const int8_t*)(base + o_imv);
    t.mv = packed ? nullptr : (const int8_t*)(base + o_mv);
    t.write_flag = packed ? nullptr : (const u8*)(base + o_wf);
    t.write_sym = packed ? nullptr : (const uint16_t*)(base + o_ws);
    t.ops = packed ? (const u8*)(base + o_mv) : nullptr;
    h2d_bytes = total;
}
void DeviceTraceOwner::upload_rows_async(cudaStream_t stream, const sezkp_trace_desc* d, u64 r0, u64 r1) {
    const u64 tau = d->tau, r = r1 - r0;
    u8* base = (u8*)buf.p;
    CUDA_CHECK(cudaMemcpyAsync(base + o_imv + r0, d->input_mv + r0, r, cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaMemcpyAsync(base + o_mv + r0 * tau, d->mv + r0 * tau, r * tau, cudaMemcpyHostToDevice, stream));
    if (packed) return;
    CUDA_CHECK(cudaMemcpyAsync(base + o_wf + r0 * tau, d->write_flag + r0 * tau, r * tau, cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaMemcpyAsync(base + o_ws + r0 * tau * 2, d->write_sym + r0 * tau, r * tau * 2, cudaMemcpyHostToDevice, stream));
}
void DeviceTraceOwner::upload(sezkp_ctx* ctx, const sezkp_trace_desc* d) {
    upload_meta(ctx, d);
    upload_rows_async(ctx->stream, d, 0, d->n_rows);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// rows [row0,row1) of every non-head column and the head columns of blocks [blk0,blk1)
void expand_columns_range(sezkp_ctx* ctx, const DeviceTrace& t, u64* cols, u64 row0, u64 row1, u64 blk0, u64 blk1, const ExpandFilter& f) {
    if (blk1 > blk0) {
        head_scan_kernel<<<blocks_for((blk1 - blk0) * t.tau * 32, 256), 256, 0, ctx->stream>>>(t, cols, blk0, blk1, f);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    }
    if (row1 > row0) {
        expand_rows_kernel<<<blocks_for(row1 - row0, 256), 256, 0, ctx->stream>>>(t, cols, row0, row1, f);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    }
}
void expand_columns_device(sezkp_ctx* ctx, const DeviceTrace& t, u64* cols) {
    expand_columns_range(ctx, t, cols, 0, t.n_rows, 0, t.n_blocks);
}

// w_n^j for j < 2^lo_bits (n = 2^L), cached per L in the context (device, read-only after the first call).
static const u64* power_table_device(sezkp_ctx* ctx, int L, int lo_bits) {
    auto it = ctx->power_tables.find(L);
    if (it != ctx->power_tables.end()) return it->second;
    const u64 w = gl::root_2exp((unsigned)L);
    std::vector<u64> h((size_t)1 << lo_bits);
    u64 x = 1;
    for (auto& e : h) {
        e = x;
        x = gl::mul(x, w);
    }
    u64* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, h.size() * 8));
    upload_table(ctx, d, h.data(), h.size() * 8);
    ctx->power_tables[L] = d;
    return d;
}

void compose_device(sezkp_ctx* ctx, const u64* cols, u64 n, u32 tau, const u64 alphas8[8], const u64* mask, size_t mask_deg,
                    u64* out, u64 row0, u64 row1) {
    if (row1 > n) row1 = n;  // default arguments: the whole base domain
    REQUIRE(mask_deg <= 8, "mask degree %zu > 8 unsupported", mask_deg);
    ComposeParams cp{};
    for (int i = 0; i < 8; i++) {
        REQUIRE(alphas8[i] < gl::P, "alpha %d is not a canonical field element", i);
        cp.a[i] = alphas8[i];
    }
    for (size_t i = 0; i < mask_deg; i++) {
        REQUIRE(mask[i] < gl::P, "mask coefficient %zu is not canonical", i);
        cp.mask[i] = mask[i];
    }
    cp.mask_deg = (int)mask_deg;
    // w_n^i = w_hi[i >> lo_bits] * w_lo[i & mask]: 64 high powers in the launch parameters, the low table on the device
    const int L = ilog2(n);
    cp.lo_bits = L > 6 ? L - 6 : 0;
    const u64 w = gl::root_2exp((unsigned)L);
    const u64 w_lo_step = gl::pow(w, 1ULL << cp.lo_bits);
    u64 x = 1;
    for (u64 j = 0; j < (n >> cp.lo_bits); j++) {
        cp.w_hi[j] = x;
        x = gl::mul(x, w_lo_step);
    }
    const u64* w_lo = power_table_device(ctx, L, cp.lo_bits);
    if (row1 <= row0) return;
    compose_kernel<<<blocks_for(row1 - row0, 128), 128, 0, ctx->stream>>>(cols, n, tau, cp, w_lo, out, row0, row1);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

bool z_on_coset(u64 z, u64 shift, int log_N) {  // v1/prover.rs:120-131
    u64 t = gl::mul(z, gl::inv(shift));
    for (int i = 0; i < log_N; i++) t = gl::sqr(t);
    return t == 1;
}

// base evaluations (device, n = 2^L; destroyed) -> out (device, 2^(L+logB)): iNTT, coset LDE, DEEP quotient.
void deep_lde_device(sezkp_ctx* ctx, u64* base_vals, u64* out, int L, int logB, u64 shift, u64 z) {
    REQUIRE(z < gl::P && shift < gl::P && shift != 0, "shift / z must be canonical, shift non-zero");
    REQUIRE(!z_on_coset(z, shift, L + logB), "OOD point z lies on the evaluation coset");
    const u64 n = 1ULL << L, N = n << logB;
    u64* tmp = (u64*)ctx->scratch[0].ensure(n * 8);
    ntt_batch_device(ctx, base_vals, tmp, L, 1, true);
    u64* inter = (u64*)ctx->scratch[1].ensure(N * 8);
    coset_lde_device(ctx, base_vals, out, inter, L, logB, shift, 1);
    deep_quotient_device(ctx, out, L + logB, shift, z);
}

// y[i] *= (shift * w^i - z)^-1 for i < 2^log_dom, w = the primitive 2^log_dom-th root (v1/lde.rs:76-93, batched inversion)
void deep_quotient_device(sezkp_ctx* ctx, u64* out, int log_dom, u64 shift, u64 z) {
    const u64 N = 1ULL << log_dom;
    const u64 w = gl::root_2exp((unsigned)log_dom);
    DeepParams dp;
    dp.shift = shift;
    dp.z = z;
    dp.w_cta = gl::pow(w, (u64)DEEP_THREADS * DEEP_PER_THREAD);
    for (int k = 0; k < DEEP_PER_THREAD; k++) dp.w_k[k] = gl::pow(w, 32ULL * k);
    for (int j = 0; j < DEEP_THREADS / 32; j++) dp.w_warp[j] = gl::pow(w, 32ULL * DEEP_PER_THREAD * j);
    for (int l = 0; l < 32; l++) dp.w_lane[l] = gl::pow(w, (u64)l);
    // shift * w^(cta * elements per CTA), cached per (log N, shift)
    const u64 n_cta = blocks_for(N, DEEP_THREADS * DEEP_PER_THREAD);
    const auto key = std::make_pair(log_dom, shift);  // n_cta is a function of log N, so the entry's length is implied
    if (!ctx->deep_tables.count(key) && ctx->deep_tables.size() >= 32) {  // callers that vary `shift`: bounded cache
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (auto& kv : ctx->deep_tables) cudaFree(kv.second);
        ctx->deep_tables.clear();
    }
    u64*& x0_table = ctx->deep_tables[key];
    if (!x0_table) {
        std::vector<u64> h(n_cta);
        u64 x = shift;
        for (auto& e : h) {
            e = x;
            x = gl::mul(x, dp.w_cta);
        }
        CUDA_CHECK(cudaMalloc(&x0_table, n_cta * 8));
        upload_table(ctx, x0_table, h.data(), n_cta * 8);
    }
    if (n_cta >= 1024 && !ctx->deep_fused) {  // large: three launches, no CTA waits for a serial inversion
        u64* others = (u64*)ctx->scratch[6].ensure((n_cta * DEEP_THREADS + n_cta) * 8);
        u64* totals = others + n_cta * DEEP_THREADS;
        deep_forward_kernel<<<(unsigned)n_cta, DEEP_THREADS, 0, ctx->stream>>>(N, dp, x0_table, others, totals);
        deep_invert_kernel<<<(unsigned)blocks_for(n_cta, 128), 128, 0, ctx->stream>>>(totals, n_cta);
        deep_apply_kernel<<<(unsigned)n_cta, DEEP_THREADS, 0, ctx->stream>>>(out, N, dp, x0_table, others, totals);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches += 3;
    } else {
        deep_kernel<<<(unsigned)n_cta, DEEP_THREADS, 0, ctx->stream>>>(out, N, dp, x0_table);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    }
}

namespace {
// out[j + B*i] = gathered[j*n + i]: the coset-major result of the sharded LDE back into the reference's natural order
__global__ void __launch_bounds__(256) coset_interleave_kernel(const u64* __restrict__ gathered, u64 n, int logB, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int B = 1 << logB;
    for (int j = 0; j < B; j++) out[((u64)i << logB) + j] = gathered[(u64)j * n + i];
}
}  // namespace

// The same DEEP-LDE with the B cosets of the extension domain split over the ranks of a sharded proof: the evaluation domain
// shift*<w_N> is the union of the cosets g_j*<w_n>, g_j = shift*w_N^j, and out[j + B*i] = f(g_j w_n^i) / (g_j w_n^i - z) — a
// plain coset evaluation of size n with shift g_j followed by the quotient on that coset.  Rank r evaluates cosets
// [r*B/world, (r+1)*B/world), the coset-major pieces are all-gathered on the device and interleaved locally.  Values are
// the unique canonical results, so the output is bit-identical to deep_lde_device's.  Needs world | B.
void deep_lde_sharded_device(sezkp_ctx* ctx, u64* base_vals, u64* out, int L, int logB, u64 shift, u64 z, int rank, int world) {
    REQUIRE(z < gl::P && shift < gl::P && shift != 0, "shift / z must be canonical, shift non-zero");
    REQUIRE(!z_on_coset(z, shift, L + logB), "OOD point z lies on the evaluation coset");
    const int B = 1 << logB, per = B / world;
    REQUIRE(world >= 1 && per >= 1 && per * world == B && ctx->allgather_dev, "internal: coset sharding needs world | blow-up and a device collective");
    const u64 n = 1ULL << L;
    u64* tmp = (u64*)ctx->scratch[0].ensure(n * 8);
    ntt_batch_device(ctx, base_vals, tmp, L, 1, true);
    u64* inter = (u64*)ctx->scratch[1].ensure(n * 8);
    u64* mine = (u64*)ctx->pool.alloc((size_t)per * n * 8);
    u64* gathered = (u64*)ctx->pool.alloc((size_t)B * n * 8);
    try {
        const u64 wN = gl::root_2exp((unsigned)(L + logB));
        for (int q = 0; q < per; q++) {
            const u64 gj = gl::mul(shift, gl::pow(wN, (u64)(rank * per + q)));
            coset_lde_device(ctx, base_vals, mine + (u64)q * n, inter, L, 0, gj, 1);
            deep_quotient_device(ctx, mine + (u64)q * n, L, gj, z);
        }
        const int32_t rc = ctx->allgather_dev(ctx->allgather_dev_user, mine, (size_t)per * n * 8, gathered, (void*)ctx->stream);
        if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "device allgather callback failed with status %d", rc);
        coset_interleave_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(gathered, n, logB, out);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    } catch (...) {
        ctx->pool.free(mine);
        ctx->pool.free(gathered);
        throw;
    }
    ctx->pool.free(mine);
    ctx->pool.free(gathered);
}

// The rank's own cosets only, coset-major and left where they are: local[q][i] = f(g w_n^i) / (g w_n^i - z) with
// g = shift * w_N^(rank*per + q).  This is layer 0 of the coset-resident FRI (fri_commit_device): no all-gather, no
// interleave.  The caller owns the returned pool allocation (per * n elements).
u64* deep_lde_coset_local_device(sezkp_ctx* ctx, u64* base_vals, int L, int logB, u64 shift, u64 z, int rank, int world) {
    REQUIRE(z < gl::P && shift < gl::P && shift != 0, "shift / z must be canonical, shift non-zero");
    REQUIRE(!z_on_coset(z, shift, L + logB), "OOD point z lies on the evaluation coset");
    const int B = 1 << logB, per = B / world;
    REQUIRE(world >= 1 && per >= 1 && per * world == B, "internal: coset sharding needs world | blow-up");
    const u64 n = 1ULL << L;
    u64* tmp = (u64*)ctx->scratch[0].ensure(n * 8);
    ntt_batch_device(ctx, base_vals, tmp, L, 1, true);
    u64* inter = (u64*)ctx->scratch[1].ensure(n * 8);
    u64* mine = (u64*)ctx->pool.alloc((size_t)per * n * 8);
    try {
        const u64 wN = gl::root_2exp((unsigned)(L + logB));
        for (int q = 0; q < per; q++) {
            const u64 gj = gl::mul(shift, gl::pow(wN, (u64)(rank * per + q)));
            coset_lde_device(ctx, base_vals, mine + (u64)q * n, inter, L, 0, gj, 1);
            deep_quotient_device(ctx, mine + (u64)q * n, L, gj, z);
        }
    } catch (...) {
        ctx->pool.free(mine);
        throw;
    }
    return mine;
}

/* ------------------------------------------------------------------------------------------ */
/* FRI                                                                                         */
/* ------------------------------------------------------------------------------------------ */
void FriLayers::release(sezkp_ctx* ctx) {
    for (auto& c : commits) c.release(ctx);
    commits.clear();
    ctx->pool.free(values);
    values = nullptr;
}

// layer0 (device, N = 2^log_N values) is copied into the retained layer buffer; every further layer is folded and
// hashed by one fused kernel.  Only root 0 is needed on the host before the betas exist (v1/prover.rs:187-198); the
// other roots are collected on the device and copied back once, then absorbed in order (v1/prover.rs:219, 235).
//
// With `shard` (one process per GPU) the leaf + chunk-tree hashing of every layer of at least
// 2^SHARD_MIN_LOG values — 98 % of the compressions — is split by chunk range: rank r hashes chunks
// [n_ch*r/world, n_ch*(r+1)/world), the 32-byte chunk (subtree) roots are all-gathered, and the few levels above them
// are reduced by every rank.  Folding is replicated (it is HBM-cheap and keeps every layer's values on every GPU, so
// FRI openings need no exchange).  Two exchanges per proof: layer 0's subtree roots (its root gates the betas) and the
// subtree roots of all other sharded layers together.  Outputs are byte-identical to the unsharded path.
namespace {
constexpr int FRI_FUSE_MIN_LOG = 20, FRI_TAIL_ONE_CTA_LOG = 14, FRI_SHARD_MIN_LOG = 20;
// Large layers retain 32-leaf sub-roots instead of 1024-leaf chunk roots: a CTA still hashes 1024 leaves, but stops
// when 32 nodes are left — below that a 256-thread CTA runs 5 levels with one partly filled warp — and the top of
// all trees is reduced by upper_reduce with full CTAs.  Openings rebuild 32 leaves instead of 1024.
constexpr int FRI_BIG_CL = 5, FRI_BIG_CTA_LOG = 10;
// own chunk range of a sharded layer, in units of the CTA granularity (2^(BIG_CTA_LOG - BIG_CL) chunks)
constexpr u64 FRI_CG = 1ULL << (FRI_BIG_CTA_LOG - FRI_BIG_CL);
inline u64 fri_range_lo(u64 n_ch, int r, int world) { return ((n_ch / FRI_CG) * (u64)r / (u64)world) * FRI_CG; }
}  // namespace

int fri_range_owner(int log_N, int l, u64 row, int world) {
    const int log_len = log_N - l;
    if (world <= 1 || log_len < FRI_SHARD_MIN_LOG) return -1;
    const u64 n_ch = (1ULL << log_len) >> FRI_BIG_CL, chunk = row >> FRI_BIG_CL;
    int r = (int)(chunk * (u64)world / n_ch);  // exact for even splits; corrected below for ragged ones
    while (r > 0 && chunk < fri_range_lo(n_ch, r, world)) r--;
    while (r + 1 < world && chunk >= fri_range_lo(n_ch, r + 1, world)) r++;
    return r;
}

void fri_commit_device(sezkp_ctx* ctx, FriLayers& fl, const u64* layer0, int log_N, const u64* betas, u8* roots_host,
                       u64* final_value, HostAbsorb* absorb, const ShardInfo* shard, const u64* coset_local0) {
    REQUIRE(log_N >= 1 && log_N <= 32, "log_N %d out of range", log_N);
    const u64 N = 1ULL << log_N;
    fl.log_N = log_N;
    // the caller may have produced layer 0 in place: fl.values pre-allocated (2N elements) and layer0 == fl.values
    if (!fl.values) fl.values = (u64*)ctx->pool.alloc(2 * N * 8);
    u8* d_roots = (u8*)ctx->scratch[10].ensure((size_t)(log_N + 1) * 32 + 64);
    const bool coset = coset_local0 != nullptr;
    if (!coset && layer0 != fl.values) CUDA_CHECK(cudaMemcpyAsync(fl.values, layer0, N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    fl.commits.resize(log_N + 1);
    std::vector<u64> beta_store;
    constexpr int FUSE_MIN_LOG = FRI_FUSE_MIN_LOG, TAIL_ONE_CTA_LOG = FRI_TAIL_ONE_CTA_LOG, SHARD_MIN_LOG = FRI_SHARD_MIN_LOG;
    constexpr int BIG_CL = FRI_BIG_CL, BIG_CTA_LOG = FRI_BIG_CTA_LOG;
    const int world = shard ? shard->world : 1, rank = shard ? shard->rank : 0;
    constexpr u64 CG = FRI_CG;
    auto range_lo = [&](u64 n_ch, int r) { return fri_range_lo(n_ch, r, world); };
    // Coset-resident mode (see the kernels above): this rank holds cosets [rank*per, (rank+1)*per) of every large layer.
    const int per = coset ? 8 / world : 0;
    GroupRank* const gr = coset ? &ctx->group->ranks[ctx->group_rank] : nullptr;
    if (coset)
        REQUIRE(shard && world > 1 && world <= FRI_GATHER_MAX_WORLD && per * world == 8 && ctx->group && ctx->group->p2p &&
                    ctx->group->world == world && log_N >= SHARD_MIN_LOG && log_N - (SHARD_MIN_LOG - 1) <= FRI_GATHER_MAX_JOBS,
                "internal: coset-resident FRI needs a context group with peer access, world | 8 and a large layer 0");
    u64* coset_loc = nullptr;  // local coset arrays of layers 1 .. first small layer
    struct LocFree {
        sezkp_ctx* c;
        u64*& p;
        ~LocFree() {
            if (p) c->pool.free(p);
        }
    } loc_free{ctx, coset_loc};
    // publish `mine_base` (this rank's coset arrays), then one launch that builds the jobs' natural-order ranges
    auto gather_ranges = [&](const u64* mine_base, FriGatherJobs& jobs) {
        const void* all[64];
        group_publish_peers(gr, mine_base, ctx->stream, all);
        for (int r = 0; r < world; r++) jobs.peer[r] = (const u64*)all[r];
        jobs.per = per;
        u32 blocks = 0;
        for (int k = 0; k < jobs.n; k++) {
            REQUIRE(jobs.j[k].cnt % 256 == 0 && jobs.j[k].cnt > 0, "internal: coset range not a multiple of 256");
            jobs.j[k].blk0 = blocks;
            blocks += (u32)(jobs.j[k].cnt / 256);
        }
        fri_gather_ranges_kernel<<<blocks, 256, 0, ctx->stream>>>(jobs);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    };
    auto own_lo = [&](u64 n_ch) { return range_lo(n_ch, rank); };
    auto own_hi = [&](u64 n_ch) { return range_lo(n_ch, rank + 1); };
    auto max_own = [&](u64 n_ch) { return ((n_ch / CG + world - 1) / world + 1) * CG; };
    // all-gather the chunk roots of `cnt` sharded layers (each rank hashed its own range) and complete level 0 of
    // their `upper` arrays on every rank
    auto gather_chunk_roots = [&](Commit* const* cms, int cnt) {
        bool even = ctx->allgather_dev != nullptr;
        for (int i = 0; i < cnt; i++) even = even && (cms[i]->n_ch / CG) % (u64)world == 0;
        if (even && cnt == 1) {  // device-side: own range -> staging -> all-gather straight into level 0 of `upper` (rank-major = chunk order)
            const u64 lo = own_lo(cms[0]->n_ch), hi = own_hi(cms[0]->n_ch);
            const size_t bytes = (size_t)(hi - lo) * 32;
            u8* stage = (u8*)ctx->scratch[6].ensure(bytes);
            CUDA_CHECK(cudaMemcpyAsync(stage, cms[0]->upper + lo * 8, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            const int32_t rc = ctx->allgather_dev(ctx->allgather_dev_user, stage, bytes, cms[0]->upper, (void*)ctx->stream);
            if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "device allgather callback failed with status %d", rc);
            return;
        }
        if (even) {  // several layers: ONE collective — pack the own ranges, all-gather, scatter with one small kernel
            REQUIRE(cnt <= GatherScatter::MAX, "internal: too many sharded FRI layers");
            GatherScatter gs{};
            gs.n = cnt;
            gs.world = world;
            size_t total = 0;
            for (int i = 0; i < cnt; i++) {
                gs.upper[i] = cms[i]->upper;
                gs.own_words[i] = (u32)((own_hi(cms[i]->n_ch) - own_lo(cms[i]->n_ch)) * 8);
                gs.off_words[i] = (u32)(total / 4);
                total += (size_t)gs.own_words[i] * 4;
            }
            gs.total_words = (u32)(total / 4);
            u8* stage = (u8*)ctx->scratch[6].ensure(total * (size_t)(world + 1));
            for (int i = 0; i < cnt; i++)
                CUDA_CHECK(cudaMemcpyAsync(stage + (size_t)gs.off_words[i] * 4, cms[i]->upper + own_lo(cms[i]->n_ch) * 8, (size_t)gs.own_words[i] * 4,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
            u8* recv = stage + total;
            const int32_t rc = ctx->allgather_dev(ctx->allgather_dev_user, stage, total, recv, (void*)ctx->stream);
            if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "device allgather callback failed with status %d", rc);
            gs.recv = (const u32*)recv;
            const u64 words = (u64)gs.total_words * world;
            gather_scatter_kernel<<<blocks_for(words / 4, 256), 256, 0, ctx->stream>>>(gs);
            CUDA_CHECK(cudaGetLastError());
            ctx->launches++;
            return;
        }
        size_t per_rank = 0;
        for (int i = 0; i < cnt; i++) per_rank += (size_t)max_own(cms[i]->n_ch) * 32;
        std::vector<u8> mine(per_rank, 0), all(per_rank * (size_t)world);
        size_t off = 0;
        for (int i = 0; i < cnt; i++) {
            const u64 lo = own_lo(cms[i]->n_ch), hi = own_hi(cms[i]->n_ch);
            if (hi > lo)
                CUDA_CHECK(cudaMemcpyAsync(mine.data() + off, cms[i]->upper + lo * 8, (hi - lo) * 32, cudaMemcpyDeviceToHost, ctx->stream));
            off += (size_t)max_own(cms[i]->n_ch) * 32;
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const int32_t rc = shard->allgather(shard->user, mine.data(), per_rank, all.data());
        if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "allgather callback failed with status %d", rc);
        off = 0;
        for (int i = 0; i < cnt; i++) {
            const u64 n_ch = cms[i]->n_ch;
            for (int r = 0; r < world; r++) {
                if (r == rank) continue;
                const u64 lo = range_lo(n_ch, r), hi = range_lo(n_ch, r + 1);
                if (hi > lo)
                    CUDA_CHECK(cudaMemcpyAsync(cms[i]->upper + lo * 8, all.data() + (size_t)r * per_rank + off, (hi - lo) * 32,
                                               cudaMemcpyHostToDevice, ctx->stream));
            }
            off += (size_t)max_own(n_ch) * 32;
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // `all` is a local
    };
    // Coset-resident mode: a rank's chunk range is a complete subtree, and the openings of a large layer are served by the
    // range owner — so instead of all-gathering the 32-leaf sub-roots of every layer (32 MB per GPU for layer 0 at T = 2^22,
    // 512 MB at T = 2^26) and reducing the whole upper tree on every rank, each rank reduces its own subtree to ONE node, the
    // ranks exchange those nodes (32 bytes per layer) and everybody reduces the log2(world) levels above them.
    auto finish_ranges = [&](Commit* const* cms, int cnt, u8* const* roots_dev) {
        for (int b0 = 0; b0 < cnt; b0 += GatherScatter::MAX) {
            const int m = std::min(GatherScatter::MAX, cnt - b0);
            commit_finish_multi_range(ctx, cms + b0, m, rank, world);
            GatherScatter gs{};
            gs.n = m;
            gs.world = world;
            const size_t total = (size_t)m * 32;
            u8* stage = (u8*)ctx->scratch[6].ensure(total * (size_t)(world + 1));
            for (int i = 0; i < m; i++) {
                const Commit& c = *cms[b0 + i];
                const int lr = ilog2(c.n_ch / (u64)world);  // level of the per-rank nodes
                gs.upper[i] = c.upper + upper_off(c.n_ch, lr) * 8;
                gs.own_words[i] = 8;
                gs.off_words[i] = (u32)(8 * i);
                CUDA_CHECK(cudaMemcpyAsync(stage + 32 * (size_t)i, gs.upper[i] + (size_t)rank * 8, 32, cudaMemcpyDeviceToDevice, ctx->stream));
            }
            gs.total_words = (u32)(total / 4);
            u8* recv = stage + total;
            const int32_t rc = ctx->allgather_dev(ctx->allgather_dev_user, stage, total, recv, (void*)ctx->stream);
            if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "device allgather callback failed with status %d", rc);
            gs.recv = (const u32*)recv;
            gather_scatter_kernel<<<blocks_for((u64)gs.total_words * world / 4, 256), 256, 0, ctx->stream>>>(gs);
            CUDA_CHECK(cudaGetLastError());
            ctx->launches++;
            commit_finish_multi_top(ctx, cms + b0, m, world, roots_dev + b0);
        }
    };
    {
        CommitOpts o;
        o.roots_host = roots_host;
        const bool big0 = log_N >= FUSE_MIN_LOG;
        if (big0) o.cta_log2 = BIG_CTA_LOG;
        if (world > 1 && log_N >= SHARD_MIN_LOG) {
            Commit& c0 = fl.commits[0];
            commit_begin(ctx, c0, fl.values, N, 1, BIG_CL, nullptr, o);
            if (coset) {  // the own index range of layer 0 in natural order, pulled from the eight coset arrays
                const u64 lo = own_lo(c0.n_ch) << BIG_CL, hi = own_hi(c0.n_ch) << BIG_CL;
                FriGatherJobs jobs{};
                jobs.n = 1;
                jobs.j[0] = FriGatherJob{0, N >> 3, lo >> 3, (hi - lo) >> 3, fl.values + lo, 0};
                gather_ranges(coset_local0, jobs);
            }
            commit_chunks(ctx, c0, own_lo(c0.n_ch), own_hi(c0.n_ch), o);
            Commit* one[1] = {&c0};
            if (coset) {
                u8* r0[1] = {d_roots};
                finish_ranges(one, 1, r0);
                CUDA_CHECK(cudaMemcpyAsync(roots_host, d_roots, 32, cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            } else {
                gather_chunk_roots(one, 1);
                commit_finish(ctx, c0, o);
            }
        } else {
            commit_build(ctx, fl.commits[0], fl.values, N, 1, big0 ? BIG_CL : 10, nullptr, o);
        }
        if (absorb) {
            absorb->on_root(0, roots_host);
            if (!betas) {
                beta_store = absorb->draw_betas(log_N);
                betas = beta_store.data();
            }
        }
    }
    REQUIRE(betas != nullptr, "internal: FRI betas missing");
    // Layers of at least 2^FUSE_MIN_LOG values: one fused kernel each (fold of the previous layer + leaf hash + chunk
    // trees).  Smaller layers are latency-bound one by one (a single wave of CTAs, ~25 us each), and hashing a layer
    // does not feed the next fold: their values are folded first (one small launch per layer, then one CTA for the
    // last 2^14), and all of them are hashed side by side in ONE launch.  The levels above the chunk roots of every
    // layer are reduced together at the end.
    u64 off = N, len = N >> 1;
    int first_small = log_N + 1;
    std::vector<Commit*> sharded;
    const int coset_ls = coset ? log_N - (SHARD_MIN_LOG - 1) : 0;  // first small layer: gathered whole, replicated from there on
    if (coset) {
        // every fold down to the first small layer on the local coset arrays (no communication), then ONE launch that
        // builds this rank's range of every large layer and the whole first small layer from all ranks' arrays
        std::vector<u64> loc_off(coset_ls + 1, 0);
        u64 total = 0;
        for (int l = 1; l <= coset_ls; l++) {
            loc_off[l] = total;
            total += (N >> l) / (u64)world;
        }
        coset_loc = (u64*)ctx->pool.alloc(total * 8);
        const u64* src = coset_local0;
        u64 m = N >> 3;  // elements per coset of the layer being folded
        for (int l = 1; l <= coset_ls; l++) {
            REQUIRE(betas[l - 1] < gl::P, "beta %d is not canonical", l - 1);
            u64* dst = coset_loc + loc_off[l];
            fri_fold_cosets_kernel<<<dim3(blocks_for(m >> 1, 256), (unsigned)per), 256, 0, ctx->stream>>>(src, m, betas[l - 1], dst);
            CUDA_CHECK(cudaGetLastError());
            ctx->launches++;
            src = dst;
            m >>= 1;
        }
        FriGatherJobs jobs{};
        u64 o2 = N;
        for (int l = 1; l <= coset_ls; l++) {
            const u64 ln = N >> l, n_ch = ln >> BIG_CL;
            const u64 lo = l < coset_ls ? own_lo(n_ch) << BIG_CL : 0, hi = l < coset_ls ? own_hi(n_ch) << BIG_CL : ln;
            jobs.j[jobs.n++] = FriGatherJob{loc_off[l], ln >> 3, lo >> 3, (hi - lo) >> 3, fl.values + o2 + lo, 0};
            o2 += ln;
        }
        gather_ranges(coset_loc, jobs);
        // nobody frees or reuses its coset arrays (layer 0 included) before every peer's gather launches have finished
        group_release_peers(gr, ctx->stream);
    }
    for (int l = 1; l <= log_N; l++) {
        REQUIRE(betas[l - 1] < gl::P, "beta %d is not canonical", l - 1);
        CommitOpts o;
        const int log_len = log_N - l;
        if (log_len >= FUSE_MIN_LOG) {
            if (coset) {  // values of the own range are in place (gathered above): hash them
                Commit& c = fl.commits[l];
                o.cta_log2 = BIG_CTA_LOG;
                commit_begin(ctx, c, fl.values + off, len, 1, BIG_CL, nullptr, o);
                commit_chunks(ctx, c, own_lo(c.n_ch), own_hi(c.n_ch), o);
                sharded.push_back(&c);
            } else if (world > 1 && log_len >= SHARD_MIN_LOG) {  // fold everywhere, hash the own chunk range
                fri_fold_kernel<<<blocks_for(len, 256), 256, 0, ctx->stream>>>(fl.values + (off - 2 * len), len, betas[l - 1], fl.values + off);
                CUDA_CHECK(cudaGetLastError());
                ctx->launches++;
                Commit& c = fl.commits[l];
                o.cta_log2 = BIG_CTA_LOG;
                commit_begin(ctx, c, fl.values + off, len, 1, BIG_CL, nullptr, o);
                commit_chunks(ctx, c, own_lo(c.n_ch), own_hi(c.n_ch), o);
                sharded.push_back(&c);
            } else {
                o.fold_src = fl.values + (off - 2 * len);
                o.fold_beta = betas[l - 1];
                o.cta_log2 = BIG_CTA_LOG;
                commit_begin(ctx, fl.commits[l], fl.values + off, len, 1, BIG_CL, nullptr, o);
                commit_chunks(ctx, fl.commits[l], 0, fl.commits[l].n_ch, o);
            }
        } else {
            if (first_small > log_N) first_small = l;
            commit_begin(ctx, fl.commits[l], fl.values + off, len, 1, 10, nullptr, o);
            if (coset && l == coset_ls) {
                // gathered whole from the coset arrays: nothing to fold
            } else if (log_len >= TAIL_ONE_CTA_LOG) {
                fri_fold_kernel<<<blocks_for(len, 256), 256, 0, ctx->stream>>>(fl.values + (off - 2 * len), len, betas[l - 1], fl.values + off);
                CUDA_CHECK(cudaGetLastError());
                ctx->launches++;
            } else if (log_len == TAIL_ONE_CTA_LOG - 1 || l == 1) {  // this and all remaining layers in one CTA
                TailBetas tb{};
                const int layers = log_len + 1;
                for (int k = 0; k < layers; k++) {
                    REQUIRE(betas[l - 1 + k] < gl::P, "beta %d is not canonical", l - 1 + k);
                    tb.b[k] = betas[l - 1 + k];
                }
                fri_tail_fold_kernel<<<1, 1024, 0, ctx->stream>>>(fl.values + (off - 2 * len), len, layers, tb);
                CUDA_CHECK(cudaGetLastError());
                ctx->launches++;
            }
        }
        off += len;
        len >>= 1;
    }
    if (first_small <= log_N) commit_chunks_multi(ctx, fl.commits.data() + first_small, log_N - first_small + 1);
    if (coset) {  // large layers 1 .. first_small-1: own subtrees, one 32-byte node per layer and rank exchanged; then the small layers
        std::vector<u8*> rd(sharded.size());
        for (size_t i = 0; i < sharded.size(); i++) rd[i] = d_roots + 32 * (i + 1);
        if (!sharded.empty()) finish_ranges(sharded.data(), (int)sharded.size(), rd.data());
        commit_finish_multi(ctx, fl.commits.data() + first_small, log_N - first_small + 1, d_roots + 32 * (size_t)first_small);
    } else {
        if (!sharded.empty()) gather_chunk_roots(sharded.data(), (int)sharded.size());
        commit_finish_multi(ctx, fl.commits.data() + 1, log_N, d_roots + 32);
    }
    CUDA_CHECK(cudaMemcpyAsync(roots_host + 32, d_roots + 32, (size_t)log_N * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(final_value, fl.values + (2 * N - 2), 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (absorb)
        for (int l = 1; l <= log_N; l++) absorb->on_root(l, roots_host + 32 * l);
}

// Requests for k query indices into layer 0 (two openings per layer per query), appended to `reqs`.
// positions [k][log_N+1] is filled; request (q, l, s) gets path offset base_off + ((q*log_N + l)*2 + s)*log_N.
void fri_open_requests(const FriLayers& fl, const u64* idx0, size_t k, u64* positions, std::vector<OpenReq>& reqs, u32 base_off,
                       int keep_rank, int world, std::vector<int32_t>* req_index) {
    const int L = fl.log_N;
    const u64 N = 1ULL << L;
    if (req_index) req_index->assign(k * (size_t)L * 2, -1);
    for (size_t q = 0; q < k; q++) {
        REQUIRE(idx0[q] < N, "FRI query %zu out of range", q);
        u64 idx = idx0[q];
        for (int l = 0; l < L; l++) {
            const u64 half = (N >> l) >> 1;
            positions[q * (L + 1) + l] = idx;
            for (int s = 0; s < 2; s++) {
                const u64 row = s ? (idx ^ half) : idx;
                if (keep_rank >= 0) {  // coset-resident layers: a large layer's values exist only on the owner of the row's range
                    const int owner = fri_range_owner(L, l, row, world);
                    if ((owner < 0 ? 0 : owner) != keep_rank) continue;
                }
                if (req_index) (*req_index)[(q * L + l) * 2 + s] = (int32_t)reqs.size();
                reqs.push_back(make_open_req(fl.commits[l], 0, row, base_off + (u32)(((q * L + l) * 2 + s) * L)));
            }
            idx %= half;  // v1/prover.rs:386-388, 430-434 (half >= 1)
        }
        positions[q * (L + 1) + L] = idx;
    }
}

// k query indices into layer 0.  positions [k][log_N+1]; values [k][log_N][2]; paths [k][log_N][2][log_N][32]
void fri_open_device(sezkp_ctx* ctx, const FriLayers& fl, const u64* idx0, size_t k, u64* positions, u64* values, u8* paths) {
    const int L = fl.log_N;
    std::vector<OpenReq> reqs;
    reqs.reserve(k * L * 2);
    fri_open_requests(fl, idx0, k, positions, reqs, 0);
    std::memset(paths, 0, k * (size_t)L * 2 * L * 32);
    std::vector<u8> cr(reqs.size() * 32);
    open_batch(ctx, reqs, k * (size_t)L * 2 * L, values, cr.data(), paths);
}

/* ------------------------------------------------------------------------------------------ */
/* prove_v1                                                                                    */
/* ------------------------------------------------------------------------------------------ */
namespace {

struct Writer {  // bincode 1.3 default: fixint little-endian, u64 lengths, arrays raw
    ProofSink& b;
    explicit Writer(ProofSink& out) : b(out) { b.len = 0; }
    void u64le(u64 v) { b.put(&v, 8); }
    void raw(const void* p, size_t n) { b.put(p, n); }
    void digest_vec(const u8* p, size_t count) {
        u64le(count);
        raw(p, count * 32);
    }
};

std::vector<std::string> column_labels(u32 tau) {  // v1/openings.rs:89-116
    std::vector<std::string> out = {"input_mv", "is_first", "is_last"};
    static const char* groups[7] = {"mv_", "wflag_", "wsym_", "head_", "winlen_", "in_off_", "out_off_"};
    for (const char* g : groups)
        for (u32 r = 0; r < tau; r++) out.push_back(std::string(g) + std::to_string(r));
    return out;
}

}  // namespace

// Host-descriptor entry: the row arrays are copied in slabs of 2^20 rows on a side stream and every slab is expanded
// and committed as soon as it has landed, so all but the first slab's H2D time hides behind the column hashing.
void prove_v1_device(sezkp_ctx* ctx, const sezkp_trace_desc* desc, const u8 manifest_root[32], ProofSink& proof_out,
                     const ShardInfo* shard) {
    validate_trace(desc);
    const u64 n = desc->n_rows, nb = desc->n_blocks, tau = desc->tau;
    constexpr u64 SLAB = 1ULL << 20;
    if (shard && shard->world > 1 && ctx->allgather_dev && n % (u64)shard->world == 0 && n / (u64)shard->world >= 1024) {
        // Sharded upload: this rank copies only its slice of the row arrays over PCIe; the compact trace is then
        // all-gathered between the GPUs (NVLink) — world replicated uploads of the whole trace share the host's PCIe /
        // memory bandwidth and were the slowest phase of an 8-GPU proof.
        DeviceTraceOwner dt;
        dt.buf = ctx->scratch[2];
        ctx->scratch[2] = DevBuf();
        try {
            dt.upload_meta(ctx, desc);
        } catch (...) {
            ctx->scratch[2] = dt.buf;
            throw;
        }
        ctx->scratch[2] = dt.buf;
        const u64 world = (u64)shard->world, rank = (u64)shard->rank, per = n / world;
        const u64 r0 = rank * per, r1 = r0 + per;
        u8* base = (u8*)dt.buf.p;
        u8* stage = (u8*)ctx->scratch[6].ensure(per * tau * 2 + 256);
        struct Arr { size_t off; size_t width; const void* host; };
        const Arr arrs[4] = {{dt.o_imv, 1, desc->input_mv}, {dt.o_mv, (size_t)tau, desc->mv},
                             {dt.o_wf, (size_t)tau, desc->write_flag}, {dt.o_ws, (size_t)tau * 2, desc->write_sym}};
        for (int a = 0; a < (dt.packed ? 2 : 4); a++) {
            const size_t bytes = (size_t)per * arrs[a].width;
            CUDA_CHECK(cudaMemcpyAsync(stage, (const u8*)arrs[a].host + (size_t)r0 * arrs[a].width, bytes, cudaMemcpyHostToDevice, ctx->stream));
            const int32_t rc = ctx->allgather_dev(ctx->allgather_dev_user, stage, bytes, base + arrs[a].off, (void*)ctx->stream);
            if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "device allgather callback failed with status %d", rc);
        }
        (void)r1;
        prove_v1_resident(ctx, dt.t, manifest_root, proof_out, shard, nullptr);
        return;
    }
    SlabPlan plan;
    DeviceTraceOwner dt;
    dt.buf = ctx->scratch[2];
    ctx->scratch[2] = DevBuf();
    try {
        dt.upload_meta(ctx, desc);
    } catch (...) {
        ctx->scratch[2] = dt.buf;
        throw;
    }
    ctx->scratch[2] = dt.buf;
    if (!ctx->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    // slab s covers rows [s*SLAB, min(n,(s+1)*SLAB)); blocks are assigned to the slab in which they END
    // Slab boundaries.  Plain descriptors (1 + 4 tau bytes per row) are copy-bound — PCIe needs ~0.6 ms per 2^20 rows,
    // expand + hashing ~0.5 ms — so equal slabs keep the exposed tail (the last slab's compute) short.  Packed
    // descriptors (1 + tau bytes per row, ~0.17 ms per 2^20 rows) are compute-bound: what stays exposed is the FIRST
    // slab's copy plus a per-slab launch overhead, so the slabs grow (n/8, n/4, the rest) — each copy hides behind the
    // previous slab's compute.
    std::vector<u64> cuts;
    if ((desc->flags & SEZKP_TRACE_PACKED_OPS) && n >= 2 * SLAB) {
        cuts = {n / 8, n / 8 + n / 4, n};
    } else {
        for (u64 r = SLAB; r < n; r += SLAB) cuts.push_back(r);
        cuts.push_back(n);
    }
    u64 blk = 0, row_acc = 0, r0 = 0;
    for (const u64 r1 : cuts) {
        SlabPlan::Slab sl;
        sl.row0 = r0;
        sl.row1 = r1;
        sl.blk0 = blk;
        while (blk < nb && row_acc + desc->block_len[blk] <= r1) row_acc += desc->block_len[blk++];
        sl.blk1 = blk;
        sl.complete_rows = row_acc;  // every block that ends at or before this row is fully on the device
        if (ctx->slab_events.size() <= plan.slabs.size()) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->slab_events.push_back(e);
        }
        sl.ready = ctx->slab_events[plan.slabs.size()];
        dt.upload_rows_async(ctx->copy_stream, desc, r0, r1);
        CUDA_CHECK(cudaEventRecord(sl.ready, ctx->copy_stream));
        plan.slabs.push_back(sl);
        r0 = r1;
    }
    (void)tau;
    try {
        prove_v1_resident(ctx, dt.t, manifest_root, proof_out, shard, &plan);
    } catch (...) {
        // the slab copies read the CALLER's host arrays: never return (even with an error) while DMA may still be in flight
        cudaStreamSynchronize(ctx->copy_stream);
        throw;
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
}

// The whole prover from a device-resident compact trace.
void prove_v1_resident(sezkp_ctx* ctx, const DeviceTrace& trace, const u8 manifest_root[32], ProofSink& proof_out,
                       const ShardInfo* shard, const SlabPlan* plan) {
    const int world = shard ? shard->world : 1, rank = shard ? shard->rank : 0;
    REQUIRE(world <= 3 + 7 * (int)trace.tau, "more ranks (%d) than committed columns (%d)", world, 3 + 7 * (int)trace.tau);
    auto exchange = [&](const void* send, size_t bytes, void* recv) {  // all-gather through the host callback
        const int32_t rc = shard->allgather(shard->user, send, bytes, recv);
        if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "allgather callback failed with status %d", rc);
    };
    ctx->timings.clear();
    double t0 = now_ms();
    const double t_begin = t0;
    // Phase clock.  The device phases (expand .. fri_commit) are bracketed by CUDA events on the prover's stream and read
    // back once at the end: no synchronisation is added for the sake of timing (three of the old per-phase
    // cudaStreamSynchronize calls drained the stream where the algorithm needs no host round trip).  A phase's interval
    // contains the host work that precedes its first launch (transcript, table look-ups), so the phases add up to the
    // device timeline.  openings / serialize are host-side phases and use the host clock.
    std::vector<const char*> ev_names;
    int n_ev = 0;
    auto mark = [&](const char* name) {
        if (ctx->phase_sync) {  // option "phase_sync": the old clock — drain the stream at every phase boundary, host time
            if (!name) return;
            cudaStreamSynchronize(ctx->stream);
            const double t1 = now_ms();
            ctx->timings.push_back({name, t1 - t0});
            t0 = t1;
            return;
        }
        if ((int)ctx->phase_events.size() <= n_ev) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreate(&e));
            ctx->phase_events.push_back(e);
        }
        CUDA_CHECK(cudaEventRecord(ctx->phase_events[n_ev++], ctx->stream));
        if (name) ev_names.push_back(name);
    };
    auto flush_marks = [&]() {  // all marked work is complete (the caller has just synchronised the stream)
        if (ctx->phase_sync) return;
        for (int i = 1; i < n_ev; i++) {
            float ms = 0;
            CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->phase_events[i - 1], ctx->phase_events[i]));
            ctx->timings.push_back({ev_names[i - 1], (double)ms});
        }
        t0 = now_ms();
    };
    auto lap = [&](const char* name) {  // host-clock phase (the stream is idle when these are taken)
        double t1 = now_ms();
        ctx->timings.push_back({name, t1 - t0});
        t0 = t1;
    };
    mark(nullptr);
    const u64 n = trace.n_rows;
    const u32 tau = trace.tau;
    const int L = ilog2(n), logB = 3, log_N = L + logB;  // BLOWUP = 8 (v1/params.rs:28)
    const u64 N = 1ULL << log_N;
    const int n_cols = 3 + 7 * (int)tau;
    constexpr int NUM_QUERIES = 30, COL_CHUNK_LOG2 = 10;  // v1/params.rs:31, 37

    // A. compact trace -> committed columns
    u64* cols = (u64*)ctx->scratch[3].ensure((size_t)n_cols * n * 8);
    // One proof over several GPUs with a device-side collective: this rank expands only what it needs — its own columns
    // (c % world == rank) for every row, and all columns for the row slice [n*rank/world, n*(rank+1)/world) whose composition
    // values it computes; the slices of base_vals are all-gathered (8n bytes in total) instead of every rank composing all
    // rows from all 59 columns.
    const bool row_sliced = !plan && world > 1 && ctx->allgather_dev && n % (u64)world == 0 && n / (u64)world >= 1024;
    const u64 slice_lo = row_sliced ? n / (u64)world * (u64)rank : 0, slice_hi = row_sliced ? slice_lo + n / (u64)world : n;
    if (!plan) {
        ExpandFilter f;
        if (row_sliced) {
            f.col_mod = (u32)world;
            f.col_rem = (u32)rank;
            f.full_lo = slice_lo;
            f.full_hi = slice_hi < n ? slice_hi + 1 : n;
            f.halo = slice_hi < n ? ~0ULL : 0;  // the last slice wraps around to row 0
        }
        expand_columns_range(ctx, trace, cols, 0, n, 0, trace.n_blocks, f);
        mark("expand_columns");
    }

    // B. transcript prelude (v1/prover.rs:67-70)
    host::Transcript tr("sezkp-stark/v1");
    tr.absorb("manifest_root", manifest_root, 32);
    tr.absorb_u64("n", n);
    tr.absorb_u64("tau", tau);

    // C. column commitments (v1/prover.rs:75-81)
    std::vector<std::string> labels = column_labels(tau);
    std::vector<const char*> label_ptrs;
    for (auto& s : labels) label_ptrs.push_back(s.c_str());
    std::vector<u8> col_roots((size_t)n_cols * 32);
    Commit cm;
    FriLayers fl;
    u64* coset_local = nullptr;  // coset-resident FRI: this rank's cosets of the DEEP-LDE (layer 0), pool-owned
    try {
        // column c is committed (and later opened) by rank c % world; local column j of rank r is global r + j*world
        const int n_local = (n_cols - rank + world - 1) / world, max_local = (n_cols + world - 1) / world;
        {
            std::vector<const char*> local_labels;
            for (int j = 0; j < n_local; j++) local_labels.push_back(label_ptrs[rank + j * world]);
            std::vector<u8> local_roots((size_t)max_local * 32, 0);
            CommitOpts copt;
            copt.dedup = true;
            copt.roots_host = local_roots.data();
            copt.col_stride = (u64)world * n;
            if (!plan) {
                commit_build(ctx, cm, cols + (u64)rank * n, n, n_local, COL_CHUNK_LOG2, local_labels.data(), copt);
            } else {  // slab pipeline: wait for the slab's copy, expand it, hash the chunks whose rows are all final
                commit_begin(ctx, cm, cols + (u64)rank * n, n, n_local, COL_CHUNK_LOG2, local_labels.data(), copt);
                u64 c_done = 0;
                for (size_t si = 0; si < plan->slabs.size(); si++) {
                    const SlabPlan::Slab& sl = plan->slabs[si];
                    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, sl.ready, 0));
                    expand_columns_range(ctx, trace, cols, sl.row0, sl.row1, sl.blk0, sl.blk1);
                    const u64 c_new = (si + 1 == plan->slabs.size()) ? cm.n_ch : (sl.complete_rows >> cm.cl);
                    commit_chunks(ctx, cm, c_done, c_new, copt);
                    if (c_new > c_done) c_done = c_new;
                }
                commit_finish(ctx, cm, copt);
            }
            if (world == 1) col_roots = local_roots;
            else {  // C1 of SURVEY §2b: all-gather of 32-byte column roots
                std::vector<u8> all((size_t)world * max_local * 32);
                exchange(local_roots.data(), local_roots.size(), all.data());
                for (int c = 0; c < n_cols; c++)
                    std::memcpy(&col_roots[32 * c], &all[((size_t)(c % world) * max_local + c / world) * 32], 32);
            }
        }
        mark("column_commit");
        tr.absorb_u64("n_cols", (u64)n_cols);
        for (int c = 0; c < n_cols; c++) tr.absorb("col_root", &col_roots[32 * c], 32);

        // D. alphas, masks (v1/params.rs:76-86, v1/masking.rs:56-79)
        u64 alphas[8];
        {
            auto by = tr.challenge("alphas", 64);
            for (int i = 0; i < 8; i++) alphas[i] = le64(&by[8 * i]) % gl::P;
        }
        u64 mask[4];
        tr.absorb("masks", "masks", 5);
        tr.absorb_u64("n_masks", 1);
        tr.absorb_u64("deg", 4);
        for (int j = 0; j < 4; j++) mask[j] = le64(tr.challenge("mask_coeff", 8).data()) % gl::P;

        // E. OOD point, nudged off the coset (v1/prover.rs:118-135)
        const u64 shift = 3;
        u64 z = le64(tr.challenge("ood_point", 8).data()) % gl::P;
        while (z_on_coset(z, shift, log_N)) z = gl::add(z, 1);

        // F. composition -> iNTT -> coset LDE -> DEEP (v1/prover.rs:142-178, v1/lde.rs:42-97)
        u64* base_vals = (u64*)ctx->scratch[4].ensure(n * 8);
        compose_device(ctx, cols, n, tau, alphas, mask, 4, base_vals, slice_lo, slice_hi);
        if (row_sliced) {
            const size_t bytes = (size_t)(slice_hi - slice_lo) * 8;
            u8* stage = (u8*)ctx->scratch[6].ensure(bytes);
            CUDA_CHECK(cudaMemcpyAsync(stage, base_vals + slice_lo, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            const int32_t rc = ctx->allgather_dev(ctx->allgather_dev_user, stage, bytes, base_vals, (void*)ctx->stream);
            if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "device allgather callback failed with status %d", rc);
        }
        mark("compose");
        fl.values = (u64*)ctx->pool.alloc(2 * N * 8);  // the DEEP-LDE lands where FRI layer 0 lives: no 8N-byte copy
        u64* lde = fl.values;
        // sharded proof with a device collective: each rank evaluates 8/world cosets of the extension domain (needs world | 8)
        // Context group with peer access: the FRI layers stay coset-resident (each rank keeps its 8/world cosets, folds them
        // locally and materialises only its own hashing range in natural order) — no all-gather of the 8N-byte layer 0.
        const bool coset_fri = row_sliced && (1 << logB) % world == 0 && world <= 8 && ctx->fri_coset && ctx->group && ctx->group->p2p &&
                               ctx->group->world == world && shard->gather_root && log_N >= FRI_SHARD_MIN_LOG;
        if (coset_fri) coset_local = deep_lde_coset_local_device(ctx, base_vals, L, logB, shift, z, rank, world);
        else if (row_sliced && (1 << logB) % world == 0) deep_lde_sharded_device(ctx, base_vals, lde, L, logB, shift, z, rank, world);
        else deep_lde_device(ctx, base_vals, lde, L, logB, shift, z);
        mark("deep_lde");

        // G. FRI fold + commit; root0 is absorbed before the betas are drawn (v1/prover.rs:184-243)
        std::vector<u8> fri_roots((size_t)(log_N + 1) * 32);
        u64 final_value = 0;
        TranscriptAbsorb ab(tr);
        fri_commit_device(ctx, fl, lde, log_N, nullptr, fri_roots.data(), &final_value, &ab, shard, coset_local);
        if (coset_local) {  // fri_commit_device ended with a stream synchronisation and every peer has finished reading
            ctx->pool.free(coset_local);
            coset_local = nullptr;
        }
        mark("fri_commit");
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // fri_commit_device ended with a synchronising copy: this returns at once
        flush_marks();

        // H. AIR row queries and column openings (v1/prover.rs:248-292)
        std::vector<u64> rows(NUM_QUERIES);
        {
            auto by = tr.challenge("row_queries", 8 * NUM_QUERIES);
            for (int i = 0; i < NUM_QUERIES; i++) rows[i] = le64(&by[8 * i]) % n;
        }
        // opening order inside one RowOpenings as serialized (v1/proof.rs:48-69): per tape 9, then is_first, is_last, input_mv
        const size_t per_row = 9 * (size_t)tau + 3, k_open = per_row * NUM_QUERIES;
        std::vector<u32> o_col(k_open);
        std::vector<u64> o_row(k_open);
        for (int qi = 0; qi < NUM_QUERIES; qi++) {
            const u64 row = rows[qi], ip1 = (row + 1 < n) ? row + 1 : 0;  // next_wrap v1/prover.rs:50-58
            size_t o = qi * per_row;
            for (u32 r = 0; r < tau; r++) {
                const u32 c_mv = 3 + r, c_wf = 3 + tau + r, c_ws = 3 + 2 * tau + r, c_hd = 3 + 3 * tau + r,
                          c_wl = 3 + 4 * tau + r, c_in = 3 + 5 * tau + r, c_out = 3 + 6 * tau + r;
                const u32 cc[9] = {c_mv, c_mv, c_wf, c_ws, c_hd, c_hd, c_wl, c_in, c_out};
                const u64 rr[9] = {row, ip1, row, row, row, ip1, row, row, row};
                for (int j = 0; j < 9; j++) {
                    o_col[o] = cc[j];
                    o_row[o++] = rr[j];
                }
            }
            o_col[o] = 1; o_row[o++] = row;  // is_first
            o_col[o] = 2; o_row[o++] = row;  // is_last
            o_col[o] = 0; o_row[o++] = row;  // input_mv
        }
        // I. FRI query indices (v1/prover.rs:297; same transcript label as the AIR rows; nothing is absorbed in between,
        //    so all openings of the proof go out in one launch)
        std::vector<u64> fri_idx(NUM_QUERIES);
        {
            auto by = tr.challenge("row_queries", 8 * NUM_QUERIES);
            for (int i = 0; i < NUM_QUERIES; i++) fri_idx[i] = le64(&by[8 * i]) % N;
        }
        const int din = cm.cl, dout = ilog2(cm.n_ch), cdepth = din + dout;
        if (!ctx->open_reqs) ctx->open_reqs = new std::vector<OpenReq>();
        std::vector<OpenReq>& reqs = *ctx->open_reqs;  // reused across proofs (see host_stage)
        reqs.clear();
        reqs.reserve(k_open + (size_t)NUM_QUERIES * log_N * 2);
        // column openings: only the locally committed columns; slots of the other ranks are filled by the exchange
        std::vector<u32> req_slot;  // request index -> opening slot
        for (size_t o = 0; o < k_open; o++)
            if ((int)(o_col[o] % world) == rank) {
                reqs.push_back(make_open_req(cm, o_col[o] / world, o_row[o], (u32)(o * cdepth)));
                req_slot.push_back((u32)o);
            }
        const size_t n_col_reqs = reqs.size();
        const size_t fri_base = k_open * (size_t)cdepth, fri_digests = (size_t)NUM_QUERIES * log_N * 2 * log_N;
        std::vector<u64> f_pos((size_t)NUM_QUERIES * (log_N + 1));
        // coset-resident FRI: a large layer's values exist only on the rank that owns the row's range; rank 0 serves the
        // small (replicated) layers.  fri_req[slot] = this rank's request for FRI opening slot (q*log_N + l)*2 + s, or -1.
        std::vector<int32_t> fri_req;
        fri_open_requests(fl, fri_idx.data(), NUM_QUERIES, f_pos.data(), reqs, (u32)fri_base, coset_fri ? rank : -1, world, &fri_req);
        u64* req_val;  // results stay in the context's pinned staging buffer until the proof is serialised
        u8 *req_cr, *all_paths;
        const double t_open0 = now_ms();
        open_batch_staged(ctx, reqs, fri_base + fri_digests, &req_val, &req_cr, &all_paths);
        const double t_open1 = now_ms();
        std::vector<u64> col_val(k_open, 0);
        std::vector<u8> col_cr(k_open * 32, 0);
        for (size_t i = 0; i < n_col_reqs; i++) {
            col_val[req_slot[i]] = req_val[i];
            std::memcpy(&col_cr[(size_t)req_slot[i] * 32], &req_cr[i * 32], 32);
        }
        // per-opening record pointers (value 8 B | chunk root 32 B | path cdepth * 32 B); unsharded: assembled views below
        std::vector<const u8*> rec_ptr;
        std::vector<const u8*> fri_ptr;  // coset-resident FRI: record (value | path) of the openings served by peers, else null
        bool deliver = true;  // this rank serialises the proof
        if (world > 1) {  // gather the opening records (value, chunk root, path): every rank sends only the ones it owns
            const size_t rec = 8 + 32 + (size_t)cdepth * 32;
            std::vector<std::vector<u32>> owned(world);
            for (size_t o = 0; o < k_open; o++) owned[o_col[o] % world].push_back((u32)o);
            size_t max_own = 0;
            for (auto& v : owned) max_own = v.size() > max_own ? v.size() : max_own;
            const size_t bytes = max_own * rec;
            const size_t frec = 8 + (size_t)log_N * 32;  // FRI opening record: value | path slot
            size_t my_fri = 0;
            if (coset_fri && rank != 0)
                for (int32_t ri : fri_req) my_fri += ri >= 0;
            // grow-only staging kept in the context (a fresh 1.7 MB vector per proof and rank thread is an mmap + page faults +
            // munmap under the process-wide address-space lock)
            std::vector<u8>& mine = ctx->host_stage[0];
            if (mine.size() < bytes + my_fri * frec) mine.resize(bytes + my_fri * frec);
            if (my_fri) {  // this rank's FRI openings behind the column records, in slot order
                size_t j = 0;
                for (size_t slot = 0; slot < fri_req.size(); slot++) {
                    if (fri_req[slot] < 0) continue;
                    u8* dst = &mine[bytes + j * frec];
                    std::memcpy(dst, &req_val[fri_req[slot]], 8);
                    std::memcpy(dst + 8, &all_paths[(fri_base + slot * (size_t)log_N) * 32], (size_t)log_N * 32);
                    j++;
                }
            }
            for (size_t j = 0; j < owned[rank].size(); j++) {
                const size_t o = owned[rank][j];
                std::memcpy(&mine[j * rec], &col_val[o], 8);
                std::memcpy(&mine[j * rec + 8], &col_cr[o * 32], 32);
                std::memcpy(&mine[j * rec + 40], &all_paths[o * (size_t)cdepth * 32], (size_t)cdepth * 32);
            }
            if (shard->gather_root) {  // context group: one rendezvous, rank 0 reads the peers' records in place
                std::vector<const void*> peers(world);
                const int32_t rc = shard->gather_root(shard->user, mine.data(), peers.data());
                if (rc != 0) sezkp_fail(SEZKP_CUDA_ECOMM, "opening-record rendezvous failed with status %d", rc);
                deliver = rank == 0;
                if (deliver) {
                    rec_ptr.resize(k_open);
                    for (int r = 0; r < world; r++)
                        for (size_t j = 0; j < owned[r].size(); j++) rec_ptr[owned[r][j]] = (const u8*)peers[r] + j * rec;
                    if (coset_fri) {  // where every peer-served FRI opening sits in that peer's staging (slot order per rank)
                        fri_ptr.assign(fri_req.size(), nullptr);
                        std::vector<size_t> cnt(world, 0);
                        for (size_t q = 0; q < (size_t)NUM_QUERIES; q++) {
                            u64 idx = fri_idx[q];
                            for (int l = 0; l < log_N; l++) {
                                const u64 half = (N >> l) >> 1;
                                for (int sgn = 0; sgn < 2; sgn++) {
                                    const int owner = fri_range_owner(log_N, l, sgn ? (idx ^ half) : idx, world);
                                    if (owner > 0) fri_ptr[(q * log_N + l) * 2 + sgn] = (const u8*)peers[owner] + bytes + cnt[owner]++ * frec;
                                }
                                idx %= half;
                            }
                        }
                    }
                }
            } else {  // one process per GPU: all-gather through the host callback, every rank assembles the whole proof
                std::vector<u8>& all = ctx->host_stage[1];
                if (all.size() < (size_t)world * bytes) all.resize((size_t)world * bytes);
                exchange(mine.data(), bytes, all.data());
                rec_ptr.resize(k_open);
                for (int r = 0; r < world; r++)
                    for (size_t j = 0; j < owned[r].size(); j++) rec_ptr[owned[r][j]] = &all[(size_t)r * bytes + j * rec];
            }
        }
        const u64* o_val = col_val.data();
        const u8* o_cr = col_cr.data();
        const u8* o_paths = all_paths;
        const u8* f_paths = all_paths + fri_base * 32;
        lap("openings");
        // sub-phases of `openings` (not part of the sum): request upload + one open_kernel launch + result download / the
        // exchange of the opening records between the ranks of a sharded proof
        ctx->timings.push_back({"openings.launch_and_copy", t_open1 - t_open0});
        ctx->timings.push_back({"openings.exchange", now_ms() - t_open1});

        // J. ProofV1 in declaration order (v1/proof.rs:80-98)
        if (!deliver) {  // context group: only rank 0 writes the proof; this rank's records stay published in its staging buffer
            proof_out.len = 0;
            ctx->timings.push_back({"serialize", 0.0});
            ctx->timings.push_back({"total", now_ms() - t_begin});
            cm.release(ctx);
            fl.release(ctx);
            return;
        }
        Writer w(proof_out);  // straight into the caller's buffer
        w.u64le(N);
        w.u64le(tau);
        w.u64le((u64)n_cols);
        for (int c = 0; c < n_cols; c++) {
            w.u64le(labels[c].size());
            w.raw(labels[c].data(), labels[c].size());
            w.raw(&col_roots[32 * c], 32);
        }
        auto put_opening = [&](size_t o) {
            const u64 row = o_row[o];
            const u8 *pv, *pc, *pp;
            if (!rec_ptr.empty()) {  // sharded: the record sits in its owner's staging buffer
                pv = rec_ptr[o];
                pc = pv + 8;
                pp = pv + 40;
            } else {
                pv = (const u8*)&o_val[o];
                pc = &o_cr[o * 32];
                pp = &o_paths[o * (size_t)cdepth * 32];
            }
            w.raw(pv, 8);
            w.u64le(row);
            w.u64le(row >> din);
            w.u64le(row & ((1ULL << din) - 1));
            w.raw(pc, 32);
            w.digest_vec(pp, din);
            w.digest_vec(pp + (size_t)din * 32, dout);
        };
        w.u64le(NUM_QUERIES);
        for (int qi = 0; qi < NUM_QUERIES; qi++) {
            w.u64le(rows[qi]);
            w.u64le(tau);
            for (size_t j = 0; j < per_row; j++) put_opening(qi * per_row + j);
        }
        w.digest_vec(fri_roots.data(), log_N + 1);
        w.u64le(NUM_QUERIES);
        for (int qi = 0; qi < NUM_QUERIES; qi++) {
            w.u64le((u64)log_N + 1);
            for (int l = 0; l <= log_N; l++) w.u64le(f_pos[(size_t)qi * (log_N + 1) + l]);
            w.u64le((u64)log_N);
            for (int l = 0; l < log_N; l++) {
                const int depth = log_N - l;
                for (int s = 0; s < 2; s++) {
                    const size_t slot = ((size_t)qi * log_N + l) * 2 + s;
                    if (!fri_ptr.empty() && fri_ptr[slot]) {  // served by a peer: its record
                        w.raw(fri_ptr[slot], 8);
                        w.digest_vec(fri_ptr[slot] + 8, depth);
                    } else {
                        w.raw(&req_val[fri_req[slot]], 8);
                        w.digest_vec(&f_paths[slot * (size_t)log_N * 32], depth);
                    }
                }
            }
        }
        w.raw(&final_value, 8);
        w.raw(manifest_root, 32);
        lap("serialize");
        ctx->timings.push_back({"total", now_ms() - t_begin});
    } catch (...) {
        if (coset_local) ctx->pool.free(coset_local);
        cm.release(ctx);
        fl.release(ctx);
        throw;
    }
    cm.release(ctx);
    fl.release(ctx);
}
