// Batched Goldilocks NTT / iNTT / coset LDE for sm_100a.
//
// Replaces the reference's radix-2 in-place transforms (crates/sezkp-ffts/src/ntt.rs:79-155) and
// `evaluate_on_coset_pow2` (coset.rs:85-102) with a multi-pass decimation-in-frequency scheme:
//   N = N_1*N_2(*N_3); pass p transforms digit p of the index (most significant first) for a tile of
//   2^b rows x R columns held in shared memory, multiplies by the inter-pass twiddle w_{M_p}^{k_p*J}
//   and writes the tile back; the last pass also applies the digit-reversal so input and output are
//   both in natural order, exactly like the reference (w_N = 7^((p-1)/N)).
// A coset LDE with blow-up B is B independent size-n NTTs of the coefficients scaled by powers of
// g_j = shift*w_N^j (j < B), whose outputs interleave: out[j + B*i] = NTT_n(c_t * g_j^t)[i].
// All values in HBM are canonical residues; results are therefore bit-identical to the reference.
#include "blake3.cuh"
#include "common.cuh"
#include "gl.cuh"
#include "ntt.cuh"

namespace {

constexpr int NTT_THREADS = 256;
constexpr int MAX_PASS_BITS = 10;
constexpr int TILE_LOG_ELEMS = 13;  // 8192 elements = 64 KiB per tile
constexpr int MAX_LOG_R = 6;

// Bounds-checked build (`make dbg` -> libsezkp_cuda_dbg.so, run by tests/test_gpu_named_shapes.py::test_bounds_checked_build):
// every global address a pass kernel forms is checked against the extent of the buffer it belongs to.  compute-sanitizer is
// closed on this pool, so this is the memory-safety evidence for the addressing of the pass descriptors.
#ifdef SEZKP_BOUNDS_CHECK
#define NTT_CHECK(off, ext, what)                                                                                          \
    do {                                                                                                                   \
        if ((u64)(off) >= (u64)(ext)) {                                                                                    \
            printf("NTT bounds violation: %s offset %llu >= extent %llu (block %u, thread %u, b=%d logR=%d)\n", what,      \
                   (unsigned long long)(off), (unsigned long long)(ext), blockIdx.x, threadIdx.x, d.b, d.logR);            \
            __trap();                                                                                                      \
        }                                                                                                                  \
    } while (0)
#else
#define NTT_CHECK(off, ext, what) ((void)0)
#endif

struct PassDesc {
    const u64* in;
    u64* out;
    int b, logR, pitch, canon_in, inv;
    u32 col_tiles, U;
    u64 total_cols;
    // input addressing
    u64 in_row_stride;
    int in_clog;
    u64 in_cs_lo, in_cs_hi, in_u_stride, in_v_stride;
    int in_v_shift;
    // output addressing
    u64 out_row_stride;
    int out_clog;
    u64 out_cs_lo, out_cs_hi, out_u_stride, out_v_stride;
    int load_rows_fast, store_rows_fast;
    // tables
    const u64* W;  // w_{2^b}^e, e < 2^b (inter-step twiddles of a two-step pass)
    const u64* tw_lo;
    const u64* tw_hi;
    int tw_lb, use_tw;
    u64 tw_stride;
    const u64* tw_full;  // optional full twiddle matrix of this pass (null: two-level tables)
    u64 tw_pitch;
    u64 scale;
    int use_scale;
    // coset pre-scale: x *= GA[j][row] * GB[j][C]
    const u64* GA;
    const u64* GB;
    int use_pre, use_gb, coset_from_col, coset_log;
    u32 ga_pitch, gb_pitch;
    // buffer extents in elements (bounds-checked build only: make dbg, -DSEZKP_BOUNDS_CHECK)
    u64 in_extent, out_extent, tw_extent, fuse_extent;
    // fused leaf hashing (K7, last LDE pass only): the pass writes 32-leaf labeled sub-roots instead of the values
    const b3::LabelTemplate* fuse_tpl;  // [batch columns]
    u32* fuse_upper;                    // level 0 of column v's retained tree at fuse_upper + v * fuse_col_words
    u64 fuse_col_words;
};

// Tile layout: element (column c, row r) lives at c*pitch + pos(r), pos(r) = r + (r >> K2) — one word of skew per
// group of B = 2^K2 rows, so that both register steps (stride B+1 in the first, contiguous runs in the second) and the
// digit-swapped read of the store phase touch all sixteen 8-byte banks evenly; the host picks pitch = 16/R (mod 16) so
// that the R columns and 16/R consecutive rows a half-warp covers fall into sixteen different banks (tile_pitch()).
//
// The 2^b-point transform of a column is the Cooley-Tukey split 2^b = A*B (A = 2^K1, B = 2^K2), rows r = B*r1 + r2,
// outputs k = k1 + A*k2:   X[k1 + A*k2] = sum_{r2} w_B^{r2*k2} * ( w_{2^b}^{r2*k1} * sum_{r1} x[B*r1 + r2] * w_A^{r1*k1} )
//   step 1: one thread = one (column, r2): A-point DFT over r1 in registers (power-of-two twiddles: shifts), then the
//           inter-step twiddle Ws[r2*k1] — the only general multiplication of the step — back to position (k1, r2);
//   step 2: one thread = one (column, k1): B-point DFT over r2 in registers, result at position (k1, k2).
// Values in the tile are canonical (<= p) when a step starts and arbitrary 64-bit representatives after the last one;
// the store phase canonicalises.
template <int K1, int K2, bool INV>
__device__ __forceinline__ void dft_step1(u64* tile, const u64* Ws, int logR, int pitch, u32 eps) {
    constexpr int A = 1 << K1, B = 1 << K2;
    constexpr int stride = K2 ? B + 1 : 1;
    const int tasks = B << logR;
    const int R = 1 << logR;
    for (int g = threadIdx.x; g < tasks; g += NTT_THREADS) {
        const int c = g & (R - 1);
        const int r2 = g >> logR;
        u64* col = tile + c * pitch + r2;
        u64 x[A];
#pragma unroll
        for (int t = 0; t < A; t++) x[t] = col[t * stride];
        gl::lazy::dft_pow2<K1, INV>(x, eps);
#pragma unroll
        for (int k = 0; k < A; k++) {
            const u64 y = x[gl::lazy::brev_bits(k, K1)];
            if (K2 == 0) col[k * stride] = y;                                  // single-step pass: stays lazy
            else if (k == 0) col[0] = gl::lazy::canon2(y);
            else col[k * stride] = gl::lazy::mulc(y, Ws[r2 * k], eps);
        }
    }
}
template <int K1, int K2, bool INV>
__device__ __forceinline__ void dft_step2(u64* tile, int logR, int pitch, u32 eps) {
    constexpr int A = 1 << K1, B = 1 << K2;
    const int tasks = A << logR;
    const int R = 1 << logR;
    for (int g = threadIdx.x; g < tasks; g += NTT_THREADS) {
        const int c = g & (R - 1);
        const int k1 = g >> logR;
        u64* col = tile + c * pitch + k1 * (B + 1);
        u64 x[B];
#pragma unroll
        for (int t = 0; t < B; t++) x[t] = col[t];
        gl::lazy::dft_pow2<K2, INV>(x, eps);
#pragma unroll
        for (int k = 0; k < B; k++) col[k] = x[gl::lazy::brev_bits(k, K2)];
    }
}

template <int K1, int K2, bool INV>
__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_pass_kernel(const PassDesc d) {
    extern __shared__ u64 smem[];
    constexpr int b = K1 + K2;
    constexpr int rows = 1 << b;
    const int pitch = d.pitch;
    const int logR = d.logR, R = 1 << logR;
    u64* tile = smem;
    u64* Ws = smem + (size_t)R * pitch;
    const u32 eps = gl::lazy::k_eps32;
    auto pos = [](int r) { return K2 ? r + (r >> K2) : r; };

    const u32 tile_id = blockIdx.x;
    const u32 ct = tile_id % d.col_tiles;
    const u32 rest = tile_id / d.col_tiles;
    const u32 u = rest % d.U;
    const u64 v = rest / d.U;
    const u64 C0 = (u64)ct << logR;
    const u64* in_base = d.in + u * d.in_u_stride + (v >> d.in_v_shift) * d.in_v_stride;
    u64* out_base = d.out + u * d.out_u_stride + v * d.out_v_stride;

    if (K2 > 0)
        for (int i = threadIdx.x; i < rows; i += NTT_THREADS) Ws[i] = d.W[i];

    // ---- load (optionally pre-scaled by the coset powers); per-thread invariants hoisted out of the loops ----
    const int tid = threadIdx.x;
    auto load_column = [&](int c, int row0, int rstep) {
        const u64 C = C0 + c;
        u64* dst = tile + c * pitch;
        if (C >= d.total_cols) {
            for (int row = row0; row < rows; row += rstep) dst[pos(row)] = 0;
            return;
        }
        const u64* src = in_base + (C & ((1ULL << d.in_clog) - 1)) * d.in_cs_lo + (C >> d.in_clog) * d.in_cs_hi;
        const u64 rs = d.in_row_stride;
        if (d.use_pre) {
            const u32 j = (u32)((d.coset_from_col ? C : v) & ((1u << d.coset_log) - 1));
            const u64* ga = d.GA + (size_t)j * d.ga_pitch;
            const u64 gbc = d.use_gb ? d.GB[(size_t)j * d.gb_pitch + C] : 1;
#pragma unroll 4
            for (int row = row0; row < rows; row += rstep) {
                u64 g = ga[row];
                if (d.use_gb) g = gl::lazy::mulc(g, gbc, eps);
                dst[pos(row)] = gl::lazy::mulc(src[(u64)row * rs], g, eps);
            }
        } else if (d.canon_in) {  // caller-supplied device data: accept any 64-bit representative
#pragma unroll 8
            for (int row = row0; row < rows; row += rstep) dst[pos(row)] = gl::lazy::canon2(src[(u64)row * rs]);
        } else {
#pragma unroll 8
            for (int row = row0; row < rows; row += rstep) dst[pos(row)] = src[(u64)row * rs];
        }
    };
    if (d.load_rows_fast && !d.use_pre) {
        // rows are contiguous in memory (last pass): flat index over (column, row), row fastest, eight independent
        // loads in flight per thread — with one load per loop trip this pass was bound by memory latency
        const int total = R << b;
        for (int i0 = tid; i0 < total; i0 += 8 * NTT_THREADS) {
            u64 w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = i0 + j * NTT_THREADS;
                const int c = i >> b, row = i & (rows - 1);
                const u64 C = C0 + c;
                w[j] = 0;
                if (i < total && C < d.total_cols)
                    w[j] = in_base[(C & ((1ULL << d.in_clog) - 1)) * d.in_cs_lo + (C >> d.in_clog) * d.in_cs_hi + (u64)row * d.in_row_stride];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = i0 + j * NTT_THREADS;
                if (i < total) tile[(i >> b) * pitch + pos(i & (rows - 1))] = d.canon_in ? gl::lazy::canon2(w[j]) : w[j];
            }
        }
    } else if (d.load_rows_fast) {
        for (int c = 0; c < R; c++) load_column(c, tid, NTT_THREADS);
    } else {
        load_column(tid & (R - 1), tid >> logR, NTT_THREADS >> logR);
    }
    __syncthreads();
    dft_step1<K1, K2, INV>(tile, Ws, logR, pitch, eps);
    __syncthreads();
    if (K2 > 0) {
        dft_step2<K1, (K2 > 0 ? K2 : 1), INV>(tile, logR, pitch, eps);
        __syncthreads();
    }
    // ---- store: output index k = k1 + A*k2 sits at position (k1, k2); inter-pass twiddle, optional scale; everything
    //      written to HBM is the canonical residue ----
    auto store_column = [&](int c, int k0, int kstep) {
        const u64 C = C0 + c;
        if (C >= d.total_cols) return;
        const u64* src = tile + c * pitch;
        u64* dst = out_base + (C & ((1ULL << d.out_clog) - 1)) * d.out_cs_lo + (C >> d.out_clog) * d.out_cs_hi;
        const u64 rs = d.out_row_stride;
        const u64* twf = d.tw_full ? d.tw_full + C : nullptr;
        const u64 Cs = C * d.tw_stride;
#pragma unroll 4
        for (int k = k0; k < rows; k += kstep) {
            u64 x = src[K2 ? (k & ((1 << K1) - 1)) * ((1 << K2) + 1) + (k >> K1) : k];
            if (d.use_tw) {
                u64 w;
                if (twf) w = twf[(u64)k * d.tw_pitch];
                else {
                    const u64 E = (u64)k * Cs;
                    w = gl::lazy::mulc(d.tw_lo[E & ((1ULL << d.tw_lb) - 1)], d.tw_hi[E >> d.tw_lb], eps);
                }
                x = gl::lazy::mulc(x, w, eps);
                if (d.use_scale) x = gl::lazy::mulc(x, d.scale, eps);
            } else if (d.use_scale) {
                x = gl::lazy::mulc(x, d.scale, eps);
            } else {
                x = gl::lazy::canon2(x);
            }
            dst[(u64)k * rs] = x;
        }
    };
    if (d.store_rows_fast) {
        for (int c = 0; c < R; c++) store_column(c, tid, NTT_THREADS);
    } else {
        store_column(tid & (R - 1), tid >> logR, NTT_THREADS >> logR);
    }
}

// Second-generation pass kernel (two-step passes, b >= 6): global loads go straight into the registers of the thread
// that runs the first register DFT on them, and the second register DFT's results go straight from registers to HBM.
// The shared-memory tile is touched once (written after step 1, read by step 2) instead of three times, and two of the
// three block barriers disappear.  Why this works without any shuffle: in the first-generation kernel the thread that
// copied rows {r2 + B*t} of column c into the tile was already the thread that ran step 1 on exactly those rows, and the
// thread that ran step 2 for (c, k1) was the one that stored rows {k1 + A*k2} — both round trips were thread-private.
//   columns contiguous in HBM (strided passes):  task = (c fastest, r2 / k1 slowest)  -> a warp touches 32/R rows x R*8 B
//   rows contiguous in HBM (last pass / single pass): task = (r2 or k1 fastest)      -> a warp touches B*8 (A*8) contiguous B
// The inter-step twiddles w_{2^b}^(r2*k1) sit in shared memory as a [A][B] matrix so that both task mappings read them
// without bank conflicts (a flat power table indexed r2*k1 is a stride-k1 access when r2 runs across the lanes).
template <int K1, int K2, bool INV, int MINB, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB) ntt_pass2_kernel(const PassDesc d) {
    extern __shared__ u64 smem[];
    constexpr int A = 1 << K1, B = 1 << K2, b = K1 + K2, rows = 1 << b;
    static_assert(K2 > 0, "two-step passes only");
    const int pitch = d.pitch, logR = d.logR, R = 1 << logR;
    u64* tile = smem;
    u64* Ws = smem + (size_t)R * pitch;  // [A][B]: Ws[k1*B + r2] = w^(r2*k1)
    const u32 eps = gl::lazy::k_eps32;
    const int tid = threadIdx.x;

    const u32 tile_id = blockIdx.x;
    const u32 ct = tile_id % d.col_tiles;
    const u32 rest = tile_id / d.col_tiles;
    const u32 u = rest % d.U;
    const u64 v = rest / d.U;
    const u64 C0 = (u64)ct << logR;
    const u64* in_base = d.in + u * d.in_u_stride + (v >> d.in_v_shift) * d.in_v_stride;
    u64* out_base = d.out + u * d.out_u_stride + v * d.out_v_stride;

    for (int i = tid; i < rows; i += THREADS) Ws[i] = d.W[(i >> K2) * (i & (B - 1))];

    // ---- step 1: A-point DFTs over r1 of rows B*r1 + r2, operands straight from HBM ----
    const int tasks1 = B << logR;
    for (int g0 = 0; g0 < tasks1; g0 += THREADS) {
        const int g = g0 + tid;
        const bool live = g < tasks1;
        int c, r2;
        if (d.load_rows_fast) {
            r2 = g & (B - 1);
            c = g >> K2;
        } else {
            c = g & (R - 1);
            r2 = g >> logR;
        }
        const u64 C = C0 + c;
        u64 x[A];
        if (live && C < d.total_cols) {
            const u64* src = in_base + (C & ((1ULL << d.in_clog) - 1)) * d.in_cs_lo + (C >> d.in_clog) * d.in_cs_hi + (u64)r2 * d.in_row_stride;
            const u64 rs = d.in_row_stride << K2;  // rows r2 + B*t
            NTT_CHECK((u64)(src - d.in) + (u64)(A - 1) * rs, d.in_extent, "pass load");
#pragma unroll
            for (int t = 0; t < A; t++) x[t] = src[(u64)t * rs];
            if (d.use_pre) {
                const u32 j = (u32)((d.coset_from_col ? C : v) & ((1u << d.coset_log) - 1));
                const u64* ga = d.GA + (size_t)j * d.ga_pitch + r2;
                if (d.use_gb) {
                    const u64 gbc = d.GB[(size_t)j * d.gb_pitch + C];
#pragma unroll
                    for (int t = 0; t < A; t++) x[t] = gl::lazy::mulc(x[t], gl::lazy::mulc(ga[t << K2], gbc, eps), eps);
                } else {
#pragma unroll
                    for (int t = 0; t < A; t++) x[t] = gl::lazy::mulc(x[t], ga[t << K2], eps);
                }
            } else if (d.canon_in) {
#pragma unroll
                for (int t = 0; t < A; t++) x[t] = gl::lazy::canon2(x[t]);
            }
        } else {
#pragma unroll
            for (int t = 0; t < A; t++) x[t] = 0;
        }
        if (g0 == 0) __syncthreads();  // Ws complete (the global loads above are already in flight)
        if (live) {
            gl::lazy::dft_pow2<K1, INV>(x, eps);
            u64* col = tile + c * pitch + r2;
            const u64* wrow = Ws + r2;
#pragma unroll
            for (int k = 0; k < A; k++) {
                const u64 y = x[gl::lazy::brev_bits(k, K1)];
                col[k * (B + 1)] = k == 0 ? gl::lazy::canon2(y) : gl::lazy::mulc(y, wrow[k << K2], eps);
            }
        }
    }
    __syncthreads();
    // ---- step 2: B-point DFTs over r2, results straight to HBM (inter-pass twiddle, scale, canonical) ----
    const int tasks2 = A << logR;
    for (int g = tid; g < tasks2; g += THREADS) {
        int c, k1;
        if (d.store_rows_fast) {
            k1 = g & (A - 1);
            c = g >> K1;
        } else {
            c = g & (R - 1);
            k1 = g >> logR;
        }
        const u64 C = C0 + c;
        if (C >= d.total_cols) continue;
        const u64* col = tile + c * pitch + k1 * (B + 1);
        u64 x[B];
#pragma unroll
        for (int t = 0; t < B; t++) x[t] = col[t];
        gl::lazy::dft_pow2<K2, INV>(x, eps);
        u64* dst = out_base + (C & ((1ULL << d.out_clog) - 1)) * d.out_cs_lo + (C >> d.out_clog) * d.out_cs_hi + (u64)k1 * d.out_row_stride;
        const u64 rs = d.out_row_stride << K1;  // rows k1 + A*k2
        NTT_CHECK((u64)(dst - d.out) + (u64)(B - 1) * rs, d.out_extent, "pass store");
        if (d.use_tw) {
            if (d.tw_full) {
                const u64* twf = d.tw_full + C + (u64)k1 * d.tw_pitch;
                const u64 tp = d.tw_pitch << K1;
                NTT_CHECK((u64)(twf - d.tw_full) + (u64)(B - 1) * tp, d.tw_extent, "twiddle matrix");
                constexpr int G = MINB >= 3 ? 4 : 8;  // twiddle loads in flight per thread (registers: x[] already holds 2*B)
#pragma unroll
                for (int k0 = 0; k0 < B; k0 += G) {
                    u64 w[G];
#pragma unroll
                    for (int q = 0; q < G; q++) w[q] = twf[(u64)(k0 + q) * tp];
#pragma unroll
                    for (int q = 0; q < G; q++) {
                        u64 y = gl::lazy::mulc(x[gl::lazy::brev_bits(k0 + q, K2)], w[q], eps);
                        if (d.use_scale) y = gl::lazy::mulc(y, d.scale, eps);
                        dst[(u64)(k0 + q) * rs] = y;
                    }
                }
            } else {
                const u64 Cs = C * d.tw_stride;
#pragma unroll  // full unroll: a run-time index into x[] would move the whole array to local memory
                for (int k2 = 0; k2 < B; k2++) {
                    const u64 E = (u64)(k1 + (k2 << K1)) * Cs;
                    const u64 w = gl::lazy::mulc(d.tw_lo[E & ((1ULL << d.tw_lb) - 1)], d.tw_hi[E >> d.tw_lb], eps);
                    u64 y = gl::lazy::mulc(x[gl::lazy::brev_bits(k2, K2)], w, eps);
                    if (d.use_scale) y = gl::lazy::mulc(y, d.scale, eps);
                    dst[(u64)k2 * rs] = y;
                }
            }
        } else if (d.use_scale) {
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) dst[(u64)k2 * rs] = gl::lazy::mulc(x[gl::lazy::brev_bits(k2, K2)], d.scale, eps);
        } else {
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) dst[(u64)k2 * rs] = gl::lazy::canon2(x[gl::lazy::brev_bits(k2, K2)]);
        }
    }
}

// K7 (north_star item 3): the last pass of a coset LDE fused with the labeled leaf hash and the first five tree levels, so
// the extended column never travels to HBM: steps 1 and 2 are those of ntt_pass2_kernel, but step 2 puts its canonical
// outputs back into the shared-memory tile in OUTPUT order — a tile row is R consecutive leaves of the extended column —
// and every thread then owns one aligned group of 32 leaves: 32 leaf compressions + 31 parent compressions as a
// streaming reduction with a five-entry stack, no communication, one 32-byte sub-root stored per thread.  The tile always
// holds 8192 leaves = 256 groups = one per thread.  Requires R >= 32 (pass width <= 8 bits), i.e. three-pass plans.
template <int K1, int K2>
__global__ void __launch_bounds__(NTT_THREADS, 2) lde_hash_pass_kernel(const PassDesc d) {
    extern __shared__ u64 smem[];
    constexpr int A = 1 << K1, B = 1 << K2, b = K1 + K2, rows = 1 << b, THREADS = NTT_THREADS;
    constexpr int PER = (1 << TILE_LOG_ELEMS) / THREADS;  // 32 values per thread in every phase
    constexpr int NIT1 = PER / A, NIT2 = PER / B;
    const int pitch = d.pitch, logR = d.logR, R = 1 << logR;
    u64* tile = smem;
    u64* Ws = smem + (size_t)R * pitch;
    const u32 eps = gl::lazy::k_eps32;
    const int tid = threadIdx.x;
    const u32 tile_id = blockIdx.x;
    const u32 ct = tile_id % d.col_tiles;
    const u32 rest = tile_id / d.col_tiles;
    const u32 u = rest % d.U;
    const u64 v = rest / d.U;
    const u64 C0 = (u64)ct << logR;
    const u64* in_base = d.in + u * d.in_u_stride + (v >> d.in_v_shift) * d.in_v_stride;
    for (int i = tid; i < rows; i += THREADS) Ws[i] = d.W[(i >> K2) * (i & (B - 1))];
    // ---- step 1 (rows are contiguous in HBM: r2 runs across the lanes) ----
#pragma unroll
    for (int it = 0; it < NIT1; it++) {
        const int g = tid + it * THREADS;
        const int r2 = g & (B - 1), c = g >> K2;
        const u64 C = C0 + c;
        const u64* src = in_base + (C & ((1ULL << d.in_clog) - 1)) * d.in_cs_lo + (C >> d.in_clog) * d.in_cs_hi + (u64)r2 * d.in_row_stride;
        const u64 rs = d.in_row_stride << K2;
        u64 x[A];
        NTT_CHECK((u64)(src - d.in) + (u64)(A - 1) * rs, d.in_extent, "fused pass load");
#pragma unroll
        for (int t = 0; t < A; t++) x[t] = src[(u64)t * rs];
        if (it == 0) __syncthreads();  // Ws complete
        gl::lazy::dft_pow2<K1, false>(x, eps);
        u64* col = tile + c * pitch + r2;
        const u64* wrow = Ws + r2;
#pragma unroll
        for (int k = 0; k < A; k++) {
            const u64 y = x[gl::lazy::brev_bits(k, K1)];
            col[k * (B + 1)] = k == 0 ? gl::lazy::canon2(y) : gl::lazy::mulc(y, wrow[k << K2], eps);
        }
    }
    __syncthreads();
    // ---- step 2: all inputs into registers first (the tile is about to be overwritten in output order) ----
    u64 xs[PER];
#pragma unroll
    for (int it = 0; it < NIT2; it++) {
        const int g = tid + it * THREADS;
        const int c = g & (R - 1), k1 = g >> logR;
        const u64* col = tile + c * pitch + k1 * (B + 1);
#pragma unroll
        for (int t = 0; t < B; t++) xs[it * B + t] = col[t];
    }
    __syncthreads();
    const int opitch = R + 1;  // output-order tile: leaf (row k, column c) at k * (R + 1) + c
#pragma unroll
    for (int it = 0; it < NIT2; it++) {
        const int g = tid + it * THREADS;
        const int c = g & (R - 1), k1 = g >> logR;
        u64 x[B];
#pragma unroll
        for (int t = 0; t < B; t++) x[t] = xs[it * B + t];
        gl::lazy::dft_pow2<K2, false>(x, eps);
#pragma unroll
        for (int k2 = 0; k2 < B; k2++) tile[(k1 + (k2 << K1)) * opitch + c] = gl::lazy::canon2(x[gl::lazy::brev_bits(k2, K2)]);
    }
    __syncthreads();
    // ---- step 3: one aligned group of 32 leaves per thread -> labeled leaves -> five tree levels -> sub-root ----
    const int parts = R >> 5;  // groups per tile row
    const int row = tid / parts, part = tid - row * parts;
    const u64* leaves = tile + row * opitch + part * 32;
    const b3::LabelTemplate tpl = d.fuse_tpl[v];
    u32 stack[5][8];
    u32 cur[8];
    B3_DISPATCH_LABELED(tpl, {
        for (int i = 0; i < 32; i++) {
            b3::leaf_labeled_w<B3W>(tpl, leaves[i], cur);
            int lvl = 0;
            for (int idx = i; idx & 1; idx >>= 1, lvl++) {
                u32 o[8];
                b3::parent(stack[lvl], cur, o);
#pragma unroll
                for (int w = 0; w < 8; w++) cur[w] = o[w];
            }
            if (lvl < 5) {
#pragma unroll
                for (int w = 0; w < 8; w++) stack[lvl][w] = cur[w];
            }
        }
    })
    const u64 leaf0 = u * d.out_u_stride + (C0 + (u64)part * 32) + (u64)row * d.out_row_stride;  // first leaf of the group in column v
    u32* dst = d.fuse_upper + v * d.fuse_col_words + (leaf0 >> 5) * 8;
    NTT_CHECK((u64)(dst - d.fuse_upper) + 7, d.fuse_extent, "sub-root store");
    NTT_CHECK((leaf0 >> 5), (d.fuse_col_words + 8) / 16, "sub-root index within the column's level 0");
    *(uint4*)dst = make_uint4(cur[0], cur[1], cur[2], cur[3]);
    *(uint4*)(dst + 4) = make_uint4(cur[4], cur[5], cur[6], cur[7]);
}

// Full inter-pass twiddle matrix T[k][J] = w^(k*J*stride) (k < rows, J < pitch): same shape as the data block, so the
// store loop reads it with the data's own coalescing; it stays in L2 across the columns of a batch.
// `scale` (N^-1 of an inverse transform, else 1) is folded into the first pass's matrix, so the last pass of an inverse
// NTT canonicalises instead of multiplying every element once more.
__global__ void tw_fill_kernel(u64* __restrict__ out, u64 count, u64 pitch, u64 stride, const u64* __restrict__ lo,
                               const u64* __restrict__ hi, int lb, u64 scale) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const u64 k = i / pitch, J = i - k * pitch;
    const u64 E = k * J * stride;
    out[i] = gl::mul(gl::mul(lo[E & ((1ULL << lb) - 1)], hi[E >> lb]), scale);
}

typedef void (*pass_fn)(const PassDesc);
template <int K1, int K2>
pass_fn get_fn(bool inv) { return inv ? ntt_pass_kernel<K1, K2, true> : ntt_pass_kernel<K1, K2, false>; }

// register-step split of a pass of b bits
void split_bits(int b, int& K1, int& K2) {
    static const int k1[11] = {0, 1, 2, 3, 4, 5, 3, 4, 4, 5, 5};
    K1 = k1[b];
    K2 = b - K1;
}
template <int K1, int K2, int MINB, int THREADS = NTT_THREADS>
pass_fn get_fn2(bool inv) { return inv ? ntt_pass2_kernel<K1, K2, true, MINB, THREADS> : ntt_pass2_kernel<K1, K2, false, MINB, THREADS>; }
// second-generation kernel (fused global I/O) for the two-step widths; `gen` 1 selects the first-generation kernel (A/B runs)
pass_fn kernel_for_bits(int b, bool inv, int gen) {
    if (gen >= 2) switch (b) {
        case 6: return get_fn2<3, 3, 3>(inv);
        case 7: return get_fn2<4, 3, 3>(inv);
        case 8: return get_fn2<4, 4, 3>(inv);
        case 9: return gen == 4 ? get_fn2<5, 4, 4, 128>(inv) : gen == 3 ? get_fn2<5, 4, 2>(inv) : get_fn2<5, 4, 3>(inv);
        case 10: return gen == 4 ? get_fn2<5, 5, 4, 128>(inv) : gen == 3 ? get_fn2<5, 5, 2>(inv) : get_fn2<5, 5, 3>(inv);
        default: break;
    }
    switch (b) {
        case 1: return get_fn<1, 0>(inv);
        case 2: return get_fn<2, 0>(inv);
        case 3: return get_fn<3, 0>(inv);
        case 4: return get_fn<4, 0>(inv);
        case 5: return get_fn<5, 0>(inv);
        case 6: return get_fn<3, 3>(inv);
        case 7: return get_fn<4, 3>(inv);
        case 8: return get_fn<4, 4>(inv);
        case 9: return get_fn<5, 4>(inv);
        case 10: return get_fn<5, 5>(inv);
        default: sezkp_fail(SEZKP_CUDA_EINVAL, "unsupported pass width %d", b);
    }
}
// shared-memory pitch of a tile column (see the layout note above the kernel)
int tile_pitch(int b, int logR) {
    int K1, K2;
    split_bits(b, K1, K2);
    const int base = (1 << b) + (K2 ? (1 << K1) : 0);
    const int want = logR >= 4 ? 1 : ((16 >> logR) & 15);  // a 64-bit access is served per half-warp: 16 lanes x 8 B
    return base + ((want - base) & 15);
}

}  // namespace

/* ------------------------------------------------------------------------------------------ */
/* tables                                                                                     */
/* ------------------------------------------------------------------------------------------ */
struct NttTables {
    int L = 0;
    bool inverse = false;
    std::vector<int> plan;    // pass widths, most significant digit first
    u64* W[MAX_PASS_BITS + 1] = {};  // device, per pass width
    u64* tw_lo = nullptr;
    u64* tw_hi = nullptr;
    u64* tw_full[3] = {nullptr, nullptr, nullptr};  // per non-last pass, built on first use when the transform is small enough
    int tw_lb = 0;
    u64 scale = 1;  // N^-1 for inverse
    bool tw0_scaled = false;  // tw_full[0] already carries `scale`
    // coset tables, keyed by (logB, shift)
    struct Coset {
        u64* GA = nullptr;
        u64* GB = nullptr;
        u32 ga_pitch = 0, gb_pitch = 0;
    };
    std::map<std::pair<int, u64>, Coset> cosets;
};

static std::vector<int> make_plan(int L) {
    std::vector<int> p;
    if (L <= MAX_PASS_BITS) p = {L};
    else if (L <= 2 * MAX_PASS_BITS) p = {(L + 1) / 2, L / 2};
    else {
        int a = (L + 2) / 3, b = (L - a + 1) / 2, c = L - a - b;
        p = {a, b, c};
    }
    for (int x : p)
        if (x < 1 || x > MAX_PASS_BITS) sezkp_fail(SEZKP_CUDA_EINVAL, "log size %d out of range", L);
    return p;
}

static u64* upload(sezkp_ctx* ctx, const std::vector<u64>& h) {
    u64* d = nullptr;
    cudaError_t e = cudaMalloc(&d, h.size() * 8);
    if (e != cudaSuccess) sezkp_fail(SEZKP_CUDA_ENOMEM, "cudaMalloc(table %zu B): %s", h.size() * 8, cudaGetErrorString(e));
    upload_table(ctx, d, h.data(), h.size() * 8);
    return d;
}

static NttTables* get_tables(sezkp_ctx* ctx, int L, bool inverse) {
    u64 key = ((u64)L << 1) | (inverse ? 1 : 0);
    auto it = ctx->ntt_tables.find(key);
    if (it != ctx->ntt_tables.end()) return it->second;
    NttTables* t = new NttTables();
    t->L = L;
    t->inverse = inverse;
    t->plan = make_plan(L);
    for (int b : t->plan) {
        if (t->W[b]) continue;
        u64 w = gl::root_2exp((unsigned)b);
        if (inverse) w = gl::inv(w);
        std::vector<u64> h((size_t)1 << b);
        u64 x = 1;
        for (auto& e : h) {
            e = x;
            x = gl::mul(x, w);
        }
        t->W[b] = upload(ctx, h);
    }
    u64 wN = gl::root_2exp((unsigned)L);
    if (inverse) wN = gl::inv(wN);
    t->tw_lb = (L + 1) / 2;
    {
        std::vector<u64> lo((size_t)1 << t->tw_lb), hi((size_t)1 << (L - t->tw_lb));
        u64 x = 1;
        for (auto& e : lo) {
            e = x;
            x = gl::mul(x, wN);
        }
        u64 step = x;  // wN^(2^lb)
        x = 1;
        for (auto& e : hi) {
            e = x;
            x = gl::mul(x, step);
        }
        t->tw_lo = upload(ctx, lo);
        t->tw_hi = upload(ctx, hi);
    }
    t->scale = inverse ? gl::inv(gl::from_u64(1ULL << L)) : 1;
    ctx->ntt_tables[key] = t;
    return t;
}

static NttTables::Coset& get_coset(sezkp_ctx* ctx, NttTables* t, int logB, u64 shift) {
    auto key = std::make_pair(logB, shift);
    auto it = t->cosets.find(key);
    if (it != t->cosets.end()) return it->second;
    const int L = t->L, B = 1 << logB;
    const int b1 = t->plan[0];
    const u64 N1 = 1ULL << b1, S1 = 1ULL << (L - b1);
    const u64 wBig = gl::root_2exp((unsigned)(L + logB));
    std::vector<u64> ga((size_t)B * N1), gb((size_t)B * S1);
    u64 g = shift;  // g_j = shift * wBig^j
    for (int j = 0; j < B; j++) {
        u64 x = 1;
        for (u64 J = 0; J < S1; J++) {
            gb[(size_t)j * S1 + J] = x;
            x = gl::mul(x, g);
        }
        const u64 gS = x;  // g^S1
        x = 1;
        for (u64 r = 0; r < N1; r++) {
            ga[(size_t)j * N1 + r] = x;
            x = gl::mul(x, gS);
        }
        g = gl::mul(g, wBig);
    }
    NttTables::Coset c;
    c.GA = upload(ctx, ga);
    c.GB = upload(ctx, gb);
    c.ga_pitch = (u32)N1;
    c.gb_pitch = (u32)S1;
    return t->cosets[key] = c;
}

void ntt_free_tables(sezkp_ctx* ctx) {
    for (auto& kv : ctx->ntt_tables) {
        NttTables* t = kv.second;
        for (auto& w : t->W)
            if (w) cudaFree(w);
        cudaFree(t->tw_lo);
        cudaFree(t->tw_hi);
        for (auto& f : t->tw_full)
            if (f) cudaFree(f);
        for (auto& c : t->cosets) {
            cudaFree(c.second.GA);
            cudaFree(c.second.GB);
        }
        delete t;
    }
    ctx->ntt_tables.clear();
}

/* ------------------------------------------------------------------------------------------ */
/* pass construction + launch                                                                  */
/* ------------------------------------------------------------------------------------------ */
// gen 4 runs the widest passes (b >= 9: 32-point register DFTs, 64 data registers per thread) with 128-thread CTAs on
// 4096-element tiles: 128 registers per thread without spills and four independent CTAs per SM
static int pass_threads(const sezkp_ctx* ctx, int b) { return ctx->ntt_gen == 4 && b >= 9 ? 128 : NTT_THREADS; }
static int pick_logR(const sezkp_ctx* ctx, int b, u64 avail_cols, int min_logR) {
    int lr = TILE_LOG_ELEMS - (pass_threads(ctx, b) == 128 ? 1 : 0) - b;
    if (lr > MAX_LOG_R) lr = MAX_LOG_R;
    while (lr > 0 && (1ULL << lr) > avail_cols) lr--;
    if (lr < min_logR) lr = min_logR;
    return lr;
}

static void launch_pass(sezkp_ctx* ctx, PassDesc& d, u64 V) {
    const u64 tiles = (u64)d.col_tiles * d.U * V;
    REQUIRE(tiles > 0 && tiles < (1ULL << 31), "NTT grid too large (%llu tiles)", (unsigned long long)tiles);
    d.pitch = tile_pitch(d.b, d.logR);
    const size_t smem = (((size_t)d.pitch << d.logR) + (d.b > 5 ? ((size_t)1 << d.b) : 0)) * 8;
    pass_fn fn = kernel_for_bits(d.b, d.inv != 0, ctx->ntt_gen);
    if (d.fuse_upper) {  // K7: last LDE pass + labeled leaf hash + five tree levels
        REQUIRE(!d.inv && (d.b == 7 || d.b == 8) && d.b + d.logR == TILE_LOG_ELEMS && d.logR >= 5, "internal: fused LDE/hash pass shape");
        fn = d.b == 8 ? (pass_fn)lde_hash_pass_kernel<4, 4> : (pass_fn)lde_hash_pass_kernel<4, 3>;
    }
    // max dynamic smem already granted per kernel — per context: the attribute is per device, and the contexts of a
    // group run on concurrent threads
    size_t& granted = ctx->func_smem[(const void*)fn];
    if (smem > 48 * 1024 && granted < smem) {
        CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        granted = smem;
    }
    fn<<<(unsigned)tiles, d.fuse_upper ? NTT_THREADS : pass_threads(ctx, d.b), smem, ctx->stream>>>(d);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

constexpr int FULL_TW_MAX_LOG = 24;  // full twiddle matrices up to 2^24 entries (128 MiB) per pass
static void attach_full_twiddles(sezkp_ctx* ctx, NttTables* t, PassDesc& d, int p, u64 rows, u64 Sp, u64 stride) {
    const u64 count = rows * Sp;
    d.tw_extent = count;
    if (count > (1ULL << FULL_TW_MAX_LOG)) return;
    if (!t->tw_full[p]) {
        u64* buf = nullptr;
        if (cudaMalloc(&buf, count * 8) != cudaSuccess) {
            cudaGetLastError();
            return;  // fall back to the two-level tables
        }
        const bool fold = t->inverse && p == 0;
        tw_fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(buf, count, Sp, stride, t->tw_lo, t->tw_hi, t->tw_lb,
                                                                                fold ? t->scale : 1);
        ctx->launches++;
        t->tw_full[p] = buf;
        if (fold) t->tw0_scaled = true;
    }
    d.tw_full = t->tw_full[p];
    d.tw_pitch = Sp;
}

static void base_desc(PassDesc& d, const NttTables* t, int b) {
    d = PassDesc{};
    d.b = b;
    d.inv = t->inverse ? 1 : 0;
    d.W = t->W[b];
    d.tw_lo = t->tw_lo;
    d.tw_hi = t->tw_hi;
    d.tw_lb = t->tw_lb;
    d.U = 1;
}

// Plain batched NTT: `cols` vectors of length 2^L at `data` (stride 2^L), result back in `data`.
// `tmp` must hold cols*2^L elements when L > MAX_PASS_BITS.
void ntt_batch_device(sezkp_ctx* ctx, u64* data, u64* tmp, int L, u64 cols, bool inverse) {
    REQUIRE(L >= 0 && L <= 3 * MAX_PASS_BITS, "log_n %d out of range", L);
    if (L == 0 || cols == 0) return;
    NttTables* t = get_tables(ctx, L, inverse);
    const std::vector<int>& plan = t->plan;
    const int m = (int)plan.size();
    const u64 N = 1ULL << L;
    PassDesc d;
    if (m == 1) {  // case C: columns of the tile are batch vectors
        base_desc(d, t, L);
        d.in = data;
        d.out = data;
        d.in_extent = d.out_extent = cols * N;
        d.logR = pick_logR(ctx, L, cols, 0);
        d.col_tiles = (u32)((cols + (1ULL << d.logR) - 1) >> d.logR);
        d.total_cols = cols;
        d.in_row_stride = 1;
        d.in_cs_hi = N;
        d.out_row_stride = 1;
        d.out_cs_hi = N;
        d.load_rows_fast = 1;
        d.store_rows_fast = 1;
        d.use_scale = inverse;
        d.scale = t->scale;
        d.canon_in = 1;
        launch_pass(ctx, d, 1);
        return;
    }
    REQUIRE(tmp != nullptr, "internal: NTT temp buffer missing");
    // buffers: m==2: data -> tmp -> data ; m==3: data -> data -> tmp -> data
    u64 S = N;  // S_{p-1}
    for (int p = 0; p < m; p++) {
        const int b = plan[p];
        const u64 Np = 1ULL << b;
        const u64 Mp = S;       // size of the current sub-problem
        const u64 Sp = Mp / Np; // stride of digit p
        base_desc(d, t, b);
        const bool last = (p == m - 1);
        if (!last) {  // case A
            d.in = (m == 3 && p == 1) ? data : data;
            d.out = (m == 2) ? tmp : (p == 0 ? data : tmp);
            d.logR = pick_logR(ctx, b, Sp, 0);
            d.col_tiles = (u32)(Sp >> d.logR);
            d.total_cols = Sp;
            d.U = (u32)(N / Mp);
            d.in_row_stride = Sp;
            d.in_cs_hi = 1;
            d.in_u_stride = Mp;
            d.in_v_stride = N;
            d.out_row_stride = Sp;
            d.out_cs_hi = 1;
            d.out_u_stride = Mp;
            d.out_v_stride = N;
            d.use_tw = 1;
            d.tw_stride = N / Mp;
            d.canon_in = (p == 0);
            attach_full_twiddles(ctx, t, d, p, Np, Sp, N / Mp);
        } else {  // case B: columns are values of k_1, rows are contiguous
            const u64 N1 = 1ULL << plan[0], S1 = N / N1;
            d.in = tmp;
            d.out = data;
            d.logR = pick_logR(ctx, b, N1, 0);
            d.col_tiles = (u32)(N1 >> d.logR);
            d.total_cols = N1;
            d.in_row_stride = 1;
            d.in_cs_hi = S1;
            d.in_v_stride = N;
            d.out_row_stride = N / Np;
            d.out_cs_hi = 1;
            d.out_v_stride = N;
            if (m == 3) {
                d.U = (u32)(1ULL << plan[1]);
                d.in_u_stride = Np;     // k_2 * N_3
                d.out_u_stride = N1;    // k_2 * N_1
            }
            d.load_rows_fast = 1;
            d.use_scale = inverse && !t->tw0_scaled;
            d.scale = t->scale;
        }
        d.in_extent = d.out_extent = cols * N;
        launch_pass(ctx, d, cols);
        S = Sp;
    }
}

// Coset LDE: coeffs [cols][n] -> out [cols][B*n]; inter must hold cols*B*n elements when log n > MAX_PASS_BITS.
bool lde_hash_fusable(int L, int logB) {  // the last pass must be 7 or 8 bits wide with at least 32 tile columns
    if (L <= 2 * MAX_PASS_BITS) return false;
    const std::vector<int> plan = make_plan(L);
    const int b = plan.back();
    return (b == 7 || b == 8) && ((u64)1 << (plan[0] + logB)) >= ((u64)1 << (TILE_LOG_ELEMS - b)) && TILE_LOG_ELEMS - b >= 5 && logB <= TILE_LOG_ELEMS - b;
}
void coset_lde_device(sezkp_ctx* ctx, const u64* coeffs, u64* out, u64* inter, int L, int logB, u64 shift, u64 cols, const LdeHashFuse* fuse) {
    REQUIRE(L >= 1 && L <= 3 * MAX_PASS_BITS, "log_n %d out of range", L);
    REQUIRE(logB >= 0 && logB <= 4 && L + logB <= 32, "log_blow %d out of range", logB);
    REQUIRE(shift != 0 && shift < gl::P, "coset shift must be a non-zero canonical field element");
    if (cols == 0) return;
    NttTables* t = get_tables(ctx, L, false);
    NttTables::Coset& cs = get_coset(ctx, t, logB, shift);
    const std::vector<int>& plan = t->plan;
    const int m = (int)plan.size();
    const u64 n = 1ULL << L, B = 1ULL << logB;
    PassDesc d;
    if (m == 1) {  // case C': columns are (vector, coset) pairs
        base_desc(d, t, L);
        d.in = coeffs;
        d.out = out;
        d.in_extent = cols * n;
        d.out_extent = cols * B * n;
        d.total_cols = cols * B;
        d.logR = pick_logR(ctx, L, d.total_cols, 0);
        d.col_tiles = (u32)((d.total_cols + (1ULL << d.logR) - 1) >> d.logR);
        d.in_row_stride = 1;
        d.in_clog = logB;
        d.in_cs_lo = 0;
        d.in_cs_hi = n;
        d.out_row_stride = B;
        d.out_clog = logB;
        d.out_cs_lo = 1;
        d.out_cs_hi = B * n;
        d.load_rows_fast = 1;
        d.store_rows_fast = 1;
        d.use_pre = 1;
        d.use_gb = 0;
        d.coset_from_col = 1;
        d.coset_log = logB;
        d.GA = cs.GA;
        d.GB = cs.GB;
        d.ga_pitch = cs.ga_pitch;
        d.gb_pitch = cs.gb_pitch;
        launch_pass(ctx, d, 1);
        return;
    }
    REQUIRE(inter != nullptr, "internal: LDE intermediate buffer missing");
    u64 S = n;
    for (int p = 0; p < m; p++) {
        const int b = plan[p];
        const u64 Np = 1ULL << b, Mp = S, Sp = Mp / Np;
        base_desc(d, t, b);
        const bool last = (p == m - 1);
        if (!last) {  // case A / A'
            d.in = (p == 0) ? coeffs : inter;
            d.out = inter;
            d.in_extent = (p == 0) ? cols * n : cols * B * n;
            d.out_extent = cols * B * n;
            d.logR = pick_logR(ctx, b, Sp, 0);
            d.col_tiles = (u32)(Sp >> d.logR);
            d.total_cols = Sp;
            d.U = (u32)(n / Mp);
            d.in_row_stride = Sp;
            d.in_cs_hi = 1;
            d.in_u_stride = Mp;
            d.in_v_stride = n;
            d.in_v_shift = (p == 0) ? logB : 0;
            d.out_row_stride = Sp;
            d.out_cs_hi = 1;
            d.out_u_stride = Mp;
            d.out_v_stride = n;
            d.use_tw = 1;
            d.tw_stride = n / Mp;
            attach_full_twiddles(ctx, t, d, p, Np, Sp, n / Mp);
            if (p == 0) {
                d.use_pre = 1;
                d.use_gb = 1;
                d.coset_from_col = 0;
                d.coset_log = logB;
                d.GA = cs.GA;
                d.GB = cs.GB;
                d.ga_pitch = cs.ga_pitch;
                d.gb_pitch = cs.gb_pitch;
            }
            launch_pass(ctx, d, cols * B);
        } else {  // case B': columns are (coset j, k_1) pairs so that stores are contiguous
            const u64 N1 = 1ULL << plan[0], S1 = n / N1;
            d.in = inter;
            d.out = out;
            d.in_extent = d.out_extent = cols * B * n;
            d.total_cols = N1 * B;
            d.logR = pick_logR(ctx, b, d.total_cols, logB);
            d.col_tiles = (u32)(d.total_cols >> d.logR);
            d.in_row_stride = 1;
            d.in_clog = logB;
            d.in_cs_lo = n;   // coset j -> intermediate vector j of this batch entry
            d.in_cs_hi = S1;  // k_1
            d.in_v_stride = B * n;
            d.out_row_stride = (n / Np) * B;
            d.out_cs_hi = 1;
            d.out_v_stride = B * n;
            if (m == 3) {
                d.U = (u32)(1ULL << plan[1]);
                d.in_u_stride = Np;
                d.out_u_stride = N1 * B;
            }
            d.load_rows_fast = 1;
            if (fuse) {
                REQUIRE(lde_hash_fusable(L, logB), "internal: LDE of 2^%d x 2^%d cannot be fused with the leaf hash", L, logB);
                d.fuse_tpl = (const b3::LabelTemplate*)fuse->templates;
                d.fuse_upper = fuse->upper;
                d.fuse_col_words = fuse->col_words;
                d.fuse_extent = cols * fuse->col_words;
            }
            launch_pass(ctx, d, cols);
        }
        S = Sp;
    }
}
