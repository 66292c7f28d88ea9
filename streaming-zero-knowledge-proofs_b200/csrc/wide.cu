// BASELINE.json configs[3] / SURVEY.md §8(d) "config 4": LDE + labeled BLAKE3 column commitment + FRI over wide column
// sets (e.g. 256 columns x 2^24 rows), column-sharded over the GPUs of a context group.
//
//   per column    interpolate_from_evals (sezkp-ffts/src/ntt.rs:173-177) -> evaluate_on_coset_pow2 (coset.rs:85-102) ->
//                 hash_field_leaves_labeled over the extended column (v1/merkle.rs:132-146) -> chunked tree root
//                 (v1/openings.rs:306-398); column c lives on (and is committed by) GPU c % world
//   C1            all-gather of the 32-byte column roots
//   transcript    new("sezkp-stark/v1"), n, n_cols, col_root x n_cols -> "alphas" (8 B each) (sezkp-crypto/src/lib.rs:74-124)
//   C3            C(i) = sum_c alpha_c * col_c[i]: every GPU sums its own columns, then ONE kernel per GPU adds the partial
//                 vectors of all GPUs, reading the peers' HBM directly over NVLink (no staging copy)
//   tail          OOD point + nudge (v1/prover.rs:118-135), DEEP coset LDE (v1/lde.rs:42-97), FRI fold + commit
//                 (v1/prover.rs:184-243) with the leaf / subtree hashing of the large layers split by chunk range (C2)
// Extended columns are never all resident: they live in a scratch buffer that is reused column group by column group.
#include <string>

#include "gl.cuh"
#include "group.cuh"
#include "hash.cuh"
#include "ntt.cuh"
#include "stark.cuh"
#include "wide.cuh"

namespace {

// SURVEY §8(d) config-4 generator: value(c, i) = one splitmix64 step from state 0x5EED ^ (c << 40) ^ i, mod p.
__global__ void __launch_bounds__(256) wide_synth_kernel(u64* __restrict__ out, u64 n, u64 seed, u64 col0, u64 col_step, u64 n_local) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n_local) return;
    const u64 j = idx / n, i = idx - j * n;
    const u64 c = col0 + j * col_step;
    u64 z = (seed ^ (c << 40) ^ i) + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    out[idx] = z >= gl::P ? z - gl::P : z;  // z < 2^64 < 2p
}

constexpr int WIDE_MAX_COLS_PER_LAUNCH = 32;
struct WideAlphas {
    u64 a[WIDE_MAX_COLS_PER_LAUNCH];
    int count;
    int accumulate;  // add to out instead of overwriting
};
// out[i] (+)= sum_j a[j] * cols[j][i]: one pass over `count` resident columns (HBM-bound: 8 B per column element).
__global__ void __launch_bounds__(256) wide_partial_kernel(const u64* __restrict__ cols, u64 n, const WideAlphas wa, u64* __restrict__ out) {
    namespace L = gl::lazy;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 acc = wa.accumulate ? out[i] : 0;
#pragma unroll 4
    for (int j = 0; j < wa.count; j++) acc = L::add1(acc, L::mul(wa.a[j], cols[(u64)j * n + i]));  // product canonical: single-correction add
    out[i] = L::canon(acc);
}
constexpr int WIDE_MAX_WORLD = 64;
struct PeerPtrs {
    const u64* p[WIDE_MAX_WORLD];
    int world;
};
// base[i] = sum_r partial_r[i] (mod p), partial_r in the HBM of GPU r: peer loads over NVLink, coalesced 8-byte reads.
__global__ void __launch_bounds__(256) wide_peer_sum_kernel(const PeerPtrs pp, u64 n, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 acc = 0;
    for (int r = 0; r < pp.world; r++) acc = gl::add(acc, pp.p[r][i]);
    out[i] = acc;
}

inline unsigned blocks_for(u64 n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void WideColumns::release() {
    evals.release();
    n_local = 0;
}

void wide_columns_synth(sezkp_ctx* ctx, WideColumns& wc, u64 seed, int c, int log_n, int rank, int world) {
    wc.c = c;
    wc.log_n = log_n;
    wc.rank = rank;
    wc.world = world;
    wc.n_local = (c - rank + world - 1) / world;
    if (wc.n_local <= 0) {
        wc.n_local = 0;
        return;
    }
    const u64 n = 1ULL << log_n;
    u64* d = (u64*)wc.evals.ensure((size_t)wc.n_local * n * 8);
    wide_synth_kernel<<<blocks_for(n * wc.n_local, 256), 256, 0, ctx->stream>>>(d, n, seed, (u64)rank, (u64)world, (u64)wc.n_local);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void wide_columns_upload(sezkp_ctx* ctx, WideColumns& wc, const u64* evals_host, int c, int log_n, int rank, int world) {
    wc.c = c;
    wc.log_n = log_n;
    wc.rank = rank;
    wc.world = world;
    wc.n_local = (c - rank + world - 1) / world;
    if (wc.n_local <= 0) {
        wc.n_local = 0;
        return;
    }
    const u64 n = 1ULL << log_n;
    u64* d = (u64*)wc.evals.ensure((size_t)wc.n_local * n * 8);
    for (int j = 0; j < wc.n_local; j++)
        CUDA_CHECK(cudaMemcpyAsync(d + (u64)j * n, evals_host + (u64)(rank + j * world) * n, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the caller's host array may go away
}

// LDE + commit of `c` resident columns: per column group iNTT -> coset LDE -> labeled leaves -> tree.  Roots go to
// d_roots (device, [c][32]).  By default (option "lde_fuse", three-pass sizes) the LDE's last pass hashes its own outputs
// (lde_hash_pass_kernel, ntt.cu) and the extended columns never reach HBM; otherwise they go through a scratch buffer that
// is reused group by group and are hashed by the chunk kernel.  DESIGN.md §4.4 has the A/B measurement.
void lde_commit_columns(sezkp_ctx* ctx, const u64* evals_dev, const char* const* labels, int c, int log_n, int log_blow, u64 shift,
                        int chunk_log2, u8* d_roots) {
    const size_t n = (size_t)1 << log_n, N = n << log_blow;
    size_t group = ((size_t)2 << 30) / (N * 8);  // ~2 GiB of extended values per group
    if (group < 1) group = 1;
    if (group > (size_t)c) group = (size_t)c;
    u64* coeffs = (u64*)ctx->scratch[4].ensure(group * n * 8);
    u64* tmp = log_n > 10 ? (u64*)ctx->scratch[0].ensure(group * n * 8) : nullptr;
    u64* inter = log_n > 10 ? (u64*)ctx->scratch[1].ensure(group * N * 8) : nullptr;
    // K7 (option "lde_fuse"): the LDE's last pass hashes its outputs and writes 32-leaf sub-roots; the extended values never
    // reach HBM.  Same roots either way.
    const bool fuse = ctx->lde_fuse && chunk_log2 == 10 && lde_hash_fusable(log_n, log_blow);
    u64* ext = fuse ? nullptr : (u64*)ctx->scratch[5].ensure(group * N * 8);
    for (size_t c0 = 0; c0 < (size_t)c; c0 += group) {
        const size_t g = (c0 + group <= (size_t)c) ? group : (size_t)c - c0;
        CUDA_CHECK(cudaMemcpyAsync(coeffs, evals_dev + c0 * n, g * n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        ntt_batch_device(ctx, coeffs, tmp, log_n, g, true);
        if (fuse) {
            Commit cm;
            CommitOpts o;
            o.roots_dev = d_roots + c0 * 32;
            try {
                commit_begin(ctx, cm, nullptr, N, (int)g, 5, labels + c0, o);
                LdeHashFuse f{cm.templates, cm.upper, (2 * cm.n_ch - 1) * 8};
                coset_lde_device(ctx, coeffs, nullptr, inter, log_n, log_blow, shift, g, &f);
                commit_finish(ctx, cm, o);
            } catch (...) {
                cm.release(ctx);
                throw;
            }
            cm.release(ctx);
            continue;
        }
        coset_lde_device(ctx, coeffs, ext, inter, log_n, log_blow, shift, g);
        Commit cm;
        CommitOpts o;
        o.dedup = false;  // extended values are high-entropy
        o.roots_dev = d_roots + c0 * 32;
        int cl = chunk_log2;
        if (chunk_log2 == 10 && N >= ((size_t)1 << 20)) {  // roots only: 32-leaf sub-roots out of each 1024-leaf CTA (see stark.cu)
            o.cta_log2 = 10;
            cl = 5;
        }
        try {
            commit_build(ctx, cm, ext, N, (int)g, cl, labels + c0, o);
        } catch (...) {
            cm.release(ctx);
            throw;
        }
        cm.release(ctx);
    }
}

void wide_commit_fri_rank(sezkp_ctx* ctx, const WideColumns& wc, const char* const* labels, int log_blow, u64 shift, int chunk_log2,
                          u8* col_roots_out, u8* fri_roots_out, u64* final_value, WideTaps* taps) {
    const int c = wc.c, L = wc.log_n, rank = wc.rank, world = wc.world, log_N = L + log_blow;
    REQUIRE(c >= 1 && L >= 1 && L <= 29 && log_blow >= 0 && log_blow <= 4 && log_N <= 32, "bad shape (c=%d, log_n=%d, log_blow=%d)", c, L, log_blow);
    REQUIRE(shift != 0 && shift < gl::P, "coset shift must be a non-zero canonical field element");
    REQUIRE(world >= 1 && world <= WIDE_MAX_WORLD, "world %d out of range", world);
    sezkp_group* grp = ctx->group;
    REQUIRE(world == 1 || (grp != nullptr && grp->world == world), "column-sharded commit needs a context group of %d GPUs", world);
    GroupRank* gr = world > 1 ? &grp->ranks[rank] : nullptr;
    const u64 n = 1ULL << L, N = 1ULL << log_N;
    const int n_local = wc.n_local, max_local = (c + world - 1) / world;
    ctx->timings.clear();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // device-side duration of this rank's part (reported next to the host laps)
    CUDA_CHECK(cudaEventCreate(&ev0));
    CUDA_CHECK(cudaEventCreate(&ev1));
    struct EvGuard {
        cudaEvent_t a, b;
        ~EvGuard() {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    } ev_guard{ev0, ev1};
    CUDA_CHECK(cudaEventRecord(ev0, ctx->stream));
    double t0 = wide_now_ms();
    const double t_begin = t0;
    auto lap = [&](const char* name) {
        cudaStreamSynchronize(ctx->stream);
        const double t1 = wide_now_ms();
        ctx->timings.push_back({name, t1 - t0});
        t0 = t1;
    };

    // 1. local columns: LDE + labeled commit
    std::vector<const char*> local_labels;
    for (int j = 0; j < n_local; j++) local_labels.push_back(labels[rank + j * world]);
    std::vector<u8> local_roots((size_t)max_local * 32, 0);
    u8* d_roots = (u8*)ctx->scratch[10].ensure((size_t)(max_local > log_N + 1 ? max_local : log_N + 1) * 32 + 64);
    if (n_local) {
        lde_commit_columns(ctx, wc.evals.as<u64>(), local_labels.data(), n_local, L, log_blow, shift, chunk_log2, d_roots);
        CUDA_CHECK(cudaMemcpyAsync(local_roots.data(), d_roots, (size_t)n_local * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    lap("lde_commit");

    // 2. C1: every rank learns all column roots (canonical column order)
    std::vector<u8> col_roots((size_t)c * 32);
    if (world == 1) std::memcpy(col_roots.data(), local_roots.data(), (size_t)c * 32);
    else {
        std::vector<u8> all((size_t)world * max_local * 32);
        if (group_allgather_host(gr, local_roots.data(), local_roots.size(), all.data()) != 0)
            sezkp_fail(SEZKP_CUDA_ECOMM, "column-root exchange failed");
        for (int k = 0; k < c; k++) std::memcpy(&col_roots[32 * (size_t)k], &all[((size_t)(k % world) * max_local + k / world) * 32], 32);
    }
    if (col_roots_out) std::memcpy(col_roots_out, col_roots.data(), col_roots.size());

    // 3. transcript -> one alpha per column
    host::Transcript tr("sezkp-stark/v1");
    tr.absorb_u64("n", n);
    tr.absorb_u64("n_cols", (u64)c);
    for (int k = 0; k < c; k++) tr.absorb("col_root", &col_roots[32 * (size_t)k], 32);
    std::vector<u64> alphas(c);
    {
        auto by = tr.challenge("alphas", 8 * (size_t)c);
        for (int k = 0; k < c; k++) alphas[k] = le64(&by[8 * (size_t)k]) % gl::P;
    }

    // 4. C3: combination on the base domain
    u64* base = (u64*)ctx->scratch[4].ensure(n * 8);
    u64* partial = world > 1 ? (u64*)ctx->pool.alloc(n * 8) : base;  // peers read it: pool memory is never cudaFree'd under them
    if (n_local == 0) CUDA_CHECK(cudaMemsetAsync(partial, 0, n * 8, ctx->stream));
    for (int j0 = 0; j0 < n_local; j0 += WIDE_MAX_COLS_PER_LAUNCH) {
        WideAlphas wa{};
        wa.count = n_local - j0 < WIDE_MAX_COLS_PER_LAUNCH ? n_local - j0 : WIDE_MAX_COLS_PER_LAUNCH;
        wa.accumulate = j0 > 0;
        for (int j = 0; j < wa.count; j++) wa.a[j] = alphas[rank + (j0 + j) * world];
        wide_partial_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(wc.evals.as<u64>() + (u64)j0 * n, n, wa, partial);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
    }
    if (world > 1) {
        if (grp->p2p) {  // one kernel: sum the partial vectors straight out of the peers' HBM
            PeerPtrs pp{};
            pp.world = world;
            const void* all[WIDE_MAX_WORLD];
            group_publish_peers(gr, partial, ctx->stream, all);
            for (int r = 0; r < world; r++) pp.p[r] = (const u64*)all[(rank + r) % world];  // start at a different peer per rank
            wide_peer_sum_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(pp, n, base);
            CUDA_CHECK(cudaGetLastError());
            ctx->launches++;
            group_release_peers(gr, ctx->stream);
        } else {  // no peer access: all-gather the partial vectors (staged copies), then the same sum locally
            u64* gathered = (u64*)ctx->pool.alloc((size_t)world * n * 8);
            if (group_allgather_dev(gr, partial, n * 8, gathered, (void*)ctx->stream) != 0)
                sezkp_fail(SEZKP_CUDA_ECOMM, "partial-sum exchange failed");
            PeerPtrs pp{};
            pp.world = world;
            for (int r = 0; r < world; r++) pp.p[r] = gathered + (u64)r * n;
            wide_peer_sum_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(pp, n, base);
            CUDA_CHECK(cudaGetLastError());
            ctx->launches++;
            ctx->pool.free(gathered);  // stream-ordered reuse: later allocations are used on the same stream
        }
        ctx->pool.free(partial);
    }
    lap("combine");

    // 5. OOD point, nudged off the coset (v1/prover.rs:118-135)
    u64 z = le64(tr.challenge("ood_point", 8).data()) % gl::P;
    while (z_on_coset(z, shift, log_N)) z = gl::add(z, 1);

    // 6. DEEP coset LDE straight into FRI layer-0 storage, then fold + commit
    FriLayers fl;
    u64* coset_local = nullptr;
    try {
        fl.values = (u64*)ctx->pool.alloc(2 * N * 8);
        // context group with peer access, blow-up 8: coset-resident FRI layers (stark.cu) — each GPU evaluates its 8/world
        // cosets of the DEEP-LDE and folds them locally instead of every GPU computing the whole single-vector tail
        const bool coset = world > 1 && grp && grp->p2p && log_blow == 3 && 8 % world == 0 && ctx->fri_coset && log_N >= 20;
        if (coset) coset_local = deep_lde_coset_local_device(ctx, base, L, log_blow, shift, z, rank, world);
        else deep_lde_device(ctx, base, fl.values, L, log_blow, shift, z);
        lap("deep_lde");
        std::vector<u8> fri_roots((size_t)(log_N + 1) * 32);
        u64 fin = 0;
        TranscriptAbsorb ab(tr);
        ShardInfo sh{rank, world, group_allgather_host, gr};
        fri_commit_device(ctx, fl, fl.values, log_N, nullptr, fri_roots.data(), &fin, &ab, world > 1 ? &sh : nullptr, coset_local);
        if (coset_local) {
            ctx->pool.free(coset_local);
            coset_local = nullptr;
        }
        lap("fri_commit");
        if (fri_roots_out) std::memcpy(fri_roots_out, fri_roots.data(), fri_roots.size());
        if (final_value) *final_value = fin;
        if (taps) {
            taps->alphas = alphas;
            taps->z = z;
        }
    } catch (...) {
        if (coset_local) ctx->pool.free(coset_local);
        fl.release(ctx);
        throw;
    }
    fl.release(ctx);
    CUDA_CHECK(cudaEventRecord(ev1, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ev1));
    float dev_ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&dev_ms, ev0, ev1));
    ctx->timings.push_back({"device_ms", (double)dev_ms});
    ctx->timings.push_back({"total", wide_now_ms() - t_begin});
}
