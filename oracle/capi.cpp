// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/gl.hpp header note).
// C entry points over the CPU restatement so tests/ and bench.py's cpu_baseline / reference arm can
// drive it through ctypes.  Every function returns 0 on success, <0 on error (message via
// oracle_last_error()).
#include <cstdio>
#include <string>

#include "stark.hpp"
#include "wide.hpp"

using namespace oracle;

static thread_local std::string g_err;
#define ORACLE_TRY(...)                      \
    try {                                    \
        __VA_ARGS__;                         \
        return 0;                            \
    } catch (const std::exception& e) {      \
        g_err = e.what();                    \
        return -1;                           \
    }

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

int oracle_blake3(const uint8_t* data, size_t n, uint8_t* out, size_t out_len) {
    ORACLE_TRY({
        Blake3 h;
        h.update(data, n);
        h.finalize_xof(out, out_len);
    })
}

/* ---- transcript handle ---- */
void* oracle_transcript_new(const char* domain) { return new Transcript(domain); }
void oracle_transcript_free(void* t) { delete (Transcript*)t; }
int oracle_transcript_absorb(void* t, const char* label, const uint8_t* bytes, size_t n) {
    ORACLE_TRY(((Transcript*)t)->absorb(label, bytes, n))
}
int oracle_transcript_challenge(void* t, const char* label, uint8_t* out, size_t n) {
    ORACLE_TRY({
        auto v = ((Transcript*)t)->challenge_bytes(label, n);
        std::memcpy(out, v.data(), n);
    })
}

/* ---- field / NTT ---- */
uint64_t oracle_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
uint64_t oracle_gl_inv(uint64_t a) { return gl_inv(a); }
uint64_t oracle_gl_pow(uint64_t a, uint64_t e) { return gl_pow(a, e); }
uint64_t oracle_gl_from_i64(int64_t x) { return gl_from_i64(x); }
uint64_t oracle_gl_root_2exp(unsigned k) { return gl_root_2exp(k); }

int oracle_ntt(uint64_t* data, int log_n, size_t cols, int inverse) {
    ORACLE_TRY({
        size_t n = (size_t)1 << log_n;
        for (size_t c = 0; c < cols; c++) {
            if (inverse) inverse_ntt_in_place(data + c * n, n);
            else forward_ntt_in_place(data + c * n, n);
        }
    })
}
int oracle_dft_naive(const uint64_t* in, int log_n, uint64_t* out) {
    ORACLE_TRY({
        size_t n = (size_t)1 << log_n;
        auto v = dft_naive(in, n, gl_root_2exp((unsigned)log_n));
        std::memcpy(out, v.data(), 8 * n);
    })
}
/* evaluate_on_coset_pow2 per column: coeffs [cols][m] -> out [cols][1<<k_log2] */
int oracle_coset_eval(const uint64_t* coeffs, size_t m, int k_log2, uint64_t shift, size_t cols, uint64_t* out) {
    ORACLE_TRY({
        size_t n = (size_t)1 << k_log2;
        for (size_t c = 0; c < cols; c++) {
            auto v = evaluate_on_coset_pow2(coeffs + c * m, m, (unsigned)k_log2, shift);
            std::memcpy(out + c * n, v.data(), 8 * n);
        }
    })
}
/* interpolate_from_evals then evaluate_on_coset_pow2, per column */
int oracle_lde_from_evals(const uint64_t* evals, int log_n, int log_blow, uint64_t shift, size_t cols, uint64_t* out) {
    ORACLE_TRY({
        size_t n = (size_t)1 << log_n, N = n << log_blow;
        for (size_t c = 0; c < cols; c++) {
            std::vector<u64> co(evals + c * n, evals + (c + 1) * n);
            inverse_ntt_in_place(co.data(), n);
            auto v = evaluate_on_coset_pow2(co.data(), n, (unsigned)(log_n + log_blow), shift);
            std::memcpy(out + c * N, v.data(), 8 * N);
        }
    })
}
int oracle_deep_lde(const uint64_t* base, int log_n, int log_blow, uint64_t shift, uint64_t z, uint64_t* out) {
    ORACLE_TRY({
        auto v = deep_coset_lde(base, (size_t)1 << log_n, (unsigned)log_blow, shift, z);
        std::memcpy(out, v.data(), 8 * v.size());
    })
}

/* ---- hashing / trees ---- */
int oracle_leaf_hash(const uint64_t* vals, size_t n, const char* label_or_null, uint8_t* out) {
    ORACLE_TRY({
        for (size_t i = 0; i < n; i++) {
            u8 le[8];
            gl_to_le(vals[i], le);
            Digest d = label_or_null ? hash_leaf_labeled(le, label_or_null) : hash_leaf(le);
            std::memcpy(out + 32 * i, d.data(), 32);
        }
    })
}
static std::vector<Digest> digests_from(const uint8_t* p, size_t n) {
    std::vector<Digest> v(n);
    for (size_t i = 0; i < n; i++) std::memcpy(v[i].data(), p + 32 * i, 32);
    return v;
}
int oracle_node_hash(const uint8_t* l, const uint8_t* r, uint8_t* out) {
    ORACLE_TRY({
        Digest a, b;
        std::memcpy(a.data(), l, 32);
        std::memcpy(b.data(), r, 32);
        Digest d = node_hash(a, b);
        std::memcpy(out, d.data(), 32);
    })
}
/* MerkleTree::from_leaves(..).root() (odd promotion; empty -> single zero leaf) */
int oracle_merkle_root(const uint8_t* leaves, size_t n, uint8_t* out) {
    ORACLE_TRY({
        Digest d = MerkleTree::from_leaves(digests_from(leaves, n)).root();
        std::memcpy(out, d.data(), 32);
    })
}
/* MerkleTree::open(idx): writes siblings bottom->top, returns count through n_sibs */
int oracle_merkle_open(const uint8_t* leaves, size_t n, size_t idx, uint8_t* sibs, size_t* n_sibs) {
    ORACLE_TRY({
        auto s = MerkleTree::from_leaves(digests_from(leaves, n)).open(idx);
        for (size_t i = 0; i < s.size(); i++) std::memcpy(sibs + 32 * i, s[i].data(), 32);
        *n_sibs = s.size();
    })
}
/* StreamingLayerBuilder root over unlabeled leaves of vals */
int oracle_streaming_layer_root(const uint64_t* vals, size_t n, uint8_t* out) {
    ORACLE_TRY({
        StreamingLayerBuilder b(n);
        for (size_t i = 0; i < n; i++) {
            u8 le[8];
            gl_to_le(vals[i], le);
            b.absorb_leaf(le);
        }
        Digest d = b.finalize();
        std::memcpy(out, d.data(), 32);
    })
}
/* sezkp_merkle::merkle_root (batch) and Frontier root over pre-hashed leaves */
int oracle_manifest_merkle_root(const uint8_t* leaves, size_t n, int use_frontier, uint8_t* out) {
    ORACLE_TRY({
        auto v = digests_from(leaves, n);
        Digest d;
        if (use_frontier) {
            Frontier f;
            for (auto& x : v) f.push_leaf(x);
            d = f.finalize_root();
        } else d = merkle_root(v);
        std::memcpy(out, d.data(), 32);
    })
}
/* sezkp_merkle::leaf_hash (sezkp-merkle/src/lib.rs:85-117) over one BlockSummary's scalar fields */
int oracle_manifest_leaf_hash(uint16_t version, uint32_t block_id, uint64_t step_lo, uint64_t step_hi, uint16_t ctrl_in,
                              uint16_t ctrl_out, int64_t in_head_in, int64_t in_head_out, uint64_t tau,
                              const int64_t* win_left, const int64_t* win_right, const uint32_t* in_off,
                              const uint32_t* out_off, uint64_t steps_len, uint8_t* out) {
    ORACLE_TRY({
        Blake3 h;
        h.update(&version, 2); h.update(&block_id, 4); h.update(&step_lo, 8); h.update(&step_hi, 8);
        h.update(&ctrl_in, 2); h.update(&ctrl_out, 2); h.update(&in_head_in, 8); h.update(&in_head_out, 8);
        h.update(&tau, 8);
        for (uint64_t r = 0; r < tau; r++) { h.update(&win_left[r], 8); h.update(&win_right[r], 8); }
        for (uint64_t r = 0; r < tau; r++) h.update(&in_off[r], 4);
        for (uint64_t r = 0; r < tau; r++) h.update(&out_off[r], 4);
        h.update(&steps_len, 8);
        h.finalize(out);
    })
}

/* ---- column commitments (a13/a14) over arbitrary columns: cols [c][n], labels[c] ---- */
int oracle_column_commit(const uint64_t* cols, const char* const* labels, size_t c, size_t n, int chunk_log2, uint8_t* roots) {
    ORACLE_TRY({
        size_t chunk = (size_t)1 << chunk_log2;
        for (size_t ci = 0; ci < c; ci++) {
            std::vector<Digest> croots, cur;
            for (size_t i = 0; i < n; i++) {
                u8 le[8];
                gl_to_le(cols[ci * n + i], le);
                cur.push_back(hash_leaf_labeled(le, labels[ci]));
                if (cur.size() == chunk) { croots.push_back(MerkleTree::from_leaves(cur).root()); cur.clear(); }
            }
            if (!cur.empty()) croots.push_back(MerkleTree::from_leaves(cur).root());
            Digest d = MerkleTree::from_leaves(croots).root();
            std::memcpy(roots + 32 * ci, d.data(), 32);
        }
    })
}

/* ---- trace columns / composition (feeder) ---- */
/* committed columns in canonical label order: out [3+7tau][n] */
int oracle_trace_columns(const sezkp_trace_desc* d, uint64_t* out) {
    ORACLE_TRY({
        TraceView tv(d);
        TraceColumns tc = TraceColumns::build(tv, false);
        size_t nc = 3 + 7 * tc.tau;
        for (size_t ci = 0; ci < nc; ci++) std::memcpy(out + ci * tc.n, tc.committed(ci).data(), 8 * tc.n);
    })
}
int oracle_compose_base(const sezkp_trace_desc* d, const uint64_t alphas8[8], const uint64_t* mask_coeffs, size_t mask_deg, uint64_t* out) {
    ORACLE_TRY({
        TraceView tv(d);
        TraceColumns tc = TraceColumns::build(tv, true);
        std::vector<std::vector<u64>> masks{std::vector<u64>(mask_coeffs, mask_coeffs + mask_deg)};
        auto v = compose_base(tc, Alphas::from8(alphas8), masks);
        std::memcpy(out, v.data(), 8 * v.size());
    })
}

/* ---- FRI fold+commit for given betas: roots [log_N+1][32], final value ---- */
int oracle_fri_commit(const uint64_t* layer0, int log_N, const uint64_t* betas, uint8_t* roots, uint64_t* final_value) {
    ORACLE_TRY({
        size_t N = (size_t)1 << log_N;
        std::vector<u64> l0(layer0, layer0 + N), b(betas, betas + log_N);
        auto layers = fri_fold_layers(l0, b);
        for (size_t l = 0; l < layers.size(); l++) {
            Digest d = MerkleTree::from_leaves(leaves_of(layers[l].data(), layers[l].size())).root();
            std::memcpy(roots + 32 * l, d.data(), 32);
        }
        *final_value = layers.back()[0];
    })
}


/* ---- config 4 (SURVEY §8d "W"): wide LDE + labeled column commit + FRI (oracle/wide.hpp) ---- */
/* value(c, i) of the 0x5EED splitmix generator: out[1 << log_n] */
int oracle_wide_column(uint64_t c, int log_n, uint64_t* out) {
    ORACLE_TRY({
        size_t n = (size_t)1 << log_n;
        for (size_t i = 0; i < n; i++) out[i] = wide_value(c, i);
    })
}
/* iNTT -> coset LDE -> labeled leaves -> root of one column of base-domain evaluations */
int oracle_lde_commit_root(const uint64_t* evals, int log_n, int log_blow, uint64_t shift, const char* label, uint8_t* root) {
    ORACLE_TRY({
        Digest d = lde_commit_root(evals, (unsigned)log_n, (unsigned)log_blow, shift, label);
        std::memcpy(root, d.data(), 32);
    })
}
/* the same for column c of the generator, label "c_{c}" */
int oracle_wide_column_root(uint64_t c, int log_n, int log_blow, uint64_t shift, uint8_t* root) {
    ORACLE_TRY({
        size_t n = (size_t)1 << log_n;
        std::vector<u64> v(n);
        for (size_t i = 0; i < n; i++) v[i] = wide_value(c, i);
        Digest d = lde_commit_root(v.data(), (unsigned)log_n, (unsigned)log_blow, shift, "c_" + std::to_string(c));
        std::memcpy(root, d.data(), 32);
    })
}
/* transcript -> alphas -> combination -> z -> DEEP LDE -> FRI roots.  evals [n_cols][n] or NULL (generator columns).
 * Outputs: alphas[n_cols], z, betas[log_n+log_blow], fri_roots[(log_n+log_blow+1)][32], final value. */
int oracle_wide_tail(const uint64_t* evals, size_t n_cols, int log_n, int log_blow, uint64_t shift, const uint8_t* col_roots,
                     uint64_t* alphas, uint64_t* z, uint64_t* betas, uint8_t* fri_roots, uint64_t* final_value) {
    ORACLE_TRY({
        size_t n = (size_t)1 << log_n;
        auto roots = digests_from(col_roots, n_cols);
        auto col = [&](size_t c) {
            std::vector<u64> v(n);
            if (evals) std::memcpy(v.data(), evals + c * n, 8 * n);
            else for (size_t i = 0; i < n; i++) v[i] = wide_value(c, i);
            return v;
        };
        WideOut w = wide_tail(col, n_cols, (unsigned)log_n, (unsigned)log_blow, shift, roots);
        if (alphas) std::memcpy(alphas, w.alphas.data(), 8 * n_cols);
        if (z) *z = w.z;
        if (betas) std::memcpy(betas, w.betas.data(), 8 * w.betas.size());
        for (size_t l = 0; l < w.fri_roots.size(); l++) std::memcpy(fri_roots + 32 * l, w.fri_roots[l].data(), 32);
        *final_value = w.final_value;
    })
}

/* ---- full prover / verifier ---- */
struct OracleProveTrace {  /* optional taps for per-stage parity tests */
    uint64_t alphas[8];
    uint64_t mask_coeffs[4];
    uint64_t z;
    uint64_t betas[64];
    uint64_t rows[30];
    uint64_t fri_rows[30];
};
int oracle_prove_v1(const sezkp_trace_desc* d, const uint8_t manifest_root[32], int faithful_cost, uint8_t* buf, size_t cap,
                    size_t* len, OracleProveTrace* taps) {
    ORACLE_TRY({
        ProveTrace t;
        ProofV1 p = prove_v1(d, manifest_root, faithful_cost != 0, &t);
        auto bytes = bincode_proof(p);
        *len = bytes.size();
        if (taps) {
            std::memcpy(taps->alphas, t.alphas, 64);
            for (int i = 0; i < 4; i++) taps->mask_coeffs[i] = t.mask_coeffs[0][i];
            taps->z = t.z;
            for (size_t i = 0; i < t.betas.size() && i < 64; i++) taps->betas[i] = t.betas[i];
            for (size_t i = 0; i < 30; i++) { taps->rows[i] = t.rows[i]; taps->fri_rows[i] = t.fri_rows[i]; }
        }
        if (buf) {
            if (cap < bytes.size()) throw std::runtime_error("proof buffer too small");
            std::memcpy(buf, bytes.data(), bytes.size());
        }
    })
}
/* 0 = accept, 1 = reject (reason in oracle_last_error), -1 = malformed */
int oracle_verify_v1(const uint8_t* proof, size_t len, const sezkp_trace_desc* d) {
    try {
        ProofV1 p = bincode_parse(proof, len);
        std::string why = verify_v1(p, d);
        if (why.empty()) return 0;
        g_err = why;
        return 1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
