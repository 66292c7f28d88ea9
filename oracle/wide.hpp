// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/gl.hpp header note).
//
// CPU restatement of BASELINE.json configs[3] / SURVEY.md §8(d) "Config 4 (W)": the wide LDE + column
// commitment + FRI pipeline.  It is not a shape `prove_v1` produces; it chains reference functions only:
//   per column   interpolate_from_evals (sezkp-ffts/src/ntt.rs:173-177) -> evaluate_on_coset_pow2
//                (coset.rs:85-102, blow-up 2^log_blow, shift) -> hash_field_leaves_labeled over the
//                extended column (v1/merkle.rs:132-146) -> Merkle root (plain binary tree; equals the
//                chunked root of v1/openings.rs:306-398 for power-of-two lengths)
//   transcript   Blake3Transcript::new("sezkp-stark/v1"), absorb_u64("n"), absorb_u64("n_cols"),
//                absorb("col_root") per column, challenge_bytes("alphas", 8*n_cols) -> from_u64 each
//                (sezkp-crypto/src/lib.rs:74-124; derivers as v1/params.rs:76-126)
//   combination  C(i) = sum_c alpha_c * col_c[i] on the base domain
//   OOD point    derive_ood_point + nudge off the coset (v1/prover.rs:118-135)
//   DEEP LDE     deep_coset_lde_stream (v1/lde.rs:42-97)
//   FRI          absorb root0, derive_betas, fold/commit every layer (v1/prover.rs:184-243)
// Parity pinning: same status as oracle/stark.hpp (every stage is one of its functions).
#pragma once
#include "stark.hpp"

namespace oracle {

// SURVEY §8(d) config 4 generator: value(c, i) = one splitmix64 step from state 0x5EED ^ (c << 40) ^ i, mod p.
inline u64 wide_value(u64 c, u64 i) {
    u64 s = (0x5EEDULL ^ (c << 40) ^ i) + 0x9E3779B97F4A7C15ULL;
    u64 z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return z % GL_P;
}

// Streaming root over labeled leaves of a power-of-two vector (the per-level stack of fri_stream.rs:55-122 with the
// labeled leaf of merkle.rs:132-146): never holds more than log2(n) digests, so a 2^27-leaf column needs no 4 GiB tree.
inline Digest labeled_root_streaming(const u64* v, size_t n, const std::string& label) {
    if (n == 0 || (n & (n - 1))) throw std::invalid_argument("labeled_root_streaming: n must be a power of two");
    std::vector<std::pair<bool, Digest>> stack;
    for (size_t i = 0; i < n; i++) {
        u8 le[8];
        gl_to_le(v[i], le);
        Digest cur = hash_leaf_labeled(le, label);
        size_t lvl = 0;
        for (;;) {
            if (stack.size() <= lvl) stack.push_back({false, Digest{}});
            if (stack[lvl].first) {
                stack[lvl].first = false;
                cur = node_hash(stack[lvl].second, cur);
                lvl++;
            } else {
                stack[lvl] = {true, cur};
                break;
            }
        }
    }
    return stack.back().second;
}
inline Digest unlabeled_root_streaming(const u64* v, size_t n) {
    StreamingLayerBuilder b(n);
    for (size_t i = 0; i < n; i++) {
        u8 le[8];
        gl_to_le(v[i], le);
        b.absorb_leaf(le);
    }
    return b.finalize();
}

// iNTT -> coset LDE -> labeled leaves -> root of one column (evals: n = 2^log_n base-domain evaluations).
inline Digest lde_commit_root(const u64* evals, unsigned log_n, unsigned log_blow, u64 shift, const std::string& label) {
    const size_t n = (size_t)1 << log_n;
    std::vector<u64> co(evals, evals + n);
    inverse_ntt_in_place(co.data(), n);
    std::vector<u64> ext = evaluate_on_coset_pow2(co.data(), n, log_n + log_blow, shift);
    return labeled_root_streaming(ext.data(), ext.size(), label);
}

struct WideOut {
    std::vector<u64> alphas;
    u64 z = 0;
    std::vector<u64> betas;
    std::vector<Digest> fri_roots;  // log_N + 1
    u64 final_value = 0;
};
// Everything after the column roots.  col(c) must return the n base-domain evaluations of column c.
template <class ColFn>
inline WideOut wide_tail(ColFn col, size_t n_cols, unsigned log_n, unsigned log_blow, u64 shift, const std::vector<Digest>& col_roots) {
    const size_t n = (size_t)1 << log_n;
    const unsigned lde_k = log_n + log_blow;
    const size_t N = (size_t)1 << lde_k;
    WideOut out;
    Transcript tr("sezkp-stark/v1");
    tr.absorb_u64("n", n);
    tr.absorb_u64("n_cols", n_cols);
    for (size_t c = 0; c < n_cols; c++) tr.absorb("col_root", col_roots[c].data(), 32);
    {
        auto by = tr.challenge_bytes("alphas", 8 * n_cols);
        for (size_t c = 0; c < n_cols; c++) out.alphas.push_back(gl_from_u64(le64(&by[8 * c])));
    }
    std::vector<u64> base(n, 0);
    for (size_t c = 0; c < n_cols; c++) {
        const std::vector<u64> v = col(c);
        const u64 a = out.alphas[c];
        for (size_t i = 0; i < n; i++) base[i] = gl_add(base[i], gl_mul(a, v[i]));
    }
    u64 z = derive_ood_point(tr);  // prover.rs:118-135
    {
        const u64 shift_inv = gl_inv(shift);
        auto on_coset = [&](u64 zz) {
            u64 t = gl_mul(zz, shift_inv);
            for (unsigned i = 0; i < lde_k; i++) t = gl_mul(t, t);
            return t == 1;
        };
        while (on_coset(z)) z = gl_add(z, 1);
    }
    out.z = z;
    std::vector<u64> cur = deep_coset_lde(base.data(), n, log_blow, shift, z);
    base = std::vector<u64>();
    Digest root0 = unlabeled_root_streaming(cur.data(), N);
    tr.absorb("fri_layer_root", root0.data(), 32);
    out.fri_roots.push_back(root0);
    out.betas = derive_betas(tr, lde_k);
    for (unsigned r = 0; r < lde_k; r++) {  // prover.rs:204-238
        const size_t half = cur.size() / 2;
        std::vector<u64> nx(half);
        for (size_t i = 0; i < half; i++) nx[i] = gl_add(cur[i], gl_mul(out.betas[r], cur[i + half]));
        cur.swap(nx);
        Digest root = unlabeled_root_streaming(cur.data(), cur.size());
        tr.absorb("fri_layer_root", root.data(), 32);
        out.fri_roots.push_back(root);
    }
    out.final_value = cur[0];
    return out;
}

}  // namespace oracle
