// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/gl.hpp header note).
//
// CPU restatement of the reference's STARK v1 commitment path, function by function.  Every
// function cites the reference file:line it follows (paths relative to /root/reference/crates).
// Parity pinning: BLAKE3 / transcript / parent combiner are pinned by the reference's shipped
// fixtures (manifest roots + v0 proof bytes, tests/golden/); NTT / LDE / FRI / ProofV1 bytes are
// *unpinned by the reference's own tests* (it ships no v1 golden output) and are pinned instead by
// exact-arithmetic definitions (naive O(n^2) DFT), by the known-answer values of SURVEY.md §8c
// (an independent Python restatement) and by replaying the reference's structural tests.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/sezkp_trace.h"
#include "blake3_ref.hpp"
#include "gl.hpp"

namespace oracle {

using Digest = std::array<u8, 32>;
using Le8 = std::array<u8, 8>;

/* ------------------------------------------------------------------------------------------ */
/* Transcript — sezkp-crypto/src/lib.rs:42 (prefix), :74-124 (Blake3Transcript)               */
/* ------------------------------------------------------------------------------------------ */
struct Transcript {
    Blake3 st;
    static void put_u32(Blake3& h, u32 v) { h.update(&v, 4); }
    explicit Transcript(const std::string& domain) {  // lib.rs:82-87
        st.update("sezkp.transcript.v0", 19);
        put_u32(st, (u32)domain.size());
        st.update(domain.data(), domain.size());
    }
    void absorb(const std::string& label, const void* bytes, size_t n) {  // lib.rs:95-99
        st.update("absorb", 6);
        put_u32(st, (u32)label.size());
        st.update(label.data(), label.size());
        put_u32(st, (u32)n);
        st.update(bytes, n);
    }
    void absorb_u64(const std::string& label, u64 x) { absorb(label, &x, 8); }  // lib.rs:53-55
    std::vector<u8> challenge_bytes(const std::string& label, size_t n) {      // lib.rs:104-120
        Blake3 c = st;  // Hasher::clone
        c.update("challenge", 9);
        put_u32(c, (u32)label.size());
        c.update(label.data(), label.size());
        std::vector<u8> out(n);
        c.finalize_xof(out.data(), n);
        st.update("after_challenge", 15);
        put_u32(st, (u32)label.size());
        st.update(label.data(), label.size());
        return out;
    }
};

/* ------------------------------------------------------------------------------------------ */
/* NTT — sezkp-ffts/src/ntt.rs, coset.rs, lib.rs (naive dft)                                   */
/* ------------------------------------------------------------------------------------------ */
inline size_t bitrev(size_t x, unsigned bits) {  // ntt.rs:19-27
    size_t y = 0;
    for (unsigned i = 0; i < bits; i++) {
        y = (y << 1) | (x & 1);
        x >>= 1;
    }
    return y;
}
inline unsigned log2_exact(size_t n) {
    unsigned k = 0;
    while (((size_t)1 << k) < n) k++;
    return k;
}
inline void bit_reverse_permute(u64* a, size_t n) {  // ntt.rs:29-40
    unsigned bits = log2_exact(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, bits);
        if (j > i) std::swap(a[i], a[j]);
    }
}
// ntt.rs:42-74: per-stage tables, stage s has 2^(s-1) powers of w_{2^s} (or its inverse); rebuilt per call.
inline std::vector<std::vector<u64>> build_twiddles(unsigned n_log2, bool inverse) {
    std::vector<std::vector<u64>> out;
    for (unsigned s = 1; s <= n_log2; s++) {
        size_t half = (size_t)1 << (s - 1);
        u64 wl = gl_root_2exp(s);
        if (inverse) wl = gl_inv(wl);
        std::vector<u64> ws(half);
        u64 w = 1;
        for (size_t i = 0; i < half; i++) {
            ws[i] = w;
            w = gl_mul(w, wl);
        }
        out.push_back(std::move(ws));
    }
    return out;
}
inline void ntt_core(u64* a, size_t n, bool inverse) {  // ntt.rs:79-111 / :117-147
    if (n <= 1) return;
    if (n & (n - 1)) throw std::invalid_argument("NTT size must be power of two");
    bit_reverse_permute(a, n);
    unsigned n_log2 = log2_exact(n);
    auto tw = build_twiddles(n_log2, inverse);
    size_t len = 2, stage = 1;
    while (len <= n) {
        size_t half = len / 2;
        const u64* ws = tw[stage - 1].data();
        for (size_t j = 0; j < n; j += len)
            for (size_t i = 0; i < half; i++) {
                u64 u = a[j + i];
                u64 v = gl_mul(a[j + i + half], ws[i]);
                a[j + i] = gl_add(u, v);
                a[j + i + half] = gl_sub(u, v);
            }
        stage++;
        len <<= 1;
    }
}
inline void forward_ntt_in_place(u64* a, size_t n) { ntt_core(a, n, false); }  // ntt.rs:79
inline void inverse_ntt_in_place(u64* a, size_t n) {                           // ntt.rs:117-155
    if (n <= 1) return;
    ntt_core(a, n, true);
    u64 inv_n = gl_inv(gl_from_u64((u64)n));
    for (size_t i = 0; i < n; i++) a[i] = gl_mul(a[i], inv_n);
}
// coset.rs:85-102
inline std::vector<u64> evaluate_on_coset_pow2(const u64* coeffs, size_t m_in, unsigned k_log2, u64 shift) {
    size_t n = (size_t)1 << k_log2;
    std::vector<u64> scaled(n, 0);
    u64 pw = 1;
    size_t m = m_in < n ? m_in : n;
    for (size_t j = 0; j < m; j++) {
        scaled[j] = gl_mul(coeffs[j], pw);
        pw = gl_mul(pw, shift);
    }
    forward_ntt_in_place(scaled.data(), n);
    return scaled;
}
// lib.rs:191-202 — the independent O(n^2) definition used to pin the fast transform.
inline std::vector<u64> dft_naive(const u64* a, size_t n, u64 omega) {
    std::vector<u64> out(n);
    for (size_t k = 0; k < n; k++) {
        u64 acc = 0;
        for (size_t j = 0; j < n; j++) acc = gl_add(acc, gl_mul(a[j], gl_pow(omega, (u64)j * (u64)k)));
        out[k] = acc;
    }
    return out;
}

/* ------------------------------------------------------------------------------------------ */
/* Leaves / parent combiner / trees — v1/merkle.rs, v1/fri_stream.rs, sezkp-merkle/src/lib.rs  */
/* ------------------------------------------------------------------------------------------ */
inline Digest hash_leaf(const u8 le[8]) {  // merkle.rs:150-159, fri_stream.rs:37-41
    Digest d;
    Blake3::hash(le, 8, d.data());
    return d;
}
inline Digest hash_leaf_labeled(const u8 le[8], const std::string& label) {  // merkle.rs:132-146
    Blake3 h;
    h.update("col_leaf", 8);  // params.rs:58 DS_COL_LEAF
    u32 llen = (u32)label.size();
    h.update(&llen, 4);
    h.update(label.data(), label.size());
    h.update(le, 8);
    Digest d;
    h.finalize(d.data());
    return d;
}
// merkle.rs:57-61 == fri_stream.rs:45-50 == sezkp-merkle/src/lib.rs:123-128 == sezkp-fold/src/fold.rs:31-37
inline Digest node_hash(const Digest& l, const Digest& r) {
    Blake3 h;
    h.update(l.data(), 32);
    h.update(r.data(), 32);
    Digest d;
    h.finalize(d.data());
    return d;
}

struct MerkleTree {  // merkle.rs:32-127
    std::vector<Digest> leaves;
    std::vector<Digest> nodes;
    static std::vector<Digest> next_level(const std::vector<Digest>& lvl) {
        std::vector<Digest> nx;
        nx.reserve((lvl.size() + 1) / 2);
        for (size_t i = 0; i < lvl.size(); i += 2) {
            if (i + 1 < lvl.size()) nx.push_back(node_hash(lvl[i], lvl[i + 1]));
            else nx.push_back(lvl[i]);  // odd promotion
        }
        return nx;
    }
    static MerkleTree from_leaves(const std::vector<Digest>& raw) {  // merkle.rs:46-71
        MerkleTree t;
        t.leaves = raw;
        if (t.leaves.empty()) t.leaves.push_back(Digest{});
        std::vector<Digest> lvl = t.leaves;
        t.nodes = lvl;
        while (lvl.size() > 1) {
            lvl = next_level(lvl);
            t.nodes.insert(t.nodes.end(), lvl.begin(), lvl.end());
        }
        return t;
    }
    Digest root() const { return nodes.back(); }
    std::vector<Digest> open(size_t idx) const {  // merkle.rs:80-108
        std::vector<Digest> sibs;
        std::vector<Digest> lvl = leaves;
        if (!lvl.empty()) idx %= lvl.size();
        while (lvl.size() > 1) {
            size_t sib = ((idx ^ 1) < lvl.size()) ? (idx ^ 1) : idx;
            sibs.push_back(lvl[sib]);
            lvl = next_level(lvl);
            idx >>= 1;
        }
        return sibs;
    }
    static bool verify(const Digest& root, const Digest& leaf, size_t idx, const std::vector<Digest>& sibs) {  // merkle.rs:110-126
        Digest cur = leaf;
        for (const Digest& s : sibs) {
            cur = (idx & 1) == 0 ? node_hash(cur, s) : node_hash(s, cur);
            idx >>= 1;
        }
        return cur == root;
    }
};

// merkle.rs:243-280
inline bool verify_chunked_open(const Digest& outer_root, const std::string& label, const u8 value_le[8],
                                const Digest& chunk_root, size_t idx_in_chunk, const std::vector<Digest>& path_in,
                                size_t chunk_idx, const std::vector<Digest>& path_to) {
    Digest leaf = hash_leaf_labeled(value_le, label);
    if (!MerkleTree::verify(chunk_root, leaf, idx_in_chunk, path_in)) return false;
    return MerkleTree::verify(outer_root, chunk_root, chunk_idx, path_to);
}

struct StreamingLayerBuilder {  // fri_stream.rs:55-122
    size_t expected, seen = 0;
    std::vector<std::pair<bool, Digest>> stack;
    explicit StreamingLayerBuilder(size_t n) : expected(n) {}
    void absorb_leaf(const u8 le[8]) {  // :75-95
        seen++;
        Digest cur = hash_leaf(le);
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
    Digest finalize() const {  // :99-121 (note: H(node, acc) order — equals MerkleTree only for powers of two)
        if (seen != expected) throw std::runtime_error("StreamingLayerBuilder: leaf count mismatch");
        bool have = false;
        Digest cur{};
        for (size_t i = stack.size(); i-- > 0;) {
            if (!stack[i].first) continue;
            if (!have) {
                cur = stack[i].second;
                have = true;
            } else cur = node_hash(stack[i].second, cur);
        }
        return cur;
    }
};

// sezkp-merkle/src/lib.rs:140-157 (batch root with odd promotion; empty -> zeros)
inline Digest merkle_root(std::vector<Digest> leaves) {
    if (leaves.empty()) return Digest{};
    while (leaves.size() > 1) leaves = MerkleTree::next_level(leaves);
    return leaves[0];
}
// sezkp-merkle/src/lib.rs:167-208 (Frontier; finalize folds high->low as H(acc, node))
struct Frontier {
    std::vector<std::pair<bool, Digest>> slots;
    void push_leaf(Digest h) {
        size_t lvl = 0;
        for (;;) {
            if (slots.size() <= lvl) slots.resize(lvl + 1, {false, Digest{}});
            if (!slots[lvl].first) {
                slots[lvl] = {true, h};
                break;
            }
            slots[lvl].first = false;
            h = node_hash(slots[lvl].second, h);
            lvl++;
        }
    }
    Digest finalize_root() const {
        bool have = false;
        Digest acc{};
        for (size_t i = slots.size(); i-- > 0;) {
            if (!slots[i].first) continue;
            if (!have) {
                acc = slots[i].second;
                have = true;
            } else acc = node_hash(acc, slots[i].second);
        }
        return acc;
    }
};

/* ------------------------------------------------------------------------------------------ */
/* Block view over the flat descriptor                                                         */
/* ------------------------------------------------------------------------------------------ */
struct TraceView {
    const sezkp_trace_desc* d;
    size_t tau, n;
    std::vector<size_t> block_start;  // row offset of each block
    explicit TraceView(const sezkp_trace_desc* desc) : d(desc), tau(desc->tau), n(desc->n_rows) {
        size_t row = 0;
        for (u64 k = 0; k < d->n_blocks; k++) {
            block_start.push_back(row);
            row += d->block_len[k];
        }
        if (row != n) throw std::invalid_argument("trace_desc: n_rows != sum(block_len)");
    }
};

// v1/openings.rs:89-116 all_labels (canonical column order)
inline std::vector<std::string> all_labels(size_t tau) {
    std::vector<std::string> out = {"input_mv", "is_first", "is_last"};
    const char* groups[7] = {"mv_", "wflag_", "wsym_", "head_", "winlen_", "in_off_", "out_off_"};
    for (const char* g : groups)
        for (size_t r = 0; r < tau; r++) out.push_back(std::string(g) + std::to_string(r));
    return out;
}

/* ------------------------------------------------------------------------------------------ */
/* TraceColumns — v1/columns.rs:252-365 (identical semantics: v1/openings.rs:193-273 RowIter)   */
/* ------------------------------------------------------------------------------------------ */
constexpr int SYM_BITS = 4;    // columns.rs:32
constexpr int HEAD_BITS = 16;  // columns.rs:34

struct TraceColumns {
    size_t n = 0, tau = 0;
    std::vector<u64> input_mv, is_first, is_last;
    std::vector<std::vector<u64>> mv, write_flag, write_sym, head, win_len, in_off, out_off;
    std::vector<std::vector<std::vector<u64>>> sym_bits, head_bits, slack_bits;  // [tau][bits][n]

    static TraceColumns build(const TraceView& tv, bool with_bits = true) {
        TraceColumns c;
        c.n = tv.n;
        c.tau = tv.tau;
        const sezkp_trace_desc* d = tv.d;
        size_t n = c.n, tau = c.tau;
        c.input_mv.assign(n, 0);
        c.is_first.assign(n, 0);
        c.is_last.assign(n, 0);
        auto mk = [&](std::vector<std::vector<u64>>& v) { v.assign(tau, std::vector<u64>(n, 0)); };
        mk(c.mv); mk(c.write_flag); mk(c.write_sym); mk(c.head); mk(c.win_len); mk(c.in_off); mk(c.out_off);
        if (with_bits) {
            c.sym_bits.assign(tau, std::vector<std::vector<u64>>(SYM_BITS, std::vector<u64>(n, 0)));
            c.head_bits.assign(tau, std::vector<std::vector<u64>>(HEAD_BITS, std::vector<u64>(n, 0)));
            c.slack_bits.assign(tau, std::vector<std::vector<u64>>(HEAD_BITS, std::vector<u64>(n, 0)));
        }
        size_t row = 0;
        for (u64 k = 0; k < d->n_blocks; k++) {
            size_t len = d->block_len[k];
            if (len == 0) continue;
            c.is_first[row] = 1;            // columns.rs:287
            c.is_last[row + len - 1] = 1;   // columns.rs:288
            std::vector<u64> wlen(tau);
            for (size_t r = 0; r < tau; r++) {  // columns.rs:291-297: (right-left).unsigned_abs()+1
                i64 diff = d->win_right[k * tau + r] - d->win_left[k * tau + r];
                u64 ad = diff < 0 ? (u64)0 - (u64)diff : (u64)diff;
                wlen[r] = ad + 1;
            }
            std::vector<i64> cur(tau, 0);  // columns.rs:300
            for (size_t j = 0; j < len; j++) {
                size_t i = row + j;
                c.input_mv[i] = gl_from_i64((i64)d->input_mv[i]);
                for (size_t r = 0; r < tau; r++) {
                    i64 m = d->mv[i * tau + r];
                    c.mv[r][i] = gl_from_i64(m);
                    c.write_flag[r][i] = gl_from_u64(d->write_flag[i * tau + r] ? 1 : 0);
                    c.write_sym[r][i] = gl_from_u64(d->write_flag[i * tau + r] ? d->write_sym[i * tau + r] : 0);
                    cur[r] += m;  // move-then-write: head is post-move (columns.rs:313)
                    c.head[r][i] = gl_from_i64(cur[r]);
                    c.win_len[r][i] = gl_from_u64(wlen[r]);
                    c.in_off[r][i] = gl_from_u64((u64)d->head_in_off[k * tau + r]);
                    c.out_off[r][i] = gl_from_u64((u64)d->head_out_off[k * tau + r]);
                    if (with_bits) {  // columns.rs:324-342
                        u64 sym_u = c.write_sym[r][i];
                        for (int b = 0; b < SYM_BITS; b++) c.sym_bits[r][b][i] = (sym_u >> b) & 1;
                        u64 head_u = c.head[r][i];
                        for (int b = 0; b < HEAD_BITS; b++) c.head_bits[r][b][i] = (head_u >> b) & 1;
                        u64 slack = gl_sub(gl_sub(c.win_len[r][i], 1), c.head[r][i]);
                        for (int b = 0; b < HEAD_BITS; b++) c.slack_bits[r][b][i] = (slack >> b) & 1;
                    }
                }
            }
            row += len;
        }
        return c;
    }
    // committed column by canonical index (order of all_labels)
    const std::vector<u64>& committed(size_t ci) const {
        if (ci == 0) return input_mv;
        if (ci == 1) return is_first;
        if (ci == 2) return is_last;
        size_t g = (ci - 3) / tau, r = (ci - 3) % tau;
        switch (g) {
            case 0: return mv[r];
            case 1: return write_flag[r];
            case 2: return write_sym[r];
            case 3: return head[r];
            case 4: return win_len[r];
            case 5: return in_off[r];
            default: return out_off[r];
        }
    }
};

/* ------------------------------------------------------------------------------------------ */
/* AIR composition — v1/air.rs:49-136; Alphas mapping v1/prover.rs:85-98                       */
/* ------------------------------------------------------------------------------------------ */
struct Alphas {
    u64 bool_flag, mv_domain, head_update, head_bits_bool, head_reconstruct, slack_bits_bool, slack_reconstruct,
        sym_bits_bool, sym_reconstruct, boundary_first, boundary_last;
    static Alphas from8(const u64 a[8]) {  // prover.rs:86-98 (note the reuse of a[0] and a[2])
        return Alphas{a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[0], a[2], a[2]};
    }
};

inline u64 compose_row(const TraceColumns& tc, size_t i, const Alphas& a) {  // air.rs:49-113
    u64 acc = 0;
    for (size_t r = 0; r < tc.tau; r++) {
        u64 mv = tc.mv[r][i], flg = tc.write_flag[r][i], head = tc.head[r][i];
        size_t ip1 = (i + 1) % tc.n;
        u64 head_next = tc.head[r][ip1], mv_next = tc.mv[r][ip1];
        acc = gl_add(acc, gl_mul(gl_mul(a.bool_flag, flg), gl_sub(flg, 1)));
        acc = gl_add(acc, gl_mul(gl_mul(gl_mul(a.mv_domain, mv), gl_sub(mv, 1)), gl_add(mv, 1)));
        u64 one_minus_last = gl_sub(1, tc.is_last[i]);
        acc = gl_add(acc, gl_mul(gl_mul(a.head_update, one_minus_last), gl_sub(gl_sub(head_next, head), mv_next)));
        u64 sum = 0, bb = 0, pw = 1;
        for (int k = 0; k < HEAD_BITS; k++) {
            u64 b = tc.head_bits[r][k][i];
            bb = gl_add(bb, gl_mul(b, gl_sub(b, 1)));
            sum = gl_add(sum, gl_mul(b, pw));
            pw = gl_mul(pw, 2);
        }
        acc = gl_add(acc, gl_mul(gl_mul(a.head_bits_bool, flg), bb));
        acc = gl_add(acc, gl_mul(gl_mul(a.head_reconstruct, flg), gl_sub(head, sum)));
        sum = 0; bb = 0; pw = 1;
        for (int k = 0; k < HEAD_BITS; k++) {
            u64 b = tc.slack_bits[r][k][i];
            bb = gl_add(bb, gl_mul(b, gl_sub(b, 1)));
            sum = gl_add(sum, gl_mul(b, pw));
            pw = gl_mul(pw, 2);
        }
        u64 slack = gl_sub(gl_sub(tc.win_len[r][i], 1), head);
        acc = gl_add(acc, gl_mul(gl_mul(a.slack_bits_bool, flg), bb));
        acc = gl_add(acc, gl_mul(gl_mul(a.slack_reconstruct, flg), gl_sub(slack, sum)));
        sum = 0; bb = 0; pw = 1;
        for (int k = 0; k < SYM_BITS; k++) {
            u64 b = tc.sym_bits[r][k][i];
            bb = gl_add(bb, gl_mul(b, gl_sub(b, 1)));
            sum = gl_add(sum, gl_mul(b, pw));
            pw = gl_mul(pw, 2);
        }
        acc = gl_add(acc, gl_mul(gl_mul(a.sym_bits_bool, flg), bb));
        acc = gl_add(acc, gl_mul(gl_mul(a.sym_reconstruct, flg), gl_sub(tc.write_sym[r][i], sum)));
    }
    return acc;
}
inline u64 compose_boundary(const TraceColumns& tc, size_t i, const Alphas& a) {  // air.rs:116-136
    u64 acc = 0;
    for (size_t r = 0; r < tc.tau; r++) {
        u64 head = tc.head[r][i], mv = tc.mv[r][i];
        acc = gl_add(acc, gl_mul(gl_mul(a.boundary_first, tc.is_first[i]), gl_sub(gl_sub(head, mv), tc.in_off[r][i])));
        acc = gl_add(acc, gl_mul(gl_mul(a.boundary_last, tc.is_last[i]), gl_sub(head, tc.out_off[r][i])));
    }
    return acc;
}
// masking.rs:86-103 (Horner, ascending coefficients)
inline u64 eval_masks_sum_at(const std::vector<std::vector<u64>>& all, u64 x) {
    u64 s = 0;
    for (const auto& co : all) {
        u64 acc = 0;
        for (size_t j = co.size(); j-- > 0;) acc = gl_add(gl_mul(acc, x), co[j]);
        s = gl_add(s, acc);
    }
    return s;
}
// the `base_eval` closure of prover.rs:142-158: C(i) + B(i) + R(w_n^i)
inline std::vector<u64> compose_base(const TraceColumns& tc, const Alphas& al,
                                     const std::vector<std::vector<u64>>& masks) {
    std::vector<u64> out(tc.n);
    u64 w_base = gl_root_2exp(log2_exact(tc.n));
    u64 x = 1;
    for (size_t i = 0; i < tc.n; i++) {
        u64 comp = gl_add(compose_row(tc, i, al), compose_boundary(tc, i, al));
        out[i] = gl_add(comp, eval_masks_sum_at(masks, x));
        x = gl_mul(x, w_base);
    }
    return out;
}

/* ------------------------------------------------------------------------------------------ */
/* DEEP coset LDE — v1/lde.rs:42-97                                                            */
/* ------------------------------------------------------------------------------------------ */
inline std::vector<u64> deep_coset_lde(const u64* base_evals, size_t n_base, unsigned blow_log2, u64 shift, u64 z) {
    if (n_base & (n_base - 1)) throw std::invalid_argument("n_base must be a power of two");  // lde.rs:51
    unsigned lde_k = log2_exact(n_base) + blow_log2;
    size_t lde_n = (size_t)1 << lde_k;
    std::vector<u64> coeffs(base_evals, base_evals + n_base);
    inverse_ntt_in_place(coeffs.data(), n_base);                                // lde.rs:65
    std::vector<u64> y = evaluate_on_coset_pow2(coeffs.data(), n_base, lde_k, shift);  // lde.rs:69
    u64 w = gl_root_2exp(lde_k), w_pow = 1;
    for (size_t i = 0; i < lde_n; i++) {  // lde.rs:80-93: per-element Fermat inverse
        u64 x = gl_mul(shift, w_pow);
        u64 denom = gl_sub(x, z);
        y[i] = gl_mul(y[i], gl_inv(denom));
        w_pow = gl_mul(w_pow, w);
    }
    return y;
}

/* ------------------------------------------------------------------------------------------ */
/* Proof objects — v1/proof.rs:17-98; bincode 1.3.3 default config (sezkp-stark/src/lib.rs:131) */
/* ------------------------------------------------------------------------------------------ */
struct Opening {
    Le8 value_le{};
    u64 index = 0, chunk_index = 0, index_in_chunk = 0;
    Digest chunk_root{};
    std::vector<Digest> path_in_chunk, path_to_chunk;
};
struct PerTapeOpen { Opening mv, next_mv, write_flag, write_sym, head, next_head, win_len, in_off, out_off; };
struct RowOpenings {
    u64 row = 0;
    std::vector<PerTapeOpen> per_tape;
    Opening is_first, is_last, input_mv;
};
struct FriPair { Le8 vi{}; std::vector<Digest> pi; Le8 vj{}; std::vector<Digest> pj; };
struct FriQuery { std::vector<u64> positions; std::vector<FriPair> pairs; };
struct ColumnRoot { std::string label; Digest root{}; };
struct ProofV1 {
    u64 domain_n = 0, tau = 0;
    std::vector<ColumnRoot> col_roots;
    std::vector<RowOpenings> queries;
    std::vector<Digest> fri_roots;
    std::vector<FriQuery> fri_queries;
    Le8 fri_final_value_le{};
    Digest manifest_root{};
};

struct BinW {
    std::vector<u8> b;
    void u64le(u64 v) { for (int i = 0; i < 8; i++) b.push_back((u8)(v >> (8 * i))); }
    void raw(const u8* p, size_t n) { b.insert(b.end(), p, p + n); }
    void digests(const std::vector<Digest>& v) { u64le(v.size()); for (auto& d : v) raw(d.data(), 32); }
    void opening(const Opening& o) {
        raw(o.value_le.data(), 8); u64le(o.index); u64le(o.chunk_index); u64le(o.index_in_chunk);
        raw(o.chunk_root.data(), 32); digests(o.path_in_chunk); digests(o.path_to_chunk);
    }
};
inline std::vector<u8> bincode_proof(const ProofV1& p) {  // field order = struct declaration order proof.rs:80-98
    BinW w;
    w.u64le(p.domain_n); w.u64le(p.tau);
    w.u64le(p.col_roots.size());
    for (auto& c : p.col_roots) { w.u64le(c.label.size()); w.raw((const u8*)c.label.data(), c.label.size()); w.raw(c.root.data(), 32); }
    w.u64le(p.queries.size());
    for (auto& q : p.queries) {
        w.u64le(q.row);
        w.u64le(q.per_tape.size());
        for (auto& t : q.per_tape) {
            w.opening(t.mv); w.opening(t.next_mv); w.opening(t.write_flag); w.opening(t.write_sym); w.opening(t.head);
            w.opening(t.next_head); w.opening(t.win_len); w.opening(t.in_off); w.opening(t.out_off);
        }
        w.opening(q.is_first); w.opening(q.is_last); w.opening(q.input_mv);
    }
    w.digests(p.fri_roots);  // FriRoots{roots}
    w.u64le(p.fri_queries.size());
    for (auto& fq : p.fri_queries) {
        w.u64le(fq.positions.size());
        for (u64 x : fq.positions) w.u64le(x);
        w.u64le(fq.pairs.size());
        for (auto& pr : fq.pairs) { w.raw(pr.vi.data(), 8); w.digests(pr.pi); w.raw(pr.vj.data(), 8); w.digests(pr.pj); }
    }
    w.raw(p.fri_final_value_le.data(), 8);
    w.raw(p.manifest_root.data(), 32);
    return w.b;
}
struct BinR {
    const u8* p; size_t n, off = 0;
    BinR(const u8* p_, size_t n_) : p(p_), n(n_) {}
    void need(size_t k) { if (off + k > n) throw std::runtime_error("bincode: truncated proof"); }
    u64 u64le() { need(8); u64 v; std::memcpy(&v, p + off, 8); off += 8; return v; }
    void raw(u8* o, size_t k) { need(k); std::memcpy(o, p + off, k); off += k; }
    std::vector<Digest> digests() {
        u64 k = u64le();
        if (k > n) throw std::runtime_error("bincode: bad length");
        std::vector<Digest> v(k);
        for (auto& d : v) raw(d.data(), 32);
        return v;
    }
    Opening opening() {
        Opening o;
        raw(o.value_le.data(), 8); o.index = u64le(); o.chunk_index = u64le(); o.index_in_chunk = u64le();
        raw(o.chunk_root.data(), 32); o.path_in_chunk = digests(); o.path_to_chunk = digests();
        return o;
    }
};
inline ProofV1 bincode_parse(const u8* buf, size_t len) {
    BinR r(buf, len);
    ProofV1 p;
    p.domain_n = r.u64le(); p.tau = r.u64le();
    u64 nc = r.u64le();
    if (nc > len) throw std::runtime_error("bincode: bad length");
    for (u64 i = 0; i < nc; i++) {
        ColumnRoot c; u64 l = r.u64le(); r.need(l); c.label.assign((const char*)r.p + r.off, l); r.off += l; r.raw(c.root.data(), 32);
        p.col_roots.push_back(c);
    }
    u64 nq = r.u64le();
    if (nq > len) throw std::runtime_error("bincode: bad length");
    for (u64 i = 0; i < nq; i++) {
        RowOpenings q; q.row = r.u64le();
        u64 nt = r.u64le();
        if (nt > len) throw std::runtime_error("bincode: bad length");
        for (u64 t = 0; t < nt; t++) {
            PerTapeOpen o;
            o.mv = r.opening(); o.next_mv = r.opening(); o.write_flag = r.opening(); o.write_sym = r.opening(); o.head = r.opening();
            o.next_head = r.opening(); o.win_len = r.opening(); o.in_off = r.opening(); o.out_off = r.opening();
            q.per_tape.push_back(o);
        }
        q.is_first = r.opening(); q.is_last = r.opening(); q.input_mv = r.opening();
        p.queries.push_back(q);
    }
    p.fri_roots = r.digests();
    u64 nf = r.u64le();
    if (nf > len) throw std::runtime_error("bincode: bad length");
    for (u64 i = 0; i < nf; i++) {
        FriQuery fq; u64 np = r.u64le();
        if (np > len) throw std::runtime_error("bincode: bad length");
        for (u64 k = 0; k < np; k++) fq.positions.push_back(r.u64le());
        u64 npr = r.u64le();
        if (npr > len) throw std::runtime_error("bincode: bad length");
        for (u64 k = 0; k < npr; k++) { FriPair pr; r.raw(pr.vi.data(), 8); pr.pi = r.digests(); r.raw(pr.vj.data(), 8); pr.pj = r.digests(); fq.pairs.push_back(pr); }
        p.fri_queries.push_back(fq);
    }
    r.raw(p.fri_final_value_le.data(), 8);
    r.raw(p.manifest_root.data(), 32);
    if (r.off != len) throw std::runtime_error("bincode: trailing bytes");
    return p;
}

/* ------------------------------------------------------------------------------------------ */
/* Challenge derivers — v1/params.rs:76-126, v1/masking.rs:56-79                                */
/* ------------------------------------------------------------------------------------------ */
inline u64 le64(const u8* b) { u64 v; std::memcpy(&v, b, 8); return v; }
inline void derive_alphas(Transcript& tr, u64 out[8]) {  // params.rs:76-86
    auto by = tr.challenge_bytes("alphas", 64);
    for (int i = 0; i < 8; i++) out[i] = gl_from_u64(le64(&by[8 * i]));
}
inline std::vector<size_t> derive_queries(Transcript& tr, size_t n, size_t k) {  // params.rs:89-100
    auto by = tr.challenge_bytes("row_queries", 8 * k);
    std::vector<size_t> out;
    for (size_t i = 0; i < k; i++) out.push_back((size_t)(le64(&by[8 * i]) % (u64)(n > 1 ? n : 1)));
    return out;
}
inline std::vector<u64> derive_betas(Transcript& tr, size_t n_layers) {  // params.rs:103-113
    auto by = tr.challenge_bytes("fri_betas", 8 * n_layers);
    std::vector<u64> out;
    for (size_t i = 0; i < n_layers; i++) out.push_back(gl_from_u64(le64(&by[8 * i])));
    return out;
}
inline u64 derive_ood_point(Transcript& tr) {  // params.rs:116-126
    auto by = tr.challenge_bytes("ood_point", 8);
    return gl_from_u64(le64(by.data()));
}
inline std::vector<std::vector<u64>> derive_mask_coeffs(Transcript& tr, size_t deg, size_t k) {  // masking.rs:56-79
    tr.absorb("masks", "masks", 5);
    tr.absorb_u64("n_masks", k);
    tr.absorb_u64("deg", deg);
    std::vector<std::vector<u64>> out(k, std::vector<u64>(deg, 0));
    for (size_t i = 0; i < k; i++)
        for (size_t j = 0; j < deg; j++) {
            auto by = tr.challenge_bytes("mask_coeff", 8);
            out[i][j] = gl_from_u64(le64(by.data()));
        }
    return out;
}

constexpr size_t BLOWUP = 8;          // params.rs:28
constexpr size_t NUM_QUERIES = 30;    // params.rs:31
constexpr size_t COL_CHUNK_LOG2 = 10; // params.rs:37

/* ------------------------------------------------------------------------------------------ */
/* Column commitments + on-demand openings — v1/openings.rs:278-498                              */
/* ------------------------------------------------------------------------------------------ */
struct ColumnOracle {
    const TraceColumns& tc;
    size_t chunk_log2, chunk_size;
    std::vector<std::string> labels;
    mutable std::vector<std::pair<bool, MerkleTree>> outer_cache;  // per label, like OnDemandOpenings::outer_cache (openings.rs:285)
    ColumnOracle(const TraceColumns& t, size_t cl2) : tc(t), chunk_log2(cl2), chunk_size((size_t)1 << cl2), labels(all_labels(t.tau)) {}

    std::vector<Digest> chunk_roots(size_t ci) const {  // openings.rs:436-460 / :306-398
        const auto& col = tc.committed(ci);
        std::vector<Digest> roots, cur;
        for (size_t i = 0; i < tc.n; i++) {
            u8 le[8]; gl_to_le(col[i], le);
            cur.push_back(hash_leaf_labeled(le, labels[ci]));
            if (cur.size() == chunk_size) { roots.push_back(MerkleTree::from_leaves(cur).root()); cur.clear(); }
        }
        if (!cur.empty()) roots.push_back(MerkleTree::from_leaves(cur).root());
        return roots;
    }
    std::vector<ColumnRoot> build_roots() const {  // openings.rs:306-398
        std::vector<ColumnRoot> out;
        for (size_t ci = 0; ci < labels.size(); ci++)
            out.push_back({labels[ci], MerkleTree::from_leaves(chunk_roots(ci)).root()});
        return out;
    }
    size_t index_of(const std::string& label) const {
        for (size_t i = 0; i < labels.size(); i++) if (labels[i] == label) return i;
        throw std::invalid_argument("unknown column label " + label);
    }
    Opening open(const std::string& label, size_t row) const {  // openings.rs:403-432, :464-497
        size_t ci = index_of(label);
        const auto& col = tc.committed(ci);
        Opening o;
        size_t chunk_idx = row / chunk_size, idx_in = row - chunk_idx * chunk_size;
        size_t start = chunk_idx * chunk_size, end = std::min(start + chunk_size, tc.n);
        std::vector<Digest> leaves;
        for (size_t i = start; i < end; i++) {
            u8 le[8]; gl_to_le(col[i], le);
            if (i == start + idx_in) std::memcpy(o.value_le.data(), le, 8);
            leaves.push_back(hash_leaf_labeled(le, label));
        }
        MerkleTree ct = MerkleTree::from_leaves(leaves);
        o.index = row; o.chunk_index = chunk_idx; o.index_in_chunk = idx_in;
        o.chunk_root = ct.root();
        o.path_in_chunk = ct.open(idx_in);
        if (outer_cache.empty()) outer_cache.resize(labels.size());
        if (!outer_cache[ci].first) outer_cache[ci] = {true, MerkleTree::from_leaves(chunk_roots(ci))};  // openings.rs:415-419
        o.path_to_chunk = outer_cache[ci].second.open(chunk_idx);
        return o;
    }
};

/* ------------------------------------------------------------------------------------------ */
/* FRI fold + commit given betas — v1/prover.rs:204-243 (same roots as v1/fri.rs:71-91)          */
/* ------------------------------------------------------------------------------------------ */
inline std::vector<Digest> leaves_of(const u64* v, size_t n) {
    std::vector<Digest> l(n);
    for (size_t i = 0; i < n; i++) { u8 le[8]; gl_to_le(v[i], le); l[i] = hash_leaf(le); }
    return l;
}
// returns all layers (layer 0 = input); roots[l] = Merkle root of layer l
inline std::vector<std::vector<u64>> fri_fold_layers(const std::vector<u64>& layer0, const std::vector<u64>& betas) {
    std::vector<std::vector<u64>> layers{layer0};
    for (size_t r = 0; r < betas.size(); r++) {
        const auto& cur = layers.back();
        size_t half = cur.size() / 2;
        std::vector<u64> nx(half);
        for (size_t i = 0; i < half; i++) nx[i] = gl_add(cur[i], gl_mul(betas[r], cur[i + half]));
        layers.push_back(std::move(nx));
    }
    return layers;
}

/* ------------------------------------------------------------------------------------------ */
/* prove_v1 — v1/prover.rs:61-462 ("compute-once" form: identical output bytes; the reference    */
/* recomputes the DEEP-LDE stream per Merkle level for layer-0 paths, fri_stream.rs:273-309,     */
/* which `faithful_cost` re-enacts for timing only)                                              */
/* ------------------------------------------------------------------------------------------ */
struct ProveTrace {  // intermediate values exposed for per-stage parity tests
    u64 alphas[8];
    std::vector<std::vector<u64>> mask_coeffs;
    u64 z = 0;
    std::vector<u64> betas, base_vals, lde_vals;
    std::vector<size_t> rows, fri_rows;
};

inline ProofV1 prove_v1(const sezkp_trace_desc* desc, const u8 manifest_root[32], bool faithful_cost = false,
                        ProveTrace* trace_out = nullptr) {
    TraceView tv(desc);
    if (tv.n == 0 || (tv.n & (tv.n - 1))) throw std::invalid_argument("n_rows must be a power of two");  // lde.rs:51
    TraceColumns tc = TraceColumns::build(tv);  // prover.rs:64

    Transcript tr("sezkp-stark/v1");  // prover.rs:67-70
    tr.absorb("manifest_root", manifest_root, 32);
    tr.absorb_u64("n", tc.n);
    tr.absorb_u64("tau", tc.tau);

    ColumnOracle odo(tc, COL_CHUNK_LOG2);  // prover.rs:75-81
    std::vector<ColumnRoot> col_roots = odo.build_roots();
    tr.absorb_u64("n_cols", col_roots.size());
    for (auto& r : col_roots) tr.absorb("col_root", r.root.data(), 32);

    u64 a8[8];
    derive_alphas(tr, a8);  // prover.rs:85-98
    Alphas alphas = Alphas::from8(a8);
    auto mask_coeffs = derive_mask_coeffs(tr, 4, 1);  // prover.rs:103 (DEFAULT_MASK_DEG=4, DEFAULT_N_MASKS=1)

    unsigned base_log2 = log2_exact(tc.n), blow_log2 = 3, lde_k = base_log2 + blow_log2;  // prover.rs:108-113
    size_t lde_n = (size_t)1 << lde_k;
    u64 shift = 3;
    u64 z = derive_ood_point(tr);  // prover.rs:120-135
    {
        u64 shift_inv = gl_inv(shift);
        auto on_coset = [&](u64 zz) {
            u64 t = gl_mul(zz, shift_inv);
            for (unsigned i = 0; i < lde_k; i++) t = gl_mul(t, t);
            return t == 1;
        };
        while (on_coset(z)) z = gl_add(z, 1);
    }

    std::vector<u64> base_vals = compose_base(tc, alphas, mask_coeffs);             // prover.rs:142-158
    std::vector<u64> lde_vals = deep_coset_lde(base_vals.data(), tc.n, blow_log2, shift, z);  // prover.rs:162-178

    std::vector<Digest> fri_roots;  // prover.rs:184-198
    {
        StreamingLayerBuilder l0(lde_n);
        for (size_t i = 0; i < lde_n; i++) { u8 le[8]; gl_to_le(lde_vals[i], le); l0.absorb_leaf(le); }
        Digest root0 = l0.finalize();
        tr.absorb("fri_layer_root", root0.data(), 32);
        fri_roots.push_back(root0);
    }
    size_t n_folds = lde_k;
    std::vector<u64> betas = derive_betas(tr, n_folds);
    std::vector<std::vector<u64>> layers = fri_fold_layers(lde_vals, betas);  // prover.rs:204-238
    for (size_t l = 1; l < layers.size(); l++) {
        Digest root = MerkleTree::from_leaves(leaves_of(layers[l].data(), layers[l].size())).root();
        tr.absorb("fri_layer_root", root.data(), 32);
        fri_roots.push_back(root);
    }
    Le8 final_le;
    gl_to_le(n_folds == 0 ? lde_vals[0] : layers.back()[0], final_le.data());  // prover.rs:242-243

    std::vector<size_t> rows = derive_queries(tr, tc.n, NUM_QUERIES);  // prover.rs:248
    std::vector<RowOpenings> queries;
    for (size_t row : rows) {  // prover.rs:252-292
        RowOpenings q;
        q.row = row;
        q.input_mv = odo.open("input_mv", row);
        q.is_first = odo.open("is_first", row);
        q.is_last = odo.open("is_last", row);
        size_t ip1 = (tc.n == 0) ? 0 : (row + 1 < tc.n ? row + 1 : 0);  // next_wrap prover.rs:50-58
        for (size_t r = 0; r < tc.tau; r++) {
            std::string s = std::to_string(r);
            PerTapeOpen t;
            t.mv = odo.open("mv_" + s, row);
            t.next_mv = odo.open("mv_" + s, ip1);
            t.write_flag = odo.open("wflag_" + s, row);
            t.write_sym = odo.open("wsym_" + s, row);
            t.head = odo.open("head_" + s, row);
            t.next_head = odo.open("head_" + s, ip1);
            t.win_len = odo.open("winlen_" + s, row);
            t.in_off = odo.open("in_off_" + s, row);
            t.out_off = odo.open("out_off_" + s, row);
            q.per_tape.push_back(std::move(t));
        }
        queries.push_back(std::move(q));
    }

    std::vector<size_t> fri_rows = derive_queries(tr, lde_n, NUM_QUERIES);  // prover.rs:297 (same label)
    size_t n_layers = fri_roots.size();
    std::vector<FriQuery> fri_queries(fri_rows.size());
    for (auto& fq : fri_queries) fq.positions.assign(n_layers, 0);

    // Trees of every layer that gets opened (layers 0..n_layers-2).
    std::vector<MerkleTree> trees;
    for (size_t l = 0; l + 1 < n_layers; l++) trees.push_back(MerkleTree::from_leaves(leaves_of(layers[l].data(), layers[l].size())));

    auto open_layer0 = [&](size_t idx) -> std::vector<Digest> {
        if (!faithful_cost) return trees[0].open(idx);
        // fri_stream.rs:260-312: one full DEEP-LDE pass + full re-hash per tree level.
        std::vector<Digest> path;
        size_t cur_len = lde_n, id = idx;
        unsigned level = 0;
        while (cur_len > 1) {
            std::vector<u64> bv = compose_base(tc, alphas, mask_coeffs);
            std::vector<u64> y = deep_coset_lde(bv.data(), tc.n, blow_log2, shift, z);
            std::vector<Digest> walk = leaves_of(y.data(), y.size());
            std::vector<Digest> want;
            for (unsigned l = 0;; l++) {  // bubbles every level, like the reference's per-level stack
                if (l == level) want = walk;
                if (walk.size() <= 1) break;
                walk = MerkleTree::next_level(walk);
            }
            size_t sib = id ^ 1;
            path.push_back(sib < cur_len ? want[sib] : want[id]);
            id >>= 1;
            cur_len = (cur_len + 1) / 2;
            level++;
        }
        return path;
    };

    {  // prover.rs:312-398
        size_t half0 = lde_n / 2;
        for (size_t qi = 0; qi < fri_rows.size(); qi++) {
            size_t idx0 = fri_rows[qi], j0 = idx0 ^ half0;
            if (lde_n < 2) throw std::invalid_argument("lde_n < 2 unsupported");
            FriPair pr;
            gl_to_le(lde_vals[idx0], pr.vi.data());
            pr.pi = open_layer0(idx0);
            gl_to_le(lde_vals[j0], pr.vj.data());
            pr.pj = open_layer0(j0);
            fri_queries[qi].positions[0] = idx0;
            if (n_layers > 1) fri_queries[qi].positions[1] = idx0 % (lde_n / 2);
            fri_queries[qi].pairs.push_back(std::move(pr));
        }
    }
    if (n_layers > 1) {  // prover.rs:401-450
        for (size_t r = 1; r + 2 <= n_layers; r++) {
            size_t cur_len = layers[r].size(), half = cur_len / 2;
            for (size_t qi = 0; qi < fri_rows.size(); qi++) {
                size_t idx_r = fri_queries[qi].positions[r], j_r = idx_r ^ half;
                FriPair pr;
                gl_to_le(layers[r][idx_r], pr.vi.data());
                pr.pi = trees[r].open(idx_r);
                gl_to_le(layers[r][j_r], pr.vj.data());
                pr.pj = trees[r].open(j_r);
                fri_queries[qi].pairs.push_back(std::move(pr));
                fri_queries[qi].positions[r + 1] = idx_r % half;  // both branches of prover.rs:430-434 write r+1
            }
        }
    }

    if (trace_out) {
        std::memcpy(trace_out->alphas, a8, sizeof a8);
        trace_out->mask_coeffs = mask_coeffs;
        trace_out->z = z;
        trace_out->betas = betas;
        trace_out->base_vals = base_vals;
        trace_out->lde_vals = lde_vals;
        trace_out->rows = rows;
        trace_out->fri_rows = fri_rows;
    }

    ProofV1 p;  // prover.rs:452-461
    p.domain_n = lde_n;
    p.tau = tc.tau;
    p.col_roots = std::move(col_roots);
    p.queries = std::move(queries);
    p.fri_roots = std::move(fri_roots);
    p.fri_queries = std::move(fri_queries);
    p.fri_final_value_le = final_le;
    std::memcpy(p.manifest_root.data(), manifest_root, 32);
    return p;
}

/* ------------------------------------------------------------------------------------------ */
/* verify_v1 — v1/verify.rs:60-196, fri_verify v1/fri.rs:130-222                                 */
/* returns "" on accept, otherwise the rejection reason                                          */
/* ------------------------------------------------------------------------------------------ */
inline std::string fri_verify(Transcript& tr, const std::vector<Digest>& roots, const std::vector<FriQuery>& queries, const Le8& final_le) {
    if (roots.empty()) return "no FRI roots";
    size_t n_layers = roots.size();
    tr.absorb("fri_layer_root", roots[0].data(), 32);
    std::vector<u64> betas = derive_betas(tr, n_layers - 1);
    if (roots[n_layers - 1] != hash_leaf(final_le.data())) return "final FRI value mismatch with last root";
    for (const FriQuery& q : queries) {
        if (q.positions.size() != n_layers) return "positions length mismatch";
        if (q.pairs.size() != n_layers - 1) return "pairs length mismatch";
        size_t idx = q.positions[0], layer_len = (size_t)1 << (n_layers - 1);
        for (size_t l = 0; l + 1 < n_layers; l++) {
            size_t half = layer_len / 2, j = idx ^ half;
            const FriPair& pr = q.pairs[l];
            bool ok_i = MerkleTree::verify(roots[l], hash_leaf(pr.vi.data()), idx, pr.pi);
            bool ok_j = MerkleTree::verify(roots[l], hash_leaf(pr.vj.data()), j, pr.pj);
            if (!(ok_i && ok_j)) return "FRI Merkle path failed at layer " + std::to_string(l);
            u64 vi = gl_from_u64(le64(pr.vi.data())), vj = gl_from_u64(le64(pr.vj.data()));
            u64 lower = idx < half ? vi : vj, upper = idx < half ? vj : vi;
            u64 v_fold = gl_add(lower, gl_mul(betas[l], upper));
            size_t next = idx % half;
            if (q.positions[l + 1] != next) return "FRI index propagation failed at layer " + std::to_string(l);
            if (l + 1 < n_layers - 1) {
                if (gl_from_u64(le64(q.pairs[l + 1].vi.data())) != v_fold) return "FRI fold mismatch at layer " + std::to_string(l);
            } else {
                Le8 f; gl_to_le(v_fold, f.data());
                if (f != final_le) return "final FRI value mismatch";
            }
            idx = next;
            layer_len = half;
        }
    }
    return "";
}

inline std::string verify_v1(const ProofV1& proof, const sezkp_trace_desc* desc) {
    if (proof.domain_n % BLOWUP != 0) return "FRI domain_n not multiple of blowup";
    size_t n = proof.domain_n / BLOWUP;
    if (n == 0 || (n & (n - 1))) return "trace length n must be a power of two";
    size_t tau = proof.tau;
    if (desc && desc->n_blocks > 0 && desc->tau != tau) return "tau mismatch vs. block windows";

    Transcript tr("sezkp-stark/v1");
    tr.absorb("manifest_root", proof.manifest_root.data(), 32);
    tr.absorb_u64("n", n);
    tr.absorb_u64("tau", tau);
    tr.absorb_u64("n_cols", proof.col_roots.size());
    for (auto& c : proof.col_roots) tr.absorb("col_root", c.root.data(), 32);
    u64 a8[8];
    derive_alphas(tr, a8);
    Alphas al = Alphas::from8(a8);
    (void)derive_mask_coeffs(tr, 4, 1);
    (void)derive_ood_point(tr);

    size_t n_layers = proof.fri_roots.size();
    Transcript tr_rows = tr;  // verify.rs:120-128
    if (n_layers > 0) {
        tr_rows.absorb("fri_layer_root", proof.fri_roots[0].data(), 32);
        (void)derive_betas(tr_rows, n_layers - 1);
        for (size_t r = 1; r < n_layers; r++) tr_rows.absorb("fri_layer_root", proof.fri_roots[r].data(), 32);
    }
    std::vector<size_t> expected = derive_queries(tr_rows, n, NUM_QUERIES);
    if (expected.size() != proof.queries.size()) return "AIR query count mismatch";
    for (size_t i = 0; i < expected.size(); i++)
        if (proof.queries[i].row != expected[i]) return "AIR query row mismatch at position " + std::to_string(i);

    auto root_of = [&](const std::string& label, Digest& out) {
        bool found = false;  // HashMap collect: last duplicate wins
        for (auto& c : proof.col_roots) if (c.label == label) { out = c.root; found = true; }
        return found;
    };
    auto check = [&](const std::string& label, const Opening& o) -> std::string {
        Digest root;
        if (!root_of(label, root)) return "missing col root for " + label;
        if (!verify_chunked_open(root, label, o.value_le.data(), o.chunk_root, o.index_in_chunk, o.path_in_chunk, o.chunk_index, o.path_to_chunk))
            return "chunked merkle path failed for column " + label + " @ " + std::to_string(o.index);
        return "";
    };
    for (const RowOpenings& q : proof.queries) {
        std::string e;
        if (!(e = check("input_mv", q.input_mv)).empty()) return e;
        if (!(e = check("is_first", q.is_first)).empty()) return e;
        if (!(e = check("is_last", q.is_last)).empty()) return e;
        for (size_t r = 0; r < q.per_tape.size(); r++) {
            const PerTapeOpen& t = q.per_tape[r];
            std::string s = std::to_string(r);
            if (!(e = check("mv_" + s, t.mv)).empty()) return e;
            if (!(e = check("mv_" + s, t.next_mv)).empty()) return e;
            if (!(e = check("wflag_" + s, t.write_flag)).empty()) return e;
            if (!(e = check("wsym_" + s, t.write_sym)).empty()) return e;
            if (!(e = check("head_" + s, t.head)).empty()) return e;
            if (!(e = check("head_" + s, t.next_head)).empty()) return e;
            if (!(e = check("winlen_" + s, t.win_len)).empty()) return e;
            if (!(e = check("in_off_" + s, t.in_off)).empty()) return e;
            if (!(e = check("out_off_" + s, t.out_off)).empty()) return e;
        }
        // openings-only AIR: air.rs:209-238
        auto F = [](const Opening& o) { return gl_from_u64(le64(o.value_le.data())); };
        u64 is_first = F(q.is_first), is_last = F(q.is_last), acc = 0;
        for (const PerTapeOpen& t : q.per_tape) {
            u64 mv = F(t.mv), flg = F(t.write_flag), head = F(t.head), head_next = F(t.next_head);
            acc = gl_add(acc, gl_mul(gl_mul(al.bool_flag, flg), gl_sub(flg, 1)));
            acc = gl_add(acc, gl_mul(gl_mul(gl_mul(al.mv_domain, mv), gl_sub(mv, 1)), gl_add(mv, 1)));
            acc = gl_add(acc, gl_mul(gl_mul(al.head_update, gl_sub(1, is_last)), gl_sub(gl_sub(head_next, head), F(t.next_mv))));
        }
        u64 bacc = 0;
        for (const PerTapeOpen& t : q.per_tape) {
            bacc = gl_add(bacc, gl_mul(gl_mul(al.boundary_first, is_first), gl_sub(gl_sub(F(t.head), F(t.mv)), F(t.in_off))));
            bacc = gl_add(bacc, gl_mul(gl_mul(al.boundary_last, is_last), gl_sub(F(t.head), F(t.out_off))));
        }
        if (gl_add(acc, bacc) != 0) return "AIR composition non-zero at row " + std::to_string(q.row);
    }
    return fri_verify(tr, proof.fri_roots, proof.fri_queries, proof.fri_final_value_le);
}

}  // namespace oracle
