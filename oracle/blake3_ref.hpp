// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/gl.hpp header note).
//
// Portable BLAKE3 (unkeyed hash mode): one-shot, incremental update, clone (plain copy) and XOF.
// The reference uses the third-party crate `blake3 = 1.8.2` (Cargo.lock:125-135), which is not
// vendored under /root/reference; this file restates the published BLAKE3 algorithm
//  and is pinned in tests against the official empty-input vector, the
// Python `blake3` module (bindings to the same Rust crate) and the reference's shipped fixtures.
// Reference call sites on the path: v1/merkle.rs:58-61,136-143,154-156; v1/fri_stream.rs:38-49;
// sezkp-merkle/src/lib.rs:86-127; sezkp-crypto/src/lib.rs:82-120.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace oracle {

struct Blake3 {
    static constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

    static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

    static inline void g(uint32_t* s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
        s[a] = s[a] + s[b] + mx;
        s[d] = rotr(s[d] ^ s[a], 16);
        s[c] = s[c] + s[d];
        s[b] = rotr(s[b] ^ s[c], 12);
        s[a] = s[a] + s[b] + my;
        s[d] = rotr(s[d] ^ s[a], 8);
        s[c] = s[c] + s[d];
        s[b] = rotr(s[b] ^ s[c], 7);
    }

    // Full 16-word compression output (lower half = chaining value, all 16 = XOF block).
    static void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                         uint32_t block_len, uint32_t flags, uint32_t out[16]) {
        static const int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
        uint32_t s[16], m[16], t[16];
        for (int i = 0; i < 8; i++) s[i] = cv[i];
        for (int i = 0; i < 4; i++) s[8 + i] = IV[i];
        s[12] = (uint32_t)counter;
        s[13] = (uint32_t)(counter >> 32);
        s[14] = block_len;
        s[15] = flags;
        for (int i = 0; i < 16; i++) m[i] = block[i];
        for (int r = 0; r < 7; r++) {
            g(s, 0, 4, 8, 12, m[0], m[1]);
            g(s, 1, 5, 9, 13, m[2], m[3]);
            g(s, 2, 6, 10, 14, m[4], m[5]);
            g(s, 3, 7, 11, 15, m[6], m[7]);
            g(s, 0, 5, 10, 15, m[8], m[9]);
            g(s, 1, 6, 11, 12, m[10], m[11]);
            g(s, 2, 7, 8, 13, m[12], m[13]);
            g(s, 3, 4, 9, 14, m[14], m[15]);
            for (int i = 0; i < 16; i++) t[i] = m[PERM[i]];
            for (int i = 0; i < 16; i++) m[i] = t[i];
        }
        for (int i = 0; i < 8; i++) {
            out[i] = s[i] ^ s[i + 8];
            out[i + 8] = s[i + 8] ^ cv[i];
        }
    }

    static inline void words_from_le(const uint8_t* b, uint32_t* w, int n) {
        for (int i = 0; i < n; i++)
            w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
                   ((uint32_t)b[4 * i + 3] << 24);
    }
    static inline void le_from_words(const uint32_t* w, uint8_t* b, int n) {
        for (int i = 0; i < n; i++) {
            b[4 * i] = (uint8_t)w[i];
            b[4 * i + 1] = (uint8_t)(w[i] >> 8);
            b[4 * i + 2] = (uint8_t)(w[i] >> 16);
            b[4 * i + 3] = (uint8_t)(w[i] >> 24);
        }
    }

    // A pending compression whose flags may still get ROOT added.
    struct Output {
        uint32_t in_cv[8];
        uint32_t block[16];
        uint64_t counter;
        uint32_t block_len;
        uint32_t flags;
        void chaining_value(uint32_t cv[8]) const {
            uint32_t o[16];
            compress(in_cv, block, counter, block_len, flags, o);
            for (int i = 0; i < 8; i++) cv[i] = o[i];
        }
        void root_bytes(uint8_t* out, size_t n) const {
            uint64_t t = 0;
            while (n > 0) {
                uint32_t o[16];
                uint8_t buf[64];
                compress(in_cv, block, t, block_len, flags | ROOT, o);
                le_from_words(o, buf, 16);
                size_t take = n < 64 ? n : 64;
                std::memcpy(out, buf, take);
                out += take;
                n -= take;
                t++;
            }
        }
    };

    // ---- chunk state ----
    uint32_t cs_cv[8];
    uint64_t cs_counter = 0;
    uint8_t cs_block[64];
    uint32_t cs_block_len = 0;
    uint32_t cs_blocks_compressed = 0;
    // ---- subtree chaining-value stack ----
    uint32_t stack[54][8];
    int stack_len = 0;

    Blake3() { reset_chunk(0); }

    void reset_chunk(uint64_t counter) {
        for (int i = 0; i < 8; i++) cs_cv[i] = IV[i];
        cs_counter = counter;
        std::memset(cs_block, 0, 64);
        cs_block_len = 0;
        cs_blocks_compressed = 0;
    }
    size_t chunk_len() const { return 64 * (size_t)cs_blocks_compressed + cs_block_len; }
    uint32_t start_flag() const { return cs_blocks_compressed == 0 ? (uint32_t)CHUNK_START : 0u; }

    void chunk_update(const uint8_t* in, size_t n) {
        while (n > 0) {
            if (cs_block_len == 64) {
                uint32_t w[16], o[16];
                words_from_le(cs_block, w, 16);
                compress(cs_cv, w, cs_counter, 64, start_flag(), o);
                for (int i = 0; i < 8; i++) cs_cv[i] = o[i];
                cs_blocks_compressed++;
                std::memset(cs_block, 0, 64);
                cs_block_len = 0;
            }
            size_t want = 64 - cs_block_len;
            size_t take = n < want ? n : want;
            std::memcpy(cs_block + cs_block_len, in, take);
            cs_block_len += (uint32_t)take;
            in += take;
            n -= take;
        }
    }
    Output chunk_output() const {
        Output o;
        for (int i = 0; i < 8; i++) o.in_cv[i] = cs_cv[i];
        words_from_le(cs_block, o.block, 16);
        o.counter = cs_counter;
        o.block_len = cs_block_len;
        o.flags = start_flag() | CHUNK_END;
        return o;
    }
    static Output parent_output(const uint32_t l[8], const uint32_t r[8]) {
        Output o;
        for (int i = 0; i < 8; i++) {
            o.in_cv[i] = IV[i];
            o.block[i] = l[i];
            o.block[8 + i] = r[i];
        }
        o.counter = 0;
        o.block_len = 64;
        o.flags = PARENT;
        return o;
    }
    void add_chunk_cv(uint32_t cv[8], uint64_t total_chunks) {
        while ((total_chunks & 1) == 0) {
            uint32_t merged[8];
            parent_output(stack[stack_len - 1], cv).chaining_value(merged);
            stack_len--;
            for (int i = 0; i < 8; i++) cv[i] = merged[i];
            total_chunks >>= 1;
        }
        for (int i = 0; i < 8; i++) stack[stack_len][i] = cv[i];
        stack_len++;
    }

    Blake3& update(const void* data, size_t n) {
        const uint8_t* in = (const uint8_t*)data;
        while (n > 0) {
            if (chunk_len() == 1024) {
                uint32_t cv[8];
                chunk_output().chaining_value(cv);
                uint64_t total = cs_counter + 1;
                add_chunk_cv(cv, total);
                reset_chunk(total);
            }
            size_t want = 1024 - chunk_len();
            size_t take = n < want ? n : want;
            chunk_update(in, take);
            in += take;
            n -= take;
        }
        return *this;
    }
    void finalize_xof(uint8_t* out, size_t n) const {
        Output o = chunk_output();
        int remaining = stack_len;
        while (remaining > 0) {
            remaining--;
            uint32_t cv[8];
            o.chaining_value(cv);
            o = parent_output(stack[remaining], cv);
        }
        o.root_bytes(out, n);
    }
    void finalize(uint8_t out[32]) const { finalize_xof(out, 32); }

    static void hash(const void* data, size_t n, uint8_t out[32]) {
        Blake3 h;
        h.update(data, n);
        h.finalize(out);
    }
};

}  // namespace oracle
