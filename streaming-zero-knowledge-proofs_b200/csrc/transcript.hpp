// Host-side Fiat-Shamir transcript of the prover (product code; never runs on the GPU — it is a
// serial sponge over < 10 KB).  Restates sezkp_crypto::Blake3Transcript
// (reference crates/sezkp-crypto/src/lib.rs:42, 74-124): a running BLAKE3 hasher with
//   absorb(label, bytes)   = "absorb" || u32 len(label) || label || u32 len(bytes) || bytes
//   challenge_bytes(l, n)  = XOF_n(clone || "challenge" || u32 len(l) || l), then the live state
//                            ratchets with "after_challenge" || u32 len(l) || l.
// BLAKE3 itself (third-party crate blake3 1.8.2 in the reference) is implemented here from the
// published specification: 1024-byte chunks of 64-byte blocks, binary tree of chunk chaining values,
// root compression re-run with an output-block counter for the XOF.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace host {

class Blake3Hasher {
  public:
    Blake3Hasher() { start_chunk(0); }
    void update(const void* data, size_t len) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        while (len) {
            if (chunk_bytes_ == 1024) {  // current chunk is full and more input follows: seal it
                uint32_t cv[8];
                seal_chunk(cv);
                push_subtree(cv, chunk_index_ + 1);
                start_chunk(chunk_index_ + 1);
            }
            if (buf_len_ == 64) {  // buffered block is full and more input follows: compress it (not the last block)
                uint32_t words[16], st[16];
                load_words(buf_, words);
                compress(cv_, words, chunk_index_, 64, block_flags(), st);
                std::memcpy(cv_, st, 32);
                blocks_done_++;
                buf_len_ = 0;
                std::memset(buf_, 0, 64);
            }
            size_t room = 64 - buf_len_;
            size_t chunk_room = 1024 - chunk_bytes_;
            size_t take = len < room ? len : room;
            if (take > chunk_room) take = chunk_room;
            std::memcpy(buf_ + buf_len_, p, take);
            buf_len_ += (uint32_t)take;
            chunk_bytes_ += (uint32_t)take;
            p += take;
            len -= take;
        }
    }
    // Extendable output; does not modify the hasher.
    void squeeze(uint8_t* out, size_t n) const {
        Node node;
        std::memcpy(node.cv, cv_, 32);
        load_words(buf_, node.block);
        node.counter = chunk_index_;
        node.block_len = buf_len_;
        node.flags = block_flags() | F_CHUNK_END;
        for (int i = depth_ - 1; i >= 0; i--) {
            uint32_t right[8], st[16];
            compress(node.cv, node.block, node.counter, node.block_len, node.flags, st);
            std::memcpy(right, st, 32);
            node = parent_node(stack_[i], right);
        }
        uint64_t block_no = 0;
        while (n) {
            uint32_t st[16];
            compress(node.cv, node.block, block_no++, node.block_len, node.flags | F_ROOT, st);
            uint8_t bytes[64];
            for (int i = 0; i < 16; i++) {
                bytes[4 * i] = (uint8_t)st[i];
                bytes[4 * i + 1] = (uint8_t)(st[i] >> 8);
                bytes[4 * i + 2] = (uint8_t)(st[i] >> 16);
                bytes[4 * i + 3] = (uint8_t)(st[i] >> 24);
            }
            size_t take = n < 64 ? n : 64;
            std::memcpy(out, bytes, take);
            out += take;
            n -= take;
        }
    }

  private:
    enum : uint32_t { F_CHUNK_START = 1, F_CHUNK_END = 2, F_PARENT = 4, F_ROOT = 8 };
    struct Node {
        uint32_t cv[8];
        uint32_t block[16];
        uint64_t counter;
        uint32_t block_len, flags;
    };
    static const uint32_t* iv() {
        static const uint32_t k[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                      0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        return k;
    }
    static uint32_t ror(uint32_t x, unsigned r) { return (x >> r) | (x << (32 - r)); }
    static void quarter(uint32_t* v, int a, int b, int c, int d, uint32_t x, uint32_t y) {
        v[a] += v[b] + x; v[d] = ror(v[d] ^ v[a], 16);
        v[c] += v[d];     v[b] = ror(v[b] ^ v[c], 12);
        v[a] += v[b] + y; v[d] = ror(v[d] ^ v[a], 8);
        v[c] += v[d];     v[b] = ror(v[b] ^ v[c], 7);
    }
    static void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t block_len, uint32_t flags,
                         uint32_t out[16]) {
        static const uint8_t sigma[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
        uint32_t v[16], m[16];
        std::memcpy(v, cv, 32);
        std::memcpy(v + 8, iv(), 16);
        v[12] = (uint32_t)counter;
        v[13] = (uint32_t)(counter >> 32);
        v[14] = block_len;
        v[15] = flags;
        std::memcpy(m, block, 64);
        for (int round = 0;; round++) {
            quarter(v, 0, 4, 8, 12, m[0], m[1]);
            quarter(v, 1, 5, 9, 13, m[2], m[3]);
            quarter(v, 2, 6, 10, 14, m[4], m[5]);
            quarter(v, 3, 7, 11, 15, m[6], m[7]);
            quarter(v, 0, 5, 10, 15, m[8], m[9]);
            quarter(v, 1, 6, 11, 12, m[10], m[11]);
            quarter(v, 2, 7, 8, 13, m[12], m[13]);
            quarter(v, 3, 4, 9, 14, m[14], m[15]);
            if (round == 6) break;
            uint32_t nm[16];
            for (int i = 0; i < 16; i++) nm[i] = m[sigma[i]];
            std::memcpy(m, nm, 64);
        }
        for (int i = 0; i < 8; i++) {
            out[i] = v[i] ^ v[i + 8];
            out[i + 8] = v[i + 8] ^ cv[i];
        }
    }
    static void load_words(const uint8_t* b, uint32_t* w) {
        for (int i = 0; i < 16; i++)
            w[i] = (uint32_t)b[4 * i] | (uint32_t)b[4 * i + 1] << 8 | (uint32_t)b[4 * i + 2] << 16 | (uint32_t)b[4 * i + 3] << 24;
    }
    static Node parent_node(const uint32_t left[8], const uint32_t right[8]) {
        Node n;
        std::memcpy(n.cv, iv(), 32);
        std::memcpy(n.block, left, 32);
        std::memcpy(n.block + 8, right, 32);
        n.counter = 0;
        n.block_len = 64;
        n.flags = F_PARENT;
        return n;
    }
    uint32_t block_flags() const { return blocks_done_ == 0 ? (uint32_t)F_CHUNK_START : 0u; }
    void start_chunk(uint64_t index) {
        std::memcpy(cv_, iv(), 32);
        chunk_index_ = index;
        std::memset(buf_, 0, 64);
        buf_len_ = 0;
        blocks_done_ = 0;
        chunk_bytes_ = 0;
    }
    void seal_chunk(uint32_t cv_out[8]) const {
        uint32_t words[16], st[16];
        load_words(buf_, words);
        compress(cv_, words, chunk_index_, buf_len_, block_flags() | F_CHUNK_END, st);
        std::memcpy(cv_out, st, 32);
    }
    // Merge completed subtrees: after `total` chunks, the stack holds one cv per set bit of `total`.
    void push_subtree(uint32_t cv[8], uint64_t total) {
        while ((total & 1) == 0) {
            Node n = parent_node(stack_[depth_ - 1], cv);
            uint32_t st[16];
            compress(n.cv, n.block, 0, 64, F_PARENT, st);
            std::memcpy(cv, st, 32);
            depth_--;
            total >>= 1;
        }
        std::memcpy(stack_[depth_++], cv, 32);
    }

    uint32_t cv_[8];
    uint64_t chunk_index_ = 0;
    uint8_t buf_[64];
    uint32_t buf_len_ = 0, blocks_done_ = 0, chunk_bytes_ = 0;
    uint32_t stack_[54][8];
    int depth_ = 0;
};

class Transcript {
  public:
    explicit Transcript(const std::string& domain) {
        h_.update("sezkp.transcript.v0", 19);
        frame(h_, domain);
    }
    void absorb(const std::string& label, const void* bytes, size_t n) {
        h_.update("absorb", 6);
        frame(h_, label);
        uint32_t len = (uint32_t)n;
        h_.update(&len, 4);
        h_.update(bytes, n);
    }
    void absorb_u64(const std::string& label, uint64_t x) { absorb(label, &x, 8); }
    std::vector<uint8_t> challenge(const std::string& label, size_t n) {
        Blake3Hasher fork = h_;
        fork.update("challenge", 9);
        frame(fork, label);
        std::vector<uint8_t> out(n);
        fork.squeeze(out.data(), n);
        h_.update("after_challenge", 15);
        frame(h_, label);
        return out;
    }

  private:
    static void frame(Blake3Hasher& h, const std::string& s) {
        uint32_t len = (uint32_t)s.size();
        h.update(&len, 4);
        h.update(s.data(), s.size());
    }
    Blake3Hasher h_;
};

}  // namespace host
