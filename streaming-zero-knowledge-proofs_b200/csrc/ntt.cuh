// Internal interface of ntt.cu (device-pointer level).
#pragma once
#include "common.cuh"

// In-place batched NTT/iNTT of `cols` vectors of length 2^L stored back to back at `data` (device).
// `tmp` (device, cols*2^L elements) is required when L > 10.
void ntt_batch_device(sezkp_ctx* ctx, u64* data, u64* tmp, int L, u64 cols, bool inverse);
// Coset LDE: coeffs [cols][2^L] -> out [cols][2^(L+logB)], out[i] = f(shift * w^i); `inter` (device,
// cols*2^(L+logB) elements) is required when L > 10.  coeffs are not modified.
// fuse != null (K7): the last pass hashes the extended values (labeled leaves, template of batch column v = templates[v]) and
// reduces every aligned group of 32 leaves to its sub-root, written to level 0 (2^5-leaf chunks) of column v's retained tree at
// upper + v * col_words; nothing is written to `out` (may be null).  Only when lde_hash_fusable(L, logB).
struct LdeHashFuse {
    const void* templates;  // device b3::LabelTemplate[cols]
    u32* upper;
    u64 col_words;
};
bool lde_hash_fusable(int L, int logB);
void coset_lde_device(sezkp_ctx* ctx, const u64* coeffs, u64* out, u64* inter, int L, int logB, u64 shift, u64 cols,
                      const LdeHashFuse* fuse = nullptr);
void ntt_free_tables(sezkp_ctx* ctx);
