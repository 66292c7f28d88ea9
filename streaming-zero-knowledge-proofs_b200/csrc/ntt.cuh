// Internal interface of ntt.cu (device-pointer level).
#pragma once
#include "common.cuh"

// In-place batched NTT/iNTT of `cols` vectors of length 2^L stored back to back at `data` (device).
// `tmp` (device, cols*2^L elements) is required when L > 10.
void ntt_batch_device(sezkp_ctx* ctx, u64* data, u64* tmp, int L, u64 cols, bool inverse);
// Coset LDE: coeffs [cols][2^L] -> out [cols][2^(L+logB)], out[i] = f(shift * w^i); `inter` (device,
// cols*2^(L+logB) elements) is required when L > 10.  coeffs are not modified.
void coset_lde_device(sezkp_ctx* ctx, const u64* coeffs, u64* out, u64* inter, int L, int logB, u64 shift, u64 cols);
void ntt_free_tables(sezkp_ctx* ctx);
