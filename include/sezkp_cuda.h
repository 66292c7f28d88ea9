/* sezkp_cuda.h — C ABI of libsezkp_cuda.so: the B200 (sm_100a) implementation of SEZKP's STARK v1
 * commitment hot path.  This is the surface a Rust `sezkp-cuda` FFI crate binds;
 * every entry point names the reference item it replaces (paths relative to the reference's crates/).
 *
 * Conventions (mirroring the reference's FFI stub, sezkp-ffi/src/lib.rs:49-94):
 *   - every function returns int32_t: 0 = OK, <0 = SEZKP_CUDA_E*; nothing aborts or throws across the ABI;
 *   - sezkp_cuda_last_error(ctx) is a ctx-owned NUL-terminated string, valid until the next call on ctx;
 *   - the caller owns every input and output buffer (explicit sizes / capacities); the library never
 *     returns memory the caller must free except opaque handles with a matching *_free;
 *   - field elements are canonical Goldilocks residues (< p = 2^64-2^32+1) as little-endian uint64_t,
 *     digests are 32 raw bytes; pointers named *_dev are device pointers on the ctx's GPU, all others host;
 *   - a ctx is bound to one GPU (sezkp_cuda_create) or to several GPUs of one box driven from this one process
 *     (sezkp_cuda_create_multi); it is not thread-safe: one call at a time per ctx;
 *   - there is no CPU fallback: without a usable CUDA device sezkp_cuda_create fails with ENODEV.
 */
#ifndef SEZKP_CUDA_H
#define SEZKP_CUDA_H
#include <stddef.h>
#include <stdint.h>

#include "sezkp_trace.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SEZKP_CUDA_ABI_VERSION 1u  /* additions only since round 1: still 1 */

#define SEZKP_CUDA_OK 0
#define SEZKP_CUDA_EINVAL (-1)  /* bad argument (non power-of-two size, label too long, NULL, ...) */
#define SEZKP_CUDA_ENOMEM (-2)  /* host or device allocation failed                               */
#define SEZKP_CUDA_ECUDA (-3)   /* CUDA runtime / kernel error                                    */
#define SEZKP_CUDA_ENODEV (-4)  /* no usable CUDA device                                          */
#define SEZKP_CUDA_ERANGE (-5)  /* output buffer too small (required size reported)               */
#define SEZKP_CUDA_ESTATE (-6)  /* call sequence error (streaming API)                            */
#define SEZKP_CUDA_ECOMM (-7)   /* the caller-supplied collective callback failed                 */

typedef struct sezkp_ctx sezkp_ctx;
typedef struct sezkp_tree sezkp_tree; /* retained column commitments (chunk roots + upper levels + values) */
typedef struct sezkp_fri sezkp_fri;   /* retained FRI layers (values + upper tree levels)                  */
typedef struct sezkp_stream sezkp_stream;
typedef struct sezkp_trace_dev sezkp_trace_dev; /* compact trace resident in HBM */
typedef struct sezkp_columns sezkp_columns;     /* base-domain columns resident in HBM, column c on GPU c % n_gpus */

/* ---------------------------------------------------------------- context ---- */
uint32_t sezkp_cuda_abi_version(void);                       /* cf. sezkp_abi_version(), sezkp-ffi/src/lib.rs:55-68 */
int32_t sezkp_cuda_create(int device_id, sezkp_ctx** out);   /* device_id < 0: current device                     */
/* One context over n_dev GPUs of this box, driven from this single process — what the reference's stateless, single-
 * process backend call (sezkp-core/src/backend.rs:41-61; CLI arms sezkp-cli/src/main.rs:503-513) needs to use a whole
 * 8-GPU box (SURVEY §8b "threading").  The library keeps one worker thread + one internal context per GPU and implements
 * every exchange itself over NVLink peer access (column roots, FRI subtree roots, opening records, compact trace, partial
 * combination sums): no NCCL, no callbacks.  With such a ctx
 *   sezkp_stark_v1_prove / _prove_resident / the streaming + JSONL entry points   shard ONE proof over the GPUs
 *       (columns c % n_dev, FRI leaf hashing by chunk range; proof bytes identical to the single-GPU proof),
 *   sezkp_lde_commit_batch, sezkp_columns_*, sezkp_lde_commit_fri                   shard columns c % n_dev,
 * and every other entry point runs on device_ids[0] as with a plain ctx.  sezkp_cuda_destroy releases all of it.
 * A device id may be listed more than once (its ranks then share that GPU): useful for testing on a single GPU. */
int32_t sezkp_cuda_create_multi(const int* device_ids, int n_dev, sezkp_ctx** out);
int32_t sezkp_cuda_group_size(const sezkp_ctx* ctx);         /* GPUs behind this ctx (1 for sezkp_cuda_create)       */
int32_t sezkp_cuda_device_count(void);                       /* usable CUDA devices in this process (0: none)        */
void sezkp_cuda_destroy(sezkp_ctx* ctx);
const char* sezkp_cuda_last_error(const sezkp_ctx* ctx);     /* ctx may be NULL: error of the last failed create   */
/* use_own != 0: (re)create a private non-blocking stream; else adopt the caller's cudaStream_t (NULL = legacy default
 * stream), e.g. torch.cuda.current_stream().cuda_stream, so that the caller's events bracket the library's work */
/* Stream contract of every *_dev entry point: the library launches on the context's stream only and never waits for
 * other streams.  Device buffers handed in must be complete with respect to that stream (write them on it, or
 * synchronise first — the private stream is non-blocking, so not even the legacy default stream orders with it), and
 * results are complete once sezkp_cuda_synchronize returns (entry points that return host data synchronise themselves).
 * The *_dev NTT/LDE entry points accept any 64-bit representative of a field element; host-buffer entry points
 * reject non-canonical values with EINVAL. */
int32_t sezkp_cuda_set_stream(sezkp_ctx* ctx, void* cuda_stream, int use_own);
int32_t sezkp_cuda_synchronize(sezkp_ctx* ctx);
/* tuning switches; "dedup" (default 1): value-aware column commit that hashes identical leaves / identical sibling
 * pairs of a chunk once (outputs are identical either way; 0 forces one compression per node); "tabled" (default 1):
 * subtree tables for structured columns; "deep_fused" (default 0): one-launch DEEP kernel also for large domains; "tab_cache" (default 1): keep the
 * subtree tables across proofs while their parameters and labels are unchanged; "ntt_gen" (default 3): NTT pass-kernel
 * generation (1 = round-1 kernel with three shared-memory round trips, 2-4 = global loads/stores fused into the register
 * DFTs, differing in register budget / CTA size; A/B measurements); "lde_fuse" (default 1): fuse the last pass of the LDE
 * with the labeled leaf hash in sezkp_lde_commit_batch / sezkp_lde_commit_fri (north_star item 3; same roots); "phase_sync"
 * (default 1): how sezkp_cuda_get_timings clocks the prover's phases — 1 = host clock with a stream synchronisation at every
 * phase boundary, 0 = CUDA events read back at the end of the proof; "fri_coset" (default 1): one proof over a context group
 * with peer access keeps the FRI layers coset-resident (local folds, each GPU materialises only its own hashing range;
 * 0 = all-gather of layer 0 and replicated folds; same proof bytes) */
int32_t sezkp_cuda_set_option(sezkp_ctx* ctx, const char* name, int64_t value);
/* number of kernels launched by this ctx since creation / since the last reset */
uint64_t sezkp_cuda_launch_count(sezkp_ctx* ctx, int reset);
/* JSON object {"phase": ms, ...} of the last sezkp_stark_v1_prove on this ctx */
int32_t sezkp_cuda_get_timings(sezkp_ctx* ctx, char* json_buf, size_t cap);
/* the same for GPU `rank` of a multi-GPU context (rank 0 = sezkp_cuda_get_timings) */
int32_t sezkp_cuda_get_timings_gpu(sezkp_ctx* ctx, int rank, char* json_buf, size_t cap);

/* ------------------------------------------------------------ NTT / LDE ------ */
/* forward_ntt_in_place / inverse_ntt_in_place per column (sezkp-ffts/src/ntt.rs:79-111, 117-155):
 * data is [cols][1<<log_n], natural order in and out, w_N = 7^((p-1)/N); inverse scales by N^-1. */
int32_t sezkp_ntt_batch(sezkp_ctx* ctx, uint64_t* data, int log_n, int cols, int inverse);
int32_t sezkp_ntt_batch_dev(sezkp_ctx* ctx, uint64_t* data_dev, int log_n, int cols, int inverse);
/* evaluate_on_coset_pow2(coeffs, log_n+log_blow, shift) per column (sezkp-ffts/src/coset.rs:85-102):
 * coeffs [cols][1<<log_n] -> out [cols][1<<(log_n+log_blow)], out[i] = f(shift * w_N^i). */
int32_t sezkp_coset_lde_batch(sezkp_ctx* ctx, const uint64_t* coeffs, int log_n, int log_blow, uint64_t shift, int cols,
                              uint64_t* out);
int32_t sezkp_coset_lde_batch_dev(sezkp_ctx* ctx, const uint64_t* coeffs_dev, int log_n, int log_blow, uint64_t shift,
                                  int cols, uint64_t* out_dev);
/* interpolate_from_evals (ntt.rs:173-177) then evaluate_on_coset_pow2, per column; evals are NOT modified. */
int32_t sezkp_lde_from_evals_batch(sezkp_ctx* ctx, const uint64_t* evals, int log_n, int log_blow, uint64_t shift,
                                   int cols, uint64_t* out);
int32_t sezkp_lde_from_evals_batch_dev(sezkp_ctx* ctx, const uint64_t* evals_dev, int log_n, int log_blow, uint64_t shift,
                                       int cols, uint64_t* out_dev);
/* deep_coset_lde_stream (sezkp-stark/src/v1/lde.rs:42-97) without the chunked callback: base_evals[n] ->
 * out[n<<log_blow], out[i] = f(shift*w^i) / (shift*w^i - z).  z on the coset -> EINVAL. */
int32_t sezkp_deep_lde(sezkp_ctx* ctx, const uint64_t* base_evals, int log_n, int log_blow, uint64_t shift, uint64_t z,
                       uint64_t* out);
int32_t sezkp_deep_lde_dev(sezkp_ctx* ctx, const uint64_t* base_evals_dev, int log_n, int log_blow, uint64_t shift,
                           uint64_t z, uint64_t* out_dev);

/* ------------------------------------------------------ hashing / Merkle ----- */
/* hash_field_leaves (label NULL; v1/merkle.rs:150-159, v1/fri_stream.rs:37-41) or
 * hash_field_leaves_labeled (v1/merkle.rs:132-146): vals[n] -> out[n][32].  strlen(label) <= 44. */
int32_t sezkp_leaf_hash(sezkp_ctx* ctx, const uint64_t* vals, size_t n, const char* label_or_null, uint8_t* out);
/* MerkleTree::from_leaves(leaves).root() (v1/merkle.rs:46-77) == sezkp_merkle::merkle_root for n >= 1
 * (sezkp-merkle/src/lib.rs:140-157): BLAKE3(left||right) parents, odd node promoted unchanged. */
int32_t sezkp_merkle_root(sezkp_ctx* ctx, const uint8_t* leaves, size_t n, uint8_t out_root[32]);
/* OnDemandOpenings::build_roots over arbitrary columns (v1/openings.rs:306-398): cols [c][n] with one label per
 * column, labeled leaves -> 2^chunk_log2-row chunk trees -> outer tree over chunk roots.  n must be a power of two
 * (then the result equals one binary tree over n labeled leaves).  keep != NULL retains what openings need.
 * labels == NULL: unlabeled leaves BLAKE3(le8), i.e. the root StreamingLayerBuilder computes for an FRI layer
 * (v1/fri_stream.rs:55-122). */
int32_t sezkp_column_commit_batch(sezkp_ctx* ctx, const uint64_t* cols, const char* const* labels, int c, size_t n,
                                  int chunk_log2, uint8_t* roots /* [c][32] */, sezkp_tree** keep_or_null);
int32_t sezkp_column_commit_batch_dev(sezkp_ctx* ctx, const uint64_t* cols_dev, const char* const* labels, int c, size_t n,
                                      int chunk_log2, uint8_t* roots /* host [c][32] */, sezkp_tree** keep_or_null);
/* LDE + commitment of extended columns (BASELINE.json configs[3]): per column interpolate_from_evals
 * (ntt.rs:173-177) -> evaluate_on_coset_pow2 (coset.rs:85-102) -> hash_field_leaves_labeled over the extended column ->
 * chunked tree root, processed in column groups so the extended columns are never all resident.  evals [c][1<<log_n]. */
int32_t sezkp_lde_commit_batch(sezkp_ctx* ctx, const uint64_t* evals, const char* const* labels, int c, int log_n, int log_blow,
                               uint64_t shift, int chunk_log2, uint8_t* roots /* [c][32] */);
int32_t sezkp_lde_commit_batch_dev(sezkp_ctx* ctx, const uint64_t* evals_dev, const char* const* labels, int c, int log_n,
                                   int log_blow, uint64_t shift, int chunk_log2, uint8_t* roots /* host [c][32] */);
/* BASELINE.json configs[3] / SURVEY §8d config 4 as one call: per column LDE + labeled commitment as above, then the
 * transcript of prove_v1 restricted to what the shape has — new("sezkp-stark/v1"), absorb n, n_cols and every col_root,
 * challenge "alphas" (8 bytes per column, from_u64 each; sezkp-crypto/src/lib.rs:74-124, v1/params.rs:76-126) — the
 * base-domain combination C(i) = sum_c alpha_c * col_c[i], the OOD point nudged off the coset (v1/prover.rs:118-135),
 * deep_coset_lde_stream (v1/lde.rs:42-97) and the FRI fold-and-commit loop (v1/prover.rs:184-243; no masks, no queries).
 * The columns are a resident set (uploaded from host arrays, or synthesised on the device with SURVEY §8d's generator
 * value(c,i) = splitmix64_step(seed ^ c<<40 ^ i) mod p, seed 0x5EED in the benchmark); on a multi-GPU ctx column c lives on
 * GPU c % n_dev and the exchanges are: all-gather of column roots (C1), one kernel per GPU that sums the partial
 * combinations straight out of the peers' HBM (C3), all-gather of FRI subtree roots (C2).
 * Outputs (host): col_roots [c][32], fri_roots [log_n+log_blow+1][32], the final FRI value. */
int32_t sezkp_columns_upload(sezkp_ctx* ctx, const uint64_t* evals /* [c][1<<log_n] */, int c, int log_n, sezkp_columns** out);
int32_t sezkp_columns_synth(sezkp_ctx* ctx, uint64_t seed, int c, int log_n, sezkp_columns** out);
void sezkp_columns_free(sezkp_ctx* ctx, sezkp_columns* cols);
int32_t sezkp_lde_commit_fri(sezkp_ctx* ctx, const sezkp_columns* cols, const char* const* labels, int log_blow, uint64_t shift,
                             int chunk_log2, uint8_t* col_roots, uint8_t* fri_roots, uint64_t* final_value);
/* OnDemandOpenings::open (v1/openings.rs:403-497) for k (column, row) pairs.  Per opening the outputs are
 * value (8 B LE), chunk_root (32 B), path_in_chunk (min(chunk_log2, log2 n) siblings), path_to_chunk (the rest);
 * sibling arrays are [k][depth][32] with depth_in / depth_out returned. */
int32_t sezkp_column_open(sezkp_ctx* ctx, const sezkp_tree* tree, const uint32_t* col_idx, const uint64_t* row_idx, size_t k,
                          uint64_t* values, uint8_t* chunk_roots, uint8_t* path_in_chunk, uint8_t* path_to_chunk,
                          int* depth_in, int* depth_out);
/* Batched verification of openings: verify_chunked_open (v1/merkle.rs:243-280; ColumnCommit / OnDemandOpenings proofs, as
 * checked by verify_v1 v1/verify.rs:60-196) and MerkleTree::verify (v1/merkle.rs:111-126; FRI layer paths, v1/fri.rs:130-222)
 * for k openings in one launch.  Per opening i: the leaf of values[i] — labeled with labels[col_idx[i]], or unlabeled when
 * labels is NULL — is walked up path_in_chunk[i] (depth_in siblings, position index_in_chunk[i]); if chunk_roots is given the
 * result must equal chunk_roots[i]; then up path_to_chunk[i] (depth_out siblings, position chunk_index[i]) to
 * col_roots[col_idx[i]].  ok[i] = 1 iff every comparison holds (and values[i] is canonical).  col_idx may be NULL when c == 1;
 * a plain Merkle path is depth_out = 0, chunk_roots = NULL.  The reference's verifier itself stays on the CPU. */
int32_t sezkp_verify_openings(sezkp_ctx* ctx, const uint8_t* col_roots /* [c][32] */, const char* const* labels_or_null, int c,
                              const uint32_t* col_idx, const uint64_t* values, const uint64_t* index_in_chunk,
                              const uint64_t* chunk_index, const uint8_t* chunk_roots /* [k][32] or NULL */,
                              const uint8_t* path_in_chunk /* [k][depth_in][32] */, int depth_in,
                              const uint8_t* path_to_chunk /* [k][depth_out][32] */, int depth_out, size_t k, uint8_t* ok /* [k] */);
void sezkp_tree_free(sezkp_ctx* ctx, sezkp_tree* tree);

/* ------------------------------------------------------------------ FRI ------ */
/* FRI fold-and-commit for given folding challenges (v1/prover.rs:184-243; same layers as v1/fri.rs:71-91):
 * layer0[1<<log_N]; betas[log_N]; y'[i] = y[i] + beta_r*y[i+half]; roots [log_N+1][32] (layer 0 first) and the final
 * value.  keep != NULL retains the layers for sezkp_fri_open. */
int32_t sezkp_fri_commit(sezkp_ctx* ctx, const uint64_t* layer0, int log_N, const uint64_t* betas, uint8_t* roots,
                         uint64_t* final_value, sezkp_fri** keep_or_null);
int32_t sezkp_fri_commit_dev(sezkp_ctx* ctx, const uint64_t* layer0_dev, int log_N, const uint64_t* betas, uint8_t* roots,
                             uint64_t* final_value, sezkp_fri** keep_or_null);
/* FRI query openings (v1/prover.rs:297-450; compat v1/fri.rs:98-127) for k layer-0 indices: per query and per layer
 * l < log_N: values[q][l][2] = (y_l[idx], y_l[idx^half]), paths[q][l][2][log_N - l][32] stored with a fixed
 * pitch of log_N siblings per path (unused tail zero), positions[q][log_N+1]. */
int32_t sezkp_fri_open(sezkp_ctx* ctx, const sezkp_fri* fri, const uint64_t* idx0, size_t k, uint64_t* positions,
                       uint64_t* values, uint8_t* paths);
void sezkp_fri_free(sezkp_ctx* ctx, sezkp_fri* fri);

/* --------------------------------------------------- feeder (columns + AIR) -- */
/* TraceColumns::build committed columns (v1/columns.rs:252-365) in all_labels order (v1/openings.rs:89-116):
 * out [3+7*tau][n_rows]. */
int32_t sezkp_trace_columns(sezkp_ctx* ctx, const sezkp_trace_desc* trace, uint64_t* out);
/* base-domain composition C(i)+B(i)+R(w^i) (compose_row/compose_boundary v1/air.rs:49-136, eval_masks_sum_at
 * v1/masking.rs:86-103, closure at v1/prover.rs:142-158): alphas8 as drawn by derive_alphas, one mask polynomial
 * of mask_deg ascending coefficients; out[n_rows]. */
int32_t sezkp_compose_base(sezkp_ctx* ctx, const sezkp_trace_desc* trace, const uint64_t alphas8[8],
                           const uint64_t* mask_coeffs, size_t mask_deg, uint64_t* out);

/* --------------------------------------------------------------- prover ------ */
/* prove_v1 + bincode::serialize (v1/prover.rs:61-462, sezkp-stark/src/lib.rs:129-142): writes the ProofV1 bytes
 * (v1/proof.rs:80-98, bincode 1.3 default config).  proof_buf may be NULL to query *len; cap too small -> ERANGE
 * with *len set.  The Fiat-Shamir transcript (sezkp-crypto/src/lib.rs:74-124) runs on the host inside the library. */
int32_t sezkp_stark_v1_prove(sezkp_ctx* ctx, const sezkp_trace_desc* trace, const uint8_t manifest_root[32],
                             uint8_t* proof_buf, size_t cap, size_t* len);
/* Column-sharded prover for the GPUs of one box, one process (and one ctx) per GPU: rank r commits and
 * opens only the columns c with c % world == r; everything that depends on all columns (composition, LDE, FRI) is
 * replicated, except the leaf / chunk-tree hashing of the large FRI layers, which is split by chunk range.  The exchange
 * steps are all-gathers of small host buffers — the 32-byte column roots, the FRI subtree roots (twice) and the opening
 * records — done through `allgather`, which the host binding implements over NCCL (or any collective):
 * it must fill recv_all[world][bytes] with every rank's `send`, rank-major, and return 0.  Every rank returns the
 * identical proof, byte-equal to sezkp_stark_v1_prove. */
typedef int32_t (*sezkp_allgather_fn)(void* user, const void* send, size_t bytes, void* recv_all);
int32_t sezkp_stark_v1_prove_sharded(sezkp_ctx* ctx, const sezkp_trace_desc* trace, const uint8_t manifest_root[32], int rank,
                                     int world, sezkp_allgather_fn allgather, void* user, uint8_t* proof_buf, size_t cap,
                                     size_t* len);
/* Same prover over a compact trace that is already resident in HBM (upload once, prove many times: lets a caller
 * overlap the next trace's H2D copy with the current proof, and separates copy time from kernel time). */
int32_t sezkp_trace_upload(sezkp_ctx* ctx, const sezkp_trace_desc* trace, sezkp_trace_dev** out);
void sezkp_trace_free(sezkp_ctx* ctx, sezkp_trace_dev* trace);
int32_t sezkp_stark_v1_prove_resident(sezkp_ctx* ctx, const sezkp_trace_dev* trace, const uint8_t manifest_root[32],
                                      uint8_t* proof_buf, size_t cap, size_t* len);
/* Optional device-side collective for the sharded prover (NCCL over NVLink in the host binding): all-gather `bytes`
 * bytes from every rank's `send_dev` into `recv_all_dev` ([world][bytes], rank-major), both DEVICE pointers, ordered on
 * `cuda_stream` (the context's stream: enqueue the collective there, or make that stream wait for it) — return 0.
 * When registered, sezkp_stark_v1_prove_sharded uploads only rows [n*r/world, n*(r+1)/world) of the trace from the
 * host and all-gathers the compact trace between the GPUs (instead of world replicated PCIe uploads), and the FRI subtree
 * roots are exchanged without a host round trip.  Needs n_rows % world == 0; otherwise the host-callback path is used. */
typedef int32_t (*sezkp_allgather_dev_fn)(void* user, const void* send_dev, size_t bytes, void* recv_all_dev, void* cuda_stream);
int32_t sezkp_cuda_set_allgather_dev(sezkp_ctx* ctx, sezkp_allgather_dev_fn fn, void* user);
/* The sharded prover over a trace every rank already holds in HBM (each rank uploaded the same trace): what the N-GPU
 * single-proof latency is without the N-fold replicated H2D copy of sezkp_stark_v1_prove_sharded. */
int32_t sezkp_stark_v1_prove_resident_sharded(sezkp_ctx* ctx, const sezkp_trace_dev* trace, const uint8_t manifest_root[32],
                                              int rank, int world, sezkp_allgather_fn allgather, void* user,
                                              uint8_t* proof_buf, size_t cap, size_t* len);
/* ProvingBackendStream (sezkp-core/src/prover.rs:21-33): begin_stream / ingest_block / finish_stream, driven like
 * StreamingProver::prove_stream_iter (prover.rs:104-150) over a block iterator (core/io.rs:111-139).  Each ingest pushes
 * one or more blocks (a descriptor whose arrays cover just those blocks); rows are packed into pinned staging buffers
 * and copied to the GPU on a side stream every 2^20 rows while the host keeps parsing, so at finish only the tail is
 * still in flight.  expected_rows (0 = unknown) pre-sizes the device trace.  finish consumes the handle on success;
 * sezkp_cuda_get_timings then also reports stream_h2d_copy_ms, stream_copy_exposed_ms and stream_copy_hidden_frac. */
int32_t sezkp_stark_v1_begin(sezkp_ctx* ctx, uint32_t tau, const uint8_t manifest_root[32], uint64_t expected_rows,
                             sezkp_stream** out);
int32_t sezkp_stark_v1_ingest(sezkp_ctx* ctx, sezkp_stream* st, const sezkp_trace_desc* blocks);
int32_t sezkp_stark_v1_finish(sezkp_ctx* ctx, sezkp_stream* st, uint8_t* proof_buf, size_t cap, size_t* len);
void sezkp_stark_v1_abort(sezkp_ctx* ctx, sezkp_stream* st);
/* Native JSONL front-end (reference: stream_block_summaries_jsonl, crates/sezkp-core/src/io_jsonl.rs:27-88; one
 * serde-JSON BlockSummary per line, blank lines skipped).  The reference's stark arm rejects .jsonl input
 * (sezkp-core/src/io.rs:78-88); these entry points are what its CLI would call to accept it.
 *  - sezkp_jsonl_parse: text[0,len) (whole lines) -> a library-owned compact trace, parsed on n_threads host threads
 *    (<= 0: all cores).  No ctx and no GPU needed.  *desc points into the handle; scalars (may be NULL) receives a
 *    pointer to n_blocks records.  Errors: EINVAL with sezkp_jsonl_last_error() = "jsonl line N: ...".
 *  - sezkp_stark_v1_ingest_jsonl: parse + ingest_block for every block of the text, in file order.
 *  - sezkp_stark_v1_prove_jsonl_file: the whole StreamingProver loop over a file taken in chunk_bytes pieces
 *    (0: 64 MiB).  Regular files are mapped: one pool of n_threads host threads parses a piece and then packs it into the
 *    pinned staging ring, whose H2D copies run on a side stream behind the parsing of the next pieces; pipes and special
 *    files are read with a double-buffered fread loop.  tau comes from the first block.  sezkp_cuda_get_timings then also
 *    reports jsonl_read_ms, jsonl_parse_ms, jsonl_pack_ms and jsonl_bytes. */
typedef struct sezkp_jsonl_trace sezkp_jsonl_trace;
int32_t sezkp_jsonl_parse(const char* text, size_t len, int n_threads, sezkp_jsonl_trace** out, sezkp_trace_desc* desc,
                          const sezkp_block_scalars** scalars);
void sezkp_jsonl_free(sezkp_jsonl_trace* t);
const char* sezkp_jsonl_last_error(void);                     /* thread-local                                       */
/* The inverse (what the reference CLI's export-jsonl writes, crates/sezkp-cli/src/main.rs): one serde_json BlockSummary
 * per line, field order of sezkp-core/src/types.rs:116-151, formatted on n_threads host threads (<= 0: all cores).
 * scalars may be NULL (version 1, block ids from 1, step ranges from the block lengths).  No ctx and no GPU needed. */
int32_t sezkp_jsonl_write_file(const char* path, const sezkp_trace_desc* trace, const sezkp_block_scalars* scalars, int n_threads,
                               uint64_t* bytes_written);
int32_t sezkp_stark_v1_ingest_jsonl(sezkp_ctx* ctx, sezkp_stream* st, const char* text, size_t len, int n_threads,
                                    uint64_t* n_blocks, uint64_t* n_rows);
int32_t sezkp_stark_v1_prove_jsonl_file(sezkp_ctx* ctx, const char* path, const uint8_t manifest_root[32], int n_threads,
                                        size_t chunk_bytes, uint64_t expected_rows, uint8_t* proof_buf, size_t cap,
                                        size_t* len);
/* upper bound of the serialized ProofV1 size for n_rows rows and tau tapes (for sizing proof_buf) */
size_t sezkp_stark_v1_proof_bound(uint64_t n_rows, uint32_t tau);

#ifdef __cplusplus
}
#endif
#endif /* SEZKP_CUDA_H */
