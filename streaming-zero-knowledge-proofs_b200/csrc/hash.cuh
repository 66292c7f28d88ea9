// Internal interface of hash.cu: chunked BLAKE3 Merkle commitments over device-resident columns.
#pragma once
#include "blake3.cuh"
#include <vector>

#include "common.cuh"

// A commitment to `cols` columns of `n` field elements each (n a power of two).  Leaves are labeled
// (templates != null) or unlabeled.  Only the chunk roots and the levels above them are retained
// (`upper`); openings rebuild the 2^cl-leaf chunk on demand, like the reference's
// OnDemandOpenings::open_within_chunk (v1/openings.rs:464-497).
struct Commit {
    const u64* values = nullptr;   // device [cols][n]
    bool owns_values = false;
    u64 n = 0;
    u64 col_stride = 0;            // elements between consecutive committed columns (n when dense)
    int cols = 0;
    int cl = 0;                    // log2 leaves per chunk = min(chunk_log2, log2 n)
    int cta_cl = 0;                // log2 leaves one CTA of the plain chunk kernel reduces (>= cl): it then writes 2^(cta_cl-cl) chunk roots
    u64 n_ch = 0;                  // chunks per column = n >> cl
    u32* upper = nullptr;          // device [cols][2*n_ch-1][8]: level l (count n_ch>>l) at offset 2*n_ch-(2*n_ch>>l)
    b3::LabelTemplate* templates = nullptr;  // device [cols] or null (unlabeled)
    std::vector<b3::LabelTemplate> templates_host;  // the same on the host (key of the subtree-table cache)
    // subtree-table state of the value-aware commit (hash.cu, "structured columns"); null/empty when not used
    void* tab_dev = nullptr;       // device: ColTab[cols], column lists, work list of chunks to redo
    std::vector<int> tab_logg;     // per column: log2 of the leaves under one table entry, 0 = generic kernel
    bool tab_classified = false;
    u64 tab_chunks_done = 0;       // chunks hashed so far (the work list refers to them)
    void release(sezkp_ctx* ctx);
};
inline u64 upper_off(u64 n_ch, int l) { return 2 * n_ch - ((2 * n_ch) >> l); }

b3::LabelTemplate make_label_template(const char* label);
struct CommitOpts {
    bool dedup = false;             // value-aware kernel (identical leaves / sibling pairs hashed once); same outputs
    const u64* fold_src = nullptr;  // fused FRI fold: values[i] = fold_src[i] + fold_beta*fold_src[i+n], written + hashed
    u64 fold_beta = 0;
    u8* roots_host = nullptr;       // [cols][32]; copying to the host synchronises the stream
    u8* roots_dev = nullptr;        // [cols][32] device copy (no synchronisation)
    u64 col_stride = 0;             // 0 = dense (n); column sharding commits every world-th column of a dense array
    int cta_log2 = 0;               // > chunk_log2: one CTA still reduces 2^cta_log2 leaves but stops at the 2^chunk_log2-leaf
                                    // sub-roots (plain / fused-fold kernel only): the thin top of every CTA tree moves to
                                    // upper_reduce, where a CTA is full again
};
// Build a commitment over device values (all launches on ctx->stream).
void commit_build(sezkp_ctx* ctx, Commit& cm, const u64* values_dev, u64 n, int cols, int chunk_log2,
                  const char* const* labels_or_null, const CommitOpts& opt);
void commit_begin(sezkp_ctx* ctx, Commit& cm, const u64* values_dev, u64 n, int cols, int chunk_log2, const char* const* labels_or_null,
                  const CommitOpts& opt);
void commit_chunks(sezkp_ctx* ctx, Commit& cm, u64 chunk0, u64 chunk1, const CommitOpts& opt);
void commit_finish(sezkp_ctx* ctx, Commit& cm, const CommitOpts& opt);
// Upper levels of `count` single-column commitments (begin + chunks done) in at most three launches; root i goes to
// roots_dev + 32*i (device memory, may be null).
constexpr int UPPER_MAX_JOBS = 40;
struct UpperJob {
    u32* upper;
    u64 n_ch;
    int l0, k;
    u32* root_out;
    u64 grp0;  // first group (of 2^k nodes at level l0) this job reduces; its CTA count is the number of groups
};
struct UpperJobs {
    UpperJob j[UPPER_MAX_JOBS];
    u32 cta0[UPPER_MAX_JOBS];
    int n;
};
void commit_finish_multi(sezkp_ctx* ctx, Commit* cms, int count, u8* roots_dev);
// Range-split finish of single-column commitments whose level-0 nodes were hashed by chunk range over `world` ranks (a power
// of two dividing n_ch): _range reduces this rank's aligned range — a complete subtree — down to its ONE node at level
// log2(n_ch / world); after the ranks have exchanged those nodes (32 bytes each), _top reduces the log2(world) levels above
// them on every rank and writes root i to roots_dev[i] (device, may be null entries).  The levels below the exchanged one
// exist only for the own range — which is where this rank's openings lie.
void commit_finish_multi_range(sezkp_ctx* ctx, Commit* const* cms, int count, int rank, int world);
void commit_finish_multi_top(sezkp_ctx* ctx, Commit* const* cms, int count, int world, u8* const* roots_dev);
// Leaf hashes + chunk trees of `count` single-column unlabeled commitments (begin done, values final) in one launch.
struct CommitJob {
    const u64* values;
    u32* upper;
    int cl;
};
struct CommitJobs {
    CommitJob j[UPPER_MAX_JOBS];
    u32 cta0[UPPER_MAX_JOBS];
    int n;
};
void commit_chunks_multi(sezkp_ctx* ctx, Commit* cms, int count);
// Generic opening request: one CTA rebuilds the chunk containing `row` of one committed column.
struct OpenReq {
    const u64* values;             // the column
    const u32* upper;              // its retained upper levels
    const b3::LabelTemplate* tpl;  // null = unlabeled
    u64 n_ch, row;
    u32 cl, depth_out;
    u32 out_off, pad;              // path position in the output (32 B units): cl in-chunk siblings then depth_out upper siblings
};
OpenReq make_open_req(const Commit& cm, u32 col, u64 row, u32 out_off);
// Same launch, results left in the context's pinned staging buffer (valid until the next call).
void open_batch_staged(sezkp_ctx* ctx, const std::vector<OpenReq>& reqs, size_t path_digests, u64** values, u8** chunk_roots, u8** paths);
void open_batch(sezkp_ctx* ctx, const std::vector<OpenReq>& reqs, size_t path_digests, u64* values, u8* chunk_roots, u8* paths_host);
// k openings; outputs are host arrays: values[k], chunk_roots[k][32], path_in[k][cl][32], path_to[k][log2(n_ch)][32].
void commit_open(sezkp_ctx* ctx, const Commit& cm, const u32* col_idx, const u64* row_idx, size_t k, u64* values,
                 u8* chunk_roots, u8* path_in, u8* path_to);
// Plain helpers
void leaf_hash_device(sezkp_ctx* ctx, const u64* vals_dev, size_t n, const char* label_or_null, u32* out_dev);
void merkle_root_device(sezkp_ctx* ctx, u32* level_dev /* n digests, destroyed */, u32* tmp_dev, size_t n, u8* root_host);
// Batched Merkle path verification (MerkleTree::verify / verify_chunked_open): see sezkp_verify_openings in the C header.
void verify_paths_device(sezkp_ctx* ctx, const u8* col_roots, const char* const* labels_or_null, int c, const u32* col_idx_or_null,
                         const u64* values, const u64* idx_in, const u64* idx_out, const u8* chunk_roots_or_null, const u8* path_in, int din,
                         const u8* path_to, int dout, size_t k, u8* ok_host);
