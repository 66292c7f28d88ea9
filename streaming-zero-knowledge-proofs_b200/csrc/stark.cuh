// Internal interface of stark.cu.
#pragma once
#include <cstring>
#include <functional>
#include <vector>

#include "common.cuh"
#include "hash.cuh"
#include "gl.cuh"
#include "transcript.hpp"

struct DeviceTrace {  // device image of sezkp_trace_desc (+ block_start prefix sums)
    u32 tau;
    u64 n_blocks, n_rows;
    const u64* block_start;
    const u64* block_len;
    const int64_t* win_left;
    const int64_t* win_right;
    const u32* head_in_off;
    const u32* head_out_off;
    const int8_t* input_mv;
    const int8_t* mv;
    const u8* write_flag;
    const uint16_t* write_sym;
    const u8* ops;  // non-null: packed per-tape ops (SEZKP_TRACE_PACKED_OPS), mv / write_flag / write_sym are null
};
struct DeviceTraceOwner {
    DevBuf buf;
    DeviceTrace t{};
    size_t h2d_bytes = 0;
    bool packed = false;
    size_t o_start = 0, o_len = 0, o_wl = 0, o_wr = 0, o_io = 0, o_oo = 0, o_imv = 0, o_mv = 0, o_wf = 0, o_ws = 0, total = 0;
    void layout(const sezkp_trace_desc* d);
    void upload_meta(sezkp_ctx* ctx, const sezkp_trace_desc* d);                                    // allocation + per-block metadata
    void upload_rows_async(cudaStream_t stream, const sezkp_trace_desc* d, u64 row0, u64 row1);     // a slab of the row arrays
    void upload(sezkp_ctx* ctx, const sezkp_trace_desc* d);                                         // everything, synchronous
};
struct SlabPlan {  // pipelined upload: per slab the rows, the blocks that end in it, and the event of its copy
    struct Slab {
        u64 row0, row1, blk0, blk1, complete_rows;
        cudaEvent_t ready;
    };
    std::vector<Slab> slabs;
};
void validate_trace(const sezkp_trace_desc* d);
void expand_columns_device(sezkp_ctx* ctx, const DeviceTrace& t, u64* cols_dev);
void compose_device(sezkp_ctx* ctx, const u64* cols_dev, u64 n, u32 tau, const u64 alphas8[8], const u64* mask, size_t mask_deg,
                    u64* out_dev, u64 row0 = 0, u64 row1 = ~0ULL);  // rows [row0, row1) only (row i reads rows i and i+1 mod n)
bool z_on_coset(u64 z, u64 shift, int log_N);
void deep_lde_device(sezkp_ctx* ctx, u64* base_vals_dev /* destroyed */, u64* out_dev, int L, int logB, u64 shift, u64 z);
void deep_quotient_device(sezkp_ctx* ctx, u64* y_dev, int log_dom, u64 shift, u64 z);
void deep_lde_sharded_device(sezkp_ctx* ctx, u64* base_vals_dev /* destroyed */, u64* out_dev, int L, int logB, u64 shift, u64 z, int rank,
                             int world);
u64* deep_lde_coset_local_device(sezkp_ctx* ctx, u64* base_vals_dev /* destroyed */, int L, int logB, u64 shift, u64 z, int rank, int world);

// Where the serialised proof goes: straight into the caller's buffer (no intermediate copy).  Bytes beyond the
// capacity are counted but not written, so a NULL / short buffer still yields the required length.
struct ProofSink {
    u8* buf = nullptr;
    size_t cap = 0, len = 0;
    ProofSink(u8* b, size_t c) : buf(b), cap(b ? c : 0) {}
    void put(const void* p, size_t n) {
        if (len + n <= cap) std::memcpy(buf + len, p, n);
        len += n;
    }
};

struct HostAbsorb {  // transcript hooks of the FRI commit loop
    virtual void on_root(int layer, const u8* root) = 0;
    virtual std::vector<u64> draw_betas(int n) = 0;
    virtual ~HostAbsorb() {}
};
// Transcript hooks of the FRI loop as prove_v1 uses them (v1/prover.rs:187, 219, 235; v1/params.rs:103-113).
inline u64 le64(const u8* p) {
    u64 v;
    std::memcpy(&v, p, 8);
    return v;
}
struct TranscriptAbsorb : HostAbsorb {
    host::Transcript& tr;
    explicit TranscriptAbsorb(host::Transcript& t) : tr(t) {}
    void on_root(int, const u8* root) override { tr.absorb("fri_layer_root", root, 32); }
    std::vector<u64> draw_betas(int n) override {
        auto by = tr.challenge("fri_betas", 8 * (size_t)n);
        std::vector<u64> out(n);
        for (int i = 0; i < n; i++) out[i] = le64(&by[8 * i]) % gl::P;
        return out;
    }
};
struct FriLayers {
    int log_N = 0;
    u64* values = nullptr;        // device: layer l (len N>>l) at offset sum_{i<l} N>>i
    std::vector<Commit> commits;  // one per layer
    void release(sezkp_ctx* ctx);
};
struct ShardInfo;
// shard != null: the hashing of the large layers is split by chunk range over the ranks (see stark.cu)
// coset_local0 != null (context group with peer access, world | 8): coset-resident layers — layer 0 arrives as this rank's
// cosets (deep_lde_coset_local_device), folding is local, only the own index range of every large layer is materialised in
// natural order (fl.values holds that range at its natural position; the rest of a large layer is never written).
void fri_commit_device(sezkp_ctx* ctx, FriLayers& fl, const u64* layer0_dev, int log_N, const u64* betas_or_null, u8* roots_host,
                       u64* final_value, HostAbsorb* absorb_or_null, const ShardInfo* shard = nullptr, const u64* coset_local0 = nullptr);
// Rank that holds (and opens) index `row` of layer l under the chunk-range split of the large layers; -1: the layer is small
// (replicated on every rank).
int fri_range_owner(int log_N, int l, u64 row, int world);
// keep_rank >= 0: only the requests this rank serves are appended (large layers: the range owner; small layers: rank 0);
// req_index[(q*log_N + l)*2 + s] = index of the request in `reqs` or -1.
void fri_open_requests(const FriLayers& fl, const u64* idx0, size_t k, u64* positions, std::vector<OpenReq>& reqs, u32 base_off,
                       int keep_rank = -1, int world = 1, std::vector<int32_t>* req_index = nullptr);
void fri_open_device(sezkp_ctx* ctx, const FriLayers& fl, const u64* idx0, size_t k, u64* positions, u64* values, u8* paths);
struct ShardInfo {  // column sharding across the GPUs of one box
    int rank, world;
    sezkp_allgather_fn allgather;
    void* user;
    // Optional (context groups: the ranks are threads of one process and only rank 0 delivers the proof): publish `send` and
    // rendezvous once; rank 0 then reads the peers' buffers in place through all_ptrs[world] — no copies, no second barrier.
    // `send` must stay valid until the whole group call has returned (context-owned staging).
    int32_t (*gather_root)(void* user, const void* send, const void** all_ptrs) = nullptr;
};
void prove_v1_resident(sezkp_ctx* ctx, const DeviceTrace& trace, const u8 manifest_root[32], ProofSink& proof_out,
                       const ShardInfo* shard = nullptr, const SlabPlan* plan = nullptr);
struct ExpandFilter {  // which (column, row) cells an expansion produces; the default produces everything
    u32 col_mod = 0, col_rem = 0;           // col_mod > 1: columns c % col_mod == col_rem for every row ...
    u64 full_lo = 0, full_hi = 0, halo = ~0ULL;  // ... plus all columns for rows [full_lo, full_hi) and row `halo`
};
// rows [row0,row1) of the non-head columns and the head columns of blocks [blk0,blk1)
void expand_columns_range(sezkp_ctx* ctx, const DeviceTrace& t, u64* cols_dev, u64 row0, u64 row1, u64 blk0, u64 blk1,
                          const ExpandFilter& f = ExpandFilter());
void prove_v1_device(sezkp_ctx* ctx, const sezkp_trace_desc* desc, const u8 manifest_root[32], ProofSink& proof_out,
                     const ShardInfo* shard = nullptr);

// stream.cu
struct sezkp_stream;
sezkp_stream* stream_begin(sezkp_ctx* ctx, u32 tau, const u8 manifest_root[32], u64 expected_rows);
void stream_ingest(sezkp_ctx* ctx, sezkp_stream* st, const sezkp_trace_desc* blocks);
// run fn(0..tasks-1) on several host threads and return when all are done (fn does not throw)
typedef std::function<void(int tasks, const std::function<void(int)>& fn)> HostParallelFor;
void stream_ingest_parts(sezkp_ctx* ctx, sezkp_stream* st, const sezkp_trace_desc* descs, size_t count, const HostParallelFor* par);
void stream_finish(sezkp_ctx* ctx, sezkp_stream* st, ProofSink& proof);
void stream_free(sezkp_ctx* ctx, sezkp_stream* st);
u32 stream_tau(const sezkp_stream* st);
