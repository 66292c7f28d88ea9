// Internal interface of wide.cu (config 4: LDE + labeled column commit + FRI over column-sharded wide column sets).
#pragma once
#include <chrono>
#include <vector>

#include "common.cuh"

struct WideColumns {  // this rank's share of a resident column set: column c of the set lives on rank c % world
    DevBuf evals;     // device [n_local][1 << log_n], local column j = global column rank + j * world
    int c = 0, log_n = 0, rank = 0, world = 1, n_local = 0;
    void release();
};
struct WideTaps {  // intermediate values for per-stage parity tests
    std::vector<u64> alphas;
    u64 z = 0;
};
inline double wide_now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
void wide_columns_synth(sezkp_ctx* ctx, WideColumns& wc, u64 seed, int c, int log_n, int rank, int world);
void wide_columns_upload(sezkp_ctx* ctx, WideColumns& wc, const u64* evals_host, int c, int log_n, int rank, int world);
// LDE + labeled commit of c resident columns (dense [c][n]); roots to d_roots (device [c][32]); no synchronisation.
void lde_commit_columns(sezkp_ctx* ctx, const u64* evals_dev, const char* const* labels, int c, int log_n, int log_blow, u64 shift,
                        int chunk_log2, u8* d_roots);
// The whole config-4 pipeline for one rank of a group (or a single GPU: world == 1).  labels: all c labels.  Host outputs
// (any may be null): col_roots_out [c][32], fri_roots_out [log_n+log_blow+1][32], final value.  Every rank computes the
// same outputs.
void wide_commit_fri_rank(sezkp_ctx* ctx, const WideColumns& wc, const char* const* labels, int log_blow, u64 shift, int chunk_log2,
                          u8* col_roots_out, u8* fri_roots_out, u64* final_value, WideTaps* taps);
