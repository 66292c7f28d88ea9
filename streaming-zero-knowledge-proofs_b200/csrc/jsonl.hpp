// Native JSONL front-end (jsonl.cpp): serde-JSON BlockSummary lines -> the flat arrays of sezkp_trace_desc.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/sezkp_trace.h"

namespace jsonl {

typedef sezkp_block_scalars Manifest;  // per-block scalars only the manifest leaf hash reads

struct Trace {
    uint32_t tau = 0;
    std::vector<uint64_t> block_len;
    std::vector<int64_t> win_left, win_right;
    std::vector<uint32_t> head_in_off, head_out_off;
    std::vector<int8_t> input_mv, mv;
    std::vector<uint8_t> write_flag;
    std::vector<uint16_t> write_sym;
    std::vector<Manifest> manifest;
    size_t n_lines = 0;  // newline-terminated (or final unterminated) lines seen, blank ones included
    void fill_desc(sezkp_trace_desc& d) const;
    void clear_keep_capacity();  // empty, but the arrays keep their memory (pieces of one file reuse them: no fresh pages per piece)
};

// Persistent parser threads for callers that parse many pieces of one file (prove_jsonl_file: a 13 GB file is ~200 pieces;
// spawning 32 threads per piece cost more than the hand-over to the GPU).  run() executes fn(0..tasks-1) on the pool's threads
// and the calling thread and returns when all are done; fn must not throw.  One run() at a time.
class WorkerPool {
public:
    explicit WorkerPool(int n_threads);
    ~WorkerPool();
    WorkerPool(const WorkerPool&) = delete;
    WorkerPool& operator=(const WorkerPool&) = delete;
    void run(int tasks, const std::function<void(int)>& fn);
    int threads() const { return n_; }

private:
    struct Impl;
    Impl* p_;
    int n_;
};

// Parse every non-blank line of text[0, len) on up to n_threads host threads; throws std::runtime_error
// ("jsonl line N: ...") on malformed input.  tau_hint = 0 takes tau from the first block.
void parse(const char* text, size_t len, int n_threads, uint32_t tau_hint, size_t first_line_no, Trace& out);
// Same, but the workers' outputs are returned separately in file order (no concatenation pass); returns the number
// of lines seen, tau_out = the common tau (tau_hint when no block was found).
size_t parse_parts(const char* text, size_t len, int n_threads, uint32_t tau_hint, size_t first_line_no, std::vector<Trace>& parts,
                   uint32_t& tau_out, WorkerPool* pool = nullptr);

// Writer (inverse of the parser): one serde-JSON BlockSummary per line, formatted on n_threads host threads; `scalars` may be
// null (version 1, block ids from 1, step ranges from the block lengths).  Returns the bytes written.
size_t write_file(const char* path, const sezkp_trace_desc& d, const sezkp_block_scalars* scalars, int n_threads);

}  // namespace jsonl
