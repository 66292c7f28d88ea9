// Native JSONL front-end (jsonl.cpp): serde-JSON BlockSummary lines -> the flat arrays of sezkp_trace_desc.
#pragma once
#include <cstddef>
#include <cstdint>
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
};

// Parse every non-blank line of text[0, len) on up to n_threads host threads; throws std::runtime_error
// ("jsonl line N: ...") on malformed input.  tau_hint = 0 takes tau from the first block.
void parse(const char* text, size_t len, int n_threads, uint32_t tau_hint, size_t first_line_no, Trace& out);
// Same, but the workers' outputs are returned separately in file order (no concatenation pass); returns the number
// of lines seen, tau_out = the common tau (tau_hint when no block was found).
size_t parse_parts(const char* text, size_t len, int n_threads, uint32_t tau_hint, size_t first_line_no, std::vector<Trace>& parts,
                   uint32_t& tau_out);

// Writer (inverse of the parser): one serde-JSON BlockSummary per line, formatted on n_threads host threads; `scalars` may be
// null (version 1, block ids from 1, step ranges from the block lengths).  Returns the bytes written.
size_t write_file(const char* path, const sezkp_trace_desc& d, const sezkp_block_scalars* scalars, int n_threads);

}  // namespace jsonl
