/* Compact trace descriptor shared by the C ABI (include/sezkp_cuda.h) and the test oracle.
 *
 * It is the flat, pointer-and-size restatement of the reference's `&[BlockSummary]` argument of
 * `ProvingBackend::prove` (crates/sezkp-core/src/backend.rs:41-61; struct at
 * crates/sezkp-core/src/types.rs:116-151), holding exactly the fields `prove_v1` reads
 * (crates/sezkp-stark/src/v1/columns.rs:252-365, v1/openings.rs:182-273): per block the row count
 * and, per tape, window bounds and entry/exit offsets; per row the input move and, per tape, the
 * head move and optional write.  ≈ 1 + 4·tau bytes per row instead of the reference's in-memory
 * BlockSummary tree — or 1 + tau bytes per row with SEZKP_TRACE_PACKED_OPS, where the three per-tape arrays are one
 * byte per (row, tape): bits 0-1 = mv + 1, bit 2 = write.is_some(), bits 3-7 = the written symbol (alphabets of up to
 * 32 symbols; the host-to-device copy is the exposed part of an end-to-end prove, so a binding whose symbols fit should
 * flatten into this form).  All arrays are caller-owned host memory.
 */
#ifndef SEZKP_TRACE_H
#define SEZKP_TRACE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sezkp_trace_desc {
    uint32_t tau;                 /* tapes per step (blocks[0].windows.len())                    */
    uint32_t flags;               /* 0 or SEZKP_TRACE_PACKED_OPS                                 */
    uint64_t n_blocks;            /* number of BlockSummary records                              */
    uint64_t n_rows;              /* Σ_k (step_hi − step_lo + 1); must be a power of two         */
    const uint64_t* block_len;    /* [n_blocks]       step_hi − step_lo + 1 (= movement_log len) */
    const int64_t*  win_left;     /* [n_blocks][tau]  windows[r].left                            */
    const int64_t*  win_right;    /* [n_blocks][tau]  windows[r].right                           */
    const uint32_t* head_in_off;  /* [n_blocks][tau]  head_in_offsets[r]                         */
    const uint32_t* head_out_off; /* [n_blocks][tau]  head_out_offsets[r]                        */
    const int8_t*   input_mv;     /* [n_rows]         steps[j].input_mv                          */
    const int8_t*   mv;           /* [n_rows][tau]    steps[j].tapes[r].mv; PACKED_OPS: the packed op bytes */
    const uint8_t*  write_flag;   /* [n_rows][tau]    steps[j].tapes[r].write.is_some()  (ignored when packed) */
    const uint16_t* write_sym;    /* [n_rows][tau]    steps[j].tapes[r].write.unwrap_or(0) (ignored when packed) */
} sezkp_trace_desc;
#define SEZKP_TRACE_PACKED_OPS 1u  /* flags bit 0: `mv` holds (mv + 1) | write.is_some() << 2 | symbol << 3 per (row, tape) */

/* The per-block scalars that only the manifest leaf hash reads (crates/sezkp-merkle/src/lib.rs:85-117); the native
 * JSONL parser returns them next to the descriptor so a caller can rebuild `manifest_root` from a .jsonl file. */
typedef struct sezkp_block_scalars {
    uint64_t step_lo, step_hi;
    int64_t  in_head_in, in_head_out;
    uint32_t block_id;
    uint16_t version, ctrl_in, ctrl_out;
    uint16_t reserved;            /* 0 */
    uint32_t reserved2;           /* 0; sizeof == 48 */
} sezkp_block_scalars;

#ifdef __cplusplus
}
#endif
#endif /* SEZKP_TRACE_H */
