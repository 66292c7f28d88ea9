"""JSONL block streams: one serde-shaped ``BlockSummary`` per line.

Reference: ``stream_block_summaries_jsonl`` (crates/sezkp-core/src/io_jsonl.rs:27-88) and the CLI's
``export-jsonl`` (crates/sezkp-cli/src/main.rs).  The reference's stark arm rejects ``.jsonl`` input
(core/io.rs:78-88); here the stream feeds ``Context.prove_v1_stream`` (``ProvingBackendStream``), so the trace is
never materialised on the host: each parsed piece is packed into pinned staging memory and copied to the GPU while the
next lines are being parsed.
"""
from __future__ import annotations

import json
from typing import Iterator

import numpy as np

from .trace import CompactTrace, blocks_to_compact


def write_jsonl(path: str, ct: CompactTrace) -> None:
    """CompactTrace -> JSONL of BlockSummary objects (field names and order of crates/sezkp-core/src/types.rs:116-151)."""
    tau = ct.tau
    row = 0
    with open(path, "w") as f:
        for k in range(ct.n_blocks):
            n = int(ct.block_len[k])
            steps = []
            for j in range(row, row + n):
                tapes = [{"write": (int(ct.write_sym[j, r]) if ct.write_flag[j, r] else None), "mv": int(ct.mv[j, r])} for r in range(tau)]
                steps.append({"input_mv": int(ct.input_mv[j]), "tapes": tapes})
            blk = {
                "version": int(ct.version[k]) if ct.version is not None else 1,
                "block_id": int(ct.block_id[k]) if ct.block_id is not None else k + 1,
                "step_lo": int(ct.step_lo[k]) if ct.step_lo is not None else row + 1,
                "step_hi": int(ct.step_hi[k]) if ct.step_hi is not None else row + n,
                "ctrl_in": int(ct.ctrl_in[k]) if ct.ctrl_in is not None else 0,
                "ctrl_out": int(ct.ctrl_out[k]) if ct.ctrl_out is not None else 0,
                "in_head_in": int(ct.in_head_in[k]) if ct.in_head_in is not None else 0,
                "in_head_out": int(ct.in_head_out[k]) if ct.in_head_out is not None else 0,
                "windows": [{"left": int(ct.win_left[k, r]), "right": int(ct.win_right[k, r])} for r in range(tau)],
                "head_in_offsets": [int(x) for x in ct.head_in_off[k]],
                "head_out_offsets": [int(x) for x in ct.head_out_off[k]],
                "movement_log": {"steps": steps},
                "pre_tags": [[0] * 16 for _ in range(tau)],
                "post_tags": [[0] * 16 for _ in range(tau)],
            }
            f.write(json.dumps(blk, separators=(",", ":")) + "\n")
            row += n


def stream_jsonl(path: str, blocks_per_piece: int = 64) -> Iterator[CompactTrace]:
    """Yield CompactTrace pieces of up to `blocks_per_piece` consecutive blocks; blank lines are skipped like the reference."""
    batch = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            batch.append(json.loads(line))
            if len(batch) == blocks_per_piece:
                yield blocks_to_compact(batch)
                batch = []
    if batch:
        yield blocks_to_compact(batch)
