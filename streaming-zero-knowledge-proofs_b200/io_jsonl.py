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
    """CompactTrace -> JSONL of BlockSummary objects (field names and order of crates/sezkp-core/src/types.rs:116-151,
    compact separators like serde_json::to_string)."""
    tau = ct.tau
    row = 0
    tags = json.dumps([[0] * 16 for _ in range(tau)], separators=(",", ":"))
    ops = {}  # (flag, sym, mv) -> '{"write":..,"mv":..}'

    def op(f, s, v):
        k = (f, s, v)
        t = ops.get(k)
        if t is None:
            t = ops[k] = '{"write":%s,"mv":%d}' % (str(s) if f else "null", v)
        return t

    opt = lambda name, k, dflt: int(getattr(ct, name)[k]) if getattr(ct, name) is not None else dflt
    with open(path, "w") as f:
        for k in range(ct.n_blocks):
            n = int(ct.block_len[k])
            fl, sy, mv, im = (ct.write_flag[row:row + n].tolist(), ct.write_sym[row:row + n].tolist(), ct.mv[row:row + n].tolist(),
                              ct.input_mv[row:row + n].tolist())
            steps = ",".join('{"input_mv":%d,"tapes":[%s]}' % (im[j], ",".join(op(fl[j][r], sy[j][r], mv[j][r]) for r in range(tau)))
                             for j in range(n))
            head = {
                "version": opt("version", k, 1), "block_id": opt("block_id", k, k + 1), "step_lo": opt("step_lo", k, row + 1),
                "step_hi": opt("step_hi", k, row + n), "ctrl_in": opt("ctrl_in", k, 0), "ctrl_out": opt("ctrl_out", k, 0),
                "in_head_in": opt("in_head_in", k, 0), "in_head_out": opt("in_head_out", k, 0),
                "windows": [{"left": int(ct.win_left[k, r]), "right": int(ct.win_right[k, r])} for r in range(tau)],
                "head_in_offsets": [int(x) for x in ct.head_in_off[k]],
                "head_out_offsets": [int(x) for x in ct.head_out_off[k]],
            }
            f.write(json.dumps(head, separators=(",", ":"))[:-1] + ',"movement_log":{"steps":[' + steps + ']},"pre_tags":' + tags +
                    ',"post_tags":' + tags + "}\n")
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
