#!/usr/bin/env python
"""ncu driver: column commit of the 59 columns of a simulated trace (T=2^log_t) — dedup kernel on real column data."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = m.Context(0)
ct = m.simulate(1 << log_t, 512, 8)
cols = ctx.trace_columns(ct)
dev = torch.from_numpy(cols.view(np.int64)).cuda()
names = ["input_mv", "is_first", "is_last"] + [f"{g}_{r}" for g in ("mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off") for r in range(8)]
for _ in range(2):
    roots = ctx.column_commit(dev, names, dev=True, n=ct.n_rows)
ctx.synchronize()
print("ok", roots[0].tobytes().hex()[:16])
