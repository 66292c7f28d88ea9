#!/usr/bin/env python
"""ncu driver: column commit of ONE column group of a simulated trace: python tools/profile_commit_group.py <group> [log_t]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
group = sys.argv[1] if len(sys.argv) > 1 else "winlen"
log_t = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = m.Context(0)
ct = m.simulate(1 << log_t, 512, 8)
cols = ctx.trace_columns(ct)
gi = ["mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off"].index(group)
sub = torch.from_numpy(cols[3 + 8 * gi: 3 + 8 * gi + 8].copy().view(np.int64)).cuda()
names = [f"{group}_{r}" for r in range(8)]
for _ in range(2):
    roots = ctx.column_commit(sub, names, dev=True, n=ct.n_rows)
ctx.synchronize()
print("ok", roots[0].tobytes().hex()[:16])
