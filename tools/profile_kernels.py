#!/usr/bin/env python
"""Small single-GPU driver for ncu captures: one column commit (59 labeled columns) and one coset LDE.

usage: python tools/profile_kernels.py [log_rows=20] [lde_cols=16]
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lde_cols = int(sys.argv[2]) if len(sys.argv) > 2 else 16
P = 0xFFFFFFFF00000001
ctx = m.Context(0)
rng = np.random.default_rng(1)
n = 1 << log_rows
tau = 8
labels = ["input_mv", "is_first", "is_last"] + [f"{g}_{r}" for g in ("mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off") for r in range(tau)]
cols = torch.from_numpy((rng.integers(0, 1 << 63, size=(len(labels), n), dtype=np.uint64) % np.uint64(P)).view(np.int64)).cuda()
for _ in range(2):
    roots = ctx.column_commit(cols, labels, dev=True, n=n)
k, lb = 20, 2
d = torch.from_numpy((rng.integers(0, 1 << 63, size=(lde_cols, 1 << k), dtype=np.uint64) % np.uint64(P)).view(np.int64)).cuda()
out = torch.empty((lde_cols, (1 << k) << lb), dtype=torch.int64, device="cuda")
for _ in range(2):
    ctx.coset_lde_dev(d, k, lb, 3, lde_cols, out)
    ctx.ntt_dev(d, k, lde_cols, False)
    ctx.ntt_dev(d, k, lde_cols, True)
l0 = torch.from_numpy((rng.integers(0, 1 << 63, size=1 << 22, dtype=np.uint64) % np.uint64(P)).view(np.int64)).cuda()
betas = rng.integers(0, 1 << 63, size=22, dtype=np.uint64) % np.uint64(P)
ctx.fri_commit(l0, betas, dev=True, log_N=22)
ctx.synchronize()
print("profile run ok", roots[0].tobytes().hex()[:16], ctx.launch_count())
