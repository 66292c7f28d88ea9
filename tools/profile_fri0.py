#!/usr/bin/env python
"""ncu driver for bench.py's roofline kernel: chunk_commit_kernel<FOLD=0> over one FRI-layer-0-sized vector
(N = 2^25 unlabeled leaves -> leaf hashes -> 1024-leaf chunk trees), one warm call + one profiled call."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
ctx = m.Context(0)
N = 1 << 25
g = torch.Generator(device="cuda")
g.manual_seed(7)
layer0 = torch.randint(0, (1 << 62), (N,), dtype=torch.int64, device="cuda", generator=g)
ctx.set_option("dedup", 0)
for _ in range(2):
    ctx.column_commit(layer0, None, dev=True, n=N, c=1)
ctx.synchronize()
print("ok")
