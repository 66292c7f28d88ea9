#!/usr/bin/env python
"""ncu driver: 64 x 2^20 forward NTT only (two launches of ntt_pass_kernel<5,5> per call), 1 warm call + 1 profiled call."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
ctx = m.Context(0)
P = 0xFFFFFFFF00000001
cols, k = 64, 20
rng = np.random.default_rng(1)
d = torch.from_numpy((rng.integers(0, 1 << 63, size=(cols, 1 << k), dtype=np.uint64) % np.uint64(P)).view(np.int64)).cuda()
for _ in range(2):
    ctx.ntt_dev(d, k, cols, False)
ctx.synchronize()
print("ok")
