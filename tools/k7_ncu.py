#!/usr/bin/env python
"""one LDE + commit of 2 columns x 2^24 rows for ncu (LDE_FUSE=0/1 selects the mode)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
import numpy as np
ctx = m.Context(0)
ctx.set_option("lde_fuse", int(os.environ.get("LDE_FUSE", "0")))
import torch
ev = torch.randint(0, 1 << 62, (2, 1 << 24), dtype=torch.int64, device="cuda")
for _ in range(2):
    r = ctx.lde_commit(ev, ["c_0", "c_1"], 3, dev=True, log_n=24)
print(r[0].tobytes().hex()[:16])
