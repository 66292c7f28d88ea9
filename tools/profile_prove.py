#!/usr/bin/env python
"""ncu driver: two resident proofs of the bench workload (T = 2^log_t, b = 512, tau = 8); profile the second one.

  ncu --set full --clock-control none --import-source on -k regex:'chunk_commit|compose|deep|ntt_pass|expand|head_scan|open_kernel' \
      --launch-skip 43 -c 18 -o gpurun_out/prof python tools/profile_prove.py 22
"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = m.Context(0)
ct = m.simulate(1 << log_t, 512, 8)
root = m.manifest_root(ct)
rt = ctx.upload_trace(ct)
for _ in range(2):
    proof = ctx.prove_v1_resident(rt, root)
ctx.synchronize()
print("ok", len(proof), ctx.timings())
