#!/usr/bin/env python
"""ncu driver for one resident STARK v1 prove step at the bench size (T=2^22, b=512, tau=8): two warm proofs, then one
between cudaProfilerStart/Stop (run ncu with --profile-from-start off)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = m.Context(0)
ct = m.simulate(1 << log_t, 512, 8)
ct.pack_ops()
root = m.manifest_root(ct)
rt = ctx.upload_trace(ct)
for _ in range(2):
    p = ctx.prove_v1_resident(rt, root)
torch.cuda.synchronize()
torch.cuda.profiler.start()
q = ctx.prove_v1_resident(rt, root)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
assert p == q
print("ok", len(q), ctx.timings())
