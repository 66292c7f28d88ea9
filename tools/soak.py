#!/usr/bin/env python
"""Soak: alternate trace shapes / entry points on one context and check that every proof of a shape is byte-identical to
its first one and that device memory does not creep (pool reuse)."""
import importlib, os, sys, hashlib, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
ctx = m.Context(0)
shapes = [(1 << 12, 64, 3), (1 << 16, 512, 8), (1 << 20, 512, 2), (1 << 18, 256, 1), (1 << 21, 512, 8)]
traces = []
for T, b, tau in shapes:
    ct = m.simulate(T, b, tau, seed=T % 97)
    traces.append((ct, m.manifest_root(ct)))
first = {}
mem = []
t0 = time.time()
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for it in range(rounds):
    for k, (ct, root) in enumerate(traces):
        mode = (it + k) % 3
        if mode == 1:
            ct.pack_ops()
        else:
            ct.ops = None
        if mode == 2:
            rt = ctx.upload_trace(ct)
            p = ctx.prove_v1_resident(rt, root)
            rt.free()
        else:
            p = ctx.prove_v1(ct, root)
        h = hashlib.sha256(p).hexdigest()
        assert first.setdefault(k, h) == h, f"proof of shape {k} changed at iteration {it} (mode {mode})"
    free, total = torch.cuda.mem_get_info()
    mem.append((total - free) / 1e9)
print(f"soak ok: {rounds * len(traces)} proofs in {time.time() - t0:.1f} s, device memory used GB first/max/last: {mem[0]:.2f} / {max(mem):.2f} / {mem[-1]:.2f}")
