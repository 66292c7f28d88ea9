#!/usr/bin/env python
"""e2e (host pinned input) and resident prove latency at T=2^22 with the two phase clocks (option phase_sync)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
b = importlib.import_module("streaming-zero-knowledge-proofs_b200.binding")
import bench
ct = bench.pin_trace(torch, m.simulate(1 << 22, 512, 8, seed=42))
root = m.manifest_root(ct)
buf = torch.empty(b.proof_size_bound(ct.n_rows, ct.tau), dtype=torch.uint8, pin_memory=True).numpy()
ctx = m.Context(0)
ctx.set_option("tab_cache", 0)
rt = ctx.upload_trace(ct)
for rnd in range(2):
    for ps in (1, 0):
        ctx.set_option("phase_sync", ps)
        for _ in range(3):
            ctx.prove_v1(ct, root, buf)
        t0 = time.perf_counter()
        for _ in range(20):
            ctx.prove_v1(ct, root, buf, view=True)
        e2e = (time.perf_counter() - t0) / 20 * 1e3
        ph = ctx.timings()
        for _ in range(3):
            ctx.prove_v1_resident(rt, root, buf)
        t0 = time.perf_counter()
        for _ in range(20):
            ctx.prove_v1_resident(rt, root, buf, view=True)
        res = (time.perf_counter() - t0) / 20 * 1e3
        print(f"phase_sync={ps}: e2e {e2e:.3f} ms  resident {res:.3f} ms  e2e phases {ph}", flush=True)
