#!/usr/bin/env python
"""One T=2^22 proof through a context group over all visible GPUs: e2e and resident latency + phase breakdown."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
b = importlib.import_module("streaming-zero-knowledge-proofs_b200.binding")
import bench
devices = list(range(torch.cuda.device_count()))
T = 1 << int(os.environ.get("LOG_T", 22))
ct = bench.pin_trace(torch, m.simulate(T, 512, 8, seed=42))
root = m.manifest_root(ct)
buf = torch.empty(b.proof_size_bound(ct.n_rows, ct.tau), dtype=torch.uint8, pin_memory=True).numpy()
one = m.Context(0)
want = one.prove_v1(ct, root)
one.close()
for coset in (1, 0, 1):  # A/B/A: coset-resident FRI layers vs all-gathered layer 0 + replicated folds
    out = bench.group_prove_bench(torch, m, devices, ct, root, buf, want, 10, {"fri_coset": coset})
    out["fri_coset"] = coset
    out.pop("resident_phases_ms_per_gpu", None)
    print(json.dumps(out), flush=True)
