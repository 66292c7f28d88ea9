#!/usr/bin/env python
"""BASELINE configs[3] at full size on one GPU: 256 columns x 2^24 rows, blow-up 8 — iNTT + coset LDE + labeled BLAKE3
column commit per column group (sezkp_lde_commit_batch_dev).  usage: python tools/wide_check.py [cols=256] [log_n=24]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
k = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ctx = m.Context(0)
n = 1 << k
g = torch.Generator(device="cuda"); g.manual_seed(0x5EED)
ev = torch.randint(0, (1 << 62), (cols, n), dtype=torch.int64, device="cuda", generator=g)
labels = [f"c_{c}" for c in range(cols)]
r0 = ctx.lde_commit(ev, labels, 3, 3, dev=True, log_n=k)
torch.cuda.synchronize()
t0 = time.perf_counter()
r1 = ctx.lde_commit(ev, labels, 3, 3, dev=True, log_n=k)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
assert (r0 == r1).all()
free, total = torch.cuda.mem_get_info()
alg = cols * (8 * n * 9 + 8 * n * 8)
print(json.dumps({"workload": f"{cols} columns x 2^{k} rows, blow-up 8: iNTT + coset LDE + labeled commit, 1 GPU", "seconds_per_step": dt,
                  "rows_per_s": n / dt, "algorithmic_GB": alg / 1e9, "GBps": alg / dt / 1e9, "leaf_compressions_per_s": cols * (2 * (n << 3) - 1) / dt,
                  "device_mem_used_GB": (total - free) / 1e9, "root0": r0[0].tobytes().hex()}))
