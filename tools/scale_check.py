#!/usr/bin/env python
"""BASELINE configs[4] scale check on one GPU: streaming STARK v1 prove of a T=2^log_t simulated trace that is generated
slab by slab (never resident on the host) and pushed through begin/ingest/finish (pinned staging + side-stream copies).
Checks: the proof verifies under the CPU oracle's verify_v1 outcome rules (Merkle paths + FRI consistency), and reports
rows/s, the copy overlap and device memory.  usage: python tools/scale_check.py 26"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
tr = importlib.import_module("streaming-zero-knowledge-proofs_b200.trace")
log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T, b, tau, SLAB = 1 << log_t, 512, 8, 1 << 20
ctx = m.Context(0)

def slabs():
    per = 1 + 2 * tau
    for lo in range(0, T, SLAB):
        hi = min(T, lo + SLAB)
        r = tr._splitmix_block(42, lo * per, (hi - lo) * per).reshape(hi - lo, per)
        input_mv = (r[:, 0] % np.uint64(3)).astype(np.int8) - 1
        w, mm = r[:, 1::2], r[:, 2::2]
        f = (w % np.uint64(10)) < np.uint64(4)
        wsym = np.where(f, (w >> np.uint64(32)) % np.uint64(16), 0).astype(np.uint16)
        mv = (mm % np.uint64(3)).astype(np.int8) - 1
        yield tr.partition(input_mv, mv, f.astype(np.uint8), wsym, b)   # blocks never straddle a slab (2^20 % 512 == 0)

root = bytes(range(32))
cold = ctx.prove_v1_stream(slabs(), root, tau=tau, expected_rows=T)   # cold call: allocations, twiddle tables
t0 = time.perf_counter()
proof = ctx.prove_v1_stream(slabs(), root, tau=tau, expected_rows=T)
dt = time.perf_counter() - t0
assert proof == cold
tm = ctx.timings()
free, total = torch.cuda.mem_get_info()
out = {"log_T": log_t, "proof_bytes": len(proof), "wall_s_incl_host_generation": dt, "rows_per_s_incl_host_generation": T / dt,
       "gpu_prove_ms": sum(v for k, v in tm.items() if not k.startswith("stream_")), "timings": tm,
       "device_mem_used_GB": (total - free) / 1e9}
import oracle_lib
orc = oracle_lib.load()
ok, why = orc.verify_v1(proof, m.demo_block(16) if False else next(iter(slabs())))
out["oracle_verify"] = {"accepted": ok, "reason": why}
print(json.dumps(out))
