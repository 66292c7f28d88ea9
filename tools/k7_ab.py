#!/usr/bin/env python
"""K7 A/B: LDE + labeled commit of W_COLS columns x 2^W_LOG_N rows with the leaf hash fused into the LDE's last pass
(option lde_fuse = 1) and as a separate kernel (0).  Same roots; prints the step time of both."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
cols, k = int(os.environ.get("W_COLS", 16)), int(os.environ.get("W_LOG_N", 24))
ctx = m.Context(0)
cs = ctx.columns_synth(cols, k)
res = {}
for fuse in (0, 1, 0, 1):
    ctx.set_option("lde_fuse", fuse)
    r = ctx.lde_commit_fri(cs)
    t0 = time.perf_counter()
    for _ in range(2):
        r = ctx.lde_commit_fri(cs)
    dt = (time.perf_counter() - t0) / 2
    tm = ctx.timings()
    res.setdefault(fuse, []).append((dt * 1e3, tm["lde_commit"], r[0][0].tobytes().hex()[:16]))
    print(f"lde_fuse={fuse}: {dt * 1e3:.1f} ms per step, lde_commit phase {tm['lde_commit']:.1f} ms ({tm['lde_commit'] / cols:.2f} ms per column), root0 {r[0][0].tobytes().hex()[:16]}", flush=True)
assert res[0][0][2] == res[1][0][2], "fused and unfused roots differ"
print("same roots")
