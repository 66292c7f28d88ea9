#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer: NTT/LDE, commits (plain + value-aware), FRI, streaming and prove on tiny inputs."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
P = 0xFFFFFFFF00000001
ctx = m.Context(0)
rng = np.random.default_rng(3)
f = lambda shape: rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) % np.uint64(P)
for k in (3, 10, 12, 21):
    v = f((2, 1 << k)) if k < 20 else f((1, 1 << k))
    assert np.array_equal(ctx.ntt(ctx.ntt(v), inverse=True), v)
ctx.coset_lde(f((3, 1 << 11)), 3, 3); ctx.lde_from_evals(f((2, 1 << 12)), 2, 3); ctx.deep_lde(f(1 << 11), 3, 3, 12345)
ctx.leaf_hash(f(1000), "out_off_7"); ctx.merkle_root(rng.integers(0, 256, (1001, 32), dtype=np.uint8))
for dd in (0, 1, 2):
    ctx.set_option("dedup", dd)
    roots, tree = ctx.column_commit(np.stack([f(4096), (np.arange(4096) // 512).astype(np.uint64), np.zeros(4096, np.uint64)]), ["a", "bb", "ccc"], keep=True)
    tree.open(np.array([0, 1, 2], np.uint32), np.array([5, 4095, 1024], np.uint64)); tree.free()
ctx.set_option("dedup", 2)
r, fin, h = ctx.fri_commit(f(1 << 13), f(13), keep=True); h.open(np.array([1, 8191], np.uint64)); h.free()
ctx.lde_commit(f((2, 1 << 11)), ["c_0", "c_1"], 3)
ct = m.simulate(1 << 12, 512, 2)
root = m.manifest_root(ct)
p1 = ctx.prove_v1(ct, root)
rt = ctx.upload_trace(ct); p2 = ctx.prove_v1_resident(rt, root); rt.free()
p3 = ctx.prove_v1_stream([ct], root)
assert p1 == p2 == p3
print("sanitize run ok", len(p1))
