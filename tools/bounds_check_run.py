#!/usr/bin/env python
"""Drive the NTT / LDE pass kernels over a spread of shapes and print one digest per case.  Run once with the normal
library and once with SEZKP_CUDA_LIB=.../libsezkp_cuda_dbg.so (the bounds-checked build: a violating address prints a
message and traps, so the process dies); tests/test_gpu_named_shapes.py::test_bounds_checked_build compares the outputs."""
import hashlib, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import det_vec_fast  # input generator only
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
ctx = m.Context(0)
dig = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
for L in (1, 2, 5, 6, 7, 8, 9, 10, 11, 13, 14, 16, 17, 19, 20, 21, 22, 23):
    for cols in ((1, 3) if L <= 20 else (1,)):
        v = np.stack([det_vec_fast(1 << L, 10 * L + c) for c in range(cols)])
        f = ctx.ntt(v)
        print("ntt", L, cols, dig(f), dig(ctx.ntt(f, inverse=True)), flush=True)
for L in (1, 4, 6, 9, 10, 11, 12, 15, 18, 20, 21):
    for lb in (0, 1, 2, 3, 4):
        if L + lb > 24:
            continue
        v = np.stack([det_vec_fast(1 << L, 7 * L + c) for c in range(2 if L < 20 else 1)])
        print("lde", L, lb, dig(ctx.coset_lde(v, lb, 3)), dig(ctx.lde_from_evals(v, lb, 7)), flush=True)
for L, lb, fuse in ((21, 3, 1), (21, 2, 1), (22, 3, 1), (21, 3, 0), (12, 3, 1), (20, 2, 1)):
    ctx.set_option("lde_fuse", fuse)
    v = np.stack([det_vec_fast(1 << L, 3 * L + c) for c in range(2)])
    print("lde_commit", L, lb, fuse, dig(ctx.lde_commit(v, ["c_0", "other"], lb)), flush=True)
print("deep", dig(ctx.deep_lde(det_vec_fast(1 << 16, 5), 3, 3, 123456789)), flush=True)
ct = m.simulate(1 << 14, 512, 8, seed=3)
print("prove", dig(np.frombuffer(ctx.prove_v1(ct, m.manifest_root(ct)), np.uint8)), flush=True)
print("done")
