import importlib, os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
ctx = m.Context(0)
ct = m.simulate(1 << 20, 512, 8)
cols = ctx.trace_columns(ct)
lib = m.load_library()
out = (C.c_ulonglong * 4)()
for group in ("winlen", "in_off", "mv"):
    gi = ["mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off"].index(group)
    sub = torch.from_numpy(cols[3 + 8 * gi: 3 + 8 * gi + 8].copy().view(np.int64)).cuda()
    lib.sezkp_debug_memo_stats(out, 1)
    ctx.column_commit(sub, [f"{group}_{r}" for r in range(8)], dev=True, n=ct.n_rows)
    ctx.synchronize()
    lib.sezkp_debug_memo_stats(out, 1)
    vals = cols[3 + 8 * gi]
    print(group, "hits", out[0], "misses", out[1], "claimed", out[2], "distinct values col0:", len(np.unique(vals)), "chunks", ct.n_rows // 1024)
