#!/usr/bin/env python
"""Time the column commit per column group of a simulated trace (dedup on / off)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = m.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
ct = m.simulate(1 << log_t, 512, 8)
cols = ctx.trace_columns(ct)
n = ct.n_rows
dev = torch.from_numpy(cols.view(np.int64)).cuda()
names = ["input_mv", "is_first", "is_last"] + [f"{g}_{r}" for g in ("mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off") for r in range(8)]
groups = {"scalars": (0, 3)}
for gi, g in enumerate(("mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off")):
    groups[g] = (3 + 8 * gi, 3 + 8 * gi + 8)
def timed(fn, k=3):
    fn(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(k): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / k
for name, (a, b) in groups.items():
    sub = dev[a:b].contiguous()
    res = []
    for dd, tb in ((2, 1), (2, 0), (0, 0)):
        ctx.set_option("dedup", dd)
        ctx.set_option("tabled", tb)
        res.append(timed(lambda: ctx.column_commit(sub, names[a:b], dev=True, n=n)))
        if tb: stats = ctx.tab_stats()
    ctx.set_option("dedup", 2); ctx.set_option("tabled", 1)
    print(f"{name:8s} cols={b-a} tabled {res[0]:.3f} ms (tab cols {stats[0]}, redone chunks {stats[1]})  dedup {res[1]:.3f} ms  plain {res[2]:.3f} ms  "
          f"({res[0]/(b-a):.3f} / {res[1]/(b-a):.3f} / {res[2]/(b-a):.3f} per column)")
for dd, tb in ((2, 1), (2, 0), (0, 0)):
    ctx.set_option("dedup", dd); ctx.set_option("tabled", tb)
    print(f"all 59 columns dedup={dd} tabled={tb}: {timed(lambda: ctx.column_commit(dev, names, dev=True, n=n)):.3f} ms", ctx.tab_stats())
