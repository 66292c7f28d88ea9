#!/usr/bin/env python
"""Config-2 microbench only: 64 x 2^20 forward NTT, inverse NTT, coset LDE x4 (+ optional other shapes)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
ctx = m.Context(0)
ctx.set_option("ntt_gen", int(os.environ.get("NTT_GEN", "2")))  # 1: first-generation pass kernel (A/B runs)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
P = 0xFFFFFFFF00000001
def timed(fn, k=5):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(k): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / k
if os.environ.get("NTT_CHECK", "1") == "1":  # parity first: config-2 column 0 against the oracle digests
    import json, blake3
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import det_vec_fast
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "named_shape_digests.json")))
    g2 = gold["config2_64x2^20_det_vec_seed2024+c"]["0"]
    v = det_vec_fast(1 << 20, 2024)
    ok = [blake3.blake3(ctx.ntt(v).tobytes()).hexdigest() == g2["forward"],
          blake3.blake3(ctx.ntt(v, inverse=True).tobytes()).hexdigest() == g2["inverse_of_input"],
          blake3.blake3(ctx.coset_lde(v, 2, 3).tobytes()).hexdigest() == g2["coset_k22_shift3"]]
    for k in (21, 22, 24):
        gk = gold["ntt_det_vec_seed7"][f"2^{k}"]
        v = det_vec_fast(1 << k, 7)
        ok.append(blake3.blake3(ctx.ntt(v).tobytes()).hexdigest() == gk["forward"])
        ok.append(blake3.blake3(ctx.ntt(v, inverse=True).tobytes()).hexdigest() == gk["inverse"])
    v = det_vec_fast(1 << 22, 100)
    ok.append(blake3.blake3(ctx.lde_from_evals(v, 3, 3).tobytes()).hexdigest() == gold["lde_2^22_x8_det_vec_seed100+c"]["0"]["lde_x8_shift3"])
    print("parity vs oracle digests (fwd, inv, coset 2^20; fwd/inv 2^21, 2^22, 2^24; lde 2^22 x8):", ok, flush=True)
shapes = [(64, 20, 2)] + [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for cols, k, lb in shapes:
    n = 1 << k
    rng = np.random.default_rng(1)
    d = torch.from_numpy((rng.integers(0, 1 << 63, size=(cols, n), dtype=np.uint64) % np.uint64(P)).view(np.int64)).cuda()
    out = torch.empty((cols, n << lb), dtype=torch.int64, device="cuda")
    f = timed(lambda: ctx.ntt_dev(d, k, cols, False)); i = timed(lambda: ctx.ntt_dev(d, k, cols, True))
    l = timed(lambda: ctx.coset_lde_dev(d, k, lb, 3, cols, out))
    gb = lambda b, ms: b / ms / 1e6
    print(f"cols={cols} n=2^{k} B={1<<lb}: fwd {f:.3f} ms ({gb(16*n*cols,f):.0f} GB/s)  inv {i:.3f} ms ({gb(16*n*cols,i):.0f} GB/s)  lde {l:.3f} ms ({gb(8*n*(1+(1<<lb))*cols,l):.0f} GB/s)")
    del d, out
