#!/usr/bin/env python
"""Run the CPU oracle's prove_v1 once at large sizes and store length + BLAKE3 of the proof bytes
(tests/golden/prove_digests.json).  The GPU parity test at the headline size (T=2^22)
compares against these digests instead of re-running minutes of CPU work.  Usage: make_prove_digests.py 20 22"""
import importlib, json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
orc = oracle_lib.load()
path = os.path.join(HERE, "prove_digests.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for lt in [int(x) for x in sys.argv[1:]]:
    ct = m.simulate(1 << lt, 512, 8, seed=42)
    root = m.manifest_root(ct)
    t = time.time()
    proof = orc.prove_v1(ct, root)
    out[f"sim_T2^{lt}_b512_tau8_seed42"] = {"log_T": lt, "manifest_root": root.hex(), "proof_len": len(proof),
                                           "proof_blake3": orc.blake3(proof).hex(), "oracle_seconds": round(time.time() - t, 1)}
    json.dump(out, open(path, "w"), indent=1)
    print(lt, out[f"sim_T2^{lt}_b512_tau8_seed42"], flush=True)
