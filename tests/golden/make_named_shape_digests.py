#!/usr/bin/env python
"""Oracle digests at the NAMED benchmark shapes (BASELINE.json configs), stored in
tests/golden/named_shape_digests.json so that the GPU parity tests can compare the CUDA path with the CPU oracle at sizes
the oracle needs minutes for.  Sections (run all, or name some on the command line):

  config2     64 columns det_vec(2^20, 2024+c) (sezkp-ffts/benches/ntt.rs:21-34): BLAKE3 over the LE bytes of the input,
              its forward NTT, the inverse NTT of the input, and evaluate_on_coset_pow2(col, 22, 3), per column
  ntt_large   det_vec(2^k, 7) for k = 21, 22, 24 (three-pass plans on the GPU): forward / inverse digests
  lde_large   two columns det_vec(2^22, 100+c): interpolate + coset LDE x8 (N = 2^25) digests
  quickstart  README quick-start shape T = 2^15, b = 512, tau = 8: proof length + BLAKE3
  wide20      config 4 generator (0x5EED splitmix, labels c_{c}) at 2^20 rows x blow-up 8: roots of columns 0..7 and 255,
              and the whole pipeline (alphas, z, FRI roots, final value) for 8 columns
  wide24      the same at the named row count 2^24 (N = 2^27): roots of columns 0..7, pipeline for 8 columns
  wide24_full config 4 exactly as named: all 256 column roots and the pipeline over 256 columns x 2^24 rows (hours of CPU;
              not in the default list)
Usage: make_named_shape_digests.py [section ...]   (SEZKP_GOLDEN_PROCS = worker processes, default all cores)"""
import importlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402

PATH = os.path.join(HERE, "named_shape_digests.json")


def b3(orc, a):
    return orc.blake3(np.ascontiguousarray(a, np.uint64).tobytes()).hex()


def job_config2(c):
    orc = oracle_lib.load()
    v = oracle_lib.det_vec_fast(1 << 20, 2024 + c)
    fwd = orc.ntt(v)
    assert np.array_equal(orc.ntt(fwd, inverse=True), v)  # criterion case (ii): inverse on (i)'s output gives the input back
    return c, {"input": b3(orc, v), "forward": b3(orc, fwd), "inverse_of_input": b3(orc, orc.ntt(v, inverse=True)),
               "coset_k22_shift3": b3(orc, orc.coset_eval(v, 22, 3))}


def job_ntt_large(k):
    orc = oracle_lib.load()
    v = oracle_lib.det_vec_fast(1 << k, 7)
    return k, {"input": b3(orc, v), "forward": b3(orc, orc.ntt(v)), "inverse": b3(orc, orc.ntt(v, inverse=True))}


def job_lde_large(c):
    orc = oracle_lib.load()
    v = oracle_lib.det_vec_fast(1 << 22, 100 + c)
    return c, {"input": b3(orc, v), "lde_x8_shift3": b3(orc, orc.lde_from_evals(v, 3, 3))}


def job_wide_root(a):
    c, log_n = a
    return c, oracle_lib.load().wide_column_root(c, log_n).hex()


def wide(pool, log_n, root_cols, tail_cols):
    t0 = time.time()
    roots = dict(pool.map(job_wide_root, [(c, log_n) for c in root_cols]))
    orc = oracle_lib.load()
    cr = np.frombuffer(b"".join(bytes.fromhex(roots[c]) for c in range(tail_cols)), np.uint8)
    w = orc.wide_tail(None, tail_cols, log_n, cr)
    return {"log_n": log_n, "log_blow": 3, "shift": 3, "column_roots": {str(c): roots[c] for c in root_cols},
            "pipeline": {"n_cols": tail_cols, "alphas": [int(x) for x in w["alphas"]], "z": int(w["z"]),
                         "betas": [int(x) for x in w["betas"]], "fri_roots": [r.tobytes().hex() for r in w["fri_roots"]],
                         "final_value": int(w["final"])},
            "oracle_seconds": round(time.time() - t0, 1)}


def main():
    want = sys.argv[1:] or ["config2", "ntt_large", "lde_large", "quickstart", "wide20", "wide24"]
    out = json.load(open(PATH)) if os.path.exists(PATH) else {}
    procs = int(os.environ.get("SEZKP_GOLDEN_PROCS", os.cpu_count() or 1))
    oracle_lib.load()  # build before forking

    def save():
        json.dump(out, open(PATH, "w"), indent=1, sort_keys=True)

    with mp.get_context("fork").Pool(procs) as pool:
        if "config2" in want:
            t0 = time.time()
            out["config2_64x2^20_det_vec_seed2024+c"] = {str(c): d for c, d in pool.map(job_config2, range(64))}
            print("config2", round(time.time() - t0, 1), "s", flush=True)
            save()
        if "ntt_large" in want:
            out["ntt_det_vec_seed7"] = {f"2^{k}": d for k, d in pool.map(job_ntt_large, [21, 22, 24])}
            save()
        if "lde_large" in want:
            out["lde_2^22_x8_det_vec_seed100+c"] = {str(c): d for c, d in pool.map(job_lde_large, [0, 1])}
            save()
        if "quickstart" in want:
            m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
            orc = oracle_lib.load()
            ct = m.simulate(1 << 15, 512, 8, seed=42)
            root = m.manifest_root(ct)
            proof = orc.prove_v1(ct, root)
            out["quickstart_T2^15_b512_tau8_seed42"] = {"manifest_root": root.hex(), "proof_len": len(proof), "proof_blake3": orc.blake3(proof).hex()}
            save()
        if "wide20" in want:
            out["wide_0x5EED_2^20"] = wide(pool, 20, list(range(8)) + [255], 8)
            print("wide20", out["wide_0x5EED_2^20"]["oracle_seconds"], "s", flush=True)
            save()
        if "wide24_full" in want:
            out["wide_0x5EED_2^24_256cols"] = wide(pool, 24, list(range(256)), 256)
            print("wide24_full", out["wide_0x5EED_2^24_256cols"]["oracle_seconds"], "s", flush=True)
            save()
        if "wide24" in want:
            out["wide_0x5EED_2^24"] = wide(pool, 24, list(range(8)), 8)
            print("wide24", out["wide_0x5EED_2^24"]["oracle_seconds"], "s", flush=True)
            save()


if __name__ == "__main__":
    main()
