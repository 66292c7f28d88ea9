#!/usr/bin/env python
"""Named-scale runs that do not fit the default bench budget, on ONE multi-GPU context (sezkp_cuda_create_multi over all
visible GPUs, or SEZKP_DEVICES=0,1,..):
  wide      BASELINE configs[3]: 256 columns x 2^24 rows, LDE + labeled commit + FRI  (W_COLS / W_LOG_N override)
  proof26   one STARK v1 proof at T = 2^26 (PROOF_LOG_T), host pinned input and resident
  jsonl26   BASELINE configs[4]: streaming JSONL prove at T = 2^26 (JSONL_LOG_T): native writer -> file -> native parser on
            all host cores -> pinned staging ring -> first GPU -> NVLink replication -> sharded prove
usage: scale_r02.py [wide] [proof26] [jsonl26]   -> one JSON object on stdout"""
import importlib, json, os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
b = importlib.import_module("streaming-zero-knowledge-proofs_b200.binding")
import bench

env = lambda k, d: int(os.environ.get(k, d))
devices = [int(x) for x in os.environ["SEZKP_DEVICES"].split(",")] if "SEZKP_DEVICES" in os.environ else list(range(torch.cuda.device_count()))
want = sys.argv[1:] or ["wide", "proof26", "jsonl26"]
out = {"devices": devices, "host_cores": len(os.sched_getaffinity(0))}
hbm, _ = bench.peaks()
if "wide" in want:
    os.environ.setdefault("SEZKP_W_COLS", str(env("W_COLS", 256)))
    os.environ.setdefault("SEZKP_W_LOG_N", str(env("W_LOG_N", 24)))
    out["wide"] = bench.wide_bench(m, devices, hbm, 2)
    print("wide done", out["wide"]["ms_per_step"], file=sys.stderr, flush=True)
if "proof26" in want or "jsonl26" in want:
    lt = env("PROOF_LOG_T", 26)
    t0 = time.time()
    ct = m.simulate(1 << lt, 512, 8, seed=42)
    root = m.manifest_root(ct)
    out["simulate_s"] = time.time() - t0
    g = m.Context(devices=devices) if len(devices) > 1 else m.Context(devices[0])
    buf = np.empty(b.proof_size_bound(ct.n_rows, ct.tau), np.uint8)
    one_shot = None
    if "proof26" in want:
        ctp = bench.pin_trace(torch, ct)
        for _ in range(2):
            p = g.prove_v1(ctp, root, buf)
        t0 = time.perf_counter()
        for _ in range(3):
            p = g.prove_v1(ctp, root, buf)
        e2e = (time.perf_counter() - t0) / 3
        ph = g.timings()
        rt = g.upload_trace(ctp)
        for _ in range(2):
            g.prove_v1_resident(rt, root, buf)
        t0 = time.perf_counter()
        for _ in range(3):
            p2 = g.prove_v1_resident(rt, root, buf)
        res = (time.perf_counter() - t0) / 3
        out["proof"] = {"log_T": lt, "n_gpus": len(devices), "e2e_ms": e2e * 1e3, "e2e_rows_per_s": (1 << lt) / e2e, "resident_ms": res * 1e3,
                        "resident_rows_per_s": (1 << lt) / res, "same_bytes": bool(p == p2), "proof_bytes": len(p), "e2e_phases_ms_gpu0": ph,
                        "resident_phases_ms_gpu0": g.timings()}
        rt.free()
        one_shot = p
        print("proof done", out["proof"]["e2e_ms"], file=sys.stderr, flush=True)
    if "jsonl26" in want:
        d = tempfile.mkdtemp(prefix="sezkp_jsonl_", dir=os.environ.get("SEZKP_TMP", "/dev/shm" if os.path.isdir("/dev/shm") else None))
        path = os.path.join(d, "blocks.jsonl")
        ct.ops = None  # the writer reads the plain arrays
        t0 = time.time()
        size = b.write_jsonl_native(path, ct)
        wsec = time.time() - t0
        try:
            T = ct.n_rows
            p1 = g.prove_v1_jsonl_file(path, root, T, 8, expected_rows=T)
            t0 = time.perf_counter()
            p2 = g.prove_v1_jsonl_file(path, root, T, 8, expected_rows=T)
            dt = time.perf_counter() - t0
            tm = g.timings()
            out["jsonl"] = {"log_T": lt, "n_gpus": len(devices), "file_GB": size / 1e9, "write_s": wsec, "e2e_s": dt, "rows_per_s": T / dt,
                            "file_GBps": size / dt / 1e9, "same_bytes_twice": bool(p1 == p2), "timings": tm,
                            "identical_to_one_shot_proof": (bool(p2 == one_shot) if one_shot is not None else None)}
        finally:
            import shutil
            shutil.rmtree(d, ignore_errors=True)
    g.close()
print(json.dumps(out))
