#!/usr/bin/env python
"""bench.py — STARK v1 prove throughput (trace-rows/s) on B200, with roofline and CPU baseline.

Headline workload (BASELINE.json configs[2]): STARK v1 prove of a simulated trace, T = 2^22 rows, b = 512,
tau = 8 (59 committed columns, LDE domain 2^25, 26 FRI layers, 2250 column openings, 30x25 FRI pairs) on one
B200.  A "step" is one complete proof.

  value   rows/s with the compact trace already resident in HBM when the timed region starts
          (sezkp_stark_v1_prove_resident), CUDA events on the library's stream, max over ranks
  e2e     rows/s through the public API StarkV1Cuda.prove / sezkp_stark_v1_prove with HOST (pinned) buffers:
          H2D of the compact trace and D2H of roots / openings / proof inside the timed region
  roofline   dominant kernel of the step (chunk_commit_kernel: labeled BLAKE3 leaves + 1024-leaf chunk trees of
          the 59 columns), timed alone with CUDA events; plus the config-2 NTT/iNTT/LDE microbench in `micro`
  cpu_baseline  the CPU oracle (port of the reference's single-threaded prover, compute-once form) on a bounded
          sample, 1 core

N > 1 (torchrun): every rank proves its own trace (independent proofs, weak scaling, no data-path collective);
the 32-byte FRI/column roots are all-gathered with NCCL only so that rank 0 can report all proofs.
`--impl reference` times the oracle port on the host cores instead (rank 0 only).
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "streaming-zero-knowledge-proofs_b200"
METRIC = "stark_v1_prove_trace_rows_per_s"
UNIT = "rows/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload():
    log_t = env_int("SEZKP_BENCH_LOG_T", 22)
    return {"workload": f"STARK v1 prove, simulated trace T=2^{log_t}, b=512, tau=8 (59 columns, blow-up 8, 30 queries)",
            "log_T": log_t, "b": 512, "tau": 8, "l2": "inputs_exceed_l2 (2 GB of committed columns per step)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "utilization.gpu")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        def num(x):
            return x.replace(".", "").isdigit()
        loaded = [r for r in self.rows if len(r) > 9 and num(r[9]) and float(r[9]) >= 50 and num(r[1])]
        sm = [float(r[1]) for r in (loaded or self.rows) if len(r) > 2 and num(r[1])]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "samples_under_load": len(loaded)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ref_worker(job):
    """One oracle proof in a worker process (reference arm): returns its wall time."""
    log_t, seed = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    m = importlib.import_module(PKG)
    orc = oracle_lib.load()
    ct = m.simulate(1 << log_t, 512, 8, seed=seed)
    root = m.manifest_root(ct)
    t0 = time.perf_counter()
    orc.prove_v1(ct, root)
    return time.perf_counter() - t0


def run_reference(args, rank):
    """Reference arm: the oracle port of the reference's CPU prover on the host cores (the Rust reference cannot be
    compiled in this image).  The reference's prover is single-threaded, so "all the host threads it can use" is one
    independent proof per core: a step is `cores` concurrent T=2^k proofs, value = cores * 2^k rows / wall time of the
    step; the one-core figure is reported next to it."""
    if rank != 0:
        return
    import multiprocessing as mp
    log_t = env_int("SEZKP_REF_LOG_T", 14)
    cores = env_int("SEZKP_REF_CORES", len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    ctxm = mp.get_context("fork")
    single = _ref_worker((log_t, 42))  # also builds / warms the oracle library before the pool forks
    with ctxm.Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):
            pool.map(_ref_worker, [(log_t, 42 + i) for i in range(cores)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, [(log_t, 42 + i) for i in range(cores)])
        dt = (time.perf_counter() - t0) / args.steps
    v = cores * (1 << log_t) / dt
    sample = (f"oracle prove_v1 (compute-once form of the reference algorithm), {cores} independent proofs at T=2^{log_t}, b=512, "
              f"tau=8 per step, one per host core")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload(),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_thread_value": (1 << log_t) / single},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def pin_trace(torch, ct):
    """Move the compact trace's arrays into pinned host memory (so the e2e H2D runs at full PCIe rate)."""
    ct.pack_ops()  # SEZKP_TRACE_PACKED_OPS: 1 + tau bytes per row over PCIe instead of 1 + 4 tau (symbols < 32)
    names = ["block_len", "win_left", "win_right", "head_in_off", "head_out_off", "input_mv", "mv", "write_flag", "write_sym"]
    if ct.ops is not None:
        names.append("ops")
    for name in names:
        a = np.ascontiguousarray(getattr(ct, name))
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
        v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        setattr(ct, name, v)
        ct._keep.append(t)
    return ct


def timed(torch, fn, steps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps  # ms


def micro_bench(torch, ctx, hbm_peak):
    """Config 2 (BASELINE.json configs[1]): 64 columns x 2^20, det_vec(seed 2024+c); forward NTT, inverse NTT, coset LDE x4."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    cols, k, lb = 64, 20, 2
    n = 1 << k
    rng = np.random.default_rng(2024)
    host = (rng.integers(0, 1 << 63, size=(cols, n), dtype=np.uint64) % np.uint64(0xFFFFFFFF00000001))
    d = torch.from_numpy(host.view(np.int64)).cuda()
    out = torch.empty((cols, n << lb), dtype=torch.int64, device="cuda")
    res = {}
    for name, fn, bytes_ in (
        ("ntt_forward", lambda: ctx.ntt_dev(d, k, cols, False), 16 * n * cols),
        ("ntt_inverse", lambda: ctx.ntt_dev(d, k, cols, True), 16 * n * cols),
        ("coset_lde_x4", lambda: ctx.coset_lde_dev(d, k, lb, 3, cols, out), 8 * n * (1 + (1 << lb)) * cols),
    ):
        for _ in range(3):
            fn()
        ms = timed(torch, fn, 5)
        gbs = bytes_ / ms / 1e6
        res[name] = {"ms": ms, "algorithmic_GB": bytes_ / 1e9, "GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
        if name.startswith("ntt"):
            # the binding unit is the ALU pipe, not HBM: 112 ALU-pipe instructions per element and pass (ncu:
            # profiles/ncu_full_r01_final_ntt_pass_kernel.csv), two passes at 2^20, 64 lanes/clk/SM on that pipe
            alu_ms = cols * n * 2 * 112 / (148 * 64 * 1.965e9) * 1e3
            res[name]["int_alu"] = {"alu_pipe_instr_per_element_pass": 112, "passes": 2, "alu_pipe_bound_ms": alu_ms,
                                    "frac_of_alu_pipe_bound": alu_ms / ms, "hbm_frac_at_alu_pipe_bound": bytes_ / alu_ms / 1e6 / hbm_peak}
    del d, out
    return res


def lde_commit_bench(torch, dist, ctx, rank, world, hbm_peak, steps):
    """BASELINE configs[3] shape family (LDE + labeled BLAKE3 column commit), columns sharded c % world, roots all-gathered
    with NCCL.  Default 64 columns x 2^22 rows, blow-up 8 (SEZKP_W_COLS / SEZKP_W_LOG_N select e.g. the full 256 x 2^24)."""
    cols, k, lb = env_int("SEZKP_W_COLS", 64), env_int("SEZKP_W_LOG_N", 22), 3
    n = 1 << k
    mine = list(range(rank, cols, world))
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED + rank)
    ev = torch.randint(0, (1 << 62), (max(len(mine), 1), n), dtype=torch.int64, device="cuda", generator=g)  # < 2^62 < p: canonical
    labels = [f"c_{c}" for c in mine]

    def step():
        roots = ctx.lde_commit(ev, labels, lb, 3, dev=True, log_n=k) if mine else np.zeros((0, 32), np.uint8)
        if world > 1:  # C1: all-gather of 32-byte column roots (padded to the per-rank maximum)
            pad = torch.zeros(((cols + world - 1) // world) * 32, dtype=torch.uint8, device="cuda")
            pad[: roots.size] = torch.from_numpy(roots.reshape(-1)).cuda()
            out = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(out, pad)
        return roots

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    alg = cols * (8 * n * (1 + (1 << lb)) + 8 * (n << lb))  # LDE (read n, write B*n) + commit (read B*n)
    del ev
    return {"workload": f"{cols} columns x 2^{k} rows, blow-up 8: iNTT + coset LDE + labeled BLAKE3 commit, columns sharded over {world} GPU(s)",
            "ms_per_step": ms, "rows_per_s": n / (ms / 1e3), "algorithmic_GB": alg / 1e9, "GBps": alg / ms / 1e6,
            "GBps_per_gpu": alg / ms / 1e6 / world, "frac_of_hbm_peak_per_gpu": alg / ms / 1e6 / world / hbm_peak,
            "leaf_compressions_per_s": cols * (2 * (n << lb) - 1) / (ms / 1e3)}


def jsonl_stream_bench(torch, ctx, m, steps):
    """BASELINE configs[4] family on one GPU (scaled: T = 2^SEZKP_JSONL_LOG_T rows, default 2^19 ~ 100 MB of JSONL):
    .jsonl file -> native multi-threaded parser -> pinned staging ring -> H2D on a side stream -> prove.  The file is
    in the page cache; parsing is the bound, so the parser's rate and the hidden fraction of the copies are reported."""
    import tempfile
    log_t = env_int("SEZKP_JSONL_LOG_T", 19)
    T = 1 << log_t
    ct = m.simulate(T, 512, 8, seed=77)
    root = m.manifest_root(ct)
    d = tempfile.mkdtemp(prefix="sezkp_jsonl_")
    path = os.path.join(d, "blocks.jsonl")
    m.io_jsonl.write_jsonl(path, ct)
    size = os.path.getsize(path)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        ref = ctx.prove_v1(ct, root)
        ok = ctx.prove_v1_jsonl_file(path, root, T, 8, threads=threads, expected_rows=T) == ref
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.prove_v1_jsonl_file(path, root, T, 8, threads=threads, expected_rows=T)
        dt = (time.perf_counter() - t0) / steps
        tm = ctx.timings()
        # the pure-Python reader this replaces, on a slice (it parses ~1e5 rows/s)
        small = m.simulate(1 << 14, 512, 8, seed=78)
        sp = os.path.join(d, "small.jsonl")
        m.io_jsonl.write_jsonl(sp, small)
        t0 = time.perf_counter()
        ctx.prove_v1_stream(m.io_jsonl.stream_jsonl(sp), m.manifest_root(small))
        py_dt = time.perf_counter() - t0
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    return {"workload": f"streaming JSONL STARK prove, T=2^{log_t}, b=512, tau=8, {size / 1e6:.0f} MB file (page cache), 1 GPU",
            "rows_per_s": T / dt, "ms_per_step": dt * 1e3, "file_MBps": size / dt / 1e6, "parser_threads": threads,
            "proof_identical_to_one_shot": bool(ok),
            "jsonl_parse_ms": tm.get("jsonl_parse_ms"), "jsonl_read_ms": tm.get("jsonl_read_ms"),
            "stream_h2d_copy_ms": tm.get("stream_h2d_copy_ms"), "stream_copy_hidden_frac": tm.get("stream_copy_hidden_frac"),
            "prove_ms_after_last_line": tm.get("total"),
            "python_reader_rows_per_s": (1 << 14) / py_dt,
            "note": "host JSON parsing bounds this path; the GPU part is prove_ms_after_last_line"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-micro", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("SEZKP_NCCL_DEBUG", "NONE")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = importlib.import_module(PKG)
    ctx = m.Context(local)
    stream = torch.cuda.Stream()  # the library launches on torch's current stream so torch events bracket its work
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # every timed proof builds its subtree tables: the cross-proof table cache (a serving optimisation: tables depend only on
    # labels and value ranges) would otherwise hit on every repetition of the same synthetic trace
    ctx.set_option("tab_cache", 0)
    hbm_peak, peak_src = peaks()

    wl = workload()
    T = 1 << wl["log_T"]
    ct = pin_trace(torch, m.simulate(T, wl["b"], wl["tau"], seed=42 + rank))
    root = m.manifest_root(ct)
    n_cols = 3 + 7 * ct.tau
    from importlib import import_module
    binding = import_module(PKG + ".binding")
    proof_buf_t = torch.empty(binding.proof_size_bound(ct.n_rows, ct.tau), dtype=torch.uint8, pin_memory=True)
    proof_buf = proof_buf_t.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ----
    rt = ctx.upload_trace(ct)
    proof = None
    for _ in range(args.warmup):
        proof = ctx.prove_v1_resident(rt, root, proof_buf)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ctx.launch_count(reset=True)
    # view=True: the proof is read from the caller's pinned buffer (what a native caller of the C ABI gets), no Python copy
    ms = timed(torch, lambda: ctx.prove_v1_resident(rt, root, proof_buf, view=True), args.steps)
    launches = ctx.launch_count()
    barrier()
    phases = ctx.timings()
    ms = max_over_ranks(ms)
    value = world * T / (ms / 1e3)

    # ---- end-to-end arm: host (pinned) buffers through the public API ----
    for _ in range(min(args.warmup, 2)):
        ctx.prove_v1(ct, root, proof_buf)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proof = ctx.prove_v1(ct, root, proof_buf, view=True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    proof = bytes(proof)
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_phases = ctx.timings()
    h2d_bytes = ct.nbytes() + 8 * ct.n_blocks
    log_N = wl["log_T"] + 3
    d2h_bytes = len(proof) + 32 * n_cols + 32 * (log_N + 1)
    rt.free()

    sharded = None
    if world > 1:
        # (a) NCCL gathers one 32-byte digest per rank so rank 0 can attest every independent proof was produced
        import hashlib
        mine = torch.frombuffer(bytearray(hashlib.sha256(proof).digest()), dtype=torch.uint8).cuda()
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        # (b) ONE proof with its columns sharded over all ranks (north_star's multi-GPU layout): commitments and openings of
        #     column c on rank c % world, composition/LDE/FRI replicated, two NCCL all-gathers of roots / opening records
        ct0 = ct if rank == 0 else None
        ct_same = pin_trace(torch, m.simulate(T, wl["b"], wl["tau"], seed=42)) if rank != 0 else ct
        root0 = m.manifest_root(ct_same)
        cb = m.parallel.dist_allgather_callback(torch.device("cuda", local))
        # device-side collective (NCCL over NVLink): sharded upload + all-gather of the compact trace, FRI subtree roots
        dev_cb = m.parallel.dist_allgather_dev_callback(torch.device("cuda", local))
        ctx.set_allgather_dev(dev_cb)
        for _ in range(2):
            p_sh = ctx.prove_v1_sharded(ct_same, root0, rank, world, cb, proof_buf)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p_sh = ctx.prove_v1_sharded(ct_same, root0, rank, world, cb, proof_buf)
        torch.cuda.synchronize()
        sh_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        same = torch.tensor([int(hashlib.sha256(p_sh).digest()[:7].hex(), 16)], dtype=torch.int64, device="cuda")
        lo, hi = same.clone(), same.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        sh_phases = ctx.timings()
        # same with the trace already resident on every GPU (what `value` is for the independent proofs): without the
        # world-fold replicated H2D of the whole trace, which is what the e2e figure above mostly measures at N = 8
        rt_sh = ctx.upload_trace(ct_same)
        for _ in range(2):
            p_rs = ctx.prove_v1_resident_sharded(rt_sh, root0, rank, world, cb, proof_buf)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p_rs = ctx.prove_v1_resident_sharded(rt_sh, root0, rank, world, cb, proof_buf)
        torch.cuda.synchronize()
        rs_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        rs_phases = ctx.timings()
        rt_sh.free()
        sharded = {"ms_per_proof": sh_ms, "rows_per_s": T / (sh_ms / 1e3), "identical_on_all_ranks": bool(lo.item() == hi.item()),
                   "phases_ms_rank0": sh_phases, "note": "one T-row proof: rows uploaded 1/world per rank + NVLink all-gather of the compact trace, columns sharded c % world, FRI hashing sharded by chunk range (e2e from host pinned input)",
                   "resident": {"ms_per_proof": rs_ms, "rows_per_s": T / (rs_ms / 1e3), "identical_to_e2e_proof": bool(p_rs == p_sh),
                                "phases_ms_rank0": rs_phases, "note": "same prover, trace already in HBM on every rank"}}

    out = None
    if rank == 0:
        # ---- dominant kernel alone: chunk_commit_kernel on FRI layer 0 (N = 8T unlabeled leaves -> BLAKE3 leaf hashes ->
        #      1024-leaf chunk trees; + the two upper_reduce launches), the largest single launch of the step ----
        N = 8 * ct.n_rows
        g = torch.Generator(device="cuda")
        g.manual_seed(7)
        layer0 = torch.randint(0, (1 << 62), (N,), dtype=torch.int64, device="cuda", generator=g)  # < 2^62 < p: canonical
        ctx.set_option("dedup", 0)  # FRI layers are high-entropy: the prover hashes them with the plain kernel
        for _ in range(3):
            ctx.column_commit(layer0, None, dev=True, n=N, c=1)
        kms = timed(torch, lambda: ctx.column_commit(layer0, None, dev=True, n=N, c=1), max(3, args.steps))
        ctx.set_option("dedup", 2)
        del layer0
        alg_bytes = 8 * N            # column commit from values, root only: 8 B per leaf read once
        achieved = alg_bytes / kms / 1e6
        compressions = 2 * N - 1
        alu_bound = 148 * 64 * 1.965e9 / 455
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tpath) and wl["log_T"] == 22:
            traffic = json.load(open(tpath)).get("chunk_commit_kernel_fri_layer0_dram_bytes_per_launch")
        # second kernel family of the step: the value-aware commit of the 59 trace columns (subtree tables)
        cols_host = ctx.trace_columns(ct)
        cols_dev = torch.from_numpy(cols_host.view(np.int64)).cuda()
        labels = ["input_mv", "is_first", "is_last"] + [f"{g}_{r}" for g in ("mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off")
                                                        for r in range(ct.tau)]
        cc = {}
        for name, dd, tb in (("tables", 2, 1), ("dedup_only", 2, 0), ("plain", 0, 0)):
            ctx.set_option("dedup", dd)
            ctx.set_option("tabled", tb)
            for _ in range(2):
                ctx.column_commit(cols_dev, labels, dev=True, n=ct.n_rows)
            cc[name] = timed(torch, lambda: ctx.column_commit(cols_dev, labels, dev=True, n=ct.n_rows), max(3, args.steps))
            if tb:
                cc["tabled_columns"], cc["chunks_redone"] = ctx.tab_stats()
        ctx.set_option("dedup", 2)
        ctx.set_option("tabled", 1)
        del cols_dev, cols_host
        col_bytes = 8 * ct.n_rows * n_cols
        roofline = {"bound": "hbm", "kernel": "chunk_commit_kernel<FOLD=0> on FRI layer 0 (+2 upper_reduce launches)", "achieved": achieved,
                    "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                    "traffic_unit": "bytes per launch (ncu)", "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": kms, "share_of_step": kms / ms,
                    "note": "The contract's roofline is HBM, but BLAKE3 leaf+tree hashing is integer-ALU bound: 8 B in, 2 "
                            "single-block compressions per leaf (455 ALU-pipe + 335 FMA-pipe instructions each).  int_alu states "
                            "the binding fraction.  The same kernel with the fused fold (FOLD=1) runs the other 25 FRI layers; "
                            "together they are ~45 % of the step.",
                    "int_alu": {"compressions_per_s": compressions / (kms / 1e3), "alu_pipe_instr_per_compression": 455,
                                "fma_pipe_instr_per_compression": 335, "alu_pipe_bound_compressions_per_s": alu_bound,
                                "frac_of_alu_pipe_bound": compressions / (kms / 1e3) / alu_bound},
                    "column_commit": {"kernel": "chunk_commit_tabled_kernel<LOGG> x4 (+ stats, table build, upper_reduce)",
                                      "ms_per_launch": cc["tables"], "algorithmic_bytes_per_launch": col_bytes,
                                      "achieved": col_bytes / cc["tables"] / 1e6, "frac": col_bytes / cc["tables"] / 1e6 / hbm_peak,
                                      "share_of_step": cc["tables"] / ms, "tabled_columns": cc["tabled_columns"],
                                      "chunks_redone_by_generic_kernel": cc["chunks_redone"],
                                      "dedup_only_ms_per_launch": cc["dedup_only"], "plain_ms_per_launch": cc["plain"],
                                      "plain_compressions_per_s": n_cols * (2 * ct.n_rows - 1) / (cc["plain"] / 1e3),
                                      "note": "value-aware: identical subtrees are hashed once per column (tables) or per chunk "
                                              "(dedup); plain = one compression per node; all three give identical roots"}}
        micro = None if args.no_micro else micro_bench(torch, ctx, hbm_peak)
        jsonl_stream = None if args.no_micro else jsonl_stream_bench(torch, ctx, m, 2)
    # BASELINE configs[3] family last: ~0.7 s of sustained hashing per step, after which the board sits at its power cap
    lde_commit = None if args.no_micro else lde_commit_bench(torch, dist, ctx, rank, world, hbm_peak, max(2, min(args.steps, 3)))
    if rank == 0:

        sampler.stop_flag = True  # the GPU is idle from here on
        sampler.join(timeout=2)
        # ---- CPU baseline: oracle port, bounded sample ----
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        orc = oracle_lib.load()
        log_c = env_int("SEZKP_CPU_LOG_T", 15)
        cct = m.simulate(1 << log_c, 512, 8)
        croot = m.manifest_root(cct)
        t0 = time.perf_counter()
        cproof = orc.prove_v1(cct, croot)
        cdt = time.perf_counter() - t0
        parity = ctx.prove_v1(cct, croot) == cproof  # same bytes on the same input (checker role of the oracle)
        # the reference's actual cost profile: its layer-0 FRI paths re-run the whole DEEP-LDE stream once per Merkle level
        # (v1/fri_stream.rs:273-309) — timed with the oracle's faithful-cost mode at a size that finishes in seconds
        log_f = env_int("SEZKP_CPU_FAITHFUL_LOG_T", 8)
        fct = m.simulate(1 << log_f, 512, 8)
        froot = m.manifest_root(fct)
        t0 = time.perf_counter()
        fproof = orc.prove_v1(fct, froot, faithful_cost=True)
        fdt = time.perf_counter() - t0
        faithful = {"value": (1 << log_f) / fdt, "unit": UNIT, "log_T": log_f, "seconds": fdt,
                    "identical_to_compute_once": bool(fproof == orc.prove_v1(fct, froot)),
                    "note": "30 queries x 2 paths x log2(8n) full recomputations of the LDE stream, like the reference"}
        cpu_baseline = {"value": (1 << log_c) / cdt, "unit": UNIT, "cores": 1, "kind": "port", "faithful_cost": faithful,
                        "sample": f"one oracle prove_v1 at T=2^{log_c}, b=512, tau=8 ({cdt:.1f} s); reference is single-threaded",
                        "gpu_proof_identical": bool(parity)}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (Goldilocks field) / u32 (BLAKE3)", "data": "synthetic",
            "config": dict(wl, parallelism=f"independent proofs x{world}" if world > 1 else "single GPU"),
            "clocks": sampler.summary(),
            "e2e": {"value": world * T / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "api": "sezkp_stark_v1_prove (host pinned buffers)"},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "phases_ms": phases, "e2e_phases_ms": e2e_phases, "proof_bytes": len(proof), "micro": micro, "lde_commit": lde_commit, "jsonl_stream": jsonl_stream, "sharded_single_proof": sharded,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
