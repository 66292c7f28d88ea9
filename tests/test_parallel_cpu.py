"""N>1 host logic on CPU: the sharding contract and the all-gather callback (gloo, world_size 2, and threads).

The GPU kernels cannot run here, so each rank computes the roots / opening records of ITS columns with the CPU
oracle (checker role), pushes them through exactly the callback the C ABI would call, and the merged result must equal
the single-rank answer.  The GPU-side use of the same callback is covered by tests/test_gpu_parity.py (threads on one GPU).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, pkg


def _call(cb, send: bytes, world: int) -> bytes:
    recv = C.create_string_buffer(world * len(send))
    rc = cb(None, C.cast(C.c_char_p(send), C.c_void_p), len(send), C.cast(recv, C.c_void_p))
    assert rc == 0
    return recv.raw


def _rank_roots(orc, m, cols, labels, rank, world):
    par = m.parallel
    mine = par.local_columns(len(labels), rank, world)
    roots = orc.column_commit(cols[mine], [labels[c] for c in mine]) if mine else np.zeros((0, 32), np.uint8)
    pad = np.zeros((par.max_local(len(labels), world), 32), np.uint8)
    pad[: len(mine)] = roots
    return pad.tobytes()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch.distributed as dist
    import oracle_lib
    m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = oracle_lib.load()
        ct = m.simulate(256, 64, 2)
        cols = orc.trace_columns(ct)
        labels = ["input_mv", "is_first", "is_last"] + [f"{g}_{r}" for g in ("mv", "wflag", "wsym", "head", "winlen", "in_off", "out_off") for r in range(2)]
        cb = m.parallel.dist_allgather_callback()
        gathered = _call(cb, _rank_roots(orc, m, cols, labels, rank, world), world)
        merged = m.parallel.merge_roots(np.frombuffer(gathered, np.uint8), len(labels), world)
        full = orc.column_commit(cols, labels)
        # opening records: value(8) of (column o % n_cols, row o) owned by column owner
        k = 40
        owner = [m.parallel.owner_of_column(o % len(labels), world) for o in range(k)]
        rec = np.zeros((k, 8), np.uint8)
        for o in range(k):
            if owner[o] == rank:
                rec[o] = np.frombuffer(int(cols[o % len(labels), o]).to_bytes(8, "little"), np.uint8)
        g2 = _call(cb, rec.tobytes(), world)
        recs = m.parallel.merge_records(np.frombuffer(g2, np.uint8), owner, world)
        ok = np.array_equal(merged, full) and all(
            int.from_bytes(recs[o].tobytes(), "little") == int(cols[o % len(labels), o]) for o in range(k))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharding_contract_indices():
    par = pkg().parallel
    for n_cols, world in [(59, 1), (59, 2), (59, 4), (59, 8), (17, 3), (3, 8)]:
        seen = sorted(c for r in range(world) for c in par.local_columns(n_cols, r, world))
        assert seen == list(range(n_cols))
        for r in range(world):
            assert all(par.owner_of_column(c, world) == r for c in par.local_columns(n_cols, r, world))
            assert len(par.local_columns(n_cols, r, world)) <= par.max_local(n_cols, world)


def test_gloo_world2_allgather_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_thread_group_allgather(oracle):
    m = pkg()
    world = 3
    tg = m.parallel.ThreadGroup(world)
    ct = m.simulate(128, 32, 1)
    cols = oracle.trace_columns(ct)
    labels = ["input_mv", "is_first", "is_last", "mv_0", "wflag_0", "wsym_0", "head_0", "winlen_0", "in_off_0", "out_off_0"]

    def rank_fn(r):
        g = _call(tg.callback(r), _rank_roots(oracle, m, cols, labels, r, world), world)
        return m.parallel.merge_roots(np.frombuffer(g, np.uint8), len(labels), world)

    outs = tg.run(rank_fn)
    full = oracle.column_commit(cols, labels)
    assert all(np.array_equal(o, full) for o in outs)
