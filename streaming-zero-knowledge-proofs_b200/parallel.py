"""Multi-GPU plumbing for the column-sharded prover (one process per GPU).

The C ABI (`sezkp_stark_v1_prove_sharded`) asks the host for exactly one collective: an all-gather of a small host
buffer (32-byte column roots, then the opening records).  This module provides that callback

  * over ``torch.distributed`` — NCCL over NVLink on the GPU box (the bytes are staged through a device tensor), or
    gloo on CPU-only machines (used by the CPU tests), and
  * over threads of one process (``ThreadGroup``), which lets a single-GPU box exercise world sizes > 1,

plus the pure index helpers that define the sharding contract (which rank owns which column, where a rank's roots sit
in the gathered buffer).  No data-path collective exists: field columns never cross GPUs.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Callable, List, Sequence

import numpy as np

ALLGATHER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
ALLGATHER_DEV_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)  # + cuda_stream


class _DevBytes:
    """`nbytes` bytes of device memory at `ptr` as a __cuda_array_interface__ object (zero-copy torch.as_tensor)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def dist_allgather_dev_callback(device):
    """Device-side all-gather (sezkp_allgather_dev_fn) over torch.distributed / NCCL: the library's device pointers are
    wrapped zero-copy and `all_gather_into_tensor` is enqueued on the library's own stream, so nothing synchronises
    with the host and later kernels of that stream see the gathered data."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()

    def _cb(user, send, nbytes, recv, stream):
        try:
            n = int(nbytes)
            st = torch.cuda.ExternalStream(int(stream or 0), device=device)
            with torch.cuda.stream(st):
                src = torch.as_tensor(_DevBytes(send, n), device=device)
                dst = torch.as_tensor(_DevBytes(recv, n * world), device=device)
                dist.all_gather_into_tensor(dst, src)
            return 0
        except Exception as e:  # never raise through the C frame
            print("sezkp device allgather callback failed:", repr(e))
            return -1

    return ALLGATHER_DEV_FN(_cb)


# ---------------------------------------------------------------------------------------------- sharding contract
def owner_of_column(c: int, world: int) -> int:
    return c % world


def local_columns(n_cols: int, rank: int, world: int) -> List[int]:
    """global column indices committed by `rank`, in local order"""
    return list(range(rank, n_cols, world))


def max_local(n_cols: int, world: int) -> int:
    return (n_cols + world - 1) // world


def merge_roots(gathered: np.ndarray, n_cols: int, world: int) -> np.ndarray:
    """gathered: [world][max_local][32] (rank-major, zero padded) -> [n_cols][32] in canonical column order"""
    g = np.asarray(gathered, np.uint8).reshape(world, max_local(n_cols, world), 32)
    return np.stack([g[c % world, c // world] for c in range(n_cols)])


def merge_records(gathered: np.ndarray, owner: Sequence[int], world: int) -> np.ndarray:
    """gathered: [world][k][rec] -> [k][rec] taking record o from rank owner[o]"""
    k = len(owner)
    g = np.asarray(gathered, np.uint8).reshape(world, k, -1)
    return np.stack([g[owner[o], o] for o in range(k)])


# ---------------------------------------------------------------------------------------------- collectives
def dist_allgather_callback(device=None):
    """All-gather callback over torch.distributed's default group (nccl: staged through `device`; gloo: CPU).

    The NCCL path keeps four grow-only buffers (pinned send / receive staging, device send / receive) and issues
    memmove -> H2D -> all_gather_into_tensor -> D2H on the current stream with one synchronisation at the end: an
    exchange of a few hundred KB costs ~0.1 ms instead of the ~0.6 ms of fresh pageable tensors per call."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    state = {"cap": 0}

    def _ensure(n):
        if n > state["cap"]:
            cap = max(2 * state["cap"], n, 1 << 16)
            state["pin_s"] = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            state["pin_r"] = torch.empty(cap * world, dtype=torch.uint8, pin_memory=True)
            state["dev_s"] = torch.empty(cap, dtype=torch.uint8, device=device)
            state["dev_r"] = torch.empty(cap * world, dtype=torch.uint8, device=device)
            state["cap"] = cap

    def _cb_nccl(user, send, nbytes, recv):
        try:
            n = int(nbytes)
            _ensure(n)
            C.memmove(state["pin_s"].data_ptr(), send, n)
            ds, dr = state["dev_s"][:n], state["dev_r"][: n * world]
            ds.copy_(state["pin_s"][:n], non_blocking=True)
            dist.all_gather_into_tensor(dr, ds)
            pr = state["pin_r"][: n * world]
            pr.copy_(dr, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            C.memmove(recv, pr.data_ptr(), n * world)
            return 0
        except Exception as e:  # never raise through the C frame
            print("sezkp allgather callback failed:", repr(e))
            return -1

    def _cb_cpu(user, send, nbytes, recv):
        try:
            src = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(nbytes,))
            t = torch.from_numpy(src.copy())
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            dst = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(world * nbytes,))
            dst[:] = torch.cat(outs).numpy()
            return 0
        except Exception as e:
            print("sezkp allgather callback failed:", repr(e))
            return -1

    return ALLGATHER_FN(_cb_nccl if device is not None else _cb_cpu)


class ThreadGroup:
    """`world` ranks as threads of one process; all-gather through a barrier (test / single-GPU emulation)."""

    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots: List[bytes] = [b""] * world

    def callback(self, rank: int):
        def _cb(user, send, nbytes, recv):
            try:
                self.slots[rank] = C.string_at(send, nbytes)
                self.barrier.wait(timeout=120)
                C.memmove(recv, b"".join(self.slots), self.world * nbytes)
                self.barrier.wait(timeout=120)
                return 0
            except Exception as e:
                print("sezkp thread allgather failed:", repr(e))
                return -1

        return ALLGATHER_FN(_cb)

    def dev_callback(self, rank: int):
        """Device-side all-gather among the threads (one GPU): every rank publishes its send pointer, all wait, every
        rank copies the `world` pieces into its own receive buffer on the library's stream and synchronises it."""
        import torch

        if not hasattr(self, "dev_slots"):
            self.dev_slots = [0] * self.world

        def _cb(user, send, nbytes, recv, stream):
            try:
                n = int(nbytes)
                st = torch.cuda.ExternalStream(int(stream or 0))
                st.synchronize()  # the send buffer is final before the others read it
                self.dev_slots[rank] = int(send)
                self.barrier.wait(timeout=120)
                with torch.cuda.stream(st):
                    dst = torch.as_tensor(_DevBytes(recv, n * self.world), device="cuda")
                    for r in range(self.world):
                        dst[r * n:(r + 1) * n].copy_(torch.as_tensor(_DevBytes(self.dev_slots[r], n), device="cuda"))
                st.synchronize()
                self.barrier.wait(timeout=120)  # nobody reuses its send buffer before everyone has copied
                return 0
            except Exception as e:
                print("sezkp thread device allgather failed:", repr(e))
                return -1

        return ALLGATHER_DEV_FN(_cb)

    def run(self, fn: Callable[[int], object]) -> list:
        """run fn(rank) on `world` threads, return the results in rank order (exceptions re-raised)"""
        out: list = [None] * self.world
        err: list = [None] * self.world

        def _t(r):
            try:
                out[r] = fn(r)
            except BaseException as e:
                err[r] = e
                self.barrier.abort()

        ts = [threading.Thread(target=_t, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out
