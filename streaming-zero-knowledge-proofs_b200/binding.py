"""ctypes binding of libsezkp_cuda.so (include/sezkp_cuda.h).

There is no CPU fallback: a missing library raises ImportError-like :class:`SezkpCudaError` at load time and a
missing GPU raises at context creation (``SEZKP_CUDA_ENODEV``).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Optional, Sequence

import numpy as np

from .trace import CompactTrace, TraceDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEZKP_CUDA_LIB") or os.path.join(_HERE, "libsezkp_cuda.so")  # override: A/B builds of the same ABI
P = 0xFFFFFFFF00000001

EXPORTS = [
    "sezkp_cuda_abi_version", "sezkp_cuda_create", "sezkp_cuda_destroy", "sezkp_cuda_last_error", "sezkp_cuda_set_stream",
    "sezkp_cuda_synchronize", "sezkp_cuda_set_option", "sezkp_cuda_launch_count", "sezkp_cuda_get_timings", "sezkp_cuda_get_timings_gpu",
    "sezkp_ntt_batch", "sezkp_ntt_batch_dev", "sezkp_coset_lde_batch", "sezkp_coset_lde_batch_dev",
    "sezkp_lde_from_evals_batch", "sezkp_lde_from_evals_batch_dev", "sezkp_deep_lde", "sezkp_deep_lde_dev",
    "sezkp_leaf_hash", "sezkp_merkle_root", "sezkp_column_commit_batch", "sezkp_column_commit_batch_dev",
    "sezkp_lde_commit_batch", "sezkp_lde_commit_batch_dev", "sezkp_column_open", "sezkp_verify_openings", "sezkp_tree_free", "sezkp_fri_commit", "sezkp_fri_commit_dev", "sezkp_fri_open", "sezkp_fri_free",
    "sezkp_trace_columns", "sezkp_compose_base", "sezkp_stark_v1_prove", "sezkp_stark_v1_begin", "sezkp_stark_v1_ingest",
    "sezkp_stark_v1_finish", "sezkp_stark_v1_abort", "sezkp_stark_v1_proof_bound", "sezkp_trace_upload", "sezkp_trace_free",
    "sezkp_stark_v1_prove_resident", "sezkp_stark_v1_prove_sharded", "sezkp_stark_v1_prove_resident_sharded",
    "sezkp_cuda_set_allgather_dev", "sezkp_cuda_create_multi", "sezkp_cuda_group_size", "sezkp_cuda_device_count", "sezkp_columns_upload", "sezkp_columns_synth",
    "sezkp_columns_free", "sezkp_lde_commit_fri", "sezkp_jsonl_parse", "sezkp_jsonl_write_file", "sezkp_jsonl_free", "sezkp_jsonl_last_error", "sezkp_stark_v1_ingest_jsonl", "sezkp_stark_v1_prove_jsonl_file",
]

ERR_NAMES = {0: "OK", -1: "EINVAL", -2: "ENOMEM", -3: "ECUDA", -4: "ENODEV", -5: "ERANGE", -6: "ESTATE", -7: "ECOMM"}


class SezkpCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sezkp_cuda {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """Load libsezkp_cuda.so; fails loudly when it has not been built (``python __graft_entry__.py`` builds it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SezkpCudaError(-4, f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                                     "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.sezkp_cuda_abi_version.restype = C.c_uint32
        lib.sezkp_cuda_last_error.restype = C.c_char_p
        lib.sezkp_cuda_last_error.argtypes = [C.c_void_p]
        lib.sezkp_cuda_launch_count.restype = C.c_uint64
        lib.sezkp_cuda_launch_count.argtypes = [C.c_void_p, C.c_int]
        lib.sezkp_cuda_destroy.argtypes = [C.c_void_p]
        lib.sezkp_tree_free.argtypes = [C.c_void_p, C.c_void_p]
        lib.sezkp_fri_free.argtypes = [C.c_void_p, C.c_void_p]
        lib.sezkp_stark_v1_abort.argtypes = [C.c_void_p, C.c_void_p]
        lib.sezkp_trace_free.argtypes = [C.c_void_p, C.c_void_p]
        lib.sezkp_columns_free.argtypes = [C.c_void_p, C.c_void_p]
        lib.sezkp_cuda_group_size.argtypes = [C.c_void_p]
        lib.sezkp_stark_v1_proof_bound.restype = C.c_size_t
        lib.sezkp_stark_v1_proof_bound.argtypes = [C.c_uint64, C.c_uint32]
        lib.sezkp_jsonl_last_error.restype = C.c_char_p
        lib.sezkp_jsonl_free.argtypes = [C.c_void_p]
        lib.sezkp_jsonl_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.sezkp_stark_v1_ingest_jsonl.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        lib.sezkp_stark_v1_prove_jsonl_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_size_t, C.c_uint64,
                                                        C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = lib
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _vp(x) -> C.c_void_p:
    """device pointer (int / torch tensor) -> c_void_p"""
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


class Context:
    """One GPU, one context (one process per GPU).  Mirrors ``sezkp_ctx``."""

    def __init__(self, device: int = -1, devices: Optional[Sequence[int]] = None):
        """device: one GPU (sezkp_cuda_create).  devices=[...]: ONE context over several GPUs of this box, driven from this
        single process (sezkp_cuda_create_multi): prove / lde_commit / lde_commit_fri then shard over all of them."""
        self.lib = load_library()
        h = C.c_void_p()
        if devices is not None:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.sezkp_cuda_create_multi(ids, C.c_int(len(devices)), C.byref(h))
        else:
            rc = self.lib.sezkp_cuda_create(C.c_int(device), C.byref(h))
        if rc != 0:
            raise SezkpCudaError(rc, self.lib.sezkp_cuda_last_error(None).decode())
        self.h = h

    @property
    def n_gpus(self) -> int:
        return int(self.lib.sezkp_cuda_group_size(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.sezkp_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise SezkpCudaError(rc, self.lib.sezkp_cuda_last_error(self.h).decode())

    # ---- plumbing ----
    def set_stream(self, cuda_stream: Optional[int]):
        """adopt a caller stream handle (0 = legacy default stream); None -> private stream"""
        if cuda_stream is None:
            self._ck(self.lib.sezkp_cuda_set_stream(self.h, None, C.c_int(1)))
        else:
            self._ck(self.lib.sezkp_cuda_set_stream(self.h, C.c_void_p(cuda_stream), C.c_int(0)))

    def synchronize(self):
        self._ck(self.lib.sezkp_cuda_synchronize(self.h))

    def set_option(self, name: str, value: int):
        self._ck(self.lib.sezkp_cuda_set_option(self.h, name.encode(), C.c_int64(value)))

    def tab_stats(self) -> tuple:
        """(columns served from subtree tables, chunks redone by the generic kernel) of the last value-aware commit"""
        out = (C.c_ulonglong * 2)()
        self.lib.sezkp_debug_tab_stats(self.h, out)
        return int(out[0]), int(out[1])

    def launch_count(self, reset=False) -> int:
        return int(self.lib.sezkp_cuda_launch_count(self.h, C.c_int(int(reset))))

    def timings(self) -> dict:
        buf = C.create_string_buffer(4096)
        self._ck(self.lib.sezkp_cuda_get_timings(self.h, buf, C.c_size_t(4096)))
        return json.loads(buf.value.decode())

    def timings_gpu(self, rank: int) -> dict:
        buf = C.create_string_buffer(4096)
        self._ck(self.lib.sezkp_cuda_get_timings_gpu(self.h, C.c_int(rank), buf, C.c_size_t(4096)))
        return json.loads(buf.value.decode())

    # ---- NTT / LDE (host buffers) ----
    def ntt(self, data, inverse=False) -> np.ndarray:
        a = np.array(data, dtype=np.uint64, copy=True, order="C")
        a2 = a.reshape(-1, a.shape[-1])
        n = a2.shape[1]
        if n & (n - 1):
            raise SezkpCudaError(-1, "NTT size must be a power of two")
        self._ck(self.lib.sezkp_ntt_batch(self.h, _p(a2), C.c_int(n.bit_length() - 1), C.c_int(a2.shape[0]), C.c_int(int(inverse))))
        return a

    def coset_lde(self, coeffs, log_blow: int, shift: int) -> np.ndarray:
        a = np.ascontiguousarray(coeffs, np.uint64)
        a2 = a.reshape(-1, a.shape[-1])
        n = a2.shape[1]
        out = np.empty((a2.shape[0], n << log_blow), np.uint64)
        self._ck(self.lib.sezkp_coset_lde_batch(self.h, _p(a2), C.c_int(n.bit_length() - 1), C.c_int(log_blow), C.c_uint64(shift),
                                                C.c_int(a2.shape[0]), _p(out)))
        return out if a.ndim == 2 else out[0]

    def lde_from_evals(self, evals, log_blow: int, shift: int) -> np.ndarray:
        a = np.ascontiguousarray(evals, np.uint64)
        a2 = a.reshape(-1, a.shape[-1])
        n = a2.shape[1]
        out = np.empty((a2.shape[0], n << log_blow), np.uint64)
        self._ck(self.lib.sezkp_lde_from_evals_batch(self.h, _p(a2), C.c_int(n.bit_length() - 1), C.c_int(log_blow),
                                                     C.c_uint64(shift), C.c_int(a2.shape[0]), _p(out)))
        return out if a.ndim == 2 else out[0]

    def deep_lde(self, base_evals, log_blow: int, shift: int, z: int) -> np.ndarray:
        a = np.ascontiguousarray(base_evals, np.uint64)
        out = np.empty(a.size << log_blow, np.uint64)
        self._ck(self.lib.sezkp_deep_lde(self.h, _p(a), C.c_int(a.size.bit_length() - 1), C.c_int(log_blow), C.c_uint64(shift),
                                         C.c_uint64(z), _p(out)))
        return out

    # ---- device-pointer variants (torch tensors or raw ints) ----
    def ntt_dev(self, data_dev, log_n: int, cols: int, inverse=False):
        self._ck(self.lib.sezkp_ntt_batch_dev(self.h, _vp(data_dev), C.c_int(log_n), C.c_int(cols), C.c_int(int(inverse))))

    def coset_lde_dev(self, coeffs_dev, log_n, log_blow, shift, cols, out_dev):
        self._ck(self.lib.sezkp_coset_lde_batch_dev(self.h, _vp(coeffs_dev), C.c_int(log_n), C.c_int(log_blow), C.c_uint64(shift),
                                                    C.c_int(cols), _vp(out_dev)))

    def lde_from_evals_dev(self, evals_dev, log_n, log_blow, shift, cols, out_dev):
        self._ck(self.lib.sezkp_lde_from_evals_batch_dev(self.h, _vp(evals_dev), C.c_int(log_n), C.c_int(log_blow),
                                                         C.c_uint64(shift), C.c_int(cols), _vp(out_dev)))

    def deep_lde_dev(self, base_dev, log_n, log_blow, shift, z, out_dev):
        self._ck(self.lib.sezkp_deep_lde_dev(self.h, _vp(base_dev), C.c_int(log_n), C.c_int(log_blow), C.c_uint64(shift),
                                             C.c_uint64(z), _vp(out_dev)))

    # ---- hashing / Merkle ----
    def leaf_hash(self, vals, label: Optional[str] = None) -> np.ndarray:
        v = np.ascontiguousarray(vals, np.uint64)
        out = np.empty((v.size, 32), np.uint8)
        self._ck(self.lib.sezkp_leaf_hash(self.h, _p(v), C.c_size_t(v.size), label.encode() if label is not None else None, _p(out)))
        return out

    def merkle_root(self, leaves) -> bytes:
        l = np.ascontiguousarray(leaves, np.uint8).reshape(-1, 32)
        out = C.create_string_buffer(32)
        self._ck(self.lib.sezkp_merkle_root(self.h, _p(l), C.c_size_t(l.shape[0]), out))
        return out.raw

    def column_commit(self, cols, labels: Optional[Sequence[str]], chunk_log2=10, keep=False, dev=False, n=None, c=None):
        """labels=None commits unlabeled leaves (FRI-layer style); then `c` gives the column count for device input."""
        if dev:
            c = len(labels) if labels is not None else c
            ptr = _vp(cols)
        else:
            a = np.ascontiguousarray(cols, np.uint64)
            c, n = a.shape
            ptr = _p(a)
        arr = (C.c_char_p * c)(*[l.encode() for l in labels]) if labels is not None else None
        roots = np.empty((c, 32), np.uint8)
        tree = C.c_void_p()
        fn = self.lib.sezkp_column_commit_batch_dev if dev else self.lib.sezkp_column_commit_batch
        self._ck(fn(self.h, ptr, arr, C.c_int(c), C.c_size_t(n), C.c_int(chunk_log2), _p(roots), C.byref(tree) if keep else None))
        return (roots, ColumnTree(self, tree)) if keep else roots

    def verify_openings(self, col_roots, labels: Optional[Sequence[str]], col_idx, values, idx_in, chunk_idx=None, chunk_roots=None,
                        path_in=None, path_to=None) -> np.ndarray:
        """Batched verify_chunked_open / MerkleTree::verify: -> bool array [k].  path_in [k][din][32], path_to [k][dout][32]."""
        roots = np.ascontiguousarray(col_roots, np.uint8).reshape(-1, 32)
        c = roots.shape[0]
        v = np.ascontiguousarray(values, np.uint64)
        k = v.size
        ci = np.ascontiguousarray(col_idx, np.uint32) if col_idx is not None else None
        ii = np.ascontiguousarray(idx_in, np.uint64)
        io = np.ascontiguousarray(chunk_idx, np.uint64) if chunk_idx is not None else None
        cr = np.ascontiguousarray(chunk_roots, np.uint8).reshape(k, 32) if chunk_roots is not None else None
        pi = np.ascontiguousarray(path_in, np.uint8).reshape(k, -1, 32) if path_in is not None and np.size(path_in) else None
        pt = np.ascontiguousarray(path_to, np.uint8).reshape(k, -1, 32) if path_to is not None and np.size(path_to) else None
        din = pi.shape[1] if pi is not None else 0
        dout = pt.shape[1] if pt is not None else 0
        arr = (C.c_char_p * c)(*[l.encode() for l in labels]) if labels is not None else None
        ok = np.zeros(k, np.uint8)
        self._ck(self.lib.sezkp_verify_openings(self.h, _p(roots), arr, C.c_int(c), _p(ci) if ci is not None else None, _p(v), _p(ii),
                                                _p(io) if io is not None else None, _p(cr) if cr is not None else None,
                                                _p(pi) if pi is not None else None, C.c_int(din), _p(pt) if pt is not None else None,
                                                C.c_int(dout), C.c_size_t(k), _p(ok)))
        return ok.astype(bool)

    def lde_commit(self, evals, labels: Sequence[str], log_blow: int, shift: int = 3, chunk_log2=10, dev=False, log_n=None) -> np.ndarray:
        """iNTT -> coset LDE -> labeled leaf hash -> root per column (extended columns never all resident)."""
        c = len(labels)
        if dev:
            ptr = _vp(evals)
        else:
            a = np.ascontiguousarray(evals, np.uint64)
            assert a.shape[0] == c
            log_n = a.shape[1].bit_length() - 1
            ptr = _p(a)
        arr = (C.c_char_p * c)(*[l.encode() for l in labels])
        roots = np.empty((c, 32), np.uint8)
        fn = self.lib.sezkp_lde_commit_batch_dev if dev else self.lib.sezkp_lde_commit_batch
        self._ck(fn(self.h, ptr, arr, C.c_int(c), C.c_int(log_n), C.c_int(log_blow), C.c_uint64(shift), C.c_int(chunk_log2), _p(roots)))
        return roots

    # ---- config 4: resident column sets, LDE + commit + FRI ----
    def columns_upload(self, evals) -> "ColumnSet":
        a = np.ascontiguousarray(evals, np.uint64)
        c, n = a.shape
        h = C.c_void_p()
        self._ck(self.lib.sezkp_columns_upload(self.h, _p(a), C.c_int(c), C.c_int(n.bit_length() - 1), C.byref(h)))
        return ColumnSet(self, h, c, n.bit_length() - 1)

    def columns_synth(self, c: int, log_n: int, seed: int = 0x5EED) -> "ColumnSet":
        """SURVEY 8d config-4 generator on the device: value(c, i) = splitmix64 step of seed ^ c<<40 ^ i, mod p."""
        h = C.c_void_p()
        self._ck(self.lib.sezkp_columns_synth(self.h, C.c_uint64(seed), C.c_int(c), C.c_int(log_n), C.byref(h)))
        return ColumnSet(self, h, c, log_n)

    def lde_commit_fri(self, cols: "ColumnSet", labels: Optional[Sequence[str]] = None, log_blow: int = 3, shift: int = 3, chunk_log2=10):
        """-> (col_roots [c][32], fri_roots [log_N+1][32], final value)"""
        labels = list(labels) if labels is not None else [f"c_{k}" for k in range(cols.c)]
        assert len(labels) == cols.c
        arr = (C.c_char_p * cols.c)(*[l.encode() for l in labels])
        lN = cols.log_n + log_blow
        cr = np.empty((cols.c, 32), np.uint8)
        fr = np.empty((lN + 1, 32), np.uint8)
        fin = C.c_uint64(0)
        self._ck(self.lib.sezkp_lde_commit_fri(self.h, cols.h, arr, C.c_int(log_blow), C.c_uint64(shift), C.c_int(chunk_log2), _p(cr), _p(fr),
                                               C.byref(fin)))
        return cr, fr, fin.value

    def fri_commit(self, layer0, betas, keep=False, dev=False, log_N=None):
        b = np.ascontiguousarray(betas, np.uint64)
        if dev:
            ptr = _vp(layer0)
        else:
            a = np.ascontiguousarray(layer0, np.uint64)
            log_N = a.size.bit_length() - 1
            ptr = _p(a)
        roots = np.empty((log_N + 1, 32), np.uint8)
        fin = C.c_uint64(0)
        h = C.c_void_p()
        fn = self.lib.sezkp_fri_commit_dev if dev else self.lib.sezkp_fri_commit
        self._ck(fn(self.h, ptr, C.c_int(log_N), _p(b), _p(roots), C.byref(fin), C.byref(h) if keep else None))
        return (roots, fin.value, FriHandle(self, h, log_N)) if keep else (roots, fin.value)

    # ---- feeder ----
    def trace_columns(self, ct: CompactTrace) -> np.ndarray:
        d = ct.as_desc()
        out = np.empty((3 + 7 * ct.tau, ct.n_rows), np.uint64)
        self._ck(self.lib.sezkp_trace_columns(self.h, C.byref(d), _p(out)))
        return out

    def compose_base(self, ct: CompactTrace, alphas8, mask_coeffs) -> np.ndarray:
        d = ct.as_desc()
        a = np.ascontiguousarray(alphas8, np.uint64)
        m = np.ascontiguousarray(mask_coeffs, np.uint64)
        out = np.empty(ct.n_rows, np.uint64)
        self._ck(self.lib.sezkp_compose_base(self.h, C.byref(d), _p(a), _p(m), C.c_size_t(m.size), _p(out)))
        return out

    # ---- prover ----
    def prove_v1(self, ct: CompactTrace, manifest_root: bytes, buf: Optional[np.ndarray] = None, view: bool = False):
        """Proof bytes.  `buf` (e.g. pinned) receives the proof; view=True returns the filled slice of `buf` instead of a
        copy (what a native caller of the C ABI gets)."""
        if len(manifest_root) != 32:
            raise SezkpCudaError(-1, "manifest_root must be 32 bytes")
        d = ct.as_desc()
        n = C.c_size_t(0)
        if buf is None:
            buf = np.empty(proof_size_bound(ct.n_rows, ct.tau), np.uint8)
        self._ck(self.lib.sezkp_stark_v1_prove(self.h, C.byref(d), manifest_root, _p(buf), C.c_size_t(buf.size), C.byref(n)))
        return buf[: n.value] if view else buf[: n.value].tobytes()

    def prove_v1_sharded(self, ct: CompactTrace, manifest_root: bytes, rank: int, world: int, allgather_cb,
                         buf: Optional[np.ndarray] = None) -> bytes:
        """Column-sharded prove: `allgather_cb` is a parallel.ALLGATHER_FN (see parallel.dist_allgather_callback)."""
        d = ct.as_desc()
        n = C.c_size_t(0)
        if buf is None:
            buf = np.empty(proof_size_bound(ct.n_rows, ct.tau), np.uint8)
        self._ck(self.lib.sezkp_stark_v1_prove_sharded(self.h, C.byref(d), manifest_root, C.c_int(rank), C.c_int(world), allgather_cb,
                                                        None, _p(buf), C.c_size_t(buf.size), C.byref(n)))
        return buf[: n.value].tobytes()

    def set_allgather_dev(self, cb) -> None:
        """Register (or clear with None) the device-side collective of the sharded prover (parallel.ALLGATHER_DEV_FN)."""
        self._allgather_dev_cb = cb  # keep the ctypes thunk alive
        self._ck(self.lib.sezkp_cuda_set_allgather_dev(self.h, cb if cb is not None else C.cast(None, C.c_void_p), None))

    def prove_v1_resident_sharded(self, rt: "ResidentTrace", manifest_root: bytes, rank: int, world: int, allgather_cb,
                                  buf: Optional[np.ndarray] = None) -> bytes:
        """Sharded prove over a trace that is already resident on this rank's GPU (no per-rank H2D of the whole trace)."""
        n = C.c_size_t(0)
        if buf is None:
            buf = np.empty(proof_size_bound(rt.n_rows, rt.tau), np.uint8)
        self._ck(self.lib.sezkp_stark_v1_prove_resident_sharded(self.h, rt.h, manifest_root, C.c_int(rank), C.c_int(world), allgather_cb,
                                                                 None, _p(buf), C.c_size_t(buf.size), C.byref(n)))
        return buf[: n.value].tobytes()

    def upload_trace(self, ct: CompactTrace) -> "ResidentTrace":
        d = ct.as_desc()
        h = C.c_void_p()
        self._ck(self.lib.sezkp_trace_upload(self.h, C.byref(d), C.byref(h)))
        return ResidentTrace(self, h, ct.n_rows, ct.tau)

    def prove_v1_resident(self, rt: "ResidentTrace", manifest_root: bytes, buf: Optional[np.ndarray] = None, view: bool = False):
        n = C.c_size_t(0)
        if buf is None:
            buf = np.empty(proof_size_bound(rt.n_rows, rt.tau), np.uint8)
        self._ck(self.lib.sezkp_stark_v1_prove_resident(self.h, rt.h, manifest_root, _p(buf), C.c_size_t(buf.size), C.byref(n)))
        return buf[: n.value] if view else buf[: n.value].tobytes()

    def prove_v1_stream(self, blocks, manifest_root: bytes, tau: Optional[int] = None, expected_rows: int = 0) -> bytes:
        """begin_stream / ingest_block / finish_stream (reference sezkp-core/src/prover.rs:21-33).  `blocks` is any
        iterable of CompactTrace pieces (one or more blocks each), e.g. a generator parsing JSONL lines."""
        it = iter(blocks)
        first = next(it)
        tau = tau or first.tau
        st = C.c_void_p()
        self._ck(self.lib.sezkp_stark_v1_begin(self.h, C.c_uint32(tau), manifest_root, C.c_uint64(expected_rows), C.byref(st)))
        try:
            rows = 0
            import itertools
            for b in itertools.chain([first], it):
                d = b.as_desc(packed=False)  # the staging ring holds the plain arrays
                self._ck(self.lib.sezkp_stark_v1_ingest(self.h, st, C.byref(d)))
                rows += b.n_rows
            buf = np.empty(proof_size_bound(rows, tau), np.uint8)
            n = C.c_size_t(0)
            self._ck(self.lib.sezkp_stark_v1_finish(self.h, st, _p(buf), C.c_size_t(buf.size), C.byref(n)))
            st = None
            return buf[: n.value].tobytes()
        finally:
            if st is not None and st.value:
                self.lib.sezkp_stark_v1_abort(self.h, st)


    def prove_v1_jsonl_file(self, path: str, manifest_root: bytes, n_rows_bound: int, tau_bound: int = 8, threads: int = 0,
                            chunk_bytes: int = 0, expected_rows: int = 0) -> bytes:
        """StreamingProver loop over a .jsonl file with the native multi-threaded parser (sezkp_stark_v1_prove_jsonl_file).
        n_rows_bound / tau_bound only size the proof buffer."""
        buf = np.empty(proof_size_bound(n_rows_bound, tau_bound), np.uint8)
        n = C.c_size_t(0)
        self._ck(self.lib.sezkp_stark_v1_prove_jsonl_file(self.h, path.encode(), manifest_root, C.c_int(threads), C.c_size_t(chunk_bytes),
                                                          C.c_uint64(expected_rows), _p(buf), C.c_size_t(buf.size), C.byref(n)))
        return buf[: n.value].tobytes()

    def prove_v1_jsonl_text(self, pieces, manifest_root: bytes, tau: int, threads: int = 0, expected_rows: int = 0) -> bytes:
        """begin_stream, then sezkp_stark_v1_ingest_jsonl for every piece of text (whole lines), then finish_stream."""
        st = C.c_void_p()
        self._ck(self.lib.sezkp_stark_v1_begin(self.h, C.c_uint32(tau), manifest_root, C.c_uint64(expected_rows), C.byref(st)))
        try:
            rows = 0
            for text in pieces:
                nb, nr = C.c_uint64(0), C.c_uint64(0)
                self._ck(self.lib.sezkp_stark_v1_ingest_jsonl(self.h, st, text, C.c_size_t(len(text)), C.c_int(threads), C.byref(nb), C.byref(nr)))
                rows += nr.value
            buf = np.empty(proof_size_bound(rows, tau), np.uint8)
            n = C.c_size_t(0)
            self._ck(self.lib.sezkp_stark_v1_finish(self.h, st, _p(buf), C.c_size_t(buf.size), C.byref(n)))
            st = None
            return buf[: n.value].tobytes()
        finally:
            if st is not None and st.value:
                self.lib.sezkp_stark_v1_abort(self.h, st)


BLOCK_SCALARS_DTYPE = np.dtype({"names": ["step_lo", "step_hi", "in_head_in", "in_head_out", "block_id", "version", "ctrl_in", "ctrl_out"],
                                "formats": ["<u8", "<u8", "<i8", "<i8", "<u4", "<u2", "<u2", "<u2"],
                                "offsets": [0, 8, 16, 24, 32, 36, 38, 40], "itemsize": 48})  # sezkp_block_scalars


def parse_jsonl(text: bytes, threads: int = 0) -> CompactTrace:
    """Native JSONL parser (sezkp_jsonl_parse): bytes of whole lines -> CompactTrace (arrays copied out of the handle).
    Needs the library but no GPU."""
    lib = load_library()
    h, d, sc = C.c_void_p(), TraceDesc(), C.c_void_p()
    rc = lib.sezkp_jsonl_parse(text, C.c_size_t(len(text)), C.c_int(threads), C.byref(h), C.byref(d), C.byref(sc))
    if rc != 0:
        raise SezkpCudaError(rc, lib.sezkp_jsonl_last_error().decode())
    try:
        nb, n, tau = int(d.n_blocks), int(d.n_rows), int(d.tau)

        def arr(ptr, dt, shape):
            cnt = int(np.prod(shape))
            if cnt == 0:
                return np.zeros(shape, dt)
            return np.frombuffer((C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt).reshape(shape).copy()

        rec = arr(sc.value, BLOCK_SCALARS_DTYPE, (nb,))
        return CompactTrace(
            tau=tau, block_len=arr(d.block_len, np.uint64, (nb,)), win_left=arr(d.win_left, np.int64, (nb, tau)),
            win_right=arr(d.win_right, np.int64, (nb, tau)), head_in_off=arr(d.head_in_off, np.uint32, (nb, tau)),
            head_out_off=arr(d.head_out_off, np.uint32, (nb, tau)), input_mv=arr(d.input_mv, np.int8, (n,)),
            mv=arr(d.mv, np.int8, (n, tau)), write_flag=arr(d.write_flag, np.uint8, (n, tau)),
            write_sym=arr(d.write_sym, np.uint16, (n, tau)),
            version=rec["version"].copy(), block_id=rec["block_id"].copy(), step_lo=rec["step_lo"].copy(),
            step_hi=rec["step_hi"].copy(), ctrl_in=rec["ctrl_in"].copy(), ctrl_out=rec["ctrl_out"].copy(),
            in_head_in=rec["in_head_in"].copy(), in_head_out=rec["in_head_out"].copy())
    finally:
        lib.sezkp_jsonl_free(h)


def write_jsonl_native(path: str, ct: CompactTrace, threads: int = 0) -> int:
    """Native multi-threaded JSONL writer (sezkp_jsonl_write_file): same bytes as io_jsonl.write_jsonl, ~100x faster."""
    lib = load_library()
    d = ct.as_desc(packed=False)
    sc = None
    if ct.version is not None:
        rec = np.zeros(ct.n_blocks, BLOCK_SCALARS_DTYPE)
        for name in ("step_lo", "step_hi", "in_head_in", "in_head_out", "block_id", "version", "ctrl_in", "ctrl_out"):
            rec[name] = getattr(ct, name)
        sc = rec
    n = C.c_uint64(0)
    rc = lib.sezkp_jsonl_write_file(path.encode(), C.byref(d), _p(sc) if sc is not None else None, C.c_int(threads), C.byref(n))
    if rc != 0:
        raise SezkpCudaError(rc, lib.sezkp_jsonl_last_error().decode())
    return int(n.value)


def proof_size_bound(n_rows: int, tau: int) -> int:
    ln = max(1, int(n_rows).bit_length() - 1)
    lN = ln + 3
    openings = 30 * (9 * tau + 3) * (8 + 24 + 32 + 16 + 32 * ln)
    fri = 30 * (16 + 8 * (lN + 1) + lN * 2 * (8 + 8 + 32 * lN))
    return 4096 + (3 + 7 * tau) * 64 + openings + fri + 32 * (lN + 1)


class ResidentTrace:
    def __init__(self, ctx: Context, h, n_rows, tau):
        self.ctx, self.h, self.n_rows, self.tau = ctx, h, n_rows, tau

    def free(self):
        if self.h:
            self.ctx.lib.sezkp_trace_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ColumnSet:
    """base-domain columns resident in HBM (column c on GPU c % n_gpus of the context)"""

    def __init__(self, ctx: Context, h, c, log_n):
        self.ctx, self.h, self.c, self.log_n = ctx, h, c, log_n

    def free(self):
        if self.h:
            self.ctx.lib.sezkp_columns_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ColumnTree:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def open(self, col_idx, row_idx):
        c = np.ascontiguousarray(col_idx, np.uint32)
        r = np.ascontiguousarray(row_idx, np.uint64)
        k = c.size
        din, dout = C.c_int(0), C.c_int(0)
        lib = self.ctx.lib
        self.ctx._ck(lib.sezkp_column_open(self.ctx.h, self.h, None, None, C.c_size_t(0), None, None, None, None, C.byref(din), C.byref(dout)))
        vals = np.empty(k, np.uint64)
        cr = np.empty((k, 32), np.uint8)
        pin = np.zeros((k, max(din.value, 1), 32), np.uint8)
        pto = np.zeros((k, max(dout.value, 1), 32), np.uint8)
        self.ctx._ck(lib.sezkp_column_open(self.ctx.h, self.h, _p(c), _p(r), C.c_size_t(k), _p(vals), _p(cr), _p(pin), _p(pto),
                                           C.byref(din), C.byref(dout)))
        return vals, cr, pin[:, : din.value], pto[:, : dout.value]

    def free(self):
        if self.h:
            self.ctx.lib.sezkp_tree_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FriHandle:
    def __init__(self, ctx: Context, h, log_N):
        self.ctx, self.h, self.log_N = ctx, h, log_N

    def open(self, idx0):
        q = np.ascontiguousarray(idx0, np.uint64)
        k, L = q.size, self.log_N
        pos = np.empty((k, L + 1), np.uint64)
        vals = np.empty((k, L, 2), np.uint64)
        paths = np.zeros((k, L, 2, L, 32), np.uint8)
        self.ctx._ck(self.ctx.lib.sezkp_fri_open(self.ctx.h, self.h, _p(q), C.c_size_t(k), _p(pos), _p(vals), _p(paths)))
        return pos, vals, paths

    def free(self):
        if self.h:
            self.ctx.lib.sezkp_fri_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
