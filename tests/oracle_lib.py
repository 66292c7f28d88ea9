"""ctypes front-end for the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
_LIB = None

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build():
    so = os.path.join(ODIR, "liboracle.so")
    srcs = [os.path.join(ODIR, f) for f in ("capi.cpp", "stark.hpp", "wide.hpp", "blake3_ref.hpp", "gl.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ODIR, "-s"])
    return so


class Taps(C.Structure):
    _fields_ = [("alphas", C.c_uint64 * 8), ("mask_coeffs", C.c_uint64 * 4), ("z", C.c_uint64),
                ("betas", C.c_uint64 * 64), ("rows", C.c_uint64 * 30), ("fri_rows", C.c_uint64 * 30)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_transcript_new.restype = C.c_void_p
        lib.oracle_transcript_new.argtypes = [C.c_char_p]
        lib.oracle_transcript_free.argtypes = [C.c_void_p]
        lib.oracle_transcript_absorb.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.oracle_transcript_challenge.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
        for f in ("oracle_gl_mul", "oracle_gl_pow"):
            getattr(lib, f).restype = C.c_uint64
            getattr(lib, f).argtypes = [C.c_uint64, C.c_uint64]
        lib.oracle_gl_inv.restype = C.c_uint64
        lib.oracle_gl_inv.argtypes = [C.c_uint64]
        lib.oracle_gl_from_i64.restype = C.c_uint64
        lib.oracle_gl_from_i64.argtypes = [C.c_int64]
        lib.oracle_gl_root_2exp.restype = C.c_uint64
        lib.oracle_gl_root_2exp.argtypes = [C.c_uint]

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib.oracle_last_error().decode())

    # ---- hashing ----
    def blake3(self, data: bytes, out_len=32) -> bytes:
        out = C.create_string_buffer(out_len)
        self._ck(self.lib.oracle_blake3(data, C.c_size_t(len(data)), out, C.c_size_t(out_len)))
        return out.raw

    def transcript(self, domain: str):
        return _Transcript(self, domain)

    def leaf_hash(self, vals, label=None) -> np.ndarray:
        vals = np.ascontiguousarray(vals, np.uint64)
        out = np.empty((vals.size, 32), np.uint8)
        self._ck(self.lib.oracle_leaf_hash(vals.ctypes.data_as(C.c_void_p), C.c_size_t(vals.size),
                                           label.encode() if label is not None else None, out.ctypes.data_as(C.c_void_p)))
        return out

    def node_hash(self, l: bytes, r: bytes) -> bytes:
        out = C.create_string_buffer(32)
        self._ck(self.lib.oracle_node_hash(l, r, out))
        return out.raw

    def merkle_root(self, leaves) -> bytes:
        leaves = np.ascontiguousarray(leaves, np.uint8).reshape(-1, 32)
        out = C.create_string_buffer(32)
        self._ck(self.lib.oracle_merkle_root(leaves.ctypes.data_as(C.c_void_p), C.c_size_t(leaves.shape[0]), out))
        return out.raw

    def merkle_open(self, leaves, idx):
        leaves = np.ascontiguousarray(leaves, np.uint8).reshape(-1, 32)
        sibs = np.zeros((64, 32), np.uint8)
        n = C.c_size_t(0)
        self._ck(self.lib.oracle_merkle_open(leaves.ctypes.data_as(C.c_void_p), C.c_size_t(leaves.shape[0]), C.c_size_t(idx),
                                             sibs.ctypes.data_as(C.c_void_p), C.byref(n)))
        return sibs[: n.value].copy()

    def streaming_layer_root(self, vals) -> bytes:
        vals = np.ascontiguousarray(vals, np.uint64)
        out = C.create_string_buffer(32)
        self._ck(self.lib.oracle_streaming_layer_root(vals.ctypes.data_as(C.c_void_p), C.c_size_t(vals.size), out))
        return out.raw

    def manifest_merkle_root(self, leaves, frontier=False) -> bytes:
        leaves = np.ascontiguousarray(leaves, np.uint8).reshape(-1, 32)
        out = C.create_string_buffer(32)
        self._ck(self.lib.oracle_manifest_merkle_root(leaves.ctypes.data_as(C.c_void_p), C.c_size_t(leaves.shape[0]),
                                                      C.c_int(int(frontier)), out))
        return out.raw

    def manifest_leaf_hash(self, ct, k) -> bytes:
        out = C.create_string_buffer(32)
        wl = np.ascontiguousarray(ct.win_left[k], np.int64)
        wr = np.ascontiguousarray(ct.win_right[k], np.int64)
        io = np.ascontiguousarray(ct.head_in_off[k], np.uint32)
        oo = np.ascontiguousarray(ct.head_out_off[k], np.uint32)
        self._ck(self.lib.oracle_manifest_leaf_hash(
            C.c_uint16(int(ct.version[k])), C.c_uint32(int(ct.block_id[k])), C.c_uint64(int(ct.step_lo[k])),
            C.c_uint64(int(ct.step_hi[k])), C.c_uint16(int(ct.ctrl_in[k])), C.c_uint16(int(ct.ctrl_out[k])),
            C.c_int64(int(ct.in_head_in[k])), C.c_int64(int(ct.in_head_out[k])), C.c_uint64(ct.tau),
            wl.ctypes.data_as(C.c_void_p), wr.ctypes.data_as(C.c_void_p), io.ctypes.data_as(C.c_void_p),
            oo.ctypes.data_as(C.c_void_p), C.c_uint64(int(ct.block_len[k])), out))
        return out.raw

    def manifest_root(self, ct) -> bytes:
        leaves = np.frombuffer(b"".join(self.manifest_leaf_hash(ct, k) for k in range(ct.n_blocks)), np.uint8)
        return self.manifest_merkle_root(leaves)

    # ---- field / NTT ----
    def ntt(self, data, inverse=False) -> np.ndarray:
        a = np.array(data, dtype=np.uint64, copy=True, order="C")
        a2 = a.reshape(-1, a.shape[-1])
        log_n = int(a2.shape[1]).bit_length() - 1
        self._ck(self.lib.oracle_ntt(a2.ctypes.data_as(C.c_void_p), C.c_int(log_n), C.c_size_t(a2.shape[0]), C.c_int(int(inverse))))
        return a

    def dft_naive(self, data) -> np.ndarray:
        a = np.ascontiguousarray(data, np.uint64)
        out = np.empty_like(a)
        self._ck(self.lib.oracle_dft_naive(a.ctypes.data_as(C.c_void_p), C.c_int(a.size.bit_length() - 1), out.ctypes.data_as(C.c_void_p)))
        return out

    def coset_eval(self, coeffs, k_log2, shift) -> np.ndarray:
        a = np.ascontiguousarray(coeffs, np.uint64)
        a2 = a.reshape(-1, a.shape[-1])
        out = np.empty((a2.shape[0], 1 << k_log2), np.uint64)
        self._ck(self.lib.oracle_coset_eval(a2.ctypes.data_as(C.c_void_p), C.c_size_t(a2.shape[1]), C.c_int(k_log2),
                                            C.c_uint64(shift), C.c_size_t(a2.shape[0]), out.ctypes.data_as(C.c_void_p)))
        return out if a.ndim == 2 else out[0]

    def lde_from_evals(self, evals, log_blow, shift) -> np.ndarray:
        a = np.ascontiguousarray(evals, np.uint64)
        a2 = a.reshape(-1, a.shape[-1])
        log_n = int(a2.shape[1]).bit_length() - 1
        out = np.empty((a2.shape[0], a2.shape[1] << log_blow), np.uint64)
        self._ck(self.lib.oracle_lde_from_evals(a2.ctypes.data_as(C.c_void_p), C.c_int(log_n), C.c_int(log_blow),
                                                C.c_uint64(shift), C.c_size_t(a2.shape[0]), out.ctypes.data_as(C.c_void_p)))
        return out if a.ndim == 2 else out[0]

    def deep_lde(self, base, log_blow, shift, z) -> np.ndarray:
        a = np.ascontiguousarray(base, np.uint64)
        log_n = a.size.bit_length() - 1
        out = np.empty(a.size << log_blow, np.uint64)
        self._ck(self.lib.oracle_deep_lde(a.ctypes.data_as(C.c_void_p), C.c_int(log_n), C.c_int(log_blow), C.c_uint64(shift),
                                          C.c_uint64(z), out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- columns / composition / FRI ----
    def column_commit(self, cols, labels, chunk_log2=10) -> np.ndarray:
        cols = np.ascontiguousarray(cols, np.uint64)
        c, n = cols.shape
        arr = (C.c_char_p * c)(*[l.encode() for l in labels])
        out = np.empty((c, 32), np.uint8)
        self._ck(self.lib.oracle_column_commit(cols.ctypes.data_as(C.c_void_p), arr, C.c_size_t(c), C.c_size_t(n),
                                               C.c_int(chunk_log2), out.ctypes.data_as(C.c_void_p)))
        return out

    def trace_columns(self, ct) -> np.ndarray:
        d = ct.as_desc(packed=False)  # the oracle reads the plain arrays
        out = np.empty((3 + 7 * ct.tau, ct.n_rows), np.uint64)
        self._ck(self.lib.oracle_trace_columns(C.byref(d), out.ctypes.data_as(C.c_void_p)))
        return out

    def compose_base(self, ct, alphas8, mask_coeffs) -> np.ndarray:
        d = ct.as_desc(packed=False)  # the oracle reads the plain arrays
        a = np.ascontiguousarray(alphas8, np.uint64)
        m = np.ascontiguousarray(mask_coeffs, np.uint64)
        out = np.empty(ct.n_rows, np.uint64)
        self._ck(self.lib.oracle_compose_base(C.byref(d), a.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p),
                                              C.c_size_t(m.size), out.ctypes.data_as(C.c_void_p)))
        return out

    def fri_commit(self, layer0, betas):
        l0 = np.ascontiguousarray(layer0, np.uint64)
        b = np.ascontiguousarray(betas, np.uint64)
        log_n = l0.size.bit_length() - 1
        assert b.size == log_n
        roots = np.empty((log_n + 1, 32), np.uint8)
        fin = C.c_uint64(0)
        self._ck(self.lib.oracle_fri_commit(l0.ctypes.data_as(C.c_void_p), C.c_int(log_n), b.ctypes.data_as(C.c_void_p),
                                            roots.ctypes.data_as(C.c_void_p), C.byref(fin)))
        return roots, fin.value

    # ---- config 4 (wide LDE + commit + FRI, oracle/wide.hpp) ----
    def wide_column(self, c, log_n) -> np.ndarray:
        out = np.empty(1 << log_n, np.uint64)
        self._ck(self.lib.oracle_wide_column(C.c_uint64(c), C.c_int(log_n), out.ctypes.data_as(C.c_void_p)))
        return out

    def lde_commit_root(self, evals, log_blow, shift, label) -> bytes:
        a = np.ascontiguousarray(evals, np.uint64)
        out = C.create_string_buffer(32)
        self._ck(self.lib.oracle_lde_commit_root(a.ctypes.data_as(C.c_void_p), C.c_int(a.size.bit_length() - 1), C.c_int(log_blow),
                                                 C.c_uint64(shift), label.encode(), out))
        return out.raw

    def wide_column_root(self, c, log_n, log_blow=3, shift=3) -> bytes:
        out = C.create_string_buffer(32)
        self._ck(self.lib.oracle_wide_column_root(C.c_uint64(c), C.c_int(log_n), C.c_int(log_blow), C.c_uint64(shift), out))
        return out.raw

    def wide_tail(self, evals, n_cols, log_n, col_roots, log_blow=3, shift=3):
        """evals [n_cols][n] or None (generator columns) -> dict(alphas, z, betas, fri_roots, final)"""
        lN = log_n + log_blow
        roots = np.ascontiguousarray(col_roots, np.uint8).reshape(n_cols, 32)
        al = np.empty(n_cols, np.uint64)
        be = np.empty(lN, np.uint64)
        fr = np.empty((lN + 1, 32), np.uint8)
        z, fin = C.c_uint64(0), C.c_uint64(0)
        ev = None
        if evals is not None:
            ev = np.ascontiguousarray(evals, np.uint64)
            assert ev.shape == (n_cols, 1 << log_n)
        self._ck(self.lib.oracle_wide_tail(ev.ctypes.data_as(C.c_void_p) if ev is not None else None, C.c_size_t(n_cols), C.c_int(log_n),
                                           C.c_int(log_blow), C.c_uint64(shift), roots.ctypes.data_as(C.c_void_p),
                                           al.ctypes.data_as(C.c_void_p), C.byref(z), be.ctypes.data_as(C.c_void_p),
                                           fr.ctypes.data_as(C.c_void_p), C.byref(fin)))
        return {"alphas": al, "z": z.value, "betas": be, "fri_roots": fr, "final": fin.value}

    # ---- prover / verifier ----
    def prove_v1(self, ct, manifest_root: bytes, faithful_cost=False, taps=False):
        d = ct.as_desc(packed=False)  # the oracle reads the plain arrays
        n = C.c_size_t(0)
        t = Taps()
        cap = 64 << 20
        buf = C.create_string_buffer(cap)
        self._ck(self.lib.oracle_prove_v1(C.byref(d), manifest_root, C.c_int(int(faithful_cost)), buf, C.c_size_t(cap),
                                          C.byref(n), C.byref(t)))
        proof = buf.raw[: n.value]
        return (proof, t) if taps else proof

    def verify_v1(self, proof: bytes, ct):
        """returns (accepted: bool, reason: str)"""
        d = ct.as_desc(packed=False)  # the oracle reads the plain arrays
        rc = self.lib.oracle_verify_v1(proof, C.c_size_t(len(proof)), C.byref(d))
        if rc == 0:
            return True, ""
        return False, self.lib.oracle_last_error().decode()

    def gl_mul(self, a, b):
        return self.lib.oracle_gl_mul(a, b)

    def gl_inv(self, a):
        return self.lib.oracle_gl_inv(a)

    def gl_pow(self, a, e):
        return self.lib.oracle_gl_pow(a, e)

    def gl_from_i64(self, x):
        return self.lib.oracle_gl_from_i64(x)

    def gl_root_2exp(self, k):
        return self.lib.oracle_gl_root_2exp(k)


class _Transcript:
    def __init__(self, o, domain):
        self.o = o
        self.h = C.c_void_p(o.lib.oracle_transcript_new(domain.encode()))

    def absorb(self, label, data: bytes):
        self.o._ck(self.o.lib.oracle_transcript_absorb(self.h, label.encode(), data, C.c_size_t(len(data))))

    def absorb_u64(self, label, x):
        self.absorb(label, int(x).to_bytes(8, "little"))

    def challenge(self, label, n) -> bytes:
        out = C.create_string_buffer(n)
        self.o._ck(self.o.lib.oracle_transcript_challenge(self.h, label.encode(), out, C.c_size_t(n)))
        return out.raw

    def __del__(self):
        try:
            self.o.lib.oracle_transcript_free(self.h)
        except Exception:
            pass


def load() -> Oracle:
    global _LIB
    if _LIB is None:
        _LIB = Oracle(C.CDLL(build()))
    return _LIB


# ---- deterministic input generators of the reference's tests / benches ----
P = 0xFFFFFFFF00000001


def det_vec(n, seed):
    """LCG vectors of crates/sezkp-ffts/tests/ntt_roundtrip.rs:14-27 == benches/ntt.rs:21-34."""
    A, Cc, M = 1664525, 1013904223, 1 << 32
    a = (A * seed + Cc) & 0xFFFFFFFFFFFFFFFF
    out = np.empty(n, np.uint64)
    for i in range(n):
        a = ((a * A + Cc) & 0xFFFFFFFFFFFFFFFF) % M
        out[i] = (a ^ ((i * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)) % P
    return out


def det_vec_fast(n, seed):
    """Vectorised det_vec (same values): LCG mod 2^32 in closed form via repeated affine doubling."""
    A, Cc = 1664525, 1013904223
    a0 = (A * seed + Cc) & 0xFFFFFFFFFFFFFFFF
    # after the first step the state is < 2^32 and everything is mod 2^32
    x = np.empty(n, np.uint64)
    cur = ((a0 * A + Cc) & 0xFFFFFFFFFFFFFFFF) % (1 << 32)
    # block-doubling: x[0]=cur; x[i+k] = A^k x[i] + C*(A^k-1)/(A-1)  (mod 2^32)
    x[0] = cur
    filled, mulk, addk = 1, A, Cc
    M32 = np.uint64(0xFFFFFFFF)
    while filled < n:
        take = min(filled, n - filled)
        x[filled:filled + take] = (x[:take] * np.uint64(mulk) + np.uint64(addk)) & M32
        addk = (addk * mulk + addk) & 0xFFFFFFFF
        mulk = (mulk * mulk) & 0xFFFFFFFF
        filled += take
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    return (x ^ i) % np.uint64(P)


def det_coeffs(n):
    """crates/sezkp-ffts/tests/coset_lde.rs:17-21"""
    with np.errstate(over="ignore"):
        return (np.arange(n, dtype=np.uint64) * np.uint64(0xDEADBEEF ^ 0x42)) % np.uint64(P)
