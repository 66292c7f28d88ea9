"""Compact trace model: the host-side mirror of the reference's ``&[BlockSummary]`` input.

* :class:`CompactTrace` packs the fields ``prove_v1`` reads (reference
  ``crates/sezkp-stark/src/v1/columns.rs:252-365``, ``crates/sezkp-core/src/types.rs:116-151``) into flat
  numpy arrays and exposes them as the C ABI's ``sezkp_trace_desc`` (``include/sezkp_trace.h``).
* :func:`simulate` is the deterministic stand-in for ``sezkp-cli simulate`` + ``partition_trace``
  (reference ``crates/sezkp-trace/src/generator.rs:38-73`` distribution, ``partition.rs:43-150`` rules
  verbatim) driven by splitmix64 instead of rand's ChaCha12.
* :func:`manifest_root` restates ``sezkp_merkle::{leaf_hash, merkle_root}`` (reference
  ``crates/sezkp-merkle/src/lib.rs:85-157``) on the host with the ``blake3`` package — the manifest is T/b
  leaves and stays on the CPU in the reference's CLI as well.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence

import numpy as np


class TraceDesc(C.Structure):
    """ctypes image of ``sezkp_trace_desc`` (include/sezkp_trace.h)."""

    _fields_ = [
        ("tau", C.c_uint32),
        ("flags", C.c_uint32),
        ("n_blocks", C.c_uint64),
        ("n_rows", C.c_uint64),
        ("block_len", C.c_void_p),
        ("win_left", C.c_void_p),
        ("win_right", C.c_void_p),
        ("head_in_off", C.c_void_p),
        ("head_out_off", C.c_void_p),
        ("input_mv", C.c_void_p),
        ("mv", C.c_void_p),
        ("write_flag", C.c_void_p),
        ("write_sym", C.c_void_p),
    ]


@dataclass
class CompactTrace:
    tau: int
    block_len: np.ndarray      # u64 [n_blocks]
    win_left: np.ndarray       # i64 [n_blocks, tau]
    win_right: np.ndarray      # i64 [n_blocks, tau]
    head_in_off: np.ndarray    # u32 [n_blocks, tau]
    head_out_off: np.ndarray   # u32 [n_blocks, tau]
    input_mv: np.ndarray       # i8  [n_rows]
    mv: np.ndarray             # i8  [n_rows, tau]
    write_flag: np.ndarray     # u8  [n_rows, tau]
    write_sym: np.ndarray      # u16 [n_rows, tau]
    # manifest-only scalars (sezkp-merkle leaf_hash); not read by the prover
    version: Optional[np.ndarray] = None     # u16
    block_id: Optional[np.ndarray] = None    # u32
    step_lo: Optional[np.ndarray] = None     # u64
    step_hi: Optional[np.ndarray] = None     # u64
    ctrl_in: Optional[np.ndarray] = None     # u16
    ctrl_out: Optional[np.ndarray] = None    # u16
    in_head_in: Optional[np.ndarray] = None  # i64
    in_head_out: Optional[np.ndarray] = None # i64
    ops: Optional[np.ndarray] = None         # u8 [n_rows, tau]: packed per-tape ops (pack_ops), sent instead of mv/flag/sym
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_blocks(self) -> int:
        return int(self.block_len.shape[0])

    @property
    def n_rows(self) -> int:
        return int(self.input_mv.shape[0])

    def nbytes(self) -> int:
        rows = (self.input_mv, self.ops) if self.ops is not None else (self.input_mv, self.mv, self.write_flag, self.write_sym)
        return sum(a.nbytes for a in (self.block_len, self.win_left, self.win_right, self.head_in_off, self.head_out_off) + rows)

    def pack_ops(self) -> bool:
        """Build the SEZKP_TRACE_PACKED_OPS form of the per-tape arrays (include/sezkp_trace.h): one byte per (row, tape)
        = (mv + 1) | written << 2 | symbol << 3.  Possible when every move is in {-1, 0, 1} and every symbol < 32;
        returns False (and leaves the trace unpacked) otherwise.  as_desc() sends the packed form once it exists."""
        mv, fl, sy = np.asarray(self.mv), np.asarray(self.write_flag), np.asarray(self.write_sym)
        if mv.size and (int(mv.min()) < -1 or int(mv.max()) > 1 or int(sy.max()) > 31 or int(fl.max()) > 1):
            self.ops = None
            return False
        self.ops = np.ascontiguousarray((mv.astype(np.int16) + 1).astype(np.uint8) | (fl.astype(np.uint8) << 2) | (sy.astype(np.uint8) << 3))
        return True

    def _c(self):
        """Contiguous, correctly typed views (kept alive on self)."""
        spec = [("block_len", np.uint64), ("win_left", np.int64), ("win_right", np.int64),
                ("head_in_off", np.uint32), ("head_out_off", np.uint32), ("input_mv", np.int8),
                ("mv", np.int8), ("write_flag", np.uint8), ("write_sym", np.uint16)]
        out = {}
        for name, dt in spec:
            a = np.ascontiguousarray(getattr(self, name), dtype=dt)
            setattr(self, name, a)
            out[name] = a
        return out

    def as_desc(self, packed: bool = True) -> TraceDesc:
        a = self._c()
        d = TraceDesc()
        d.tau = self.tau
        d.flags = 0
        d.n_blocks = self.n_blocks
        d.n_rows = self.n_rows
        for name, arr in a.items():
            setattr(d, name, arr.ctypes.data if arr.size else None)
        self._desc_arrays = list(a.values())
        if packed and self.ops is not None:  # packed per-tape ops replace the three arrays
            ops = np.ascontiguousarray(self.ops, dtype=np.uint8)
            self.ops = ops
            d.flags = 1
            d.mv = ops.ctypes.data if ops.size else None
            d.write_flag = None
            d.write_sym = None
            self._desc_arrays.append(ops)
        return d


def blocks_to_compact(blocks: Sequence[dict]) -> CompactTrace:
    """Pack serde-shaped BlockSummary dicts (as read from the reference's CBOR/JSON) into a CompactTrace."""
    nb = len(blocks)
    tau = len(blocks[0]["windows"]) if nb else 0
    lens = np.array([b["step_hi"] - b["step_lo"] + 1 for b in blocks], dtype=np.uint64)
    n = int(lens.sum())
    ct = CompactTrace(
        tau=tau,
        block_len=lens,
        win_left=np.zeros((nb, tau), np.int64), win_right=np.zeros((nb, tau), np.int64),
        head_in_off=np.zeros((nb, tau), np.uint32), head_out_off=np.zeros((nb, tau), np.uint32),
        input_mv=np.zeros(n, np.int8), mv=np.zeros((n, tau), np.int8),
        write_flag=np.zeros((n, tau), np.uint8), write_sym=np.zeros((n, tau), np.uint16),
        version=np.zeros(nb, np.uint16), block_id=np.zeros(nb, np.uint32), step_lo=np.zeros(nb, np.uint64),
        step_hi=np.zeros(nb, np.uint64), ctrl_in=np.zeros(nb, np.uint16), ctrl_out=np.zeros(nb, np.uint16),
        in_head_in=np.zeros(nb, np.int64), in_head_out=np.zeros(nb, np.int64),
    )
    row = 0
    for k, b in enumerate(blocks):
        steps = b["movement_log"]["steps"]
        if len(steps) != int(lens[k]):
            raise ValueError(f"block {k}: movement_log has {len(steps)} steps, step range says {int(lens[k])}")
        if len(b["windows"]) != tau:
            raise ValueError(f"block {k}: tau mismatch")
        for r in range(tau):
            ct.win_left[k, r] = b["windows"][r]["left"]
            ct.win_right[k, r] = b["windows"][r]["right"]
            ct.head_in_off[k, r] = b["head_in_offsets"][r]
            ct.head_out_off[k, r] = b["head_out_offsets"][r]
        for name in ("version", "block_id", "step_lo", "step_hi", "ctrl_in", "ctrl_out", "in_head_in", "in_head_out"):
            getattr(ct, name)[k] = b[name]
        for j, st in enumerate(steps):
            ct.input_mv[row + j] = st["input_mv"]
            for r, op in enumerate(st["tapes"]):
                ct.mv[row + j, r] = op["mv"]
                if op["write"] is not None:
                    ct.write_flag[row + j, r] = 1
                    ct.write_sym[row + j, r] = op["write"]
        row += len(steps)
    return ct


# ----------------------------------------------------------------------------------------------
# deterministic simulate + partition: splitmix64 stream, state s0 = seed
# ----------------------------------------------------------------------------------------------
_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def _splitmix_block(seed: int, start: int, count: int) -> np.ndarray:
    """Outputs number start+1 .. start+count of splitmix64 seeded with `seed` (1-based call index)."""
    with np.errstate(over="ignore"):
        k = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def simulate(T: int, b: int, tau: int, seed: int = 42) -> CompactTrace:
    """Deterministic toy trace with the reference generator's distribution, partitioned into blocks of b.

    Per step, in stream order: ``input_mv = next()%3 - 1``; then per tape ``w = next()``:
    ``write = Some((w>>32)%16) if w%10 < 4 else None``; ``mv = next()%3 - 1``.
    Partition rules follow reference ``crates/sezkp-trace/src/partition.rs:61-147``.
    """
    if T <= 0 or b <= 0:
        raise ValueError("T and b must be positive")
    per = 1 + 2 * tau
    input_mv = np.empty(T, np.int8)
    mv = np.empty((T, tau), np.int8)
    wflag = np.empty((T, tau), np.uint8)
    wsym = np.empty((T, tau), np.uint16)
    CH = 1 << 18
    for lo in range(0, T, CH):
        hi = min(T, lo + CH)
        r = _splitmix_block(seed, lo * per, (hi - lo) * per).reshape(hi - lo, per)
        input_mv[lo:hi] = (r[:, 0] % np.uint64(3)).astype(np.int8) - 1
        w = r[:, 1::2]
        m = r[:, 2::2]
        f = (w % np.uint64(10)) < np.uint64(4)
        wflag[lo:hi] = f
        wsym[lo:hi] = np.where(f, (w >> np.uint64(32)) % np.uint64(16), 0).astype(np.uint16)
        mv[lo:hi] = (m % np.uint64(3)).astype(np.int8) - 1
    return partition(input_mv, mv, wflag, wsym, b)


# ---- exact reproduction of `sezkp-cli simulate` (reference crates/sezkp-trace/src/generator.rs:38-73) -------------------
# rand 0.9.2: StdRng = ChaCha12, seed_from_u64(42) expands the seed with PCG32 (rand_core), BlockRng hands out the words of
# four ChaCha blocks at a time; random_range(0..=2) (i32) and random_range(0u16..=15) draw one u32 and take the high half of
# sample * range (a second draw only when the low half exceeds 2^32 - range: Canon's single-extra-sample correction);
# random_bool(0.4) compares one u64 (two words, low first) with floor(0.4 * 2^64).
def _chacha12_words(key, first_block: int, n_blocks: int) -> np.ndarray:
    """ChaCha12 keystream words of blocks [first_block, first_block + n_blocks), 16 u32 each, vectorised over blocks."""
    M = np.uint64(0xFFFFFFFF)
    ctr = np.arange(first_block, first_block + n_blocks, dtype=np.uint64)
    init = [np.full(n_blocks, c, np.uint64) for c in (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)]
    init += [np.full(n_blocks, k, np.uint64) for k in key]
    init += [ctr & M, ctr >> np.uint64(32), np.zeros(n_blocks, np.uint64), np.zeros(n_blocks, np.uint64)]
    w = [x.copy() for x in init]

    def rotl(x, r):
        return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & M

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & M; w[d] = rotl(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & M; w[b] = rotl(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & M; w[d] = rotl(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & M; w[b] = rotl(w[b] ^ w[c], 7)

    for _ in range(6):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    out = np.stack([(w[i] + init[i]) & M for i in range(16)], axis=1)
    return out.reshape(-1)


def _std_rng_seed(state: int):
    """rand_core SeedableRng::seed_from_u64: eight PCG32 outputs = the 32-byte ChaCha key (little-endian words)."""
    key = []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & 0xFFFFFFFFFFFFFFFF
        xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        key.append(((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
    return key


def simulate_exact(T: int, b: int, tau: int) -> CompactTrace:
    """Byte-identical inputs to `sezkp-cli simulate --t T --b b --tau tau` (generate_trace + partition_trace); checked
    against the reference's shipped blocks.cbor (tests/golden/fixture_root_T64.json).  The draw sequence is data dependent
    (a symbol is drawn only when the write coin comes up), so the consumer is a scalar loop over a vectorised keystream:
    about 1e5 steps/s — use `simulate` (splitmix, same distribution) for the large benchmark traces."""
    if T <= 0 or b <= 0 or not (1 <= tau <= 255):
        raise ValueError("T and b must be positive, 1 <= tau <= 255")
    key = _std_rng_seed(42)
    P_INT = int(0.4 * 2.0 ** 64)  # Bernoulli::new(0.4).p_int
    input_mv = np.empty(T, np.int8)
    mv = np.empty((T, tau), np.int8)
    wflag = np.zeros((T, tau), np.uint8)
    wsym = np.zeros((T, tau), np.uint16)
    # BlockRng buffer: 64 words (4 blocks) per refill; next_u64 at index 63 pairs the last word (low) with the first word of
    # the next buffer (high) — i.e. the stream is consumed strictly word by word, so one flat stream is equivalent
    words, pos, next_block = [], 0, 0

    def need(k):
        nonlocal words, pos, next_block
        if pos + k > len(words):
            blocks = max(4096, (k + 15) // 16)
            words = words[pos:] + _chacha12_words(key, next_block, blocks).tolist()
            pos = 0
            next_block += blocks

    def u32():
        nonlocal pos
        v = words[pos]
        pos += 1
        return v

    def rng_range(r):
        m = u32() * r
        hi, lo = m >> 32, m & 0xFFFFFFFF
        if lo > ((-r) & 0xFFFFFFFF):
            hi += 1 if lo + ((u32() * r) >> 32) > 0xFFFFFFFF else 0
        return hi

    for i in range(T):
        need(2 + 6 * tau)
        input_mv[i] = rng_range(3) - 1
        for r in range(tau):
            lo = u32()
            if (lo | (u32() << 32)) < P_INT:
                wflag[i, r] = 1
                wsym[i, r] = rng_range(16)
            mv[i, r] = rng_range(3) - 1
    return partition(input_mv, mv, wflag, wsym, b)


def partition(input_mv, mv, wflag, wsym, b: int) -> CompactTrace:
    """reference partition_trace (partition.rs:43-150) on flat step arrays."""
    T, tau = mv.shape
    nb = (T + b - 1) // b
    lens = np.full(nb, b, np.uint64)
    lens[-1] = T - (nb - 1) * b
    starts = np.arange(nb, dtype=np.int64) * b
    # per-block relative heads: cumulative sum restarted at each block start
    cs = np.cumsum(mv.astype(np.int64), axis=0)
    base = np.zeros((nb, tau), np.int64)
    base[1:] = cs[starts[1:] - 1]
    rel = cs - np.repeat(base, lens.astype(np.int64), axis=0)
    min_pos = np.minimum(np.minimum.reduceat(rel, starts, axis=0), 0)   # min/max start at 0 (entry head)
    max_pos = np.maximum(np.maximum.reduceat(rel, starts, axis=0), 0)
    ends = starts + lens.astype(np.int64) - 1
    cur = rel[ends]
    gi = np.cumsum(input_mv.astype(np.int64))
    in_head_out = gi[ends]
    in_head_in = np.concatenate([[0], in_head_out[:-1]])

    def to_u32(x):  # u32::try_from(..).unwrap_or(u32::MAX)
        return np.where((x < 0) | (x > 0xFFFFFFFF), 0xFFFFFFFF, x).astype(np.uint32)

    return CompactTrace(
        tau=tau, block_len=lens, win_left=min_pos, win_right=max_pos,
        head_in_off=to_u32(-min_pos), head_out_off=to_u32(cur - min_pos),
        input_mv=np.ascontiguousarray(input_mv, np.int8), mv=np.ascontiguousarray(mv, np.int8),
        write_flag=np.ascontiguousarray(wflag, np.uint8), write_sym=np.ascontiguousarray(wsym, np.uint16),
        version=np.ones(nb, np.uint16), block_id=np.arange(1, nb + 1, dtype=np.uint32),
        step_lo=(starts + 1).astype(np.uint64), step_hi=(ends + 1).astype(np.uint64),
        ctrl_in=np.zeros(nb, np.uint16), ctrl_out=np.zeros(nb, np.uint16),
        in_head_in=in_head_in.astype(np.int64), in_head_out=in_head_out.astype(np.int64),
    )


def demo_block(T: int) -> CompactTrace:
    """The hand-built tau=1 block of the reference's tests (crates/sezkp-stark/tests/air_ok.rs,
    stream_fri_equiv.rs): mv = 1,0,1,0,...; write 5 when i%3==0; window [0, T-1]; in_off 0; out_off = sum(mv)."""
    i = np.arange(T)
    mv = (i % 2 == 0).astype(np.int8).reshape(T, 1)
    wf = (i % 3 == 0).astype(np.uint8).reshape(T, 1)
    ws = np.where(i % 3 == 0, 5, 0).astype(np.uint16).reshape(T, 1)
    return CompactTrace(
        tau=1, block_len=np.array([T], np.uint64), win_left=np.zeros((1, 1), np.int64),
        win_right=np.full((1, 1), T - 1, np.int64), head_in_off=np.zeros((1, 1), np.uint32),
        head_out_off=np.array([[int(mv.sum())]], np.uint32), input_mv=np.zeros(T, np.int8), mv=mv,
        write_flag=wf, write_sym=ws, version=np.ones(1, np.uint16), block_id=np.ones(1, np.uint32),
        step_lo=np.ones(1, np.uint64), step_hi=np.array([T], np.uint64), ctrl_in=np.zeros(1, np.uint16),
        ctrl_out=np.zeros(1, np.uint16), in_head_in=np.zeros(1, np.int64), in_head_out=np.zeros(1, np.int64),
    )


def manifest_leaf_preimage(ct: CompactTrace, k: int) -> bytes:
    """sezkp_merkle::leaf_hash preimage (reference crates/sezkp-merkle/src/lib.rs:85-117)."""
    tau = ct.tau
    out = struct.pack("<HIQQHHqq", int(ct.version[k]), int(ct.block_id[k]), int(ct.step_lo[k]), int(ct.step_hi[k]),
                      int(ct.ctrl_in[k]), int(ct.ctrl_out[k]), int(ct.in_head_in[k]), int(ct.in_head_out[k]))
    out += struct.pack("<Q", tau)
    for r in range(tau):
        out += struct.pack("<qq", int(ct.win_left[k, r]), int(ct.win_right[k, r]))
    out += b"".join(struct.pack("<I", int(x)) for x in ct.head_in_off[k])
    out += b"".join(struct.pack("<I", int(x)) for x in ct.head_out_off[k])
    out += struct.pack("<Q", int(ct.block_len[k]))
    return out


def manifest_root(ct: CompactTrace) -> bytes:
    """sezkp_merkle::commit_blocks root (batch merkle_root with odd promotion, lib.rs:140-157, 215-222)."""
    import blake3  # the official crate's Python binding; host-side only (T/b leaves)

    lvl = [blake3.blake3(manifest_leaf_preimage(ct, k)).digest() for k in range(ct.n_blocks)]
    if not lvl:
        return bytes(32)
    while len(lvl) > 1:
        nxt = [blake3.blake3(lvl[i] + lvl[i + 1]).digest() if i + 1 < len(lvl) else lvl[i] for i in range(0, len(lvl), 2)]
        lvl = nxt
    return lvl[0]
