"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol,
the header and the binding agree, and there is no silent fallback without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, pkg


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sezkp_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sezkp_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    m = pkg()
    if not os.path.exists(m.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = m.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libsezkp_cuda.so does not export {s}"
    assert sorted(m.EXPORTS) == syms, "binding.EXPORTS and include/sezkp_cuda.h disagree"
    assert lib.sezkp_cuda_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = pkg()
    with pytest.raises(m.SezkpCudaError) as ei:
        m.Context()
    assert ei.value.code == -4  # ENODEV
    with pytest.raises(m.SezkpCudaError):
        m.StarkV1Cuda.prove(m.demo_block(16), bytes(32))
    with pytest.raises(m.SezkpCudaError) as ei:
        m.Context(devices=[0, 1])  # sezkp_cuda_create_multi
    assert ei.value.code == -4


def test_trace_desc_layout_matches_header():
    m = pkg()
    # struct sezkp_trace_desc: u32,u32,u64,u64, 9 pointers
    assert C.sizeof(m.TraceDesc) == 4 + 4 + 8 + 8 + 9 * 8
    ct = m.demo_block(16)
    d = ct.as_desc()
    assert d.tau == 1 and d.n_blocks == 1 and d.n_rows == 16 and d.flags == 0
    # packed per-tape ops (SEZKP_TRACE_PACKED_OPS): one byte per (row, tape) = (mv + 1) | written << 2 | symbol << 3
    assert ct.pack_ops()
    dp = ct.as_desc()
    assert dp.flags == 1 and dp.write_flag is None and dp.write_sym is None
    i = np.arange(16)
    want = ((i % 2 == 0).astype(np.uint8) + 1) | ((i % 3 == 0).astype(np.uint8) << 2) | (np.where(i % 3 == 0, 5, 0).astype(np.uint8) << 3)
    assert np.array_equal(ct.ops[:, 0], want)
    assert ct.as_desc(packed=False).flags == 0


def test_simulate_generator_and_partition():
    m = pkg()
    from importlib import import_module
    tr = import_module("streaming-zero-knowledge-proofs_b200.trace")
    assert tr._splitmix_block(42, 0, 3).tolist() == [13679457532755275413, 2949826092126892291, 5139283748462763858]
    ct = m.simulate(1 << 12, 512, 8)
    assert ct.n_rows == 4096 and ct.n_blocks == 8 and ct.tau == 8
    assert set(np.unique(ct.mv)) <= {-1, 0, 1} and set(np.unique(ct.input_mv)) <= {-1, 0, 1}
    assert int(ct.write_sym.max()) <= 15 and np.all(ct.write_sym[ct.write_flag == 0] == 0)
    # partition invariants (reference partition.rs:61-147): window covers [min,max] of post-move heads, entry at 0
    for k in range(ct.n_blocks):
        rows = slice(k * 512, (k + 1) * 512)
        heads = np.cumsum(ct.mv[rows].astype(np.int64), axis=0)
        assert np.array_equal(ct.win_left[k], np.minimum(heads.min(axis=0), 0))
        assert np.array_equal(ct.win_right[k], np.maximum(heads.max(axis=0), 0))
        assert np.array_equal(ct.head_in_off[k].astype(np.int64), -ct.win_left[k])
        assert np.array_equal(ct.head_out_off[k].astype(np.int64), heads[-1] - ct.win_left[k])
    assert ct.step_lo.tolist() == [1 + 512 * k for k in range(8)] and ct.step_hi.tolist() == [512 * (k + 1) for k in range(8)]
    # ragged last block
    ct2 = m.simulate(1000, 512, 2)
    assert ct2.block_len.tolist() == [512, 488]


def test_artifact_envelope_matches_reference_fixture_shape(tmp_path):
    """CBOR/JSON envelope round trip; same map shape as the reference's shipped proof_stark.cbor (v0 artifact)."""
    import cbor2
    from conftest import load_fixture
    m = pkg()
    fx = load_fixture("fixture_root_T64.json")["proof_v0"]
    art = m.ProofArtifact("stark", bytes.fromhex(fx["manifest_root"]), bytes.fromhex(fx["proof_bytes"]),
                          {"tau": 2, "proto": "stark-v1", "domain_n": 512})
    for ext in ("cbor", "json"):
        p = str(tmp_path / f"proof.{ext}")
        m.artifact.write_proof_auto(p, art)
        back = m.artifact.read_proof_auto(p)
        assert back == art
    obj = cbor2.load(open(str(tmp_path / "proof.cbor"), "rb"))
    assert list(obj.keys()) == ["backend", "manifest_root", "proof_bytes", "meta"]
    assert isinstance(obj["proof_bytes"], list) and all(isinstance(x, int) for x in obj["proof_bytes"][:8])  # Vec<u8> without serde_bytes
    assert list(obj["meta"].keys()) == ["domain_n", "proto", "tau"]  # serde_json map: alphabetical
    assert len(obj["manifest_root"]) == 32


def test_simulate_exact_reproduces_the_reference_fixture():
    """f4: `sezkp-cli simulate` inputs (rand 0.9.2 StdRng = ChaCha12, seed_from_u64(42), generator.rs:38-73) and
    partition_trace (partition.rs:61-147) reproduced exactly: every field of the reference's shipped blocks.cbor (T=64, b=8,
    tau=2) and its manifest root."""
    from conftest import load_fixture
    m = pkg()
    fx = load_fixture("fixture_root_T64.json")
    ct = m.simulate_exact(64, 8, 2)
    assert ct.n_blocks == len(fx["blocks"])
    row = 0
    for k, bk in enumerate(fx["blocks"]):
        steps = bk["movement_log"]["steps"]
        assert int(ct.block_len[k]) == len(steps)
        assert (int(ct.version[k]), int(ct.block_id[k]), int(ct.step_lo[k]), int(ct.step_hi[k])) == (bk["version"], bk["block_id"], bk["step_lo"], bk["step_hi"])
        assert (int(ct.ctrl_in[k]), int(ct.ctrl_out[k]), int(ct.in_head_in[k]), int(ct.in_head_out[k])) == (bk["ctrl_in"], bk["ctrl_out"], bk["in_head_in"], bk["in_head_out"])
        assert [int(x) for x in ct.win_left[k]] == [w["left"] for w in bk["windows"]]
        assert [int(x) for x in ct.win_right[k]] == [w["right"] for w in bk["windows"]]
        assert [int(x) for x in ct.head_in_off[k]] == bk["head_in_offsets"] and [int(x) for x in ct.head_out_off[k]] == bk["head_out_offsets"]
        for j, s in enumerate(steps):
            assert int(ct.input_mv[row + j]) == s["input_mv"]
            for r, tp in enumerate(s["tapes"]):
                assert int(ct.mv[row + j, r]) == tp["mv"]
                assert bool(ct.write_flag[row + j, r]) == (tp["write"] is not None) and int(ct.write_sym[row + j, r]) == (tp["write"] or 0)
        row += len(steps)
    assert m.manifest_root(ct).hex() == fx["manifest"]["root"]
