"""Pin the CPU oracle before anything is compared against it (CPU-only; no GPU).

Sources of truth, in decreasing authority:
  1. the reference's shipped fixtures (manifest roots, v0 proof bytes)  -> BLAKE3, transcript, parent combiner
  2. Python `blake3` (bindings to the crate the reference uses) known answers committed in tests/golden
  3. exact mathematical definitions (naive DFT, field identities)
  4. SURVEY.md §8c known-answer values from an independent Python restatement of prove_v1
"""
import hashlib
import json

import numpy as np
import pytest

from conftest import load_fixture, pkg
from oracle_lib import P, det_coeffs, det_vec, det_vec_fast

H = bytes.fromhex


def test_blake3_kats(oracle):
    kats = load_fixture("blake3_kats.json")
    for c in kats["cases"]:
        data = bytes((i * 251 + 7) % 256 for i in range(c["len"]))
        assert oracle.blake3(data).hex() == c["out32"], c["len"]
        assert oracle.blake3(data, 300).hex() == c["xof300"], c["len"]
    assert oracle.blake3(b"").hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"


def test_blake3_against_module_if_present(oracle):
    blake3 = pytest.importorskip("blake3")
    rng = np.random.default_rng(1)
    for n in [0, 3, 64, 100, 1024, 1500, 4097, 9999]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.blake3(d, 240) == blake3.blake3(d).digest(length=240)


@pytest.mark.parametrize("name", ["fixture_root_T64.json", "fixture_riscv_T32.json"])
def test_reference_fixture_manifest_root_and_v0_proof(oracle, name):
    """sezkp-merkle leaf_hash/node_hash/merkle_root + Blake3Transcript framing, against shipped files."""
    fx = load_fixture(name)
    ct = pkg().blocks_to_compact(fx["blocks"])
    root = oracle.manifest_root(ct)
    assert root.hex() == fx["manifest"]["root"]
    assert pkg().manifest_root(ct).hex() == fx["manifest"]["root"]  # host-side restatement in the package
    # v0 StarkIOP proof bytes: sezkp-stark/src/lib.rs:66-92, commit.rs:47-90, witness.rs:33-105
    tau = ct.tau
    tr = oracle.transcript("sezkp-stark/v0/row-stream")
    tr.absorb_u64("tau", tau)
    rows = bytearray()
    n_rows = 0
    chunks = []
    for j in range(ct.n_rows):
        rows.append(int(ct.input_mv[j]) & 0xFF)
        for r in range(tau):
            rows.append((int(ct.mv[j, r]) + 1) & 0xFF)
            rows.append(int(ct.write_flag[j, r]))
        n_rows += 1
        if n_rows % 4096 == 0:
            chunks.append(bytes(rows))
            rows = bytearray()
    if rows:
        chunks.append(bytes(rows))
    for ch in chunks:
        tr.absorb("rows", ch)
    commit_root = tr.challenge("root", 32)
    tr2 = oracle.transcript("sezkp-stark-v0")
    tr2.absorb("manifest_root", root)
    tr2.absorb("commit_root", commit_root)
    tr2.absorb_u64("n_rows", n_rows)
    tr2.absorb_u64("tau", tau)
    proof = tr2.challenge("alpha", 32) + tr2.challenge("beta", 32)
    assert proof.hex() == fx["proof_v0"]["proof_bytes"]
    assert fx["proof_v0"]["manifest_root"] == fx["manifest"]["root"]


def test_field_kats(oracle):
    assert oracle.gl_root_2exp(1) == 0xFFFFFFFF00000000
    assert oracle.gl_root_2exp(2) == 0x0001000000000000
    assert oracle.gl_root_2exp(3) == 0xFFFFFFFEFF000001
    for k, v in [(15, 0xF6B2CFFE2306BAAC), (18, 0x81281A7B05F9BEAC), (20, 0x30BA2ECD5E93E76D),
                 (22, 0x4B2A18ADE67246B5), (25, 0x4BBAF5976ECFEFD8), (32, 0x185629DCDA58878C)]:
        assert oracle.gl_root_2exp(k) == v
    assert oracle.gl_inv(3) == 12297829379609722881
    assert oracle.gl_inv(1 << 20) == 18446726477228544001
    assert oracle.gl_from_i64(-1) == 18446744069414584320
    # python big-int cross-check of mul/pow/inv
    rng = np.random.default_rng(7)
    for _ in range(200):
        a, b = int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64))
        assert oracle.gl_mul(a, b) == a * b % P
        assert oracle.gl_pow(a, b) == pow(a, b, P)
        if a:
            assert oracle.gl_inv(a) == pow(a, P - 2, P)


def test_ntt_kats_and_naive_dft(oracle):
    out = oracle.ntt(np.arange(1, 9, dtype=np.uint64))
    assert out.tolist() == [36, 18445622567621360637, 18445618169507741693, 1130298020461564,
                            18446744069414584317, 18445613771394122749, 1125899906842620, 1121501793223676]
    ev = oracle.coset_eval(np.array([1, 2, 3, 4], np.uint64), 3, 3)
    assert ev.tolist() == [142, 7481077014752257, 18418033621790097383, 18439137646161692162,
                           18446744069414584235, 7718571727623169, 28710447624486886, 18439150843925101058]
    for k in range(1, 9):  # fast transform == O(n^2) definition (sezkp-ffts/src/lib.rs:191-202)
        v = det_vec(1 << k, 1337)
        assert np.array_equal(oracle.ntt(v), oracle.dft_naive(v)), k
    assert det_vec(4, 2024).tolist() == [3706503514, 11400714820744568036, 4354685562416586166, 15755400382988054644]
    assert np.array_equal(det_vec(5000, 2024), det_vec_fast(5000, 2024))


def _b3(a):
    import blake3 as _b  # noqa
    return _b.blake3(np.ascontiguousarray(a, np.uint64).tobytes()).hexdigest()


def test_ntt_digest_kats(oracle):
    """BLAKE3 over LE bytes of det_vec / forward NTT / coset eval."""
    for n, kc, exp in [
        (1 << 10, 12, ("a275d0be8ea7e69be799bbaad9bb007e11867e491f2dd6f9947fd614339954f1",
                       "a053894941d273c9145e9b03916b38214b000ef17b4a575c238a5b0ec8497ce0",
                       "d1eea764a7935e9b8bf1ef527846b4824d41f35669d6222f17acb587ac8e9c9b")),
        (1 << 12, 14, ("ab81158cca5e8c8f409015d6764313c62e06eee7ddcfc2edf35158e6c89c9741",
                       "d9c2fde786acee2089bc76bc0355184b6a96827a032668b2c6b07cfb737f90fd",
                       "86c5428800c396ab638ad510b08620e84dacdd8ca67367612f9430f1505e3fdd")),
    ]:
        v = det_vec_fast(n, 2024)
        assert oracle.blake3(v.tobytes()).hex() == exp[0]
        assert oracle.blake3(oracle.ntt(v).tobytes()).hex() == exp[1]
        assert oracle.blake3(oracle.coset_eval(v, kc, 3).tobytes()).hex() == exp[2]


def test_leaf_and_node_kats(oracle):
    l = oracle.leaf_hash(np.array([0, 1], np.uint64))
    assert l[0].tobytes().hex() == "71e0a99173564931c0b8acc52d2685a8e39c64dc52e3d02390fdac2a12b155cb"
    assert l[1].tobytes().hex() == "1a0d12016999e47689dae5744d2b8c1903faf7ca2886a658150083100ef2c8ee"
    assert oracle.leaf_hash(np.array([1], np.uint64), "is_first")[0].tobytes().hex() == \
        "24403bd2529ac61a592963d5be5045db5ed01cf396b129563ff97fde7d97d5e7"
    assert oracle.leaf_hash(np.array([P - 1], np.uint64), "mv_0")[0].tobytes().hex() == \
        "35e213688025f488d133edc3f1ab9cd2c5ad2b03be42037a3e9006d6151395c8"
    z = bytes(32)
    assert oracle.node_hash(z, z).hex() == "4d006976636a8696d909a630a4081aad4d7c50f81afdee04020bf05086ab6a55"
    assert oracle.node_hash(l[0].tobytes(), l[1].tobytes()).hex() == \
        "48e7bffbeecd0579c8ec8df002a3cc435f6b7feebc842e05468bc8f2d78652d4"
    # definitions: leaf = BLAKE3(le8); labeled = BLAKE3("col_leaf"||u32(len)||label||le8); node = BLAKE3(l||r)
    v = 0x0123456789ABCDEF % P
    le = v.to_bytes(8, "little")
    assert oracle.leaf_hash(np.array([v], np.uint64))[0].tobytes() == oracle.blake3(le)
    lab = "out_off_7"
    pre = b"col_leaf" + len(lab).to_bytes(4, "little") + lab.encode() + le
    assert oracle.leaf_hash(np.array([v], np.uint64), lab)[0].tobytes() == oracle.blake3(pre)


PROVE_KATS = {
    "fixture_root_T64.json": dict(n=64, n_cols=17, col_root0="b7193761afcfcea3f9934df7276d9343fe28c734f9cf2ed694a7e671a8b467b3",
                                  fri0="9d954c557d437788be88202e6127bd2a163cd74173b27f07a71970de6208ee81",
                                  fri1="dbadea344153261ab712972612d9cac2d72e82a827b314353e4ae60e9757d9ca",
                                  final=13304568597758088877, rows6=[60, 24, 1, 42, 24, 61],
                                  fri6=[446, 349, 391, 489, 203, 256], blen=270967,
                                  b3="7852805e6d6c64027aafe0c2672927715bcd92aaa08f83383de1fc6a183ba6ca"),
    "fixture_riscv_T32.json": dict(n=32, fri0="2fc583a4b74b4c5532148636c792b5220b79c54d15055f8513e3e82c46923834",
                                   final=10818276303087028547, blen=232295,
                                   b3="6bc9eedf2415d4af75277bb226e077618fbb1028683f15287028a0a50dd96dd4"),
}


def parse_head(proof):
    """minimal bincode peek: domain_n, tau, n_cols, first col label/root"""
    import struct
    dn, tau, nc = struct.unpack_from("<QQQ", proof, 0)
    off = 24
    (ll,) = struct.unpack_from("<Q", proof, off)
    label = proof[off + 8: off + 8 + ll].decode()
    root0 = proof[off + 8 + ll: off + 8 + ll + 32]
    return dn, tau, nc, label, root0


@pytest.mark.parametrize("name", list(PROVE_KATS))
def test_prove_v1_kats_on_reference_fixtures(oracle, name):
    """Cross-restatement check: C++ oracle == SURVEY's independent Python restatement on the shipped blocks."""
    fx = load_fixture(name)
    k = PROVE_KATS[name]
    ct = pkg().blocks_to_compact(fx["blocks"])
    proof, taps = oracle.prove_v1(ct, H(fx["manifest"]["root"]), taps=True)
    assert len(proof) == k["blen"]
    assert oracle.blake3(proof).hex() == k["b3"]
    dn, tau, nc, label, root0 = parse_head(proof)
    assert dn == 8 * k["n"] and tau == ct.tau and label == "input_mv"
    if "col_root0" in k:
        assert nc == k["n_cols"] and root0.hex() == k["col_root0"]
        assert list(taps.rows)[:6] == k["rows6"] and list(taps.fri_rows)[:6] == k["fri6"]


@pytest.mark.parametrize("T,fri0,final,blen,b3", [
    (16, "0b828ccf9d47869ea432343a7b41f8f1f89c297ad64b015fa7a57497c43af97e", 9294559106242659263, 139055,
     "a952a59ad878a59bdc5626b2cf53469eb9226a6c439f37b074eb0472d0c6fe08"),
    (64, "4d050b95a89907706ba756d4df92caba94418248ed958845f5a80d7684c4819d", 4171081071171934075, 197199,
     "733e0d80c093a79a7dab5464741660950997b31c1f14ab255177f120827f2d43"),
])
def test_prove_v1_demo_block_kats_and_verify(oracle, T, fri0, final, blen, b3):
    ct = pkg().demo_block(T)
    proof = oracle.prove_v1(ct, bytes([7] * 32))
    assert len(proof) == blen and oracle.blake3(proof).hex() == b3
    assert final.to_bytes(8, "little") == proof[-40:-32]
    assert H(fri0) in proof
    ok, why = oracle.verify_v1(proof, ct)  # reference tests air_ok.rs / stream_fri_equiv.rs: accept
    assert ok, why
    # faithful-cost mode (per-level recomputation like fri_stream.rs:273-309) yields identical bytes
    if T == 16:
        assert oracle.prove_v1(ct, bytes([7] * 32), faithful_cost=True) == proof


def test_verify_rejects_tampered_and_bad_air(oracle):
    ct = pkg().demo_block(16)
    proof = bytearray(oracle.prove_v1(ct, bytes([7] * 32)))
    bad = bytearray(proof)
    bad[-40] ^= 1  # final value
    assert not oracle.verify_v1(bytes(bad), ct)[0]
    # reference test air_fail_endpoint.rs: wrong out_off -> reject (when the last row is sampled; here the
    # boundary term is non-zero on row T-1, and with T=16 and 30 queries it is sampled with overwhelming odds)
    ct2 = pkg().demo_block(16)
    ct2.head_out_off[0, 0] += 1
    p2 = oracle.prove_v1(ct2, bytes([7] * 32))
    ok, why = oracle.verify_v1(p2, ct2)
    assert (not ok) and "AIR composition non-zero" in why
