"""Multi-GPU context groups (sezkp_cuda_create_multi): ONE process, one worker thread + context per GPU, every exchange
inside the library.  Outputs must be byte-identical to the single-GPU path.  A device may be listed more than once, so
the whole group logic (sharded upload, column sharding, FRI chunk sharding, opening exchange, peer sums) is exercised on a
single GPU too; with >= 2 GPUs the same tests also run over real peers."""
import numpy as np
import pytest

from conftest import pkg
from oracle_lib import P

pytestmark = pytest.mark.gpu


def device_lists():
    import torch
    n = torch.cuda.device_count()
    out = [[0, 0], [0, 0, 0]]
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    if n >= 8:
        out.append(list(range(8)))
    return out


@pytest.fixture(scope="module")
def single():
    return pkg().Context()


@pytest.fixture(scope="module", params=device_lists(), ids=lambda d: "gpus" + "".join(map(str, d)))
def group(request):
    g = pkg().Context(devices=request.param)
    assert g.n_gpus == len(request.param)
    yield g
    g.close()


@pytest.mark.parametrize("T,b,tau", [(1 << 10, 128, 1), (1 << 12, 512, 2), (1 << 14, 512, 8), (1 << 18, 512, 3)])
def test_group_prove_is_byte_identical(single, group, T, b, tau):
    """host-descriptor prove: sharded upload + all-gather of the compact trace, columns c % world, FRI hashing by chunk range
    (T = 2^18 has sharded layers of 2^21 and 2^20 values), opening records exchanged between the rank threads"""
    m = pkg()
    ct = m.simulate(T, b, tau, seed=5)
    root = m.manifest_root(ct)
    want = single.prove_v1(ct, root)
    assert group.prove_v1(ct, root) == want
    ct.pack_ops()
    assert group.prove_v1(ct, root) == want


def test_group_prove_resident(single, group):
    m = pkg()
    ct = m.simulate(1 << 16, 512, 8, seed=6)
    root = m.manifest_root(ct)
    want = single.prove_v1(ct, root)
    rt = group.upload_trace(ct)
    for _ in range(2):
        assert group.prove_v1_resident(rt, root) == want
    rt.free()


def test_group_lde_commit_and_wide_pipeline(single, group, oracle):
    rng = np.random.default_rng(77)
    c, log_n = 7, 12  # 7 columns over 2 / 3 GPUs: ragged column shares
    ev = (rng.integers(0, 1 << 63, size=(c, 1 << log_n), dtype=np.uint64) * np.uint64(2) + np.uint64(1)) % np.uint64(P)
    labels = [f"c_{k}" for k in range(c)]
    want_roots = single.lde_commit(ev, labels, 3)
    assert np.array_equal(group.lde_commit(ev, labels, 3), want_roots)
    cu_s, cu_g = single.columns_upload(ev), group.columns_upload(ev)
    a, b = single.lde_commit_fri(cu_s), group.lde_commit_fri(cu_g)
    cu_s.free(), cu_g.free()
    assert np.array_equal(a[0], want_roots)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    w = oracle.wide_tail(ev, c, log_n, want_roots)
    assert np.array_equal(b[1], w["fri_roots"]) and b[2] == w["final"]


def test_group_wide_pipeline_sharded_fri_layers(single, group):
    """2^18 rows x blow-up 8: FRI layers of 2^21 and 2^20 values are hashed by chunk range across the group"""
    cs_s, cs_g = single.columns_synth(5, 18), group.columns_synth(5, 18)
    a, b = single.lde_commit_fri(cs_s), group.lde_commit_fri(cs_g)
    cs_s.free(), cs_g.free()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_group_fewer_columns_than_gpus(single, group):
    cs_s, cs_g = single.columns_synth(1, 10), group.columns_synth(1, 10)
    a, b = single.lde_commit_fri(cs_s), group.lde_commit_fri(cs_g)
    cs_s.free(), cs_g.free()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_group_error_on_one_call_leaves_group_usable(single, group):
    m = pkg()
    bad = m.simulate(1000, 100, 2)  # not a power of two
    with pytest.raises(m.SezkpCudaError) as ei:
        group.prove_v1(bad, bytes(32))
    assert ei.value.code == -1
    ct = m.simulate(1 << 10, 128, 2, seed=9)
    root = m.manifest_root(ct)
    assert group.prove_v1(ct, root) == single.prove_v1(ct, root)


def test_group_plain_entry_points_run_on_first_gpu(group, oracle):
    v = np.arange(1 << 12, dtype=np.uint64)
    assert np.array_equal(group.ntt(v), oracle.ntt(v))


def test_group_streaming_and_jsonl(single, group, tmp_path):
    """ProvingBackendStream + the native JSONL front-end on a context group: rows are ingested on the first GPU, replicated
    to the peers over NVLink at finish, and every GPU proves its share — same bytes as the one-shot single-GPU proof."""
    m = pkg()
    ct = m.simulate(1 << 14, 512, 4, seed=11)
    root = m.manifest_root(ct)
    want = single.prove_v1(ct, root)

    def pieces():
        for k in range(0, ct.n_blocks, 5):
            e = min(ct.n_blocks, k + 5)
            r = slice(k * 512, e * 512)
            yield m.CompactTrace(tau=ct.tau, block_len=ct.block_len[k:e], win_left=ct.win_left[k:e], win_right=ct.win_right[k:e],
                                 head_in_off=ct.head_in_off[k:e], head_out_off=ct.head_out_off[k:e], input_mv=ct.input_mv[r],
                                 mv=ct.mv[r], write_flag=ct.write_flag[r], write_sym=ct.write_sym[r])

    assert group.prove_v1_stream(pieces(), root) == want
    path = str(tmp_path / "blocks.jsonl")
    m.io_jsonl.write_jsonl(path, ct)
    assert group.prove_v1_jsonl_file(path, root, ct.n_rows, ct.tau, threads=4) == want
    assert "stream_replicate_ms" in group.timings()


def test_new_api_error_paths(single):
    """bad arguments of the round-2 entry points come back as status codes; the context stays usable"""
    import ctypes as C
    m = pkg()
    lib = m.load_library()
    # device list problems
    with pytest.raises(m.SezkpCudaError):
        m.Context(devices=[0, 4096])
    with pytest.raises(m.SezkpCudaError):
        m.Context(devices=[])
    # non-canonical column values are rejected on upload
    bad = np.full((1, 16), P, np.uint64)
    with pytest.raises(m.SezkpCudaError) as ei:
        single.columns_upload(bad)
    assert ei.value.code == -1
    # a column set belongs to the context (group size) that created it
    cs = single.columns_synth(2, 8)
    g = m.Context(devices=[0, 0])
    try:
        with pytest.raises(m.SezkpCudaError) as ei:
            g.lde_commit_fri(cs)
        assert ei.value.code == -1
        # shift must be a non-zero canonical field element
        cg = g.columns_synth(2, 8)
        with pytest.raises(m.SezkpCudaError):
            g.lde_commit_fri(cg, shift=0)
        a = g.lde_commit_fri(cg)  # the group is still usable after the failed call
        b = single.lde_commit_fri(cs)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
        cg.free()
    finally:
        cs.free()
        g.close()
    # verify_openings: k = 0 is fine, a column index out of range is EINVAL
    assert single.verify_openings(np.zeros((1, 32), np.uint8), None, None, np.zeros(0, np.uint64), np.zeros(0, np.uint64)).size == 0
    with pytest.raises(m.SezkpCudaError) as ei:
        single.verify_openings(np.zeros((2, 32), np.uint8), ["a", "b"], np.array([2], np.uint32), np.zeros(1, np.uint64), np.zeros(1, np.uint64))
    assert ei.value.code == -1
    assert lib.sezkp_cuda_device_count() >= 1


@pytest.mark.parametrize("devs", [[0, 0], [0] * 4, [0] * 8], ids=lambda d: "ranks%d" % len(d))
def test_coset_resident_fri_layers(single, devs):
    """One proof over a context group: the DEEP-LDE stays split by coset, the large FRI layers are folded locally and only
    each rank's hashing range is materialised (fri_gather_ranges_kernel reads the peers' coset arrays), FRI openings are
    served by the range owners and gathered by rank 0.  Bytes must equal the single-GPU proof, with the mode on and off,
    from host descriptors and from resident traces; 2^17 rows has one large layer (2^20), 2^19 rows three."""
    import torch
    m = pkg()
    n = torch.cuda.device_count()
    lists = [devs] + ([list(range(len(devs)))] if n >= len(devs) else [])
    for dl in lists:
        g = m.Context(devices=dl)
        try:
            for T, b, tau in [(1 << 17, 512, 2), (1 << 19, 512, 3)]:
                ct = m.simulate(T, b, tau, seed=31)
                root = m.manifest_root(ct)
                want = single.prove_v1(ct, root)
                assert g.prove_v1(ct, root) == want
                rt = g.upload_trace(ct)
                assert g.prove_v1_resident(rt, root) == want
                g.set_option("fri_coset", 0)
                assert g.prove_v1_resident(rt, root) == want
                g.set_option("fri_coset", 1)
                assert g.prove_v1_resident(rt, root) == want
                rt.free()
            # the config-4 pipeline's single-vector tail (DEEP-LDE + FRI) uses the same coset-resident layers
            cs_s, cs_g = single.columns_synth(9, 18), g.columns_synth(9, 18)
            a, b = single.lde_commit_fri(cs_s), g.lde_commit_fri(cs_g)
            g.set_option("fri_coset", 0)
            c = g.lde_commit_fri(cs_g)
            g.set_option("fri_coset", 1)
            cs_s.free(), cs_g.free()
            for x in (b, c):
                assert np.array_equal(a[0], x[0]) and np.array_equal(a[1], x[1]) and a[2] == x[2]
        finally:
            g.close()
