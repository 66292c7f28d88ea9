"""Native JSONL front-end (csrc/jsonl.cpp) against the Python restatement of the reference's reader
(`stream_block_summaries_jsonl`, crates/sezkp-core/src/io_jsonl.rs:27-88): same blocks, same order, blank lines
skipped, malformed lines rejected with their line number.  Host logic only — no GPU, no compute calls."""
import json
import os

import numpy as np
import pytest

from conftest import pkg

FIELDS = ("block_len", "win_left", "win_right", "head_in_off", "head_out_off", "input_mv", "mv", "write_flag", "write_sym",
          "version", "block_id", "step_lo", "step_hi", "ctrl_in", "ctrl_out", "in_head_in", "in_head_out")


def same(a, b):
    assert a.tau == b.tau
    for f in FIELDS:
        x, y = np.asarray(getattr(a, f)), np.asarray(getattr(b, f))
        assert x.shape == y.shape and (x.astype(np.int64) == y.astype(np.int64)).all(), f


def jsonl_bytes(m, ct, tmp_path, name="t.jsonl"):
    p = str(tmp_path / name)
    m.io_jsonl.write_jsonl(p, ct)
    return open(p, "rb").read()


@pytest.mark.parametrize("T,b,tau,threads", [(64, 16, 1, 1), (512, 64, 3, 2), (4096, 512, 8, 8), (1000, 37, 2, 5), (8192, 512, 8, 0)])
def test_native_parser_matches_python_reader(tmp_path, T, b, tau, threads):
    m = pkg()
    ct = m.simulate(T, b, tau, seed=7 + T)
    text = jsonl_bytes(m, ct, tmp_path)
    got = m.binding.parse_jsonl(text, threads)
    same(got, ct)
    ref = m.blocks_to_compact([json.loads(l) for l in text.decode().splitlines() if l.strip()])
    same(got, ref)
    assert m.manifest_root(got) == m.manifest_root(ct)


def test_blank_lines_whitespace_key_order_and_unknown_keys(tmp_path):
    m = pkg()
    ct = m.simulate(256, 32, 2, seed=3)
    lines = jsonl_bytes(m, ct, tmp_path).decode().splitlines()
    out = []
    for i, l in enumerate(lines):
        d = json.loads(l)
        if i % 2:
            d = dict(reversed(list(d.items())))  # serde accepts any key order
            d["movement_log"] = {"note": "x\\\"}]", "steps": d["movement_log"]["steps"]}
            d["future_field"] = {"a": [1.5e3, True, None, "s{"], "b": {}}
        out.append(json.dumps(d, indent=(None if i % 3 else 1)).replace("\n", " ") if i % 3 == 0 else json.dumps(d))
        if i % 4 == 0:
            out.append("   \t ")
            out.append("")
    text = ("\r\n".join(out) + "\n\n").encode()
    same(m.binding.parse_jsonl(text, 3), ct)
    same(m.binding.parse_jsonl(text.rstrip(b"\n"), 1), ct)  # last line without a newline


def test_empty_input():
    m = pkg()
    got = m.binding.parse_jsonl(b"", 2)
    assert got.n_blocks == 0 and got.n_rows == 0
    got = m.binding.parse_jsonl(b"\n  \n", 2)
    assert got.n_blocks == 0


@pytest.mark.parametrize("mutate,needle", [
    (lambda d: d.pop("windows"), "missing"),
    (lambda d: d.__setitem__("step_hi", d["step_hi"] + 1), "movement_log length"),
    (lambda d: d["movement_log"]["steps"][1]["tapes"].pop(), "tapes"),
    (lambda d: d["movement_log"]["steps"][0].__setitem__("input_mv", 1.0), "float"),
    (lambda d: d["movement_log"]["steps"][0]["tapes"][0].__setitem__("mv", 300), "out of range"),
    (lambda d: d["head_in_offsets"].append(0), "tau"),
    (lambda d: d["movement_log"]["steps"][0]["tapes"][0].__setitem__("write", -1), "out of range"),
])
def test_malformed_lines_are_rejected_with_line_number(tmp_path, mutate, needle):
    m = pkg()
    ct = m.simulate(128, 32, 2, seed=5)
    lines = jsonl_bytes(m, ct, tmp_path).decode().splitlines()
    d = json.loads(lines[2])
    mutate(d)
    lines[2] = json.dumps(d)
    with pytest.raises(m.SezkpCudaError) as ei:
        m.binding.parse_jsonl(("\n".join(lines) + "\n").encode(), 2)
    assert ei.value.code == -1 and "line 3" in str(ei.value) and needle in str(ei.value), str(ei.value)


def test_truncated_and_trailing_garbage(tmp_path):
    m = pkg()
    ct = m.simulate(64, 32, 1, seed=9)
    text = jsonl_bytes(m, ct, tmp_path)
    first = text.split(b"\n")[0]
    for bad in (first[:-1], first + b" x", first[: len(first) // 2], b"[1,2]", b"{\"step_lo\":1"):
        with pytest.raises(m.SezkpCudaError) as ei:
            m.binding.parse_jsonl(bad + b"\n", 1)
        assert "line 1" in str(ei.value)


def test_block_scalars_layout():
    m = pkg()
    assert m.binding.BLOCK_SCALARS_DTYPE.itemsize == 48


def test_errors_in_a_multi_worker_parse_report_the_first_bad_line(tmp_path):
    """text large enough for several workers (>= 64 KiB each): the earliest malformed line in file order is reported,
    with its line number counted across blank lines; a tau change between workers' ranges is caught as well."""
    m = pkg()
    ct = m.simulate(8192, 512, 8, seed=21)
    lines = jsonl_bytes(m, ct, tmp_path).decode().splitlines()
    assert len(lines) == 16 and sum(map(len, lines)) > (1 << 20)
    bad = list(lines)
    bad.insert(3, "")                      # blank line: counted, not parsed
    bad[11] = bad[11].replace('"mv":', '"mv":1.5e0,"x":', 1)
    bad[14] = bad[14][:-1]                 # truncated too, but later in the file
    with pytest.raises(m.SezkpCudaError) as ei:
        m.binding.parse_jsonl(("\n".join(bad) + "\n").encode(), 6)
    assert "line 12:" in str(ei.value) and "float" in str(ei.value), str(ei.value)
    other = m.simulate(512, 512, 3, seed=1)
    odd = jsonl_bytes(m, other, tmp_path, "o.jsonl").decode().splitlines()[0]
    mixed = lines[:13] + [odd] + lines[13:]
    with pytest.raises(m.SezkpCudaError) as ei:
        m.binding.parse_jsonl(("\n".join(mixed) + "\n").encode(), 8)
    assert "tau" in str(ei.value) or "tapes" in str(ei.value), str(ei.value)


def test_native_writer_matches_python_writer_and_round_trips(tmp_path):
    """sezkp_jsonl_write_file (the export-jsonl direction) emits byte-for-byte what io_jsonl.write_jsonl does, also for ragged
    blocks and non-power-of-two traces, and the native parser reads it back."""
    import filecmp
    from importlib import import_module
    m = pkg()
    b = import_module("streaming-zero-knowledge-proofs_b200.binding")
    for T, blk, tau, threads in ((1 << 12, 512, 3, 4), (1000, 300, 2, 3), (64, 64, 1, 1), (5000, 7, 8, 16)):
        ct = m.simulate(T, blk, tau, seed=T)
        pa, pb = str(tmp_path / "a.jsonl"), str(tmp_path / "b.jsonl")
        m.io_jsonl.write_jsonl(pa, ct)
        n = b.write_jsonl_native(pb, ct, threads)
        assert n == os.path.getsize(pa) and filecmp.cmp(pa, pb, shallow=False)
        back = b.parse_jsonl(open(pb, "rb").read(), threads)
        assert np.array_equal(back.mv, ct.mv) and np.array_equal(back.write_sym, ct.write_sym) and np.array_equal(back.block_len, ct.block_len)


def test_bulk_fast_path_boundaries(tmp_path):
    """The bulk steps path (fast_steps in csrc/jsonl.cpp) accepts only serde-exact bytes with 1-2 digit symbols and one-digit
    moves in its branch-free form; everything else inside an otherwise exact line must fall through to the slower forms and
    still give the same arrays or the same errors: 3-5 digit symbols, |mv| up to 128, "-0", whitespace in the middle of the
    steps array, a steps array that ends the line (no slack bytes behind the last op), leading zeros, floats."""
    m = pkg()
    ct = m.simulate(256, 64, 3, seed=11)
    lines = jsonl_bytes(m, ct, tmp_path).decode().splitlines()
    docs = [json.loads(l) for l in lines]
    rows = [s for d in docs for s in d["movement_log"]["steps"]]
    # unusual but valid numbers, spread over many steps
    edits = {5: (0, 65535, -128), 6: (1, 100, 127), 70: (2, 9999, -1), 71: (0, 12345, 0), 130: (1, 15, 100), 255: (2, 65535, -128)}
    for i, (r, sym, mv) in edits.items():
        rows[i]["tapes"][r] = {"write": sym, "mv": mv}
    text = "\n".join(json.dumps(d, separators=(",", ":")) for d in docs) + "\n"
    assert '"write":65535,"mv":-128' in text
    ref = m.blocks_to_compact(docs)
    same(m.binding.parse_jsonl(text.encode(), 1), ref)
    # "-0" is a valid integer; whitespace after a comma in the middle of the steps array; both must parse to the same arrays
    t2 = text.replace('"mv":0}', '"mv":-0}', 3).replace('},{"input_mv"', '}, {"input_mv"', 2).replace('},{"write"', '} ,{"write"', 2)
    assert t2 != text
    same(m.binding.parse_jsonl(t2.encode(), 1), ref)
    # movement_log as the LAST key: the steps array is followed by "]}}" only, so the last ops have no slack bytes behind them
    tail = []
    for d in docs:
        d2 = {k: v for k, v in d.items() if k != "movement_log"}
        d2["movement_log"] = d["movement_log"]
        tail.append(json.dumps(d2, separators=(",", ":")))
    assert all(l.endswith("]}}") for l in tail)
    same(m.binding.parse_jsonl(("\n".join(tail)).encode(), 1), ref)
    # malformed numbers inside exact bytes are still rejected, with the reference's line number
    for old, new, needle in (('"mv":1}', '"mv":01}', "leading zero"), ('"mv":1}', '"mv":1.0}', "float"), ('"mv":1}', '"mv":128}', "out of range"),
                             ('"write":null', '"write":nul', None), ('"write":null', '"write":65536', "out of range"),
                             ('"input_mv":', '"input_mv":-129,"x":', "out of range")):
        ls = text.splitlines()
        assert old in ls[1]
        ls[1] = ls[1].replace(old, new, 1)
        with pytest.raises(m.SezkpCudaError) as ei:
            m.binding.parse_jsonl(("\n".join(ls) + "\n").encode(), 1)
        assert "line 2" in str(ei.value) and (needle is None or needle in str(ei.value)), str(ei.value)


def test_parser_under_sanitizers(tmp_path):
    """tools/jsonl_fuzz.cpp: the parser compiled with -fsanitize=address,undefined parses ~7000 truncated / corrupted variants
    of a small file from exact-size heap buffers; any out-of-bounds read of the fixed-width fast path aborts the run."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    m = pkg()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, m.__name__, "csrc")
    exe = str(tmp_path / "jsonl_fuzz")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         "-I", csrc, os.path.join(root, "tools", "jsonl_fuzz.cpp"), os.path.join(csrc, "jsonl.cpp"), "-o", exe],
                        capture_output=True, text=True)
    if cc.returncode != 0 and ("asan" in cc.stderr or "sanitize" in cc.stderr):
        pytest.skip("sanitizer runtime not installed")
    assert cc.returncode == 0, cc.stderr[-2000:]
    ct = m.simulate(4096, 512, 8, seed=5)
    path = str(tmp_path / "f.jsonl")
    m.io_jsonl.write_jsonl(path, ct)
    run = subprocess.run([exe, path], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "parsed ok" in run.stdout and "piecewise pool parse ok" in run.stdout, (run.stdout + run.stderr)[-3000:]
    # the persistent worker pool + output arrays reused from piece to piece (the prover's file path) under ThreadSanitizer
    exe_t = str(tmp_path / "jsonl_fuzz_tsan")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-I", csrc, os.path.join(root, "tools", "jsonl_fuzz.cpp"),
                         os.path.join(csrc, "jsonl.cpp"), "-o", exe_t], capture_output=True, text=True)
    if cc.returncode != 0 and ("tsan" in cc.stderr or "sanitize" in cc.stderr):
        return  # no TSan runtime: the ASan run above already covered the same code single-checked
    assert cc.returncode == 0, cc.stderr[-2000:]
    run = subprocess.run([exe_t, path, "pool"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "piecewise pool parse ok" in run.stdout and "ThreadSanitizer" not in run.stderr, (run.stdout + run.stderr)[-3000:]
