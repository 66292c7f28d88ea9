#!/usr/bin/env python
"""Regenerate tests/golden/*.json from the reference checkout (run in the build container only).

Sources (read-only, /root/reference):
  * blocks.cbor + manifest.cbor + proof_stark.cbor            (repo root fixture, T=64, b=8, tau=2)
  * examples/minimal-riscv/{blocks,manifest,proof_stark}.cbor (T=32, b=4, tau=2)
    -> pin sezkp-merkle leaf_hash/node_hash/merkle_root and the Blake3Transcript framing (v0 proof bytes).
  * Python `blake3` 1.0.8 (PyO3 bindings to the official Rust `blake3` crate the reference depends on,
    Cargo.lock:125) -> hash / XOF known answers for the portable oracle BLAKE3.
The GPU box has no /root/reference: tests only read the JSON written here.
"""
import json
import os
import sys

import blake3
import cbor2

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def fixture(name, d):
    blocks = cbor2.load(open(os.path.join(d, "blocks.cbor"), "rb"))
    man = cbor2.load(open(os.path.join(d, "manifest.cbor"), "rb"))
    proof = cbor2.load(open(os.path.join(d, "proof_stark.cbor"), "rb"))
    for b in blocks:  # tags are advisory and unused on the path
        b.pop("pre_tags", None)
        b.pop("post_tags", None)
    out = {
        "source": os.path.relpath(d, REF) or ".",
        "blocks": blocks,
        "manifest": {"version": man["version"], "root": bytes(man["root"]).hex(), "n_leaves": man["n_leaves"]},
        "proof_v0": {
            "backend": proof["backend"],
            "manifest_root": bytes(proof["manifest_root"]).hex(),
            "proof_bytes": bytes(proof["proof_bytes"]).hex(),
            "meta": proof["meta"],
        },
    }
    json.dump(out, open(os.path.join(HERE, name), "w"), separators=(",", ":"))


def blake3_kats():
    cases = []
    for n in [0, 1, 8, 31, 63, 64, 65, 127, 128, 1023, 1024, 1025, 2048, 2049, 3072, 3073, 4096, 5000, 7168, 10000]:
        data = bytes((i * 251 + 7) % 256 for i in range(n))
        cases.append({"len": n, "out32": blake3.blake3(data).hexdigest(), "xof300": blake3.blake3(data).hexdigest(length=300)})
    json.dump({"input_rule": "byte i = (i*251+7) % 256", "blake3_module": blake3.__version__, "cases": cases},
              open(os.path.join(HERE, "blake3_kats.json"), "w"), indent=0)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; golden files are committed, nothing to do")
    fixture("fixture_root_T64.json", REF)
    fixture("fixture_riscv_T32.json", os.path.join(REF, "examples/minimal-riscv"))
    blake3_kats()
    print("golden written")
