"""ProofArtifact / CommitManifest wire formats, so GPU-produced proofs are consumable by the unchanged `sezkp-cli verify`.

Reference: ``ProofArtifact`` (crates/sezkp-core/src/artifact.rs:34-68) written by ``write_proof_auto``
(crates/sezkp-core/src/io.rs:146-240): CBOR (ciborium) or JSON by file extension.  Serde details that matter:
``backend`` is the lowercase variant name, ``manifest_root`` a 32-element array, ``proof_bytes`` a ``Vec<u8>`` WITHOUT
``serde_bytes`` — i.e. an array of small integers, not a CBOR byte string — and ``meta`` a serde_json map whose keys are
emitted in alphabetical order (serde_json without ``preserve_order``).  ``CommitManifest`` is
``{version:u32=1, root:[u8;32], n_leaves:u32}`` (crates/sezkp-merkle/src/lib.rs:66-74).
"""
from __future__ import annotations

import json
from typing import Union

from .backend import ProofArtifact


def artifact_to_obj(a: ProofArtifact) -> dict:
    return {
        "backend": a.backend,
        "manifest_root": list(a.manifest_root),
        "proof_bytes": list(a.proof_bytes),
        "meta": {k: a.meta[k] for k in sorted(a.meta)},
    }


def artifact_from_obj(o: dict) -> ProofArtifact:
    return ProofArtifact(str(o["backend"]).lower(), bytes(o["manifest_root"]), bytes(o["proof_bytes"]), dict(o.get("meta") or {}))


def write_proof_auto(path: str, a: ProofArtifact) -> None:
    obj = artifact_to_obj(a)
    if path.endswith(".json"):
        with open(path, "w") as f:
            json.dump(obj, f, separators=(",", ":"))
    elif path.endswith(".cbor"):
        import cbor2
        with open(path, "wb") as f:
            cbor2.dump(obj, f)
    else:
        raise ValueError("proof path must end with .cbor or .json (reference io.rs:146-160)")


def read_proof_auto(path: str) -> ProofArtifact:
    if path.endswith(".json"):
        return artifact_from_obj(json.load(open(path)))
    if path.endswith(".cbor"):
        import cbor2
        return artifact_from_obj(cbor2.load(open(path, "rb")))
    raise ValueError("proof path must end with .cbor or .json")


def manifest_obj(root: bytes, n_leaves: int) -> dict:
    return {"version": 1, "root": list(root), "n_leaves": int(n_leaves)}
