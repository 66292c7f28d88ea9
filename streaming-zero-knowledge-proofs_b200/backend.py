"""Host-side mirror of the reference's backend façade for the STARK v1 path.

Reference: ``trait ProvingBackend { prove(blocks, manifest_root) -> ProofArtifact; verify(..) }``
(crates/sezkp-core/src/backend.rs:41-61) implemented by ``StarkV1`` (crates/sezkp-stark/src/lib.rs:126-190).
``StarkV1Cuda.prove`` returns the same artifact (backend "stark", proof_bytes = bincode(ProofV1),
meta {"domain_n","proto":"stark-v1","tau"} with alphabetically ordered keys as serde_json emits them).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Union

from .binding import Context, SezkpCudaError
from .trace import CompactTrace, blocks_to_compact


@dataclass
class ProofArtifact:  # crates/sezkp-core/src/artifact.rs:34-68
    backend: str
    manifest_root: bytes
    proof_bytes: bytes
    meta: dict = field(default_factory=dict)


class StarkV1Cuda:
    """Stateless like the reference backend; a process-wide context is created on first use."""

    _ctx: Optional[Context] = None

    @classmethod
    def context(cls) -> Context:
        if cls._ctx is None:
            cls._ctx = Context()
        return cls._ctx

    @classmethod
    def _compact(cls, blocks: Union[CompactTrace, Sequence[dict]]) -> CompactTrace:
        return blocks if isinstance(blocks, CompactTrace) else blocks_to_compact(blocks)

    @classmethod
    def prove(cls, blocks, manifest_root: bytes) -> ProofArtifact:
        ct = cls._compact(blocks)
        proof = cls.context().prove_v1(ct, bytes(manifest_root))
        return ProofArtifact("stark", bytes(manifest_root), proof,
                             {"domain_n": ct.n_rows * 8, "proto": "stark-v1", "tau": ct.tau})

    @classmethod
    def prove_streaming(cls, blocks, manifest_root: bytes) -> ProofArtifact:  # crates/sezkp-stark/src/lib.rs:170-190
        art = cls.prove(blocks, manifest_root)
        art.meta = {"domain_n": art.meta["domain_n"], "mode": "streaming", "proto": "stark-v1", "tau": art.meta["tau"]}
        return art
