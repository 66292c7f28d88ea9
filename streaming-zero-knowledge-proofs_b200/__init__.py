"""sezkp-b200: B200-native STARK v1 commitment path for SEZKP (host-side mirror of the reference API).

The package directory name carries a hyphen (it mirrors the upstream repository name), so import it with
``importlib.import_module("streaming-zero-knowledge-proofs_b200")`` (tests/conftest.py and bench.py do exactly that).
"""
from .trace import CompactTrace, TraceDesc, blocks_to_compact, demo_block, manifest_root, simulate, simulate_exact  # noqa: F401

from .binding import Context, SezkpCudaError, load_library, EXPORTS, LIB_PATH  # noqa: F401,E402
from .backend import ProofArtifact, StarkV1Cuda  # noqa: F401,E402
from . import artifact, io_jsonl, parallel  # noqa: F401,E402

__all__ = ["Context", "SezkpCudaError", "StarkV1Cuda", "ProofArtifact", "CompactTrace", "TraceDesc", "blocks_to_compact", "demo_block", "manifest_root", "simulate", "simulate_exact"]
