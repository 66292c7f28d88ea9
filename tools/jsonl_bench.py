"""Streaming-JSONL leg of bench.py alone, at one or more sizes: tools/jsonl_bench.py [log_T ...] (default 21 24).
Prints one JSON object per size (the fields of bench.py's `jsonl_stream`).  The file goes to tempfile's directory: set
TMPDIR=/dev/shm to measure a tmpfs-backed file (every page of the mapping is faulted and unmapped individually there)."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

m = importlib.import_module(bench.PKG)
ctx = m.Context(0)
for lt in [int(a) for a in sys.argv[1:]] or [21, 24]:
    os.environ["SEZKP_JSONL_LOG_T"] = str(lt)
    out = bench.jsonl_stream_bench(torch, ctx, m, 3)
    print(json.dumps(out), flush=True)
