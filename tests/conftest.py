import ctypes as C
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pkg():
    return importlib.import_module("streaming-zero-knowledge-proofs_b200")


@pytest.fixture(scope="session")
def sezkp():
    return pkg()


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand from oracle/ with g++."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    return oracle_lib.load()


def load_fixture(name):
    return json.load(open(os.path.join(GOLDEN, name)))
