"""A non-Python caller of the C ABI: tools/c_driver.c links libsezkp_cuda.so directly (gcc, no torch, no ctypes) and checks
the SURVEY §8c known answers; the proof it writes must equal the one the Python binding returns."""
import os
import subprocess

import pytest

from conftest import ROOT, pkg

EXE = os.path.join(ROOT, "tools", "c_driver")


def build_driver():
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(EXE + ".c"):
        import __graft_entry__
        __graft_entry__.build()


def test_c_driver_links_and_reports_enodev_without_gpu():
    import torch
    build_driver()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and r.stdout.startswith("ENODEV"), r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n_gpus", [1, 2])
def test_c_driver_known_answers_and_proof_bytes(tmp_path, oracle, n_gpus):
    build_driver()
    m = pkg()
    out = str(tmp_path / "proof.bin")
    r = subprocess.run([EXE, str(n_gpus), out, "same"], capture_output=True, text=True, timeout=300)  # "same": both ranks on GPU 0
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
    assert f"gpus={n_gpus}" in r.stdout
    proof = open(out, "rb").read()
    assert len(proof) == 197199
    assert oracle.blake3(proof).hex() == "733e0d80c093a79a7dab5464741660950997b31c1f14ab255177f120827f2d43"  # SURVEY §8c †KAT
    assert proof == m.Context().prove_v1(m.demo_block(64), bytes([7]) * 32)
