"""The header-only C++ adaptor (dune-iga_b200/host/igab200_adaptor.hh) that keeps the reference's method names:
compiled on the CPU box (not gpu), executed against the oracle on the GPU box (gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_adaptor")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-c", os.path.join(ROOT, "oracle", "igab_oracle.c"), "-o",
                           os.path.join(os.path.dirname(EXE), "igab_oracle.o")], env=env)
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "cpp", "test_adaptor.cpp"),
                           os.path.join(os.path.dirname(EXE), "igab_oracle.o"), "-o", EXE,
                           "-L" + os.path.join(ROOT, "dune-iga_b200"), "-ligab200",
                           "-Wl,-rpath," + os.path.join(ROOT, "dune-iga_b200"), "-Wl,-rpath,/usr/local/cuda/lib64",
                           "-L/usr/local/cuda/lib64", "-lm"], env=env)


def test_adaptor_compiles_against_the_c_abi():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_adaptor_quadrature_loop_matches_oracle():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "adaptor test: ok" in r.stdout
