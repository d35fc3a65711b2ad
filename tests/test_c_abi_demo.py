"""The C ABI is usable from plain C (no torch, no Python): examples/c_abi_demo.c is compiled with gcc against
include/rpo_b200.h + librpo_b200.so; on a GPU it projects the IEEE-14 operating point and checks the published
power-flow solution and the reference's loop-test quantities."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "c_abi_demo")


def _compile():
    from rpo_b200 import _lib
    _lib.load()                                   # builds librpo_b200.so if it is missing
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", EXE, "-L" + os.path.join(ROOT, "rpo_b200"), "-lrpo_b200",
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "rpo_b200")]
    subprocess.check_call(cmd)


def test_header_compiles_as_c_and_links():
    """The header is C (not C++) clean under -Wall -Werror and every symbol the demo uses resolves at link time."""
    _compile()
    assert os.path.isfile(EXE)


@pytest.mark.gpu
def test_c_demo_runs_on_gpu():
    _compile()
    out = subprocess.run([EXE, "8192"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout and "sm_100a" in out.stdout
