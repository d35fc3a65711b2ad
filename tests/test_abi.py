"""The C-ABI shared library loads on a box without a GPU and exports every symbol include/rpo_b200.h declares.
No compute calls here (those are the -m gpu tests)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "rpo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rpo_[a-z0-9_]+)\s*\(", hdr)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from rpo_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    names = _declared()
    assert len(names) >= 25
    lib = built.load()
    assert sorted(built.SIGNATURES) == names, "ctypes table and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", built.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    for n in names:
        assert n in exported, n
        assert hasattr(lib, n)


def test_metadata_calls(built):
    lib = built.load()
    assert lib.rpo_version() >= 100
    assert lib.rpo_build_arch() == b"sm_100a"


def test_sass_is_sm100a_only(built):
    out = subprocess.check_output(["cuobjdump", "-lelf", built.LIB_PATH]).decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_without_gpu(built):
    """Invalid arguments are rejected before any launch (error convention of the header)."""
    lib = built.load()
    assert lib.rpo_evopf_complete_fwd(None, 57, None, None, 4, 1e-5, 50, None, None) == -1
    assert b"null" in lib.rpo_last_error()
    assert lib.rpo_evopf_complete_fwd(None, 57, None, None, -1, 1e-5, 50, None, None) == -1
    assert lib.rpo_evopf_complete_fwd(None, 57, None, None, 0, 1e-5, 50, None, None) == 0      # empty batch is a no-op
    assert lib.rpo_cart_complete_fwd(1, 1, 4, 2, None) == -1                                   # partial_idx must be 0/1


def test_ops_fail_loudly_on_cpu_tensors(built):
    import torch
    from rpo_b200 import ops
    with pytest.raises(built.RpoLibraryError):
        ops.evopf_complete_fwd(torch.zeros(2, 57), torch.zeros(2, 14))
    with pytest.raises(built.RpoLibraryError):
        ops.cart_complete_fwd(torch.zeros(2, 1), 0)


def test_missing_library_is_an_error(built, monkeypatch):
    monkeypatch.setattr(built, "_lib", None)
    monkeypatch.setattr(built, "LIB_PATH", "/nonexistent/librpo_b200.so")
    with pytest.raises(built.RpoLibraryError):
        built.load()
