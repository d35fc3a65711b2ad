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


def _prototypes():
    """name -> list of C parameter types, parsed from the header."""
    hdr = open(os.path.join(ROOT, "include", "rpo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|const char\*)\s+(rpo_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, [re.sub(r"\s*\b[A-Za-z_][A-Za-z0-9_]*$", "", p).strip() for p in params])
    return protos


def test_ctypes_signatures_match_the_header(built):
    """Every row of the ctypes table has the header's arity and a compatible type in every position."""
    import ctypes
    kinds = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float}
    protos = _prototypes()
    assert sorted(protos) == sorted(built.SIGNATURES)
    for name, (ret, params) in protos.items():
        restype, argtypes = built.SIGNATURES[name]
        assert restype is (ctypes.c_int if ret == "int" else ctypes.c_char_p), name
        assert len(argtypes) == len(params), (name, params, argtypes)
        for c_type, ct in zip(params, argtypes):
            if "*" in c_type:
                assert ct is ctypes.c_void_p, (name, c_type)
            else:
                assert ct is kinds[c_type], (name, c_type, ct)


def test_generated_header_is_up_to_date(tmp_path):
    """rpo_b200/csrc/evopf_gen.cuh is committed; it must be exactly what tools/gen_evopf.py emits today."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_evopf", os.path.join(ROOT, "tools", "gen_evopf.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    txt = gen.Gen().generate()
    assert txt == open(os.path.join(ROOT, "rpo_b200", "csrc", "evopf_gen.cuh")).read()


def test_box_argument_validation_without_gpu(built):
    lib = built.load()
    P = 1                                                     # any non-null pointer value: rejected before it is used
    assert lib.rpo_box_fwd(P, P, P, 14, P, 57, 5, P, 9, 4, None) == -1 and b"mode" in lib.rpo_last_error()
    assert lib.rpo_box_fwd(P, P, P, 14, P, 57, 3, P, 0, 4, None) == -1 and b"n_vol" in lib.rpo_last_error()
    assert lib.rpo_box_fwd(P, P, P, 14, None, 57, 5, P, 0, 4, None) == -1          # volatile box without soc
    assert lib.rpo_box_fwd(P, P, P, 0, None, 0, 0, P, 0, 4, None) == -1            # width
    assert lib.rpo_box_fwd(None, P, P, 14, P, 57, 5, P, 0, 4, None) == -1          # x missing for an x-mode
    assert lib.rpo_box_bwd(P, P, P, P, 14, P, 57, 5, P, 4, 4, None) == -1          # no backward for the query modes
    assert lib.rpo_box_fwd(P, P, P, 14, P, 57, 5, P, 0, 0, None) == 0              # empty batch


def test_round2_entry_points_validate_their_arguments_without_gpu(built):
    """Argument checks of the entry points added in round 2 run before any CUDA call: exercised here without a GPU."""
    lib = built.load()
    A = 0x1000                                                # a 16-byte aligned non-null pointer value (never dereferenced)
    # alignment contract of the single-pass classic entry points (float2 / float4 accesses)
    assert lib.rpo_cart_direction(A + 4, A, 8, 0, None) == -1 and b"8-byte aligned" in lib.rpo_last_error()
    assert lib.rpo_cart_resid(A + 4, A, A, 8, None) == -1
    assert lib.rpo_cart_complete_bwd(A + 4, A, 8, 1, None) == -1
    assert lib.rpo_pen_step(A + 8, A, A, A, A, A, None, None, 8, None) == -1 and b"16-byte aligned" in lib.rpo_last_error()
    assert lib.rpo_pen_project(A, A, A + 4, A, 8, 2e-3, 0.0, 1e-5, 10, None, None) == -1
    # residual backward / fused backward
    assert lib.rpo_evopf_resid_bwd(A, 57, A, None, A, 2, A, 8, None) == -1 and b"ineq_mode" in lib.rpo_last_error()
    assert lib.rpo_evopf_resid_bwd(A, 20, A, A, None, 0, A, 8, None) == -1 and b"ld_state" in lib.rpo_last_error()
    assert lib.rpo_evopf_resid_bwd(None, 57, A, A, None, 0, A, 8, None) == -1
    assert lib.rpo_evopf_complete_bwd_penalty(A, 57, A, A, None, None, A, None, 8, None) == -1          # dual weights missing
    assert lib.rpo_evopf_complete_bwd_penalty(A, 30, A, A, A, None, A, None, 8, None) == -1             # no state of charge in the row
    assert lib.rpo_evopf_complete_bwd_squashed(A, 57, A, A, A, A, A, 5, None, None, A, None, 8, None) == -1 and b"box mode" in lib.rpo_last_error()
    assert lib.rpo_evopf_complete_bwd_squashed(A, 57, A, A, None, A, A, 1, None, None, A, None, 8, None) == -1
    assert lib.rpo_evopf_project_squashed(A, 57, A, None, A, 1, None, A, 8, 1e-5, 50, 1e-4, 0.0, 1e-5, 10, None, None, None) == -1
    assert lib.rpo_evopf_project_squashed(A, 30, A, A, A, 1, None, A, 8, 1e-5, 50, 1e-4, 0.0, 1e-5, 10, None, None, None) == -1
    assert lib.rpo_evopf_complete_fwd_squashed(A, 57, A, A, A, 7, A, 8, 1e-5, 50, None, None) == -1
    assert lib.rpo_evopf_complete_fwd_plain(A, 20, A, A, 8, 1e-5, 50, None, None) == -1
    assert lib.rpo_cart_resid_bwd(None, A, A, 0, A, 8, None) == -1                                      # g_ineq needs the action
    assert lib.rpo_pen_resid_bwd(A, A, A, None, 3, A, 8, None) == -1
    # empty batches are a no-op everywhere
    assert lib.rpo_evopf_resid_bwd(A, 57, A, A, A, 1, A, 0, None) == 0
    assert lib.rpo_evopf_complete_bwd_squashed(A, 57, A, A, A, A, A, 1, None, None, A, None, 0, None) == 0
