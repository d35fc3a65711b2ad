"""ctypes binding of librpo_b200.so (the C ABI declared in include/rpo_b200.h).

There is no CPU fallback: if the shared library is missing or was not built for sm_100a, importing the
ops raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C rpo_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# RPO_B200_LIB overrides the path for kernel-variant experiments (tools/variant_sweep.py); the product uses the in-tree build.
LIB_PATH = os.environ.get("RPO_B200_LIB") or os.path.join(_HERE, "librpo_b200.so")

RPO_OK = 0

_P = c_void_p

# name -> (restype, argtypes); mirrors include/rpo_b200.h one to one
SIGNATURES = {
    "rpo_version": (c_int, []),
    "rpo_last_error": (c_char_p, []),
    "rpo_build_arch": (c_char_p, []),
    "rpo_evopf_complete_fwd": (c_int, [_P, c_int, _P, _P, c_int64, c_float, c_int, _P, _P]),
    "rpo_evopf_complete_fwd_plain": (c_int, [_P, c_int, _P, _P, c_int64, c_float, c_int, _P, _P]),
    "rpo_evopf_complete_bwd": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "rpo_evopf_direction": (c_int, [_P, c_int, _P, _P, c_int64, _P]),
    "rpo_evopf_correct": (c_int, [_P, c_int, _P, _P, c_int64, c_float, c_float, c_float, c_int, _P, _P]),
    "rpo_evopf_project": (c_int, [_P, c_int, _P, _P, _P, c_int64, c_float, c_int, c_float, c_float, c_float, c_int,
                                  _P, _P, _P]),
    "rpo_evopf_resid": (c_int, [_P, c_int, _P, _P, _P, c_int64, _P]),
    "rpo_evopf_resid_bwd": (c_int, [_P, c_int, _P, _P, _P, c_int, _P, c_int64, _P]),
    "rpo_evopf_complete_bwd_penalty": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "rpo_evopf_project_squashed": (c_int, [_P, c_int, _P, _P, _P, c_int, _P, _P, c_int64, c_float, c_int, c_float, c_float,
                                           c_float, c_int, _P, _P, _P]),
    "rpo_evopf_complete_fwd_squashed": (c_int, [_P, c_int, _P, _P, _P, c_int, _P, c_int64, c_float, c_int, _P, _P]),
    "rpo_evopf_complete_bwd_squashed": (c_int, [_P, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_int64, _P]),
    "rpo_evopf_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "rpo_cart_complete_fwd": (c_int, [_P, _P, c_int64, c_int, _P]),
    "rpo_cart_complete_bwd": (c_int, [_P, _P, c_int64, c_int, _P]),
    "rpo_cart_direction": (c_int, [_P, _P, c_int64, c_int, _P]),
    "rpo_cart_correct": (c_int, [_P, _P, c_int64, c_int, c_float, c_float, c_float, c_int, _P, _P]),
    "rpo_cart_project": (c_int, [_P, _P, _P, c_int64, c_int, c_float, c_float, c_float, c_int, _P, _P]),
    "rpo_cart_resid": (c_int, [_P, _P, _P, c_int64, _P]),
    "rpo_cart_resid_bwd": (c_int, [_P, _P, _P, c_int, _P, c_int64, _P]),
    "rpo_cart_step": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, _P]),
    "rpo_pen_complete_fwd": (c_int, [_P, _P, _P, c_int64, _P]),
    "rpo_pen_complete_bwd": (c_int, [_P, _P, _P, c_int64, _P]),
    "rpo_pen_direction": (c_int, [_P, _P, _P, c_int64, _P]),
    "rpo_pen_correct": (c_int, [_P, _P, _P, c_int64, c_float, c_float, c_float, c_int, _P, _P]),
    "rpo_pen_project": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_int, _P, _P]),
    "rpo_pen_resid": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "rpo_pen_resid_bwd": (c_int, [_P, _P, _P, _P, c_int, _P, c_int64, _P]),
    "rpo_pen_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "rpo_box_fwd": (c_int, [_P, _P, _P, c_int, _P, c_int, c_int, _P, c_int, c_int64, _P]),
    "rpo_box_bwd": (c_int, [_P, _P, _P, _P, c_int, _P, c_int, c_int, _P, c_int, c_int64, _P]),
    "rpo_copy_rows_h2d": (c_int, [_P, c_int, _P, c_int, c_int, c_int64, _P]),
    "rpo_copy_rows_d2h": (c_int, [_P, c_int, _P, c_int, c_int, c_int64, _P]),
    "rpo_fp32_peak_probe": (c_int, [_P, c_int, c_int, c_int, _P]),
}

_lib = None


class RpoLibraryError(RuntimeError):
    pass


def _try_build():
    """The .so is a build artefact (git-ignored).  If it is missing but the CUDA toolchain is present, compile it for
    sm_100a exactly as __graft_entry__.build() does.  This is a build step, not a fallback: without nvcc we still raise."""
    import shutil
    import subprocess
    nvcc = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(_HERE, "csrc", "rpo_kernels.cu")
    if not (os.path.isfile(nvcc) and os.path.isfile(src)):
        return
    cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-DRPO_BUILD_ARCH=\"sm_100a\"", "-shared", "-o", LIB_PATH, src, "-lcudart"]
    try:
        subprocess.check_call(cmd)
    except Exception:
        pass


def load():
    """Load the shared library (once) and bind every symbol of the header; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH) and not os.environ.get("RPO_B200_LIB") and not os.environ.get("RPO_B200_NO_AUTOBUILD"):
        _try_build()
    if not os.path.isfile(LIB_PATH):
        raise RpoLibraryError(
            "%s not found: the CUDA extension is not built (run __graft_entry__.build() or make -C rpo_b200/csrc). "
            "rpo_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RpoLibraryError("symbol %s missing from %s" % (name, LIB_PATH)) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != RPO_OK:
        msg = load().rpo_last_error()
        raise RpoLibraryError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
