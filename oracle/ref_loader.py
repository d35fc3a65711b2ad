"""Import the UNMODIFIED reference (`/root/reference/rpo`) through the stub modules in oracle/stubs.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  In the build container the tree is /root/reference; on the GPU box
(where that path does not exist) it is the verbatim copy oracle/make_ref_copy.sh placed under oracle/_ref/ (git-ignored,
shipped by gpurun, SHA-256 list of the original beside it).  Used to (i) validate the torch restatement in
oracle/rpo_oracle.py, (ii) generate the golden vectors in tests/golden/ (oracle/make_golden.py), (iii) run the reference's
own RPODDPG / RPOSAC classes over the CUDA kernels in the GPU tests and (iv) time the reference's CPU path as the
baseline of bench.py.  Nothing under rpo_b200/ imports it.
"""
from __future__ import annotations

import os
import sys
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")
_REPO = os.path.dirname(_HERE)
# the verbatim copy oracle/make_ref_copy.sh places for the GPU box (git-ignored, shipped by gpurun)
REF_COPY = os.path.join(_HERE, "_ref")


def _has_tree(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "rpo", "env", "electrical_grid", "evopf.py"))


def _pick_root() -> str:
    env = os.environ.get("RPO_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if _has_tree("/root/reference") else REF_COPY


REFERENCE_ROOT = _pick_root()
DATA_DIR = os.path.join(REFERENCE_ROOT, "rpo", "env", "electrical_grid", "data", "crawlers")


def reference_available() -> bool:
    return _has_tree(REFERENCE_ROOT)


def verify_copy(root: str = REF_COPY) -> int:
    """Check every file of a make_ref_copy.sh copy against the SHA-256 list written from the original tree: proves the
    copy that runs on the GPU box is the unmodified reference.  Returns the number of files checked."""
    import hashlib
    n = 0
    with open(os.path.join(root, "SHA256SUMS")) as fh:
        for line in fh:
            digest, rel = line.split(None, 1)
            rel = rel.strip().lstrip("*")
            with open(os.path.join(root, rel), "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != digest:
                    raise RuntimeError("reference copy differs from the original: %s" % rel)
            n += 1
    return n


def load_reference():
    """Return the reference's top-level ``rpo`` module (cpu device forced at env construction)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, "bool"):
        np.bool = bool  # reference uses the alias removed in numpy>=1.24 (ddpg_pa.py:71)
    for p in (_REPO, _STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # stubs must shadow nothing real: gym/pypower/igraph are absent in this image
    import rpo  # noqa: F401  (the reference package)
    import rpo.env  # noqa: F401
    import rpo.algo  # noqa: F401
    return rpo


def make_cpu_env(name: str):
    """Construct a reference env with its device pinned to CPU (envs pick cuda if available)."""
    import torch
    rpo = load_reference()
    with mock.patch.object(torch.cuda, "is_available", return_value=False):
        if name == "evopf":
            return rpo.env.EVOPFEnv()
        if name == "cart":
            return rpo.env.CartSafeEnv()
        if name == "pen":
            return rpo.env.SpringPendulumEnv()
    raise KeyError(name)
