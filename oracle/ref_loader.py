"""Import the UNMODIFIED reference (`/root/reference/rpo`) through the stub modules in oracle/stubs.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container); it is used
to (i) validate the torch restatement in oracle/rpo_oracle.py and (ii) generate the golden vectors in
tests/golden/ (oracle/make_golden.py).  Nothing that runs on the GPU box calls this.
"""
from __future__ import annotations

import os
import sys
from unittest import mock

REFERENCE_ROOT = os.environ.get("RPO_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rpo", "env", "electrical_grid", "evopf.py"))


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
