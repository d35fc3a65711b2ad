"""Seeded synthetic inputs of each environment's shape.

EVOPF states follow the reference loaders' distributions without needing their pickled curves
(rpo/env/electrical_grid/data/demand.py:47-67, price.py:40-47): per-bus demand = case share mixed
with a Dirichlet(1) draw over the 11 load buses (rho = 0.5) times a total in [2.14, 2.95] pu,
qd = pd * tan(acos(U(0.9, 1))) * sign(Qd_case), soc ~ U(0.1, 0.8), price window ~ 0.49 (1+N)(1+0.1N).
Partial actions: D1 = uniform in the state-dependent partial box (evopf.py:769-783), D2 = its central 30 %.
"""
from __future__ import annotations

import numpy as np
import torch

from .case14 import PD, QD, case14
from . import evopf_constants as K


def evopf_states(n: int, seed: int = 0, device="cpu") -> torch.Tensor:
    rng = np.random.default_rng(seed)
    bus = case14()["bus"]
    pd_case = bus[:, PD]
    idx = np.nonzero(pd_case)[0]
    share = pd_case / pd_case.sum()
    total = rng.uniform(2.14, 2.95, size=(n, 1))
    ratio = np.repeat(share[None] * 0.5, n, axis=0)
    ratio[:, idx] += 0.5 * rng.dirichlet(np.ones(len(idx)), size=n)
    pd = ratio * total
    pf = rng.uniform(0.9, 1.0, size=(n, 14))
    qd = pd * np.tan(np.arccos(pf)) * np.sign(bus[:, QD])[None]
    soc = rng.uniform(0.1, 0.8, size=(n, 5))
    price = 0.49 * (1.0 + rng.standard_normal((n, 24))) * (1.0 + 0.1 * rng.standard_normal((n, 1)))
    st = np.concatenate([pd, qd, soc, price], axis=1).astype(np.float32)
    return torch.from_numpy(st).to(device)


_BOX_CACHE = {}


def evopf_partial_box(state: torch.Tensor):
    """(low, high) of the volatile partial-action box, float32, same device as ``state``
    (EVOPFEnv.update evopf.py:769-783 with Battery.update_bound :127-143).  The constant rows are cached per device,
    so the function performs no host->device copy after its first call (CUDA-graph capturable)."""
    key = str(state.device)
    if key not in _BOX_CACHE:
        _BOX_CACHE[key] = (torch.tensor(K.ACTION_LOW[K.PARTIAL_ACTIONS], dtype=torch.float32, device=state.device),
                           torch.tensor(K.ACTION_HIGH[K.PARTIAL_ACTIONS], dtype=torch.float32, device=state.device))
    lo0, hi0 = _BOX_CACHE[key]
    n = state.shape[0]
    soc = state[:, 28:33].double()
    hi_e = ((1.0 / K.ETA_IN) * torch.clamp(K.B_HIGH - soc, max=K.B_PMAX)).float()
    lo_e = (K.ETA_OUT * torch.clamp(K.B_LOW - soc, min=K.B_PMIN)).float()
    low = torch.cat([lo0[:-5].expand(n, 9), lo_e], dim=1)
    high = torch.cat([hi0[:-5].expand(n, 9), hi_e], dim=1)
    return low, high


def evopf_partial_actions(state: torch.Tensor, seed: int = 1, dist: str = "D1") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    u = torch.rand(state.shape[0], 14, generator=g, dtype=torch.float32).to(state.device)
    if dist == "D2":
        u = 0.35 + 0.3 * u
    low, high = evopf_partial_box(state)
    return low + (high - low) * u


def upstream_grad(n: int, dim: int, seed: int = 3, device="cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(n, dim, generator=g, dtype=torch.float32).to(device)


def cart_inputs(n: int, seed: int = 0, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    state = (torch.rand(n, 6, generator=g) - 0.5) * 0.1           # cartpole.py:233
    z = (torch.rand(n, 1, generator=g) - 0.5) * 20.0
    return state.float().to(device), z.float().to(device)


def pen_inputs(n: int, seed: int = 0, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)           # pendulum.py:131-140
    u = torch.rand(n, 5, generator=g, dtype=torch.float64)
    th = (u[:, 0] * 2 - 1) * (np.pi / 12)
    thd = u[:, 1] * 2 - 1
    l = 0.95 + 0.1 * u[:, 2]
    ld = (u[:, 3] * 2 - 1) * 0.05
    obs = torch.stack([torch.cos(th), torch.sin(th), thd, l, ld], dim=1).float()
    z = ((u[:, 4] * 2 - 1) * 6.0).float().view(-1, 1)
    return obs.to(device), z.to(device)
