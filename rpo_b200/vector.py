"""Device-resident rollout plumbing for the "next" rows of SURVEY.md §8(f): vectorised EVOPF replicas stepped by
the batched step kernel (f-2), a device-side replay ring buffer (f-2) and a device-side exogenous-row sampler (f-4).

The reference rolls out ONE environment on the host (rpo_ddpg.py:88-118), pushes every transition through numpy
(buffer.py:22-29) and re-uploads seven arrays per training step (rpo_ddpg.py:165-174).  Here N replicas live on the
GPU: states, actions, rewards and the replay memory never leave HBM.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import evopf_constants as K
from . import ops
from .case14 import PD, QD, case14


def _daily_curves(demand_curve=None, price_curve=None, T: int = 24):
    """The 24-step daily curves of the reference loaders (data/demand.py:14-27, data/price.py:13-27): averaged from the
    reference's recorded series when its package is importable (rpo_b200.env.loaders.find_reference_data_dir), else the
    synthetic profiles of rpo_b200.env.loaders."""
    import numpy as np
    from .env import loaders
    d = loaders.find_reference_data_dir() if (demand_curve is None or price_curve is None) else None
    if demand_curve is None:
        demand_curve = loaders._daily_curve(d + "/demand.pickle", "value", T) if d else loaders.synthetic_curve(T, level=1000.0)
    if price_curve is None:
        price_curve = loaders._daily_curve(d + "/price.pickle", "SYS", T) if d else loaders.synthetic_curve(T, level=49.0, swing=0.25)
    return np.asarray(demand_curve, dtype=np.float64), np.asarray(price_curve, dtype=np.float64)


class EvopfExoSampler(object):
    """Exogenous rows [pd(14) | qd(14) | price window(24)] drawn on the device with the reference loaders' semantics.

    Demand (data/demand.py:47-67): hour h's total is sum(Pd) * curve[h] / mean(curve); bus shares = case share mixed with a
    Dirichlet(1) draw over the 11 load buses (rho = .5), qd = pd tan(acos(U(.9,1))) sign(Qd) — independent per hour, so
    drawing them step by step is the same distribution as the reference's per-episode table.
    Price (data/price.py:29-55): ONE realisation per episode, price[0:48] = mag_r * [curve | 0] * (1 + N(0,1)^48) / 100 with
    mag_r = 1 + 0.1 N(0,1) drawn once; the observation at hour h is the window price[h:h+24] of that single realisation,
    so the look-ahead IS the future price (state[t].price[1] == state[t+1].price[0]).  `new_episode(mask)` redraws the
    realisation of the replicas whose day restarts."""

    def __init__(self, device, seed: int = 0, T: int = 24, default_generator: bool = False, demand_curve=None,
                 price_curve=None, n: int = 0):
        """default_generator=True draws from torch's default generator of the device (which CUDA-graph capture
        supports) instead of a private seeded one."""
        self.device = torch.device(device)
        self.T = T
        self.gen = None if default_generator else torch.Generator(device=self.device).manual_seed(seed)
        self.ar24 = torch.arange(24, device=self.device)
        bus = case14()["bus"]
        pd_case = torch.tensor(bus[:, PD], dtype=torch.float64)
        self.share = (pd_case / pd_case.sum()).to(self.device, torch.float32)
        self.load_mask = (pd_case != 0).to(self.device)
        self.q_sign = torch.tensor(bus[:, QD], dtype=torch.float64).sign().to(self.device, torch.float32)
        self.total = float(pd_case.sum() / 100.0)                      # 2.59 pu
        dc, pc = _daily_curves(demand_curve, price_curve, T)
        self.demand_curve = torch.tensor(dc / dc.mean(), dtype=torch.float32, device=self.device)   # demand.py:51-52
        self.price_curve = torch.tensor(pc / 100.0, dtype=torch.float32, device=self.device)        # price / baseMVA, evopf.py:57
        self.price_ep = torch.zeros(n, 2 * T, device=self.device)      # the current episode's price realisation per replica
        if n:
            self.new_episode(None)

    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.gen, device=self.device)

    def _randn(self, *shape):
        return torch.randn(*shape, generator=self.gen, device=self.device)

    def new_episode(self, mask):
        """Redraw the episode price realisation (price.py:40-47) of the replicas in `mask` ((N,) bool; None = all).
        Always draws for all N rows and selects: no data-dependent shapes, so the call can sit inside a CUDA graph."""
        n = self.price_ep.shape[0]
        mag = 1.0 + 0.1 * self._randn(n, 1)
        fresh = torch.zeros_like(self.price_ep)
        fresh[:, :self.T] = mag * self.price_curve[None] * (1.0 + self._randn(n, self.T))
        if mask is None:
            self.price_ep.copy_(fresh)
        else:
            self.price_ep.copy_(torch.where(mask[:, None], fresh, self.price_ep))

    def rows(self, hour: torch.Tensor) -> torch.Tensor:
        """hour (N,) int64 in [0, T) -> (N, 52) float32."""
        n = hour.shape[0]
        e = -torch.log(self._rand(n, 14).clamp_min(1e-12)) * self.load_mask        # Dirichlet(1) via normalised Exp(1)
        dirich = e / e.sum(dim=1, keepdim=True)
        ratio = 0.5 * self.share[None] + 0.5 * dirich
        pd = ratio * (self.total * self.demand_curve[hour])[:, None]
        pf = 0.9 + 0.1 * self._rand(n, 14)
        qd = pd * torch.tan(torch.acos(pf)) * self.q_sign[None]
        if self.price_ep.shape[0] != n:                                # stand-alone use: one fresh realisation per call
            self.price_ep = torch.zeros(n, 2 * self.T, device=self.device)
            self.new_episode(None)
        price = torch.gather(self.price_ep, 1, hour[:, None] + self.ar24[None])     # window [h, h+24) of the realisation
        return torch.cat([pd, qd, price], dim=1).float()


class TableExo(object):
    """Exogenous rows from precomputed per-replica day tables (N, T, 52): e.g. the reference loaders' own draws, so that a
    batched rollout can be compared with the reference's sequential one on identical data (tests)."""

    def __init__(self, table: torch.Tensor):
        self.table = table
        self.device = table.device
        self.T = table.shape[1]
        self._ar = torch.arange(table.shape[0], device=table.device)

    def new_episode(self, mask):
        return None

    def rows(self, hour: torch.Tensor) -> torch.Tensor:
        return self.table[self._ar, hour]


class EvopfVectorEnv(object):
    """N independent EVOPF-v0 replicas on one GPU; `step` = one launch of rpo_evopf_step (evopf.py:348-366 row-wise).
    Episodes last T = 24 steps (demand.py:69-75) and restart automatically (soc back to 0.2, evopf.py:62-63,243; a new
    price realisation, price.py:29-38)."""

    def __init__(self, n: int, device, seed: int = 0, default_generator: bool = False, exo=None):
        self.n, self.device = n, torch.device(device)
        self.exo = exo if exo is not None else EvopfExoSampler(self.device, seed, default_generator=default_generator, n=n)
        self.T = self.exo.T
        self.hour = torch.zeros(n, dtype=torch.int64, device=self.device)
        self.state = None

    def reset(self) -> torch.Tensor:
        self.hour.zero_()
        self.exo.new_episode(None)
        rows = self.exo.rows(self.hour)
        soc = torch.full((self.n, 5), K.B_LOW + 0.1, device=self.device)
        self.state = torch.cat([rows[:, :28], soc, rows[:, 28:]], dim=1).contiguous()
        return self.state

    def step(self, action: torch.Tensor, inplace: bool = False):
        """action (N,43) -> next_state (N,57), reward (N,), done (N,) bool, eq_viol (N,28), ineq_viol (N,58).
        inplace=True keeps `self.state` in ONE buffer (needed under CUDA-graph capture): the pre-step state is saved in
        `self.prev_state` and the buffer is overwritten with the next state."""
        nxt_hour = self.hour + 1
        done = nxt_hour >= self.T
        self.exo.new_episode(done)                                     # replicas whose day is over start a new realisation
        rows = self.exo.rows(torch.where(done, torch.zeros_like(nxt_hour), nxt_hour))
        ns, rew, eqv, iqv = ops.evopf_step(self.state, action, rows)
        # episode restart without a host round trip: soc back to its initial value where the day is over
        soc = torch.where(done[:, None], torch.full_like(ns[:, 28:33], K.B_LOW + 0.1), ns[:, 28:33])
        ns = torch.cat([ns[:, :28], soc, ns[:, 33:]], dim=1)
        self.hour.copy_(torch.where(done, torch.zeros_like(nxt_hour), nxt_hour))
        if inplace:
            self.prev_state = self.state.clone()
            self.state.copy_(ns)
        else:
            self.prev_state = self.state
            self.state = ns
        return ns, rew, done, eqv, iqv


class DeviceReplay(object):
    """Ring buffer of transitions in HBM with uniform sampling on the device (reference: utils/buffer.py:3-34).
    Write pointer and fill level live on the device too, so extend() / sample() issue no host synchronisation and can
    be captured in a CUDA graph (`default_generator=True` then draws from torch's graph-safe default generator)."""

    def __init__(self, capacity: int, device, default_generator: bool = False, **shapes: int):
        self.capacity, self.device = int(capacity), torch.device(device)
        self.buf: Dict[str, torch.Tensor] = {k: torch.zeros(self.capacity, d, device=self.device) for k, d in shapes.items()}
        self.ptr = torch.zeros((), dtype=torch.int64, device=self.device)
        self.fill = torch.zeros((), dtype=torch.int64, device=self.device)
        self.gen = None if default_generator else torch.Generator(device=self.device).manual_seed(1234)

    def __len__(self):
        return int(self.fill.item())

    def extend(self, **rows: torch.Tensor):
        n = next(iter(rows.values())).shape[0]
        idx = (self.ptr + torch.arange(n, device=self.device)) % self.capacity
        for k, v in rows.items():
            self.buf[k].index_copy_(0, idx, v.reshape(n, -1).to(self.buf[k].dtype))
        self.ptr.copy_((self.ptr + n) % self.capacity)
        self.fill.copy_(torch.clamp(self.fill + n, max=self.capacity))

    def sample(self, num: int) -> Dict[str, torch.Tensor]:
        u = torch.rand(num, generator=self.gen, device=self.device)
        idx = torch.clamp((u * self.fill).long(), max=self.capacity - 1)
        return {k: v.index_select(0, idx) for k, v in self.buf.items()}


@torch.no_grad()
def evaluate_batched(venv: EvopfVectorEnv, policy, projector, max_len: int = 500):
    """Batched counterpart of RPODDPG.eval (rpo_ddpg.py:207-264, SURVEY f-4): the reference plays 10 episodes one
    after the other with B = 1 projections (eval_steps = 50 correction steps each); here all replicas of `venv` play
    one episode in lock-step, every step being ONE fused projection launch with train=False semantics.

    policy(state) -> partial action (N,14); projector.process_action(state, z, train=False) -> (action, steps).
    Returns the reference's ten statistics: mean/std of episode reward, of the per-episode running mean of the max
    inequality / equality violation, and of the per-episode maxima."""
    state = venv.reset()
    n = venv.n
    dev = state.device
    total = torch.zeros(n, device=dev)
    mean_ineq = torch.zeros(n, device=dev)
    mean_eq = torch.zeros(n, device=dev)
    max_ineq = torch.zeros(n, device=dev)
    max_eq = torch.zeros(n, device=dev)
    alive = torch.ones(n, dtype=torch.bool, device=dev)
    for i in range(max_len):
        action, _ = projector.process_action(state, policy(state), train=False)
        state, rew, done, eqv, iqv = venv.step(action)
        iq, eq = iqv.max(dim=1).values, eqv.abs().max(dim=1).values
        w = alive.float()
        total += w * rew
        mean_ineq += w * (iq - mean_ineq) / (i + 1)
        mean_eq += w * (eq - mean_eq) / (i + 1)
        max_ineq = torch.where(alive, torch.maximum(max_ineq, iq), max_ineq)
        max_eq = torch.where(alive, torch.maximum(max_eq, eq), max_eq)
        alive = alive & ~done
        if not bool(alive.any()):
            break
    out = []
    for v in (total, mean_ineq, mean_eq, max_ineq, max_eq):
        out += [float(v.mean()), float(v.std(unbiased=False))]
    return tuple(out)
