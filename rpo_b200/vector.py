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


class EvopfExoSampler(object):
    """Draws exogenous rows [pd(14) | qd(14) | price window(24)] on the device with the reference loaders' distributions
    (data/demand.py:47-67, data/price.py:40-47): case share mixed with Dirichlet(1) over the 11 load buses (rho = .5),
    total load following a daily curve, qd = pd tan(acos(U(.9,1))) sign(Qd), price = curve (1+N)(1+.1 N)/100."""

    def __init__(self, device, seed: int = 0, T: int = 24, default_generator: bool = False):
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
        h = torch.arange(T, dtype=torch.float32)
        self.demand_curve = (1.0 + 0.16 * torch.sin((h - 9.0) * 2 * math.pi / T)).to(self.device)
        self.price_curve = (0.49 * (1.0 + 0.25 * torch.sin((h - 15.0) * 2 * math.pi / T))).to(self.device)

    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.gen, device=self.device)

    def _randn(self, *shape):
        return torch.randn(*shape, generator=self.gen, device=self.device)

    def rows(self, hour: torch.Tensor) -> torch.Tensor:
        """hour (N,) int64 in [0, T) -> (N, 52) float32."""
        n = hour.shape[0]
        e = -torch.log(self._rand(n, 14).clamp_min(1e-12)) * self.load_mask        # Dirichlet(1) via normalised Exp(1)
        dirich = e / e.sum(dim=1, keepdim=True)
        ratio = 0.5 * self.share[None] + 0.5 * dirich
        pd = ratio * (self.total * self.demand_curve[hour])[:, None]
        pf = 0.9 + 0.1 * self._rand(n, 14)
        qd = pd * torch.tan(torch.acos(pf)) * self.q_sign[None]
        idx = (hour[:, None] + self.ar24[None]) % self.T
        ahead_valid = (hour[:, None] + self.ar24[None]) < self.T
        price = self.price_curve[idx] * (1 + self._randn(n, 24)) * (1 + 0.1 * self._randn(n, 1)) * ahead_valid
        return torch.cat([pd, qd, price], dim=1).float()


class EvopfVectorEnv(object):
    """N independent EVOPF-v0 replicas on one GPU; `step` = one launch of rpo_evopf_step (evopf.py:348-366 row-wise).
    Episodes last T = 24 steps (demand.py:69-75) and restart automatically (soc back to 0.2, evopf.py:62-63,243)."""

    def __init__(self, n: int, device, seed: int = 0, default_generator: bool = False):
        self.n, self.device = n, torch.device(device)
        self.exo = EvopfExoSampler(self.device, seed, default_generator=default_generator)
        self.T = self.exo.T
        self.hour = torch.zeros(n, dtype=torch.int64, device=self.device)
        self.state = None

    def reset(self) -> torch.Tensor:
        self.hour.zero_()
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
