"""CartSafe-v0 drop-in: the API of rpo.env.CartSafeEnv (rpo/env/classic_control/cartpole.py) over the CUDA kernels."""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import ops
from .spaces import Box, EnvBase


class _CompleteFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, partial_idx):
        ctx.partial_idx = partial_idx
        return ops.cart_complete_fwd(z, partial_idx)

    @staticmethod
    def backward(ctx, dl_da):
        return ops.cart_complete_bwd(dl_da.contiguous(), ctx.partial_idx), None


class _ResidFn(torch.autograd.Function):
    """eq_resid / ineq_resid / ineq_dist with autograd w.r.t. the action, as the reference's torch expressions have
    (cartpole.py:375-387; differentiated by the actor loss, rpo_ddpg.py:312-319).  kind: 0 eq | 1 ineq_resid | 2 ineq_dist."""

    @staticmethod
    def forward(ctx, action, kind):
        ctx.kind = kind
        ctx.save_for_backward(action)
        if kind == 0:
            return ops.cart_resid(action, want_ineq=False)[0]
        r = ops.cart_resid(action, want_eq=False)[1]
        return torch.clamp_(r, 0) if kind == 2 else r

    @staticmethod
    def backward(ctx, g):
        (action,) = ctx.saved_tensors
        g = g.contiguous()
        if ctx.kind == 0:
            return ops.cart_resid_bwd(action, g_eq=g), None
        return ops.cart_resid_bwd(action, g_ineq=g, ineq_mode=1 if ctx.kind == 2 else 0), None


class CartSafeEnv(EnvBase):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 50}

    def __init__(self, device=None, partial_idx=None):
        self.gravity, self.masscart, self.masspole = 9.8, 1.0, 0.1
        self.total_mass = self.masspole + self.masscart
        self.length = 0.5
        self.polemass_length = self.masspole * self.length
        self.tau = 0.02
        self.mu_c, self.mu_p = 0.1, 0.01
        self.delta = np.array([np.pi / 3, -np.pi / 6])
        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4
        fmax = np.finfo(np.float32).max
        high = np.array([self.x_threshold * 2, fmax, fmax, self.theta_threshold_radians * 2, fmax, fmax], dtype=np.float32)
        self.observation_space = Box(-high, high, dtype=np.float32)
        self.action_space = Box(np.array([-10, -10]).astype(np.float32), np.array([10, 10]).astype(np.float32))
        self.state_dim, self.action_dim, self.eq_num, self.ineq_num = 6, 2, 1, 6
        self.seed()
        self.state = None
        # cartpole.py:117: the policy-controlled force is drawn at random at construction (global numpy RNG)
        if partial_idx is None:
            self.partial_actions = np.random.choice(self.action_dim, self.action_dim - self.eq_num, replace=False)
        else:
            self.partial_actions = np.array([int(partial_idx)])
        self.other_actions = np.setdiff1d(np.arange(self.action_dim), self.partial_actions)
        self.steps_beyond_done = None
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.volatile = False
        self.update = None
        self.action = None
        self.last_state = None
        self.holding_eq = False                                   # cartpole.py:138-139
        self.holding_ineq = False

    @property
    def _p(self):
        return int(self.partial_actions[0])

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def _t(self, a):
        t = a.to(device=self.device, dtype=torch.float32) if isinstance(a, torch.Tensor) else \
            torch.as_tensor(np.asarray(a), dtype=torch.float32, device=self.device)
        return t.view(1, -1) if t.dim() == 1 else t

    # ---- hard-constraint protocol (cartpole.py:369-408) -------------------------------------------------
    def complete_partial(self, state, action_partial):
        z = action_partial.view(-1, 1) if action_partial.dim() == 1 else action_partial
        return _CompleteFn.apply(z, self._p)

    def eq_resid(self, state, action):
        return _ResidFn.apply(action, 0)

    def ineq_resid(self, state, action):
        return _ResidFn.apply(action, 1)

    def eq_dist(self, state, action):
        return torch.abs(self.eq_resid(state, action))

    def ineq_dist(self, state, action):
        return _ResidFn.apply(action, 2)

    def ineq_partial_grad(self, state, action, eps=0):
        return ops.cart_direction(action, self._p)

    def ineq_dist_np(self, state, action):
        return self.ineq_dist(None, self._t(action)).cpu().numpy()[0]

    def eq_resid_np(self, state, action):
        return self.eq_resid(None, self._t(action)).cpu().numpy()[0]

    def ineq_resid_np(self, state, action):                       # cartpole.py:410-411
        return self.ineq_resid(None, self._t(action)).cpu().numpy()[0]

    def eq_dist_np(self, state, action):                          # cartpole.py:420-422
        return np.abs(self.eq_resid_np(state, action))

    # Penalty-gradient forms used only by corr_mode != 0 (rpo_ddpg.py:275-278; no reference script): two tiny torch
    # products over the constant constraint matrices, kept for protocol completeness (cartpole.py:389-394).
    def _constraint_matrices(self, device):
        if getattr(self, "_mats", None) is None or self._mats[0].device != device:
            delta = torch.tensor([np.pi / 3, -np.pi / 6], dtype=torch.float32, device=device)
            E = torch.sin(delta).view(1, 2)
            A = torch.cat([torch.cos(delta).view(1, 2), -torch.cos(delta).view(1, 2),
                           torch.tensor([[1.0, 0.0], [-1.0, 0.0], [0.0, 1.0], [0.0, -1.0]], device=device)], dim=0)
            self._mats = (E, A)
        return self._mats

    def eq_grad(self, state, action):
        E, _ = self._constraint_matrices(action.device)
        return 2 * (action @ E.T) @ E

    def ineq_grad(self, state, action):
        _, A = self._constraint_matrices(action.device)
        return 2 * self.ineq_dist(state, action) @ A

    # The reference caches nothing for this environment either; the hold / release switches exist for API parity
    # (cartpole.py:424-434).
    def hold_eq(self):
        self.holding_eq = True

    def release_eq(self):
        self.holding_eq = False

    def hold_ineq(self):
        self.holding_ineq = True

    def release_ineq(self):
        self.holding_ineq = False

    def correct(self, state, action, lr, momentum, corr_eps, n_steps):
        return ops.cart_correct(action, self._p, lr, momentum, corr_eps, n_steps)

    def project(self, state, action_partial, lr, momentum, corr_eps, n_steps):
        z = action_partial.view(-1, 1) if action_partial.dim() == 1 else action_partial
        ac, a, flags = ops.cart_project(z, self._p, lr, momentum, corr_eps, n_steps, want_a=True)
        return ac, a, flags

    @property
    def box_constraint(self):
        return self.action_space.low, self.action_space.high

    @property
    def box_constraint_partial(self):
        return self.action_space.low[self.partial_actions], self.action_space.high[self.partial_actions]

    # ---- gym API (cartpole.py:166-239) ---------------------------------------------------------------------
    def reset(self):
        self.state = self.np_random.uniform(low=-0.05, high=0.05, size=(6,))
        self.action = None
        self.last_state = None
        self.steps_beyond_done = None
        return np.array(self.state)

    def step(self, action):
        action = np.asarray(action, dtype=np.float32)
        ns, done, eqv, iqv = ops.cart_step(self._t(self.state), self._t(action))
        self.last_state = np.array(self.state)
        self.action = action.copy()
        self.state = tuple(ns[0].double().cpu().numpy())
        done = bool(done[0].item())
        if not done:
            reward = 1.0
        elif self.steps_beyond_done is None:
            self.steps_beyond_done = 0
            reward = 1.0
        else:
            self.steps_beyond_done += 1
            reward = 0.0
        info = {"ineq_viol": iqv.cpu().numpy()[0], "eq_viol": eqv.cpu().numpy()[0]}
        return np.array(self.state), reward, done, info

    def step_batch(self, state, action):
        """(B,6), (B,2) device tensors -> next_state, done(int32), eq_viol, ineq_viol."""
        return ops.cart_step(state, action)

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the hot path")

    def close(self):
        pass
