"""SpringPendulum-v0 drop-in: the API of rpo.env.SpringPendulumEnv (rpo/env/classic_control/pendulum.py) over the kernels.

Row-wise semantics: the reference's batched ineq_partial_grad multiplies (B,1) by (1,B) (pendulum.py:337-339) and so
couples rows for B > 1; every kernel here evaluates a row as the reference evaluates a batch of one.
The reference caches diff_eq* on the env between calls (pendulum.py:264-288); the kernels recompute them from the
observation (5 flops), so this class is stateless between protocol calls.

`reference_batch_coupling=True` (opt-in) reproduces the reference's batched behaviour literally instead: the correction
direction of row i then sums the reduced gradients d_j of ALL rows j with a_p[i] d_j > b_i (the (B,B) product), and
`correct` / `project` run the reference's loop step by step around it.  That path is a handful of torch operations
(O(B^2) like the reference's), not a kernel: it exists so that a batched call can be compared with the reference
result for result; the row-wise kernels stay the default and the fast path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .spaces import Box, EnvBase


class _CompleteFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, z):
        ctx.save_for_backward(obs)
        return ops.pen_complete_fwd(obs, z)

    @staticmethod
    def backward(ctx, dl_da):
        (obs,) = ctx.saved_tensors
        return None, ops.pen_complete_bwd(obs, dl_da.contiguous())


class _ResidFn(torch.autograd.Function):
    """eq_resid / ineq_resid / ineq_dist with autograd w.r.t. the action (pendulum.py:296-311; the equality coefficients
    are detached copies of the observation there, :273-288, so no gradient reaches the state).  kind: 0 | 1 | 2 as in
    the other environments."""

    @staticmethod
    def forward(ctx, obs, action, kind):
        ctx.kind = kind
        ctx.save_for_backward(obs, action)
        if kind == 0:
            return ops.pen_resid(obs, action, want_ineq=False)[0]
        r = ops.pen_resid(obs, action, want_eq=False)[1]
        return torch.clamp_(r, 0) if kind == 2 else r

    @staticmethod
    def backward(ctx, g):
        obs, action = ctx.saved_tensors
        g = g.contiguous()
        if ctx.kind == 0:
            return None, ops.pen_resid_bwd(obs, action, g_eq=g), None
        return None, ops.pen_resid_bwd(obs, action, g_ineq=g, ineq_mode=1 if ctx.kind == 2 else 0), None


class SpringPendulumEnv(EnvBase):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}

    def __init__(self, g=10.0, device=None, reference_batch_coupling=False):
        self.reference_batch_coupling = bool(reference_batch_coupling)
        self.max_speed, self.max_torque, self.max_length_delta = 8.0, 6.0, 0.5
        self.max_summation = 32.0
        self.dt, self.g, self.m, self.k, self.l0 = 0.05, g, 0.5, 1.0, 1.0
        if g != 10.0:
            raise NotImplementedError("the kernels bake g = 10 (the value every reference script uses)")
        self.m_dt = self.m / self.dt
        high = np.array([1.0, 1.0, self.max_speed, 1.5, 0.05], dtype=np.float32)
        low = np.array([-1.0, -1.0, -self.max_speed, 0.5, -0.05], dtype=np.float32)
        ah = self.max_torque * np.array([1.0, 1.0], dtype=np.float32)
        self.action_space = Box(low=-ah, high=ah, shape=(2,), dtype=np.float32)
        self.observation_space = Box(low=low, high=high, dtype=np.float32)
        self.state_dim, self.action_dim, self.eq_num, self.ineq_num = 5, 2, 1, 1
        self.fx_idx, self.fy_idx = 0, 1
        self.seed()
        self.partial_actions = np.array([0])
        self.other_actions = np.setdiff1d(np.arange(self.action_dim), self.partial_actions)
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.counter = None
        self.volatile = False
        self.update = None
        self.state = None
        self.holding_eq = False                                   # pendulum.py:61-62
        self.holding_ineq = False

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def _t(self, a):
        t = a.to(device=self.device, dtype=torch.float32) if isinstance(a, torch.Tensor) else \
            torch.as_tensor(np.asarray(a), dtype=torch.float32, device=self.device)
        return t.view(1, -1) if t.dim() == 1 else t

    # ---- hard-constraint protocol (pendulum.py:256-343) ---------------------------------------------------
    def complete_partial(self, state, action_partial):
        z = action_partial.view(-1, 1) if action_partial.dim() == 1 else action_partial
        return _CompleteFn.apply(state, z)

    def eq_resid(self, state, action):
        return _ResidFn.apply(state, action, 0)

    def ineq_resid(self, state, action):
        return _ResidFn.apply(state, action, 1)

    def eq_dist(self, state, action):
        return torch.abs(self.eq_resid(state, action))

    def ineq_dist(self, state, action):
        return _ResidFn.apply(state, action, 2)

    def ineq_partial_grad(self, state, action, eps=0):
        if self.reference_batch_coupling and action.shape[0] > 1:
            return self._direction_coupled(state, action)
        return ops.pen_direction(state, action)

    def _direction_coupled(self, state, action, rows_per_block: int = 4096):
        """pendulum.py:331-343 for a batch, operation for operation (float32): row i is coupled to every row j through
        `action[:, p] @ diff_grad_partial.T` (B,1)@(1,B).  Evaluated in row blocks to bound the (B,B) temporary."""
        cos, sin, thd, l, ld = (state[:, i:i + 1] for i in range(5))
        e_p, inv_o = sin, 1 / cos                                              # diff_eq_partial, diff_eq_other_inv (:273-281)
        bias = -self.m_dt * ld - (l * self.m * thd ** 2 - self.k * (l - self.l0) - self.m * self.g * cos)      # :287
        D = 2 * action.detach()                                                # set_ineq :294
        d = D[:, 0:1] - D[:, 1:2] * (inv_o * e_p)                              # diff_grad_partial :334-335
        b = self.max_summation - (bias * inv_o) * D[:, 1:2]                    # bias_grad_partial :336
        grad = torch.empty_like(d)
        for r0 in range(0, action.shape[0], rows_per_block):
            m = torch.clamp(action[r0:r0 + rows_per_block, 0:1] @ d.T - b[r0:r0 + rows_per_block], 0)     # :337
            grad[r0:r0 + rows_per_block] = (m > 0).to(torch.float32) @ d                                  # :339
        return torch.cat([grad, -(grad * e_p) * inv_o], dim=1)                 # :340-342

    def _correct_coupled(self, state, action, lr, momentum, corr_eps, n_steps):
        """RPODDPG.grad_steps (rpo_ddpg.py:266-286) step by step, always n_steps steps like the kernels; flags[k] = the
        reference's loop test before step k."""
        a = action.detach().clone()
        prev = torch.zeros_like(a)
        flags = torch.zeros(max(n_steps, 1), dtype=torch.int32, device=a.device)
        for k in range(n_steps):
            unsat = (self.eq_resid(state, a).abs().max() > corr_eps) | (self.ineq_dist(state, a).max() > corr_eps)
            flags[k] = unsat.to(torch.int32)
            step = lr * self._direction_coupled(state, a) + momentum * prev
            a = a - step
            prev = step
        return a, flags

    def ineq_dist_np(self, state, action):
        return self.ineq_dist(self._t(state), self._t(action)).cpu().numpy()

    def eq_resid_np(self, state, action):
        return self.eq_resid(self._t(state), self._t(action)).cpu().numpy()

    # The reference caches diff_eq* / diff_ineq on the env and lets callers freeze them (pendulum.py:264-296, 345-355);
    # the kernels recompute them from the observation, so these are switches without effect, kept for API parity.
    def set_eq(self, state, action_partial=None):
        return None

    def set_ineq(self, state, action):
        return None

    def hold_eq(self):
        self.holding_eq = True

    def release_eq(self):
        self.holding_eq = False

    def hold_ineq(self):
        self.holding_ineq = True

    def release_ineq(self):
        self.holding_ineq = False

    def correct(self, state, action, lr, momentum, corr_eps, n_steps):
        if self.reference_batch_coupling and action.shape[0] > 1:
            return self._correct_coupled(state, action, lr, momentum, corr_eps, n_steps)
        return ops.pen_correct(state, action, lr, momentum, corr_eps, n_steps)

    def project(self, state, action_partial, lr, momentum, corr_eps, n_steps):
        z = action_partial.view(-1, 1) if action_partial.dim() == 1 else action_partial
        if self.reference_batch_coupling and z.shape[0] > 1:
            a = ops.pen_complete_fwd(state, z)
            ac, flags = self._correct_coupled(state, a, lr, momentum, corr_eps, n_steps)
            return ac, a, flags
        ac, a, flags = ops.pen_project(state, z, lr, momentum, corr_eps, n_steps, want_a=True)
        return ac, a, flags

    @property
    def box_constraint(self):
        return self.action_space.low, self.action_space.high

    @property
    def box_constraint_partial(self):
        return self.action_space.low[self.partial_actions], self.action_space.high[self.partial_actions]

    # ---- gym API (pendulum.py:80-140) ----------------------------------------------------------------------
    def reset(self):
        high = np.array([np.pi / 12, 1.0, 1.05, 0.05], dtype=np.float32)
        low = np.array([-np.pi / 12, -1.0, 0.95, -0.05], dtype=np.float32)
        self.state = self.np_random.uniform(low=low, high=high)
        self.counter = 0
        return self._get_obs()

    def _get_obs(self):
        theta, thetadot, l, ldot = self.state
        return np.array([np.cos(theta), np.sin(theta), thetadot, l, ldot])

    def step(self, action):
        self.counter += 1
        action = np.asarray(action, dtype=np.float32)
        nst, nobs, rew, done, eqv, iqv = ops.pen_step(self._t(self.state), self._t(action))
        self.state = nst[0].double().cpu().numpy()
        info = {"ineq_viol": iqv.cpu().numpy(), "eq_viol": eqv.cpu().numpy()}
        return nobs[0].double().cpu().numpy(), float(rew[0].item()), bool(done[0].item()), info

    def step_batch(self, st, action):
        """(B,4) internal states, (B,2) actions -> next_st, next_obs, reward, done(int32), eq_viol, ineq_viol."""
        return ops.pen_step(st, action)

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the hot path")

    def close(self):
        pass
