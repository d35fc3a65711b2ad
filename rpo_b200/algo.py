"""Drivers of the projection: process_action / grad_steps with the reference's signatures
(rpo/algo/rpo_ddpg.py:72-77,266-305; rpo/algo/rpo_sac.py:76-81,280-319).

`grad_steps` runs the whole post-correction loop in ONE kernel launch.  The reference decides after every step,
on the host, whether the *global* max of |eq_resid| / ineq_dist over the batch still exceeds corr_eps (two
device->host syncs per step).  The kernel instead records, per step, whether any instance failed that test; the
host reads the K flags once and, in the rare case that the reference would have stopped early (k < K), re-runs the
kernel with n_steps = k from the original action — same trajectory, so the result equals the reference's loop.
"""
from __future__ import annotations

from typing import Tuple

import torch


def reference_step_count(flags: torch.Tensor, limit: int) -> int:
    """Steps the reference's while-loop takes: `step == 0 or unsat` and `step < limit` (rpo_ddpg.py:271-272)."""
    if limit <= 0:
        return 0
    fl = flags[:limit].tolist()          # one small device->host read
    for k in range(1, limit):
        if fl[k] == 0:
            return k
    return limit


def grad_steps_env(env, state, action, lr: float, momentum: float, corr_eps: float, limit: int,
                   exact: bool = True) -> Tuple[torch.Tensor, int]:
    """Post-correction loop for corr_mode == 0 on any rpo_b200 env.  Returns (action, steps)."""
    if limit <= 0:
        return action, 0
    out, flags = env.correct(state, action.detach(), lr, momentum, corr_eps, limit)
    if not exact:
        return out, limit
    k = reference_step_count(flags, limit)
    if k < limit:
        out, _ = env.correct(state, action.detach(), lr, momentum, corr_eps, k)
    return out, k


class _StraightThrough(torch.autograd.Function):
    """value = `corrected` (bit for bit), gradient = identity to `action` (see RPOProjectionMixin.grad_steps)."""

    @staticmethod
    def forward(ctx, action, corrected):
        return corrected.view_as(corrected)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _CompletePenaltyFn(torch.autograd.Function):
    """complete_partial + nju(ineq_dist) of the actor loss (rpo_ddpg.py:309-319) as one autograd node whose backward is
    ONE kernel (rpo_evopf_complete_bwd_penalty): the penalty's row selector is added to dL/dy in registers before the
    transposed solve, so neither the (B,58) gradient of ineq_dist nor the sum dL/dy + penalty term exists in HBM."""

    @staticmethod
    def forward(ctx, state, z, weight, tol, max_iters):
        from . import ops
        y = ops.evopf_complete_fwd(state, z, tol=tol, max_iters=max_iters)
        dist = torch.clamp_(ops.evopf_resid(state, y, want_eq=False, want_ineq=True)[1], 0)
        ctx.save_for_backward(state, y, dist, weight)
        return y, dist @ weight.reshape(-1, 1)

    @staticmethod
    def backward(ctx, dl_dy, dl_dpen):
        from . import ops
        state, y, dist, weight = ctx.saved_tensors
        if dl_dy is None:
            dl_dy = torch.zeros_like(y)
        if dl_dpen is None:
            dl_dpen = torch.zeros(y.shape[0], 1, dtype=y.dtype, device=y.device)
        dz, dx = ops.evopf_complete_bwd_penalty(state, y, dl_dy.contiguous(), weight.detach(), dl_dpen.contiguous(),
                                                need_dx=ctx.needs_input_grad[0])
        d_state = None
        if ctx.needs_input_grad[0]:
            d_state = torch.zeros_like(state)
            d_state[:, :28] = dx
        d_w = (dl_dpen.reshape(1, -1) @ dist).reshape(weight.shape) if ctx.needs_input_grad[2] else None
        return d_state, dz, d_w, None, None


def complete_partial_with_penalty(env, state, action_partial, nju_weight):
    """EVOPF: (action, nju(ineq_dist(state, action))) with the fused backward.  Same value and gradients as
    `a = env.complete_partial(state, z); pen = F.linear(env.ineq_dist(state, a), nju_weight)` (rpo_ddpg.py:309-319,
    rpo/algo/model/dual.py:63-65), two launches fewer in the backward."""
    if action_partial.dim() == 1:
        action_partial = action_partial.view(1, -1)
    return _CompletePenaltyFn.apply(state, action_partial, nju_weight, env.newton_tol, env.newton_max_iters)


class _SquashedCompletePenaltyFn(torch.autograd.Function):
    """box(u, state) -> complete_partial -> nju(ineq_dist) as one autograd node: forward = rpo_evopf_complete_fwd_squashed
    (+ the residual kernel when a penalty is wanted), backward = ONE launch of rpo_evopf_complete_bwd_squashed (GRG
    transposed solve with the penalty selector added in registers and the box derivative applied to the 14 reduced
    gradients).  Replaces four launches and two (B,14) / (B,43) round trips of the unfused graph."""

    @staticmethod
    def forward(ctx, state, u, lo, hi, mode, weight, tol, max_iters):
        from . import ops
        y = ops.evopf_complete_fwd_squashed(state, u, lo, hi, mode, tol=tol, max_iters=max_iters)
        ctx.mode = mode
        if weight is None:
            ctx.save_for_backward(state, u, lo, hi, y)
            ctx.has_w = False
            return y, y.new_zeros(y.shape[0], 1)
        dist = torch.clamp_(ops.evopf_resid(state, y, want_eq=False, want_ineq=True)[1], 0)
        ctx.save_for_backward(state, u, lo, hi, y, dist, weight)
        ctx.has_w = True
        return y, dist @ weight.reshape(-1, 1)

    @staticmethod
    def backward(ctx, dl_dy, dl_dpen):
        from . import ops
        if ctx.has_w:
            state, u, lo, hi, y, dist, weight = ctx.saved_tensors
        else:
            state, u, lo, hi, y = ctx.saved_tensors
            dist = weight = None
        if dl_dy is None:
            dl_dy = torch.zeros_like(y)
        rs = None
        if ctx.has_w:
            rs = dl_dpen.contiguous() if dl_dpen is not None else torch.zeros(y.shape[0], 1, dtype=y.dtype, device=y.device)
        du, dx = ops.evopf_complete_bwd_squashed(state, y, dl_dy.contiguous(), u, lo, hi, ctx.mode,
                                                 None if weight is None else weight.detach(), rs, need_dx=ctx.needs_input_grad[0])
        d_state = None
        if ctx.needs_input_grad[0]:
            d_state = torch.zeros_like(state)
            d_state[:, :28] = dx
        d_w = (rs.reshape(1, -1) @ dist).reshape(weight.shape) if (ctx.has_w and ctx.needs_input_grad[5]) else None
        return d_state, du, None, None, None, d_w, None, None


def complete_partial_squashed(env, box, state, u, nju_weight=None, tanh: bool = True):
    """EVOPF: (action, penalty) = (complete_partial(state, box(tanh(u), state)), nju(ineq_dist)) with the action box of a
    `rpo_b200.box.BoxConstraint` fused into the completion kernel and its derivative into the GRG backward kernel.
    tanh=True: u is the policy's raw output (policy.py:30-31); tanh=False: u is already in [-1, 1] (style 'tanh' affine)."""
    from . import ops
    if box.style != "tanh" or not box.volatile:
        raise ValueError("the fused form covers the volatile 'tanh'-style EVOPF partial box")
    mode = ops.BOX_TANH if tanh else ops.BOX_AFFINE
    if u.dim() == 1:
        u = u.view(1, -1)
    return _SquashedCompletePenaltyFn.apply(state, u, box.cmin_torch, box.cmax_torch, mode, nju_weight, env.newton_tol,
                                            env.newton_max_iters)


class RPOProjectionMixin(object):
    """Mix into (or patch onto) rpo.algo.RPODDPG / RPOSAC: same attribute names, same return conventions."""

    exact_loop_exit = True

    def grad_steps(self, state, action, train=True):
        if getattr(self, "corr_mode", 0) != 0:
            raise NotImplementedError("corr_mode != 0 is not on the B200 path (unused by every reference script)")
        if train:
            a, _ = grad_steps_env(self.env, state, action, self.corr_lr, self.corr_momentum, self.corr_eps,
                                  self.max_steps, self.exact_loop_exit)
            if torch.is_grad_enabled() and action.requires_grad:
                # The reference's train-mode loop is differentiable (action - lr * ineq_partial_grad(action),
                # rpo_ddpg.py:280-281); no reference script differentiates it (the call in actor_loss is commented out,
                # rpo_ddpg.py:315).  The kernel output carries no graph, so keep the caller's graph alive with a
                # straight-through estimator: value = corrected action, d/d action = identity, i.e. the correction's
                # own Jacobian lr * d(direction)/d(action) = O(corr_lr) is dropped.
                a = _StraightThrough.apply(action, a)
            return a
        with torch.no_grad():
            return grad_steps_env(self.env, state, action, self.eval_lr, self.corr_momentum, self.corr_eps,
                                  self.eval_steps, self.exact_loop_exit)

    def process_action(self, state, action_partial, train=True):
        needs_grad = torch.is_grad_enabled() and (action_partial.requires_grad or state.requires_grad)
        if needs_grad or getattr(self, "corr_mode", 0) != 0:
            try:                           # EVOPF: the fused kernel's own Newton sequence, so both paths give the same action
                action = self.env.complete_partial(state, action_partial, refine=False)
            except TypeError:
                action = self.env.complete_partial(state, action_partial)
            return self.grad_steps(state, action, train=train)
        # no-grad call sites (rpo_ddpg.py:100,106,233,330): fused completion + correction, one launch
        lr, limit = (self.corr_lr, self.max_steps) if train else (self.eval_lr, self.eval_steps)
        yc, y, flags = self.env.project(state, action_partial, lr, self.corr_momentum, self.corr_eps, limit)
        k = limit
        if self.exact_loop_exit:
            k = reference_step_count(flags, limit)
            if k < limit:
                yc, _ = self.env.correct(state, y, lr, self.corr_momentum, self.corr_eps, k)
        return yc if train else (yc, k)


def process_action_squashed(projector, box, state, u, train=True, tanh=True):
    """process_action(state, box(tanh(u), state)) for no-grad call sites in ONE launch (rpo_evopf_project_squashed): the
    rollout / critic-target projections of rpo_ddpg.py:96-106,327-332 with the policy's squash fused in."""
    from . import ops
    env = projector.env
    lr, limit = (projector.corr_lr, projector.max_steps) if train else (projector.eval_lr, projector.eval_steps)
    mode = ops.BOX_TANH if tanh else ops.BOX_AFFINE
    yc, y, flags, _ = ops.evopf_project_squashed(state, u, box.cmin_torch, box.cmax_torch, mode, lr, projector.corr_momentum,
                                                 projector.corr_eps, limit, tol=env.newton_tol, max_iters=env.newton_max_iters,
                                                 want_y=projector.exact_loop_exit)
    k = limit
    if projector.exact_loop_exit:
        k = reference_step_count(flags, limit)
        if k < limit:
            yc, _ = env.correct(state, y, lr, projector.corr_momentum, projector.corr_eps, k)
    return yc if train else (yc, k)


def install(rpo_module=None, register_envs: bool = True, box: bool = True):
    """Route the reference's own RPODDPG / RPOSAC through the B200 path: patches
    process_action / grad_steps on the two classes, re-registers the gym ids to rpo_b200.env classes and (box=True)
    makes both agents construct rpo_b200.box.BoxConstraint (one launch per squash / clip) instead of
    rpo/algo/model/utils.py's host round trip (rpo_ddpg.py:28,63; rpo_sac.py:29,67)."""
    if rpo_module is None:
        import rpo as rpo_module           # the reference package
    import rpo.algo as ralgo
    for cls in (ralgo.RPODDPG, ralgo.RPOSAC):
        cls.process_action = RPOProjectionMixin.process_action
        cls.grad_steps = RPOProjectionMixin.grad_steps
        cls.exact_loop_exit = True
    if box:
        import importlib
        from .box import BoxConstraint
        for name in ("rpo.algo.rpo_ddpg", "rpo.algo.rpo_sac"):
            importlib.import_module(name).BoxConstraint = BoxConstraint
    if register_envs:
        from .env import register_all
        register_all(force=True)
    return rpo_module
