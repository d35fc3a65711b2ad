"""Drop-in `BoxConstraint` — same constructor, methods and attribute names as the reference's
rpo/algo/model/utils.py:5-118, evaluated by ONE kernel launch (rpo_box_fwd / rpo_box_bwd) instead of the reference's
per-call host round trip (state -> numpy -> Battery.update_bound in float64 -> four fresh H2D tensors, utils.py:75-90).

Usage inside a policy is unchanged: ``box(torch.tanh(x), state)`` (policy.py:30-31); ``box.squash(x, state)`` fuses
the tanh into the same launch.  All tensor paths are differentiable w.r.t. x (the state receives no gradient, as in
the reference).  numpy inputs are accepted as the reference accepts them and return numpy arrays.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class _BoxFn(torch.autograd.Function):
    """x (B,W) -> (B,W); backward = one launch of rpo_box_bwd."""

    @staticmethod
    def forward(ctx, x, lo, hi, state, n_vol, mode):
        ctx.save_for_backward(x, lo, hi, state if state is not None else x.new_empty(0))
        ctx.n_vol, ctx.mode, ctx.has_state = n_vol, mode, state is not None
        return ops.box_fwd(x, lo, hi, mode, state, n_vol)

    @staticmethod
    def backward(ctx, g):
        x, lo, hi, state = ctx.saved_tensors
        gx = ops.box_bwd(x, g.contiguous(), lo, hi, ctx.mode, state if ctx.has_state else None, ctx.n_vol)
        return gx, None, None, None, None, None


class BoxConstraint(object):
    """cmin / cmax: numpy (W,) box (env.box_constraint or env.box_constraint_partial); volatile=True makes the last five
    columns follow the EVOPF batteries (`update` is accepted for signature compatibility and used only to answer
    `update_box` on numpy states the way the reference does)."""

    N_VOL = 5

    def __init__(self, cmin, cmax, device, style="tanh", volatile=False, update=None, full=False):
        if style not in ("tanh", "sigmoid"):
            raise ValueError("style must be 'tanh' or 'sigmoid'")
        self.style, self.volatile, self.update, self.full = style, volatile, update, full
        self.cmin, self.cmax = np.asarray(cmin), np.asarray(cmax)
        if style == "sigmoid":
            self.base, self.scale = self.cmin, self.cmax - self.cmin
        else:
            self.scale = (self.cmax - self.cmin) / 2
            self.base = self.cmin + self.scale
        self.device = torch.device(device)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.base_torch, self.scale_torch = torch.tensor(self.base, **f32), torch.tensor(self.scale, **f32)
        self.cmin_torch, self.cmax_torch = torch.tensor(self.cmin, **f32), torch.tensor(self.cmax, **f32)
        self._affine = ops.BOX_AFFINE01 if style == "sigmoid" else ops.BOX_AFFINE

    # reference API: utils.py:33-38
    def cuda(self):
        self.to("cuda")

    def to(self, *args, **kwargs):
        for k in ("base_torch", "scale_torch", "cmin_torch", "cmax_torch"):
            setattr(self, k, getattr(self, k).to(*args, **kwargs))
        self.device = self.cmin_torch.device

    # ---- the one launch ----------------------------------------------------------------------------------
    def _n_vol(self):
        return self.N_VOL if self.volatile else 0

    def _state(self, state):
        if not self.volatile:
            return None
        if state is None:
            raise ValueError("a volatile BoxConstraint needs the state")
        if not isinstance(state, torch.Tensor):
            state = torch.as_tensor(np.asarray(state), dtype=torch.float32)
        state = state.to(self.device, torch.float32)
        return state.view(1, -1) if state.dim() == 1 else state

    def _apply(self, x, state, mode):
        is_np = not isinstance(x, torch.Tensor)
        xt = torch.as_tensor(np.asarray(x), dtype=torch.float32).to(self.device) if is_np else x
        x2 = xt.view(1, -1) if xt.dim() == 1 else xt
        st = self._state(state)
        if st is not None and st.shape[0] != x2.shape[0]:             # the reference broadcasts (n,W) against x
            if st.shape[0] == 1:
                st = st.expand(x2.shape[0], -1)
            elif x2.shape[0] == 1:
                x2 = x2.expand(st.shape[0], -1)
            else:
                raise ValueError("state and x disagree on the batch size")
        out = _BoxFn.apply(x2, self.cmin_torch, self.cmax_torch, st, self._n_vol(), mode)
        if st is None and xt.dim() == 1:
            out = out.view(-1)
        return out.detach().cpu().numpy() if is_np else out

    def __call__(self, x, state=None):                               # utils.py:40-51
        return self._apply(x, state, self._affine)

    def squash(self, x, state=None):
        """box(tanh(x), state) in one launch (policy.py:30-31); style 'sigmoid' has no fused form."""
        if self.style != "tanh":
            return self(torch.sigmoid(x), state)
        return self._apply(x, state, ops.BOX_TANH)

    def clip(self, x, state=None):                                   # utils.py:92-104
        return self._apply(x, state, ops.BOX_CLIP)

    def _query(self, state, mode):
        st = self._state(state)
        B = 1 if st is None else st.shape[0]
        return ops.box_fwd(None, self.cmin_torch, self.cmax_torch, mode, st, self._n_vol(), B=B)

    def action_scale(self, state=None):                              # utils.py:106-118
        if not self.volatile:
            return self.scale_torch if isinstance(state, torch.Tensor) else self.scale
        s = self._query(state, ops.BOX_SCALE)                        # (hi - lo) / 2
        if self.style == "sigmoid":
            s = s * 2
        return s if isinstance(state, torch.Tensor) else s.cpu().numpy()

    def sample(self, state):                                         # utils.py:53-62
        tmp = torch.rand_like(self.base_torch)
        if self.style == "tanh":
            tmp = 2 * tmp - 1
        return self(tmp, state) if self.volatile else self.scale_torch * tmp + self.base_torch

    def sample_np(self, state):                                      # utils.py:64-73
        tmp = np.random.rand(*self.base.shape)
        if self.style == "tanh":
            tmp = 2 * tmp - 1
        return self(tmp, state) if self.volatile else self.scale * tmp + self.base

    def update_box(self, x):                                         # utils.py:75-90
        """Refreshes the reference's cached *_vol attributes from the device (kept for callers that read them)."""
        self.cmin_torch_vol = self._query(x, ops.BOX_LOW)
        self.cmax_torch_vol = self._query(x, ops.BOX_HIGH)
        if self.style == "sigmoid":
            self.base_torch_vol, self.scale_torch_vol = self.cmin_torch_vol, self.cmax_torch_vol - self.cmin_torch_vol
        else:
            self.scale_torch_vol = (self.cmax_torch_vol - self.cmin_torch_vol) / 2
            self.base_torch_vol = self.cmin_torch_vol + self.scale_torch_vol
        self.cmin_vol, self.cmax_vol = self.cmin_torch_vol.cpu().numpy(), self.cmax_torch_vol.cpu().numpy()
        self.base_vol, self.scale_vol = self.base_torch_vol.cpu().numpy(), self.scale_torch_vol.cpu().numpy()
