"""EVOPF-v0 drop-in: the API of rpo.env.EVOPFEnv (rpo/env/electrical_grid/evopf.py:190-783) over the CUDA kernels.

Attributes and method signatures follow the reference so that rpo.algo.RPODDPG / RPOSAC (and anything else that
speaks the hard-constraint protocol) can use this class unchanged; every protocol method is one
kernel launch through the C ABI (rpo_b200.ops).  Tensors live on ``self.device`` (CUDA): there is no CPU path.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import evopf_constants as K
from .. import ops
from ..case14 import case14
from .loaders import DemandLoader, PriceLoader
from .spaces import Box, EnvBase


class _CompleteFn(torch.autograd.Function):
    """PFFunctionFn (evopf.py:787-910).  Saves only y (B,43): the backward kernel rebuilds the Jacobian from it,
    where the reference saves a (B,28,43) Jacobian and a (B,22,22) inverse (evopf.py:851-853)."""

    @staticmethod
    def forward(ctx, x, z, tol, max_iters, refine=True):
        y = ops.evopf_complete_fwd(x, z, tol=tol, max_iters=max_iters, refine=refine)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dl_dy):
        (y,) = ctx.saved_tensors
        dz, dx = ops.evopf_complete_bwd(y, dl_dy.contiguous(), need_dx=ctx.needs_input_grad[0])
        return dx, dz, None, None, None


class _ResidFn(torch.autograd.Function):
    """eq_resid / ineq_resid / ineq_dist with the autograd the reference's plain-torch versions have (evopf.py:520-567):
    the actor loss differentiates through them (rpo_ddpg.py:312-319, rpo_sac.py:326-334).  Forward = rpo_evopf_resid
    (one launch), backward = rpo_evopf_resid_bwd (one launch; the Jacobian is rebuilt from the action, nothing but
    (state, action) is saved).  kind: 0 eq_resid | 1 ineq_resid | 2 ineq_dist.  As in the reference, no gradient reaches
    the state through the inequality rows (the battery bounds pass through numpy, evopf.py:104-121); eq_resid's state
    gradient is -g on the [pd|qd] columns (:532,:537)."""

    @staticmethod
    def forward(ctx, state, action, kind):
        ctx.kind = kind
        ctx.save_for_backward(state, action)
        if kind == 0:
            return ops.evopf_resid(state, action, want_eq=True, want_ineq=False)[0]
        r = ops.evopf_resid(state, action, want_eq=False, want_ineq=True)[1]
        return torch.clamp_(r, 0) if kind == 2 else r

    @staticmethod
    def backward(ctx, g):
        state, action = ctx.saved_tensors
        g = g.contiguous()
        d_state = d_action = None
        if ctx.needs_input_grad[1]:
            if ctx.kind == 0:
                d_action = ops.evopf_resid_bwd(state, action, g_eq=g)
            else:
                d_action = ops.evopf_resid_bwd(state, action, g_ineq=g, ineq_mode=1 if ctx.kind == 2 else 0)
        if ctx.needs_input_grad[0] and ctx.kind == 0:
            d_state = torch.zeros_like(state)
            d_state[:, :K.XDIM] = -g
        return d_state, d_action, None


class Battery(object):
    """Host-side bookkeeping of the 5 batteries (evopf.py:23-149); the arithmetic of step / ineq_resid runs in the kernels."""

    def __init__(self, p_data, num, genbase, low=K.B_LOW, high=K.B_HIGH, p_min=K.B_PMIN, p_max=K.B_PMAX,
                 eta_in=K.ETA_IN, eta_out=K.ETA_OUT, init_strategy="empty", residual=1.0):
        self.p_data, self.num, self.genbase = p_data, num, genbase
        self.low, self.high, self.p_min, self.p_max = low, high, p_min, p_max
        self.eta_in, self.eta_out = eta_in, eta_out
        self.init_strategy, self.residual = init_strategy, residual
        self.state = None

    def reset(self):                                            # evopf.py:54-72
        state_p, _ = self.p_data.reset()
        state_p = state_p / self.genbase
        if self.init_strategy == "random":
            state_b = np.random.dirichlet(np.ones(self.num)) * self.residual + self.low
        elif self.init_strategy == "empty":
            state_b = np.ones(self.num) * (self.low + 0.1)
        elif self.init_strategy == "full":
            state_b = np.ones(self.num) * self.high
        else:
            raise NotImplementedError
        self.state = np.concatenate([state_b, state_p])
        return self.state.copy()

    def update_bound(self, soc: np.ndarray):                    # evopf.py:127-143 (float64 numpy)
        soc = np.asarray(soc, dtype=np.float64)                 # the reference subtracts from float64 arrays (:135-138)
        new_p_max = (1.0 / self.eta_in) * np.minimum(self.p_max, self.high - soc)
        new_p_min = self.eta_out * np.maximum(self.p_min, self.low - soc)
        return new_p_max, new_p_min


class EVOPFEnv(EnvBase):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 50}

    def __init__(self, device=None, demand_curve=None, price_curve=None, data_dir=None):
        self._device = torch.device(device) if device is not None else torch.device("cuda")
        ppc = case14()
        self.ppc = ppc
        # exogenous series: explicit argument > RPO_EVOPF_DATA_DIR > the importable reference's own pickles (evopf.py:212-216)
        data_dir = data_dir or os.environ.get("RPO_EVOPF_DATA_DIR")
        if not data_dir and demand_curve is None and price_curve is None:
            from .loaders import find_reference_data_dir
            data_dir = find_reference_data_dir()
            if not data_dir:
                import warnings
                warnings.warn("EVOPFEnv: reference demand/price pickles not found (no `rpo` package importable, no data_dir, "
                              "no RPO_EVOPF_DATA_DIR): using the synthetic daily curves of rpo_b200.env.loaders", stacklevel=2)
        self.data_dir = data_dir
        dp = os.path.join(data_dir, "demand.pickle") if data_dir else None
        pp = os.path.join(data_dir, "price.pickle") if data_dir else None
        self.data = DemandLoader(ppc=ppc, rho=0.5, q_rand=True, T=24, min_pf=0.9, max_pf=1.0, regularized=True,
                                 curve=demand_curve, pickle_path=dp)                     # evopf.py:212
        self.nahead = 24
        self.p_data = PriceLoader(T=24, ahead=self.nahead, regularized=True, reg_bias=0.1, reg_sigma=0.05,
                                  curve=price_curve, pickle_path=pp)                      # evopf.py:216
        self.nbus, self.ng, self.ne = K.NBUS, K.NG, K.NE
        self._box_dev = {}
        self._ineq_jac = {}
        self.genbase, self.baseMVA = K.GENBASE.copy(), K.BASE_MVA
        self.slack, self.pv, self.spv, self.pq = K.SLACK.copy(), K.PV.copy(), K.SPV.copy(), K.PQ.copy()
        self.slack_, self.pv_, self.spv_ = K.SLACK_.copy(), K.PV_.copy(), K.SPV_.copy()
        self.nslack, self.npv = K.NSLACK, K.NPV
        self.evs = Battery(self.p_data, num=self.ne, genbase=self.baseMVA, init_strategy="empty")
        self.we, self.wg = K.WE, K.WG
        self._xdim, self._ydim = K.XDIM, K.YDIM
        self._neq, self._nineq, self._nknowns = 2 * K.NBUS, 4 * K.NG + 2 * K.NBUS, K.NSLACK
        self.pg_start_yidx, self.qg_start_yidx, self.vm_start_yidx = K.PG0, K.QG0, K.VM0
        self.va_start_yidx, self.pe_start_yidx = K.VA0, K.PE0
        self._partial_vars, self._other_vars = K.PARTIAL_VARS.copy(), K.OTHER_VARS.copy()
        self._partial_actions, self._other_actions = K.PARTIAL_ACTIONS.copy(), K.OTHER_ACTIONS.copy()
        self.pg_pv_zidx, self.vm_spv_zidx = K.PG_PV_ZIDX.copy(), K.VM_SPV_ZIDX.copy()
        self.pe_pv_zidx, self.pe_spv_zidx = K.PE_PV_ZIDX.copy(), K.PE_SPV_ZIDX.copy()
        high = np.array([10.0] * 14 + [10.0] * 14 + [self.evs.high] * 5 + [240.0] * 24)
        low = np.array([-10.0] * 14 + [-10.0] * 14 + [self.evs.low] * 5 + [0.0] * 24)
        self.action_space = Box(low=K.ACTION_LOW.copy(), high=K.ACTION_HIGH.copy(), dtype=np.float32)
        self.observation_space = Box(low=low, high=high, dtype=np.float32)
        self.state_dim, self.action_dim = K.STATE_DIM, K.YDIM
        self.eq_num, self.ineq_num = K.EQ_NUM, K.INEQ_NUM
        self.newton_tol, self.newton_max_iters = K.NEWTON_TOL, K.NEWTON_MAX_ITERS         # evopf.py:786
        self.state = None
        self.action = None
        self.volatile = True

    # ---- properties of the reference (evopf.py:457-499) -------------------------------------------------
    partial_actions = property(lambda self: self._partial_actions)
    other_actions = property(lambda self: self._other_actions)
    partial_vars = property(lambda self: self._partial_vars)
    other_vars = property(lambda self: self._other_vars)
    xdim = property(lambda self: self._xdim)
    ydim = property(lambda self: self._ydim)
    neq = property(lambda self: self._neq)
    nineq = property(lambda self: self._nineq)
    nknowns = property(lambda self: self._nknowns)
    partial_unknown_vars = property(lambda self: np.concatenate([K.PG0 + K.PV_, K.VM0 + K.SPV]))      # evopf.py:291,474
    device = property(lambda self: self._device)

    @property
    def box_constraint(self):
        return self.action_space.low, self.action_space.high

    @property
    def box_constraint_partial(self):
        return self.action_space.low[self.partial_actions], self.action_space.high[self.partial_actions]

    # ---- helpers ----------------------------------------------------------------------------------------
    def _t(self, a):
        if isinstance(a, torch.Tensor):
            t = a.to(device=self._device, dtype=torch.float32)
        else:
            t = torch.as_tensor(np.asarray(a), dtype=torch.float32, device=self._device)
        return t.view(1, -1) if t.dim() == 1 else t

    def decompose(self, state):                                 # evopf.py:454-455
        return state[:2 * self.nbus], state[2 * self.nbus:]

    def obj_fn(self, action):                                   # evopf.py:509-518 (float64 result: genbase is a float64 array there)
        """Generation cost of `action` (the reward of `step` is computed inside rpo_evopf_step; this method is the
        reference's stand-alone helper, a few torch operations)."""
        pg = self._t(action)[:, :self.ng].double()
        pg_mw = pg * torch.as_tensor(self.genbase, dtype=torch.float64, device=pg.device)
        quad = torch.as_tensor(K.QUAD_COSTS, dtype=torch.float32, device=pg.device).double()
        lin = torch.as_tensor(K.LIN_COSTS, dtype=torch.float32, device=pg.device).double()
        cost = (quad * pg_mw ** 2).sum(dim=1) + (lin * pg_mw).sum(dim=1) + K.CONST_COST
        return cost / (float(np.mean(self.genbase)) ** 2)

    def get_action_vars(self, action):                          # evopf.py:501-507
        return (action[:, :self.ng], action[:, self.ng:2 * self.ng], action[:, 2 * self.ng:2 * self.ng + self.nbus],
                action[:, -self.ne - self.nbus:-self.ne], action[:, -self.ne:])

    # ---- hard-constraint protocol -----------------------------------------------------------------------
    def complete_partial(self, state, action_partial, refine=True):          # evopf.py:724-727
        """refine=True (default): float32 Newton finished by a float64-residual chord step (equality residual below the
        float32 reference's); refine=False: the plain float32 Newton sequence of the fused projection kernel, which
        rpo_b200.algo.process_action uses so that its autograd path and its fused no-grad path give identical actions."""
        if action_partial.dim() == 1:
            action_partial = action_partial.view(1, -1)
        return _CompleteFn.apply(state[:, :self.xdim], action_partial, self.newton_tol, self.newton_max_iters, refine)

    def eq_resid(self, state, action):                          # evopf.py:520-546 (differentiable, like the reference's)
        return _ResidFn.apply(state, action, 0)

    def ineq_resid(self, state, action):                        # evopf.py:548-563
        return _ResidFn.apply(state, action, 1)

    def ineq_dist(self, state, action):                         # evopf.py:565-567
        return _ResidFn.apply(state, action, 2)

    def ineq_partial_grad(self, state, action, eps=0.0):        # evopf.py:596-612 (eps is ignored there too)
        return ops.evopf_direction(state, action)

    def ineq_grad_new(self, state, action, eps=0.0):            # evopf.py:590-594: ineq_jac^T . 1[ineq_dist > 0]
        """The constant +-1 selector `ineq_jac` (evopf.py:663-707) is never built: the product is the sign-of-violation
        vector, assembled on the device from the residual kernel's output."""
        m = (self.ineq_resid(state, action) > 0).to(torch.float32)
        return self._fold_ineq_rows(m)

    def _fold_ineq_rows(self, w):
        """ineq_jac^T . w for a (B,58) weight on the inequality rows: the selector has one +1 (upper bound) and one -1
        (lower bound) per bounded variable, so the product is a difference of column blocks."""
        ng, nb = self.ng, self.nbus
        d = torch.zeros(w.shape[0], self.ydim, dtype=w.dtype, device=w.device)
        d[:, 0:ng] = w[:, 0:ng] - w[:, ng:2 * ng]
        d[:, ng:2 * ng] = w[:, 2 * ng:3 * ng] - w[:, 3 * ng:4 * ng]
        d[:, 2 * ng:2 * ng + nb] = w[:, 4 * ng:4 * ng + nb] - w[:, 4 * ng + nb:4 * ng + 2 * nb]
        d[:, -self.ne:] = w[:, 4 * ng + 2 * nb:5 * ng + 2 * nb] - w[:, 5 * ng + 2 * nb:]
        return d

    def ineq_grad(self, state, action, eps=0.0):                # evopf.py:584-588 (corr_mode != 0 only)
        dist = self.ineq_dist(state, action)
        return 2 * self._fold_ineq_rows(dist + (dist > 0) * eps)

    def ineq_jac(self, state, action):                          # evopf.py:663-707: constant (58,43) selector, expanded
        key = str(action.device)
        if key not in self._ineq_jac:
            eye = torch.eye(self.ineq_num, dtype=torch.float32, device=action.device)
            self._ineq_jac[key] = self._fold_ineq_rows(eye).contiguous()          # row r = ineq_jac[r, :]
        jac = self._ineq_jac[key]
        return jac.unsqueeze(0).expand(action.shape[0], *jac.shape)

    def ineq_dist_np(self, state, action):                      # evopf.py:569-572
        return self.ineq_dist(self._t(state), self._t(action)).cpu().numpy()

    def eq_resid_np(self, state, action):                       # evopf.py:574-577
        return self.eq_resid(self._t(state), self._t(action)).cpu().numpy()

    # fused kernels used by rpo_b200.algo (no counterpart method in the reference: they fuse its call sequences)
    def correct(self, state, action, lr, momentum, corr_eps, n_steps):
        return ops.evopf_correct(state, action, lr, momentum, corr_eps, n_steps)

    def project(self, state, action_partial, lr, momentum, corr_eps, n_steps):
        if action_partial.dim() == 1:
            action_partial = action_partial.view(1, -1)
        yc, y, flags, _ = ops.evopf_project(state, action_partial, lr, momentum, corr_eps, n_steps, tol=self.newton_tol,
                                            max_iters=self.newton_max_iters, want_y=True)
        return yc, y, flags

    def eq_jac(self, action):                                   # evopf.py:614-661 -> (B,28,43)
        """Dense d eq_resid / d action, for inspection and for corr_mode != 0 callers.  The kernels never build it (its
        entries live inside the generated sparse LU); this is the same branch-term formulation written as a
        dozen dense torch operations, NOT a hot-path routine: 4.8 KB per instance."""
        _, _, vm, va, _ = self.get_action_vars(action)
        dev, dt = action.device, action.dtype
        G = torch.as_tensor(K.G, dtype=dt, device=dev)
        Bm = torch.as_tensor(K.B, dtype=dt, device=dev)
        vr, vi = vm * torch.cos(va), vm * torch.sin(va)
        al = G * vr[:, None, :] - Bm * vi[:, None, :]                      # alpha_ij, beta_ij: Y_ij (vr_j + j vi_j)
        be = Bm * vr[:, None, :] + G * vi[:, None, :]
        T1 = vr[:, :, None] * al + vi[:, :, None] * be
        T2 = vi[:, :, None] * al - vr[:, :, None] * be
        P, Q = torch.diag_embed(T1.sum(dim=2)), torch.diag_embed(T2.sum(dim=2))
        inv_v = (1.0 / vm)[:, None, :]
        dP_dva, dP_dvm = T2 - Q, (T1 + P) * inv_v
        dQ_dva, dQ_dvm = P - T1, (T2 + Q) * inv_v
        n, nb, ng = action.shape[0], self.nbus, self.ng
        sel = torch.zeros(nb, ng, dtype=dt, device=dev)
        sel[torch.as_tensor(self.spv, device=dev), torch.arange(ng, device=dev)] = 1.0
        sel = sel.expand(n, nb, ng)
        zero = torch.zeros(n, nb, ng, dtype=dt, device=dev)
        return torch.cat([torch.cat([sel, zero, -dP_dvm, -dP_dva, -sel], dim=2),         # -I for pe: the reference's sign (:639-640)
                          torch.cat([zero, sel, -dQ_dvm, -dQ_dva, zero], dim=2)], dim=1)

    def eq_grad(self, state, action):                           # evopf.py:579-582
        r = self.eq_resid(state, action)
        return 2 * self.eq_jac(action).transpose(1, 2).bmm(r.unsqueeze(-1)).squeeze(-1)

    # ---- volatile action box (evopf.py:769-783, Battery.update_bound :127-143) --------------------------
    def update(self, state, full=False):
        if isinstance(state, torch.Tensor):
            s = state.detach().cpu().numpy()
        else:
            s = np.asarray(state)
        if s.ndim == 1:
            s = s.reshape(1, -1)
        n = s.shape[0]
        soc = s[:, -self.ne - self.nahead:-self.nahead]
        new_p_max, new_p_min = self.evs.update_bound(soc)
        low, high = self.box_constraint if full else self.box_constraint_partial
        low = low[None].repeat(n, axis=0)
        high = high[None].repeat(n, axis=0)
        low[:, -self.ne:] = new_p_min
        high[:, -self.ne:] = new_p_max
        return low, high

    def update_device(self, state: torch.Tensor, full: bool = False):
        """Same box as update(), computed on the device without a host round trip: two launches of rpo_box_fwd
        (bounds in float64 then rounded, as numpy does).  Policies should not need it: `rpo_b200.box.BoxConstraint`
        squashes / clips against the volatile box in one launch."""
        key = (str(state.device), bool(full))
        if key not in self._box_dev:
            lo, hi = self.box_constraint if full else self.box_constraint_partial
            self._box_dev[key] = (torch.tensor(lo, dtype=torch.float32, device=state.device),
                                  torch.tensor(hi, dtype=torch.float32, device=state.device))
        lo, hi = self._box_dev[key]
        return (ops.box_fwd(None, lo, hi, ops.BOX_LOW, state, self.ne), ops.box_fwd(None, lo, hi, ops.BOX_HIGH, state, self.ne))

    # ---- gym API (evopf.py:348-380) -----------------------------------------------------------------------
    def reset(self):
        state_grid, _ = self.data.reset()
        state_evs = self.evs.reset()
        self.state = np.concatenate([state_grid, state_evs])
        self.action = None
        return self.state.copy()

    def step(self, action):
        action = np.asarray(action)
        price_next, _ = self.p_data.fetch()
        price_next = price_next / self.baseMVA
        grid_next, done = self.data.fetch()
        exo = np.concatenate([grid_next, price_next]).astype(np.float32)
        ns, rew, eqv, iqv = ops.evopf_step(self._t(self.state), self._t(action), self._t(exo))
        next_state = ns[0].double().cpu().numpy()
        info = {"ineq_viol": iqv.cpu().numpy(), "eq_viol": eqv.cpu().numpy()}
        self.state = next_state
        self.evs.state = next_state[28:].copy()
        self.action = action
        return next_state, float(rew[0].item()), done, info

    def step_batch(self, state, action, next_exo):
        """Vectorised step over independent replicas: (B,57), (B,43), (B,52) device tensors -> next_state, reward,
        eq_viol, ineq_viol (all device tensors); the exogenous rows come from the caller's loaders / generators."""
        return ops.evopf_step(state, action, next_exo)

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the hot path (SURVEY.md §2 #17)")

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]
