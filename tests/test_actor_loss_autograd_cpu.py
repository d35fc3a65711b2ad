"""The actor loss of the reference differentiates through ineq_dist / eq_resid (rpo/algo/rpo_ddpg.py:307-324,
rpo/algo/rpo_sac.py:321-340: `loss = -Q(s, a) + nju(ineq_dist(s, a))`).  The drop-in environments must give the same
d(actor_loss)/dz as the UNMODIFIED reference's autograd — for EVOPF, Safe Cartpole and Spring Pendulum.

No GPU here: `rpo_b200.ops` is the host simulator (the very per-instance code the kernels run, g++-compiled).  The GPU
twin (kernels against the oracle's autograd) is tests/test_actor_loss_autograd_gpu.py."""
import numpy as np
import pytest
import torch

from oracle.ref_loader import reference_available
from rpo_b200 import synth
from tests.parity import GRAD_RTOL, rel_err

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


@pytest.fixture()
def hs(monkeypatch):
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)
    return hostsim_ops


def _actor_loss_grad(env, state, z, q_w, nju_w, lamb_w=None, use_resid=False):
    """d/dz of the reference's actor loss with a linear stand-in critic: (-a.q_w + nju(ineq_dist))[.mean()]."""
    z = z.clone().requires_grad_(True)
    a = env.complete_partial(state, z)
    viol = env.ineq_resid(state, a) if use_resid else env.ineq_dist(state, a)
    loss = -(a @ q_w.view(-1, 1)) + viol @ nju_w.view(-1, 1)
    if lamb_w is not None:
        # rpo_ddpg.py:320 has lamb(|eq_viol|) (commented out); |.| of a residual that sits at float32 noise after the
        # completion has a random sign, so the linear form is what can be compared
        loss = loss + env.eq_resid(state, a) @ lamb_w.view(-1, 1)
    loss.mean().backward()
    return z.grad.detach().clone(), viol.detach()


def _weights(n_ineq, n_eq, n_a, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n_a, generator=g), torch.rand(n_ineq, generator=g), torch.rand(n_eq, generator=g))


@pytest.mark.parametrize("dist", ["D1", "D2"])
def test_evopf_actor_loss_gradient_matches_reference_autograd(hs, dist):
    from oracle.ref_loader import make_cpu_env
    from rpo_b200.env import EVOPFEnv
    ref = make_cpu_env("evopf")
    env = EVOPFEnv(device="cpu")
    B = 192
    st = synth.evopf_states(B, seed=5)
    z = synth.evopf_partial_actions(st, seed=6, dist=dist)
    q_w, nju_w, lamb_w = _weights(58, 28, 43, 7)
    for kw in (dict(), dict(use_resid=True), dict(lamb_w=lamb_w)):
        g_ref, viol_ref = _actor_loss_grad(ref, st, z, q_w, nju_w, **kw)
        g_b, viol_b = _actor_loss_grad(env, st, z, q_w, nju_w, **kw)
        assert bool(env.ineq_dist(st, env.complete_partial(st, z.clone().requires_grad_(True))).requires_grad)
        # rows whose violated set agrees (an inequality residual within float32 round-off of 0 may flip)
        same = ((viol_ref > 0) == (viol_b > 0)).all(dim=1)
        assert float(same.float().mean()) > 0.97
        err = rel_err(g_b[same], g_ref[same])
        assert float(err.max()) <= GRAD_RTOL, (kw.keys(), float(err.max()))
        # the penalty term really contributes: without it the gradient is different
        g_nop, _ = _actor_loss_grad(env, st, z, q_w, torch.zeros(58))
        assert float((g_b - g_nop).abs().max()) > 1e-4


def test_evopf_resid_functions_backward_individually(hs):
    from oracle.ref_loader import make_cpu_env
    from rpo_b200.env import EVOPFEnv
    ref = make_cpu_env("evopf")
    env = EVOPFEnv(device="cpu")
    B = 64
    st = synth.evopf_states(B, seed=11)
    z = synth.evopf_partial_actions(st, seed=12)
    y = env.complete_partial(st, z).detach()
    gen = torch.Generator().manual_seed(13)
    for name, n in (("eq_resid", 28), ("ineq_resid", 58), ("ineq_dist", 58)):
        up = torch.randn(B, n, generator=gen)
        grads = []
        for e in (ref, env):
            yy = y.clone().requires_grad_(True)
            ss = st.clone().requires_grad_(name == "eq_resid")
            out = getattr(e, name)(ss, yy)
            out.backward(up)
            grads.append((yy.grad.clone(), None if ss.grad is None else ss.grad.clone(), out.detach()))
        (gy_r, gs_r, o_r), (gy_b, gs_b, o_b) = grads
        assert float((o_r - o_b).abs().max()) < 2e-5, name
        assert float(rel_err(gy_b, gy_r).max()) <= GRAD_RTOL, name
        if name == "eq_resid":
            assert torch.equal(gs_b, gs_r)                     # -g on [pd|qd], zero elsewhere


def test_evopf_fused_penalty_backward_equals_composition(hs):
    """rpo_evopf_complete_bwd_penalty = GRG backward of (dl_dy + ineq_dist's autograd), in one launch."""
    from rpo_b200 import ops
    B = 96
    st = synth.evopf_states(B, seed=21)
    z = synth.evopf_partial_actions(st, seed=22)
    y = ops.evopf_complete_fwd(st, z)
    g = synth.upstream_grad(B, 43, seed=23)
    gen = torch.Generator().manual_seed(24)
    w = torch.rand(58, generator=gen)
    rs = torch.rand(B, generator=gen)
    for row_scale in (None, rs):
        up = w.view(1, -1).expand(B, 58) * (1.0 if row_scale is None else row_scale.view(-1, 1))
        dy = ops.evopf_resid_bwd(st, y, g_ineq=up.contiguous(), ineq_mode=1)
        dz_c, dx_c = ops.evopf_complete_bwd(y, g + dy)
        dz_f, dx_f = ops.evopf_complete_bwd_penalty(st, y, g, w, row_scale)
        assert float(rel_err(dz_f, dz_c).max()) < 1e-5 and float(rel_err(dx_f, dx_c).max()) < 1e-5


@pytest.mark.parametrize("p", [0, 1])
def test_cart_actor_loss_gradient_matches_reference_autograd(hs, p):
    from oracle.ref_loader import make_cpu_env
    from rpo_b200.env import CartSafeEnv
    ref = make_cpu_env("cart")
    ref.partial_actions = np.array([p])
    ref.other_actions = np.array([1 - p])
    ref.diff_eq_partial = ref.diff_eq[:, ref.partial_actions]
    ref.diff_eq_other_inv = torch.inverse(ref.diff_eq[:, ref.other_actions])
    env = CartSafeEnv(device="cpu", partial_idx=p)
    B = 256
    st, z = synth.cart_inputs(B, seed=31)
    z = z * 1.6                                              # push a good share of rows past the force limits
    q_w, nju_w, lamb_w = _weights(6, 1, 2, 32)
    for kw in (dict(), dict(use_resid=True), dict(lamb_w=lamb_w)):
        g_ref, v = _actor_loss_grad(ref, st, z, q_w, nju_w, **kw)
        g_b, _ = _actor_loss_grad(env, st, z, q_w, nju_w, **kw)
        assert float((v > 0).float().mean()) > 0.02
        assert float((g_b - g_ref).abs().max()) <= GRAD_RTOL * float(g_ref.abs().max())


def test_pendulum_actor_loss_gradient_matches_reference_autograd(hs):
    from oracle.ref_loader import make_cpu_env
    from rpo_b200.env import SpringPendulumEnv
    ref = make_cpu_env("pen")
    env = SpringPendulumEnv(device="cpu")
    B = 256
    obs, z = synth.pen_inputs(B, seed=41)
    q_w, nju_w, lamb_w = _weights(1, 1, 2, 42)
    for kw in (dict(), dict(use_resid=True), dict(lamb_w=lamb_w)):
        g_ref, v = _actor_loss_grad(ref, obs, z, q_w, nju_w, **kw)
        g_b, _ = _actor_loss_grad(env, obs, z, q_w, nju_w, **kw)
        assert float((v > 0).float().mean()) > 0.02
        assert float((g_b - g_ref).abs().max()) <= GRAD_RTOL * float(g_ref.abs().max())


def test_clamp_gradient_convention_at_zero(hs):
    """torch.clamp(resid, 0) passes the gradient where resid >= 0 (not > 0): pg == pmin == 0 exactly is the common
    case of a saturated policy output.  The kernels' mask must use the same convention as the reference's autograd."""
    from oracle.ref_loader import make_cpu_env
    from rpo_b200.env import EVOPFEnv
    ref = make_cpu_env("evopf")
    env = EVOPFEnv(device="cpu")
    st = synth.evopf_states(4, seed=51)
    y = env.complete_partial(st, synth.evopf_partial_actions(st, seed=52, dist="D2")).detach()
    y[:, 1:5] = 0.0                                          # pg_pv exactly on pmin = 0: rows pmin - pg == 0
    up = torch.ones(4, 58)
    out = []
    for e in (ref, env):
        yy = y.clone().requires_grad_(True)
        e.ineq_dist(st, yy).backward(up)
        out.append(yy.grad.clone())
    assert torch.equal(out[0][:, 1:5], out[1][:, 1:5])
    assert float(out[0][:, 1:5].abs().max()) == 1.0          # the reference does pass -1 through the lower-bound rows


def test_fused_completion_with_penalty_matches_reference_actor_loss(hs):
    """rpo_b200.algo.complete_partial_with_penalty (one backward launch) against the reference's autograd of
    -Q + nju(ineq_dist) w.r.t. z AND w.r.t. the dual weights."""
    from oracle.ref_loader import make_cpu_env
    from rpo_b200.algo import complete_partial_with_penalty
    from rpo_b200.env import EVOPFEnv
    ref = make_cpu_env("evopf")
    env = EVOPFEnv(device="cpu")
    B = 128
    st = synth.evopf_states(B, seed=61)
    z = synth.evopf_partial_actions(st, seed=62)
    q_w, nju_w, _ = _weights(58, 28, 43, 63)
    z_r, w_r = z.clone().requires_grad_(True), nju_w.view(1, -1).clone().requires_grad_(True)
    a = ref.complete_partial(st, z_r)
    viol_r = ref.ineq_dist(st, a)
    (-(a @ q_w.view(-1, 1)) + torch.nn.functional.linear(viol_r, w_r)).mean().backward()
    z_b, w_b = z.clone().requires_grad_(True), nju_w.view(1, -1).clone().requires_grad_(True)
    a_b, pen = complete_partial_with_penalty(env, st, z_b, w_b)
    (-(a_b @ q_w.view(-1, 1)) + pen).mean().backward()
    viol_b = env.ineq_dist(st, a_b.detach())
    same = ((viol_r > 0) == (viol_b > 0)).all(dim=1)
    assert float(same.float().mean()) > 0.97
    assert float(rel_err(z_b.grad[same], z_r.grad[same]).max()) <= GRAD_RTOL
    assert float((w_b.grad - w_r.grad).abs().max()) <= 1e-4 * float(w_r.grad.abs().max())
