"""GPU twin of tests/test_actor_loss_autograd_cpu.py: d(actor_loss)/dz through the drop-in environments (CUDA kernels
via the C ABI) against the oracle's torch autograd of the same loss (rpo/algo/rpo_ddpg.py:307-324: -Q + nju(ineq_dist)),
with nju ~ U(0,1), for EVOPF, Safe Cartpole and Spring Pendulum.  Tolerance: reduced gradients 1e-3 relative."""
import pytest
import torch

from rpo_b200 import synth
from tests.parity import GRAD_RTOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _weights(n_ineq, n_eq, n_a, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_a, generator=g), torch.rand(n_ineq, generator=g), torch.rand(n_eq, generator=g)


def _loss_grad_env(env, state, z, q_w, nju_w, lamb_w=None):
    z = z.clone().requires_grad_(True)
    a = env.complete_partial(state, z)
    viol = env.ineq_dist(state, a)
    assert viol.requires_grad and a.requires_grad
    loss = -(a @ q_w.view(-1, 1)) + viol @ nju_w.view(-1, 1)
    if lamb_w is not None:
        loss = loss + env.eq_resid(state, a) @ lamb_w.view(-1, 1)
    loss.mean().backward()
    return z.grad.detach().cpu(), viol.detach().cpu()


@pytest.mark.parametrize("dist", ["D1", "D2"])
@pytest.mark.parametrize("B", [1, 7, 256, 4096])
def test_evopf_actor_loss_gradient(evopf_oracle, B, dist):
    from rpo_b200.env import EVOPFEnv
    orc = evopf_oracle
    env = EVOPFEnv(device=DEV)
    st = synth.evopf_states(B, seed=5)
    z = synth.evopf_partial_actions(st, seed=6, dist=dist)
    q_w, nju_w, lamb_w = _weights(58, 28, 43, 7)
    for lw in (None, lamb_w):
        g_b, viol_b = _loss_grad_env(env, st.to(DEV), z.to(DEV), q_w.to(DEV), nju_w.to(DEV), None if lw is None else lw.to(DEV))
        # oracle: completion with saved Jacobian + torch autograd of the residual functions at y
        y_o, jac, inv, _ = orc.complete_partial_fwd(st, z, return_saved=True)
        yy = y_o.clone().requires_grad_(True)
        viol_o = orc.ineq_dist(st, yy)
        loss = -(yy @ q_w.view(-1, 1)) + viol_o @ nju_w.view(-1, 1)
        if lw is not None:
            loss = loss + orc.eq_resid(st, yy) @ lw.view(-1, 1)
        loss.mean().backward()
        _, dz_o = orc.complete_partial_bwd(yy.grad, jac, inv)
        same = ((viol_o.detach() > 0) == (viol_b > 0)).all(dim=1)
        assert float(same.float().mean()) >= (0.97 if B >= 256 else 0.5)
        if bool(same.any()):
            assert float(rel_err(g_b[same], dz_o[same]).max()) <= GRAD_RTOL


def test_evopf_fused_penalty_backward_on_device():
    """one launch (rpo_evopf_complete_bwd_penalty) == resid_bwd + add + complete_bwd, on full TMA tiles and a ragged tail"""
    from rpo_b200 import ops
    for B in (5, 128 * 3 + 17, 1 << 17):
        st = synth.evopf_states(B, seed=21).to(DEV)
        z = synth.evopf_partial_actions(st.cpu(), seed=22).to(DEV)
        y = ops.evopf_complete_fwd(st, z)
        g = synth.upstream_grad(B, 43, seed=23).to(DEV)
        gen = torch.Generator().manual_seed(24)
        w = torch.rand(58, generator=gen).to(DEV)
        rs = torch.rand(B, generator=gen).to(DEV)
        for row_scale in (None, rs):
            up = (w.view(1, -1).expand(B, 58) * (1.0 if row_scale is None else row_scale.view(-1, 1))).contiguous()
            dy = ops.evopf_resid_bwd(st, y, g_ineq=up, ineq_mode=1)
            dz_c, dx_c = ops.evopf_complete_bwd(y, g + dy)
            dz_f, dx_f = ops.evopf_complete_bwd_penalty(st, y, g, w, row_scale)
            assert float(rel_err(dz_f, dz_c).max()) < 1e-5 and float(rel_err(dx_f, dx_c).max()) < 1e-5
        # state passed as the compact [pd|qd|soc] view as well (leading dimension 57, 33 columns used)
        dz_v, _ = ops.evopf_complete_bwd_penalty(st[:, :33], y, g, w, None)
        dz_f, _ = ops.evopf_complete_bwd_penalty(st, y, g, w, None)
        assert torch.equal(dz_v, dz_f)


def test_evopf_resid_bwd_matches_oracle_autograd(evopf_oracle):
    from rpo_b200 import ops
    orc = evopf_oracle
    for B in (3, 128, 1000):
        st = synth.evopf_states(B, seed=31)
        z = synth.evopf_partial_actions(st, seed=32)
        y = orc.complete_partial_fwd(st, z)
        gen = torch.Generator().manual_seed(33)
        ge, gi = torch.randn(B, 28, generator=gen), torch.randn(B, 58, generator=gen)
        for mode in (0, 1):
            yy = y.clone().requires_grad_(True)
            out = (orc.eq_resid(st, yy) * ge).sum() + ((orc.ineq_dist(st, yy) if mode else orc.ineq_resid(st, yy)) * gi).sum()
            out.backward()
            dy = ops.evopf_resid_bwd(st.to(DEV), y.to(DEV), g_eq=ge.to(DEV), g_ineq=gi.to(DEV), ineq_mode=mode)
            assert float(rel_err(dy, yy.grad).max()) <= GRAD_RTOL
            dy_e = ops.evopf_resid_bwd(st.to(DEV), y.to(DEV), g_eq=ge.to(DEV))
            dy_i = ops.evopf_resid_bwd(st.to(DEV), y.to(DEV), g_ineq=gi.to(DEV), ineq_mode=mode)
            assert float((dy_e + dy_i - dy).abs().max()) <= 1e-5 * float(dy.abs().max())


@pytest.mark.parametrize("p", [0, 1])
def test_cart_actor_loss_gradient(p):
    from oracle.rpo_oracle import CartOracle
    from rpo_b200.env import CartSafeEnv
    orc = CartOracle(partial_idx=p)
    env = CartSafeEnv(device=DEV, partial_idx=p)
    B = 4099
    st, z = synth.cart_inputs(B, seed=31)
    z = z * 1.6
    q_w, nju_w, lamb_w = _weights(6, 1, 2, 32)
    g_b, viol = _loss_grad_env(env, st.to(DEV), z.to(DEV), q_w.to(DEV), nju_w.to(DEV), lamb_w.to(DEV))
    zz = z.clone().requires_grad_(True)
    a = orc.complete_partial(st, zz)
    loss = -(a @ q_w.view(-1, 1)) + orc.ineq_dist(st, a) @ nju_w.view(-1, 1) + orc.eq_resid(st, a) @ lamb_w.view(-1, 1)
    loss.mean().backward()
    assert float((viol > 0).float().mean()) > 0.02
    assert float((g_b - zz.grad).abs().max()) <= GRAD_RTOL * float(zz.grad.abs().max())


def test_pendulum_actor_loss_gradient():
    from oracle.rpo_oracle import PenOracle
    from rpo_b200.env import SpringPendulumEnv
    orc = PenOracle()
    env = SpringPendulumEnv(device=DEV)
    B = 4099
    obs, z = synth.pen_inputs(B, seed=41)
    q_w, nju_w, lamb_w = _weights(1, 1, 2, 42)
    g_b, viol = _loss_grad_env(env, obs.to(DEV), z.to(DEV), q_w.to(DEV), nju_w.to(DEV), lamb_w.to(DEV))
    zz = z.clone().requires_grad_(True)
    a = orc.complete_partial(obs, zz)
    loss = -(a @ q_w.view(-1, 1)) + orc.ineq_dist(obs, a) @ nju_w.view(-1, 1) + orc.eq_resid(obs, a) @ lamb_w.view(-1, 1)
    loss.mean().backward()
    assert float((viol > 0).float().mean()) > 0.02
    assert float((g_b - zz.grad).abs().max()) <= GRAD_RTOL * float(zz.grad.abs().max())
