"""Runs the UNMODIFIED reference beside the oracle (only where /root/reference exists: the build container)."""
import pytest
import torch

from oracle.ref_loader import reference_available
from tests.parity import rel_err

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def test_evopf_oracle_equals_reference_code(evopf_oracle):
    from oracle.ref_loader import load_reference, make_cpu_env
    from oracle.rpo_oracle import grad_steps
    from rpo_b200 import synth
    rpo = load_reference()
    env = make_cpu_env("evopf")
    orc = evopf_oracle
    st = synth.evopf_states(300, seed=7)
    z = synth.evopf_partial_actions(st, seed=8)
    g = synth.upstream_grad(300, 43, seed=9)
    zr, sr = z.clone().requires_grad_(True), st.clone().requires_grad_(True)
    y_ref = env.complete_partial(sr, zr)
    y_ref.backward(g)
    y, jac, inv, _ = orc.complete_partial_fwd(st, z, return_saved=True)
    dx, dz = orc.complete_partial_bwd(g, jac, inv)
    assert rel_err(y, y_ref).max() < 2e-5
    assert rel_err(dz, zr.grad).max() < 1e-4 and rel_err(dx, sr.grad[:, :28]).max() < 1e-4
    yd = y_ref.detach()
    assert (env.eq_jac(yd) - orc.eq_jac(yd)).abs().max() < 2e-5
    assert (env.eq_resid(st, yd) - orc.eq_resid(st, yd)).abs().max() == 0
    assert (env.ineq_grad_new(st, yd) - orc.ineq_grad_new(st, yd)).abs().max() == 0
    assert rel_err(orc.ineq_partial_grad(st, yd), env.ineq_partial_grad(st, yd)).max() < 1e-4

    class A(object):
        pass
    a = A()
    a.env, a.corr_eps, a.max_steps, a.corr_mode, a.grad_eps = env, 1e-5, 10, 0, 0.02
    a.corr_lr, a.corr_momentum, a.eval_lr, a.eval_steps = 1e-4, 0.0, 1e-4, 50
    yc_ref = rpo.algo.RPODDPG.grad_steps(a, st, yd, train=True)
    yc, k = grad_steps(orc, st, yd, 1e-4, 0.0, 1e-5, 10)
    assert k == 10 and rel_err(yc, yc_ref).max() < 2e-5
    yc_sac = rpo.algo.RPOSAC.grad_steps(a, st, yd, train=True)       # byte-identical logic (rpo_sac.py:280-319)
    assert torch.equal(yc_sac, yc_ref)
