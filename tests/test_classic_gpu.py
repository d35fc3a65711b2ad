"""GPU parity of the Safe-Cartpole / Spring-Pendulum kernels against the oracle and the reference goldens."""
import os

import numpy as np
import pytest
import torch

from tests.parity import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("p", [0, 1])
def test_cart_golden(p):
    from rpo_b200 import ops
    g = np.load(os.path.join(GOLD, "cart_p%d.npz" % p))
    assert int(g["partial_idx"]) == p
    z = torch.tensor(g["z"]).cuda()
    a = ops.cart_complete_fwd(z, p)
    assert np.abs(a.cpu().numpy() - g["a"]).max() <= 1e-6
    dz = ops.cart_complete_bwd(torch.tensor(g["g"]).cuda(), p)
    assert np.abs(dz.cpu().numpy() - g["dz"]).max() <= 1e-6
    eq, iq = ops.cart_resid(a)
    assert np.abs(eq.cpu().numpy() - g["eq_resid"]).max() <= 4e-6          # 2 ulp at |resid| ~ 30
    assert np.abs(iq.cpu().numpy() - g["ineq_resid"]).max() <= 4e-6
    d = ops.cart_direction(torch.tensor(g["a"]).cuda(), p)
    assert np.abs(d.cpu().numpy() - g["dir"]).max() <= 1e-5
    ac, flags = ops.cart_correct(torch.tensor(g["a"]).cuda(), p, 2e-2, 0.0, 1e-5, 10)
    assert np.abs(ac.cpu().numpy() - g["a_corr"]).max() <= 1e-4
    ap, a1, _ = ops.cart_project(z, p, 2e-2, 0.0, 1e-5, 10, want_a=True)
    assert torch.equal(a1, a)
    assert np.abs(ap.cpu().numpy() - g["a_corr"]).max() <= 1e-4
    # eval branch: the reference stops when the global max meets corr_eps (rpo_ddpg.py:290-291)
    ace, fl = ops.cart_correct(torch.tensor(g["a"]).cuda(), p, 2e-2, 0.0, 1e-5, 50)
    fl = fl.cpu().numpy()
    k = next((i for i in range(1, 50) if fl[i] == 0), 50)
    assert k == int(g["eval_steps"])
    ace, _ = ops.cart_correct(torch.tensor(g["a"]).cuda(), p, 2e-2, 0.0, 1e-5, k)
    assert np.abs(ace.cpu().numpy() - g["a_corr_eval"]).max() <= 1e-4
    # step
    ns, done, eqv, iqv = ops.cart_step(torch.tensor(g["step_state"], dtype=torch.float32).cuda(),
                                       torch.tensor(g["step_action"]).cuda())
    assert np.abs(ns.cpu().numpy() - g["step_next"]).max() <= 1e-5 * max(1.0, np.abs(g["step_next"]).max())
    assert (done.cpu().numpy().astype(bool) == g["step_done"]).all()
    assert np.abs(iqv.cpu().numpy() - g["step_ineq_viol"]).max() <= 1e-5
    assert np.abs(eqv.cpu().numpy().ravel() - g["step_eq_viol"].ravel()).max() <= 1e-5


def test_cart_oracle_large():
    from oracle.rpo_oracle import CartOracle, grad_steps
    from rpo_b200 import ops, synth
    for p in (0, 1):
        orc = CartOracle(partial_idx=p)
        st, z = synth.cart_inputs(100003, seed=9)
        z = z * 1.3
        a_o = orc.complete_partial(st, z)
        ac_o, k = grad_steps(orc, st, a_o, 2e-2, 0.0, 1e-5, 10)
        ac, a, flags = ops.cart_project(z.cuda(), p, 2e-2, 0.0, 1e-5, 10, want_a=True)
        assert (a.cpu() - a_o).abs().max() <= 1e-5
        assert (ac.cpu() - ac_o).abs().max() <= 1e-4


def test_pen_golden():
    from rpo_b200 import ops
    g = np.load(os.path.join(GOLD, "pen.npz"))
    obs = torch.tensor(g["obs"]).cuda()
    z = torch.tensor(g["z"]).cuda()
    a = ops.pen_complete_fwd(obs, z)
    assert rel_err(a.cpu(), torch.tensor(g["a"])).max() <= 1e-5
    dz = ops.pen_complete_bwd(obs, torch.tensor(g["g"]).cuda())
    assert rel_err(dz.cpu(), torch.tensor(g["dz"])).max() <= 1e-5
    ag = torch.tensor(g["a"]).cuda()
    eq, iq = ops.pen_resid(obs, ag)
    assert np.abs(eq.cpu().numpy() - g["eq_resid"]).max() <= 1e-5
    assert np.abs(iq.cpu().numpy() - g["ineq_resid"]).max() <= 1e-4
    d = ops.pen_direction(obs, ag)
    assert rel_err(d.cpu(), torch.tensor(g["dir"])).max() <= 1e-5          # row-wise semantics (reference at B == 1)
    # The golden rows were produced one at a time (B == 1), so each row stopped on ITS OWN loop test after
    # eval_steps[i] steps; the kernel runs n_steps on every row (global-exit semantics), so rows are grouped by count.
    ks = g["eval_steps"]
    for k in np.unique(ks):
        sel = torch.tensor(np.nonzero(ks == k)[0]).cuda()
        ac, fl = ops.pen_correct(obs[sel], ag[sel], 2e-3, 0.0, 1e-5, int(min(k, 10)))
        assert rel_err(ac.cpu(), torch.tensor(g["a_corr"])[sel.cpu()]).max() <= 1e-4
        ace, fl = ops.pen_correct(obs[sel], ag[sel], 2e-3, 0.0, 1e-5, int(k))
        assert rel_err(ace.cpu(), torch.tensor(g["a_corr_eval"])[sel.cpu()]).max() <= 1e-4
        ap, a1, _ = ops.pen_project(obs[sel], z[sel], 2e-3, 0.0, 1e-5, int(min(k, 10)), want_a=True)
        assert torch.equal(a1, a[sel])
        assert rel_err(ap.cpu(), torch.tensor(g["a_corr"])[sel.cpu()]).max() <= 1e-4
    # per-row loop test reproduced from the flags when a single row is passed
    for i in range(8):
        _, fl = ops.pen_correct(obs[i:i + 1], ag[i:i + 1], 2e-3, 0.0, 1e-5, 50)
        fl = fl.cpu().numpy()
        assert next((j for j in range(1, 50) if fl[j] == 0), 50) == int(ks[i])
    nst, nobs, rew, done, eqv, iqv = ops.pen_step(torch.tensor(g["step_state"], dtype=torch.float32).cuda(),
                                                  torch.tensor(g["step_action"]).cuda())
    assert np.abs(nst.cpu().numpy() - g["step_next"]).max() <= 1e-5
    assert np.abs(nobs.cpu().numpy() - g["step_obs"]).max() <= 1e-5
    assert np.abs(rew.cpu().numpy() - g["step_reward"]).max() <= 1e-5
    assert (done.cpu().numpy().astype(bool) == g["step_done"]).all()
    assert np.abs(iqv.cpu().numpy().ravel() - g["step_ineq_viol"].ravel()).max() <= 1e-4
    assert np.abs(eqv.cpu().numpy().ravel() - g["step_eq_viol"].ravel()).max() <= 1e-4


def test_pen_oracle_large():
    from oracle.rpo_oracle import PenOracle
    from rpo_b200 import ops, synth
    orc = PenOracle()
    obs, z = synth.pen_inputs(100003, seed=12)
    z = z * 1.5
    a_o = orc.complete_partial(obs, z)
    y = a_o
    for _ in range(10):
        y = y - 2e-3 * orc.ineq_partial_grad(obs, y)
    ac, a, _ = ops.pen_project(obs.cuda(), z.cuda(), 2e-3, 0.0, 1e-5, 10, want_a=True)
    assert rel_err(a.cpu(), a_o).max() <= 1e-5
    assert rel_err(ac.cpu(), y).max() <= 1e-4


def test_loop_without_flags_equals_loop_with_flags():
    """Callers that do not need the reference's loop-exit count pass no flag buffer; the kernels then skip the loop test
    (separate loop in the kernel): the corrected actions must be bit-equal either way, also beyond 64 steps."""
    from rpo_b200 import ops, synth
    B = 70001
    _, z = synth.cart_inputs(B, seed=7, device="cuda")
    z = z * 1.3
    obs, zp = synth.pen_inputs(B, seed=8, device="cuda")
    zp = zp * 1.5
    for n_steps, mom in ((10, 0.0), (50, 0.4), (70, 0.0)):
        a1, _, f1 = ops.cart_project(z, 1, 2e-2, mom, 1e-5, n_steps)
        a0, _, f0 = ops.cart_project(z, 1, 2e-2, mom, 1e-5, n_steps, want_flags=False)
        assert torch.equal(a0, a1) and f0 is None and f1.shape[0] == n_steps
        p1, _, g1 = ops.pen_project(obs, zp, 2e-3, mom, 1e-5, n_steps)
        p0, _, g0 = ops.pen_project(obs, zp, 2e-3, mom, 1e-5, n_steps, want_flags=False)
        assert torch.equal(p0, p1) and g0 is None
        # the flags are monotone: once no instance fails the test any more, none fails later (no step is taken)
        if mom == 0.0:
            for f in (f1.tolist(), g1.tolist()):
                assert all(a >= b for a, b in zip(f, f[1:])), f
        c1, h1 = ops.cart_correct(a1, 1, 2e-2, mom, 1e-5, n_steps)
        c0, _ = ops.cart_correct(a1, 1, 2e-2, mom, 1e-5, n_steps, want_flags=False)
        assert torch.equal(c0, c1)
