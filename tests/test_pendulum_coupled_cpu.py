"""Opt-in `reference_batch_coupling` of the SpringPendulum drop-in: the reference's BATCHED direction / grad_steps couple
the rows through a (B,1)@(1,B) product (pendulum.py:337-339).  The default path is row-wise; this mode
reproduces the batched reference result for result (tests/golden/pen_batched.npz = output of the unmodified reference)."""
import os

import numpy as np
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pen.npz"))
GB = np.load(os.path.join(os.path.dirname(__file__), "golden", "pen_batched.npz"))


def _algo(env, lr, steps, momentum=0.0, eval_steps=50):
    from rpo_b200.algo import RPOProjectionMixin

    class A(RPOProjectionMixin):
        pass
    a = A()
    a.env, a.corr_lr, a.eval_lr, a.max_steps, a.eval_steps = env, lr, lr, steps, eval_steps
    a.corr_eps, a.corr_momentum, a.corr_mode = 1e-5, momentum, 0
    return a


def test_coupled_mode_reproduces_the_batched_reference(monkeypatch):
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)
    from rpo_b200.env import SpringPendulumEnv
    env = SpringPendulumEnv(device="cpu", reference_batch_coupling=True)
    obs, a = torch.from_numpy(G["obs"]), torch.from_numpy(G["a"])
    d = env.ineq_partial_grad(obs, a)
    np.testing.assert_allclose(d.numpy(), GB["dir"], rtol=1e-6, atol=1e-6)
    # small row blocks give the same result as one (B,B) product
    np.testing.assert_allclose(env._direction_coupled(obs, a, rows_per_block=7).numpy(), GB["dir"], rtol=1e-6, atol=1e-6)
    ac = _algo(env, 2e-3, 10).grad_steps(obs, a, train=True)
    np.testing.assert_allclose(ac.numpy(), GB["a_corr"], rtol=1e-5, atol=1e-5)
    ace, k = _algo(env, 2e-3, 10, momentum=0.5).grad_steps(obs, a, train=False)
    assert k == int(GB["eval_steps"])
    np.testing.assert_allclose(ace.numpy(), GB["a_corr_eval_momentum"], rtol=1e-4, atol=1e-4)
    # fused entry point: completion + coupled correction
    z = torch.from_numpy(G["z"])
    yc = _algo(env, 2e-3, 10).process_action(obs, z)
    np.testing.assert_allclose(yc.numpy(), GB["a_corr"], rtol=1e-5, atol=1e-5)


def test_default_is_row_wise_and_differs_from_the_batched_reference(monkeypatch):
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)
    from rpo_b200.env import SpringPendulumEnv
    env = SpringPendulumEnv(device="cpu")
    obs, a = torch.from_numpy(G["obs"]), torch.from_numpy(G["a"])
    d = env.ineq_partial_grad(obs, a)
    np.testing.assert_allclose(d.numpy(), G["dir"], rtol=1e-5, atol=1e-6)          # rows as batches of one
    assert np.abs(d.numpy() - GB["dir"]).max() > 1.0                               # the batched reference couples rows
    # a batch of one is the same in both modes
    env_c = SpringPendulumEnv(device="cpu", reference_batch_coupling=True)
    assert torch.equal(env_c.ineq_partial_grad(obs[:1], a[:1]), env.ineq_partial_grad(obs[:1], a[:1]))
