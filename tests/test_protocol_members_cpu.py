"""The small protocol members outside the hot path (penalty gradients of the cart, numpy helpers, hold / release switches,
EVOPF obj_fn / decompose / partial_unknown_vars) against the UNMODIFIED reference classes (build container only)."""
import numpy as np
import pytest
import torch

from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def test_cart_members(monkeypatch):
    from oracle.ref_loader import make_cpu_env
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)
    from rpo_b200.env import CartSafeEnv
    ref = make_cpu_env("cart")
    env = CartSafeEnv(device="cpu", partial_idx=int(ref.partial_actions[0]))
    g = torch.Generator().manual_seed(3)
    a = (torch.rand(32, 2, generator=g) - 0.5) * 30
    st = torch.zeros(32, 6)
    assert torch.allclose(env.eq_grad(st, a), ref.eq_grad(st, a), rtol=1e-6, atol=1e-6)
    assert torch.allclose(env.ineq_grad(st, a), ref.ineq_grad(st, a), rtol=1e-6, atol=1e-5)
    a1 = a[5].numpy()
    np.testing.assert_allclose(env.ineq_resid_np(None, a1), ref.ineq_resid_np(None, a1), rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(env.eq_dist_np(None, a1), np.ravel(ref.eq_dist_np(None, a1)), rtol=1e-6, atol=1e-6)
    for name in ("hold_eq", "release_eq", "hold_ineq", "release_ineq"):
        getattr(env, name)()
        getattr(ref, name)()
    assert env.holding_eq == ref.holding_eq and env.holding_ineq == ref.holding_ineq


def test_pendulum_switches():
    from rpo_b200.env import SpringPendulumEnv
    env = SpringPendulumEnv(device="cpu")
    env.hold_eq(); env.hold_ineq(); env.release_eq(); env.release_ineq()
    assert env.holding_eq is False and env.holding_ineq is False
    assert env.set_eq(None) is None and env.set_ineq(None, None) is None


def test_evopf_members():
    from oracle.ref_loader import make_cpu_env
    from rpo_b200 import synth
    from rpo_b200.env import EVOPFEnv
    ref = make_cpu_env("evopf")
    env = EVOPFEnv(device="cpu")
    assert np.array_equal(env.partial_unknown_vars, ref.partial_unknown_vars)
    s = np.arange(57.0)
    for a, b in zip(env.decompose(s), ref.decompose(s)):
        assert np.array_equal(a, b)
    st = synth.evopf_states(16, seed=1)
    y = torch.rand(16, 43, generator=torch.Generator().manual_seed(2))
    mine, theirs = env.obj_fn(y), ref.obj_fn(y.numpy())
    assert mine.dtype == theirs.dtype == torch.float64
    assert torch.allclose(mine, theirs, rtol=1e-12, atol=0)


def test_evopf_ineq_jac_and_grad(monkeypatch):
    from oracle.ref_loader import make_cpu_env
    from rpo_b200 import synth
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)
    from rpo_b200.env import EVOPFEnv
    ref = make_cpu_env("evopf")
    env = EVOPFEnv(device="cpu")
    st = synth.evopf_states(24, seed=4)
    lo, hi = env.box_constraint
    y = torch.from_numpy(lo + (hi - lo) * np.random.default_rng(0).uniform(-0.2, 1.2, size=(24, 43))).float()
    assert torch.equal(env.ineq_jac(st, y), ref.ineq_jac(st, y))
    for eps in (0.0, 0.02):
        assert torch.allclose(env.ineq_grad(st, y, eps=eps), ref.ineq_grad(st, y, eps=eps), rtol=1e-6, atol=1e-6)
    assert torch.equal(env.ineq_grad_new(st, y), ref.ineq_grad_new(st, y))
    y[:, 24:38] = (torch.rand(24, 14, generator=torch.Generator().manual_seed(5)) - 0.5) * 0.6        # angles in radians
    J, Jr = env.eq_jac(y), ref.eq_jac(y)
    assert J.shape == (24, 28, 43) and float((J - Jr).abs().max()) <= 1e-6 * float(Jr.abs().max())
    assert torch.allclose(env.eq_grad(st, y), ref.eq_grad(st, y), rtol=1e-4, atol=1e-3)
    from oracle.rpo_oracle import EvopfOracle
    J64 = env.eq_jac(y.double())                             # float64 in, float64 out (the reference itself is float32 only)
    assert J64.dtype == torch.float64
    assert float((J64 - EvopfOracle(dtype=torch.float64, const_dtype=torch.float32).eq_jac(y.double())).abs().max()) < 1e-5
