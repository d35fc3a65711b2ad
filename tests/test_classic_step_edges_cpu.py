"""Cart / pendulum env steps against the UNMODIFIED reference on edge cases: forces beyond the action box (clipped before
the dynamics, cartpole.py:170-173), states next to the termination thresholds, friction sign changes (x_dot around 0),
pendulum length / angle limits.  Step code = the kernels' per-instance code compiled by g++ (tests/hostsim)."""
import numpy as np
import pytest
import torch

from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def test_cart_step_edges():
    from oracle.ref_loader import make_cpu_env
    from tests import hostsim_ops as hs
    ref = make_cpu_env("cart")
    rng = np.random.default_rng(3)
    N = 80
    s = rng.uniform(-0.05, 0.05, size=(N, 6))
    s[:, 0] = rng.uniform(-2.45, 2.45, size=N)                       # cart position around the +-2.4 threshold
    s[:, 3] = rng.uniform(-0.22, 0.22, size=N)                       # pole angle around +-12 degrees (0.2094 rad)
    s[:10, 1] = np.array([0.0, 1e-9, -1e-9, 1e-3, -1e-3, 0.5, -0.5, 0.0, 0.0, 2.0])   # friction sign(n_c * x_dot)
    s[:, 4] = rng.uniform(-1.5, 1.5, size=N)
    a = rng.uniform(-14, 14, size=(N, 2)).astype(np.float32)          # beyond the +-10 action box
    a[0] = [10.0, -10.0]
    a[1] = [10.000001, -10.000001]
    ns, done, eqv, iqv = hs.cart_step(torch.from_numpy(s).float(), torch.from_numpy(a))
    s32 = s.astype(np.float32).astype(np.float64)                    # the kernel reads float32 states
    for i in range(N):
        ref.reset()
        ref.state = tuple(s32[i])
        ref.steps_beyond_done = None
        ns_r, _, done_r, info = ref.step(a[i].copy())
        np.testing.assert_allclose(ns[i].numpy(), ns_r, rtol=2e-6, atol=2e-6, err_msg=str(i))
        margin = min(abs(abs(ns_r[0]) - 2.4), abs(abs(ns_r[3]) - 12 * 2 * np.pi / 360))
        if margin > 1e-6:                                            # float32 output vs float64 threshold test
            assert bool(done[i]) == done_r, i
        np.testing.assert_allclose(iqv[i].numpy(), np.ravel(info["ineq_viol"]), rtol=1e-6, atol=2e-6)
        np.testing.assert_allclose(eqv[i].numpy(), np.ravel(info["eq_viol"]), rtol=1e-6, atol=2e-6)


def test_pendulum_step_edges():
    from oracle.ref_loader import make_cpu_env
    from tests import hostsim_ops as hs
    ref = make_cpu_env("pen")
    rng = np.random.default_rng(4)
    N = 80
    st = np.stack([rng.uniform(-0.27, 0.27, size=N),                 # angle around +-pi/12 = 0.2618
                   rng.uniform(-1.5, 1.5, size=N),
                   rng.uniform(0.48, 1.52, size=N),                  # length around the (0.5, 1.5) limits
                   rng.uniform(-0.08, 0.08, size=N)], axis=1)
    a = rng.uniform(-9, 9, size=(N, 2)).astype(np.float32)            # beyond the +-6 box and the |f|^2 <= 32 disc
    nst, nobs, rew, done, eqv, iqv = hs.pen_step(torch.from_numpy(st).float(), torch.from_numpy(a))
    st32 = st.astype(np.float32).astype(np.float64)
    for i in range(N):
        ref.reset()
        ref.state = st32[i].copy()
        obs_r, r_r, done_r, info = ref.step(a[i].copy())
        np.testing.assert_allclose(nst[i].numpy(), ref.state, rtol=3e-6, atol=3e-6, err_msg=str(i))
        np.testing.assert_allclose(nobs[i].numpy(), obs_r, rtol=3e-6, atol=3e-6)
        assert abs(float(rew[i]) - r_r) <= 2e-4 * max(1.0, abs(r_r))          # 1 / (100 |theta| + 1) amplifies theta's rounding
        l, th = ref.state[2], ref.state[0]
        margin = min(abs(l - 0.5), abs(l - 1.5), abs(abs(th) - np.pi / 12))
        if margin > 1e-6:
            assert bool(done[i]) == done_r, i
        np.testing.assert_allclose(iqv[i].numpy(), np.ravel(info["ineq_viol"]), rtol=2e-6, atol=1e-5)
        np.testing.assert_allclose(eqv[i].numpy(), np.ravel(info["eq_viol"]), rtol=2e-5, atol=2e-5)
