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


def test_multi_step_trajectories_of_the_scalar_gym_api_follow_the_reference(monkeypatch):
    """the drop-in `step()` keeps its persistent state through the kernels' float32 I/O (the reference keeps
    python floats / float64).  Over whole episodes with a fixed action sequence the drift this causes stays at float32
    round-off of the state: the drop-in CartSafeEnv / SpringPendulumEnv / EVOPFEnv trajectories track the unmodified
    reference's step for step, and the episodes end on the same step."""
    import os
    from oracle.ref_loader import DATA_DIR, make_cpu_env
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)
    from rpo_b200.env import CartSafeEnv, EVOPFEnv, SpringPendulumEnv
    rng = np.random.default_rng(11)
    # cart: stabilising-ish random forces for up to 200 steps
    ref, env = make_cpu_env("cart"), CartSafeEnv(device="cpu", partial_idx=0)
    s0 = rng.uniform(-0.05, 0.05, size=6)
    ref.reset(); env.reset()
    ref.state, env.state = tuple(s0), tuple(s0)
    ref.steps_beyond_done = env.steps_beyond_done = None
    worst = 0.0
    for t in range(200):
        a = rng.uniform(-3, 3, size=2).astype(np.float32)
        o_r, r_r, d_r, _ = ref.step(a.copy())
        o_b, r_b, d_b, _ = env.step(a.copy())
        worst = max(worst, float(np.abs(np.asarray(o_b) - np.asarray(o_r)).max() / max(1.0, np.abs(o_r).max())))
        assert d_r == d_b and r_r == r_b, t
        if d_r:
            break
    assert t >= 5 and worst < 2e-5, (t, worst)
    # pendulum
    ref, env = make_cpu_env("pen"), SpringPendulumEnv(device="cpu")
    ref.reset(); env.reset()
    st0 = np.array([0.05, -0.2, 1.01, 0.01])
    ref.state, env.state = st0.copy(), st0.copy()
    worst = 0.0
    for t in range(200):
        a = rng.uniform(-2, 2, size=2).astype(np.float32) + np.array([0.0, 5.0], np.float32)     # roughly holds the weight
        o_r, r_r, d_r, _ = ref.step(a.copy())
        o_b, r_b, d_b, _ = env.step(a.copy())
        worst = max(worst, float(np.abs(np.asarray(o_b) - np.asarray(o_r)).max()))
        assert d_r == d_b, t
        assert abs(r_r - r_b) <= 1e-3 * max(1.0, abs(r_r))
        if d_r:
            break
    assert t >= 5 and worst < 5e-5, (t, worst)
    # EVOPF: a whole 24-step day, state of charge accumulating through the battery efficiency
    np.random.seed(5)
    ref = make_cpu_env("evopf")
    np.random.seed(5)
    env = EVOPFEnv(device="cpu", data_dir=DATA_DIR)
    np.random.seed(6); s_r = ref.reset()
    np.random.seed(6); s_b = env.reset()
    assert np.abs(s_r - s_b).max() < 1e-12
    for t in range(24):
        a = ref.action_space.sample().astype(np.float32)
        a[38:] = rng.uniform(-0.25, 0.25, size=5)                         # battery power beyond its box: clipped by Battery.step
        ns_r, r_r, d_r, i_r = ref.step(a.copy())
        ns_b, r_b, d_b, i_b = env.step(a.copy())
        assert d_r == d_b
        assert np.abs(ns_b - ns_r).max() < 2e-6, (t, np.abs(ns_b - ns_r).max())       # soc drift stays at float32 round-off
        assert abs(r_r - r_b) <= 2e-5 * max(1.0, abs(r_r))
    assert d_r
