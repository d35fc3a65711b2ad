"""The drop-in boundary on the GPU: rpo_b200.env classes + rpo_b200.algo drivers used the way rpo.algo uses them
(rpo/algo/rpo_ddpg.py:72-77, 96-118, 266-337)."""
import copy
import os

import numpy as np
import pytest
import torch

from tests.parity import GRAD_RTOL, Y_RTOL, assert_close_or_marginal, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class _Algo(object):
    """What RPODDPG carries for the projection (rpo_ddpg.py:35-47) with the B200 drivers mixed in."""

    def __init__(self, env, lr, max_steps=10, eval_steps=50, momentum=0.0):
        self.env, self.corr_lr, self.eval_lr = env, lr, lr
        self.max_steps, self.eval_steps = max_steps, eval_steps
        self.corr_eps, self.corr_momentum, self.corr_mode, self.grad_eps = 1e-5, momentum, 0, 0.02


def _mk(env, *a, **k):
    from rpo_b200.algo import RPOProjectionMixin

    class A(_Algo, RPOProjectionMixin):
        pass
    return A(env, *a, **k)


def test_evopf_env_protocol_and_autograd(evopf_oracle):
    from rpo_b200 import synth
    from rpo_b200.env import EVOPFEnv
    orc = evopf_oracle
    env = EVOPFEnv()
    assert (env.state_dim, env.action_dim, env.eq_num, env.ineq_num) == (57, 43, 28, 58)
    assert list(env.partial_actions) == list(orc.partial_actions)
    copy.deepcopy(env)                                            # rpo_ddpg.py:61
    B = 300
    st = synth.evopf_states(B, seed=40)
    z = synth.evopf_partial_actions(st, seed=41)
    g = synth.upstream_grad(B, 43, seed=42)
    s_d = st.cuda().requires_grad_(True)
    z_d = z.cuda().requires_grad_(True)
    y = env.complete_partial(s_d, z_d)
    (y * g.cuda()).sum().backward()
    y_o, jac, inv, _ = orc.complete_partial_fwd(st, z, return_saved=True)
    dx_o, dz_o = orc.complete_partial_bwd(g, jac, inv)
    assert rel_err(y, y_o).max() <= Y_RTOL
    assert rel_err(z_d.grad, dz_o).max() <= GRAD_RTOL
    assert rel_err(s_d.grad[:, :28], dx_o).max() <= GRAD_RTOL
    assert float(s_d.grad[:, 28:].abs().max()) == 0.0
    yd = y.detach()
    assert (env.eq_resid(s_d.detach(), yd).cpu() - orc.eq_resid(st, yd.cpu())).abs().max() < 2e-5
    assert (env.ineq_dist(s_d.detach(), yd).cpu() - orc.ineq_dist(st, yd.cpu())).abs().max() < 1e-6
    assert torch.equal(env.ineq_grad_new(s_d.detach(), yd).cpu(), orc.ineq_grad_new(st, yd.cpu()))
    lo, hi = env.update(st)
    lo_o, hi_o = orc.partial_box(st)
    assert np.array_equal(lo, lo_o) and np.array_equal(hi, hi_o)
    lo_d, hi_d = env.update_device(st.cuda())
    assert np.abs(lo_d.cpu().numpy() - lo_o).max() == 0 and np.abs(hi_d.cpu().numpy() - hi_o).max() == 0
    jac = env.eq_jac(yd)                                     # dense diagnostic form (torch), not a kernel
    assert jac.shape == (yd.shape[0], 28, 43)
    assert (jac.cpu() - orc.eq_jac(yd.cpu())).abs().max() < 2e-4          # entries up to ~48: 4e-6 relative


def test_process_action_train_and_eval(evopf_oracle):
    from oracle.rpo_oracle import grad_steps_margins
    from rpo_b200 import synth
    from rpo_b200.env import EVOPFEnv
    orc = evopf_oracle
    env = EVOPFEnv()
    algo = _mk(env, 1e-4)
    B = 256                                                      # replay batch of the scripts (scripts/evopf_exp.py:29)
    st = synth.evopf_states(B, seed=50)
    z = synth.evopf_partial_actions(st, seed=51)
    y_o = orc.complete_partial_fwd(st, z)
    with torch.no_grad():
        a = algo.process_action(st.cuda(), z.cuda())             # fused path
    yc_o, k, margin = grad_steps_margins(orc, st, y_o, 1e-4, 0.0, 1e-5, 10)
    assert_close_or_marginal(rel_err(a, yc_o), margin, Y_RTOL)
    # with autograd enabled the same call goes through complete_partial (+ autograd) and grad_steps
    zg = z.cuda().requires_grad_(True)
    a2 = algo.process_action(st.cuda(), zg)
    assert torch.equal(a2.detach(), a)
    # rollout shape: B == 1, eval branch returns (action, steps) (rpo_ddpg.py:233,288-305)
    with torch.no_grad():
        a1, steps = algo.process_action(st[:1].cuda(), z[:1].cuda(), train=False)
    yc1_o, k1, m1 = grad_steps_margins(orc, st[:1], y_o[:1], 1e-4, 0.0, 1e-5, 50)
    assert steps == k1
    assert_close_or_marginal(rel_err(a1, yc1_o), m1, Y_RTOL)
    # 1-D partial action (complete_partial reshapes it, evopf.py:725-726)
    with torch.no_grad():
        a1b = algo.process_action(st[:1].cuda(), z[0].cuda())
    assert a1b.shape == (1, 43)


def test_cart_eval_early_exit_matches_reference_golden():
    from rpo_b200.env import CartSafeEnv
    for p in (0, 1):
        g = np.load(os.path.join(GOLD, "cart_p%d.npz" % p))
        env = CartSafeEnv(partial_idx=p)
        algo = _mk(env, 2e-2)
        st = torch.tensor(g["state"]).cuda()
        z = torch.tensor(g["z"]).cuda()
        with torch.no_grad():
            a, k = algo.process_action(st, z, train=False)
            at = algo.process_action(st, z, train=True)
        assert k == int(g["eval_steps"])
        assert np.abs(a.cpu().numpy() - g["a_corr_eval"]).max() <= 1e-4
        assert np.abs(at.cpu().numpy() - g["a_corr"]).max() <= 1e-4
        zg = z.clone().requires_grad_(True)
        out = env.complete_partial(st, zg)
        out.backward(torch.tensor(g["g"]).cuda())
        assert np.abs(zg.grad.cpu().numpy() - g["dz"]).max() <= 1e-6


def test_pen_env_autograd_and_rollout():
    from rpo_b200.env import SpringPendulumEnv
    g = np.load(os.path.join(GOLD, "pen.npz"))
    env = SpringPendulumEnv()
    obs = torch.tensor(g["obs"]).cuda()
    zg = torch.tensor(g["z"]).cuda().requires_grad_(True)
    a = env.complete_partial(obs, zg)
    a.backward(torch.tensor(g["g"]).cuda())
    assert rel_err(zg.grad, torch.tensor(g["dz"])).max() <= 1e-5
    algo = _mk(env, 2e-3)
    for i in range(6):                                            # B == 1 rollouts: per-row loop exit as the reference
        with torch.no_grad():
            ai, k = algo.process_action(obs[i:i + 1], zg.detach()[i:i + 1], train=False)
        assert k == int(g["eval_steps"][i])
        assert rel_err(ai, torch.tensor(g["a_corr_eval"][i:i + 1])).max() <= 1e-4
    np.random.seed(0)
    o = env.reset()
    for t in range(5):
        o, r, done, info = env.step(np.array([0.5, 5.2], dtype=np.float32))
        assert o.shape == (5,) and info["ineq_viol"].shape == (1, 1) and isinstance(done, bool)


def test_evopf_gym_step_matches_reference_golden():
    from rpo_b200.env import EVOPFEnv
    g = np.load(os.path.join(GOLD, "evopf_step.npz"))
    env = EVOPFEnv()
    env.reset()
    for t in range(g["state"].shape[0]):
        env.state = g["state"][t].copy()
        nxt = g["next_state"][t]
        # feed the loaders' next rows of the golden trajectory (exogenous inputs of step)
        env.data.cache = np.tile(nxt[:28], (25, 1)); env.data.counter = 0
        env.p_data.cache = np.tile(nxt[33:] * 100.0, 1).repeat(1); env.p_data.cache = np.concatenate([nxt[33:] * 100.0, np.zeros(24)])
        env.p_data.counter = 0
        ns, r, done, info = env.step(g["action"][t])
        assert np.abs(ns - nxt).max() <= 2e-6 * max(1.0, np.abs(nxt).max())
        assert abs(r - float(g["reward"][t])) <= 1e-5 * max(1.0, abs(float(g["reward"][t])))
        assert np.abs(info["ineq_viol"][0] - g["ineq_viol"][t]).max() <= 1e-6
        assert np.abs(info["eq_viol"][0] - g["eq_viol"][t]).max() <= 2e-5


def test_actor_critic_iteration_runs():
    """One DDPG-style iteration through the boundary: critic target via process_action (no grad, fused kernel),
    actor loss through complete_partial + GRG backward (rpo_ddpg.py:307-337)."""
    from rpo_b200 import synth
    from rpo_b200.env import EVOPFEnv
    torch.manual_seed(0)
    env = EVOPFEnv()
    algo = _mk(env, 1e-4)
    B = 256
    st = synth.evopf_states(B, seed=60).cuda()
    actor = torch.nn.Sequential(torch.nn.Linear(57, 64), torch.nn.ReLU(), torch.nn.Linear(64, 14), torch.nn.Tanh()).cuda()
    critic = torch.nn.Sequential(torch.nn.Linear(57 + 43, 64), torch.nn.ReLU(), torch.nn.Linear(64, 1)).cuda()
    lo, hi = env.update_device(st)

    def policy(s):
        return lo + (hi - lo) * (0.5 + 0.3 * actor(s))
    with torch.no_grad():
        nxt = algo.process_action(st, policy(st))
        q_t = critic(torch.cat([st, nxt], 1))
    a = env.complete_partial(st, policy(st))
    loss = (-critic(torch.cat([st, a], 1)) + env.ineq_dist(st, a).sum(1, keepdim=True) * 0.1).mean()
    loss.backward()
    gsum = sum(float(p.grad.abs().sum()) for p in actor.parameters())
    assert np.isfinite(gsum) and gsum > 0 and bool(torch.isfinite(q_t).all())


def test_host_projector_matches_resident_path():
    """Host-buffer pipeline (pinned in / pinned out, chunked over streams, strided state upload) == resident kernels."""
    from rpo_b200 import ops, synth
    from rpo_b200.host_api import HostProjector
    B = 5000                                           # not a multiple of the chunk
    st = synth.evopf_states(B, seed=70)
    z = synth.evopf_partial_actions(st, seed=71)
    g = synth.upstream_grad(B, 43, seed=72)
    hp = HostProjector("cuda:0", chunk=1024, n_streams=3)
    yc_h, dz_h = hp.pinned(B, 43), hp.pinned(B, 14)
    hp.project_backward(st.pin_memory(), z.pin_memory(), g.pin_memory(), yc_h, dz_h)
    yc, y, _, _ = ops.evopf_project(st.cuda(), z.cuda(), 1e-4, 0.0, 1e-5, 10, want_y=True)
    dz, _ = ops.evopf_complete_bwd(y, g.cuda(), need_dx=False)
    assert torch.equal(yc_h, yc.cpu()) and torch.equal(dz_h, dz.cpu())
    assert hp.h2d_bytes == B * (57 + 14 + 43) * 4 and hp.d2h_bytes == B * (43 + 14) * 4
    hp.project_backward(st[:, :33].contiguous().pin_memory(), z.pin_memory(), None, yc_h, None)   # compact state, forward only
    assert torch.equal(yc_h, yc.cpu())
    # the strided-DMA helpers of the C ABI (kept for callers that want them)
    from rpo_b200 import _lib
    lib = _lib.load()
    d = torch.zeros(B, 33, device="cuda")
    sp = st.pin_memory()
    _lib.check(lib.rpo_copy_rows_h2d(d.data_ptr(), 33, sp.data_ptr(), 57, 33, B, torch.cuda.current_stream().cuda_stream), "h2d")
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), st[:, :33])


def test_vector_env_and_replay():
    """f-2 / f-4: N replicas stepped by the batched step kernel, device replay, device exogenous sampler."""
    from rpo_b200 import ops, synth
    from rpo_b200.vector import DeviceReplay, EvopfVectorEnv
    env = EvopfVectorEnv(512, "cuda:0", seed=3)
    s = env.reset()
    assert s.shape == (512, 57) and float(s[:, 0].abs().max()) == 0.0 and bool((s[:, 28:33] == 0.2).all())
    assert 2.0 < float(s[:, :14].sum(1).mean()) < 3.2
    rb = DeviceReplay(2048, "cuda:0", state=57, action=43, next_state=57, reward=1, done=1)
    for t in range(30):
        z = synth.evopf_partial_actions(s, seed=t)
        a, _, _, _ = ops.evopf_project(s, z, 1e-4, 0.0, 1e-5, 10)
        ns, r, done, eqv, iqv = env.step(a)
        rb.extend(state=s, action=a, next_state=ns, reward=r, done=done.float())
        assert bool(done.all()) == (t % 24 == 23)
        if t % 24 == 23:
            assert bool((ns[:, 28:33] == 0.2).all())
        s = ns
    assert len(rb) == 2048
    b = rb.sample(256)
    assert b["state"].shape == (256, 57) and b["reward"].shape == (256, 1) and bool(torch.isfinite(b["reward"]).all())
