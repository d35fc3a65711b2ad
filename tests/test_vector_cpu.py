"""rpo_b200.vector (vectorised replicas, device replay, exogenous sampler, batched eval) on CPU through the host
simulator ops (test scope only; the GPU twin is tests/test_env_api_gpu.py::test_vector_env_and_replay)."""
import numpy as np
import torch


class _Proj(object):
    pass


def test_vector_env_replay_and_batched_eval(monkeypatch):
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)
    from rpo_b200.algo import RPOProjectionMixin
    from rpo_b200.env import EVOPFEnv
    from rpo_b200.vector import DeviceReplay, EvopfExoSampler, EvopfVectorEnv, evaluate_batched
    smp = EvopfExoSampler("cpu", seed=1)
    rows = smp.rows(torch.arange(0, 24).repeat(50))
    pd, qd, price = rows[:, :14], rows[:, 14:28], rows[:, 28:]
    assert rows.shape == (1200, 52) and bool((pd[:, [0, 6, 7]] == 0).all()) and bool((pd >= 0).all())
    tot = pd.sum(1)
    assert 2.0 < float(tot.min()) and float(tot.max()) < 3.2           # 2.59 pu x daily curve (demand.py:47-52)
    assert float((qd.abs() / pd.clamp_min(1e-9))[pd > 0].max()) <= float(np.tan(np.arccos(0.9))) + 1e-5
    assert bool((price[0::24, :] != 0).all()) and bool((price[23::24, 1:] == 0).all())   # window runs past the day end

    env = EVOPFEnv(device="cpu")

    class P(_Proj, RPOProjectionMixin):
        pass
    proj = P()
    proj.env, proj.corr_lr, proj.eval_lr, proj.max_steps, proj.eval_steps = env, 1e-4, 1e-4, 10, 50
    proj.corr_eps, proj.corr_momentum, proj.corr_mode = 1e-5, 0.0, 0
    venv = EvopfVectorEnv(32, "cpu", seed=2)
    s = venv.reset()
    rb = DeviceReplay(100, "cpu", state=57, action=43, reward=1)

    def policy(st):
        lo, hi = env.update_device(st)
        return lo + (hi - lo) * 0.5
    for t in range(5):
        a = proj.process_action(s, policy(s))
        ns, r, done, eqv, iqv = venv.step(a)
        rb.extend(state=s, action=a, reward=r)
        assert ns.shape == (32, 57) and not bool(done.any())
        # soc moves by the clipped, efficiency-scaled battery power (Battery.step evopf.py:86-96)
        pe = a[:, 38:]
        hi = (1 / 0.9) * torch.clamp(0.8 - s[:, 28:33].double(), max=0.2)
        lo = 0.9 * torch.clamp(0.1 - s[:, 28:33].double(), min=-0.2)
        pc = torch.minimum(torch.maximum(pe.double(), lo), hi)
        pc = torch.where(pc >= 0, pc * 0.9, pc / 0.9)
        assert float((ns[:, 28:33].double() - (s[:, 28:33].double() + pc)).abs().max()) < 1e-6
        s = ns
    assert len(rb) == 100 and rb.sample(8)["action"].shape == (8, 43)
    stats = evaluate_batched(EvopfVectorEnv(8, "cpu", seed=3), policy, proj, max_len=30)
    assert len(stats) == 10 and all(np.isfinite(stats))
    assert stats[0] < 0                                                   # generation costs money: negative reward
