"""Shared bodies of the f-2 / f-4 parity tests (device replay, exogenous generators, batched eval) — run on CPU through
the host simulator (tests/test_rollout_parity_cpu.py) and on the GPU through the kernels (tests/test_rollout_parity_gpu.py)."""
import contextlib
import os
import warnings

import numpy as np
import torch


def check_exo_sampler_matches_reference_loaders(device):
    """EvopfExoSampler against the reference's DemandLoader / PriceLoader (data/demand.py:35-75, data/price.py:29-55):
    same daily curves, per-hour moments of the demand rows, and the per-episode price realisation (sliding window,
    constant magnitude, zero tail)."""
    from oracle import ref_loader
    from rpo_b200.vector import EvopfExoSampler, EvopfVectorEnv
    rpo = ref_loader.load_reference()
    from rpo.env.electrical_grid.data.demand import DemandLoader
    from rpo.env.electrical_grid.data.price import PriceLoader
    from pypower.api import case14
    np.random.seed(5)
    dl = DemandLoader(ppc=case14(), rho=0.5, q_rand=True, T=24, min_pf=0.9, max_pf=1.0, regularized=True)     # evopf.py:212
    pl = PriceLoader(T=24, ahead=24, regularized=True, reg_bias=0.1, reg_sigma=0.05)                          # evopf.py:216
    n_ep = 400
    ref_d = np.zeros((n_ep, 24, 28))
    ref_p = np.zeros((n_ep, 24, 24))
    for e in range(n_ep):
        row, _ = dl.reset()
        win, _ = pl.reset()
        for h in range(24):
            ref_d[e, h], ref_p[e, h] = row, win / 100.0
            if h < 23:
                row, _ = dl.fetch()
                win, _ = pl.fetch()
    n = 2048
    smp = EvopfExoSampler(device, seed=3, n=n)
    # the daily curves are the reference's (loaded from its pickles next to the importable package)
    assert np.allclose(smp.demand_curve.cpu().numpy(), dl.curve / dl.curve.mean(), rtol=1e-6)
    assert np.allclose(smp.price_curve.cpu().numpy(), pl.curve / 100.0, rtol=1e-6)
    venv = EvopfVectorEnv(n, device, seed=4)
    s = venv.reset()
    days = [s.cpu().numpy()]
    zero_a = torch.zeros(n, 43, device=device)
    zero_a[:, 10:24] = 1.0
    for h in range(1, 26):
        ns, *_ = venv.step(zero_a)
        days.append(ns.cpu().numpy())
    days = np.stack(days, 1)                                             # (n, 26, 57): hour 0..23 of day 1, hours 0,1 of day 2
    pd, qd, price = days[:, :, :14], days[:, :, 14:28], days[:, :, 33:]
    # demand: per-hour mean total and per-bus mean / std against the reference's draws
    for h in (0, 7, 13, 23):
        assert abs(pd[:, h].sum(1).mean() - ref_d[:, h, :14].sum(1).mean()) < 2e-3 * ref_d[:, h, :14].sum(1).mean()
        assert np.abs(pd[:, h].mean(0) - ref_d[:, h, :14].mean(0)).max() < 0.02
        assert np.abs(pd[:, h].std(0) - ref_d[:, h, :14].std(0)).max() < 0.02
        assert np.abs(np.abs(qd[:, h]).mean(0) - np.abs(ref_d[:, h, 14:]).mean(0)).max() < 0.02
    assert (np.sign(qd[:, 3]) == np.sign(ref_d[0, 3, 14:])[None]).all()
    # price: ONE realisation per episode — the look-ahead is the true future price, the tail past the day end is zero
    assert np.array_equal(price[:, 0, 1:], price[:, 1, :-1]) and np.array_equal(price[:, 10, 1:], price[:, 11, :-1])
    assert (price[:, 23, 1:] == 0).all() and (price[:, 5, 19:] == 0).all() and (price[:, 5, :19] != 0).any()
    assert not np.array_equal(price[:, 24, :5], price[:, 0, :5])         # day 2 is a new realisation
    assert np.array_equal(price[:, 24, 1:], price[:, 25, :-1])
    # moments of the first window against the reference (mean curve, std ~ curve * sqrt(1.01 * 2 - 1))
    m_ref, s_ref = ref_p[:, 0].mean(0), ref_p[:, 0].std(0)
    assert np.abs(price[:, 0].mean(0) - m_ref).max() < 0.12 * np.abs(m_ref).max()
    assert np.abs(price[:, 0].std(0) - s_ref).max() < 0.15 * s_ref.max()
    # mag_r is one scalar per episode: the ratio of two realisations' hour-h noise is not constant, but every window of
    # an episode shares the realisation, so the correlation between price[t][k] and price[t+k][0] is exactly 1
    assert np.array_equal(price[:, 0, 7], price[:, 7, 0])


def check_device_replay_matches_reference_buffer(device):
    """DeviceReplay against rpo.utils.buffer.ReplayBuffer (buffer.py:3-34): ring pointer, fill level, overwrite order,
    uniform sampling over the filled part."""
    from collections import namedtuple
    from oracle import ref_loader
    ref_loader.load_reference()
    from rpo.utils.buffer import ReplayBuffer
    from rpo_b200.vector import DeviceReplay
    Type = namedtuple("Type", ("shape", "dtype"))
    cap = 37
    rb = ReplayBuffer(cap, state=Type((5,), np.float32), reward=Type((1,), np.float32))
    dr = DeviceReplay(cap, device, state=5, reward=1)
    g = np.random.default_rng(0)
    t = 0
    for chunk in (1, 7, 16, 30, 3, 37, 12):                               # wraps around several times
        st, rw = g.standard_normal((chunk, 5)).astype(np.float32), g.standard_normal((chunk, 1)).astype(np.float32)
        for i in range(chunk):
            rb.add(state=st[i], reward=rw[i])                             # buffer.py:22-29
        dr.extend(state=torch.tensor(st, device=device), reward=torch.tensor(rw, device=device))
        t += chunk
        assert len(dr) == rb.size == min(t, cap) and int(dr.ptr) == rb.pointer
        assert np.array_equal(dr.buf["state"].cpu().numpy(), rb.buffer["state"])
        assert np.array_equal(dr.buf["reward"].cpu().numpy(), rb.buffer["reward"])
    # partially filled buffer: samples come from [0, size) only, all slots reachable, roughly uniform
    dr2 = DeviceReplay(64, device, x=1)
    dr2.extend(x=torch.arange(10, device=device, dtype=torch.float32).view(-1, 1))
    smp = dr2.sample(20000)["x"].view(-1).cpu().numpy()
    assert smp.min() == 0 and smp.max() == 9
    counts = np.bincount(smp.astype(int), minlength=10)
    assert counts.min() > 1700 and counts.max() < 2300                   # np.random.randint(0, size) is uniform (buffer.py:32)


def check_batched_eval_matches_reference_eval(device, tol=2e-3):
    """evaluate_batched (10 replicas in lock-step, ONE fused projection per step, K = 50 eval semantics) against the
    UNMODIFIED reference's RPODDPG.eval (rpo_ddpg.py:207-264: 10 episodes one after the other, B = 1) on identical
    exogenous data: the reference loaders' own draws are replayed through TableExo, the policy is the reference agent's
    deterministic actor with the reference BoxConstraint."""
    from oracle import ref_loader
    from rpo_b200.algo import RPOProjectionMixin
    from rpo_b200.box import BoxConstraint
    from rpo_b200.env import EVOPFEnv
    from rpo_b200.env.loaders import DemandLoader, PriceLoader
    from rpo_b200.case14 import case14
    from rpo_b200.vector import EvopfVectorEnv, TableExo, evaluate_batched
    rpo = ref_loader.load_reference()
    import importlib
    src = importlib.import_module("rpo.algo.rpo_ddpg")
    with open(src.__file__) as fh:
        code = compile(fh.read(), src.__file__, "exec")
    ns = dict(src.__dict__)
    from rpo.algo.model.utils import BoxConstraint as RefBox
    ns["BoxConstraint"] = RefBox
    exec(code, ns)
    RefDDPG = ns["RPODDPG"]                                              # pristine class (install() may have patched the module's)

    class _Log(object):
        def add(self, **kw):
            pass
    hp = dict(batch_size=32, max_steps=10, warmup=0, lr_dual=2e-2, corr_lr=1e-4, eps=0.0001, eps_start=0.0001, eps_epoch=20000,
              eval_lr=1e-4, eval_steps=50, grad_eps=0.02, corr_momentum=0.0, policy_fre=4, ex_action_dim=1, gamma=0.95,
              max_epochs=1, capacity=200, clip_thres=0.2, shared_param=False, value_type="cat", device=torch.device("cpu"))
    np.random.seed(21)
    torch.manual_seed(22)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):
            algo = RefDDPG(ref_loader.make_cpu_env("evopf"), "./test", name="t", logger=_Log(), **hp)
            np.random.seed(77)
            ref_stats = algo.eval()                                      # consumes numpy RNG only in the 10 env resets
    # replay the same draws: env.reset() = data.reset() then evs.reset() -> p_data.reset() (evopf.py:368-375, :54-72)
    np.random.seed(77)
    ppc = case14()
    dl = DemandLoader(ppc=ppc, rho=0.5, q_rand=True, T=24, min_pf=0.9, max_pf=1.0, regularized=True,
                      pickle_path=os.path.join(ref_loader.DATA_DIR, "demand.pickle"))
    pl = PriceLoader(T=24, ahead=24, regularized=True, reg_bias=0.1, reg_sigma=0.05,
                     pickle_path=os.path.join(ref_loader.DATA_DIR, "price.pickle"))
    table = np.zeros((10, 24, 52), np.float32)
    for e in range(10):
        row, _ = dl.reset()
        win, _ = pl.reset()
        for h in range(24):
            table[e, h, :28], table[e, h, 28:] = row, win / 100.0
            if h < 23:
                win, _ = pl.fetch()                                      # step(): price first, then demand (evopf.py:350-353)
                row, _ = dl.fetch()
    venv = EvopfVectorEnv(10, device, exo=TableExo(torch.tensor(table, device=device)))
    env = EVOPFEnv(device=device, data_dir=ref_loader.DATA_DIR)

    class P(RPOProjectionMixin):
        pass
    proj = P()
    proj.env, proj.corr_lr, proj.eval_lr, proj.max_steps, proj.eval_steps = env, 1e-4, 1e-4, 10, 50
    proj.corr_eps, proj.corr_momentum, proj.corr_mode = 1e-5, 0.0, 0
    actor = algo.agent.actor
    import copy
    actor_d = copy.deepcopy(actor).to(device)
    actor_d.box_constraint = BoxConstraint(*env.box_constraint_partial, device=torch.device(device), volatile=env.volatile,
                                           update=env.update)

    def policy(state):
        return actor_d(state)                                            # take_action(deterministic=True), ddpg_pa.py:101-115
    stats = evaluate_batched(venv, policy, proj, max_len=500)
    ref_stats = [float(v) for v in ref_stats]
    # (reward mean, std, mean-ineq mean, std, mean-eq mean, std, max-ineq mean, std, max-eq mean, std)
    for i, (a, b) in enumerate(zip(stats, ref_stats)):
        scale = max(abs(b), 1e-3 if i >= 2 else 1.0)
        assert abs(a - b) <= tol * scale + (2e-6 if i >= 4 else 0.0), (i, a, b)
    return stats, ref_stats
