"""BASELINE.json configs[2] / configs[3] literally, on hardware: the UNMODIFIED reference classes rpo.algo.RPODDPG and
RPOSAC (agents, MLPs, replay buffer, `run()`, `train()`, `eval()` — scripts/evopf_exp.py:27-32, scripts/evopf_exp_sac.py:
30-35 settings) driving the CUDA kernels through `rpo_b200.algo.install()`, next to the all-reference CPU run on the same
seeds.  The reference tree on the GPU box is the verbatim copy oracle/make_ref_copy.sh placed under oracle/_ref/
(checked file by file against the SHA-256 list of the original); gym / pypower / igraph come from oracle/stubs.

CPU twin (host simulator instead of the kernels): tests/test_dropin_reference_cpu.py."""
import os
import time
import warnings

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.reference_available(),
                                 reason="no reference tree: run oracle/make_ref_copy.sh in the build container before gpurun")]

HP_DDPG = dict(batch_size=256, max_steps=10, warmup=0, lr_dual=2e-2, corr_lr=1e-4, eps_epoch=20000, eval_lr=1e-4,
               eval_steps=50, grad_eps=0.02, corr_momentum=0.0, policy_fre=4, ex_action_dim=1, gamma=0.95,
               capacity=20000, clip_thres=0.2, shared_param=False, value_type="cat")            # scripts/evopf_exp.py:29-31
HP_SAC = dict(batch_size=256, max_steps=10, warmup=0, lr_dual=2e-2, corr_lr=1e-4, lr_actor=1e-4, lr_critic=3e-4,
              eps_epoch=20000, eval_lr=1e-4, eval_steps=50, grad_eps=0.1, corr_momentum=0.0, policy_fre=4, ex_action_dim=1,
              fixed=False, init_lamb=0.0, init_nju=0.0, alpha=0.001, automatic_entropy_tuning=False, capacity=20000,
              clip_thres=0.2, shared_param=False, value_type="cat")                              # scripts/evopf_exp_sac.py:32-35


class _Log(object):
    def add(self, **kw):
        pass


def _seed():
    np.random.seed(111)                                              # scripts/evopf_exp.py:10-13
    torch.manual_seed(123)
    torch.set_default_dtype(torch.float32)


def _pristine(modname, clsname):
    """A fresh class object compiled from the reference file, with the reference's own BoxConstraint (install() patches
    the imported module in place)."""
    import importlib
    src = importlib.import_module(modname)
    with open(src.__file__) as fh:
        code = compile(fh.read(), src.__file__, "exec")
    ns = dict(src.__dict__)
    from rpo.algo.model.utils import BoxConstraint as RefBox
    ns["BoxConstraint"] = RefBox
    exec(code, ns)
    return ns[clsname]


@pytest.fixture()
def installed():
    rpo = ref_loader.load_reference()
    if os.path.isdir(ref_loader.REF_COPY) and ref_loader.REFERENCE_ROOT == ref_loader.REF_COPY:
        assert ref_loader.verify_copy() >= 40                       # the copy IS the unmodified reference
    import gym
    import importlib
    import rpo.algo as ralgo
    import rpo_b200.algo as balgo
    saved = {cls: (cls.process_action, cls.grad_steps) for cls in (ralgo.RPODDPG, ralgo.RPOSAC)}
    saved_registry = dict(gym.envs.registration.registry)
    mods = [importlib.import_module(n) for n in ("rpo.algo.rpo_ddpg", "rpo.algo.rpo_sac")]
    saved_box = [m.BoxConstraint for m in mods]
    balgo.install(rpo)
    yield rpo
    for m, b in zip(mods, saved_box):
        m.BoxConstraint = b
    for cls, (pa, gs) in saved.items():
        cls.process_action, cls.grad_steps = pa, gs
        if "exact_loop_exit" in cls.__dict__:
            del cls.exact_loop_exit
    gym.envs.registration.registry.clear()
    gym.envs.registration.registry.update(saved_registry)


def _traj(algo):
    rb = algo.agent.replay_buffer
    return {k: rb.buffer[k][:rb.size].copy() for k in ("state", "action", "next_state", "reward", "eq_viol", "ineq_viol")}


def _run(algo):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        algo.run(eval=False)
        torch.cuda.synchronize()
        return time.perf_counter() - t0


def test_reference_rpoddpg_drives_the_cuda_kernels(installed, capsys):
    """200 iterations of the reference's RPODDPG.run (rpo_ddpg.py:79-161) with scripts/evopf_exp.py settings: (b) the
    reference agent on gym.make('EVOPF-v0') -> rpo_b200.env.EVOPFEnv on cuda:0, (a) the all-reference CPU run."""
    import gym
    import rpo_b200.algo as balgo
    from rpo_b200 import _lib
    from rpo_b200.env import EVOPFEnv
    rpo = installed
    N = 200
    # exploration noise off: torch.randn_like draws from different generators on cpu and cuda (ddpg_pa.py:113)
    noise = dict(eps=0.0, eps_start=0.0)
    _seed()
    env_b = gym.make("EVOPF-v0")
    assert isinstance(env_b, EVOPFEnv) and env_b.device.type == "cuda"
    assert env_b.data_dir and os.path.samefile(env_b.data_dir, ref_loader.DATA_DIR)     # found the reference's pickles by itself
    assert rpo.algo.RPODDPG.process_action is balgo.RPOProjectionMixin.process_action
    algo_b = rpo.algo.RPODDPG(env_b, "./test", name="t", logger=_Log(), max_epochs=N, device=torch.device("cuda"),
                              **HP_DDPG, **noise)
    assert _lib._lib is not None or _lib.load()                      # the CUDA extension is the thing that runs
    dt_b = _run(algo_b)
    traj_b = _traj(algo_b)
    w_b = torch.cat([p.detach().reshape(-1).cpu() for p in algo_b.agent.actor.parameters()])

    RefDDPG = _pristine("rpo.algo.rpo_ddpg", "RPODDPG")
    _seed()
    env_a = ref_loader.make_cpu_env("evopf")
    algo_a = RefDDPG(env_a, "./test", name="t", logger=_Log(), max_epochs=N, device=torch.device("cpu"), **HP_DDPG, **noise)
    torch.set_num_threads(os.cpu_count() or 1)
    dt_a = _run(algo_a)
    traj_a = _traj(algo_a)
    w_a = torch.cat([p.detach().reshape(-1) for p in algo_a.agent.actor.parameters()])

    assert traj_a["state"].shape == traj_b["state"].shape == (N, 57)
    assert np.abs(traj_a["state"][0] - traj_b["state"][0]).max() < 1e-6          # same loaders, same RNG order
    rel = np.abs(traj_a["action"] - traj_b["action"]).max(1) / np.abs(traj_a["action"]).max(1)
    # closed loop: tight until the first instance grazes an inequality bound (mask discontinuity) or the two
    # float32 MLP implementations (CPU / cuBLAS) round differently; close afterwards
    n_tight = int(np.argmax(rel > 1e-4)) if (rel > 1e-4).any() else N
    assert n_tight >= 4, rel[:10]
    assert np.abs(traj_a["reward"][:n_tight] - traj_b["reward"][:n_tight]).max() <= 1e-4 * np.abs(traj_a["reward"]).max()
    assert np.abs(traj_a["ineq_viol"][:n_tight] - traj_b["ineq_viol"][:n_tight]).max() < 1e-4
    # whole run: every episode has the same exogenous series (identical numpy RNG streams) and both agents stay in the
    # same regime: equality violations at the Newton tolerance scale after correction, rewards and actions close
    assert np.abs(traj_a["state"][:, :28] - traj_b["state"][:, :28]).max() < 1e-6
    assert np.median(rel) < 5e-2
    r_a, r_b = traj_a["reward"].mean(), traj_b["reward"].mean()
    assert abs(r_a - r_b) <= 0.05 * abs(r_a), (r_a, r_b)
    assert abs(np.abs(traj_a["eq_viol"]).max(1).mean() - np.abs(traj_b["eq_viol"]).max(1).mean()) < 5e-3
    assert abs(traj_a["ineq_viol"].max(1).mean() - traj_b["ineq_viol"].max(1).mean()) < 5e-2
    assert bool(torch.isfinite(w_b).all()) and float((w_a - w_b).abs().max()) < 5e-2
    with capsys.disabled():
        print("\n[configs[2]] reference RPODDPG.run, %d iterations, batch 256: rpo_b200 on cuda:0 %.1f it/s | all-reference on "
              "%d CPU threads %.1f it/s | tight prefix %d steps, median action rel diff %.2g, mean reward %.4f vs %.4f"
              % (N, N / dt_b, torch.get_num_threads(), N / dt_a, n_tight, float(np.median(rel)), r_b, r_a))


def test_reference_rposac_and_eval_on_the_cuda_kernels(installed, capsys):
    """configs[3]'s agent (RPOSAC, rpo_sac.py:83-219) for 200 iterations on the kernels, then the reference's own eval()
    (rpo_sac.py:221-278: 10 episodes, process_action(train=False) with K = 50 -> (action, step))."""
    import gym
    rpo = installed
    N = 200
    _seed()
    env = gym.make("EVOPF-v0")
    sac = rpo.algo.RPOSAC(env, "./test", name="t", logger=_Log(), max_epochs=N, eps=0.0001, eps_start=0.0001,
                          device=torch.device("cuda"), **HP_SAC)
    dt = _run(sac)
    rb = sac.agent.replay_buffer
    assert rb.size == N and np.isfinite(rb.buffer["reward"][:N]).all() and np.isfinite(rb.buffer["action"][:N]).all()
    # after the correction the stored actions satisfy the power-flow equations to the linearisation error of 10 steps
    assert np.abs(rb.buffer["eq_viol"][:N]).max() < 5e-2
    w = torch.cat([p.detach().reshape(-1) for p in sac.agent.actor.parameters()])
    assert bool(torch.isfinite(w).all())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = sac.eval()
    assert len(out) == 10 and all(np.isfinite(v) for v in out)
    with capsys.disabled():
        print("\n[configs[3] agent, 1 GPU] reference RPOSAC.run, %d iterations, batch 256 on cuda:0: %.1f it/s; eval(): "
              "reward %.3f +- %.3f, max ineq viol %.2e, max eq viol %.2e" % (N, N / dt, out[0], out[1], out[6], out[8]))


@pytest.mark.parametrize("id_,lr,script", [("CartSafe-v0", 2e-2, "cart_exp.py"), ("SpringPendulum-v0", 2e-3, "pen_exp.py")])
def test_reference_ddpg_on_classic_envs_on_the_cuda_kernels(installed, id_, lr, script):
    """configs[0] / configs[1]: scripts/cart_exp.py:26-28, scripts/pen_exp.py:26-29 settings, 150 iterations + eval()."""
    import gym
    rpo = installed
    _seed()
    env = gym.make(id_)
    ddpg = rpo.algo.RPODDPG(env, "./test", name="t", logger=_Log(), batch_size=64, max_steps=10, warmup=0, corr_lr=lr,
                            eval_lr=lr, eval_steps=50, grad_eps=0.1, corr_momentum=0.0, policy_fre=4, max_epochs=150,
                            capacity=2000, shared_param=True, value_type="add", device=torch.device("cuda"))
    _run(ddpg)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = ddpg.eval()
    rb = ddpg.agent.replay_buffer
    assert rb.size == 150 and np.isfinite(out[0])
    # the projected actions the agent stored satisfy the equality exactly (closed-form completion) and, after correction,
    # the force limits to the loop tolerance in the overwhelming majority of steps
    assert np.abs(rb.buffer["eq_viol"][:150]).max() < 1e-4
