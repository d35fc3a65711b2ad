#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (the GPU box has no reference tree):  python oracle/make_golden.py
The vectors are committed; tests compare both the oracle restatement and the CUDA kernels against them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference, make_cpu_env  # noqa: E402
from rpo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _Algo(object):
    """Bare holder of the attributes RPODDPG.grad_steps reads (rpo_ddpg.py:266-305)."""

    def __init__(self, env, lr, steps, momentum=0.0, eval_steps=50):
        self.env = env
        self.corr_eps = 1e-5
        self.max_steps = steps
        self.corr_mode = 0
        self.grad_eps = 0.02
        self.corr_lr = lr
        self.eval_lr = lr
        self.corr_momentum = momentum
        self.eval_steps = eval_steps


def evopf():
    rpo = load_reference()
    env = make_cpu_env("evopf")
    gs = rpo.algo.RPODDPG.grad_steps
    # (1) case-14 set-point
    ppc = env.ppc
    state = torch.zeros(1, 57)
    state[0, :14] = torch.tensor(ppc["bus"][:, 2] / 100.0, dtype=torch.float32)
    state[0, 14:28] = torch.tensor(ppc["bus"][:, 3] / 100.0, dtype=torch.float32)
    state[0, 28:33] = 0.2
    z = torch.tensor([[0.4, 0, 0, 0, 1.06, 1.045, 1.01, 1.07, 1.09, 0, 0, 0, 0, 0]], dtype=torch.float32)
    zr = z.clone().requires_grad_(True)
    y = env.complete_partial(state, zr)
    y.backward(torch.ones_like(y))
    yd = y.detach()
    np.savez(os.path.join(OUT, "evopf_setpoint.npz"), state=state.numpy(), z=z.numpy(), y=yd.numpy(),
             dz_ones=zr.grad.numpy(), eq_resid=env.eq_resid(state, yd).numpy(), ineq_resid=env.ineq_resid(state, yd).numpy(),
             dir=env.ineq_partial_grad(state, yd).numpy(),
             y_corr=gs(_Algo(env, 1e-4, 10), state, yd, train=True).numpy(),
             G=env.Ybusr.numpy(), B=env.Ybusi.numpy())
    # (2) seeded random batches, both input distributions
    for dist in ("D1", "D2"):
        Bn = 96
        st = synth.evopf_states(Bn, seed=100)
        zz = synth.evopf_partial_actions(st, seed=101, dist=dist)
        g = synth.upstream_grad(Bn, 43, seed=102)
        zr = zz.clone().requires_grad_(True)
        sr = st.clone().requires_grad_(True)
        y = env.complete_partial(sr, zr)
        y.backward(g)
        yd = y.detach()
        yc = gs(_Algo(env, 1e-4, 10), st, yd, train=True)
        ycm = gs(_Algo(env, 1e-4, 5, momentum=0.5), st, yd, train=True)
        yce, k = gs(_Algo(env, 1e-4, 10, eval_steps=50), st[:4], yd[:4], train=False)
        low, high = env.update(st)
        np.savez(os.path.join(OUT, "evopf_%s.npz" % dist), state=st.numpy(), z=zz.numpy(), g=g.numpy(), y=yd.numpy(),
                 dz=zr.grad.numpy(), dx=sr.grad[:, :28].numpy(), eq_resid=env.eq_resid(st, yd).numpy(),
                 ineq_resid=env.ineq_resid(st, yd).numpy(), dir=env.ineq_partial_grad(st, yd).numpy(),
                 y_corr=yc.numpy(), y_corr_mom=ycm.numpy(), y_corr_eval=yce.numpy(), eval_steps=np.array(k),
                 box_low=low.astype(np.float32), box_high=high.astype(np.float32))
    # (3) env.step on a few single instances (numpy path of the reference)
    rows = []
    np.random.seed(7)
    e2 = make_cpu_env("evopf")
    s = e2.reset()
    for t in range(6):
        st_t = torch.tensor(s, dtype=torch.float32).view(1, -1)
        zz = synth.evopf_partial_actions(st_t, seed=200 + t, dist="D1")
        a = e2.complete_partial(st_t, zz).squeeze(0).numpy()
        ns, r, done, info = e2.step(a)
        rows.append(dict(state=s.copy(), action=a.copy(), next_state=ns.copy(), reward=r,
                         ineq_viol=info["ineq_viol"][0], eq_viol=info["eq_viol"][0]))
        s = ns
    np.savez(os.path.join(OUT, "evopf_step.npz"), **{k: np.stack([np.asarray(r[k]) for r in rows]) for k in rows[0]})


def cart():
    rpo = load_reference()
    gs = rpo.algo.RPODDPG.grad_steps
    for seed in (123, 5):            # seed 123 -> partial_actions [1]; seed 5 -> the other one, if it differs
        np.random.seed(seed)
        env = make_cpu_env("cart")
        p = int(env.partial_actions[0])
        st, z = synth.cart_inputs(257, seed=1)
        z = z * 1.3                  # beyond the box so that inequalities bite
        zr = z.clone().requires_grad_(True)
        a = env.complete_partial(st, zr)
        g = synth.upstream_grad(257, 2, seed=2)
        a.backward(g)
        ad = a.detach()
        ac = gs(_Algo(env, 2e-2, 10), st, ad, train=True)
        ace, k = gs(_Algo(env, 2e-2, 10, eval_steps=50), st, ad, train=False)
        # step
        env.reset()
        sts, acts, nss, dones, iv, ev = [], [], [], [], [], []
        for t in range(40):
            s0 = np.array(env.state)
            act = np.random.uniform(-12, 12, size=2).astype(np.float32)
            ns, r, done, info = env.step(act)
            sts.append(s0); acts.append(act); nss.append(ns); dones.append(done)
            iv.append(info["ineq_viol"]); ev.append(info["eq_viol"])
            if done:
                env.reset()
        np.savez(os.path.join(OUT, "cart_p%d.npz" % p), partial_idx=np.array(p), state=st.numpy(), z=z.numpy(), g=g.numpy(),
                 a=ad.numpy(), dz=zr.grad.numpy(), eq_resid=env.eq_resid(st, ad).numpy(), ineq_resid=env.ineq_resid(st, ad).numpy(),
                 dir=env.ineq_partial_grad(st, ad).numpy(), a_corr=ac.numpy(), a_corr_eval=ace.numpy(), eval_steps=np.array(k),
                 step_state=np.stack(sts), step_action=np.stack(acts), step_next=np.stack(nss), step_done=np.array(dones),
                 step_ineq_viol=np.stack(iv), step_eq_viol=np.stack(ev))


def pen():
    rpo = load_reference()
    gs = rpo.algo.RPODDPG.grad_steps
    env = make_cpu_env("pen")
    obs, z = synth.pen_inputs(64, seed=3)
    z = z * 1.5
    # the reference couples rows in ineq_partial_grad for B > 1: golden rows are generated one at a time
    a_l, dz_l, dir_l, ac_l, eq_l, iq_l, ace_l, k_l = [], [], [], [], [], [], [], []
    g = synth.upstream_grad(64, 2, seed=4)
    for i in range(64):
        o, zz = obs[i:i + 1], z[i:i + 1].clone().requires_grad_(True)
        a = env.complete_partial(o, zz)
        a.backward(g[i:i + 1])
        ad = a.detach()
        a_l.append(ad.numpy()); dz_l.append(zz.grad.numpy())
        eq_l.append(env.eq_resid(o, ad).numpy()); iq_l.append(env.ineq_resid(o, ad).numpy())
        dir_l.append(env.ineq_partial_grad(o, ad).numpy())
        ac_l.append(gs(_Algo(env, 2e-3, 10), o, ad, train=True).numpy())
        ace, k = gs(_Algo(env, 2e-3, 10, eval_steps=50), o, ad, train=False)
        ace_l.append(ace.numpy()); k_l.append(k)
    # literal batched call (cross-batch coupling) kept to document the divergence
    ad_all = torch.tensor(np.concatenate(a_l))
    dir_batched = env.ineq_partial_grad(obs, ad_all).numpy()
    env.reset()
    sts, acts, nsts, nobs, rews, dones, iv, ev = [], [], [], [], [], [], [], []
    np.random.seed(11)
    for t in range(40):
        s0 = env.state.copy()
        act = np.random.uniform(-7, 7, size=2).astype(np.float32)
        o2, r, done, info = env.step(act)
        sts.append(s0); acts.append(act); nsts.append(env.state.copy()); nobs.append(o2); rews.append(r); dones.append(done)
        iv.append(info["ineq_viol"][0]); ev.append(info["eq_viol"][0])
        if done:
            env.reset()
    np.savez(os.path.join(OUT, "pen.npz"), obs=obs.numpy(), z=z.numpy(), g=g.numpy(), a=np.concatenate(a_l), dz=np.concatenate(dz_l),
             eq_resid=np.concatenate(eq_l), ineq_resid=np.concatenate(iq_l), dir=np.concatenate(dir_l), dir_batched_reference=dir_batched,
             a_corr=np.concatenate(ac_l), a_corr_eval=np.concatenate(ace_l), eval_steps=np.array(k_l),
             step_state=np.stack(sts), step_action=np.stack(acts), step_next=np.stack(nsts), step_obs=np.stack(nobs),
             step_reward=np.array(rews), step_done=np.array(dones), step_ineq_viol=np.stack(iv), step_eq_viol=np.stack(ev))


def box():
    """BoxConstraint (rpo/algo/model/utils.py:5-118) on the EVOPF volatile partial box and on static boxes."""
    rpo = load_reference()
    from rpo.algo.model.utils import BoxConstraint
    import contextlib
    import io
    env = make_cpu_env("evopf")
    cpu = torch.device("cpu")
    with contextlib.redirect_stdout(io.StringIO()):           # the constructor prints its arrays
        bx = BoxConstraint(*env.box_constraint_partial, device=cpu, volatile=env.volatile, update=env.update)
        bx_sig = BoxConstraint(*env.box_constraint_partial, device=cpu, style="sigmoid", volatile=True, update=env.update)
        bx_full = BoxConstraint(*env.box_constraint, device=cpu)
    Bn = 64
    st = synth.evopf_states(Bn, seed=300)
    st[:4, 28:33] = torch.tensor([[0.1, 0.8, 0.6, 0.3, 0.75], [0.05, 0.9, 0.45, 0.1, 0.8], [0.6000001, 0.3, 0.2999999, 0.5, 0.7],
                                  [0.2, 0.2, 0.2, 0.2, 0.2]])
    g = torch.Generator().manual_seed(301)
    pre = (1.5 * torch.randn(Bn, 14, generator=g)).float()
    up = torch.randn(Bn, 14, generator=g).float()
    out = {"state": st.numpy(), "pre": pre.numpy(), "up": up.numpy()}
    x = pre.clone().requires_grad_(True)
    z = bx(torch.tanh(x), st)
    z.backward(up)
    out["tanh_call"], out["tanh_call_grad"] = z.detach().numpy(), x.grad.numpy()
    t = torch.tanh(pre).requires_grad_(True)
    z = bx(t, st)
    z.backward(up)
    out["call"], out["call_grad"], out["call_in"] = z.detach().numpy(), t.grad.numpy(), t.detach().numpy()
    lo, hi = env.update(st)
    out["low"], out["high"] = lo, hi
    raw = (torch.tensor(lo) + (torch.tensor(hi) - torch.tensor(lo)) * (torch.rand(Bn, 14, generator=g) * 1.6 - 0.3)).float()
    raw[0, :3] = torch.tensor(lo[0, :3])                      # exactly on the bounds: torch.clip passes the gradient
    raw[0, 9:12] = torch.tensor(hi[0, 9:12])
    r = raw.clone().requires_grad_(True)
    c = bx.clip(r, st)
    c.backward(up)
    out["raw"], out["clip"], out["clip_grad"] = raw.numpy(), c.detach().numpy(), r.grad.numpy()
    out["scale"] = bx.action_scale(st).numpy()
    u01 = torch.rand(Bn, 14, generator=g).float()
    out["u01"], out["sigmoid_call"] = u01.numpy(), bx_sig(u01, st).numpy()
    x43 = (torch.rand(Bn, 43, generator=g) * 2 - 1).float()
    out["x43"], out["full_call"] = x43.numpy(), bx_full(x43).numpy()
    out["full_clip"] = bx_full.clip(3 * x43).numpy()
    out["full_low"], out["full_high"] = env.box_constraint
    np.savez(os.path.join(OUT, "box.npz"), **out)


def pen_batched():
    """The reference's BATCHED pendulum calls (rows coupled through the (B,B) product, pendulum.py:337-339): direction and
    RPODDPG.grad_steps on the whole batch of tests/golden/pen.npz at once.  Pins the opt-in reference_batch_coupling mode."""
    rpo = load_reference()
    env = make_cpu_env("pen")
    gs = rpo.algo.RPODDPG.grad_steps
    g = np.load(os.path.join(OUT, "pen.npz"))
    obs, a = torch.tensor(g["obs"]), torch.tensor(g["a"])
    out = {"dir": env.ineq_partial_grad(obs, a).numpy()}
    out["a_corr"] = gs(_Algo(env, 2e-3, 10), obs, a, train=True).numpy()
    ace, k = gs(_Algo(env, 2e-3, 10, momentum=0.5, eval_steps=50), obs, a, train=False)
    out["a_corr_eval_momentum"], out["eval_steps"] = ace.numpy(), np.array(k)
    np.savez(os.path.join(OUT, "pen_batched.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    which = sys.argv[1:] or ["evopf", "cart", "pen", "box", "pen_batched"]
    for name in which:
        {"evopf": evopf, "cart": cart, "pen": pen, "box": box, "pen_batched": pen_batched}[name]()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
