#!/usr/bin/env python
"""Data-parallel RPO-DDPG / RPO-SAC on EVOPF-v0 with the B200 projection path (BASELINE.json configs[2] / configs[3]).

  python examples/train_evopf_dp.py                                   # 1 GPU, DDPG
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_evopf_dp.py --algo sac --graph

Every rank owns `--envs` EVOPF replicas stepped on the GPU (rpo_b200.vector) and a device-resident replay buffer;
the projection (completion + correction, GRG backward) runs with NO communication; the only collective is ONE
all-reduce of the flat actor / critic gradient bucket per optimiser step (rpo_b200.dist.FlatGradBucket).
The update rules follow rpo/algo/rpo_ddpg.py:163-205,307-337 (critic smooth-L1 on the projected target action,
actor loss -Q + nju . ineq_dist through complete_partial, soft target update, dual ascent on nju) or, with --algo sac,
rpo/algo/rpo_sac.py:167-219,322-355 with the tanh-Gaussian policy of rpo/algo/model/policy.py:34-66 (twin critics,
entropy term alpha * log_pi with the box's action_scale in the log-density, automatic temperature).
This is a harness around the hot path, not a re-implementation of the reference's agent layer.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from rpo_b200.algo import RPOProjectionMixin, complete_partial_squashed, process_action_squashed  # noqa: E402
from rpo_b200.box import BoxConstraint  # noqa: E402
from rpo_b200.dist import FlatGradBucket, broadcast_parameters  # noqa: E402
from rpo_b200.env import EVOPFEnv  # noqa: E402
from rpo_b200.vector import DeviceReplay, EvopfVectorEnv  # noqa: E402


class Projector(RPOProjectionMixin):
    def __init__(self, env, a):
        self.env = env
        self.corr_lr = self.eval_lr = a.corr_lr
        self.max_steps, self.eval_steps = a.max_steps, 50
        self.corr_eps, self.corr_momentum, self.corr_mode = 1e-5, 0.0, 0
        self.exact_loop_exit = False          # every shard takes max_steps steps: no host sync inside the training step


def mlp(i, o, h=256, out_act=None):
    layers = [nn.Linear(i, h), nn.ReLU(), nn.Linear(h, h), nn.ReLU(), nn.Linear(h, o)]
    if out_act is not None:
        layers.append(out_act)
    return nn.Sequential(*layers)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--corr-lr", type=float, default=1e-4)
    ap.add_argument("--max-steps", type=int, default=10)
    ap.add_argument("--policy-fre", type=int, default=4)
    ap.add_argument("--algo", choices=("ddpg", "sac"), default="ddpg")
    ap.add_argument("--graph", action="store_true", help="capture the training iteration in CUDA graphs (SURVEY f-3)")
    ap.add_argument("--graph-full", action="store_true",
                    help="with --graph: ONE graph per iteration, the two NCCL all-reduces captured inside it")
    ap.add_argument("--profile", type=int, default=0, metavar="N",
                    help="rank 0: torch.profiler (CUPTI) kernel table of N iterations after the timed loop -> --profile-out")
    ap.add_argument("--profile-out", default="gpurun_out/train_profile.md")
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1000 + rank)                      # different seeds: broadcast_parameters must equalise the replicas
    env = EVOPFEnv(device=dev)
    proj = Projector(env, a)
    venv = EvopfVectorEnv(a.envs, dev, seed=rank, default_generator=a.graph)
    box = BoxConstraint(*env.box_constraint_partial, device=dev, volatile=env.volatile, update=env.update)
    sac = a.algo == "sac"
    n_q = 2 if sac else 1                               # twin critics as one network with two outputs (rpo_sac.py:344-352)
    actor, critic = mlp(57, 28 if sac else 14).to(dev), mlp(57 + 43, n_q).to(dev)
    actor_t, critic_t = mlp(57, 28 if sac else 14).to(dev), mlp(57 + 43, n_q).to(dev)
    log_alpha = torch.zeros(1, device=dev, requires_grad=True)
    target_entropy = -14.0
    nju = nn.Linear(58, 1, bias=False).to(dev)
    nn.init.zeros_(nju.weight)
    broadcast_parameters([actor, critic, nju])
    actor_t.load_state_dict(actor.state_dict())
    critic_t.load_state_dict(critic.state_dict())
    b_critic = FlatGradBucket(critic.parameters())
    b_actor = FlatGradBucket(list(actor.parameters()) + list(nju.parameters()) + ([log_alpha] if sac else []))
    opt_c = torch.optim.Adam(critic.parameters(), lr=3e-4, capturable=a.graph)
    opt_a = torch.optim.Adam(actor.parameters(), lr=1e-4, capturable=a.graph)
    opt_n = torch.optim.Adam(nju.parameters(), lr=2e-2, maximize=True, capturable=a.graph)
    opt_al = torch.optim.Adam([log_alpha], lr=3e-4, capturable=a.graph)
    replay = DeviceReplay(a.envs * 64, dev, default_generator=a.graph, state=57, action=43, next_state=57, reward=1, done=1)

    def pre_box(net, s, noise=0.0):
        """The policy output BEFORE the action box, and whether the box still has to apply the tanh: the box itself is
        fused into the projection / completion kernels (rpo_evopf_project_squashed, SURVEY f-1; policy.py:30-31)."""
        if sac:
            return sample(net, s)[0], False
        if not noise:
            return net(s), True
        u = torch.tanh(net(s))
        return (u + noise * torch.randn_like(u)).clamp(-1, 1), False

    def sample(net, s):
        """tanh-Gaussian policy (policy.py:48-64): returns (tanh sample in [-1, 1] — the box is fused downstream —, log_pi (B,1))."""
        mean, log_std = net(s).split(14, dim=1)
        log_std = (log_std - 3).clamp(-20, 2)
        u = mean + log_std.exp() * torch.randn_like(mean)
        y = torch.tanh(u)
        logp = -0.5 * ((u - mean) / log_std.exp()) ** 2 - log_std - 0.9189385332046727
        logp = logp - torch.log(box.action_scale(s) * (1 - y.pow(2)) + 1e-6)
        return y, logp.sum(1, keepdim=True)

    def project(net, s, noise=0.0):
        """box + completion + correction of the policy output in ONE launch (no-grad call sites, rpo_ddpg.py:96-106,327-332)."""
        u, need_tanh = pre_box(net, s, noise)
        return process_action_squashed(proj, box, s, u, train=True, tanh=need_tanh)

    venv.reset()
    last_rew = torch.zeros(a.envs, device=dev)

    hold = {}

    def seg_rollout_critic():
        """Rollout step of all replicas + critic forward/backward; ends where the critic all-reduce goes."""
        with torch.no_grad():                           # rollout of all replicas (rpo_ddpg.py:96-118, vectorised)
            state = venv.state
            act = project(actor, state, noise=0.1)
            nxt, rew, done, _, _ = venv.step(act, inplace=True)
            replay.extend(state=venv.prev_state, action=act, next_state=nxt, reward=rew, done=done.float())
            last_rew.copy_(rew)
        smp = replay.sample(a.batch)
        hold["state"] = smp["state"]
        with torch.no_grad():                           # critic target (rpo_ddpg.py:327-332 / rpo_sac.py:340-347)
            if sac:
                nz, nlogp = sample(actor, smp["next_state"])
                na = process_action_squashed(proj, box, smp["next_state"], nz, train=True, tanh=False)
                nq = critic_t(torch.cat([smp["next_state"], na], 1)).min(dim=1, keepdim=True).values - log_alpha.exp() * nlogp
            else:
                na = project(actor_t, smp["next_state"])
                nq = critic_t(torch.cat([smp["next_state"], na], 1))
            tq = smp["reward"] + 0.95 * (1 - smp["done"]) * nq
        b_critic.zero_()
        q = critic(torch.cat([smp["state"], smp["action"]], 1))
        F.smooth_l1_loss(q, tq.expand_as(q), reduction="sum").div(q.shape[0]).backward()

    def soft_update(pairs, tau=0.005):
        with torch.no_grad():
            for t, o in pairs:
                torch._foreach_lerp_(list(t.parameters()), list(o.parameters()), tau)

    def seg_critic_step(do_actor: bool):
        """Critic optimiser step (+ actor forward/backward through the GRG kernel); ends at the actor all-reduce."""
        opt_c.step()
        if sac:                                         # SAC: the critic target moves every step (rpo_sac.py:219)
            soft_update(((critic_t, critic),))
        if do_actor:                                    # actor loss (rpo_ddpg.py:307-324)
            b_actor.zero_()
            s = hold["state"]
            if sac:                                     # rpo_sac.py:322-338
                zs, logp = sample(actor, s)
                # autograd: ONE forward kernel (box + Newton completion) and ONE backward kernel (GRG solve + nju . ineq_dist
                # penalty + box derivative) between the policy sample and the critic (rpo_sac.py:324-334)
                act_c, pen = complete_partial_squashed(env, box, s, zs, nju.weight, tanh=False)
                qmin = critic(torch.cat([s, act_c], 1)).min(dim=1, keepdim=True).values
                loss = (log_alpha.exp().detach() * logp - qmin + pen).mean()
                # automatic temperature (rpo_sac.py:207-214); its gradient rides in the actor bucket's all-reduce
                loss = loss - (log_alpha * (logp.detach() + target_entropy)).mean()
            else:
                act_c, pen = complete_partial_squashed(env, box, s, actor(s), nju.weight, tanh=True)   # rpo_ddpg.py:309-319
                loss = (-critic(torch.cat([s, act_c], 1)) + pen).mean()
            loss.backward()

    def seg_actor_step():
        opt_a.step()                                    # rpo_ddpg.py:199-205
        opt_n.step()
        if sac:
            opt_al.step()
        with torch.no_grad():
            nju.weight.clamp_(min=0)
        if not sac:                                     # DDPG: targets move on actor-update steps only (rpo_ddpg.py:205)
            soft_update(((actor_t, actor), (critic_t, critic)))

    def iteration(do_actor: bool, run=None):
        """Three segments separated by the two collectives, which always run eagerly (capturing NCCL collectives inside
        the graph hung in our 2-GPU trial; an eager all-reduce of a 0.3-1.3 MB bucket costs one launch)."""
        run = run or (lambda f, *args: f(*args))
        run(seg_rollout_critic)
        b_critic.all_reduce_mean()                      # the one collective of the critic step
        run(seg_critic_step, do_actor)
        if do_actor:
            b_actor.all_reduce_mean()                   # the one collective of the actor step
            run(seg_actor_step)

    graphs = None
    if a.graph:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                   # warm-up outside capture (allocator, lazy init, optimiser state)
            for w in range(3):
                iteration(True)
                iteration(False)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graphs, pool = {}, torch.cuda.graph_pool_handle()
        if a.graph_full:
            # the whole iteration, collectives included, as one graph per variant.  capture_error_mode="thread_local": NCCL's
            # watchdog thread polls CUDA events, which a GLOBAL-mode capture forbids process-wide (the round-1 trial hung there)
            for key, flag in (("it1", True), ("it0", False)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                    iteration(flag)
                graphs[key] = g
        else:
            for key, fn, args in (("rc", seg_rollout_critic, ()), ("cs1", seg_critic_step, (True,)), ("cs0", seg_critic_step, (False,)),
                                  ("as", seg_actor_step, ())):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    fn(*args)
                graphs[key] = g

        def run_graph(fn, *args):
            key = {seg_rollout_critic: "rc", seg_actor_step: "as"}.get(fn) or ("cs1" if args[0] else "cs0")
            graphs[key].replay()
    else:
        for w in range(3):
            iteration(True)
    n_proj = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    def one(it):
        do_actor = it % a.policy_fre == 0
        if graphs is not None and a.graph_full:
            graphs["it1" if do_actor else "it0"].replay()
        else:
            iteration(do_actor, run_graph if graphs is not None else None)
        return do_actor

    for it in range(a.iters):
        do_actor = one(it)
        n_proj += a.envs + a.batch + (a.batch if do_actor else 0)
    rew = last_rew
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # replicas must still be identical (same broadcast start + same averaged gradients)
    w = torch.cat([p.detach().reshape(-1) for p in list(actor.parameters()) + list(critic.parameters())])
    drift = torch.zeros(1, device=dev)
    if world > 1:
        w0 = w.clone()
        dist.broadcast(w0, src=0)
        drift = (w - w0).abs().max().view(1)
        dist.all_reduce(drift, op=dist.ReduceOp.MAX)
        tot = torch.tensor([float(n_proj)], device=dev)
        dist.all_reduce(tot)
        n_all = float(tot.item())
    else:
        n_all = float(n_proj)
    if a.profile:                                        # CUPTI kernel table of N iterations (rank 0 writes it; all ranks run them)
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for it in range(a.profile):
                one(it)
            torch.cuda.synchronize()
        if rank == 0:
            rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
            if not rows:
                rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
            tot = sum(r[2] for r in rows) or 1.0
            rows.sort(key=lambda r: -r[2])
            os.makedirs(os.path.dirname(a.profile_out) or ".", exist_ok=True)
            with open(a.profile_out, "w") as fh:
                fh.write("# kernel table of %d training iterations (%s, world %d, %s), rank 0, torch.profiler / CUPTI\n\n" %
                         (a.profile, a.algo, world, "CUDA graphs%s" % (" incl. collectives" if a.graph_full else "") if a.graph else "eager"))
                fh.write("flat gradient buckets: critic %d B, actor+nju%s %d B\n\n" % (b_critic.flat.numel() * 4, "+log_alpha" if sac else "", b_actor.flat.numel() * 4))
                fh.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
                for k, c, t in rows[:40]:
                    fh.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k[:110], c, t, 100.0 * t / tot))
                fh.write("\nGPU time of all kernels: %.1f us over %d iterations = %.1f us / iteration\n" % (tot, a.profile, tot / a.profile))
    if rank == 0:
        print(json.dumps({"world": world, "iters": a.iters, "envs_per_gpu": a.envs, "batch_per_gpu": a.batch,
                          "algo": a.algo, "cuda_graph": bool(a.graph), "collectives_in_graph": bool(a.graph and a.graph_full), "iters_per_s": a.iters / dt, "projected_actions_per_s": n_all / dt,
                          "env_steps_per_s": world * a.envs * a.iters / dt, "max_replica_drift": float(drift.item()),
                          "final_reward_mean": float(rew.mean()), "finite": bool(torch.isfinite(w).all())}))
    if world > 1:
        if a.graph and a.graph_full:
            # graphs that captured NCCL kernels keep the communicator busy: release them before tearing it down (the first
            # trial hung in destroy_process_group), and do not wait for NCCL's own teardown
            sys.stdout.flush()
            graphs.clear()
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
