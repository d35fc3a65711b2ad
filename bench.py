#!/usr/bin/env python
"""Benchmark of the RPO projection hot path (BASELINE.json metric): projected actions / second.

One step = one pass of the hot path over one synthetic EVOPF-v0 batch of 2^20 instances per GPU:
  fused completion (Newton power flow) + 10 post-correction steps   (rpo_evopf_project, 1 kernel)
  + GRG backward of the completion                                  (rpo_evopf_complete_bwd, 1 kernel).
Instances shard across ranks with no collective (weak scaling: 2^20 instances per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--log2-batch n] [--dist D1|D2] [--strong]            (configs[4] sweep points)
                  [--workload train [--algo ddpg|sac]]                   (configs[2] / configs[3]'s agent: the unmodified
                                                                          reference RPODDPG.run / RPOSAC.run over the kernels)
  (N > 1: launched by torch.distributed.run, one rank per GPU.)

`--impl reference` times the UNMODIFIED reference (verbatim copy under oracle/_ref, placed by oracle/make_ref_copy.sh;
never /root/reference, which does not exist on the GPU box) on the host cores; if that copy is missing it falls back
to the oracle port and says so (`cpu_baseline.kind`).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "RPO projected actions/sec (completion+correction, fwd+bwd) at 1/2/4/8 B200"      # BASELINE.json "metric", verbatim
UNIT = "projected actions/s"
LOG2_B = 20
CORR = dict(lr=1e-4, momentum=0.0, corr_eps=1e-5, n_steps=10)      # scripts/evopf_exp.py:29-31
NEWTON = dict(tol=1e-5, max_iters=50)                               # evopf.py:786


# algorithmic flops per instance (dense-minimal formulation, 1 FMA = 2 flop)
def flops_project(n_newton: float, k: int) -> float:
    return 12.9e3 * n_newton + 1.8e3 + 17.3e3 * k


FLOPS_BWD = 13.5e3
BYTES_PER_ACTION = 932.0       # 188 read + 172 + 172 written fwd; 344 read + 56 written bwd


class ClockSampler(object):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML; nvidia-smi as a fallback)."""

    def __init__(self, index: int, period: float = 0.01):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self.backend = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.backend = "nvml"
        except Exception:
            self.nv = None

    def _reasons(self, mask: int):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        return [k for k, bit in names.items() if mask & bit]

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons.update(self._reasons(mask))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        if not self.samples:
            return self._smi_once()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "backend": self.backend}

    def _smi_once(self):
        try:
            import subprocess
            out = subprocess.check_output(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                 "--format=csv,noheader,nounits"], timeout=10).decode().strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [out[2].strip()], "samples": 1,
                    "backend": "nvidia-smi (after the timed region)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "backend": None}


def cpu_sample(n_inst: int, chunk: int, threads: int):
    """Time the oracle port (torch CPU restatement of the reference path) on a bounded sample; returns inst/s."""
    from oracle.rpo_oracle import EvopfOracle, grad_steps
    from rpo_b200 import synth
    torch.set_num_threads(threads)
    orc = EvopfOracle()
    st = synth.evopf_states(n_inst, seed=0)
    z = synth.evopf_partial_actions(st, seed=1)
    g = synth.upstream_grad(n_inst, 43)
    t0 = time.perf_counter()
    for b0 in range(0, n_inst, chunk):
        s, zz, gg = st[b0:b0 + chunk], z[b0:b0 + chunk], g[b0:b0 + chunk]
        y, jac, inv, _ = orc.complete_partial_fwd(s, zz, return_saved=True)     # PFFunction.forward
        grad_steps(orc, s, y, CORR["lr"], CORR["momentum"], CORR["corr_eps"], CORR["n_steps"])   # grad_steps(train)
        orc.complete_partial_bwd(gg, jac, inv)                                  # PFFunction.backward
    dt = time.perf_counter() - t0
    return n_inst / dt, dt


REF_COPY = os.path.join(ROOT, "oracle", "_ref")


def reference_copy_available() -> bool:
    return os.path.isfile(os.path.join(REF_COPY, "rpo", "env", "electrical_grid", "evopf.py"))


def load_reference_copy():
    """The unmodified reference package from oracle/_ref (stubs for gym / pypower / igraph from oracle/stubs)."""
    os.environ["RPO_REFERENCE_ROOT"] = REF_COPY
    from oracle import ref_loader
    assert ref_loader.REFERENCE_ROOT == REF_COPY, "bench.py must not read /root/reference"
    return ref_loader.load_reference(), ref_loader


class _NullLog(object):
    def add(self, **kw):
        pass


TRAIN_HP = {                                                          # scripts/evopf_exp.py:29-31, scripts/evopf_exp_sac.py:32-35
    "ddpg": dict(batch_size=256, max_steps=10, warmup=0, lr_dual=2e-2, corr_lr=1e-4, eps=0.0001, eps_start=0.0001,
                 eps_epoch=20000, eval_lr=1e-4, eval_steps=50, grad_eps=0.02, corr_momentum=0.0, policy_fre=4,
                 ex_action_dim=1, gamma=0.95, capacity=20000, clip_thres=0.2, shared_param=False, value_type="cat"),
    "sac": dict(batch_size=256, max_steps=10, warmup=0, lr_dual=2e-2, corr_lr=1e-4, eps=0.0001, eps_start=0.0001,
                lr_actor=1e-4, lr_critic=3e-4, eps_epoch=20000, eval_lr=1e-4, eval_steps=50, grad_eps=0.1,
                corr_momentum=0.0, policy_fre=4, ex_action_dim=1, fixed=False, init_lamb=0.0, init_nju=0.0, alpha=0.001,
                automatic_entropy_tuning=False, capacity=20000, clip_thres=0.2, shared_param=False, value_type="cat"),
}


def _reference_algo(device, algo="ddpg", max_epochs=1, env=None):
    """An instance of the reference's own RPODDPG / RPOSAC with the script settings on `device`."""
    import warnings
    rpo, ref_loader = load_reference_copy()
    if env is None:
        env = ref_loader.make_cpu_env("evopf")
    cls = rpo.algo.RPODDPG if algo == "ddpg" else rpo.algo.RPOSAC
    os.makedirs(os.path.join(ROOT, "test"), exist_ok=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):     # the reference's BoxConstraint prints its box
            return cls(env, "./test", name="bench", logger=_NullLog(), max_epochs=max_epochs, device=device, **TRAIN_HP[algo])


_REF_ALGO = {}


def cpu_sample_reference(n_inst: int, chunk: int, threads: int, dist: str = "D1"):
    """Time the UNMODIFIED reference on a bounded sample: env.complete_partial (autograd on) ->
    RPODDPG.grad_steps(train=True) under no_grad -> y.backward(dL/dy), in chunks; returns (inst/s, seconds)."""
    import warnings
    from rpo_b200 import synth
    torch.set_num_threads(threads)
    if "a" not in _REF_ALGO:
        _REF_ALGO["a"] = _reference_algo(torch.device("cpu"))
    algo = _REF_ALGO["a"]
    st = synth.evopf_states(n_inst, seed=0)
    z = synth.evopf_partial_actions(st, seed=1, dist=dist)
    g = synth.upstream_grad(n_inst, 43)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        for b0 in range(0, n_inst, chunk):
            s, zz, gg = st[b0:b0 + chunk], z[b0:b0 + chunk].clone().requires_grad_(True), g[b0:b0 + chunk]
            y = algo.env.complete_partial(s, zz)                         # evopf.py:724-727 -> PFFunctionFn.forward
            with torch.no_grad():
                algo.grad_steps(s, y.detach(), train=True)               # rpo_ddpg.py:266-286
            y.backward(gg)                                               # PFFunctionFn.backward, evopf.py:857-910
        dt = time.perf_counter() - t0
    return n_inst / dt, dt


def cpu_baseline_sample(n_inst, chunk, threads, dist="D1"):
    """(value, seconds, kind, description): the reference itself when its copy is present, else the oracle port."""
    if reference_copy_available():
        v, dt = cpu_sample_reference(n_inst, chunk, threads, dist)
        return v, dt, "reference", ("the UNMODIFIED reference (oracle/_ref): EVOPFEnv.complete_partial + RPODDPG.grad_steps(10) "
                                    "+ PFFunction.backward, torch CPU")
    v, dt = cpu_sample(n_inst, chunk, threads)
    return v, dt, "port", "torch-CPU oracle port of complete_partial + grad_steps(10) + PFFunction.backward (oracle/_ref missing)"


def run_reference(args, rank: int):
    """--impl reference: the reference's own CPU implementation of the path on all host threads, same metric / config,
    a bounded sample per step.  Rank 0 only (the other ranks of a torchrun launch exit without work)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.workload == "train":
        return run_reference_train(args, threads)
    n_step, chunk = 2048, 1024
    for _ in range(max(args.warmup, 1)):
        _, _, kind, what = cpu_baseline_sample(n_step, chunk, threads, args.dist)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_baseline_sample(n_step, chunk, threads, args.dist)
    dt = time.perf_counter() - t0
    val = n_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1 << args.log2_batch, device="cpu",
                                  reference_sample="each step times %d instances of that workload on the host CPU (bounded sample)" % n_step),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d instances per step, chunks of %d; %s" % (n_step, chunk, what)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, B, **extra):
    cfg = {"workload": "synthetic EVOPF-v0 completion+correction (fwd+bwd), batch 2^%d instances %s" %
                       (args.log2_batch, "in total (strong scaling)" if args.strong else "per GPU"),
           "input_dist": "D1 (uniform in the state-dependent partial box)" if args.dist == "D1" else "D2 (central 30 % of the box)",
           "newton": NEWTON, "correction": CORR}
    cfg.update(extra)
    return cfg


TRAIN_METRIC = "RPODDPG.run iterations/sec (1 env step + 1 train() call, replay batch 256), EVOPF-v0, scripts/evopf_exp.py settings"


def _train_line(args, value, dt, world, device, extra):
    name = "RPODDPG" if args.algo == "ddpg" else "RPOSAC"
    return {
        "metric": TRAIN_METRIC.replace("RPODDPG", name).replace("evopf_exp.py", "evopf_exp.py" if args.algo == "ddpg" else "evopf_exp_sac.py"),
        "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (the reference's demand / price generators)",
        "config": dict({"workload": "BASELINE.json configs[%d]: the UNMODIFIED reference %s.run on EVOPF-v0, hyper-parameters of scripts/%s"
                                    % (2 if args.algo == "ddpg" else 3, name, "evopf_exp.py:29-31" if args.algo == "ddpg" else "evopf_exp_sac.py:32-35"),
                        "device": device, "hyper_parameters": {k: v for k, v in TRAIN_HP[args.algo].items()},
                        "parallelism": "replicas only (the reference has no distributed code): %d independent run(s)" % world}, **extra),
    }


def run_reference_train(args, threads):
    """All-reference configs[2] on the host cores: RPODDPG.run of the unmodified classes and environment, device cpu."""
    import warnings
    torch.set_num_threads(threads)
    steps = min(args.steps, 40)                                       # ~5-15 it/s on the host: a bounded sample
    algo = _reference_algo(torch.device("cpu"), args.algo, max_epochs=max(args.warmup, 1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):
            algo.run(eval=False)
            algo.max_epochs = steps
            t0 = time.perf_counter()
            algo.run(eval=False)
            dt = time.perf_counter() - t0
    args.steps = steps
    line = _train_line(args, steps / dt, dt, args.gpus, "cpu", {})
    line.update({"impl": "reference", "cpu_baseline": {"value": steps / dt, "unit": "iterations/s", "cores": threads, "kind": "reference",
                                                       "sample": "%d iterations of the unmodified reference on the host cores" % steps},
                 "e2e": {"value": steps / dt, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
    print(json.dumps(line), flush=True)


def run_train(args, rank, local_rank, world):
    """--workload train: the reference's own RPODDPG / RPOSAC classes (agents, MLPs, replay buffer, run / train loops,
    all unmodified) through rpo_b200.algo.install(): gym.make('EVOPF-v0') resolves to the drop-in environment on the GPU,
    process_action / grad_steps / BoxConstraint run on the kernels.  N > 1: N independent replicas (no collective)."""
    import warnings
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rpo, _ = load_reference_copy()
    import gym
    import numpy as np
    import rpo_b200.algo as balgo
    balgo.install(rpo)
    np.random.seed(111 + rank)
    torch.manual_seed(123 + rank)
    env = gym.make("EVOPF-v0", device=dev)
    algo = _reference_algo(dev, args.algo, max_epochs=max(args.warmup, 3), env=env)
    sampler = ClockSampler(local_rank)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):
            algo.run(eval=False)                                      # W untimed warm-up iterations
            algo.max_epochs = args.steps
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            sampler.start()
            t0 = time.perf_counter()
            algo.run(eval=False)                                      # exactly K timed iterations
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
    clocks = sampler.stop()
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_max = float(t.item())
    if rank == 0:
        rb = algo.agent.replay_buffer
        line = _train_line(args, world * args.steps / dt_max, dt_max, world, "cuda", {
            "timed_region": "RPODDPG.run for K iterations after a W-iteration run of the same object, wall clock between "
                            "synchronizes (the loop is host-driven), max over ranks",
            "replay_size": int(rb.size), "max_eq_viol_in_buffer": float(abs(rb.buffer["eq_viol"][:rb.size]).max())})
        line.update({"clocks": clocks, "e2e": {"value": world * args.steps / dt_max, "unit": "iterations/s",
                                               "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                               "note": "the reference's loop keeps states / replay on the host: every iteration already "
                                                       "crosses PCIe both ways inside the timed region"},
                     "gpu_launches": None, "cpu_baseline": None})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="project", choices=["project", "train"])
    ap.add_argument("--algo", default="ddpg", choices=["ddpg", "sac"], help="--workload train: RPODDPG (configs[2]) or RPOSAC (configs[3]'s agent)")
    ap.add_argument("--log2-batch", type=int, default=LOG2_B)
    ap.add_argument("--dist", default="D1", choices=["D1", "D2"])
    ap.add_argument("--strong", action="store_true", help="2^log2-batch instances in TOTAL, split over the ranks (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        # project: 400 x 1.35 ms = a 0.54 s timed region (the clock sampler needs more than a handful of samples);
        # reference arm: a few bounded samples; train: 200 iterations of the host-driven loop
        args.steps = 400 if (args.impl == "b200" and args.workload == "project") else (200 if args.impl == "b200" else 10)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "train":
        run_train(args, rank, local_rank, world)
        return
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE (%d) != --gpus (%d)" % (world, args.gpus))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs torch.distributed.run (one rank per GPU)" % args.gpus)
    args.warmup = max(args.warmup, 3)

    from rpo_b200 import ops, synth
    from rpo_b200.host_api import HostProjector
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        # NCCL may print its version banner on fd 1 (NCCL_DEBUG=VERSION on some boxes): stdout must carry the JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        # NCCL builds its communicator lazily at the first collective (hundreds of ms with the GPU idle): do it here,
        # not at the barrier that opens the timed region — the idle GPU would drop its clocks and the first ~50 ms of
        # the timed region would measure the ramp (seen as 0.43 ms instead of 0.15 ms on the backward kernel).
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)

    B = (1 << args.log2_batch) // (world if args.strong else 1)
    # each rank owns its own contiguous shard of the global batch (seeded per rank); no data-path collective
    st_h = synth.evopf_states(B, seed=1000 + rank)
    z_h = synth.evopf_partial_actions(st_h, seed=2000 + rank, dist=args.dist)
    g_h = synth.upstream_grad(B, 43, seed=3000 + rank)
    st, z, g = st_h.to(dev), z_h.to(dev), g_h.to(dev)

    def step():
        yc, y, flags, status = ops.evopf_project(st, z, CORR["lr"], CORR["momentum"], CORR["corr_eps"], CORR["n_steps"],
                                                 tol=NEWTON["tol"], max_iters=NEWTON["max_iters"], want_y=True,
                                                 want_flags=True, want_status=False)
        dz, _ = ops.evopf_complete_bwd(y, g, need_dx=False)
        return yc, y, flags, dz

    # one step (memset + k_evopf_project + k_evopf_bwd) captured in a CUDA graph: the timed region replays it K times,
    # so host-side jitter (8 ranks sharing one host) cannot starve the GPU between 1.4 ms kernels
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            out = step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    # untimed clock ramp: an idle B200 sits at 120 MHz and needs a few hundred ms of load to reach steady clocks
    # (SM and memory side); a short timed region (K x 1.6 ms) would otherwise measure the ramp, not the kernels
    # (seen as +20 % and 3x on the latency-sensitive backward kernel).
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < 1.0:
        for _ in range(20):
            graph.replay()
        torch.cuda.synchronize()
    # workload facts for the roofline: mean Newton iterations, steps the reference's loop test would take
    _, status = ops.evopf_complete_fwd(st, z, return_status=True, refine=False)      # the fused kernel's own Newton sequence
    n_newton = float((status & 0xff).float().mean())
    not_converged = int(((status >> 8) != 0).sum())      # infeasible power flows exist at ~1e-6 of D1 (reference diverges too)
    fl = out[2].cpu().tolist()
    k_ref = next((i for i in range(1, CORR["n_steps"]) if fl[i] == 0), CORR["n_steps"])
    assert not_converged <= max(4, B // 100000), "too many non-converged instances: %d" % not_converged

    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ranks are aligned by the barrier BEFORE the W warm-up steps: the GPUs idle while waiting at a barrier, and the
    # warm-up steps then bring them back to steady state before the first timed step
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        graph.replay()
    sampler.start()
    t_wall = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        graph.replay()
    ev1.record()
    t_enqueue = time.perf_counter() - t_wall
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    # per-kernel durations: instrumented pass of the same step right after the timed region (CUDA events around each
    # launch on the launching stream; a captured graph cannot carry timing events)
    n_ins = max(3, min(args.steps, 50))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * n_ins + 1)]
    ev[0].record()
    for i in range(n_ins):
        yc, y, flags, status2 = ops.evopf_project(st, z, CORR["lr"], CORR["momentum"], CORR["corr_eps"], CORR["n_steps"],
                                                  tol=NEWTON["tol"], max_iters=NEWTON["max_iters"], want_y=True,
                                                  want_flags=True)
        ev[3 * i + 1].record()
        dz, _ = ops.evopf_complete_bwd(y, g, need_dx=False)
        ev[3 * i + 2].record()
        ev[3 * i + 3].record()
    torch.cuda.synchronize()
    proj_ms = sorted(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(n_ins))[n_ins // 2]
    bwd_ms = sorted(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(n_ins))[n_ins // 2]
    if os.environ.get("RPO_BENCH_DEBUG"):
        print("[rank %d] total %.3f ms/step  project %.3f  bwd %.3f  wall %.3f ms/step  host enqueue %.3f ms/step" %
              (rank, total_ms / args.steps, proj_ms, bwd_ms, t_wall / args.steps * 1e3, t_enqueue / args.steps * 1e3),
              file=sys.stderr, flush=True)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * B * args.steps / (total_ms_max * 1e-3)

    # FP32 roofline denominator measured live on this GPU (MEASURED_PEAKS.json has no FP32 figure)
    blocks, threads, iters = 148 * 8, 256, 8192
    for _ in range(2):
        ops.fp32_peak_probe(dev, blocks, threads, iters)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fl_probe = 0.0
    for _ in range(5):
        f, _ = ops.fp32_peak_probe(dev, blocks, threads, iters)
        fl_probe += f
    e1.record()
    torch.cuda.synchronize()
    fp32_peak = fl_probe / (e0.elapsed_time(e1) * 1e-3) / 1e12

    # end-to-end through the host-buffer API: pinned host inputs -> H2D -> kernels -> D2H results, every step.
    # Headline layout: the caller holds the state as the compact [pd|qd|soc] matrix (B,33) — the 33 columns the path
    # reads; the reference's full observation rows (B,57: + the 24-float price window the projection never reads) are
    # timed beside it (`full_rows`).  Both are H2D-link bound (bytes below / time = the link rate actually achieved).
    e2e = None
    if not args.no_e2e:
        bound = False
        if world > 1 and os.environ.get("RPO_BIND_NUMA", "1") != "0":
            from rpo_b200.host_api import bind_host_thread_to_device
            bound = bind_host_thread_to_device(dev)           # pinned buffers below land in the GPU's NUMA node
        hp = HostProjector(dev, chunk=1 << 17, n_streams=4, corr_lr=CORR["lr"], momentum=CORR["momentum"],
                           corr_eps=CORR["corr_eps"], n_steps=CORR["n_steps"], **NEWTON)
        pz, pg = z_h.pin_memory(), g_h.pin_memory()
        yc_out, dz_out = hp.pinned(B, 43), hp.pinned(B, 14)
        e2e_steps = max(3, min(args.steps, 10))

        def e2e_run(pst):
            for _ in range(2):
                hp.project_backward(pst, pz, pg, yc_out, dz_out)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                hp.project_backward(pst, pz, pg, yc_out, dz_out)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            sec = float(dt.item()) / e2e_steps
            return {"value": world * B / sec, "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes, "d2h_bytes_per_step": hp.d2h_bytes,
                    "h2d_gb_per_s_per_gpu": hp.h2d_bytes / sec / 1e9, "d2h_gb_per_s_per_gpu": hp.d2h_bytes / sec / 1e9,
                    "max_abs_diff_vs_resident": float((yc_out.to(dev) - out[0]).abs().max())}

        full = e2e_run(st_h.pin_memory())
        e2e = e2e_run(st_h[:, :33].contiguous().pin_memory())
        e2e.update({"steps": e2e_steps, "chunk": hp.chunk, "streams": len(hp.streams), "host_thread_bound_to_gpu_numa": bound,
                    "host_layout": "state as compact (B,33) [pd|qd|soc], z (B,14), dL/dy (B,43) in; y_corr (B,43), dL/dz (B,14) out",
                    "full_rows": dict(full, host_layout="state as the reference's (B,57) observation rows")})

    # BASELINE configs[0] / configs[1] beside the headline: every rank runs its own 2^24-instance shard (no collective)
    other = None
    if not args.no_other_configs:
        other = classic_configs(dev)
        for cfg in other.values():
            t = torch.tensor([cfg["ms_per_step"]], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cfg["ms_per_step"] = float(t.item())
            cfg["value"] = world * cfg["instances_per_gpu"] / (cfg["ms_per_step"] * 1e-3)
            cfg["hbm_frac"] = cfg["instances_per_gpu"] * cfg["bytes_per_instance"] / (cfg["ms_per_step"] * 1e-3) / 1e9 / 6552.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    f_proj = flops_project(n_newton, CORR["n_steps"])
    dense_achieved = B * f_proj / (proj_ms * 1e-3) / 1e12
    # flops the kernel EXECUTES: the fixed sparsity of the network is exploited (static count of the generated code by
    # tools/gen_evopf.py, FMA = 2).  This is the efficiency figure: `frac` = executed FP32 flop/s over the FP32 peak
    # measured in this run.  SURVEY §8(d)'s dense-minimal model (dense 22x22 LU, dense Jacobian) is kept beside it as
    # `dense_model`: the kernel does that model's work in 6x fewer flops, so its ratio to the peak exceeds 1.
    gs = json.load(open(os.path.join(ROOT, "rpo_b200", "csrc", "evopf_gen_stats.json")))
    fx = gs["newton_iter_flop"] * n_newton + gs["corr_step_flop"] * CORR["n_steps"] + gs["tail_flop"]
    achieved = B * fx / (proj_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_evopf_project_dram_bytes_per_launch")
        except Exception:
            traffic = None
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(ppath):
        peaks = json.load(open(ppath))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "fp32", "kernel": "k_evopf_project", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved / fp32_peak, "traffic": traffic,
        "flops_counted": "executed FP32 flops of the generated sparse code (rpo_b200/csrc/evopf_gen_stats.json): "
                         "%d x n_newton + %d x K + %d per instance" % (gs["newton_iter_flop"], gs["corr_step_flop"], gs["tail_flop"]),
        "flops_per_instance": fx,
        "peak_source": "rpo_fp32_peak_probe measured in this run (dependency-free FFMA micro-kernel, 8 chains x 256 threads x "
                       "8 CTAs/SM); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5; MEASURED_PEAKS.json carries no FP32 figure",
        "mean_newton_iters": n_newton, "correction_steps": CORR["n_steps"],
        "kernel_ms": proj_ms, "bwd_kernel_ms": bwd_ms,
        "kernel_ms_source": "median over an instrumented pass (CUDA events around each launch) right after the timed "
                            "region; the timed region replays a CUDA graph of the step (memset + 2 kernels)",
        "dense_model": {"flops_per_instance": f_proj, "achieved": dense_achieved, "frac_dense_model": dense_achieved / fp32_peak,
                        "note": "SURVEY §8(d) algorithmic count 12.9k n_N + 1.8k + 17.3k K (dense LU / dense Jacobian); not an "
                                "efficiency figure: the kernel eliminates on the fixed sparsity pattern"},
        "bwd_kernel": {"bound": "hbm", "achieved": B * (344.0 + 56.0) / (bwd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": B * 400.0 / (bwd_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_instance": 400},
        "hbm": {"achieved": value / world * BYTES_PER_ACTION / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": value / world * BYTES_PER_ACTION / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
    }
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = 65536
        cpu_baseline_sample(1024, 1024, threads, args.dist)
        v, dt, kind, what = cpu_baseline_sample(n, 2048, threads, args.dist)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": "%d of the %d instances (same generator), chunks of 2048, %.1f s; %s" % (n, B, dt, what)}
        if kind == "reference":                               # the oracle port beside it (it is the faster of the two)
            cpu_sample(1024, 1024, threads)
            vp, dtp = cpu_sample(16384, 2048, threads)
            cpu_baseline["oracle_port"] = {"value": vp, "sample": "16384 instances, %.1f s" % dtp}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B, instances_per_gpu=B, reference_loop_steps=k_ref, non_converged_instances=not_converged,
                                  l2="inputs larger than L2 (state %d MB + z %d MB per step)" % (B * 57 * 4 >> 20, B * 14 * 4 >> 20),
                                  parallelism="instance shards, no collective",
                                  timed_region="K replays of a CUDA graph of one step, CUDA events at both ends, max over ranks"),
        "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps * world,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "wall_s_timed_region": t_wall,
    }
    if other is not None:
        line["other_configs"] = other
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def classic_configs(dev):
    """BASELINE.json configs[0] / configs[1] on the GPU (reported beside the headline, same metric): closed-form completion +
    10 correction steps fused (`rpo_*_project`) + completion backward, 2^24 instances resident, reference settings
    (scripts/cart_exp.py, scripts/pen_exp.py: corr_lr 2e-2 / 2e-3), CUDA events over 20 steps after 3 warm-up steps."""
    import torch
    from rpo_b200 import ops, synth
    n = 1 << 24
    out = {}

    def timed(step):
        for _ in range(3):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    g = synth.upstream_grad(n, 2, seed=5, device=dev)
    _, z = synth.cart_inputs(n, seed=1, device=dev)

    def cart_step():
        ops.cart_project(z, 1, 2e-2, 0.0, 1e-5, 10, want_a=True)
        ops.cart_complete_bwd(g, 1)
    ms = timed(cart_step)
    out["safe_cartpole"] = {"workload": "CartSafe-v0 completion+correction (fwd+bwd), 2^24 instances per GPU", "value": n / (ms * 1e-3),
                            "unit": UNIT, "ms_per_step": ms, "bytes_per_instance": 40, "instances_per_gpu": n}
    obs, zp = synth.pen_inputs(n, seed=2, device=dev)

    def pen_step():
        ops.pen_project(obs, zp, 2e-3, 0.0, 1e-5, 10, want_a=True)
        ops.pen_complete_bwd(obs, g)
    ms = timed(pen_step)
    out["spring_pendulum"] = {"workload": "SpringPendulum-v0 completion+correction (fwd+bwd), 2^24 instances per GPU", "value": n / (ms * 1e-3),
                              "unit": UNIT, "ms_per_step": ms, "bytes_per_instance": 88, "instances_per_gpu": n}
    return out


if __name__ == "__main__":
    main()
