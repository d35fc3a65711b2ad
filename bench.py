#!/usr/bin/env python
"""Benchmark of the RPO projection hot path (BASELINE.json metric): projected actions / second.

One step = one pass of the hot path over one synthetic EVOPF-v0 batch of 2^20 instances per GPU:
  fused completion (Newton power flow) + 10 post-correction steps   (rpo_evopf_project, 1 kernel)
  + GRG backward of the completion                                  (rpo_evopf_complete_bwd, 1 kernel).
Instances shard across ranks with no collective (weak scaling: 2^20 instances per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU.)
"""
from __future__ import annotations

import argparse
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


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path (oracle port — the reference is Python, nothing to compile;
    /root/reference does not exist on the GPU box) on all host threads, same metric/config, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_step = 2048
    for _ in range(max(args.warmup, 1)):
        cpu_sample(n_step, 1024, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(n_step, 1024, threads)
    dt = time.perf_counter() - t0
    val = n_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic EVOPF-v0 completion+correction (fwd+bwd), batch 2^%d instances per GPU" % args.log2_batch,
                   "reference_sample": "each step times %d instances of that workload on the host CPU (bounded sample)" % n_step,
                   "input_dist": "D1 (uniform in the state-dependent partial box)", "newton": NEWTON, "correction": CORR,
                   "device": "cpu"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d instances per step, chunks of 1024, torch CPU oracle port of the reference path" % n_step},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-batch", type=int, default=LOG2_B)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
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

    B = 1 << args.log2_batch
    # each rank owns its own contiguous shard of the global batch (seeded per rank); no data-path collective
    st_h = synth.evopf_states(B, seed=1000 + rank)
    z_h = synth.evopf_partial_actions(st_h, seed=2000 + rank)
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
    _, status = ops.evopf_complete_fwd(st, z, return_status=True)
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

    # end-to-end through the host-buffer API: pinned host inputs -> H2D -> kernels -> D2H results, every step
    e2e = None
    if not args.no_e2e:
        bound = False
        if world > 1 and os.environ.get("RPO_BIND_NUMA", "1") != "0":
            from rpo_b200.host_api import bind_host_thread_to_device
            bound = bind_host_thread_to_device(dev)           # pinned buffers below land in the GPU's NUMA node
        hp = HostProjector(dev, chunk=1 << 17, n_streams=4, corr_lr=CORR["lr"], momentum=CORR["momentum"],
                           corr_eps=CORR["corr_eps"], n_steps=CORR["n_steps"], **NEWTON)
        pst, pz, pg = st_h.pin_memory(), z_h.pin_memory(), g_h.pin_memory()
        yc_out, dz_out = hp.pinned(B, 43), hp.pinned(B, 14)
        e2e_steps = max(3, min(args.steps, 10))
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
        e2e = {"value": world * B * e2e_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes,
               "d2h_bytes_per_step": hp.d2h_bytes, "steps": e2e_steps, "chunk": hp.chunk, "streams": len(hp.streams),
               "host_thread_bound_to_gpu_numa": bound,
               "max_abs_diff_vs_resident": float((yc_out.to(dev) - out[0]).abs().max())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    f_proj = flops_project(n_newton, CORR["n_steps"])
    achieved = B * f_proj / (proj_ms * 1e-3) / 1e12
    # flops the kernel actually executes: the fixed sparsity of the network is exploited (tools/gen_evopf.py counts them)
    executed = None
    spath = os.path.join(ROOT, "rpo_b200", "csrc", "evopf_gen_stats.json")
    if os.path.isfile(spath):
        gs = json.load(open(spath))
        fx = gs["newton_iter_flop"] * n_newton + gs["corr_step_flop"] * CORR["n_steps"] + gs["tail_flop"]
        executed = {"flops_per_instance": fx, "achieved": B * fx / (proj_ms * 1e-3) / 1e12,
                    "frac": B * fx / (proj_ms * 1e-3) / 1e12 / fp32_peak,
                    "note": "static FP32 op count of the generated sparse code (FMA = 2); the SURVEY figure above assumes dense 22x22 LU"}
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
        "peak_source": "rpo_fp32_peak_probe measured in this run (dependency-free FFMA micro-kernel)",
        "algorithmic_flops_per_instance": f_proj, "mean_newton_iters": n_newton, "correction_steps": CORR["n_steps"],
        "kernel_ms": proj_ms, "bwd_kernel_ms": bwd_ms, "executed": executed,
        "kernel_ms_source": "median over an instrumented pass (CUDA events around each launch) right after the timed "
                            "region; the timed region replays a CUDA graph of the step (memset + 2 kernels)",
        "whole_step": {"achieved": B * (f_proj + FLOPS_BWD) / ((proj_ms + bwd_ms) * 1e-3) / 1e12,
                       "frac": B * (f_proj + FLOPS_BWD) / ((proj_ms + bwd_ms) * 1e-3) / 1e12 / fp32_peak},
        "hbm": {"achieved": value / world * BYTES_PER_ACTION / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": value / world * BYTES_PER_ACTION / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
    }
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = 131072
        cpu_sample(1024, 1024, threads)
        v, dt = cpu_sample(n, 2048, threads)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d of the %d instances (same generator), chunks of 2048, %.1f s; torch-CPU oracle port of "
                                  "complete_partial + grad_steps(10) + PFFunction.backward" % (n, B, dt)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic EVOPF-v0 completion+correction (fwd+bwd), batch 2^%d instances per GPU" % args.log2_batch,
                   "instances_per_gpu": B, "input_dist": "D1 (uniform in the state-dependent partial box)",
                   "newton": NEWTON, "correction": CORR, "reference_loop_steps": k_ref, "non_converged_instances": not_converged,
                   "l2": "inputs larger than L2 (state %d MB + z %d MB per step)" % (B * 57 * 4 >> 20, B * 14 * 4 >> 20),
                   "parallelism": "instance shards, no collective",
                   "timed_region": "K replays of a CUDA graph of one step, CUDA events at both ends, max over ranks"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps * world,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "wall_s_timed_region": t_wall,
    }
    if world == 1:
        line["other_configs"] = classic_configs(dev)
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
    out["safe_cartpole"] = {"workload": "CartSafe-v0 completion+correction (fwd+bwd), 2^24 instances", "value": n / (ms * 1e-3),
                            "unit": UNIT, "ms_per_step": ms, "bytes_per_instance": 40}
    obs, zp = synth.pen_inputs(n, seed=2, device=dev)

    def pen_step():
        ops.pen_project(obs, zp, 2e-3, 0.0, 1e-5, 10, want_a=True)
        ops.pen_complete_bwd(obs, g)
    ms = timed(pen_step)
    out["spring_pendulum"] = {"workload": "SpringPendulum-v0 completion+correction (fwd+bwd), 2^24 instances", "value": n / (ms * 1e-3),
                              "unit": UNIT, "ms_per_step": ms, "bytes_per_instance": 88}
    return out


if __name__ == "__main__":
    main()
