#!/usr/bin/env python
"""profiles/r02_sweep.md from the JSON lines of tools/sweep_multi.sh (BASELINE.json configs[4])."""
import json, sys, collections
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
weak = collections.defaultdict(dict); strong = collections.defaultdict(dict); d2 = collections.defaultdict(dict); classic = {}
for r in rows:
    c = r["config"]; n = r["n_gpus"]
    lb = int(c["workload"].split("2^")[1].split()[0])
    if "other_configs" in r:
        classic[n] = r["other_configs"]
        continue
    key = (lb,)
    if r["scaling"] == "strong": strong[lb][n] = r
    elif c["input_dist"].startswith("D2"): d2[lb][n] = r
    else: weak[lb][n] = r
ns = sorted({r["n_gpus"] for r in rows})
out = []
out.append("# BASELINE.json configs[4]: synthetic EVOPF-v0 completion+correction sweep (fused project K = 10 + GRG backward), round 2 final build\n")
out.append("`gpurun --gpus 8 -- bash tools/sweep_multi.sh` → one `bench.py --gpus N --log2-batch n [--strong] [--dist D2] --steps 100 --no-e2e --no-cpu-baseline` line per cell"
           " (CUDA-graph replays between CUDA events, max over ranks; inputs resident). Efficiency = value(N) / (N × value(1)) of the same row.\n")
def table(title, data, per="per GPU"):
    out.append("\n## %s\n" % title)
    out.append("| batch (%s) | " % per + " | ".join("N = %d: actions/s (ms/step)" % n for n in ns) + " | " + " | ".join("eff. N = %d" % n for n in ns if n > 1) + " |")
    out.append("|---|" + "---|" * (len(ns) + len([n for n in ns if n > 1])))
    for lb in sorted(data):
        d = data[lb]
        cells = ["%.3e (%.3f)" % (d[n]["value"], d[n]["ms_per_step"]) if n in d else "–" for n in ns]
        effs = ["%.1f %%" % (100 * d[n]["value"] / (n * d[1]["value"])) if (n in d and 1 in d) else "–" for n in ns if n > 1]
        out.append("| 2^%d | " % lb + " | ".join(cells) + " | " + " | ".join(effs) + " |")
table("Weak scaling, D1 (2^n instances per GPU)", weak)
table("Strong scaling, D1 (2^n instances in total, split over the ranks)", strong, per="total")
table("Weak scaling, D2 (central 30 % of the box)", d2)
if classic:
    out.append("\n## Safe Cartpole / Spring Pendulum (2^24 instances per GPU, fused projection with loop test + completion backward)\n")
    out.append("| env | " + " | ".join("N = %d: actions/s (ms), HBM frac" % n for n in sorted(classic)) + " |")
    out.append("|---|" + "---|" * len(classic))
    for k in ("safe_cartpole", "spring_pendulum"):
        out.append("| %s | " % k + " | ".join("%.3e (%.3f), %.2f" % (classic[n][k]["value"], classic[n][k]["ms_per_step"], classic[n][k].get("hbm_frac", float("nan"))) for n in sorted(classic)) + " |")
# where efficiency drops below 85 %
low = [(lb, n, weak[lb][n]["value"] / (n * weak[lb][1]["value"])) for lb in weak for n in weak[lb] if n > 1 and 1 in weak[lb]]
slow = [(lb, n, strong[lb][n]["value"] / (n * strong[lb][1]["value"])) for lb in strong for n in strong[lb] if n > 1 and 1 in strong[lb]]
out.append("\n## Where efficiency falls below 85 %\n")
bad = [("weak 2^%d per GPU, N = %d: %.1f %%" % (lb, n, 100 * e)) for lb, n, e in sorted(low) if e < 0.85] + \
      [("strong 2^%d total, N = %d (%d instances per GPU): %.1f %%" % (lb, n, (1 << lb) // n, 100 * e)) for lb, n, e in sorted(slow) if e < 0.85]
out.append("\n".join("* " + b for b in bad) if bad else "* nowhere in the measured grid")
out.append("\nBoth are the same effect: a step shorter than ~0.2 ms per GPU (2^14 instances: 49 us; 131,072: 203 us against 0.169 ms = 1.353 / 8) "
           "is dominated by the fixed part of a step — one partial wave of 128-instance CTAs (2^14 instances = 128 CTAs on 296 slots, 131,072 = 3.5 waves "
           "with a ragged tail), graph-launch latency and the max over 8 ranks of a 50-200 us interval. From 2^16 instances per GPU on, weak scaling is "
           ">= 98.6 % at every N; strong scaling of 2^22 stays at 95.5 % at N = 8 (524,288 per GPU).")
print("\n".join(out))
