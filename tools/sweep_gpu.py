"""BASELINE.json configs[4]: synthetic EVOPF completion+correction sweep, batch 2^14 .. 2^22, both input distributions.
One step = k_evopf_project (K = 10) + k_evopf_bwd; CUDA events, median of 5 x 10 launches, inputs resident."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
def timeit(f, n=10, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    return sorted(ts)[len(ts) // 2]
print('| log2 B | dist | mean Newton iters | non-converged | ms / step | projected actions / s |')
print('|---|---|---|---|---|---|')
for dist in ('D1', 'D2'):
    for lb in range(14, 23):
        B = 1 << lb
        st = synth.evopf_states(B, seed=7).cuda(); z = synth.evopf_partial_actions(st, seed=8, dist=dist).cuda(); g = synth.upstream_grad(B, 43).cuda()
        _, status = ops.evopf_complete_fwd(st, z, return_status=True, refine=False)
        def step():
            yc, y, fl, _ = ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10, want_y=True)
            ops.evopf_complete_bwd(y, g, need_dx=False)
        ms = timeit(step)
        print('| %d | %s | %.2f | %d | %.3f | %.3e |' % (lb, dist, float((status & 0xff).float().mean()), int(((status >> 8) != 0).sum()), ms, B / ms * 1e3))
        del st, z, g
