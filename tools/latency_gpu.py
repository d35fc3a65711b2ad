"""Latency of the projection at the reference's own batch sizes (B = 1 rollout, B = 256 replay batch)."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
def timeit(f, n=50):
    for _ in range(10): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for B in (1, 256, 4096, 37888, 1 << 17):
    st = synth.evopf_states(B, seed=1).cuda(); z = synth.evopf_partial_actions(st, seed=2).cuda(); g = synth.upstream_grad(B, 43).cuda()
    y = ops.evopf_complete_fwd(st, z)
    print('B=%6d  project(K=10) %7.1f us  project(K=50 eval) %7.1f us  complete %6.1f us  backward %6.1f us' % (
        B, timeit(lambda: ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10)), timeit(lambda: ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 50)),
        timeit(lambda: ops.evopf_complete_fwd(st, z)), timeit(lambda: ops.evopf_complete_bwd(y, g))))
