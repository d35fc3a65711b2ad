"""GRG backward at 2^20: plain / + penalty / + box / both (CUDA events, median of 5 x 20 launches)."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import evopf_constants as K, ops, synth
B = 1 << 20
st = synth.evopf_states(B, seed=1000).cuda(); z = synth.evopf_partial_actions(st, seed=2000).cuda(); g = synth.upstream_grad(B, 43).cuda()
y = ops.evopf_complete_fwd(st, z, refine=False)
w = torch.rand(58, device='cuda'); rs = torch.rand(B, device='cuda'); u = torch.randn(B, 14, device='cuda')
lo = torch.tensor(K.ACTION_LOW[K.PARTIAL_ACTIONS], dtype=torch.float32, device='cuda'); hi = torch.tensor(K.ACTION_HIGH[K.PARTIAL_ACTIONS], dtype=torch.float32, device='cuda')
def timeit(f, n=20, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / n)
    return sorted(ts)[len(ts) // 2]
print("bwd plain %.4f ms | + penalty %.4f | + box %.4f | + penalty + box %.4f | resid_bwd(ineq_dist) %.4f" % (
    timeit(lambda: ops.evopf_complete_bwd(y, g, need_dx=False)),
    timeit(lambda: ops.evopf_complete_bwd_penalty(st, y, g, w, rs, need_dx=False)),
    timeit(lambda: ops.evopf_complete_bwd_squashed(st, y, g, u, lo, hi, ops.BOX_TANH)),
    timeit(lambda: ops.evopf_complete_bwd_squashed(st, y, g, u, lo, hi, ops.BOX_TANH, w, rs)),
    timeit(lambda: ops.evopf_resid_bwd(st, y, g_ineq=torch.ones(1, device='cuda').expand(B, 58).contiguous(), ineq_mode=1), n=5)))
