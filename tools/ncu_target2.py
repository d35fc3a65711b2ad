"""One launch of each round-2 kernel at 2^20 instances for `ncu -k regex:... -c 1` captures (after 2 warm-up launches)."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
B = 1 << 20
st = synth.evopf_states(B, seed=1000).cuda(); z = synth.evopf_partial_actions(st, seed=2000).cuda(); g = synth.upstream_grad(B, 43).cuda()
w = torch.rand(58, device='cuda'); rs = torch.rand(B, device='cuda'); gi = torch.rand(B, 58, device='cuda')
for _ in range(3):
    yc, y, _, _ = ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10, want_y=True)
    yr = ops.evopf_complete_fwd(st, z)                      # refined
    yp = ops.evopf_complete_fwd(st, z, refine=False)
    ops.evopf_complete_bwd(y, g, need_dx=False)
    ops.evopf_complete_bwd_penalty(st, y, g, w, rs, need_dx=False)
    ops.evopf_resid_bwd(st, y, g_ineq=gi, ineq_mode=1)
torch.cuda.synchronize()
