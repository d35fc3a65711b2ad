"""A few launches of the secondary kernels (GRG backward, action box, residuals) for `ncu -k regex:… -c 1` captures."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth, evopf_constants as K
B = 1 << 20
st = synth.evopf_states(B, seed=1000).cuda(); z = synth.evopf_partial_actions(st, seed=2000).cuda(); g = synth.upstream_grad(B, 43).cuda()
y = ops.evopf_complete_fwd(st, z)
lo = torch.tensor(K.ACTION_LOW[K.PARTIAL_ACTIONS], dtype=torch.float32, device='cuda')
hi = torch.tensor(K.ACTION_HIGH[K.PARTIAL_ACTIONS], dtype=torch.float32, device='cuda')
pre = torch.randn(B, 14, device='cuda')
for _ in range(3):
    ops.evopf_complete_bwd(y, g, need_dx=False)
    ops.box_fwd(pre, lo, hi, ops.BOX_TANH, st, 5)
    ops.evopf_resid(st, y)
torch.cuda.synchronize()
