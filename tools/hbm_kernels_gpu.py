"""Achieved HBM bandwidth of the memory-bound kernels (classic envs, env steps, resid); bytes are algorithmic."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
B = 1 << 24
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
    ts.sort()
    return ts[len(ts) // 2]
def rep(name, ms, bytes_per_inst, B):
    print('%-28s %8.3f ms  %7.1f GB/s  (%d B/inst, B=2^%d)' % (name, ms, B * bytes_per_inst / ms / 1e6, bytes_per_inst, B.bit_length() - 1))
st, z = synth.cart_inputs(B, seed=1); st, z = st.cuda(), (z * 1.3).cuda()
a = ops.cart_complete_fwd(z, 1); g = torch.randn(B, 2, device='cuda')
rep('cart_complete_fwd', timeit(lambda: ops.cart_complete_fwd(z, 1)), 12, B)
rep('cart_complete_bwd', timeit(lambda: ops.cart_complete_bwd(g, 1)), 12, B)
rep('cart_correct(10)', timeit(lambda: ops.cart_correct(a, 1, 2e-2, 0.0, 1e-5, 10)), 16, B)
rep('cart_project(10)', timeit(lambda: ops.cart_project(z, 1, 2e-2, 0.0, 1e-5, 10, want_a=True)), 20, B)
rep('cart_resid', timeit(lambda: ops.cart_resid(a)), 8 + 28, B)
rep('cart_step', timeit(lambda: ops.cart_step(st, a)), 24 + 8 + 24 + 4 + 28, B)
obs, zp = synth.pen_inputs(B, seed=2); obs, zp = obs.cuda(), (zp * 1.5).cuda()
ap = ops.pen_complete_fwd(obs, zp)
rep('pen_complete_fwd', timeit(lambda: ops.pen_complete_fwd(obs, zp)), 32, B)
rep('pen_complete_bwd', timeit(lambda: ops.pen_complete_bwd(obs, g)), 32, B)
rep('pen_correct(10)', timeit(lambda: ops.pen_correct(obs, ap, 2e-3, 0.0, 1e-5, 10)), 36, B)
rep('pen_project(10)', timeit(lambda: ops.pen_project(obs, zp, 2e-3, 0.0, 1e-5, 10, want_a=True)), 40, B)
st4 = torch.stack([torch.atan2(obs[:, 1], obs[:, 0]), obs[:, 2], obs[:, 3], obs[:, 4]], 1).contiguous()
rep('pen_step', timeit(lambda: ops.pen_step(st4, ap)), 16 + 8 + 16 + 20 + 4 + 4 + 8, B)
del st, z, a, g, obs, zp, ap, st4
B2 = 1 << 21
s = synth.evopf_states(B2, seed=3).cuda(); zz = synth.evopf_partial_actions(s, seed=4).cuda()
y = ops.evopf_complete_fwd(s, zz)
exo = torch.cat([s[:, :28], s[:, 33:]], 1).contiguous()
rep('evopf_resid', timeit(lambda: ops.evopf_resid(s, y)), 132 + 172 + 112 + 232, B2)
rep('evopf_step (+viol)', timeit(lambda: ops.evopf_step(s, y, exo)), 228 + 172 + 208 + 228 + 4 + 112 + 232, B2)
rep('evopf_step (no viol)', timeit(lambda: ops.evopf_step(s, y, exo, want_viol=False)), 228 + 172 + 208 + 228 + 4, B2)
from rpo_b200.box import BoxConstraint
from rpo_b200 import evopf_constants as K
bx = BoxConstraint(K.ACTION_LOW[K.PARTIAL_ACTIONS], K.ACTION_HIGH[K.PARTIAL_ACTIONS], device='cuda', volatile=True)
pre = torch.randn(B2, 14, device='cuda'); up = torch.randn(B2, 14, device='cuda')
rep('box squash (tanh, volatile)', timeit(lambda: ops.box_fwd(pre, bx.cmin_torch, bx.cmax_torch, ops.BOX_TANH, s, 5)), 56 + 56 + 20, B2)
rep('box clip (volatile)', timeit(lambda: ops.box_fwd(pre, bx.cmin_torch, bx.cmax_torch, ops.BOX_CLIP, s, 5)), 56 + 56 + 20, B2)
rep('box squash bwd', timeit(lambda: ops.box_bwd(pre, up, bx.cmin_torch, bx.cmax_torch, ops.BOX_TANH, s, 5)), 56 * 3 + 20, B2)
def torch_box():
    lo, hi = synth.evopf_partial_box(s); sc = (hi - lo) / 2
    return sc * torch.tanh(pre) + (lo + sc)
rep('  same in torch ops (device)', timeit(torch_box), 56 + 56 + 20, B2)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
src = torch.empty(1 << 28, dtype=torch.float32, device='cuda'); dst = torch.empty_like(src)
for _ in range(3): dst.copy_(src)
e0.record(); [dst.copy_(src) for _ in range(5)]; e1.record(); torch.cuda.synchronize()
print('torch copy_ reference %.1f GB/s' % (5 * 2 * src.numel() * 4 / e0.elapsed_time(e1) / 1e6))
