"""Small invocation of every kernel, run under compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
for B in (1, 300):
    st = synth.evopf_states(B, seed=1).cuda(); z = synth.evopf_partial_actions(st, seed=2).cuda(); g = synth.upstream_grad(B, 43).cuda()
    yc, y, fl, status = ops.evopf_project(st, z, 1e-4, 0.5, 1e-5, 3, want_y=True, want_status=True)
    yc, y, fl, status = ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 3, want_y=True, want_status=True)
    y2 = ops.evopf_complete_fwd(st[:, :28], z)
    ops.evopf_complete_bwd(y, g); ops.evopf_direction(st, y); ops.evopf_correct(st, y, 1e-4, 0.0, 1e-5, 2); ops.evopf_resid(st, y)
    exo = torch.cat([st[:, :28], st[:, 33:]], 1).contiguous(); ops.evopf_step(st, y, exo)
    s6, zc = synth.cart_inputs(B, seed=3); s6, zc = s6.cuda(), zc.cuda()
    ac, a, _ = ops.cart_project(zc, 1, 2e-2, 0.0, 1e-5, 3, want_a=True); ops.cart_complete_bwd(a, 0); ops.cart_direction(a, 1); ops.cart_resid(a); ops.cart_step(s6, a); ops.cart_correct(a, 0, 2e-2, 0.5, 1e-5, 2)
    ob, zp = synth.pen_inputs(B, seed=4); ob, zp = ob.cuda(), zp.cuda()
    ap_c, ap, _ = ops.pen_project(ob, zp, 2e-3, 0.0, 1e-5, 3, want_a=True); ops.pen_complete_bwd(ob, ap); ops.pen_direction(ob, ap); ops.pen_resid(ob, ap); ops.pen_correct(ob, ap, 2e-3, 0.5, 1e-5, 2)
    st4 = torch.rand(B, 4, device='cuda') * 0.1 + torch.tensor([0, 0, 1.0, 0], device='cuda'); ops.pen_step(st4, ap)
# persistent double-buffered backward kernel (run with RPO_BWD_PIPE=force: 10 tiles on 4 CTAs would need a big batch
# otherwise) + a ragged tail, and the action-box kernels (row kernel, flat kernel, unaligned view)
import os
B = 128 * 10 + 37
st = synth.evopf_states(B, seed=5).cuda(); z = synth.evopf_partial_actions(st, seed=6).cuda(); g = synth.upstream_grad(B, 43).cuda()
y = ops.evopf_complete_fwd(st, z)
dz, dx = ops.evopf_complete_bwd(y, g)
# (initcheck does not model cp.async.bulk stores as initialising writes: buffers written by the TMA path are round-tripped
# through the host before an LSU kernel reads them, and results are compared on the host)
dz1, dx1 = ops.evopf_complete_bwd(y[:100].cpu().cuda(), g[:100])
assert torch.equal(dz.cpu()[:100], dz1.cpu()) and torch.equal(dx.cpu()[:100], dx1.cpu()), 'pipelined backward differs from the one-pass kernel'
from rpo_b200 import evopf_constants as K
lo = torch.tensor(K.ACTION_LOW[K.PARTIAL_ACTIONS], dtype=torch.float32, device='cuda'); hi = torch.tensor(K.ACTION_HIGH[K.PARTIAL_ACTIONS], dtype=torch.float32, device='cuda')
x = torch.randn(B, 14, device='cuda')
for mode in range(7):
    ops.box_fwd(x if mode < 4 else None, lo, hi, mode, st, 5, B=B)
    if mode < 4:
        ops.box_bwd(x, g[:, :14].contiguous(), lo, hi, mode, st, 5)
buf = torch.zeros(B * 14 + 1, device='cuda'); xv = buf[1:].view(B, 14); xv.copy_(x)
ops.box_fwd(xv, lo, hi, 1, st, 5); ops.box_fwd(torch.randn(B, 43, device='cuda'), torch.zeros(43, device='cuda'), torch.ones(43, device='cuda'), 0)
torch.cuda.synchronize()
print('sanitizer smoke done, RPO_BWD_PIPE =', os.environ.get('RPO_BWD_PIPE'))
