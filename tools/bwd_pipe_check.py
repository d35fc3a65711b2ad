"""The persistent double-buffered backward kernel against the one-pass kernel (which small batches still take): bit-equal."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
for B in ((1 << 20) + 77, 1 << 18, (1 << 17) + 128 * 5):
    st = synth.evopf_states(B, seed=11).cuda(); z = synth.evopf_partial_actions(st, seed=12).cuda(); g = synth.upstream_grad(B, 43).cuda()
    y = ops.evopf_complete_fwd(st, z)
    for need_dx in (False, True):
        dz, dx = ops.evopf_complete_bwd(y, g, need_dx=need_dx)
        C = 1 << 14
        dz_c, dx_c = [], []
        for b0 in range(0, B, C):
            a, b = ops.evopf_complete_bwd(y[b0:b0 + C], g[b0:b0 + C], need_dx=need_dx)
            dz_c.append(a); dx_c.append(b)
        ok = torch.equal(dz, torch.cat(dz_c)) and (not need_dx or torch.equal(dx, torch.cat(dx_c)))
        print('B', B, 'dx', need_dx, 'bit-equal', ok, 'finite', bool(torch.isfinite(dz).all()))
        assert ok
print('PIPE CHECK OK')
