"""Cost of the per-step loop-test flag stores in the cart / pendulum K-step kernels."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
B = 1 << 24
def timeit(f, n=30):
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
_, z = synth.cart_inputs(B, seed=1, device='cuda')
obs, zp = synth.pen_inputs(B, seed=2, device='cuda')
for wf in (True, False, True, False, False, True):
    print('flags', wf, 'cart_project %.3f ms' % timeit(lambda: ops.cart_project(z, 1, 2e-2, 0.0, 1e-5, 10, want_a=True, want_flags=wf)),
          'pen_project %.3f ms' % timeit(lambda: ops.pen_project(obs, zp, 2e-3, 0.0, 1e-5, 10, want_a=True, want_flags=wf)))
r = ops.cart_project(z, 1, 2e-2, 0.0, 1e-5, 10, want_a=True); print('cart flags', r[2].tolist(), 'checksum %.6f' % float(r[0].double().sum()))
r = ops.pen_project(obs, zp, 2e-3, 0.0, 1e-5, 10, want_a=True); print('pen flags', r[2].tolist(), 'checksum %.6f' % float(r[0].double().sum()))
