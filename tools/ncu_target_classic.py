"""A few launches of the cart / pendulum K-step kernels at 2^24 instances for `ncu -k regex:k_cart|k_pen` captures."""
import sys
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth
n = 1 << 24
_, z = synth.cart_inputs(n, seed=1, device='cuda')
obs, zp = synth.pen_inputs(n, seed=2, device='cuda')
for _ in range(3):
    ops.cart_project(z, 1, 2e-2, 0.0, 1e-5, 10, want_a=True)
    ops.pen_project(obs, zp, 2e-3, 0.0, 1e-5, 10, want_a=True)
torch.cuda.synchronize()
