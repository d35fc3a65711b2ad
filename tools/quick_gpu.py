import sys, time
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth, _lib
print('arch', _lib.load().rpo_build_arch(), torch.cuda.get_device_name(0))
B = 1<<20
st = synth.evopf_states(B, seed=0).cuda(); z = synth.evopf_partial_actions(st, seed=1).cuda(); g = synth.upstream_grad(B, 43).cuda()
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
t_f = timeit(lambda: ops.evopf_complete_fwd(st, z))
y, status = ops.evopf_complete_fwd(st, z, return_status=True)
print('iters hist', torch.bincount((status & 0xff).long()).tolist(), 'flags', int((status>>8).max()))
t_c = timeit(lambda: ops.evopf_correct(st, y, 1e-4, 0.0, 1e-5, 10))
t_b = timeit(lambda: ops.evopf_complete_bwd(y, g))
t_p = timeit(lambda: ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10))
print('B', B, 'fwd ms', t_f, 'corr ms', t_c, 'bwd ms', t_b, 'project ms', t_p)
print('projected actions/s (project+bwd):', B/((t_p+t_b)*1e-3))
flops, _ = ops.fp32_peak_probe('cuda', 148*8, 256, 4096)
t = timeit(lambda: ops.fp32_peak_probe('cuda', 148*8, 256, 4096))
print('fp32 peak TFLOP/s', flops/(t*1e-3)/1e12)
