import sys, time
sys.path.insert(0, '.')
import torch
from rpo_b200 import ops, synth, _lib
print('arch', _lib.load().rpo_build_arch(), torch.cuda.get_device_name(0))
B = 1 << 20
st = synth.evopf_states(B, seed=1000).cuda(); z = synth.evopf_partial_actions(st, seed=2000).cuda(); g = synth.upstream_grad(B, 43).cuda()
def timeit(f, n=10, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)/n)
    ts.sort()
    return ts[len(ts)//2]
y, status = ops.evopf_complete_fwd(st, z, return_status=True, refine=False)
print('iters mean', float((status & 0xff).float().mean()), 'flagged', int(((status>>8)!=0).sum()))
t_f = timeit(lambda: ops.evopf_complete_fwd(st, z))
t_c = timeit(lambda: ops.evopf_correct(st, y, 1e-4, 0.0, 1e-5, 10))
t_b = timeit(lambda: ops.evopf_complete_bwd(y, g, need_dx=False))
t_p = timeit(lambda: ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10))
t_py = timeit(lambda: ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10, want_y=True))
t_d = timeit(lambda: ops.evopf_direction(st, y))
t_r = timeit(lambda: ops.evopf_resid(st, y))
print('B %d  fwd %.3f  corr %.3f  bwd %.3f  project %.3f  project+y %.3f  direction %.3f resid %.3f ms' % (B, t_f, t_c, t_b, t_p, t_py, t_d, t_r))
t_nf = timeit(lambda: ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10, want_flags=False))
print('project without flag buffer %.3f ms (with: %.3f)' % (t_nf, t_p))
yc = ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10)[0]
print('projected actions/s (project+y + bwd): %.4g  checksum %.10f' % (B/((t_py+t_b)*1e-3), float(yc.double().sum())))
