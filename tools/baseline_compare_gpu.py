"""Times the north-star-literal mapping (one warp per instance, dense 22x22 LU with partial pivoting in shared memory,
warp-shuffle reductions: tools/baseline/warp_lu_kernel.cu) against the product's completion kernel on identical inputs."""
import ctypes, os, subprocess, sys
sys.path.insert(0, '.')
import numpy as np, torch
from rpo_b200 import ops, synth
from rpo_b200 import evopf_constants as K
here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline")
so = os.path.join(here, "libwarp_baseline.so")
if not os.path.isfile(so):
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                           "-shared", "-o", so, os.path.join(here, "warp_lu_kernel.cu"), "-lcudart"])
lib = ctypes.CDLL(so)
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
nonslack = sorted(list(K.PV) + list(K.PQ))
eq_kind = np.array([0] * 13 + [1] * 9, np.int32); eq_bus = np.array(nonslack + list(K.PQ), np.int32)
G = K.G.astype(np.float32).copy(); Bm = K.B.astype(np.float32).copy()
spv = K.SPV.astype(np.int32).copy(); vm0 = K.VM_INIT.astype(np.float32).copy(); va0 = K.VA_INIT.astype(np.float32).copy()
torch.cuda.init(); torch.zeros(1, device='cuda')
assert lib.warp_baseline_init(P(G), P(Bm), P(eq_kind), P(eq_bus), P(eq_kind), P(eq_bus), P(spv), P(vm0), P(va0)) == 0
def timeit(f, n=5, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    return sorted(ts)[len(ts) // 2]
for lb in (16, 20):
    B = 1 << lb
    st = synth.evopf_states(B, seed=1000).cuda(); z = synth.evopf_partial_actions(st, seed=2000).cuda()
    y_w = torch.zeros(B, 43, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    run_w = lambda: lib.warp_baseline_complete(ctypes.c_void_p(st.data_ptr()), 57, ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(y_w.data_ptr()), ctypes.c_longlong(B), ctypes.c_float(1e-5), 50, ctypes.c_void_p(s))
    assert run_w() == 0
    y, status = ops.evopf_complete_fwd(st, z, return_status=True)
    torch.cuda.synchronize()
    ok = (status >> 8) == 0
    rel = ((y_w - y).abs().max(1).values / y.abs().max(1).values)[ok]
    t_w = timeit(run_w); t_p = timeit(lambda: ops.evopf_complete_fwd(st, z))
    print("B=2^%d  warp-per-instance dense LU (smem, shuffles): %.3f ms   thread-per-instance sparse static LU: %.3f ms   ratio %.1fx   max rel diff of y %.2e"
          % (lb, t_w, t_p, t_w / t_p, float(rel.max())))
