"""End-to-end (host buffers) throughput of HostProjector against chunk size / stream count."""
import sys, time
sys.path.insert(0, '.')
import torch
from rpo_b200 import synth
from rpo_b200.host_api import HostProjector
B = 1 << 20
st = synth.evopf_states(B, seed=1); z = synth.evopf_partial_actions(st, seed=2); g = synth.upstream_grad(B, 43)
pst, pz, pg = st.pin_memory(), z.pin_memory(), g.pin_memory()
for chunk, ns in ((1 << 17, 4), (1 << 16, 4), (1 << 15, 4), (1 << 15, 8), (1 << 14, 8), (1 << 16, 2)):
    hp = HostProjector('cuda:0', chunk=chunk, n_streams=ns)
    yc, dz = hp.pinned(B, 43), hp.pinned(B, 14)
    for _ in range(2):
        hp.project_backward(pst, pz, pg, yc, dz)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        hp.project_backward(pst, pz, pg, yc, dz)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print('chunk 2^%d streams %d: %.3f ms  %.4g /s  H2D %.1f GB/s' % (chunk.bit_length() - 1, ns, dt * 1e3, B / dt, hp.h2d_bytes / dt / 1e9))
    del hp
