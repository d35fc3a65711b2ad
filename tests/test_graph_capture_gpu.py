"""The entry points allocate nothing and never synchronise, so a step can be captured in a CUDA graph (bench.py's timed
region and the training harness rely on it): replays must reproduce the eager results bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_project_backward_and_box_replay_equal_eager():
    from rpo_b200 import evopf_constants as K
    from rpo_b200 import ops, synth
    B = 1 << 17                                           # large enough for the persistent backward kernel
    st = synth.evopf_states(B, seed=31).cuda()
    g = synth.upstream_grad(B, 43, seed=33).cuda()
    pre = torch.randn(B, 14, generator=torch.Generator().manual_seed(32)).cuda()
    lo = torch.tensor(K.ACTION_LOW[K.PARTIAL_ACTIONS], dtype=torch.float32, device="cuda")
    hi = torch.tensor(K.ACTION_HIGH[K.PARTIAL_ACTIONS], dtype=torch.float32, device="cuda")

    def step():
        z = ops.box_fwd(pre, lo, hi, ops.BOX_TANH, st, 5)
        yc, y, flags, _ = ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10, want_y=True)
        dz, _ = ops.evopf_complete_bwd(y, g, need_dx=False)
        return yc, dz, flags

    eager = [t.clone() for t in step()]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                            # warm-up on the capture stream (one-time attribute calls)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(3):
        for t in out:
            t.zero_()
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip(eager, out):
            assert torch.equal(a, b)
    # new inputs through the same graph: the captured pointers are the tensors' storage
    pre.mul_(-0.5)
    graph.replay()
    torch.cuda.synchronize()
    fresh = step()
    for a, b in zip(fresh, out):
        assert torch.equal(a, b)
