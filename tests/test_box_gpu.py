"""Action-box kernels (rpo_box_fwd / rpo_box_bwd) on the GPU against vectors of the UNMODIFIED reference
BoxConstraint + EVOPFEnv.update (tests/golden/box.npz): bit-exact for the affine / clip / bound modes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "box.npz"))


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def box():
    from rpo_b200.box import BoxConstraint
    return BoxConstraint(G["low"][3].copy(), G["high"][3].copy(), device="cuda", volatile=True)


def test_golden_bit_exact(box):
    st = _dev(G["state"])
    x = _dev(G["call_in"]).requires_grad_(True)
    z = box(x, st)
    z.backward(_dev(G["up"]))
    assert np.array_equal(z.detach().cpu().numpy(), G["call"]) and np.array_equal(x.grad.cpu().numpy(), G["call_grad"])
    r = _dev(G["raw"]).requires_grad_(True)
    c = box.clip(r, st)
    c.backward(_dev(G["up"]))
    assert np.array_equal(c.detach().cpu().numpy(), G["clip"]) and np.array_equal(r.grad.cpu().numpy(), G["clip_grad"])
    assert np.array_equal(box.action_scale(st).cpu().numpy(), G["scale"])
    box.update_box(st)
    assert np.array_equal(box.cmin_vol, G["low"]) and np.array_equal(box.cmax_vol, G["high"])


def test_fused_tanh_and_styles(box):
    from rpo_b200.box import BoxConstraint
    st = _dev(G["state"])
    p = _dev(G["pre"]).requires_grad_(True)
    z = box.squash(p, st)
    z.backward(_dev(G["up"]))
    np.testing.assert_allclose(z.detach().cpu().numpy(), G["tanh_call"], rtol=0, atol=3e-7)      # tanhf: 1-2 ulp
    np.testing.assert_allclose(p.grad.cpu().numpy(), G["tanh_call_grad"], rtol=3e-6, atol=2e-7)
    sig = BoxConstraint(G["low"][3].copy(), G["high"][3].copy(), device="cuda", style="sigmoid", volatile=True)
    assert np.array_equal(sig(_dev(G["u01"]), st).cpu().numpy(), G["sigmoid_call"])
    full = BoxConstraint(G["full_low"], G["full_high"], device="cuda")
    assert np.array_equal(full(_dev(G["x43"])).cpu().numpy(), G["full_call"])
    assert np.array_equal(full.clip(_dev(3 * G["x43"])).cpu().numpy(), G["full_clip"])


@pytest.mark.parametrize("B", [1, 3, 255, 4097])
def test_ragged_and_unaligned_match_torch(box, B):
    """Batch sizes whose element count is not a multiple of four, and views that are not 16-byte aligned, take the
    scalar path; every path must equal the torch expression of the reference (utils.py:40-51,92-104)."""
    from rpo_b200 import ops, synth
    st = synth.evopf_states(B, seed=5, device="cuda")
    lo, hi = synth.evopf_partial_box(st)
    g = torch.Generator(device="cuda").manual_seed(B)
    x = torch.rand(B, 14, generator=g, device="cuda") * 2 - 1
    scale = (hi - lo) / 2
    want = scale * x + (lo + scale)
    assert torch.equal(box(x, st), want)
    buf = torch.zeros(B * 14 + 1, device="cuda")
    xv = buf[1:].view(B, 14)                                   # 4-byte aligned only
    xv.copy_(x)
    assert torch.equal(ops.box_fwd(xv, box.cmin_torch, box.cmax_torch, ops.BOX_AFFINE, st, 5), want)
    raw = lo + (hi - lo) * (torch.rand(B, 14, generator=g, device="cuda") * 1.6 - 0.3)
    assert torch.equal(box.clip(raw, st), torch.minimum(torch.maximum(raw, lo), hi))
    # in place
    y = x.clone()
    ops.box_fwd(y, box.cmin_torch, box.cmax_torch, ops.BOX_AFFINE, st, 5, out=y)
    assert torch.equal(y, want)


def test_full_size_inside_box(box):
    from rpo_b200 import synth
    B = 1 << 20
    st = synth.evopf_states(B, seed=9, device="cuda")
    lo, hi = synth.evopf_partial_box(st)
    z = box.squash(torch.randn(B, 14, device="cuda") * 3, st)
    # scale * (+-1) + (lo + scale) may land one rounding step outside the box, in the reference as well
    assert bool((z >= lo - 2e-7).all()) and bool((z <= hi + 2e-7).all())
    c = box.clip(z + 0.05 * torch.randn_like(z), st)
    assert bool((c >= lo).all()) and bool((c <= hi).all())
    assert torch.equal(box.clip(c, st), c)                      # idempotent


def test_argument_errors(box):
    from rpo_b200 import _lib, ops
    st = torch.zeros(4, 57, device="cuda")
    with pytest.raises(ValueError):
        box(torch.zeros(4, 14, device="cuda"))
    with pytest.raises(_lib.RpoLibraryError):
        ops.box_fwd(torch.zeros(4, 14, device="cuda"), box.cmin_torch, box.cmax_torch, 9, st, 5)
    with pytest.raises(_lib.RpoLibraryError):
        ops.box_fwd(torch.zeros(4, 14, device="cuda"), box.cmin_torch, box.cmax_torch, 0, st, 3)
    with pytest.raises(_lib.RpoLibraryError):
        ops.box_fwd(torch.zeros(4, 14), box.cmin_torch, box.cmax_torch, 0, st, 5)


def test_box_fused_into_projection_and_backward_is_bit_identical_to_the_separate_launches():
    """rpo_evopf_project_squashed / _complete_fwd_squashed / _complete_bwd_squashed (SURVEY f-1 fused into a1/a5/a6) against
    rpo_box_fwd -> rpo_evopf_project, rpo_box_fwd -> rpo_evopf_complete_fwd and rpo_evopf_complete_bwd(_penalty) ->
    rpo_box_bwd: same bits, on full TMA tiles, a ragged tail and a state passed as the compact (B,33) view."""
    from rpo_b200 import evopf_constants as K, ops, synth
    dev = DEV
    lo = torch.tensor(K.ACTION_LOW[K.PARTIAL_ACTIONS], dtype=torch.float32, device=dev)
    hi = torch.tensor(K.ACTION_HIGH[K.PARTIAL_ACTIONS], dtype=torch.float32, device=dev)
    for B in (3, 128 * 5 + 77, 1 << 16):
        st = synth.evopf_states(B, seed=80).to(dev)
        g = torch.Generator().manual_seed(81)
        u = (1.5 * torch.randn(B, 14, generator=g)).to(dev)
        gy = synth.upstream_grad(B, 43, seed=82).to(dev)
        w = torch.rand(58, generator=g).to(dev)
        rs = torch.rand(B, generator=g).to(dev)
        for mode, x in ((ops.BOX_TANH, u), (ops.BOX_AFFINE, torch.tanh(u)), (ops.BOX_CLIP, 0.3 * u)):
            z = ops.box_fwd(x, lo, hi, mode, st, 5)
            yc_s, y_s, fl_s, _ = ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10, want_y=True)
            yc_f, y_f, fl_f, _ = ops.evopf_project_squashed(st, x, lo, hi, mode, 1e-4, 0.0, 1e-5, 10, want_y=True)
            assert torch.equal(yc_f, yc_s) and torch.equal(y_f, y_s) and torch.equal(fl_f, fl_s)
            yr_s = ops.evopf_complete_fwd(st, z)
            yr_f = ops.evopf_complete_fwd_squashed(st, x, lo, hi, mode)
            assert torch.equal(yr_f, yr_s)
            assert torch.equal(ops.evopf_complete_fwd_squashed(st[:, :33], x, lo, hi, mode), yr_s)
            dz, dx = ops.evopf_complete_bwd(yr_s, gy)
            du_s = ops.box_bwd(x, dz, lo, hi, mode, st, 5)
            du_f, dx_f = ops.evopf_complete_bwd_squashed(st, yr_s, gy, x, lo, hi, mode, need_dx=True)
            assert torch.equal(du_f, du_s) and torch.equal(dx_f, dx)
            dzp, _ = ops.evopf_complete_bwd_penalty(st, yr_s, gy, w, rs, need_dx=False)
            dup_s = ops.box_bwd(x, dzp, lo, hi, mode, st, 5)
            dup_f, _ = ops.evopf_complete_bwd_squashed(st, yr_s, gy, x, lo, hi, mode, w, rs)
            assert torch.equal(dup_f, dup_s)


def test_fused_squashed_completion_autograd_matches_the_unfused_graph():
    """rpo_b200.algo.complete_partial_squashed (one forward + one backward launch for box, completion and penalty) gives the
    same actor-loss gradients as BoxConstraint.squash -> env.complete_partial -> nju(env.ineq_dist)."""
    from rpo_b200 import synth
    from rpo_b200.algo import complete_partial_squashed
    from rpo_b200.box import BoxConstraint
    from rpo_b200.env import EVOPFEnv
    dev = torch.device(DEV)
    env = EVOPFEnv(device=dev, demand_curve=np.ones(24), price_curve=np.ones(24))
    box = BoxConstraint(*env.box_constraint_partial, device=dev, volatile=True, update=env.update)
    B = 512
    st = synth.evopf_states(B, seed=90).to(dev)
    gen = torch.Generator().manual_seed(91)
    u0 = torch.randn(B, 14, generator=gen).to(dev)
    q_w = torch.randn(43, 1, generator=gen).to(dev)
    w0 = torch.rand(1, 58, generator=gen).to(dev)
    u1, w1 = u0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    a1 = env.complete_partial(st, box.squash(u1, st))
    (-(a1 @ q_w) + torch.nn.functional.linear(env.ineq_dist(st, a1), w1)).mean().backward()
    u2, w2 = u0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    a2, pen = complete_partial_squashed(env, box, st, u2, w2)
    (-(a2 @ q_w) + pen).mean().backward()
    assert torch.equal(a2, a1)
    assert float((u2.grad - u1.grad).abs().max()) <= 1e-6 * float(u1.grad.abs().max())
    assert float((w2.grad - w1.grad).abs().max()) <= 1e-5 * float(w1.grad.abs().max())
