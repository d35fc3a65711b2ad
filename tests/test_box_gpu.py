"""Action-box kernels (rpo_box_fwd / rpo_box_bwd) on the GPU against vectors of the UNMODIFIED reference
BoxConstraint + EVOPFEnv.update (tests/golden/box.npz): bit-exact for the affine / clip / bound modes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

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
