"""Action box without a GPU: the element code the kernel runs (box_core.cuh, compiled by g++ in
tests/hostsim) against vectors produced by the UNMODIFIED reference BoxConstraint + EVOPFEnv.update
(tests/golden/box.npz, oracle/make_golden.py box), and the drop-in class driven through the host simulator."""
import os

import numpy as np
import pytest
import torch

from tests import hostsim_ops as hs

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "box.npz"))
LO = np.concatenate([G["low"][0, :9], np.zeros(5, np.float32)])     # constant columns; the last five are volatile
HI = np.concatenate([G["high"][0, :9], np.zeros(5, np.float32)])


def test_bounds_bit_exact():
    st = torch.from_numpy(G["state"])
    lo = hs.box_fwd(None, LO, HI, 5, st, 5).numpy()
    hi = hs.box_fwd(None, LO, HI, 6, st, 5).numpy()
    assert np.array_equal(lo, G["low"]) and np.array_equal(hi, G["high"])
    assert np.array_equal(hs.box_fwd(None, LO, HI, 4, st, 5).numpy(), G["scale"])


def test_affine_and_clip_bit_exact():
    st = torch.from_numpy(G["state"])
    assert np.array_equal(hs.box_fwd(G["call_in"], LO, HI, 0, st, 5).numpy(), G["call"])
    assert np.array_equal(hs.box_bwd(G["call_in"], G["up"], LO, HI, 0, st, 5).numpy(), G["call_grad"])
    assert np.array_equal(hs.box_fwd(G["raw"], LO, HI, 2, st, 5).numpy(), G["clip"])
    assert np.array_equal(hs.box_bwd(G["raw"], G["up"], LO, HI, 2, st, 5).numpy(), G["clip_grad"])
    assert np.array_equal(hs.box_fwd(G["u01"], LO, HI, 3, st, 5).numpy(), G["sigmoid_call"])
    assert np.array_equal(hs.box_fwd(G["x43"], G["full_low"], G["full_high"], 0).numpy(), G["full_call"])
    assert np.array_equal(hs.box_fwd(3 * G["x43"], G["full_low"], G["full_high"], 2).numpy(), G["full_clip"])


def test_fused_tanh():
    st = torch.from_numpy(G["state"])
    z = hs.box_fwd(G["pre"], LO, HI, 1, st, 5).numpy()
    np.testing.assert_allclose(z, G["tanh_call"], rtol=0, atol=2e-7)        # libm tanhf vs torch.tanh: <= 1 ulp of a value <= 1.4
    gx = hs.box_bwd(G["pre"], G["up"], LO, HI, 1, st, 5).numpy()
    np.testing.assert_allclose(gx, G["tanh_call_grad"], rtol=2e-6, atol=1e-7)


def test_dropin_class_through_hostsim(monkeypatch):
    hs.install(monkeypatch)
    from rpo_b200.box import BoxConstraint
    cmin, cmax = G["low"][3].copy(), G["high"][3].copy()     # row 3: soc = 0.2 -> any static part is the same for all rows
    bx = BoxConstraint(cmin, cmax, device="cpu", volatile=True, update=None)
    st = torch.from_numpy(G["state"])
    x = torch.from_numpy(G["call_in"]).requires_grad_(True)
    z = bx(x, st)
    z.backward(torch.from_numpy(G["up"]))
    assert np.array_equal(z.detach().numpy(), G["call"]) and np.array_equal(x.grad.numpy(), G["call_grad"])
    r = torch.from_numpy(G["raw"]).requires_grad_(True)
    c = bx.clip(r, st)
    c.backward(torch.from_numpy(G["up"]))
    assert np.array_equal(c.detach().numpy(), G["clip"]) and np.array_equal(r.grad.numpy(), G["clip_grad"])
    assert np.array_equal(bx.action_scale(st).numpy(), G["scale"])
    p = torch.from_numpy(G["pre"])
    np.testing.assert_allclose(bx.squash(p, st).numpy(), G["tanh_call"], atol=2e-7, rtol=0)
    bx.update_box(st)
    assert np.array_equal(bx.cmin_vol, G["low"]) and np.array_equal(bx.cmax_vol, G["high"])
    # numpy in -> numpy out, single row (rollout call shape: state (1,S))
    z1 = bx(G["call_in"][5], G["state"][5:6])
    assert isinstance(z1, np.ndarray) and np.array_equal(z1.reshape(-1), G["call"][5])
    s = bx.sample(st[:1])
    assert s.shape[-1] == 14 and bool((s >= torch.from_numpy(G["low"][:1]) - 1e-6).all()) and bool((s <= torch.from_numpy(G["high"][:1]) + 1e-6).all())
    # static box
    fb = BoxConstraint(G["full_low"], G["full_high"], device="cpu")
    assert np.array_equal(fb(torch.from_numpy(G["x43"])).numpy(), G["full_call"])
    assert np.array_equal(fb.clip(torch.from_numpy(3 * G["x43"])).numpy(), G["full_clip"])
    with pytest.raises(ValueError):
        bx(torch.from_numpy(G["call_in"]))                   # volatile box without a state
