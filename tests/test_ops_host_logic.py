"""Host-side logic of rpo_b200.ops (shapes, leading dimensions, optional outputs, batch-size checks) without a GPU: the
library is replaced by a stub that accepts every call, so only the Python around the C ABI runs."""
import contextlib

import pytest
import torch


class _StubLib(object):
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*args):
            self.calls.append((name, args))
            return 0
        return f


@pytest.fixture()
def ops(monkeypatch):
    from rpo_b200 import _lib
    from rpo_b200 import ops as o
    stub = _StubLib()
    monkeypatch.setattr(_lib, "load", lambda: stub)
    monkeypatch.setattr(o, "_is_device_tensor", lambda t: isinstance(t, torch.Tensor))
    monkeypatch.setattr(o, "_stream", lambda t: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    o.stub = stub
    return o


def test_every_wrapper_accepts_consistent_shapes(ops):
    B = 5
    st, z, y, g = torch.zeros(B, 57), torch.zeros(B, 14), torch.zeros(B, 43), torch.zeros(B, 43)
    assert ops.evopf_complete_fwd(st, z).shape == (B, 43)
    y2, status = ops.evopf_complete_fwd(st[:, :28], z, return_status=True)
    assert status.shape == (B,) and ops.stub.calls[-1][1][1] == 57          # a column slice keeps the row stride as ld
    dz, dx = ops.evopf_complete_bwd(y, g)
    assert dz.shape == (B, 14) and dx.shape == (B, 28) and ops.evopf_complete_bwd(y, g, need_dx=False)[1] is None
    assert ops.evopf_direction(st, y).shape == (B, 43)
    out, flags = ops.evopf_correct(st, y, 1e-4, 0.0, 1e-5, 10)
    assert out.shape == (B, 43) and flags.shape == (10,) and ops.evopf_correct(st, y, 1e-4, 0.0, 1e-5, 10, want_flags=False)[1] is None
    yc, yy, fl, stt = ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 10, want_y=True, want_status=True)
    assert yc.shape == yy.shape == (B, 43) and fl.shape == (10,) and stt.shape == (B,)
    eq, iq = ops.evopf_resid(st, y)
    assert eq.shape == (B, 28) and iq.shape == (B, 58)
    ns, rew, eqv, iqv = ops.evopf_step(st, y, torch.zeros(B, 52))
    assert ns.shape == (B, 57) and rew.shape == (B,) and eqv.shape == (B, 28) and iqv.shape == (B, 58)
    # single rows given as 1-D tensors
    assert ops.evopf_complete_fwd(st[0], z[0]).shape == (1, 43)
    # classic envs
    zc, a, s6 = torch.zeros(B, 1), torch.zeros(B, 2), torch.zeros(B, 6)
    assert ops.cart_complete_fwd(zc, 1).shape == (B, 2) and ops.cart_complete_bwd(a, 0).shape == (B, 1)
    assert ops.cart_direction(a, 1).shape == (B, 2) and ops.cart_correct(a, 1, 2e-2, 0.0, 1e-5, 3)[1].shape == (3,)
    ac, aa, fl = ops.cart_project(zc, 0, 2e-2, 0.0, 1e-5, 3, want_a=True)
    assert ac.shape == aa.shape == (B, 2)
    assert ops.cart_resid(a)[1].shape == (B, 6) and ops.cart_step(s6, a)[0].shape == (B, 6)
    obs, st4 = torch.zeros(B, 5), torch.zeros(B, 4)
    assert ops.pen_complete_fwd(obs, zc).shape == (B, 2) and ops.pen_complete_bwd(obs, a).shape == (B, 1)
    assert ops.pen_direction(obs, a).shape == (B, 2) and ops.pen_correct(obs, a, 2e-3, 0.0, 1e-5, 3)[0].shape == (B, 2)
    assert ops.pen_project(obs, zc, 2e-3, 0.0, 1e-5, 3)[0].shape == (B, 2) and ops.pen_resid(obs, a)[0].shape == (B, 1)
    assert ops.pen_step(st4, a)[1].shape == (B, 5)
    # action box
    lo, hi = torch.zeros(14), torch.ones(14)
    assert ops.box_fwd(z, lo, hi, ops.BOX_TANH, st, 5).shape == (B, 14)
    assert ops.box_fwd(None, lo, hi, ops.BOX_LOW, st, 5).shape == (B, 14)
    assert ops.box_bwd(z, z, lo, hi, ops.BOX_CLIP, st, 5).shape == (B, 14)
    assert ops.box_fwd(torch.zeros(B, 2), torch.zeros(2), torch.ones(2)).shape == (B, 2)


def test_batch_size_mismatches_are_rejected_before_the_launch(ops):
    st, z, y = torch.zeros(4, 57), torch.zeros(5, 14), torch.zeros(5, 43)
    n0 = len(ops.stub.calls)
    for call in (lambda: ops.evopf_complete_fwd(st, z), lambda: ops.evopf_direction(st, y),
                 lambda: ops.evopf_correct(st, y, 1e-4, 0.0, 1e-5, 2), lambda: ops.evopf_project(st, z, 1e-4, 0.0, 1e-5, 2),
                 lambda: ops.evopf_resid(st, y), lambda: ops.evopf_complete_bwd(y, torch.zeros(4, 43)),
                 lambda: ops.evopf_step(torch.zeros(5, 57), y, torch.zeros(4, 52)),
                 lambda: ops.cart_step(torch.zeros(4, 6), torch.zeros(5, 2)),
                 lambda: ops.pen_complete_fwd(torch.zeros(4, 5), torch.zeros(5, 1)),
                 lambda: ops.pen_correct(torch.zeros(4, 5), torch.zeros(5, 2), 2e-3, 0.0, 1e-5, 2),
                 lambda: ops.pen_step(torch.zeros(4, 4), torch.zeros(5, 2)),
                 lambda: ops.box_fwd(z, torch.zeros(14), torch.ones(14), 0, st, 5),
                 lambda: ops.box_bwd(z, torch.zeros(4, 14), torch.zeros(14), torch.ones(14), 0)):
        with pytest.raises(ValueError):
            call()
    assert len(ops.stub.calls) == n0
    with pytest.raises(ValueError):
        ops.evopf_complete_fwd(torch.zeros(5, 20), z)                  # too few state columns
    with pytest.raises(TypeError):
        ops.evopf_complete_fwd(torch.zeros(5, 57, dtype=torch.float64), z)


def test_column_offset_views_are_copied_before_a_bulk_row_copy(ops):
    """the kernels bulk-copy whole `ld`-float row blocks; a view like s[:, 24:57] of a (B,57) tensor starts 24
    floats into the allocation, so the block of the last rows would end 24 floats past it.  Such views must reach the
    kernel as a contiguous copy; views that start at column 0 keep their stride (no copy)."""
    B = 256
    st, y = torch.zeros(B, 57), torch.zeros(B, 43)
    ops.evopf_resid(st[:, :33], y)
    assert ops.stub.calls[-1][1][0] == st.data_ptr() and ops.stub.calls[-1][1][1] == 57        # prefix view: passed as is
    wide = torch.zeros(B, 90)
    ops.evopf_resid(wide[:, 24:57], y)
    name, args = ops.stub.calls[-1]
    assert args[1] == 33 and args[0] != wide.data_ptr() + 24 * 4                               # offset view: compact copy
    ops.evopf_resid(wide[:, 24:90], y)                                                         # same: a row block would overrun
    assert ops.stub.calls[-1][1][1] == 66
    ops.evopf_resid(wide[: B - 1, 24:60], y[: B - 1])                                          # a spare row behind it: no copy
    assert ops.stub.calls[-1][1][1] == 90
    # the new backward wrappers
    assert ops.evopf_resid_bwd(st, y, g_ineq=torch.zeros(B, 58), ineq_mode=1).shape == (B, 43)
    dz, dx = ops.evopf_complete_bwd_penalty(st, y, y, torch.zeros(58), torch.zeros(B))
    assert dz.shape == (B, 14) and dx.shape == (B, 28)
    assert ops.cart_resid_bwd(torch.zeros(B, 2), g_eq=torch.zeros(B, 1)).shape == (B, 2)
    assert ops.pen_resid_bwd(torch.zeros(B, 5), torch.zeros(B, 2), g_ineq=torch.zeros(B, 1), ineq_mode=1).shape == (B, 2)
    with pytest.raises(ValueError):
        ops.evopf_resid_bwd(st, y, g_eq=torch.zeros(B - 1, 28))
