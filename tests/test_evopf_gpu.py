"""GPU parity of the EVOPF kernels (through the C ABI) against the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest
import torch

from tests.parity import EQ_TOL, GRAD_RTOL, OUTLIER_FRAC, Y_RTOL, assert_close_or_marginal, rel_err

pytestmark = pytest.mark.gpu


def _inputs(B, dist, seed=0):
    from rpo_b200 import synth
    st = synth.evopf_states(B, seed=seed)
    z = synth.evopf_partial_actions(st, seed=seed + 1, dist=dist)
    return st, z


@pytest.mark.parametrize("B", [1, 7, 256, 4096])
@pytest.mark.parametrize("dist", ["D1", "D2"])
def test_complete_fwd_bwd(B, dist, evopf_oracle, evopf_oracle64):
    from rpo_b200 import ops, synth
    orc = evopf_oracle
    st, z = _inputs(B, dist)
    y, status = ops.evopf_complete_fwd(st.cuda(), z.cuda(), return_status=True)
    torch.cuda.synchronize()
    y = y.cpu()
    status = status.cpu().numpy()
    y_o, jac, inv, iters = orc.complete_partial_fwd(st, z, return_saved=True)
    assert rel_err(y, y_o).max() <= Y_RTOL
    assert (status >> 8).max() == 0, "non-converged / small-pivot flags set"
    assert (status & 0xff).max() <= 8
    # equality residual of OUR y (mixed-precision refined completion), evaluated in float64 (removes the checker's own
    # float32 noise): at or below the Newton tolerance on every row, and not above the float32 reference's on the same rows
    r = evopf_oracle64.eq_resid(st.double(), y.double()).abs().max(dim=1).values
    r_ref = evopf_oracle64.eq_resid(st.double(), y_o.double()).abs().max(dim=1).values
    assert r.max() <= EQ_TOL
    assert r.max() <= r_ref.max() + (2e-6 if B < 256 else 0.0)            # tiny batches: a maximum over a handful of rows
    if B >= 256:
        assert r.median() <= r_ref.median() and (r > EQ_TOL).double().mean() <= (r_ref > EQ_TOL).double().mean()
    # the plain float32 sequence (what the fused projection kernel runs before its correction steps): same noise class
    y_plain = ops.evopf_complete_fwd(st.cuda(), z.cuda(), refine=False).cpu()
    r_plain = evopf_oracle64.eq_resid(st.double(), y_plain.double()).abs().max(dim=1).values
    assert r_plain.max() <= 1.5 * EQ_TOL and rel_err(y_plain, y_o).max() <= Y_RTOL
    # GRG backward
    g = synth.upstream_grad(B, 43)
    dz, dx = ops.evopf_complete_bwd(y.cuda(), g.cuda())
    dx_o, dz_o = orc.complete_partial_bwd(g, jac, inv)
    assert rel_err(dz.cpu(), dz_o).max() <= GRAD_RTOL
    assert rel_err(dx.cpu(), dx_o).max() <= GRAD_RTOL


@pytest.mark.parametrize("B", [1, 7, 256, 4096])
@pytest.mark.parametrize("dist", ["D1", "D2"])
def test_direction_correct_project(B, dist, evopf_oracle):
    from oracle.rpo_oracle import grad_steps, grad_steps_margins
    from rpo_b200 import ops
    orc = evopf_oracle
    st, z = _inputs(B, dist, seed=10)
    y_o = orc.complete_partial_fwd(st, z)
    d = ops.evopf_direction(st.cuda(), y_o.cuda()).cpu()
    d_o = orc.ineq_partial_grad(st, y_o)
    e = rel_err(d, d_o)
    assert (e > GRAD_RTOL).double().mean() <= OUTLIER_FRAC + 1.0 / max(B, 1) * 0
    yc, flags = ops.evopf_correct(st.cuda(), y_o.cuda(), 1e-4, 0.0, 1e-5, 10)
    yc_o, k = grad_steps(orc, st, y_o, 1e-4, 0.0, 1e-5, 10)
    e = rel_err(yc.cpu(), yc_o)
    assert (e > Y_RTOL).double().mean() <= OUTLIER_FRAC
    fl = flags.cpu().numpy()
    steps = 10
    for i in range(1, 10):
        if fl[i] == 0:
            steps = i
            break
    assert steps == k
    # momentum variant
    ycm, _ = ops.evopf_correct(st.cuda(), y_o.cuda(), 1e-4, 0.5, 1e-5, 5)
    ycm_o, _ = grad_steps(orc, st, y_o, 1e-4, 0.5, 1e-5, 5)
    assert (rel_err(ycm.cpu(), ycm_o) > Y_RTOL).double().mean() <= OUTLIER_FRAC
    # fused projection == complete followed by correct (same kernels' math, bit-for-bit)
    yp, y1, fl2, status = ops.evopf_project(st.cuda(), z.cuda(), 1e-4, 0.0, 1e-5, 10, want_y=True, want_status=True)
    y_sep = ops.evopf_complete_fwd(st.cuda(), z.cuda(), refine=False)      # the fused kernel's own (plain float32) Newton sequence
    yc_sep, _ = ops.evopf_correct(st.cuda(), y_sep, 1e-4, 0.0, 1e-5, 10)
    assert torch.equal(y1, y_sep)
    assert torch.equal(yp, yc_sep)
    # the fused path corrects OUR completion (|y - y_oracle| ~ 5e-6): instances that graze a bound may flip a mask (H6)
    yc_o2, _, margin = grad_steps_margins(orc, st, orc.complete_partial_fwd(st, z), 1e-4, 0.0, 1e-5, 10)
    assert_close_or_marginal(rel_err(yp.cpu(), yc_o2), margin, Y_RTOL)


def test_resid_and_step(evopf_oracle):
    from rpo_b200 import ops, synth
    orc = evopf_oracle
    B = 1000
    st, z = _inputs(B, "D1", seed=20)
    y = orc.complete_partial_fwd(st, z)
    y = y + 0.01 * torch.randn(B, 43, generator=torch.Generator().manual_seed(0))
    eq, iq = ops.evopf_resid(st.cuda(), y.cuda())
    assert (eq.cpu() - orc.eq_resid(st, y)).abs().max() < 2e-5
    assert (iq.cpu() - orc.ineq_resid(st, y)).abs().max() < 1e-6
    nxt = synth.evopf_states(B, seed=21)
    exo = torch.cat([nxt[:, :28], nxt[:, 33:]], dim=1).contiguous()
    ns, rew, eqv, iqv = ops.evopf_step(st.cuda(), y.cuda(), exo.cuda())
    ns_o, rew_o = orc.step(st.numpy(), y.numpy(), exo.numpy())
    assert np.abs(ns.cpu().numpy() - ns_o).max() < 1e-6
    assert np.abs(rew.cpu().numpy() - rew_o).max() <= 1e-5 * max(1.0, np.abs(rew_o).max())
    assert (eqv.cpu() - orc.eq_resid(st, y)).abs().max() < 2e-5
    assert (iqv.cpu() - orc.ineq_dist(st, y)).abs().max() < 1e-6


def test_strided_state_and_empty(evopf_oracle):
    from rpo_b200 import ops
    st, z = _inputs(64, "D1", seed=30)
    y_full = ops.evopf_complete_fwd(st.cuda(), z.cuda())
    y_slice = ops.evopf_complete_fwd(st.cuda()[:, :28], z.cuda())     # non-contiguous view, ld = 57
    assert torch.equal(y_full, y_slice)
    y0 = ops.evopf_complete_fwd(st.cuda()[:0], z.cuda()[:0])
    assert y0.shape == (0, 43)
    with pytest.raises(Exception):
        ops.evopf_complete_fwd(st, z)            # CPU tensors: no CPU path, must fail loudly


def test_golden_setpoint():
    """case-14 set-point golden vector (reference output)."""
    import os
    from rpo_b200 import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "evopf_setpoint.npz"))
    st = torch.tensor(g["state"]).cuda()
    z = torch.tensor(g["z"]).cuda()
    y = ops.evopf_complete_fwd(st, z)
    assert rel_err(y.cpu(), torch.tensor(g["y"])).max() <= Y_RTOL
    dz, dx = ops.evopf_complete_bwd(y, torch.ones_like(y))
    assert rel_err(dz.cpu(), torch.tensor(g["dz_ones"])).max() <= GRAD_RTOL
    d = ops.evopf_direction(st, torch.tensor(g["y"]).cuda())
    assert rel_err(d.cpu(), torch.tensor(g["dir"])).max() <= GRAD_RTOL
    yc, _ = ops.evopf_correct(st, torch.tensor(g["y"]).cuda(), 1e-4, 0.0, 1e-5, 10)
    assert rel_err(yc.cpu(), torch.tensor(g["y_corr"])).max() <= Y_RTOL


def test_unaligned_views_take_the_fallback_path():
    """Pointers that are only 4-byte aligned (views into larger buffers) must not reach the TMA / float4 paths."""
    import ctypes
    from rpo_b200 import _lib, ops
    lib = _lib.load()
    B = 700
    st, z = _inputs(B, "D1", seed=44)
    st_d, z_d = st.cuda(), z.cuda()
    y_ref = ops.evopf_complete_fwd(st_d, z_d)
    # inputs and outputs placed at a 4-byte offset inside bigger allocations
    big_s = torch.zeros(B * 57 + 1, device="cuda"); big_s[1:] = st_d.reshape(-1)
    big_z = torch.zeros(B * 14 + 1, device="cuda"); big_z[1:] = z_d.reshape(-1)
    big_y = torch.zeros(B * 43 + 1, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    rc = lib.rpo_evopf_complete_fwd(big_s.data_ptr() + 4, 57, big_z.data_ptr() + 4, big_y.data_ptr() + 4, B, 1e-5, 50, None, s)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(big_y[1:].view(B, 43), y_ref)
    g = torch.randn(B, 43, device="cuda")
    big_g = torch.zeros(B * 43 + 1, device="cuda"); big_g[1:] = g.reshape(-1)
    big_dz = torch.zeros(B * 14 + 1, device="cuda")
    rc = lib.rpo_evopf_complete_bwd(big_y.data_ptr() + 4, big_g.data_ptr() + 4, big_dz.data_ptr() + 4, None, B, s)
    assert rc == 0
    torch.cuda.synchronize()
    dz_ref, _ = ops.evopf_complete_bwd(y_ref, g, need_dx=False)
    assert torch.equal(big_dz[1:].view(B, 14), dz_ref)
    big_yc = torch.zeros(B * 43 + 1, device="cuda")
    rc = lib.rpo_evopf_project(big_s.data_ptr() + 4, 57, big_z.data_ptr() + 4, None, big_yc.data_ptr() + 4, B, 1e-5, 50,
                               1e-4, 0.0, 1e-5, 10, None, None, s)
    assert rc == 0
    torch.cuda.synchronize()
    yc_ref, _, _, _ = ops.evopf_project(st_d, z_d, 1e-4, 0.0, 1e-5, 10)
    assert torch.equal(big_yc[1:].view(B, 43), yc_ref)


def test_step_and_resid_optional_outputs():
    """Nullable outputs of the C ABI (eq / ineq / violations) on the TMA path (full CTAs) and the ragged path."""
    from rpo_b200 import ops, synth
    for B in (128 * 5, 128 * 5 + 37):
        st, z = _inputs(B, "D1", seed=55)
        st_d = st.cuda()
        y = ops.evopf_complete_fwd(st_d, z.cuda())
        eq, iq = ops.evopf_resid(st_d, y)
        eq1, none1 = ops.evopf_resid(st_d, y, want_ineq=False)
        none2, iq2 = ops.evopf_resid(st_d, y, want_eq=False)
        assert none1 is None and none2 is None and torch.equal(eq, eq1) and torch.equal(iq, iq2)
        nxt = synth.evopf_states(B, seed=56)
        exo = torch.cat([nxt[:, :28], nxt[:, 33:]], dim=1).contiguous().cuda()
        ns, rew, eqv, iqv = ops.evopf_step(st_d, y, exo)
        ns2, rew2, e2, i2 = ops.evopf_step(st_d, y, exo, want_viol=False)
        assert e2 is None and i2 is None and torch.equal(ns, ns2) and torch.equal(rew, rew2)
        assert torch.equal(eqv, eq) and torch.equal(iqv, torch.clamp(iq, 0))


def test_small_pivot_flag_fires_on_a_collapsed_voltage_profile():
    """RPO_STATUS_SMALL_PIVOT (reference: torch.inverse would raise / return garbage on a singular Jacobian, evopf.py:831):
    generator voltages of 1e-9 pu make the flow-Jacobian rows of those buses O(V), so their static-order pivots fall below 1e-6."""
    from rpo_b200 import ops, synth
    st = synth.evopf_states(64, seed=70)
    z = synth.evopf_partial_actions(st, seed=71)
    z[::2, 4:9] = 1e-9                                          # vm_spv of every second instance
    y, status = ops.evopf_complete_fwd(st.cuda(), z.cuda(), return_status=True)
    status = status.cpu().numpy()
    assert ((status[::2] & 0x200) != 0).all(), status[:8]
    assert ((status[1::2] >> 8) == 0).all()                      # the untouched instances are unaffected (no cross-instance coupling)
    _, status_p = ops.evopf_complete_fwd(st.cuda(), z.cuda(), return_status=True, refine=False)
    assert ((status_p.cpu().numpy()[::2] & 0x200) != 0).all()
