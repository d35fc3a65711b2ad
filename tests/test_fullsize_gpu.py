"""BASELINE.json's full size (2^20 instances) through size-independent properties, plus oracle parity on a
strided subsample (the CPU oracle does ~3k instances/s, so it cannot check every row)."""
import numpy as np
import pytest
import torch

from tests.parity import EQ_TOL, GRAD_RTOL, Y_RTOL, assert_close_or_marginal, rel_err

pytestmark = pytest.mark.gpu
B = 1 << 20
LR, MOM, EPS, K = 1e-4, 0.0, 1e-5, 10


@pytest.fixture(scope="module")
def data():
    from rpo_b200 import ops, synth
    st = synth.evopf_states(B, seed=123)
    z = synth.evopf_partial_actions(st, seed=124)
    g = synth.upstream_grad(B, 43, seed=125)
    st_d, z_d, g_d = st.cuda(), z.cuda(), g.cuda()
    yc, y, flags, status = ops.evopf_project(st_d, z_d, LR, MOM, EPS, K, want_y=True, want_status=True)
    dz, dx = ops.evopf_complete_bwd(y, g_d)
    torch.cuda.synchronize()
    return dict(st=st, z=z, g=g, st_d=st_d, z_d=z_d, g_d=g_d, y=y, yc=yc, flags=flags, status=status, dz=dz, dx=dx)


def test_convergence_and_residual_at_full_size(data, evopf_oracle64):
    from rpo_b200 import ops
    status = data["status"]
    bad = (status >> 8) != 0
    n_bad = int(bad.sum())
    assert n_bad <= 8, "infeasible power flows are ~1e-6 of D1 draws; got %d" % n_bad
    ok = ~bad
    iters = (status & 0xff)[ok]
    assert int(iters.max()) <= 12 and float(iters.float().mean()) < 5.5
    assert bool(torch.isfinite(data["y"][ok]).all()) and bool(torch.isfinite(data["yc"][ok]).all())
    # equality residual of the completion, evaluated by our own residual kernel on every row ...
    eq, _ = ops.evopf_resid(data["st_d"], data["y"], want_ineq=False)
    r = eq.abs().max(dim=1).values[ok]
    assert float((r > EQ_TOL).float().mean()) <= 2e-3 and float(r.max()) <= 3 * EQ_TOL
    # ... and by the float64 oracle on a strided subsample (removes the float32 noise of the checker)
    idx = torch.arange(0, B, 257)[:4000]
    idx = idx[ok.cpu()[idx]]
    r64 = evopf_oracle64.eq_resid(data["st"][idx].double(), data["y"].cpu()[idx].double()).abs().max(dim=1).values
    assert float(r64.max()) <= 1.5 * EQ_TOL and float((r64 > EQ_TOL).double().mean()) <= 2e-3
    assert float(r64.median()) <= 4e-6
    assert data["flags"].cpu().tolist() == [1] * K                      # the reference would take all 10 steps


def test_refined_completion_residual_not_above_reference_at_full_size(data, evopf_oracle, evopf_oracle64):
    """north_star: "equality residual at or below the reference's Newton tolerance".  The stand-alone completion
    (rpo_evopf_complete_fwd: float32 Newton + float64-residual chord step) over all 2^20 instances; its equality residual,
    evaluated in float64, against the float32 reference's on the same 65,536 rows: maximum, median, and fraction above 1e-5."""
    from oracle.rpo_oracle import EvopfOracle
    from rpo_b200 import ops
    y, status = ops.evopf_complete_fwd(data["st_d"], data["z_d"], return_status=True)
    ok = ((status >> 8) == 0)
    assert int((~ok).sum()) <= 8
    # every row, by our own float32 residual kernel (its evaluation noise is ~2e-6 on top of the true residual)
    eq, _ = ops.evopf_resid(data["st_d"], y, want_ineq=False)
    r32 = eq.abs().max(dim=1).values[ok]
    assert float(r32.max()) <= 1.2 * EQ_TOL and float((r32 > EQ_TOL).float().mean()) <= 1e-5
    idx = torch.arange(0, B, 16)
    idx = idx[ok.cpu()[idx]]
    st, z = data["st"][idx], data["z"][idx]
    y_ref = evopf_oracle.complete_partial_fwd(st, z)
    chk32 = EvopfOracle(dtype=torch.float64, const_dtype=torch.float32)     # the reference's own (float32) network constants
    for chk in (chk32, evopf_oracle64):
        r = chk.eq_resid(st.double(), y.cpu()[idx].double()).abs().max(dim=1).values
        r_ref = chk.eq_resid(st.double(), y_ref.double()).abs().max(dim=1).values
        assert float(r.max()) <= float(r_ref.max()) and float(r.max()) <= EQ_TOL
        assert float(r.median()) <= float(r_ref.median())
        assert float((r > EQ_TOL).double().mean()) <= float((r_ref > EQ_TOL).double().mean())
        assert float(torch.quantile(r, 0.999)) <= float(torch.quantile(r_ref, 0.999))
    assert rel_err(y.cpu()[idx], y_ref).max() <= Y_RTOL
    # refined and plain completions agree to the float32 noise of the plain one
    assert rel_err(y.cpu()[idx], data["y"].cpu()[idx]).max() <= 2e-5


def test_oracle_parity_on_subsample(data, evopf_oracle):
    from oracle.rpo_oracle import grad_steps_margins
    orc = evopf_oracle
    idx = torch.arange(5, B, 1021)[:1024]
    keep = ((data["status"].cpu()[idx] >> 8) == 0)
    idx = idx[keep]
    st, z, g = data["st"][idx], data["z"][idx], data["g"][idx]
    y_o, jac, inv, _ = orc.complete_partial_fwd(st, z, return_saved=True)
    assert rel_err(data["y"].cpu()[idx], y_o).max() <= Y_RTOL
    dx_o, dz_o = orc.complete_partial_bwd(g, jac, inv)
    assert rel_err(data["dz"].cpu()[idx], dz_o).max() <= GRAD_RTOL
    assert rel_err(data["dx"].cpu()[idx], dx_o).max() <= GRAD_RTOL
    yc_o, k, margin = grad_steps_margins(orc, st, y_o, LR, MOM, EPS, K)
    assert k == K
    assert_close_or_marginal(rel_err(data["yc"].cpu()[idx], yc_o), margin, Y_RTOL)


def test_shard_equivalence_and_permutation(data):
    """No cross-instance coupling: any row range / any row order gives bit-identical rows (SURVEY §4 'Distributed')."""
    from rpo_b200 import ops
    from rpo_b200.dist import shard_rows
    parts = []
    for r in range(8):
        lo, hi = shard_rows(B, r, 8)
        yc, _, _, _ = ops.evopf_project(data["st_d"][lo:hi], data["z_d"][lo:hi], LR, MOM, EPS, K)
        parts.append(yc)
    conv = ((data["status"] >> 8) == 0)
    # the Newton vote is per CTA of 256 rows: shards start at multiples of 2^17, so CTAs see the same rows
    assert torch.equal(torch.cat(parts)[conv], data["yc"][conv])
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).cuda()
    dz_p, _ = ops.evopf_complete_bwd(data["y"][perm], data["g_d"][perm], need_dx=False)
    assert torch.equal(dz_p, data["dz"][perm])                           # backward has no vote at all: exact
    yc_p, _ = ops.evopf_correct(data["st_d"][perm], data["y"][perm], LR, MOM, EPS, K, want_flags=False)
    yc_s, _ = ops.evopf_correct(data["st_d"], data["y"], LR, MOM, EPS, K, want_flags=False)
    assert torch.equal(yc_p, yc_s[perm])


def test_fused_equals_separate_and_backward_linearity(data):
    from rpo_b200 import ops
    y_sep = ops.evopf_complete_fwd(data["st_d"], data["z_d"], refine=False)
    assert torch.equal(y_sep, data["y"])
    yc_sep, _ = ops.evopf_correct(data["st_d"], y_sep, LR, MOM, EPS, K, want_flags=False)
    assert torch.equal(yc_sep, data["yc"])
    # GRG backward is linear in dL/dy
    g2 = torch.randn(B, 43, generator=torch.Generator().manual_seed(9)).cuda()
    dz2, _ = ops.evopf_complete_bwd(data["y"], g2, need_dx=False)
    dz3, _ = ops.evopf_complete_bwd(data["y"], 0.5 * data["g_d"] - 2.0 * g2, need_dx=False)
    conv = ((data["status"] >> 8) == 0)
    lin = 0.5 * data["dz"] - 2.0 * dz2
    err = (dz3 - lin)[conv].abs().max(dim=1).values / lin[conv].abs().max(dim=1).values.clamp_min(1e-20)
    assert float(err.max()) <= 1e-3 and float(err.median()) <= 1e-5


def test_pipelined_backward_equals_one_pass_kernel(data):
    """Large aligned batches take the persistent double-buffered backward kernel (k_evopf_bwd_pipe); small batches and
    the ragged remainder take the one-pass kernel.  Same per-instance code: the results must be bit-equal."""
    from rpo_b200 import ops
    n = B - 51                                              # 8191 full tiles + a 77-row remainder
    y, g = data["y"][:n], data["g_d"][:n]
    for need_dx in (False, True):
        dz, dx = ops.evopf_complete_bwd(y, g, need_dx=need_dx)
        C = 1 << 14                                         # 128 tiles per call: below the persistent kernel's threshold
        parts = [ops.evopf_complete_bwd(y[b0:b0 + C], g[b0:b0 + C], need_dx=need_dx) for b0 in range(0, n, C)]
        assert torch.equal(dz, torch.cat([p[0] for p in parts]))
        if need_dx:
            assert torch.equal(dx, torch.cat([p[1] for p in parts]))
    assert torch.equal(data["dz"][:n], dz)


def test_correction_is_idempotent_on_feasible_points(data):
    """rpo_ddpg.py:271: one step is always taken, but it is a zero step where no inequality is violated."""
    from rpo_b200 import ops
    _, iq = ops.evopf_resid(data["st_d"], data["yc"], want_eq=False)
    feas = (iq.max(dim=1).values <= 0) & ((data["status"] >> 8) == 0)
    assert int(feas.sum()) > 1000
    st_f, y_f = data["st_d"][feas], data["yc"][feas]
    y2, flags = ops.evopf_correct(st_f, y_f, LR, MOM, EPS, 3)
    assert torch.equal(y2, y_f)


def test_classic_envs_at_full_size():
    from oracle.rpo_oracle import CartOracle, PenOracle
    from rpo_b200 import ops, synth
    st, z = synth.cart_inputs(B, seed=5)
    ac, a, _ = ops.cart_project(z.cuda() * 1.3, 1, 2e-2, 0.0, 1e-5, 10, want_a=True)
    eq, iq = ops.cart_resid(ac)
    assert float(eq.abs().max()) <= 2e-5                                 # correction moves along the equality manifold
    orc = CartOracle(partial_idx=1)
    idx = torch.arange(0, B, 97)
    a_o = orc.complete_partial(st[idx], z[idx] * 1.3)
    assert (a.cpu()[idx] - a_o).abs().max() <= 1e-5
    obs, zp = synth.pen_inputs(B, seed=6)
    ap_c, ap, _ = ops.pen_project(obs.cuda(), zp.cuda() * 1.5, 2e-3, 0.0, 1e-5, 10, want_a=True)
    eqp, _ = ops.pen_resid(obs.cuda(), ap)
    assert float(eqp.abs().max()) <= 1e-4
    po = PenOracle()
    assert rel_err(ap.cpu()[idx], po.complete_partial(obs[idx], zp[idx] * 1.5)).max() <= 1e-5
