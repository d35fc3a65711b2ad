"""The per-instance device code (rpo_b200/csrc/*_core.cuh + generated evopf_gen.cuh) compiled for the HOST by g++
(tests/hostsim) and checked against the oracle: validates the generated sparse LU / Newton / GRG math on a box
without a GPU.  float64 build: formulas to ~1e-9; float32 build: the tolerances the GPU tests use."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from tests.parity import GRAD_RTOL, Y_RTOL, assert_close_or_marginal, rel_err

HS = os.path.join(os.path.dirname(__file__), "hostsim")


@pytest.fixture(scope="module")
def libs():
    subprocess.check_call(["make", "-s", "-C", HS])
    out = {}
    for prec in ("f32", "f64"):
        lib = ctypes.CDLL(os.path.join(HS, "libhostsim_%s.so" % prec))
        assert lib.hs_sizeof_real() == (4 if prec == "f32" else 8)
        out[prec] = lib
    return out


def _P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _complete(lib, st, z, npd, creal, tol=1e-5, refine=1):
    B = st.shape[0]
    y = np.zeros((B, 43), npd)
    status = np.zeros(B, np.int32)
    lib.hs_evopf_complete(_P(st), st.shape[1], _P(z), _P(y), B, creal(tol), 50, _P(status), refine)
    return y, status


@pytest.mark.parametrize("dist", ["D1", "D2"])
def test_generated_math_float64_exact(libs, dist):
    from oracle.rpo_oracle import EvopfOracle, grad_steps
    from rpo_b200 import synth
    lib, orc = libs["f64"], EvopfOracle(dtype=torch.float64, const_dtype=torch.float32)
    B = 400
    st = synth.evopf_states(B, seed=3).double()
    z = synth.evopf_partial_actions(st.float(), seed=4, dist=dist).double()
    s_np, z_np = st.numpy().copy(), z.numpy().copy()
    y, status = _complete(lib, s_np, z_np, np.float64, ctypes.c_double, tol=1e-13)
    y_o, jac, inv, _ = orc.complete_partial_fwd(st, z, tol=1e-13, return_saved=True)
    assert (status >> 8).max() == 0
    assert rel_err(torch.tensor(y), y_o).max() < 1e-10
    g = synth.upstream_grad(B, 43, seed=5).double().numpy().copy()
    dz, dx = np.zeros((B, 14)), np.zeros((B, 28))
    yo = y_o.numpy().copy()
    lib.hs_evopf_backward(_P(yo), _P(g), _P(dz), _P(dx), B)
    # the oracle keeps the Jacobian of the last PRE-update iterate (evopf.py:838); at tol 1e-13 both coincide
    dx_o, dz_o = orc.complete_partial_bwd(torch.tensor(g), jac, inv)
    assert rel_err(torch.tensor(dz), dz_o).max() < 1e-8 and rel_err(torch.tensor(dx), dx_o).max() < 1e-8
    out, em, im = np.zeros((B, 43)), np.zeros(B), np.zeros(B)
    lib.hs_evopf_direction(_P(s_np), 57, _P(yo), _P(out), _P(em), _P(im), B)
    d_o = orc.ineq_partial_grad(st, y_o)
    assert rel_err(torch.tensor(out), d_o).max() < 1e-9
    assert np.abs(em - orc.eq_resid(st, y_o).abs().max(1).values.numpy()).max() < 1e-12
    # float32 inputs converted to float64: battery bounds are rounded to `real` = double here, so compare loosely
    assert np.abs(im - orc.ineq_dist(st, y_o).max(1).values.numpy()).max() < 1e-7
    yc = yo.copy()
    flags = np.zeros(10, np.int32)
    lib.hs_evopf_correct(_P(s_np), 57, _P(yc), B, ctypes.c_double(1e-4), ctypes.c_double(0.0), ctypes.c_double(1e-5), 10, _P(flags))
    yc_o, k = grad_steps(orc, st, y_o, 1e-4, 0.0, 1e-5, 10)
    assert k == 10 and flags.tolist() == [1] * 10
    assert rel_err(torch.tensor(yc), yc_o).max() < 1e-7
    eq, iq = np.zeros((B, 28)), np.zeros((B, 58))
    lib.hs_evopf_resid(_P(s_np), 57, _P(yc), _P(eq), _P(iq), B)
    assert (torch.tensor(eq) - orc.eq_resid(st, torch.tensor(yc))).abs().max() < 1e-12


@pytest.mark.parametrize("dist", ["D1", "D2"])
def test_generated_math_float32_tolerances(libs, dist, evopf_oracle, evopf_oracle64):
    from oracle.rpo_oracle import grad_steps_margins
    from rpo_b200 import synth
    lib, orc = libs["f32"], evopf_oracle
    B = 3000
    st = synth.evopf_states(B, seed=11)
    z = synth.evopf_partial_actions(st, seed=12, dist=dist)
    s_np, z_np = st.numpy().copy(), z.numpy().copy()
    y, status = _complete(lib, s_np, z_np, np.float32, ctypes.c_float)
    y_o, jac, inv, _ = orc.complete_partial_fwd(st, z, return_saved=True)
    assert (status >> 8).max() == 0 and (status & 0xff).max() <= 8
    yt = torch.tensor(y)
    assert rel_err(yt, y_o).max() <= Y_RTOL
    r = evopf_oracle64.eq_resid(st.double(), yt.double()).abs().max(1).values
    assert r.max() <= 1.5e-5 and (r > 1e-5).double().mean() <= 1e-3
    g = synth.upstream_grad(B, 43, seed=13).numpy().copy()
    dz, dx = np.zeros((B, 14), np.float32), np.zeros((B, 28), np.float32)
    lib.hs_evopf_backward(_P(y), _P(g), _P(dz), _P(dx), B)
    dx_o, dz_o = orc.complete_partial_bwd(torch.tensor(g), jac, inv)
    assert rel_err(torch.tensor(dz), dz_o).max() <= GRAD_RTOL and rel_err(torch.tensor(dx), dx_o).max() <= GRAD_RTOL
    yc = y.copy()
    flags = np.zeros(10, np.int32)
    lib.hs_evopf_correct(_P(s_np), 57, _P(yc), B, ctypes.c_float(1e-4), ctypes.c_float(0.0), ctypes.c_float(1e-5), 10, _P(flags))
    yc_o, k, margin = grad_steps_margins(orc, st, y_o, 1e-4, 0.0, 1e-5, 10)
    assert_close_or_marginal(rel_err(torch.tensor(yc), yc_o), margin, Y_RTOL)


def test_hostsim_matches_reference_goldens(libs):
    lib = libs["f32"]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "evopf_D1.npz"))
    st, z = g["state"].copy(), g["z"].copy()
    y, status = _complete(lib, st, z, np.float32, ctypes.c_float)
    assert rel_err(torch.tensor(y), torch.tensor(g["y"])).max() <= Y_RTOL
    B = st.shape[0]
    dz, dx = np.zeros((B, 14), np.float32), np.zeros((B, 28), np.float32)
    yg, gg = g["y"].copy(), g["g"].copy()
    lib.hs_evopf_backward(_P(yg), _P(gg), _P(dz), _P(dx), B)
    assert rel_err(torch.tensor(dz), torch.tensor(g["dz"])).max() <= GRAD_RTOL
    assert rel_err(torch.tensor(dx), torch.tensor(g["dx"])).max() <= GRAD_RTOL
    yc = yg.copy()
    lib.hs_evopf_correct(_P(st), 57, _P(yc), B, ctypes.c_float(1e-4), ctypes.c_float(0.0), ctypes.c_float(1e-5), 10, None)
    assert rel_err(torch.tensor(yc), torch.tensor(g["y_corr"])).max() <= Y_RTOL
    ycm = yg.copy()
    lib.hs_evopf_correct(_P(st), 57, _P(ycm), B, ctypes.c_float(1e-4), ctypes.c_float(0.5), ctypes.c_float(1e-5), 5, None)
    assert rel_err(torch.tensor(ycm), torch.tensor(g["y_corr_mom"])).max() <= Y_RTOL


def test_infeasible_power_flow_is_flagged(libs, evopf_oracle):
    """A collapsed-voltage instance (found at index 397174 of seed 0): the reference's Newton diverges silently
    (evopf.py:837 discards `converged`); the kernel code reports NOT_CONVERGED in the status word."""
    from rpo_b200 import synth
    st = synth.evopf_states(1 << 20, seed=0)[397174:397175].clone()
    z = synth.evopf_partial_actions(synth.evopf_states(1 << 20, seed=0), seed=1)[397174:397175].clone()
    y, status = _complete(libs["f32"], st.numpy().copy(), z.numpy().copy(), np.float32, ctypes.c_float)
    assert (status[0] & 0xff) == 50 and (status[0] & 0x100)
    y_o, _, _, iters = evopf_oracle.complete_partial_fwd(st, z, return_saved=True)
    assert iters == [50]


def test_fused_single_loop_equals_complete_then_correct(libs):
    """evopf_project_loop (one loop, shared eval+LU code) must equal evopf_complete followed by the correction steps
    bit for bit: same arithmetic, only the code layout differs."""
    from rpo_b200 import synth
    lib = libs["f32"]
    B = 500
    st = synth.evopf_states(B, seed=21).numpy().copy()
    z = synth.evopf_partial_actions(torch.tensor(st), seed=22).numpy().copy()
    y, status = _complete(lib, st, z, np.float32, ctypes.c_float, refine=0)     # the plain Newton sequence the fused loop runs
    yc = y.copy()
    flags = np.zeros(10, np.int32)
    lib.hs_evopf_correct(_P(st), 57, _P(yc), B, ctypes.c_float(1e-4), ctypes.c_float(0.5), ctypes.c_float(1e-5), 7, _P(flags))
    y2, yc2 = np.zeros_like(y), np.zeros_like(y)
    status2, flags2 = np.zeros(B, np.int32), np.zeros(10, np.int32)
    lib.hs_evopf_project(_P(st), 57, _P(z), _P(y2), _P(yc2), B, ctypes.c_float(1e-5), 50, ctypes.c_float(1e-4),
                         ctypes.c_float(0.5), ctypes.c_float(1e-5), 7, _P(flags2), _P(status2))
    assert np.array_equal(y, y2) and np.array_equal(yc, yc2)
    assert np.array_equal(status, status2) and np.array_equal(flags, flags2)


@pytest.mark.parametrize("scale", [0.4, 1.0, 1.5])
def test_load_scale_sweep(scale, evopf_oracle):
    """Far outside the synthetic distribution (40 % ... 150 % of the nominal load): the generated Newton completion keeps
    tracking the oracle on every instance it reports as converged, and reports the others in the status word."""
    from rpo_b200 import synth
    from tests import hostsim_ops as hs
    B = 384
    st = synth.evopf_states(B, seed=int(scale * 100))
    st[:, :28] *= scale
    z = synth.evopf_partial_actions(st, seed=7, dist="D1")
    y_h, status = hs.evopf_complete_fwd(st, z, return_status=True)
    y_o = evopf_oracle.complete_partial_fwd(st, z)
    conv = ((status >> 8) == 0) & torch.isfinite(y_o).all(dim=1)
    assert int(conv.sum()) >= B - 2                      # at most a couple of infeasible power flows at 150 %
    rel = (y_h - y_o).abs().max(dim=1).values / y_o.abs().max(dim=1).values
    assert float(rel[conv].max()) <= Y_RTOL / 5
    assert int((status[conv] & 0xFF).max()) <= 7


@pytest.mark.parametrize("K,mom,lr", [(50, 0.5, 1e-4), (50, 0.0, 5e-4)])
def test_long_correction_and_momentum(K, mom, lr, evopf_oracle):
    """eval-length loops (eval_steps = 50, rpo_ddpg.py:288-305), momentum and a 5x larger step: the generated correction
    step stays on the oracle's trajectory (mask flips aside)."""
    from oracle.rpo_oracle import grad_steps
    from rpo_b200 import synth
    from tests import hostsim_ops as hs
    B = 64
    st = synth.evopf_states(B, seed=77)
    z = synth.evopf_partial_actions(st, seed=78)
    y = evopf_oracle.complete_partial_fwd(st, z)
    yc_o, k = grad_steps(evopf_oracle, st, y, lr, mom, 1e-5, K)
    yc_h, flags = hs.evopf_correct(st, y, lr, mom, 1e-5, K)
    assert k == K and flags.tolist() == [1] * K
    rel = (yc_h - yc_o).abs().max(dim=1).values / yc_o.abs().max(dim=1).values
    assert float(rel.median()) <= 1e-6 and int((rel > Y_RTOL).sum()) <= 1


@pytest.mark.parametrize("dist", ["D1", "D2"])
def test_refined_completion_residual_is_below_the_float32_reference(libs, dist, evopf_oracle):
    """The mixed-precision refinement (float32 Newton + one chord step on the float64-accumulated residual): same stop rule,
    y within tolerance of the reference, and an equality residual (evaluated in float64 on the reference's float32
    constants) BELOW the float32 reference's on the same rows — in the median, the upper quantiles and the maximum."""
    from oracle.rpo_oracle import EvopfOracle
    from rpo_b200 import synth
    lib = libs["f32"]
    B = 8192
    st = synth.evopf_states(B, seed=31)
    z = synth.evopf_partial_actions(st, seed=32, dist=dist)
    y_ref = evopf_oracle.complete_partial_fwd(st, z)
    s_np, z_np = st.numpy().copy(), z.numpy().copy()
    y_r, status_r = _complete(lib, s_np, z_np, np.float32, ctypes.c_float, refine=1)
    y_p, status_p = _complete(lib, s_np, z_np, np.float32, ctypes.c_float, refine=0)
    assert (status_r >> 8).max() == 0 and (status_p >> 8).max() == 0
    chk = EvopfOracle(dtype=torch.float64, const_dtype=torch.float32)
    res = lambda y: chk.eq_resid(st.double(), torch.as_tensor(y).double()).abs().max(dim=1).values
    r_ref, r_r, r_p = res(y_ref), res(y_r), res(y_p)
    q = torch.tensor([0.5, 0.99, 0.999], dtype=torch.float64)
    for a, b in zip(torch.quantile(r_r, q), torch.quantile(r_ref, q)):
        assert float(a) <= 0.85 * float(b), (float(a), float(b))
    assert float(r_r.max()) <= 0.8 * float(r_ref.max()) and float(r_r.max()) <= 0.6e-5
    assert float((r_r > 1e-5).double().mean()) <= float((r_ref > 1e-5).double().mean())
    assert float(r_p.max()) <= 1.3 * float(r_ref.max())                       # the plain sequence: same noise class as the reference
    assert rel_err(torch.as_tensor(y_r), y_ref).max() <= 2e-5
    # the refinement replaces Newton iterations, it does not add to them: applied updates per instance (self vote)
    assert (status_r & 0xff).mean() <= (status_p & 0xff).mean() + 0.5


@pytest.mark.parametrize("p", [0, 1])
def test_cart_reduced_correction_step_is_bit_identical_to_the_row_form(libs, p):
    """The K-step cart kernels run cart_correct_step_red (3 compares on |a_p| against rounding boundaries found from the
    fused test itself, pairwise |.| in the loop test); it must reproduce the row-by-row form of cartpole.py:396-408 bit
    for bit — including on the floats adjacent to every boundary, with momentum, and when a large step overshoots."""
    lib = libs["f32"]
    f = ctypes.c_float
    rng = np.random.default_rng(p)
    t, A = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib.hs_cart_thresholds(p, _P(t), _P(A))
    assert (A > 0).all() and (t > 0).all()
    B = 200000
    z = rng.uniform(-14, 14, B).astype(np.float32)
    pos = 0
    for j in range(3):                                   # 40 floats on each side of +-t_j
        for sgn in (1.0, -1.0):
            for away in (100.0, 0.0):
                v = np.float32(sgn * t[j])
                for _ in range(40):
                    z[pos] = v
                    pos += 1
                    v = np.nextafter(v, np.float32(sgn * away), dtype=np.float32)
    a = np.zeros((B, 2), np.float32)
    lib.hs_cart_complete(_P(z), _P(a), B, p)
    for lr, mom, K in ((2e-2, 0.0, 10), (2e-2, 0.5, 7), (0.3, 0.0, 50)):
        a1, a2 = a.copy(), a.copy()
        f1, f2 = np.zeros(64, np.int32), np.zeros(64, np.int32)
        lib.hs_cart_correct(_P(a1), B, p, f(lr), f(mom), f(1e-5), K, _P(f1))
        lib.hs_cart_correct_red(_P(a2), B, p, f(lr), f(mom), f(1e-5), K, _P(f2))
        assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32)) and np.array_equal(f1, f2)
        assert np.abs(a1 - a).max() > 0.1


def test_residual_vjp_against_finite_differences_in_float64(libs):
    """evopf_eq_vjp / evopf_ineq_vjp (the autograd of eq_resid / ineq_resid / ineq_dist the kernels implement) in the float64
    build of the generated code against central finite differences of the float64 residual functions themselves:
    (d r / d y)^T g for random g, every one of the 43 action components."""
    from rpo_b200 import synth
    lib = libs["f64"]
    B = 6
    st = synth.evopf_states(B, seed=41).double().numpy().copy()
    z = synth.evopf_partial_actions(torch.tensor(st).float(), seed=42).double().numpy().copy()
    y, _ = _complete(lib, st, z, np.float64, ctypes.c_double, tol=1e-12)
    rng = np.random.default_rng(43)
    y = y + 0.01 * rng.standard_normal(y.shape)                # off the solution manifold, away from the inequality kinks
    ge, gi = rng.standard_normal((B, 28)), rng.standard_normal((B, 58))

    def resid(yy):
        eq, iq = np.zeros((B, 28)), np.zeros((B, 58))
        lib.hs_evopf_resid(_P(st), st.shape[1], _P(np.ascontiguousarray(yy)), _P(eq), _P(iq), B)
        return eq, iq

    eq0, iq0 = resid(y)
    assert np.abs(iq0).min() > 1e-5                              # no inequality row sits on its kink
    for mode in (0, 1):
        dy = np.zeros((B, 43))
        lib.hs_evopf_resid_bwd(_P(st), st.shape[1], _P(y), _P(ge), _P(gi), mode, _P(dy), B)
        fd = np.zeros((B, 43))
        h = 1e-6
        for j in range(43):
            yp, ym = y.copy(), y.copy()
            yp[:, j] += h
            ym[:, j] -= h
            (ep, ip), (em, im) = resid(yp), resid(ym)
            if mode == 1:
                ip, im = np.maximum(ip, 0), np.maximum(im, 0)
            fd[:, j] = (((ep - em) * ge).sum(1) + ((ip - im) * gi).sum(1)) / (2 * h)
        assert np.abs(dy - fd).max() <= 1e-7 * max(1.0, np.abs(fd).max()), (mode, np.abs(dy - fd).max())


def test_cart_reduced_step_bit_identity_on_arbitrary_float_bit_patterns(libs):
    """Same property over the whole float32 range: random bit patterns for both forces (denormals, huge values, +-0, inf,
    NaN included), fed to the correction loops directly — the reduced form must equal the row form bit for bit, and raise
    the same loop-test flags."""
    lib = libs["f32"]
    f = ctypes.c_float
    rng = np.random.default_rng(99)
    B = 300000
    bits = rng.integers(0, 2 ** 32, size=(B, 2), dtype=np.uint64).astype(np.uint32)
    a = bits.view(np.float32).copy()
    a[:1000] = rng.uniform(-12, 12, size=(1000, 2)).astype(np.float32)
    a[1000:1008] = np.array([[0.0, -0.0], [-0.0, 0.0], [np.inf, 1.0], [-np.inf, 1.0], [np.nan, 1.0], [1.0, np.nan],
                             [3.4e38, -3.4e38], [1e-45, -1e-45]], np.float32)
    for p in (0, 1):
        for lr, mom, K in ((2e-2, 0.0, 10), (2e-2, 0.9, 4)):
            a1, a2 = a.copy(), a.copy()
            f1, f2 = np.zeros(64, np.int32), np.zeros(64, np.int32)
            lib.hs_cart_correct(_P(a1), B, p, f(lr), f(mom), f(1e-5), K, _P(f1))
            lib.hs_cart_correct_red(_P(a2), B, p, f(lr), f(mom), f(1e-5), K, _P(f2))
            same = (a1.view(np.uint32) == a2.view(np.uint32)) | (np.isnan(a1) & np.isnan(a2))
            assert same.all(), (p, lr, mom, a[~same.all(1)][:4], a1[~same.all(1)][:4], a2[~same.all(1)][:4])
            assert np.array_equal(f1, f2)
