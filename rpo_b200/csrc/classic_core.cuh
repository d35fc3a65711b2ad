// Per-instance math of the two classic-control environments and of the batched env steps.
// `__host__ __device__` for the same reason as evopf_core.cuh (host simulator in tests/hostsim).
#pragma once

#include "rpo_math.cuh"

// ------------------------------------------------------------------------------------------------
// Safe Cartpole — rpo/env/classic_control/cartpole.py
//   equality  sin(delta) . a = 0, delta = (pi/3, -pi/6)                         (:85,124-127)
//   inequalities A a <= b, A = [cos d; -cos d; e1; -e1; e2; -e2], b = (8,8,10,10,10,10)  (:129-136)
// The float32 constants are the values torch produces for sin/cos(tensor(delta, float32)).
// ------------------------------------------------------------------------------------------------
struct CartK {
    real E[2];      // diff_eq
    real C[2];      // cos(delta)
    real inv_o;     // diff_eq_other_inv
    real ep;        // diff_eq_partial
    real Ared[6];   // diff_grad_partial (:397-398)
    int p, o;
};

RPO_HD CartK cart_constants(int partial_idx) {
    CartK k;
    k.E[0] = RC(0x1.bb67aep-1);   // sin(pi/3)
    k.E[1] = RC(-0.5);            // sin(-pi/6)
    k.C[0] = RC(0x1.fffffep-2);   // cos(pi/3)
    k.C[1] = RC(0x1.bb67aep-1);   // cos(-pi/6)
    k.p = partial_idx;
    k.o = 1 - partial_idx;
    k.inv_o = RC(1.0) / (partial_idx == 0 ? k.E[1] : k.E[0]);
    k.ep = partial_idx == 0 ? k.E[0] : k.E[1];
    const real m = k.inv_o * k.ep;
    // A[:, p] - A[:, o] * (inv_o * ep)
    const real A0[6] = {k.C[0], -k.C[0], RC(1.0), RC(-1.0), RC(0.0), RC(0.0)};
    const real A1[6] = {k.C[1], -k.C[1], RC(0.0), RC(0.0), RC(1.0), RC(-1.0)};
    for (int r = 0; r < 6; ++r) {
        const real ap = k.p == 0 ? A0[r] : A1[r];
        const real ao = k.p == 0 ? A1[r] : A0[r];
        k.Ared[r] = ap - ao * m;
    }
    return k;
}

// NOTE: the action is indexed with selects, never with a run-time subscript (a[k.p] would force the 2-element array
// into local memory and every loop iteration through the LSU: measured 3x slower).
RPO_HD real cart_sel(bool first, real x, real y) { return first ? x : y; }

// complete_partial (:369-373): a_p = z, a_o = (0 - z*E_p) * inv(E_o)
RPO_HD void cart_complete(const CartK& k, real z, real* a) {
    const real ao = (RC(0.0) - z * k.ep) * k.inv_o;
    a[0] = cart_sel(k.p == 0, z, ao);
    a[1] = cart_sel(k.p == 0, ao, z);
}

// autograd of complete_partial
RPO_HD real cart_complete_bwd(const CartK& k, const real* da) {
    const real dp = cart_sel(k.p == 0, da[0], da[1]), dn = cart_sel(k.p == 0, da[1], da[0]);
    return dp + (dn * k.inv_o) * (-k.ep);
}

// eq_resid (:375-376) and ineq_resid (:378-379); returns the loop test max(|eq|, relu(ineq)) pieces
RPO_HD void cart_resid(const CartK& k, const real* a, real* eq, real* ineq) {
    *eq = RC(0.0) - FMA(a[0], k.E[0], a[1] * k.E[1]);
    const real ca = FMA(a[0], k.C[0], a[1] * k.C[1]);
    ineq[0] = ca - RC(8.0);
    ineq[1] = -ca - RC(8.0);          // row 1 of A is -row 0 (cartpole.py:135)
    ineq[2] = a[0] - RC(10.0);
    ineq[3] = -a[0] - RC(10.0);
    ineq[4] = a[1] - RC(10.0);
    ineq[5] = -a[1] - RC(10.0);
}

// autograd of eq_resid (:375-376: bias - a E^T) and ineq_resid / ineq_dist (:378-387: a A^T - b, clamp(., 0)) w.r.t. the
// action: da = -g_eq E + A^T (g_ineq . m), m = [resid >= 0] for ineq_dist (relu), ones for ineq_resid.  g_eq / g_ineq
// may be null.  The actor loss differentiates these (rpo_ddpg.py:312-319).
RPO_HD void cart_resid_vjp(const CartK& k, const real* a, const real* g_eq, const real* g_ineq, int relu, real* da) {
    real d0 = RC(0.0), d1 = RC(0.0);
    if (g_eq) { d0 = -(g_eq[0] * k.E[0]); d1 = -(g_eq[0] * k.E[1]); }
    if (g_ineq) {
        real eq, iq[6], m[6];
        cart_resid(k, a, &eq, iq);
        for (int r = 0; r < 6; ++r) m[r] = (!relu || iq[r] >= RC(0.0)) ? g_ineq[r] : RC(0.0);
        const real gc = m[0] - m[1];
        d0 += FMA(gc, k.C[0], m[2] - m[3]);
        d1 += FMA(gc, k.C[1], m[4] - m[5]);
    }
    da[0] = d0; da[1] = d1;
}

// ineq_partial_grad (:396-408)
RPO_HD void cart_direction(const CartK& k, const real* a, real* out) {
    const real b[6] = {RC(8.0), RC(8.0), RC(10.0), RC(10.0), RC(10.0), RC(10.0)};
    const real ap = cart_sel(k.p == 0, a[0], a[1]);
    real g = RC(0.0);
    for (int r = 0; r < 6; ++r) {
        const real bm = FMA(ap, k.Ared[r], -b[r]);
        g += (bm > RC(0.0) ? k.Ared[r] : RC(0.0));
    }
    const real go = -(g * k.ep) * k.inv_o;
    out[0] = cart_sel(k.p == 0, g, go);
    out[1] = cart_sel(k.p == 0, go, g);
}

template <bool MOMENTUM>
RPO_HD bool cart_correct_step(const CartK& k, real* a, real* prev, real lr, real momentum, real corr_eps) {
    real eq, iq[6], dir[2];
    cart_resid(k, a, &eq, iq);
    real im = RC(0.0);
    for (int r = 0; r < 6; ++r) im = FMAX(im, iq[r]);
    cart_direction(k, a, dir);
    for (int i = 0; i < 2; ++i) {
        real s = lr * dir[i];
        if (MOMENTUM) {
            s = s + momentum * prev[i];
            prev[i] = s;
        }
        a[i] = a[i] - s;
    }
    return (FABS(eq) > corr_eps) || (im > corr_eps);
}

// ---- reduced form of the correction step (same results bit for bit, ~2x fewer instructions) ---------------------------
// The six inequality rows come in +- pairs (row 2k+1 = -row 2k, cartpole.py:129-136), so
//   * max over the rows of a pair is |.| - b:  max ineq_resid = max(|c.a| - 8, |a0| - 10, |a1| - 10);
//   * in the direction at most one row of a pair is active, and which one depends only on the sign of a_p:
//       FMA(a_p,  A_k, -b_k) > 0  <=>  a_p >  t_k,     FMA(a_p, -A_k, -b_k) > 0  <=>  a_p < -t_k      (A_k = Ared[2k] > 0)
//     with t_k the largest float for which the fused test is still false.  The sum over the rows in the reference's order,
//     ((0 + s_0) + s_1) + s_2 with s_k in {+A_k, -A_k, 0}, is sign(a_p) ((m_0 + m_1) + m_2), m_k = [|a_p| > t_k] A_k
//     (adding zeros is exact, negation commutes with rounding).
// t_k is found from the fused test itself (start at b_k / A_k, walk to the boundary), so the equivalence holds by
// construction for every float a_p.
struct CartRed {
    real t[3], A[3];
};

RPO_HD CartRed cart_reduced(const CartK& k) {
    const real b[3] = {RC(8.0), RC(10.0), RC(10.0)};
    CartRed r;
    for (int j = 0; j < 3; ++j) {
        const real A = k.Ared[2 * j];
        real t = b[j] / A;
        for (int it = 0; it < 8 && FMA(t, A, -b[j]) > RC(0.0); ++it) t = rpo_next(t, false);
        for (int it = 0; it < 8 && !(FMA(rpo_next(t, true), A, -b[j]) > RC(0.0)); ++it) t = rpo_next(t, true);
        r.t[j] = t;
        r.A[j] = A;
    }
    return r;
}

// One correction step in the reduced form; a_p / a_o = policy-controlled / completed force.  Returns the loop test.
template <bool MOMENTUM, bool WANT_TEST>
RPO_HD bool cart_correct_step_red(const CartK& k, const CartRed& r, real& ap, real& ao, real& prev_p, real& prev_o, real lr,
                                  real momentum, real corr_eps) {
    bool unsat = false;
    if (WANT_TEST) {
        const real a0 = cart_sel(k.p == 0, ap, ao), a1 = cart_sel(k.p == 0, ao, ap);
        const real eqv = FMA(a0, k.E[0], a1 * k.E[1]);                 // eq_resid = 0 - eqv
        const real ca = FMA(a0, k.C[0], a1 * k.C[1]);
        const real im = FMAX(FMAX(FABS(ca) - RC(8.0), FABS(a0) - RC(10.0)), FMAX(FABS(a1) - RC(10.0), RC(0.0)));
        unsat = (FABS(eqv) > corr_eps) || (im > corr_eps);
    }
    const real u = FABS(ap);
    const real gm = ((u > r.t[0] ? r.A[0] : RC(0.0)) + (u > r.t[1] ? r.A[1] : RC(0.0))) + (u > r.t[2] ? r.A[2] : RC(0.0));
    const real g = ap < RC(0.0) ? -gm : gm;
    const real go = -(g * k.ep) * k.inv_o;
    real sp = lr * g, so = lr * go;
    if (MOMENTUM) {
        sp = sp + momentum * prev_p; prev_p = sp;
        so = so + momentum * prev_o; prev_o = so;
    }
    ap = ap - sp;
    ao = ao - so;
    return unsat;
}

// CartSafeEnv.step (:166-229), float64 dynamics as the reference's numpy path; float32 I/O.
RPO_HD int cart_step(const real* s, const real* a, real* ns) {
    const double g = 9.8, mc = 1.0, mp = 0.1, l = 0.5, tau = 0.02, mu_c = 0.1, mu_p = 0.01;
    const double M = mp + mc, pml = mp * l;
    const double d0 = 1.0471975511965976, d1 = -0.5235987755982988;   // pi/3, -pi/6
    double a0 = (double)a[0], a1 = (double)a[1];
    a0 = a0 < -10.0 ? -10.0 : (a0 > 10.0 ? 10.0 : a0);
    a1 = a1 < -10.0 ? -10.0 : (a1 > 10.0 ? 10.0 : a1);
    const double x = s[0], xd = s[1], th = s[3], thd = s[4], thacc0 = s[5];
    const double fx = a0 * cos(d0) + a1 * cos(d1);
    const double fy = a0 * sin(d0) + a1 * sin(d1);
    const double ct = cos(th), st = sin(th);
    const double n_c = fy + M * g - pml * (thacc0 * st + thd * thd * ct);
    const double v = n_c * xd;
    const double sg = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0);
    const double temp = (fx + pml * thd * thd * (st + mu_c * sg * ct)) / M + mu_c * g * sg;
    const double thacc = (g * st - ct * temp - mu_p * thd / pml) / (l * (4.0 / 3.0 - mp * ct * (ct - mu_c * g * sg) / M));
    const double xacc = (fx + pml * (thd * thd * st - thacc * ct) - mu_c * n_c * sg) / M;
    const double x2 = x + tau * xd, xd2 = xd + tau * xacc, th2 = th + tau * thd, thd2 = thd + tau * thacc;
    ns[0] = (real)x2; ns[1] = (real)xd2; ns[2] = (real)xacc; ns[3] = (real)th2; ns[4] = (real)thd2; ns[5] = (real)thacc;
    const double thr = 12.0 * 2.0 * 3.141592653589793 / 360.0;
    return (x2 < -2.4 || x2 > 2.4 || th2 < -thr || th2 > thr) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Spring Pendulum — rpo/env/classic_control/pendulum.py (row-wise semantics)
//   equality  fx sin(th) + fy cos(th) = beta(obs)   (set_eq :264-288)
//   inequality fx^2 + fy^2 <= 32                     (:302-303)
// ------------------------------------------------------------------------------------------------
struct PenEq {
    real ep, eo, inv, bias;
};

RPO_HD PenEq pen_set_eq(const real* obs) {
    PenEq e;
    const real c = obs[0], s = obs[1], thd = obs[2], l = obs[3], ld = obs[4];
    e.ep = s;
    e.eo = c;
    e.inv = RC(1.0) / c;
    // - m_dt*ldot - (l*m*thd^2 - k*(l-l0) - m*g*cos)   with m_dt = 10, m = .5, k = 1, l0 = 1, m*g = 5
    e.bias = -RC(10.0) * ld - ((l * RC(0.5)) * (thd * thd) - RC(1.0) * (l - RC(1.0)) - RC(5.0) * c);
    return e;
}

RPO_HD void pen_complete(const PenEq& e, real z, real* a) {
    a[0] = z;
    a[1] = (e.bias - z * e.ep) * e.inv;
}

RPO_HD real pen_complete_bwd(const PenEq& e, const real* da) { return da[0] + (da[1] * e.inv) * (-e.ep); }

RPO_HD void pen_resid(const PenEq& e, const real* a, real* eq, real* ineq) {
    *eq = e.bias - FMA(a[0], e.ep, a[1] * e.eo);
    *ineq = FMA(a[0], a[0], a[1] * a[1]) - RC(32.0);
}

// autograd of eq_resid (:296-298: bias - a . [sin, cos]) and ineq_resid / ineq_dist (:300-311: a.a - 32) w.r.t. the action
RPO_HD void pen_resid_vjp(const PenEq& e, const real* a, const real* g_eq, const real* g_ineq, int relu, real* da) {
    real d0 = RC(0.0), d1 = RC(0.0);
    if (g_eq) { d0 = -(g_eq[0] * e.ep); d1 = -(g_eq[0] * e.eo); }
    if (g_ineq) {
        real eq, iq;
        pen_resid(e, a, &eq, &iq);
        const real m = (!relu || iq >= RC(0.0)) ? g_ineq[0] : RC(0.0);
        d0 = FMA(RC(2.0) * a[0], m, d0);
        d1 = FMA(RC(2.0) * a[1], m, d1);
    }
    da[0] = d0; da[1] = d1;
}

RPO_HD void pen_direction(const PenEq& e, const real* a, real* out) {
    const real D0 = RC(2.0) * a[0], D1 = RC(2.0) * a[1];
    const real dgp = FMA(-D1, e.inv * e.ep, D0);
    const real bgp = FMA(-(e.bias * e.inv), D1, RC(32.0));
    const real bm = FMA(a[0], dgp, -bgp);
    const real g = (bm > RC(0.0) ? RC(1.0) : RC(0.0)) * dgp;
    out[0] = g;
    out[1] = -(g * e.ep) * e.inv;
}

template <bool MOMENTUM>
RPO_HD bool pen_correct_step(const PenEq& e, real* a, real* prev, real lr, real momentum, real corr_eps) {
    real eq, iq, dir[2];
    pen_resid(e, a, &eq, &iq);
    pen_direction(e, a, dir);
    for (int i = 0; i < 2; ++i) {
        real s = lr * dir[i];
        if (MOMENTUM) {
            s = s + momentum * prev[i];
            prev[i] = s;
        }
        a[i] = a[i] - s;
    }
    return (FABS(eq) > corr_eps) || (iq > corr_eps);
}

// SpringPendulumEnv.step (:80-128), float64 dynamics; st = (th, thdot, l, ldot)
RPO_HD int pen_step(const real* st, const real* a, real* nst, real* nobs, real* reward) {
    const double g = 10.0, m = 0.5, dt = 0.05, l0 = 1.0, k = 1.0, PI = 3.141592653589793;
    const double th = st[0], thd = st[1], l = st[2], ld = st[3];
    double fx = (double)a[0], fy = (double)a[1];
    fx = fx < -6.0 ? -6.0 : (fx > 6.0 ? 6.0 : fx);
    fy = fy < -6.0 ? -6.0 : (fy > 6.0 ? 6.0 : fy);
    const double sth = sin(th), cth = cos(th);
    const double fth = -fy * sth + fx * cth;
    const double fl = fy * cth + fx * sth;
    double w = fmod(th + PI, 2.0 * PI);
    if (w < 0.0) w += 2.0 * PI;              // python % semantics
    const double costs = fabs(w - PI);
    const double thacc = (fth - m * (g * sth + 2.0 * ld * thd)) / (l * m);
    const double lacc = (fl - m * g * cth + m * l * thd * thd - k * (l - l0)) / m;
    double nthd = thd + thacc * dt;
    const double nld = ld + lacc * dt;
    const double nth = th + nthd * dt;
    const double nl = l + ld * dt;
    nthd = nthd < -8.0 ? -8.0 : (nthd > 8.0 ? 8.0 : nthd);
    nst[0] = (real)nth; nst[1] = (real)nthd; nst[2] = (real)nl; nst[3] = (real)nld;
    nobs[0] = (real)cos(nth); nobs[1] = (real)sin(nth); nobs[2] = (real)nthd; nobs[3] = (real)nl; nobs[4] = (real)nld;
    *reward = (real)(1.0 / (100.0 * costs + 1.0));
    return (nl <= 0.5 || nl >= 1.5 || nth >= PI / 12.0 || nth <= -PI / 12.0) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// EVOPFEnv.step dynamics part (evopf.py:348-366, Battery.step :74-102, obj_fn :509-518); float64 as the
// reference's numpy path.  next_state = [next_exo[0:28] | soc' | next_exo[28:52]].
// ------------------------------------------------------------------------------------------------
RPO_HD void evopf_step_dyn(const real* state, const real* y, const real* next_exo, real* next_state, real* reward) {
    const double c2[5] = {0.0430293, 0.25, 0.01, 0.01, 0.01};
    const double c1[5] = {20.0, 20.0, 40.0, 40.0, 40.0};
    double cost = 0.0;
    for (int g = 0; g < 5; ++g) {
        const double mw = (double)y[g] * 100.0;
        cost += c2[g] * mw * mw;
    }
    double lin = 0.0;
    for (int g = 0; g < 5; ++g) lin += c1[g] * ((double)y[g] * 100.0);
    cost = cost + lin + 0.0;
    const double reward_grid = -(cost / (100.0 * 100.0));
    const double price0 = (double)state[33];
    double ev = 0.0;
    for (int g = 0; g < 5; ++g) {
        const double soc = (double)state[28 + g];
        const double pe = (double)y[38 + g];
        const double hi = (1.0 / 0.9) * fmin(0.2, 0.8 - soc);
        const double lo = 0.9 * fmax(-0.2, 0.1 - soc);
        double pc = pe < lo ? lo : (pe > hi ? hi : pe);
        pc = (pc >= 0.0) ? pc * 0.9 : pc / 0.9;
        next_state[28 + g] = (real)(soc + pc);
        ev += -pe * price0;
    }
    *reward = (real)(1.0 * reward_grid + 5.0 * ev);
    for (int i = 0; i < 28; ++i) next_state[i] = next_exo[i];
    for (int i = 0; i < 24; ++i) next_state[33 + i] = next_exo[28 + i];
}
