// Host-side simulator of the per-instance device code — TEST INFRASTRUCTURE ONLY.
// Compiles rpo_b200/csrc/{evopf,classic}_core.cuh with g++ (RPO_REAL = float or double) so the CPU test-suite
// can check the generated math against the oracle where no GPU exists.  Never loaded by rpo_b200/.
#include <stdint.h>
#include "../../rpo_b200/csrc/evopf_core.cuh"
#include "../../rpo_b200/csrc/classic_core.cuh"
#include "../../rpo_b200/csrc/box_core.cuh"

extern "C" {

void hs_evopf_complete(const real* state, int ld, const real* z, real* y, int B, real tol, int max_iters, int* status, int refine) {
    for (int b = 0; b < B; ++b) {
        int st = refine ? evopf_complete<true>(state + (long)b * ld, z + (long)b * 14, y + (long)b * 43, tol, max_iters, RpoVoteSelf())
                        : evopf_complete<false>(state + (long)b * ld, z + (long)b * 14, y + (long)b * 43, tol, max_iters, RpoVoteSelf());
        if (status) status[b] = st;
    }
}

void hs_evopf_backward(const real* y, const real* g, real* dz, real* dx, int B) {
    for (int b = 0; b < B; ++b)
        evopf_backward(y + (long)b * 43, g + (long)b * 43, dz + (long)b * 14, dx ? dx + (long)b * 28 : (real*)0);
}

void hs_evopf_direction(const real* state, int ld, const real* y, real* out, real* eqmax, real* ineqmax, int B) {
    for (int b = 0; b < B; ++b) {
        real pe_max[5], pe_min[5];
        evopf_pe_bounds(state + (long)b * ld + 28, pe_max, pe_min);
        evopf_direction(state + (long)b * ld, pe_max, pe_min, y + (long)b * 43, out + (long)b * 43, eqmax + b, ineqmax + b);
    }
}

// flags[k] |= 1 if any instance fails the loop test before step k (k = 0..n_steps-1)
void hs_evopf_correct(const real* state, int ld, real* y, int B, real lr, real momentum, real corr_eps, int n_steps,
                      int* flags) {
    for (int b = 0; b < B; ++b) {
        real prev[43], pe_max[5], pe_min[5];
        for (int i = 0; i < 43; ++i) prev[i] = 0;
        evopf_pe_bounds(state + (long)b * ld + 28, pe_max, pe_min);
        for (int k = 0; k < n_steps; ++k) {
            real mp;
            bool unsat = evopf_correct_step<true>(state + (long)b * ld, pe_max, pe_min, y + (long)b * 43, prev, lr, momentum,
                                                  corr_eps, &mp);
            if (unsat && flags) flags[k] = 1;
        }
    }
}

void hs_evopf_resid(const real* state, int ld, const real* y, real* eq, real* ineq, int B) {
    for (int b = 0; b < B; ++b) {
        EvopfPt pt;
        evopf_eval_flows(y + (long)b * 43, pt);
        evopf_eq_resid(state + (long)b * ld, y + (long)b * 43, pt, eq + (long)b * 28);
        real pe_max[5], pe_min[5];
        evopf_pe_bounds(state + (long)b * ld + 28, pe_max, pe_min);
        evopf_ineq(pe_max, pe_min, y + (long)b * 43, ineq + (long)b * 58, (real*)0);
    }
}

struct HsOnY {
    real* y_out; int* status;
    void operator()(const real* y, int st) const {
        if (y_out) for (int i = 0; i < 43; ++i) y_out[i] = y[i];
        if (status) *status = st;
    }
};
struct HsOnFlag {
    int* flags;
    void operator()(int k, bool unsat) const { if (flags && unsat) flags[k] = 1; }
};

// fused single-loop projection (evopf_project_loop): y (nullable) = completion output, yc = corrected action
void hs_evopf_project(const real* state, int ld, const real* z, real* y, real* yc, int B, real tol, int max_iters, real lr,
                      real momentum, real corr_eps, int n_steps, int* flags, int* status) {
    for (int b = 0; b < B; ++b) {
        real yy[43], scratch[EVOPF_SCRATCH];
        HsOnY oy = {y ? y + (long)b * 43 : (real*)0, status ? status + b : (int*)0};
        HsOnFlag of = {flags};
        evopf_project_loop<true>(state + (long)b * ld, z + (long)b * 14, yy, scratch, tol, max_iters, lr, momentum, corr_eps, n_steps,
                                 RpoVoteSelf(), oy, of);
        for (int i = 0; i < 43; ++i) yc[(long)b * 43 + i] = yy[i];
    }
}

// ---- env steps and the classic-control environments (same per-instance code as the kernels) -------------------
void hs_evopf_step(const real* state, const real* y, const real* exo, real* ns, real* reward, real* eqv, real* iqv, int B) {
    for (int b = 0; b < B; ++b) {
        const real* s = state + (long)b * 57;
        const real* yy = y + (long)b * 43;
        evopf_step_dyn(s, yy, exo + (long)b * 52, ns + (long)b * 57, reward + b);
        if (eqv) {
            EvopfPt pt;
            evopf_eval_flows(yy, pt);
            evopf_eq_resid(s, yy, pt, eqv + (long)b * 28);
        }
        if (iqv) {
            real pe_max[5], pe_min[5];
            evopf_pe_bounds(s + 28, pe_max, pe_min);
            evopf_ineq(pe_max, pe_min, yy, iqv + (long)b * 58, (real*)0);
            for (int i = 0; i < 58; ++i) iqv[(long)b * 58 + i] = FMAX(iqv[(long)b * 58 + i], RC(0.0));
        }
    }
}

void hs_cart_complete(const real* z, real* a, int B, int p) {
    const CartK k = cart_constants(p);
    for (int b = 0; b < B; ++b) cart_complete(k, z[b], a + 2 * (long)b);
}
void hs_cart_complete_bwd(const real* da, real* dz, int B, int p) {
    const CartK k = cart_constants(p);
    for (int b = 0; b < B; ++b) dz[b] = cart_complete_bwd(k, da + 2 * (long)b);
}
void hs_cart_direction(const real* a, real* out, int B, int p) {
    const CartK k = cart_constants(p);
    for (int b = 0; b < B; ++b) cart_direction(k, a + 2 * (long)b, out + 2 * (long)b);
}
void hs_cart_correct(real* a, int B, int p, real lr, real momentum, real eps, int n_steps, int* flags) {
    const CartK k = cart_constants(p);
    for (int b = 0; b < B; ++b) {
        real prev[2] = {0, 0};
        for (int s = 0; s < n_steps; ++s)
            if (cart_correct_step<true>(k, a + 2 * (long)b, prev, lr, momentum, eps) && flags) flags[s] = 1;
    }
}
// reduced-form correction loop (the kernels' fast path), for the bit-exactness test against hs_cart_correct
void hs_cart_correct_red(real* a, int B, int p, real lr, real momentum, real eps, int n_steps, int* flags) {
    const CartK k = cart_constants(p);
    const CartRed r = cart_reduced(k);
    for (int b = 0; b < B; ++b) {
        real ap = a[2 * (long)b + p], ao = a[2 * (long)b + 1 - p], pp = 0, po = 0;
        for (int s = 0; s < n_steps; ++s)
            if (cart_correct_step_red<true, true>(k, r, ap, ao, pp, po, lr, momentum, eps) && flags) flags[s] = 1;
        a[2 * (long)b + p] = ap; a[2 * (long)b + 1 - p] = ao;
    }
}
void hs_cart_thresholds(int p, real* t, real* A) {
    const CartRed r = cart_reduced(cart_constants(p));
    for (int j = 0; j < 3; ++j) { t[j] = r.t[j]; A[j] = r.A[j]; }
}
void hs_cart_resid(const real* a, real* eq, real* ineq, int B) {
    const CartK k = cart_constants(0);
    for (int b = 0; b < B; ++b) {
        real e, iq[6];
        cart_resid(k, a + 2 * (long)b, &e, iq);
        if (eq) eq[b] = e;
        if (ineq) for (int r = 0; r < 6; ++r) ineq[(long)b * 6 + r] = iq[r];
    }
}
void hs_cart_step(const real* s, const real* a, real* ns, int* done, int B) {
    for (int b = 0; b < B; ++b) done[b] = cart_step(s + 6 * (long)b, a + 2 * (long)b, ns + 6 * (long)b);
}

void hs_pen_complete(const real* obs, const real* z, real* a, int B) {
    for (int b = 0; b < B; ++b) pen_complete(pen_set_eq(obs + 5 * (long)b), z[b], a + 2 * (long)b);
}
void hs_pen_complete_bwd(const real* obs, const real* da, real* dz, int B) {
    for (int b = 0; b < B; ++b) dz[b] = pen_complete_bwd(pen_set_eq(obs + 5 * (long)b), da + 2 * (long)b);
}
void hs_pen_direction(const real* obs, const real* a, real* out, int B) {
    for (int b = 0; b < B; ++b) pen_direction(pen_set_eq(obs + 5 * (long)b), a + 2 * (long)b, out + 2 * (long)b);
}
void hs_pen_correct(const real* obs, real* a, int B, real lr, real momentum, real eps, int n_steps, int* flags) {
    for (int b = 0; b < B; ++b) {
        const PenEq e = pen_set_eq(obs + 5 * (long)b);
        real prev[2] = {0, 0};
        for (int s = 0; s < n_steps; ++s)
            if (pen_correct_step<true>(e, a + 2 * (long)b, prev, lr, momentum, eps) && flags) flags[s] = 1;
    }
}
void hs_pen_resid(const real* obs, const real* a, real* eq, real* ineq, int B) {
    for (int b = 0; b < B; ++b) {
        real e, q;
        pen_resid(pen_set_eq(obs + 5 * (long)b), a + 2 * (long)b, &e, &q);
        if (eq) eq[b] = e;
        if (ineq) ineq[b] = q;
    }
}
void hs_pen_step(const real* st, const real* a, real* nst, real* nobs, real* reward, int* done, int B) {
    for (int b = 0; b < B; ++b)
        done[b] = pen_step(st + 4 * (long)b, a + 2 * (long)b, nst + 4 * (long)b, nobs + 5 * (long)b, reward + b);
}

// action box (box_core.cuh), same element order as k_box
static void hs_box(int bwd, const real* x, const real* g, const real* lo_c, const real* hi_c, int W, const real* soc, int ld_soc,
                   int n_vol, real* out, int mode, int B) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < W; ++c) {
            real lo, hi;
            if (c >= W - n_vol) evopf_pe_bound1(soc[(long)b * ld_soc + (c - (W - n_vol))], &hi, &lo);
            else { lo = lo_c[c]; hi = hi_c[c]; }
            const long e = (long)b * W + c;
            const real xv = x ? x[e] : RC(0.0);
            out[e] = bwd ? box_bwd(xv, lo, hi, g[e], mode) : box_fwd(xv, lo, hi, mode);
        }
}
void hs_box_fwd(const real* x, const real* lo, const real* hi, int W, const real* soc, int ld_soc, int n_vol, real* out, int mode, int B) {
    hs_box(0, x, (const real*)0, lo, hi, W, soc, ld_soc, n_vol, out, mode, B);
}
void hs_box_bwd(const real* x, const real* g, const real* lo, const real* hi, int W, const real* soc, int ld_soc, int n_vol, real* gx,
                int mode, int B) {
    hs_box(1, x, g, lo, hi, W, soc, ld_soc, n_vol, gx, mode, B);
}

void hs_evopf_resid_bwd(const real* state, int ld, const real* y, const real* g_eq, const real* g_ineq, int mode, real* dy, int B) {
    for (int b = 0; b < B; ++b)
        evopf_resid_vjp(state + (long)b * ld, y + (long)b * 43, g_eq ? g_eq + (long)b * 28 : (const real*)0,
                        g_ineq ? g_ineq + (long)b * 58 : (const real*)0, mode, dy + (long)b * 43);
}

void hs_evopf_backward_penalty(const real* state, int ld, const real* y, const real* g, const real* w, const real* row_scale,
                               real* dz, real* dx, int B) {
    for (int b = 0; b < B; ++b)
        evopf_backward_penalty(state + (long)b * ld + 28, y + (long)b * 43, g + (long)b * 43, w, row_scale ? row_scale[b] : RC(1.0),
                               dz + (long)b * 14, dx ? dx + (long)b * 28 : (real*)0);
}

void hs_cart_resid_bwd(const real* a, const real* g_eq, const real* g_ineq, int mode, real* da, int B) {
    const CartK k = cart_constants(0);
    for (int b = 0; b < B; ++b)
        cart_resid_vjp(k, a + 2 * (long)b, g_eq ? g_eq + b : (const real*)0, g_ineq ? g_ineq + 6 * (long)b : (const real*)0, mode,
                       da + 2 * (long)b);
}

void hs_pen_resid_bwd(const real* obs, const real* a, const real* g_eq, const real* g_ineq, int mode, real* da, int B) {
    for (int b = 0; b < B; ++b)
        pen_resid_vjp(pen_set_eq(obs + 5 * (long)b), a + 2 * (long)b, g_eq ? g_eq + b : (const real*)0,
                      g_ineq ? g_ineq + b : (const real*)0, mode, da + 2 * (long)b);
}

int hs_sizeof_real() { return (int)sizeof(real); }
}
