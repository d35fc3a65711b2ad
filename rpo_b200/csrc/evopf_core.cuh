// Per-instance drivers of the EVOPF hot path on top of the generated math (evopf_gen.cuh).
// `__host__ __device__`: the CUDA kernels (evopf_kernels.cu) are the product; tests/hostsim compiles the
// same code with g++ to validate it against the oracle on machines without a GPU.
#pragma once

#include "rpo_math.cuh"
#include "evopf_gen.cuh"

// per-instance status word (optional int32 output of the C-ABI; the reference computes `converged`
// and discards it, evopf.py:818,837 — we surface it)
#define RPO_ST_ITERS_MASK 0xff          // Newton iterations used
#define RPO_ST_NOT_CONVERGED 0x100      // ||delta|| >= tol after max_iters
#define RPO_ST_SMALL_PIVOT 0x200        // |pivot| < 1e-6 met in a static-order LU
#define RPO_ST_NONFINITE 0x400          // NaN/Inf in the result

#define RPO_PIVOT_FLOOR RC(1.0e-6)

struct RpoVoteSelf {
    RPO_HDM bool operator()(bool p) const { return p; }
};

// Mixed-precision refinement of the completion.  Once every lane of the voting group has ||delta||_2 < RPO_REFINE_SWITCH the
// float32 Newton iteration is in its quadratic regime and its sparse LU is good enough to be REUSED: the last update is
// a chord step  delta = LU^{-1} r(y)  whose right-hand side is the power-flow residual accumulated in float64 (FP64 pipe,
// otherwise idle).  The stop rule is the reference's (evopf.py:826-835: iterate until the applied update has
// ||delta||_2 < tol); what changes is how the last update is computed: no Jacobian rebuild / refactorisation, and the
// float32 evaluation noise of the residual no longer ends up in the converged point (equality residual of the result,
// evaluated in float64 over 32k D1 instances: max 4.9e-6 / median 1.6e-6 against 8.0e-6 / 1.8e-6 for the float32
// reference).  The closed-form tail (qg, slack pg; evopf.py:844-848) comes from the same float64 flows.
// Returns whether every lane of the group has converged; `it` counts applied updates (Newton + chord).
#ifndef RPO_REFINE_SWITCH
#define RPO_REFINE_SWITCH RC(1.0e-3)
#endif
#ifndef RPO_REFINE_MAX_PASSES
#define RPO_REFINE_MAX_PASSES 1
#endif
template <class Vote>
RPO_HD bool evopf_refine(const real* x, real* y, const EvopfPt& pt, const real* lu, const real* ipiv, real tol, int max_iters,
                         int& it, bool& conv, Vote any) {
    double P[14], Q[14];
    real r[EVOPF_NN];
    bool all = false;
    for (int pass = 0; pass < RPO_REFINE_MAX_PASSES && !all && it < max_iters; ++pass) {   // uniform over the voting group
        evopf_flows_f64(y, P, Q);
        evopf_newton_rhs_f64(x, y, P, Q, r);
        evopf_solve(lu, ipiv, r);
        const real n2 = evopf_newton_update(r, y);
        ++it;
        conv = sqrt(n2) < tol;
        all = !any(!conv);
    }
    evopf_tail_refined(x, y, pt, P, Q, r);
    return all;
}

// Equality completion = PFFunctionFn.forward (evopf.py:789-855): flat start, Newton on the 22 power-flow
// equations, closed-form qg / slack pg.  `any` is a vote over the lanes that iterate together (the CTA on
// the GPU; the reference iterates a chunk of 256 together, evopf.py:821-835): everybody keeps iterating
// until every member has ||delta||_2 < tol, exactly as the reference does inside a chunk.
// REFINE = true: finish with the mixed-precision refinement above (the stand-alone completion: rpo_evopf_complete_fwd,
// EVOPFEnv.complete_partial).  REFINE = false: plain float32 Newton, the very sequence the fused projection kernel runs
// before its correction steps (evopf_project_loop; there the completion point is an intermediate whose equality
// residual the 10 linearised correction steps raise to ~1e-3 anyway, so the refinement would buy nothing).
template <bool REFINE, class Vote>
RPO_HD int evopf_complete(const real* x, const real* z, real* y, real tol, int max_iters, Vote any) {
    evopf_init(z, y);
    int it = 0;
    bool conv = false, tail_ok = false;
    real minpiv = RC(3.0e38);
    while (true) {
        EvopfPt pt;
        real lu[EVOPF_NLU], ipiv[EVOPF_NN], r[EVOPF_NN];
        evopf_eval_full(y, pt);
        minpiv = FMIN(minpiv, evopf_factor(pt, lu, ipiv));
        evopf_newton_rhs(x, y, pt, r);
        evopf_solve(lu, ipiv, r);
        const real dn = sqrt(evopf_newton_update(r, y));
        ++it;
        conv = dn < tol;
        tail_ok = false;
        bool all;
        if (REFINE && !any(!(dn < RPO_REFINE_SWITCH)) && it < max_iters) {
            all = evopf_refine(x, y, pt, lu, ipiv, tol, max_iters, it, conv, any);
            tail_ok = true;
        } else {
            all = !any(!conv);
        }
        if (all || it >= max_iters) break;
    }
    if (!tail_ok) {                                    // no refinement pass ran last (max_iters hit / tol above the switch)
        EvopfPt pt;
        evopf_eval_flows(y, pt);
        evopf_tail(x, y, pt);
    }
    int st = it & RPO_ST_ITERS_MASK;
    if (!conv) st |= RPO_ST_NOT_CONVERGED;
    if (!(minpiv >= RPO_PIVOT_FLOOR)) st |= RPO_ST_SMALL_PIVOT;
    return st;
}

// action_step = lr*dir + momentum*prev;  y -= action_step  (rpo_ddpg.py:280-283).  Without momentum this is one FMA per
// component (the reference rounds lr*dir before subtracting: <= 0.5 ulp of y apart).
template <bool MOMENTUM>
RPO_HD void evopf_apply_step(real* y, const real* dir, real* prev, real lr, real momentum) {
    for (int i = 0; i < 43; ++i) {
        if (MOMENTUM) {
            const real s = FMA(momentum, prev[i], lr * dir[i]);
            prev[i] = s;
            y[i] = y[i] - s;
        } else {
            y[i] = FMA(-lr, dir[i], y[i]);
        }
    }
}

// One step of the post-correction loop body (rpo_ddpg.py:274,280-283) for corr_mode == 0:
//   dir = ineq_partial_grad(state, y);  step = lr*dir + momentum*prev;  y -= step;  prev = step.
// Returns whether this instance fails the loop test *before* the step (|eq_resid| or ineq_dist above eps).
template <bool MOMENTUM>
RPO_HD bool evopf_correct_step(const real* x, const real* pe_max, const real* pe_min, real* y, real* prev, real lr,
                               real momentum, real corr_eps, real* minpiv) {
    real dir[43], eqmax, ineqmax;
    *minpiv = evopf_direction(x, pe_max, pe_min, y, dir, &eqmax, &ineqmax);
    evopf_apply_step<MOMENTUM>(y, dir, prev, lr, momentum);
    return (eqmax > corr_eps) || (ineqmax > corr_eps);
}

// Backward of the residual functions w.r.t. the action = what autograd computes for the reference's plain-torch
// eq_resid / ineq_resid / ineq_dist (evopf.py:520-567), which the actor loss differentiates (rpo_ddpg.py:312-319,
// rpo_sac.py:326-334):  dy = (d eq_resid/dy)^T g_eq + (d ineq/dy)^T (g_ineq . m),  m = [resid >= 0] for ineq_dist
// (ineq_mode 1), all ones for ineq_resid (ineq_mode 0).  g_eq / g_ineq may be null.  x = [pd|qd|soc] row.
RPO_HD void evopf_resid_vjp(const real* x, const real* y, const real* g_eq, const real* g_ineq, int ineq_mode, real* dy) {
    for (int i = 0; i < 43; ++i) dy[i] = RC(0.0);
    if (g_eq) {
        EvopfPt pt;
        evopf_eval_full(y, pt);
        evopf_eq_vjp(y, pt, g_eq, g_eq + 14, dy);
    }
    if (g_ineq) {
        real pe_max[5], pe_min[5];
        evopf_pe_bounds(x + 28, pe_max, pe_min);
        evopf_ineq_vjp(pe_max, pe_min, y, g_ineq, RC(1.0), ineq_mode, dy);
    }
}

// GRG backward with the inequality penalty of the actor loss folded in (rpo_ddpg.py:312-319: loss = -Q + nju(ineq_dist)):
//   dL/dz, dL/dx of  L(y) + scale * w . relu(ineq_resid(state, y))  given g = dL/dy  — the penalty's +-1 row selector is
// added to g in registers before the transposed solve, so the (B,58) ineq_dist gradient never exists in HBM.
RPO_HD real evopf_backward_penalty(const real* soc, const real* y, const real* g, const real* w, real scale, real* dz, real* dx) {
    real gt[43], pe_max[5], pe_min[5];
    for (int i = 0; i < 43; ++i) gt[i] = g[i];
    evopf_pe_bounds(soc, pe_max, pe_min);
    evopf_ineq_vjp(pe_max, pe_min, y, w, scale, 1, gt);
    return evopf_backward(y, gt, dz, dx);
}

// Fused projection = complete_partial followed by n_steps >= 1 correction steps (RPODDPG.process_action,
// rpo_ddpg.py:72-77) as ONE loop whose body holds a single copy of the two big shared pieces (point evaluation and
// sparse LU): a Newton iteration and a correction step differ only in what follows the factorisation.  Besides
// removing the completion's separate tail evaluation (the first correction pass evaluates the final Newton point
// anyway), this halves the instruction footprint of the fused kernel, which is instruction-fetch sensitive
// (DESIGN.md §5).  The phase switch is uniform over the voting group, so there is no divergence.  (Sharing the plain
// triangular solve between the phases as well — evopf_direction_pre / evopf_solve / evopf_direction_fin — was measured:
// 2 % fewer instructions, more spills, 7 % slower; the two phases keep their own inlined copy.)
//   any(pred)        vote over the lanes that iterate together
//   on_y(y, status)  called once, by every lane, when the completion output is final (before the first step)
//   on_flag(k, unsat) called once per correction step with this lane's loop-test result
//   scratch          >= EVOPF_SCRATCH reals of per-lane storage for values that are cold during the factorisation
//                    (violation signs d[43], battery bounds): shared memory on the device, a local array on the host
#define EVOPF_SCRATCH 53
template <bool MOMENTUM, class Vote, class OnY, class OnFlag>
RPO_HD void evopf_project_loop(const real* x, const real* z, real* y, real* scratch, real tol, int max_iters, real lr,
                               real momentum, real corr_eps, int n_steps, Vote any, OnY on_y, OnFlag on_flag) {
    real* d = scratch;
    real* pe_max = scratch + 43;
    real* pe_min = scratch + 48;
    evopf_pe_bounds(x + 28, pe_max, pe_min);
    real prev[MOMENTUM ? 43 : 1];
    if (MOMENTUM) {
        for (int i = 0; i < 43; ++i) prev[i] = RC(0.0);
    }
    evopf_init(z, y);
    int it = 0, k = 0;
    int newton = 1;
    bool conv = false, first = true;
    real minpiv = RC(3.0e38);
    while (true) {
        RPO_OPAQUE(newton);
        EvopfPt pt;
        real lu[EVOPF_NLU], ipiv[EVOPF_NN], eqmax = RC(0.0), ineqmax = RC(0.0);
        evopf_eval_full(y, pt);
        if (!newton) {
            if (first) {
                evopf_tail(x, y, pt);
                int st = it & RPO_ST_ITERS_MASK;
                if (!conv) st |= RPO_ST_NOT_CONVERGED;
                if (!(minpiv >= RPO_PIVOT_FLOOR)) st |= RPO_ST_SMALL_PIVOT;
                on_y(y, st);
                first = false;
            }
            eqmax = evopf_eq_resid(x, y, pt, (real*)0);
            ineqmax = evopf_ineq(pe_max, pe_min, y, (real*)0, d);
        }
        const real mp = evopf_factor(pt, lu, ipiv);
        if (newton) {
            real r[EVOPF_NN];
            minpiv = FMIN(minpiv, mp);
            evopf_newton_rhs(x, y, pt, r);
            evopf_solve(lu, ipiv, r);
            const real n2 = evopf_newton_update(r, y);
            ++it;
            conv = sqrt(n2) < tol;
            if (!any(!conv) || it >= max_iters) newton = 0;
        } else {
            real dir[43];
            evopf_direction_post(y, pt, lu, ipiv, d, dir);
            evopf_apply_step<MOMENTUM>(y, dir, prev, lr, momentum);
            on_flag(k, (eqmax > corr_eps) || (ineqmax > corr_eps));
            if (++k >= n_steps) break;
        }
    }
}
