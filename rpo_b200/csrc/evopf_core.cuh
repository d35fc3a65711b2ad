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

// Equality completion = PFFunctionFn.forward (evopf.py:789-855): flat start, Newton on the 22 power-flow
// equations, closed-form qg / slack pg.  `any` is a vote over the lanes that iterate together (the warp on
// the GPU; the reference iterates a chunk of 256 together, evopf.py:821-835): everybody keeps iterating
// until every member has ||delta||_2 < tol, exactly as the reference does inside a chunk.
template <class Vote>
RPO_HD int evopf_complete(const real* x, const real* z, real* y, real tol, int max_iters, Vote any) {
    EvopfPt pt;
    evopf_init(z, y);
    int it = 0;
    bool conv = false;
    real minpiv = RC(3.0e38);
    while (true) {
        real mp;
        const real n2 = evopf_newton_iter(x, y, pt, &mp);
        minpiv = FMIN(minpiv, mp);
        ++it;
        conv = sqrt(n2) < tol;
        if (!any(!conv) || it >= max_iters) break;
    }
    evopf_eval_flows(y, pt);
    evopf_tail(x, y, pt);
    int st = it & RPO_ST_ITERS_MASK;
    if (!conv) st |= RPO_ST_NOT_CONVERGED;
    if (!(minpiv >= RPO_PIVOT_FLOOR)) st |= RPO_ST_SMALL_PIVOT;
    return st;
}

// One step of the post-correction loop body (rpo_ddpg.py:274,280-283) for corr_mode == 0:
//   dir = ineq_partial_grad(state, y);  step = lr*dir + momentum*prev;  y -= step;  prev = step.
// Returns whether this instance fails the loop test *before* the step (|eq_resid| or ineq_dist above eps).
template <bool MOMENTUM>
RPO_HD bool evopf_correct_step(const real* x, const real* pe_max, const real* pe_min, real* y, real* prev, real lr,
                               real momentum, real corr_eps, real* minpiv) {
    real dir[43], eqmax, ineqmax;
    *minpiv = evopf_direction(x, pe_max, pe_min, y, dir, &eqmax, &ineqmax);
    for (int i = 0; i < 43; ++i) {
        real s = lr * dir[i];
        if (MOMENTUM) {
            s = s + momentum * prev[i];
            prev[i] = s;
        }
        y[i] = y[i] - s;
    }
    return (eqmax > corr_eps) || (ineqmax > corr_eps);
}
