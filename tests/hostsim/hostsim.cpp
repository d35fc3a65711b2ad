// Host-side simulator of the per-instance device code — TEST INFRASTRUCTURE ONLY.
// Compiles rpo_b200/csrc/{evopf,classic}_core.cuh with g++ (RPO_REAL = float or double) so the CPU test-suite
// can check the generated math against the oracle where no GPU exists.  Never loaded by rpo_b200/.
#include <stdint.h>
#include "../../rpo_b200/csrc/evopf_core.cuh"
#include "../../rpo_b200/csrc/classic_core.cuh"

extern "C" {

void hs_evopf_complete(const real* state, int ld, const real* z, real* y, int B, real tol, int max_iters, int* status) {
    for (int b = 0; b < B; ++b) {
        int st = evopf_complete(state + (long)b * ld, z + (long)b * 14, y + (long)b * 43, tol, max_iters, RpoVoteSelf());
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

int hs_sizeof_real() { return (int)sizeof(real); }
}
