/* Plain-C caller of the rpo_b200 C ABI: no torch, no Python — cudaMalloc'ed buffers, the case-14 set-point as input.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/c_abi_demo.c -o build/c_abi_demo \
 *       -Lrpo_b200 -lrpo_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/rpo_b200
 *
 * Projects B copies of the IEEE-14 operating point (loads of the case file, generator set-points of the published
 * solution) through completion + 10 correction steps, then checks what the reference's own loop
 * test looks at: max |eq_resid| after the completion and max ineq_dist before / after the correction (rpo/algo/rpo_ddpg.py:271-272).  Exit code 0 = all checks passed. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rpo_b200.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } } while (0)
#define RPO(call) do { int rc_ = (call); if (rc_ != RPO_OK) { fprintf(stderr, "%s: %d %s\n", #call, rc_, rpo_last_error()); return 3; } } while (0)

/* IEEE 14-bus loads in per unit (MATPOWER case14, bus columns PD / QD divided by baseMVA = 100) */
static const float PD[14] = {0.f, .217f, .942f, .478f, .076f, .112f, 0.f, 0.f, .295f, .09f, .035f, .061f, .135f, .149f};
static const float QD[14] = {0.f, .127f, .19f, -.039f, .016f, .075f, 0.f, 0.f, .166f, .058f, .018f, .016f, .058f, .05f};
/* z = [pg of the 4 PV generators | vm of the 5 generator buses | pe of the 5 batteries] */
static const float Z[14] = {0.4f, 0.f, 0.f, 0.f, 1.06f, 1.045f, 1.01f, 1.07f, 1.09f, 0.f, 0.f, 0.f, 0.f, 0.f};

int main(int argc, char** argv) {
    const long long B = argc > 1 ? atoll(argv[1]) : 4096;
    printf("rpo_b200 version %d, built for %s, B = %lld\n", rpo_version(), rpo_build_arch(), B);
    float *h_state = (float*)calloc((size_t)B * RPO_EVOPF_STATE_DIM, sizeof(float));
    float *h_z = (float*)malloc((size_t)B * RPO_EVOPF_Z_DIM * sizeof(float));
    for (long long b = 0; b < B; ++b) {
        float* s = h_state + b * RPO_EVOPF_STATE_DIM;
        const float load = 1.0f + 0.1f * (float)(b % 7 - 3) / 3.0f;            /* +-10 % load variation across rows */
        for (int i = 0; i < 14; ++i) { s[i] = PD[i] * load; s[14 + i] = QD[i] * load; }
        for (int i = 0; i < 5; ++i) s[28 + i] = 0.2f;                           /* state of charge */
        for (int i = 0; i < 14; ++i) h_z[b * 14 + i] = Z[i];
    }
    float *d_state, *d_z, *d_y, *d_yc, *d_eq, *d_ineq;
    int32_t *d_status, *d_flags;
    CK(cudaMalloc((void**)&d_state, (size_t)B * 57 * 4));
    CK(cudaMalloc((void**)&d_z, (size_t)B * 14 * 4));
    CK(cudaMalloc((void**)&d_y, (size_t)B * 43 * 4));
    CK(cudaMalloc((void**)&d_yc, (size_t)B * 43 * 4));
    CK(cudaMalloc((void**)&d_eq, (size_t)B * 28 * 4));
    CK(cudaMalloc((void**)&d_ineq, (size_t)B * 58 * 4));
    CK(cudaMalloc((void**)&d_status, (size_t)B * 4));
    CK(cudaMalloc((void**)&d_flags, 10 * 4));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    CK(cudaMemcpyAsync(d_state, h_state, (size_t)B * 57 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_z, h_z, (size_t)B * 14 * 4, cudaMemcpyHostToDevice, st));
    /* RPODDPG.process_action: complete_partial + grad_steps(train=True) with the settings of scripts/evopf_exp.py:29-31 */
    RPO(rpo_evopf_project(d_state, 57, d_z, d_y, d_yc, B, 1e-5f, 50, 1e-4f, 0.f, 1e-5f, 10, d_flags, d_status, st));
    /* the completion solves the equalities (eq_resid at y); the correction steps reduce the inequality violation (ineq at y_corr
     * vs at y) — the set-point's reactive dispatch violates the generator Q limits, so there is something to correct */
    float* d_ineq0;
    CK(cudaMalloc((void**)&d_ineq0, (size_t)B * 58 * 4));
    RPO(rpo_evopf_resid(d_state, 57, d_y, d_eq, d_ineq0, B, st));
    RPO(rpo_evopf_resid(d_state, 57, d_yc, NULL, d_ineq, B, st));
    float *h_y = (float*)malloc((size_t)B * 43 * 4), *h_eq = (float*)malloc((size_t)B * 28 * 4), *h_ineq = (float*)malloc((size_t)B * 58 * 4);
    float *h_ineq0 = (float*)malloc((size_t)B * 58 * 4);
    CK(cudaMemcpyAsync(h_ineq0, d_ineq0, (size_t)B * 58 * 4, cudaMemcpyDeviceToHost, st));
    int32_t *h_status = (int32_t*)malloc((size_t)B * 4), h_flags[10];
    CK(cudaMemcpyAsync(h_y, d_y, (size_t)B * 43 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_eq, d_eq, (size_t)B * 28 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_ineq, d_ineq, (size_t)B * 58 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double eq_max = 0, ineq_max = 0, ineq0_max = 0, iters = 0;
    long long bad = 0;
    for (long long b = 0; b < B; ++b) {
        for (int i = 0; i < 28; ++i) eq_max = fmax(eq_max, fabs(h_eq[b * 28 + i]));
        for (int i = 0; i < 58; ++i) { ineq_max = fmax(ineq_max, h_ineq[b * 58 + i]); ineq0_max = fmax(ineq0_max, h_ineq0[b * 58 + i]); }
        iters += h_status[b] & RPO_STATUS_ITERS_MASK;
        bad += (h_status[b] & (RPO_STATUS_NOT_CONVERGED | RPO_STATUS_SMALL_PIVOT)) != 0;
    }
    /* row 3 carries the unscaled case loads: slack generation of the published solution is 2.324 pu, |V_4| = 1.018 */
    const float* y3 = h_y + 3 * 43;
    printf("Newton iterations (mean) %.2f, flagged rows %lld, max|eq_resid(y)| %.3g, max ineq_resid %.3g -> %.3g after 10 steps\n",
           iters / B, bad, eq_max, ineq0_max, ineq_max);
    printf("row 3: pg_slack %.4f (published 2.3239), vm[3] %.4f (published 1.0177), loop-test flags:", y3[0], y3[10 + 3]);
    for (int k = 0; k < 10; ++k) printf(" %d", h_flags[k]);
    printf("\n");
    const int ok = bad == 0 && eq_max < 2e-5 && ineq_max < 0.5 * ineq0_max && fabs(y3[0] - 2.3239f) < 2e-3 && fabs(y3[13] - 1.0177f) < 1e-3;
    /* argument errors come back as codes, never as crashes */
    const int rc = rpo_evopf_project(d_state, 20, d_z, d_y, d_yc, B, 1e-5f, 50, 1e-4f, 0.f, 1e-5f, 10, d_flags, d_status, st);
    printf("bad leading dimension -> rc %d (%s)\n", rc, rpo_last_error());
    printf(ok && rc == RPO_ERR_INVALID_ARG ? "OK\n" : "FAILED\n");
    return ok && rc == RPO_ERR_INVALID_ARG ? 0 : 1;
}
