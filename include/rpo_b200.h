/*
 * rpo_b200 — C ABI of the B200-native RPO projection hot path.
 *
 * The reference (wadx2019/rpo) is pure Python/torch and has no FFI; its de-facto plug-in boundary is the
 * duck-typed "hard-constraint protocol" that rpo.algo calls on an environment plus the grad_steps loop
 * (SURVEY.md §8b).  Every entry point below replaces one of those methods; the reference file:line it
 * replaces is cited on each declaration (paths relative to the reference root).
 *
 * Conventions
 *   - all array arguments are DEVICE pointers to row-major float32 (int32 for status/flags), unless the
 *     function name ends in _host; rows are contiguous except `state`, which carries a leading dimension
 *     `ld_state` (in floats) so that a (B,57) observation tensor or a (B,28) [pd|qd] slice can be passed;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream, allocates
 *     nothing and never synchronises;
 *   - alignment: 4 bytes is enough for the EVOPF entry points, the action box and rpo_cart_correct / rpo_cart_project
 *     (they take vector / TMA paths only when the pointers allow); the single-pass cart / pendulum entry points need their
 *     (B,2) arrays 8-byte aligned and the pendulum's (B,4) states 16-byte aligned and return RPO_ERR_INVALID_ARG otherwise;
 *   - return value 0 = ok, negative = error (RPO_ERR_*); rpo_last_error() gives a message;
 *   - a pointer documented "nullable" may be NULL to skip that output.
 */
#ifndef RPO_B200_H
#define RPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPO_OK 0
#define RPO_ERR_INVALID_ARG (-1)
#define RPO_ERR_CUDA (-2)

/* per-instance status word written by the completion kernels (optional output) */
#define RPO_STATUS_ITERS_MASK 0xff       /* Newton iterations used */
#define RPO_STATUS_NOT_CONVERGED 0x100   /* ||delta||_2 >= tol after max_iters (reference: `converged`, evopf.py:837, discarded) */
#define RPO_STATUS_SMALL_PIVOT 0x200     /* |pivot| < 1e-6 in the static-order LU (reference: torch.inverse would raise on singular) */

/* EVOPF-v0 dimensions (evopf.py:270-337) */
#define RPO_EVOPF_STATE_DIM 57
#define RPO_EVOPF_X_DIM 28
#define RPO_EVOPF_Y_DIM 43
#define RPO_EVOPF_Z_DIM 14
#define RPO_EVOPF_EQ_NUM 28
#define RPO_EVOPF_INEQ_NUM 58

int rpo_version(void);
const char* rpo_last_error(void);
/* name of the compiled SM target, e.g. "sm_100a" */
const char* rpo_build_arch(void);

/* ------------------------------------------------------------------ EVOPF-v0 ---------------------------------- */

/* EVOPFEnv.complete_partial -> PFFunctionFn.forward            rpo/env/electrical_grid/evopf.py:724-727, 789-855
 * state: rows hold [pd(14) | qd(14) | ...] (only the first 28 floats of a row are read); z (B,14); y (B,43) out.
 * tol/max_iters: Newton stop rule (reference: 1e-5 / 50, evopf.py:786).  status: nullable (B) int32. */
int rpo_evopf_complete_fwd(const float* state, int ld_state, const float* z, float* y, int64_t B, float tol,
                           int max_iters, int32_t* status, void* stream);
/* rpo_evopf_complete_fwd finishes with a mixed-precision refinement: once every instance of a CTA has ||delta|| < 1e-3 the
 * last update is a chord step on the float32 LU whose right-hand side is the residual accumulated in float64, which
 * removes the float32 evaluation noise from the converged point (same stop rule as evopf.py:826-835: the applied update
 * must have ||delta||_2 < tol).  The _plain variant is the float32 Newton sequence alone — the one rpo_evopf_project
 * runs before its correction steps; use it where bit-identity with rpo_evopf_project's `y` output is wanted. */
int rpo_evopf_complete_fwd_plain(const float* state, int ld_state, const float* z, float* y, int64_t B, float tol,
                                 int max_iters, int32_t* status, void* stream);

/* PFFunctionFn.backward (GRG / implicit-function backward)      rpo/env/electrical_grid/evopf.py:857-910
 * y (B,43) = output of the forward; dl_dy (B,43); dl_dz (B,14) out; dl_dx (B,28) out, nullable.
 * The Jacobian is recomputed from y (the reference saves (B,28,43)+(B,22,22) tensors, evopf.py:851-853). */
int rpo_evopf_complete_bwd(const float* y, const float* dl_dy, float* dl_dz, float* dl_dx, int64_t B, void* stream);

/* EVOPFEnv.ineq_partial_grad (reduced correction direction)     rpo/env/electrical_grid/evopf.py:596-612
 * state rows: [pd|qd|soc(5)|...] (first 33 floats read); y (B,43); out (B,43). */
int rpo_evopf_direction(const float* state, int ld_state, const float* y, float* out, int64_t B, void* stream);

/* RPODDPG.grad_steps / RPOSAC.grad_steps, corr_mode == 0        rpo/algo/rpo_ddpg.py:266-305, rpo/algo/rpo_sac.py:280-319
 * Runs exactly n_steps correction steps on every instance (y_out may alias y_in).  unsat_flags (n_steps int32,
 * nullable) is zeroed by the call and unsat_flags[k] is set to 1 when ANY instance fails the loop test
 * (max|eq_resid| > corr_eps or max ineq_dist > corr_eps) before step k — the host uses it to reproduce the
 * reference's global-max loop exit (see rpo_b200/algo.py). */
int rpo_evopf_correct(const float* state, int ld_state, const float* y_in, float* y_out, int64_t B, float lr,
                      float momentum, float corr_eps, int n_steps, int32_t* unsat_flags, void* stream);

/* RPODDPG.process_action = complete_partial + grad_steps fused  rpo/algo/rpo_ddpg.py:72-77 (no-grad call sites :100,:106,:330)
 * y (B,43) nullable: completion output before correction; y_corr (B,43): corrected action. */
int rpo_evopf_project(const float* state, int ld_state, const float* z, float* y, float* y_corr, int64_t B, float tol,
                      int max_iters, float lr, float momentum, float corr_eps, int n_steps, int32_t* unsat_flags,
                      int32_t* status, void* stream);

/* EVOPFEnv.eq_resid / ineq_resid                                 rpo/env/electrical_grid/evopf.py:520-546, 548-563 (+Battery.ineq_resid :104-121)
 * eq (B,28) nullable; ineq (B,58) nullable (ineq_dist = max(ineq, 0)). */
int rpo_evopf_resid(const float* state, int ld_state, const float* y, float* eq, float* ineq, int64_t B, void* stream);

/* Autograd of EVOPFEnv.eq_resid / ineq_resid / ineq_dist w.r.t. the action   rpo/env/electrical_grid/evopf.py:520-567
 * The reference's residual functions are plain torch code and its actor loss differentiates through them
 * (rpo/algo/rpo_ddpg.py:312-319, rpo/algo/rpo_sac.py:326-334: loss += nju(ineq_dist(state, actions))).
 * dl_dy (B,43) out = (d eq_resid/dy)^T g_eq + (d ineq/dy)^T (g_ineq . m); g_eq (B,28) / g_ineq (B,58) nullable upstream
 * gradients; ineq_mode 0: ineq_resid (m = 1), 1: ineq_dist = clamp(ineq_resid, 0) (m = [ineq_resid >= 0], torch.clamp's
 * gradient).  The eq part is the TRUE derivative (+I for pe, evopf.py:532), not eq_jac's -I (evopf.py:639-640). */
int rpo_evopf_resid_bwd(const float* state, int ld_state, const float* y, const float* g_eq, const float* g_ineq,
                        int ineq_mode, float* dl_dy, int64_t B, void* stream);

/* PFFunctionFn.backward with the actor loss's inequality penalty folded in  rpo/algo/rpo_ddpg.py:307-324 + evopf.py:857-910
 * Computes dl_dz / dl_dx of  L(y) + row_scale_b * sum_r w_r * relu(ineq_resid_r(state, y))  given dl_dy = dL/dy:
 * the +-1 row selector of the violated rows (evopf.py:663-707) is added to dl_dy in registers before the transposed
 * solve, so the (B,58) ineq_dist gradient is never materialised.  w (58) = Dual.weight (rpo/algo/model/dual.py:47-65);
 * row_scale (B) nullable (NULL = 1): upstream gradient of the per-row penalty (1/B for loss.mean()). */
int rpo_evopf_complete_bwd_penalty(const float* state, int ld_state, const float* y, const float* dl_dy, const float* w,
                                   const float* row_scale, float* dl_dz, float* dl_dx, int64_t B, void* stream);

/* The action box fused into the completion / projection and into the GRG backward   rpo/algo/model/utils.py:40-51,75-101,
 * rpo/algo/model/policy.py:30-31 (box(tanh(x), state)) followed by rpo/algo/rpo_ddpg.py:72-77 / evopf.py:724-727.
 * u (B,14) = the policy's raw output; z = box(u) against the state-dependent EVOPF partial box (box_lo / box_hi: the 14
 * constant bounds, whose last 5 entries are replaced by the battery bounds of the row's state of charge, exactly as
 * rpo_box_fwd with n_vol = 5); box_mode 0-3 as rpo_box_fwd.  Results equal rpo_box_fwd -> rpo_evopf_project /
 * rpo_evopf_complete_fwd bit for bit; the backward returns dL/du = dL/dz . d box / d u (rpo_box_bwd's arithmetic) and
 * optionally folds the actor-loss penalty in as rpo_evopf_complete_bwd_penalty does (w nullable: no penalty). */
int rpo_evopf_project_squashed(const float* state, int ld_state, const float* u, const float* box_lo, const float* box_hi,
                               int box_mode, float* y, float* y_corr, int64_t B, float tol, int max_iters, float lr,
                               float momentum, float corr_eps, int n_steps, int32_t* unsat_flags, int32_t* status,
                               void* stream);
int rpo_evopf_complete_fwd_squashed(const float* state, int ld_state, const float* u, const float* box_lo,
                                    const float* box_hi, int box_mode, float* y, int64_t B, float tol, int max_iters,
                                    int32_t* status, void* stream);
int rpo_evopf_complete_bwd_squashed(const float* state, int ld_state, const float* y, const float* dl_dy, const float* u,
                                    const float* box_lo, const float* box_hi, int box_mode, const float* w,
                                    const float* row_scale, float* dl_du, float* dl_dx, int64_t B, void* stream);

/* EVOPFEnv.step, batched row-wise                                rpo/env/electrical_grid/evopf.py:348-366 (+Battery.step :74-102, obj_fn :509-518)
 * state (B,57); y (B,43) action; next_exo (B,52) = next [pd|qd](28) + next price window (24) from the data loaders;
 * next_state (B,57) out; reward (B) out; eq_viol (B,28) nullable; ineq_viol (B,58) nullable (both at the pre-step state). */
int rpo_evopf_step(const float* state, const float* y, const float* next_exo, float* next_state, float* reward,
                   float* eq_viol, float* ineq_viol, int64_t B, void* stream);

/* ------------------------------------------------------------------ Safe Cartpole ----------------------------- */
/* partial_idx in {0,1}: which force is the policy output (CartSafeEnv.partial_actions is drawn at random at
 * construction, rpo/env/classic_control/cartpole.py:117). */

/* CartSafeEnv.complete_partial                                   rpo/env/classic_control/cartpole.py:369-373 ; z (B,1) -> a (B,2) */
int rpo_cart_complete_fwd(const float* z, float* a, int64_t B, int partial_idx, void* stream);
/* autograd of the above: dl_da (B,2) -> dl_dz (B,1) */
int rpo_cart_complete_bwd(const float* dl_da, float* dl_dz, int64_t B, int partial_idx, void* stream);
/* CartSafeEnv.ineq_partial_grad                                  cartpole.py:396-408 */
int rpo_cart_direction(const float* a, float* out, int64_t B, int partial_idx, void* stream);
/* grad_steps on CartSafeEnv                                      rpo/algo/rpo_ddpg.py:266-305 */
int rpo_cart_correct(const float* a_in, float* a_out, int64_t B, int partial_idx, float lr, float momentum,
                     float corr_eps, int n_steps, int32_t* unsat_flags, void* stream);
/* fused complete + correct */
int rpo_cart_project(const float* z, float* a, float* a_corr, int64_t B, int partial_idx, float lr, float momentum,
                     float corr_eps, int n_steps, int32_t* unsat_flags, void* stream);
/* CartSafeEnv.eq_resid / ineq_resid                              cartpole.py:375-379 ; eq (B,1), ineq (B,6), nullable */
int rpo_cart_resid(const float* a, float* eq, float* ineq, int64_t B, void* stream);
/* autograd of CartSafeEnv.eq_resid / ineq_resid / ineq_dist w.r.t. the action   cartpole.py:375-387 (consumed by
 * rpo_ddpg.py:312-319); g_eq (B,1), g_ineq (B,6) nullable; ineq_mode as rpo_evopf_resid_bwd; a may be NULL when g_ineq is. */
int rpo_cart_resid_bwd(const float* a, const float* g_eq, const float* g_ineq, int ineq_mode, float* dl_da, int64_t B, void* stream);
/* CartSafeEnv.step, batched row-wise                             cartpole.py:166-229 ; state (B,6), a (B,2) ->
 * next_state (B,6), done (B) int32, eq_viol (B,1) / ineq_viol (B,6) nullable (unclipped action). */
int rpo_cart_step(const float* state, const float* a, float* next_state, int32_t* done, float* eq_viol, float* ineq_viol,
                  int64_t B, void* stream);

/* ------------------------------------------------------------------ Spring Pendulum --------------------------- */
/* obs (B,5) = [cos th, sin th, thdot, l, ldot]; the policy output is fx (partial_actions = [0], pendulum.py:48).
 * Row-wise semantics: each row is evaluated as the reference evaluates a batch of one. */

/* SpringPendulumEnv.complete_partial (+ set_eq)                  rpo/env/classic_control/pendulum.py:256-288 */
int rpo_pen_complete_fwd(const float* obs, const float* z, float* a, int64_t B, void* stream);
int rpo_pen_complete_bwd(const float* obs, const float* dl_da, float* dl_dz, int64_t B, void* stream);
/* SpringPendulumEnv.ineq_partial_grad                            pendulum.py:331-343 */
int rpo_pen_direction(const float* obs, const float* a, float* out, int64_t B, void* stream);
int rpo_pen_correct(const float* obs, const float* a_in, float* a_out, int64_t B, float lr, float momentum,
                    float corr_eps, int n_steps, int32_t* unsat_flags, void* stream);
int rpo_pen_project(const float* obs, const float* z, float* a, float* a_corr, int64_t B, float lr, float momentum,
                    float corr_eps, int n_steps, int32_t* unsat_flags, void* stream);
/* SpringPendulumEnv.eq_resid / ineq_resid                        pendulum.py:298-303 ; eq (B,1), ineq (B,1), nullable */
int rpo_pen_resid(const float* obs, const float* a, float* eq, float* ineq, int64_t B, void* stream);
/* autograd of SpringPendulumEnv.eq_resid / ineq_resid / ineq_dist w.r.t. the action   pendulum.py:296-311 */
int rpo_pen_resid_bwd(const float* obs, const float* a, const float* g_eq, const float* g_ineq, int ineq_mode, float* dl_da,
                      int64_t B, void* stream);
/* SpringPendulumEnv.step, batched row-wise                       pendulum.py:80-128 ; st (B,4) = (th, thdot, l, ldot), a (B,2) ->
 * next_st (B,4), next_obs (B,5), reward (B), done (B) int32, eq_viol (B,1) / ineq_viol (B,1) nullable. */
int rpo_pen_step(const float* st, const float* a, float* next_st, float* next_obs, float* reward, int32_t* done,
                 float* eq_viol, float* ineq_viol, int64_t B, void* stream);

/* ------------------------------------------------------------------ action box ------------- */
/* BoxConstraint: squash / clip of a policy output into the (state-dependent) action box, one launch, no host round
 * trip.  Replaces rpo/algo/model/utils.py:40-51 (__call__), :92-104 (clip), :106-118 (action_scale), :75-90 (update_box)
 * and, for EVOPF, EVOPFEnv.update rpo/env/electrical_grid/evopf.py:769-783 + Battery.update_bound :127-143, which the
 * reference evaluates through numpy on the host on every actor forward (D2H + 4 x H2D).
 * x, out (B, width); lo / hi (width) device arrays with the constant bounds; the LAST n_vol columns (0, or 5 = the EVOPF
 * batteries) take their bounds from soc (B rows, leading dimension ld_soc floats; pass state + 28, ld 57):
 * hi = min(0.2, 0.8 - soc) / 0.9, lo = 0.9 max(-0.2, 0.1 - soc) in float64, rounded to float32.
 * mode: 0 scale*x+base (style "tanh") | 1 scale*tanh(x)+base (policy.py:30-31 fused) | 2 clip | 3 (hi-lo)*x+lo (style
 * "sigmoid") | 4 scale | 5 lo | 6 hi (x may be NULL for 4-6).  out may alias x. */
int rpo_box_fwd(const float* x, const float* lo, const float* hi, int width, const float* soc, int ld_soc, int n_vol,
                float* out, int mode, int64_t B, void* stream);
/* grad_x = grad_out . d out / d x for modes 0-3 (clip: torch.clip's mask lo <= x <= hi).  No gradient reaches the state,
 * as in the reference (its bounds leave the autograd graph through numpy). */
int rpo_box_bwd(const float* x, const float* grad_out, const float* lo, const float* hi, int width, const float* soc,
                int ld_soc, int n_vol, float* grad_x, int mode, int64_t B, void* stream);

/* ------------------------------------------------------------------ host-buffer helpers ----------------------- */
/* Copy the first `ncols` floats of each of `nrows` rows between a HOST matrix (pinned for asynchrony; leading dimension
 * ld_host floats) and a DEVICE matrix (leading dimension ld_dev floats): one strided DMA (cudaMemcpy2DAsync) on `stream`.
 * Used by the host-buffer pipeline to upload only the [pd|qd|soc] columns (33 of 57) of the reference's state rows
 * (rpo/algo/rpo_ddpg.py:93,168: states live on the host between steps). */
int rpo_copy_rows_h2d(float* dst_dev, int ld_dev, const float* src_host, int ld_host, int ncols, int64_t nrows, void* stream);
int rpo_copy_rows_d2h(float* dst_host, int ld_host, const float* src_dev, int ld_dev, int ncols, int64_t nrows, void* stream);

/* ------------------------------------------------------------------ measurement helpers ----------------------- */
/* Dependency-free FP32 FMA micro-kernel used by bench.py to measure the FP32 roofline denominator on the box
 * (SURVEY.md §8d).  Launches `blocks` x `threads`, `iters` x 8 independent FMA chains x 4 per thread; out (blocks*threads) floats.
 * flops executed = 2 * 32 * iters * blocks * threads. */
int rpo_fp32_peak_probe(float* out, int blocks, int threads, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RPO_B200_H */
