// Scalar-type and intrinsic shims shared by the generated per-instance math (evopf_gen.cuh), the
// hand-written per-instance drivers (evopf_core.cuh, classic_core.cuh) and the kernels.
//
// The per-instance code is `__host__ __device__` so that the very same source can be compiled by g++
// into a host-side simulator (tests/hostsim) — used ONLY by the CPU test-suite to check the generated
// math against the oracle in a container without a GPU.  The product path is the CUDA build.
#pragma once

#include <math.h>

#ifndef RPO_REAL
#define RPO_REAL float
#endif
typedef RPO_REAL real;

#if defined(__CUDACC__)
#define RPO_HD __host__ __device__ __forceinline__
#define RPO_HDM __host__ __device__ __forceinline__
#define RPO_D __device__ __forceinline__
#else
#define RPO_HD static inline
#define RPO_HDM inline
#endif

#define RC(x) ((real)(x))

// Hides an integer from the optimiser (no instruction emitted).  Used on the phase flag of evopf_project_loop so that
// the compiler cannot split the single loop back into two specialised loops (jump threading), which would duplicate
// the shared evaluation + LU code the loop exists to share.
#if defined(__GNUC__) || defined(__CUDACC__)
#define RPO_OPAQUE(x) asm volatile("" : "+r"(x))
#else
#define RPO_OPAQUE(x) do { } while (0)
#endif

// RPO_ALIGN(n): markers the generator drops every 32 statements of the big straight-line functions.  On the device,
// when RPO_ALIGN_EVERY > 0, every RPO_ALIGN_EVERY-th marker is a block barrier that keeps the warps of a CTA within
// one instruction-cache window of each other (the kernels are instruction-fetch bound).
#ifndef RPO_ALIGN_EVERY
#define RPO_ALIGN_EVERY 0
#endif
#if defined(__CUDA_ARCH__) && RPO_ALIGN_EVERY > 0
#define RPO_ALIGN(n) do { if ((n) % RPO_ALIGN_EVERY == 0) __syncthreads(); } while (0)
#else
#define RPO_ALIGN(n) do { } while (0)
#endif

// ---- branch-free single-precision sincos -------------------------------------------------------------
// Cody-Waite 3-term reduction by pi/2 (exact for |x| up to ~1e5 rad; voltage angles live in [-pi, pi]) and the
// Cephes minimax polynomials on [-pi/4, pi/4] (max error ~1 ulp).  Same code on host (hostsim) and device; it
// replaces CUDA's sincosf, whose inlined Payne-Hanek slow path made up ~1/3 of the kernel's instructions.
RPO_HD void rpo_sincosf(float x, float* sp, float* cp) {
    const float kf = rintf(x * 0.63661977236758134f);
    float r = fmaf(kf, -0x1.921fb6p+0f, x);          // pi/2 = C1 + C2 + C3, C1 = float(pi/2); the FMA keeps k*C exact
    r = fmaf(kf, 0x1.777a5cp-25f, r);
    r = fmaf(kf, 0x1.ee59dap-50f, r);
    const float r2 = r * r;
    float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, r2, -1.6666654611e-1f);
    const float s = fmaf(r * r2, ps, r);
    float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, r2, 4.166664568298827e-2f);
    pc = fmaf(pc, r2, -0.5f);
    const float c = fmaf(pc, r2, 1.0f);
    const int q = (int)kf;
    const float ss = (q & 1) ? c : s;
    const float cc = (q & 1) ? s : c;
    *sp = (q & 2) ? -ss : ss;
    *cp = ((q + 1) & 2) ? -cc : cc;
}

// sincos without range reduction for |x| <= 1.6 (pi/2 + margin): near-minimax fits of (sin x - x)/x^3 and
// (cos x - 1)/x^2 in x^2 (max abs error 1.5e-7 over the interval, evaluated in float32); 11 FP32 instructions.
RPO_HD void rpo_sincosf_small(float x, float* sp, float* cp) {
    const float x2 = x * x;
    float ps = fmaf(x2, 0x1.60ff98p-19f, -0x1.9fae94p-13f);
    ps = fmaf(ps, x2, 0x1.111036p-7f);
    ps = fmaf(ps, x2, -0x1.555554p-3f);
    *sp = fmaf(x * x2, ps, x);
    float pc = fmaf(x2, -0x1.19d6b8p-22f, 0x1.9f9cb4p-16f);
    pc = fmaf(pc, x2, -0x1.6c14f8p-10f);
    pc = fmaf(pc, x2, 0x1.55555p-5f);
    pc = fmaf(pc, x2, -0.5f);
    *cp = fmaf(pc, x2, 1.0f);
}

// ---- sincos for the mixed-precision refinement of the completion ----------------------------------------
// sin x = x + x^3 ps(x^2), cos x = 1 + x^2 pc(x^2): the leading terms are exact, so evaluating the small corrections
// ps / pc in float32 (the polynomials of rpo_sincosf_small) and only the final combination in float64 gives an absolute
// error of 6e-9 for |x| <= 0.6 and 2e-8 for |x| <= 1 (pure float32: 4e-8 / 5e-8 plus the rounding of the result) at the
// price of 4 FP64 instructions.  Larger angles (never met in power-flow states) fall back to the float32 values.
RPO_HD void rpo_sincos_mixed(float x, double* sp, double* cp) {
    if (fabsf(x) <= 1.6f) {
        const float x2 = x * x;
        float ps = fmaf(x2, 0x1.60ff98p-19f, -0x1.9fae94p-13f);
        ps = fmaf(ps, x2, 0x1.111036p-7f);
        ps = fmaf(ps, x2, -0x1.555554p-3f);
        float pc = fmaf(x2, -0x1.19d6b8p-22f, 0x1.9f9cb4p-16f);
        pc = fmaf(pc, x2, -0x1.6c14f8p-10f);
        pc = fmaf(pc, x2, 0x1.55555p-5f);
        pc = fmaf(pc, x2, -0.5f);
        const double xd = (double)x, x2d = xd * xd;
        *sp = fma(xd * x2d, (double)ps, xd);
        *cp = fma(x2d, (double)pc, 1.0);
    } else {
        float s_, c_;
        rpo_sincosf(x, &s_, &c_);
        *sp = (double)s_;
        *cp = (double)c_;
    }
}

// next representable `real` above / below x (the reduced cart correction walks to a rounding boundary with it)
RPO_HD float rpo_next(float x, bool up) { return nextafterf(x, up ? 3.0e38f : -3.0e38f); }
#if !defined(__CUDA_ARCH__)
static inline double rpo_next(double x, bool up) { return nextafter(x, up ? 1.0e300 : -1.0e300); }
#endif

#if defined(__CUDA_ARCH__)
// device, float only (the CUDA build is always float)
RPO_D float rpo_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define SINCOS(x, s, c) rpo_sincosf((x), (s), (c))
#define SINCOS_SMALL(x, s, c) rpo_sincosf_small((x), (s), (c))
#define SINCOS_REFINE(x, s, c) rpo_sincos_mixed((x), (s), (c))
#define RCP(x) rpo_rcp_approx(x)
#define FMIN(a, b) fminf((a), (b))
#define FMAX(a, b) fmaxf((a), (b))
#define FABS(a) fabsf(a)
#else
static inline float rpo_fma_(float a, float b, float c) { return fmaf(a, b, c); }
static inline double rpo_fma_(double a, double b, double c) { return fma(a, b, c); }
static inline void rpo_sincos_(float x, float* s, float* c) { rpo_sincosf(x, s, c); }
static inline void rpo_sincos_(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
static inline void rpo_sincos_small_(float x, float* s, float* c) { rpo_sincosf_small(x, s, c); }
static inline void rpo_sincos_small_(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
static inline void rpo_sincos_refine_(float x, double* s, double* c) { rpo_sincos_mixed(x, s, c); }
static inline void rpo_sincos_refine_(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }   // float64 validation build
#define SINCOS_REFINE(x, s, c) rpo_sincos_refine_((x), (s), (c))
#define FMA(a, b, c) rpo_fma_((real)(a), (real)(b), (real)(c))
#define SINCOS(x, s, c) rpo_sincos_((x), (s), (c))
#define SINCOS_SMALL(x, s, c) rpo_sincos_small_((x), (s), (c))
#define RCP(x) (RC(1.0) / (x))
#define FMIN(a, b) ((a) < (b) ? (a) : (b))
#define FMAX(a, b) ((a) > (b) ? (a) : (b))
#define FABS(a) ((a) < 0 ? -(a) : (a))
#endif
