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

#if defined(__CUDA_ARCH__)
// device, float only (the CUDA build is always float)
#define FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define SINCOS(x, s, c) sincosf((x), (s), (c))
#define RCP(x) __frcp_rn(x)
#define FMIN(a, b) fminf((a), (b))
#define FMAX(a, b) fmaxf((a), (b))
#define FABS(a) fabsf(a)
#else
static inline float rpo_fma_(float a, float b, float c) { return fmaf(a, b, c); }
static inline double rpo_fma_(double a, double b, double c) { return fma(a, b, c); }
static inline void rpo_sincos_(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
static inline void rpo_sincos_(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
#define FMA(a, b, c) rpo_fma_((real)(a), (real)(b), (real)(c))
#define SINCOS(x, s, c) rpo_sincos_((x), (s), (c))
#define RCP(x) (RC(1.0) / (x))
#define FMIN(a, b) ((a) < (b) ? (a) : (b))
#define FMAX(a, b) ((a) > (b) ? (a) : (b))
#define FABS(a) ((a) < 0 ? -(a) : (a))
#endif
