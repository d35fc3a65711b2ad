// Action-box squash / clip of one element, shared by the CUDA kernels and the host simulator.
//
// Reference: rpo/algo/model/utils.py BoxConstraint — `__call__` :40-51 (scale * x + base), `clip` :92-104,
// `action_scale` :106-118, `update_box` :75-90 (scale = (cmax - cmin) / 2, base = cmin + scale for style "tanh";
// scale = cmax - cmin, base = cmin for style "sigmoid"), called from the policies as box(tanh(x), s)
// (rpo/algo/model/policy.py:30-31,61-65).  For EVOPF the last five columns of the box are volatile: they follow the
// battery state of charge (EVOPFEnv.update evopf.py:769-783, Battery.update_bound :127-143) — computed in float64 and
// rounded when numpy stores them into the float32 box rows; everything after that is float32, multiply THEN add.
#pragma once

#include "rpo_math.cuh"

enum {
    RPO_BOX_AFFINE = 0,     // scale * x + base, style "tanh" (x is already in [-1, 1])
    RPO_BOX_TANH = 1,       // scale * tanh(x) + base (the policy's tanh fused in)
    RPO_BOX_CLIP = 2,       // min(max(x, lo), hi)
    RPO_BOX_AFFINE01 = 3,   // (hi - lo) * x + lo, style "sigmoid" (x in [0, 1])
    RPO_BOX_SCALE = 4,      // (hi - lo) / 2, x ignored          (action_scale, style "tanh")
    RPO_BOX_LOW = 5,        // lo, x ignored                     (update_box)
    RPO_BOX_HIGH = 6,       // hi, x ignored
    RPO_BOX_N_MODES = 7
};

#if defined(__CUDA_ARCH__)
#define RPO_TANH(x) tanhf(x)
#else
static inline float rpo_tanh_(float x) { return tanhf(x); }
static inline double rpo_tanh_(double x) { return tanh(x); }
#define RPO_TANH(x) rpo_tanh_(x)
#endif

RPO_HD real box_fwd(real x, real lo, real hi, int mode) {
    switch (mode) {
    case RPO_BOX_CLIP: return FMIN(FMAX(x, lo), hi);
    case RPO_BOX_AFFINE01: return (hi - lo) * x + lo;
    case RPO_BOX_LOW: return lo;
    case RPO_BOX_HIGH: return hi;
    default: break;
    }
    const real scale = (hi - lo) * RC(0.5);
    if (mode == RPO_BOX_SCALE) return scale;
    const real base = lo + scale;
    const real t = mode == RPO_BOX_TANH ? RPO_TANH(x) : x;
    return scale * t + base;
}

// d out / d x . g  (no gradient reaches the state: the reference's bounds go through numpy)
RPO_HD real box_bwd(real x, real lo, real hi, real g, int mode) {
    switch (mode) {
    case RPO_BOX_CLIP: return (x >= lo && x <= hi) ? g : RC(0.0);          // torch.clip backward
    case RPO_BOX_AFFINE01: return g * (hi - lo);
    case RPO_BOX_AFFINE: return g * ((hi - lo) * RC(0.5));
    case RPO_BOX_TANH: {
        const real t = RPO_TANH(x);
        return (g * ((hi - lo) * RC(0.5))) * (RC(1.0) - t * t);
    }
    default: return RC(0.0);
    }
}
