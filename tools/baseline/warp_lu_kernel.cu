// MEASUREMENT BASELINE, not part of the product library.
//
// The mapping BASELINE.json's north-star sketches, implemented literally for the EVOPF completion
// (PFFunctionFn.forward, rpo/env/electrical_grid/evopf.py:789-855): ONE WARP PER CONSTRAINT INSTANCE, the dense
// 22x22 Newton Jacobian staged in SHARED MEMORY, LU with partial pivoting (warp arg-max by shuffles, pivot row
// broadcast from shared memory, one lane per row), back-substitution with warp-shuffle reductions.
// tools/baseline_compare_gpu.py times it against rpo_evopf_complete_fwd on identical inputs and checks that both
// give the same y; the result is recorded in profiles/ and explains the thread-per-instance design.
#include <cuda_runtime.h>
#include <stdint.h>

#define NB 14
#define NN 22
#define LDA 24            // 22 columns + rhs + pad
#define WARPS_PER_CTA 4

__constant__ float cG[NB * NB], cB[NB * NB];
__constant__ int c_eq_kind[NN], c_eq_bus[NN], c_unk_kind[NN], c_unk_bus[NN];   // kind 0 = P / theta, 1 = Q / V
__constant__ int c_spv[5], c_pv_gen[4], c_pv_bus[4], c_pq[9];
__constant__ float c_vm_init[NB], c_va_init[NB];

struct WarpSmem {
    float A[NN * LDA];
    float vm[NB], va[NB], vr[NB], vi[NB], P[NB], Q[NB];
    float pinj[NB], qinj[NB];
    float x[NN];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
k_warp_complete(const float* __restrict__ state, int ld, const float* __restrict__ z, float* __restrict__ y_out, long long B,
                float tol, int max_iters) {
    __shared__ WarpSmem sm[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long inst = (long long)blockIdx.x * WARPS_PER_CTA + w;
    if (inst >= B) return;
    WarpSmem& s = sm[w];
    const float* x = state + inst * ld;
    const float* zz = z + inst * 14;
    // step 0: scatter z, flat start (evopf.py:794-807); injections
    float pg_l = 0.f, pe_l = 0.f;
    if (lane < NB) {
        s.vm[lane] = c_vm_init[lane];
        s.va[lane] = c_va_init[lane];
        s.pinj[lane] = -x[lane];
        s.qinj[lane] = -x[NB + lane];
    }
    __syncwarp();
    if (lane < 5) {
        s.vm[c_spv[lane]] = zz[4 + lane];
        pe_l = zz[9 + lane];
        pg_l = lane == 0 ? 0.f : zz[lane - 1];
        s.pinj[c_spv[lane]] += pg_l + pe_l;
    }
    __syncwarp();
    for (int it = 0; it < max_iters; ++it) {
        if (lane < NB) {
            float sn, cs;
            sincosf(s.va[lane], &sn, &cs);
            s.vr[lane] = s.vm[lane] * cs;
            s.vi[lane] = s.vm[lane] * sn;
        }
        __syncwarp();
        if (lane < NB) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const float g = cG[lane * NB + j], bb = cB[lane * NB + j];
                a = fmaf(g, s.vr[j], a); a = fmaf(-bb, s.vi[j], a);
                b = fmaf(bb, s.vr[j], b); b = fmaf(g, s.vi[j], b);
            }
            s.P[lane] = s.vr[lane] * a + s.vi[lane] * b;
            s.Q[lane] = s.vi[lane] * a - s.vr[lane] * b;
        }
        __syncwarp();
        // one lane per equation row: dense Jacobian row (V columns scaled by V_j) + right-hand side
        if (lane < NN) {
            const int ek = c_eq_kind[lane], i = c_eq_bus[lane];
            const float vri = s.vr[i], vii = s.vi[i];
            for (int c = 0; c < NN; ++c) {
                const int uk = c_unk_kind[c], j = c_unk_bus[c];
                const float g = cG[i * NB + j], bb = cB[i * NB + j];
                const float al = g * s.vr[j] - bb * s.vi[j], be = bb * s.vr[j] + g * s.vi[j];
                const float t1 = vri * al + vii * be, t2 = vii * al - vri * be;
                float v;
                if (ek == 0) v = (uk == 0) ? t2 : t1; else v = (uk == 0) ? -t1 : t2;
                if (i == j) {
                    if (ek == 0) v += (uk == 0) ? -s.Q[i] : s.P[i]; else v += (uk == 0) ? s.P[i] : s.Q[i];
                }
                s.A[lane * LDA + c] = v;
            }
            s.A[lane * LDA + NN] = (ek == 0) ? (s.pinj[i] - s.P[i]) : (s.qinj[i] - s.Q[i]);
        }
        __syncwarp();
        // LU with partial pivoting on the augmented matrix
        for (int k = 0; k < NN; ++k) {
            float best = (lane >= k && lane < NN) ? fabsf(s.A[lane * LDA + k]) : -1.f;
            int arg = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (arg != k && lane <= NN) {                       // swap rows k and arg, one column per lane
                const float t = s.A[k * LDA + lane];
                s.A[k * LDA + lane] = s.A[arg * LDA + lane];
                s.A[arg * LDA + lane] = t;
            }
            __syncwarp();
            const float inv = 1.0f / s.A[k * LDA + k];
            if (lane > k && lane < NN) {
                const float m = s.A[lane * LDA + k] * inv;
                for (int j = k + 1; j <= NN; ++j) s.A[lane * LDA + j] = fmaf(-m, s.A[k * LDA + j], s.A[lane * LDA + j]);
            }
            __syncwarp();
        }
        // back-substitution: lane j holds x_j; the dot product of row k with x is a warp-shuffle reduction
        float xj = 0.f;
        for (int k = NN - 1; k >= 0; --k) {
            const float part = (lane > k && lane < NN) ? s.A[k * LDA + lane] * xj : 0.f;
            const float dot = warp_sum(part);
            if (lane == k) xj = (s.A[k * LDA + NN] - dot) / s.A[k * LDA + k];
        }
        // update (scaled V columns) and ||delta||_2 by shuffle reduction
        float d = 0.f;
        if (lane < NN) {
            const int j = c_unk_bus[lane];
            if (c_unk_kind[lane] == 0) { d = xj; s.va[j] += xj; }
            else { d = s.vm[j] * xj; s.vm[j] += d; }
        }
        const float n2 = warp_sum(d * d);
        __syncwarp();
        if (sqrtf(n2) < tol) break;
    }
    // step 2 (evopf.py:844-848): flows at the final point
    if (lane < NB) {
        float sn, cs;
        sincosf(s.va[lane], &sn, &cs);
        s.vr[lane] = s.vm[lane] * cs;
        s.vi[lane] = s.vm[lane] * sn;
    }
    __syncwarp();
    if (lane < NB) {
        float a = 0.f, b = 0.f;
        for (int j = 0; j < NB; ++j) {
            const float g = cG[lane * NB + j], bb = cB[lane * NB + j];
            a = fmaf(g, s.vr[j], a); a = fmaf(-bb, s.vi[j], a);
            b = fmaf(bb, s.vr[j], b); b = fmaf(g, s.vi[j], b);
        }
        s.P[lane] = s.vr[lane] * a + s.vi[lane] * b;
        s.Q[lane] = s.vi[lane] * a - s.vr[lane] * b;
    }
    __syncwarp();
    float* y = y_out + inst * 43;
    if (lane < 5) {
        const int bus = c_spv[lane];
        y[5 + lane] = x[NB + bus] + s.Q[bus];
        y[lane] = lane == 0 ? (x[bus] - pe_l) + s.P[bus] : pg_l;
        y[38 + lane] = pe_l;
    }
    if (lane < NB) {
        y[10 + lane] = s.vm[lane];
        y[24 + lane] = s.va[lane];
    }
}

extern "C" {
int warp_baseline_init(const float* G, const float* Bm, const int* eq_kind, const int* eq_bus, const int* unk_kind,
                       const int* unk_bus, const int* spv, const float* vm_init, const float* va_init) {
    cudaMemcpyToSymbol(cG, G, sizeof(float) * NB * NB);
    cudaMemcpyToSymbol(cB, Bm, sizeof(float) * NB * NB);
    cudaMemcpyToSymbol(c_eq_kind, eq_kind, sizeof(int) * NN);
    cudaMemcpyToSymbol(c_eq_bus, eq_bus, sizeof(int) * NN);
    cudaMemcpyToSymbol(c_unk_kind, unk_kind, sizeof(int) * NN);
    cudaMemcpyToSymbol(c_unk_bus, unk_bus, sizeof(int) * NN);
    cudaMemcpyToSymbol(c_spv, spv, sizeof(int) * 5);
    cudaMemcpyToSymbol(c_vm_init, vm_init, sizeof(float) * NB);
    cudaMemcpyToSymbol(c_va_init, va_init, sizeof(float) * NB);
    return (int)cudaGetLastError();
}

int warp_baseline_complete(const float* state, int ld, const float* z, float* y, long long B, float tol, int max_iters, void* stream) {
    const unsigned grid = (unsigned)((B + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    k_warp_complete<<<grid, 32 * WARPS_PER_CTA, 0, (cudaStream_t)stream>>>(state, ld, z, y, B, tol, max_iters);
    return (int)cudaGetLastError();
}
}
