// fastmath.cuh -- table-driven FP64 log/exp for the impulse evaluation.
//
// The continuous sweeps are bound by the FP64 pipe: with libdevice log/exp/div a LogitNormal pair
// costs ~80 DP instructions.  The parity contract is 1e-10 relative on intensities, so an
// implementation with ~1e-15 absolute error in log and ~2e-16 relative error in exp is exact for the
// purpose while costing 11 + 10 DP instructions (tables in shared memory, integer work on the ALU
// pipe).  LogitNormal pair: 2 log + 1 exp + 8 = ~40 DP instructions; Exponential pair: ~13.
//   log(x)  = e ln2 + (-log c_i) + log1p(m c_i - 1),   c_i ~ 1/m on the i-th of 128 mantissa intervals
//   exp(x)  = 2^(k/64) (1 + expm1(r)),                  k = round(64 x / ln2), r = x - k ln2/64
// Tables are correctly rounded (tools/gen_fastmath_tables.py).
#pragma once
#include <cuda_runtime.h>
#include "fastmath_tables.h"

struct FastTables {
    double2 logtab[128];  // {c, -log c}
    double exptab[64];    // 2^(j/64)
};

// one copy per translation unit (2.5 KB), read once per CTA through L2
static __device__ unsigned long long g_nhp_logtab[256];
static __device__ unsigned long long g_nhp_exptab[64];
// host: copy the generated tables into this translation unit's device arrays (idempotent)
static inline cudaError_t fast_tables_upload(cudaStream_t s) {
    cudaError_t e = cudaMemcpyToSymbolAsync(g_nhp_logtab, NHP_LOGTAB_BITS, sizeof(NHP_LOGTAB_BITS), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(g_nhp_exptab, NHP_EXPTAB_BITS, sizeof(NHP_EXPTAB_BITS), 0, cudaMemcpyHostToDevice, s);
}

// cooperative load of the tables into shared memory (2.5 KB); caller synchronises afterwards
__device__ __forceinline__ void fast_tables_load(FastTables *ft) {
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(ft);
    for (int i = threadIdx.x; i < 256 + 64; i += blockDim.x) dst[i] = i < 256 ? g_nhp_logtab[i] : g_nhp_exptab[i - 256];
}

// rare-path fallbacks kept out of line so the hot loops stay small
static __device__ __noinline__ double slow_log(double x) { return log(x); }
static __device__ __noinline__ double slow_exp(double x) { return exp(x); }

// x must be a positive normal double (callers guarantee 0 < x < inf; subnormals take the libdevice path)
__device__ __forceinline__ double fast_log(double x, const FastTables *ft) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    if (hi < 0x00100000 || hi >= 0x7ff00000) return slow_log(x);
    int e = (hi >> 20) - 1023;
    int idx = (hi >> 13) & 127;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    double2 t = ft->logtab[idx];
    double r = fma(m, t.x, -1.0);
    double q = fma(r, -1.0 / 6.0, 0.2);
    q = fma(r, q, -0.25);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    double r2 = r * r;
    double p = fma(r2, q, r);
    double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;  // (double)e
    double a = fma(ed, NHP_LN2_HI, t.y);
    return a + fma(ed, NHP_LN2_LO, p);
}

// flushes to 0 below -707 (the omitted mass is < 1e-307) and defers to libdevice above 709 / for NaN
__device__ __forceinline__ double fast_exp(double x, const FastTables *ft) {
    if (!(x >= -707.0)) return x < -707.0 ? 0.0 : x;  // NaN propagates
    if (x > 709.0) return slow_exp(x);
    double t = fma(x, NHP_64_OVER_LN2, 6755399441055744.0);
    int k = __double2loint(t);
    double kf = t - 6755399441055744.0;
    double r = fma(kf, -NHP_LN2_64_HI, x);
    r = fma(kf, -NHP_LN2_64_LO, r);
    double T = ft->exptab[k & 63];
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    double r2 = r * r;
    double p = fma(r2, q, r);
    double v = fma(T, p, T);
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}
