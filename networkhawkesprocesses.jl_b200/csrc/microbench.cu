// microbench.cu -- in-library roofline denominators that MEASURED_PEAKS.json does not carry:
// the FP64 FMA peak and the register-resident pairs/s ceiling of each impulse function
// (SURVEY.md section 8d asks the builder to measure both and quote % roofline against them).
#include "nhp_internal.cuh"

// 8 independent DFMA chains per thread, no memory traffic
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[0] = s;
}

// FP64 tensor-core rate: mma.sync.aligned.m8n8k4.f64 (DMMA), 4 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) k_dmma_peak(double *out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-3, b = 0.5 - threadIdx.x * 1e-3;
    double c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    for (int i = 0; i < iters; i++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
    }
    double s = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
    if (s == 12345.678) out[0] = s;
}

// the device impulse function on register inputs (dt marches through the support)
template <typename E> __global__ void __launch_bounds__(256) k_pair_peak(double *out, int iters, E e, double D) {
    __shared__ FastTables s_ft;
    fast_tables_load(&s_ft);
    __syncthreads();
    const FastTables *ft = &s_ft;
    double dt0 = (threadIdx.x + 1) * (D / 300.0), acc = 0.0, step = D * 1e-7;
    double dt1 = dt0 * 0.5;
    for (int i = 0; i < iters; i++) {
        acc += pair_value(e, dt0, D, ft);
        acc += pair_value(e, dt1, D, ft);
        dt0 += step; dt1 += step;
    }
    if (acc == 12345.678) out[0] = acc;
}

// which: 0 = DFMA TFLOP/s (2 flops per FMA), 1 = LogitNormal pairs/s, 2 = Exponential pairs/s, 3 = DMMA m8n8k4 TFLOP/s
extern "C" int nhp_bench_fp64(nhp_ctx *ctx, int which, double *result) {
    NHP_CHECK(ctx, ctx != nullptr, NHP_ERR_INVALID, "ctx is NULL");
    NHP_CHECK(ctx, result != nullptr && which >= 0 && which <= 3, NHP_ERR_INVALID, "nhp_bench_fp64: bad argument");
    NHP_CUDA(ctx, cudaSetDevice(ctx->device));
    void *scratch;
    NHP_TRY(nhp_scratch(ctx, 64, &scratch));
    NHP_CUDA(ctx, fast_tables_upload(ctx->stream));
    const int blocks = ctx->sm_count * 8, iters = which == 0 ? 20000 : 2000;
    double best = 1e30;
    for (int rep = 0; rep < 4; rep++) {
        NHP_TRY(nhp_timer_begin(ctx));
        if (which == 0) k_dfma_peak<<<blocks, 256, 0, ctx->stream>>>((double *)scratch, iters, 0.999999, 1e-9);
        else if (which == 3) k_dmma_peak<<<blocks, 256, 0, ctx->stream>>>((double *)scratch, iters);
        else if (which == 1) { EntryLN e{0.3, 0.1, 0.6, 0.0}; k_pair_peak<EntryLN><<<blocks, 256, 0, ctx->stream>>>((double *)scratch, iters, e, 1.0); }
        else { EntryEX e{0.3, 1.1}; k_pair_peak<EntryEX><<<blocks, 256, 0, ctx->stream>>>((double *)scratch, iters, e, 1.0); }
        NHP_LAUNCHED(ctx);
        NHP_TRY(nhp_timer_end(ctx));
        if (rep > 0 && ctx->last_ms < best) best = ctx->last_ms;
    }
    double work = (double)blocks * 256.0 * iters * (which == 0 ? 16.0 : 2.0);
    if (which == 3) work = (double)blocks * 8.0 * iters * 4.0 * 512.0;  // 8 warps x 4 DMMA x (8*8*4*2 flops)
    *result = work / (best * 1e-3) / ((which == 0 || which == 3) ? 1e12 : 1.0);
    return NHP_OK;
}

// test hook: out[i] = fast_log(x[i]) (which = 0) or fast_exp(x[i]) (which = 1); HOST pointers
__global__ void k_fastmath_eval(const double *x, double *out, int64_t n, int which) {
    __shared__ FastTables s_ft;
    fast_tables_load(&s_ft);
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = which == 0 ? fast_log(x[i], &s_ft) : fast_exp(x[i], &s_ft);
}
extern "C" int nhp_test_fastmath(nhp_ctx *ctx, int which, const double *x, int64_t n, double *out) {
    NHP_CHECK(ctx, ctx != nullptr, NHP_ERR_INVALID, "ctx is NULL");
    NHP_CHECK(ctx, x && out && n > 0 && (which == 0 || which == 1), NHP_ERR_INVALID, "nhp_test_fastmath: bad argument");
    NHP_CUDA(ctx, cudaSetDevice(ctx->device));
    NHP_CUDA(ctx, fast_tables_upload(ctx->stream));
    void *scratch;
    NHP_TRY(nhp_scratch(ctx, (size_t)n * 16, &scratch));
    double *dx = (double *)scratch, *dout = dx + n;
    NHP_CUDA(ctx, cudaMemcpyAsync(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_fastmath_eval<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dx, dout, n, which);
    NHP_LAUNCHED(ctx);
    NHP_CUDA(ctx, cudaMemcpyAsync(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    NHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NHP_OK;
}
