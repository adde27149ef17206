// cont_query.cu -- intensity(process, data, times::Vector{Float64})  (continuous.jl:76-96):
// lambda_k(t0) for all K nodes at arbitrary query times; strict window  t0 - dtmax < t_j < t0.
// One CTA per query time: the window bounds come from two binary searches on the sorted times,
// the per-entry logs (shared by all K children) are computed once into shared memory, then each
// thread owns child nodes and sums the window in event order, baseline added last as the
// reference does (`intensity(baseline, time) .+ lambda`).
#include "cont_sweep.cuh"

constexpr int QCHUNK = 256;

struct QueryArgs {
    const double *t; const int *c; int64_t n;
    const double *tq; int64_t nq;
    int K; const void *table; const double *lambda0;
    const double *gx, *gv; int G;  // grid baseline (G == 0: homogeneous lambda0)
    double D, horizon;
    double *out;  // [nq*K], out[q + nq*k]
};

__device__ __forceinline__ int64_t lower_bound_ge(const double *t, int64_t n, double x) {  // first j with t[j] >= x
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (t[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int64_t upper_bound_gt(const double *t, int64_t n, double x) {  // first j with t[j] > x
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (t[mid] <= x) lo = mid + 1; else hi = mid; }
    return lo;
}

template <int KIND> __global__ void __launch_bounds__(256) k_intensity_query(const QueryArgs a) {
    typedef typename EntryOf<KIND>::type E;
    __shared__ FastTables s_ft;
    __shared__ double s_x[QCHUNK], s_y[QCHUNK];  // LogitNormal: log dt, log(D - dt); Exponential: dt
    __shared__ int s_c[QCHUNK];
    __shared__ int64_t s_rng[2];
    fast_tables_load(&s_ft);
    const int64_t q = blockIdx.x;
    const double t0 = a.tq[q];
    if (threadIdx.x == 0) {
        s_rng[0] = upper_bound_gt(a.t, a.n, t0 - a.horizon);
        s_rng[1] = lower_bound_ge(a.t, a.n, t0);
    }
    __syncthreads();
    const int64_t lo = s_rng[0], hi = s_rng[1];
    constexpr int MAXC = 8;  // children per thread per pass (K <= 2048 in one pass)
    for (int cbase = 0; cbase < a.K; cbase += 256 * MAXC) {
        double acc[MAXC];
#pragma unroll
        for (int m = 0; m < MAXC; m++) acc[m] = 0.0;
        for (int64_t j0 = lo; j0 < hi; j0 += QCHUNK) {
            int cnt = (int)min((int64_t)QCHUNK, hi - j0);
            __syncthreads();
            if ((int)threadIdx.x < cnt) {
                double dt = t0 - a.t[j0 + threadIdx.x];
                s_c[threadIdx.x] = a.c[j0 + threadIdx.x];
                if (KIND == NHP_LOGITNORMAL) {
                    bool ok = dt > 0.0 && dt < a.D;
                    s_x[threadIdx.x] = ok ? fast_log(dt, &s_ft) : NAN;  // NaN marks pdf == 0 (outside 0 < x < 1)
                    s_y[threadIdx.x] = ok ? fast_log(a.D - dt, &s_ft) : 0.0;
                } else s_x[threadIdx.x] = dt;
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < MAXC; m++) {
                int ch = cbase + m * 256 + threadIdx.x;
                if (ch >= a.K) break;
                const E *col = reinterpret_cast<const E *>(a.table) + (size_t)ch * a.K;
                double s = acc[m];
                for (int e = 0; e < cnt; e++) {
                    if (KIND == NHP_LOGITNORMAL) {
                        double la = s_x[e];
                        if (la == la) {
                            EntryLN en = load_entry(reinterpret_cast<const EntryLN *>(col) + s_c[e]);
                            double lb = s_y[e], dz = (la - lb) - en.mu, hd = en.h * dz;
                            s += en.cf * fast_exp(fma(-hd, dz, -(la + lb)), &s_ft);
                        }
                    } else {
                        EntryEX en = load_entry(reinterpret_cast<const EntryEX *>(col) + s_c[e]);
                        s += pair_value(en, s_x[e], a.D, &s_ft);
                    }
                }
                acc[m] = s;
            }
        }
#pragma unroll
        for (int m = 0; m < MAXC; m++) {
            int ch = cbase + m * 256 + threadIdx.x;
            if (ch < a.K) {
                double b0 = a.lambda0[ch];
                if (a.G > 0) {  // lambda0_ch(t0) on the grid (interpolation.jl:27-36; the host has checked the support)
                    const double *y = a.gv + (int64_t)ch * a.G;
                    int lo = 0, hi = a.G - 1;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.gx[mid] <= t0) lo = mid; else hi = mid; }
                    b0 = t0 >= a.gx[a.G - 1] ? y[a.G - 1] : (y[lo + 1] * (t0 - a.gx[lo]) + y[lo] * (a.gx[lo + 1] - t0)) / (a.gx[lo + 1] - a.gx[lo]);
                }
                a.out[q + a.nq * (int64_t)ch] = b0 + acc[m];
            }
        }
    }
}

extern "C" int nhp_cont_intensity(nhp_ctx *ctx, nhp_events *ev, const double *times, int64_t nq, double *out) {
    NHP_CHECK(ctx, ctx != nullptr, NHP_ERR_INVALID, "ctx is NULL");
    NHP_CHECK(ctx, ctx->cont_set, NHP_ERR_STATE, "continuous parameters not set (call nhp_cont_params_set)");
    NHP_CHECK(ctx, ev != nullptr && ev->K == ctx->K, NHP_ERR_INVALID, "nhp_cont_intensity: bad events handle");
    NHP_CHECK(ctx, nq >= 0 && (nq == 0 || (times && out)), NHP_ERR_INVALID, "nhp_cont_intensity: NULL times/out");
    for (int64_t q = 0; q < nq; q++)
        NHP_CHECK(ctx, times[q] >= 0.0, NHP_ERR_INVALID, "intensity: time must be non-negative (baselines.jl:111)");
    if (nq == 0) return NHP_OK;
    NHP_CUDA(ctx, cudaSetDevice(ctx->device));
    NHP_CUDA(ctx, fast_tables_upload(ctx->stream));
    void *scratch;
    size_t bytes_q = (size_t)nq * sizeof(double), bytes_o = (size_t)nq * ctx->K * sizeof(double);
    NHP_TRY(nhp_scratch(ctx, bytes_q + bytes_o, &scratch));
    double *dq = (double *)scratch, *dout = dq + nq;
    NHP_CUDA(ctx, cudaMemcpyAsync(dq, times, bytes_q, cudaMemcpyHostToDevice, ctx->stream));
    QueryArgs a;
    a.t = ev->d_t; a.c = ev->d_c; a.n = ev->n; a.tq = dq; a.nq = nq; a.K = (int)ctx->K; a.table = ctx->d_table; a.lambda0 = ctx->d_lambda0;
    a.D = ctx->dtmax; a.horizon = nhp_cont_horizon_value(ctx, ev->index_base + ev->n, 0); a.out = dout;
    a.gx = ctx->d_bgrid_x; a.gv = ctx->d_bgrid_v; a.G = (int)ctx->bgrid_n;
    if (ctx->bgrid_n > 0) {
        std::vector<double> gx((size_t)ctx->bgrid_n);
        NHP_CUDA(ctx, cudaMemcpy(gx.data(), ctx->d_bgrid_x, gx.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t q = 0; q < nq; q++)
            NHP_CHECK(ctx, times[q] >= gx.front() && times[q] <= gx.back(), NHP_ERR_INVALID, "intensity: time outside the interpolation support of the baseline grid (interpolation.jl:29)");
    }
    NHP_TRY(nhp_timer_begin(ctx));
    if (ctx->kind == NHP_LOGITNORMAL) k_intensity_query<NHP_LOGITNORMAL><<<(unsigned)nq, 256, 0, ctx->stream>>>(a);
    else k_intensity_query<NHP_EXPONENTIAL><<<(unsigned)nq, 256, 0, ctx->stream>>>(a);
    NHP_LAUNCHED(ctx);
    NHP_CUDA(ctx, cudaGetLastError());
    NHP_TRY(nhp_timer_end(ctx));
    NHP_CUDA(ctx, cudaMemcpyAsync(out, dout, bytes_o, cudaMemcpyDeviceToHost, ctx->stream));
    NHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NHP_OK;
}
