// cont_exp_scan.cu -- recursive_loglikelihood (continuous.jl:241-276 / 407-442) for the Exponential impulse as a
// chunked scan.  The reference carries the K x K state R[p,c](t) = sum_{j: c_j = p, t_j < t} exp(-theta[p,c] (t - t_j))
// from event to event (O(N K) exponentials, strictly sequential).  The window sweeps reproduce it with a cut-off horizon
// H (nhp_cont_horizon_value): exact to 1e-14, but H * rate pairs per event -- thousands when the stream is dense and K is
// small.  Here the stream is cut into chunks of M events:
//   k_exp_chunk_aggregate   L_b[p,c]  = sum_{j in chunk b, c_j = p} exp(-theta[p,c] (T_{b+1} - t_j))      (M K exps per chunk, parallel over chunks)
//   k_exp_chunk_scan        S_{b+1}   = S_b exp(-theta (T_{b+1} - T_b)) + L_b                               (K^2 threads, sequential over the chunks)
//   k_exp_chunk_loglik      lambda_i  = lambda0 + sum_p wt[p,c] S_b[p,c] exp(-theta[p,c] (t_i - T_b))       (K exps: everything before the chunk)
//                                     + sum_{j in chunk b, j < i} wt[c_j,c] exp(-theta[c_j,c] (t_i - t_j))   (the chunk's own pairs)
// with T_b the time of the first event of chunk b: about M/2 + 2K exponentials per event and no truncation at all.
// Quirks kept: events at t == 0.0 never act as parents (Q6: index < n_t0), a finite dtmax is ignored (Q7), the
// compensator ignores A (Q3, through a.rowsum).  Used for unsharded data when it is cheaper than the horizon window.
#include "cont_sweep.cuh"
#include <algorithm>

struct ExpScanArgs {
    SweepArgs s;
    int M;              // events per chunk
    int64_t nchunks;
    double *S;          // [nchunks][K][K]: carry at the start of every chunk, child-major ([c][p]) like the table
};

// L_b into slot b + 1 of S (the scan turns it into S_{b+1}); slot 0 stays zero
__global__ void __launch_bounds__(256) k_exp_chunk_aggregate(const ExpScanArgs x) {
    extern __shared__ double s_L[];  // [K][K] child-major
    const SweepArgs &a = x.s;
    const int K = a.K;
    const int64_t b = blockIdx.x;   // chunks 0 .. nchunks - 2
    const int64_t i0 = a.first + b * x.M, i1 = i0 + x.M;  // a full chunk (the last, possibly partial, chunk feeds nobody)
    const EntryEX *table = reinterpret_cast<const EntryEX *>(a.table);
    for (int k = threadIdx.x; k < K * K; k += blockDim.x) s_L[k] = 0.0;
    __syncthreads();
    const double Tn = a.t[i1];      // first event of the next chunk
    for (int64_t j = max(i0, a.jmin); j < i1; j++) {
        const int p = a.c[j];
        const double d = Tn - a.t[j];
        for (int c = threadIdx.x; c < K; c += blockDim.x) s_L[c * K + p] += exp(-table[(size_t)c * K + p].theta * d);  // thread c owns column c: no races
    }
    __syncthreads();
    double *out = x.S + (size_t)(b + 1) * K * K;
    for (int k = threadIdx.x; k < K * K; k += blockDim.x) out[k] = s_L[k];
}

__global__ void k_exp_chunk_scan(const ExpScanArgs x) {
    const SweepArgs &a = x.s;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t KK = (int64_t)a.K * a.K;
    if (k >= KK) return;
    const double th = reinterpret_cast<const EntryEX *>(a.table)[k].theta;
    double S = 0.0;
    x.S[k] = 0.0;
    double Tb = a.t[a.first];
    for (int64_t b = 0; b + 1 < x.nchunks; b++) {
        const double Tn = a.t[a.first + (b + 1) * x.M];
        S = S * exp(-th * (Tn - Tb)) + x.S[(size_t)(b + 1) * KK + k];
        x.S[(size_t)(b + 1) * KK + k] = S;
        Tb = Tn;
    }
}

__global__ void __launch_bounds__(256) k_exp_chunk_loglik(const ExpScanArgs x) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[16];
    const SweepArgs &a = x.s;
    const int K = a.K;
    double *s_t = reinterpret_cast<double *>(smem_raw);   // [M]
    int *s_c = reinterpret_cast<int *>(s_t + x.M);        // [M]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const EntryEX *table = reinterpret_cast<const EntryEX *>(a.table);
    double sum_log = 0.0, sum_row = 0.0;
    for (int64_t b = blockIdx.x; b < x.nchunks; b += gridDim.x) {
        const int64_t i0 = a.first + b * x.M;
        const int m = (int)min((int64_t)x.M, a.n - i0);
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += blockDim.x) { s_t[k] = a.t[i0 + k]; s_c[k] = a.c[i0 + k]; }
        __syncthreads();
        const double Tb = s_t[0];
        const double *Sb = x.S + (size_t)b * K * K;
        const int jl = (int)max((int64_t)0, a.jmin - i0);  // first admissible parent inside the chunk (quirk Q6)
        for (int e = warp; e < m; e += 8) {
            const int c = s_c[e];
            const double ti = s_t[e];
            const EntryEX *col = table + (size_t)c * K;
            double acc = 0.0;
            if (b > 0) {  // everything before the chunk, through the carried state
                const double dT = ti - Tb;
                const double *Sc = Sb + (size_t)c * K;
                for (int p = lane; p < K; p += 32) {
                    const EntryEX en = load_entry(col + p);
                    if (en.wt != 0.0) acc += en.wt * Sc[p] * exp(-en.theta * dT);
                }
            }
            for (int j = jl + lane; j < e; j += 32) {  // the chunk's own pairs
                const EntryEX en = load_entry(col + s_c[j]);
                if (en.wt != 0.0) acc += en.wt * exp(-en.theta * (ti - s_t[j]));
            }
            acc = warp_sum(acc);
            if (lane == 0) sum_log += log(acc + __ldg(a.lambda0 + c));
        }
    }
    block_sum2(sum_log, sum_row, red);
    if (threadIdx.x == 0) { a.partials[2 * (size_t)blockIdx.x] = sum_log; a.partials[2 * (size_t)blockIdx.x + 1] = 0.0; }
}

// Returns NHP_OK after launching (grid in *grid_out), 1 if the chunked scan does not apply or would not pay, < 0 on error.
int nhp_cont_try_exp_scan(nhp_ctx *ctx, nhp_events *ev, SweepArgs &a, int *grid_out) {
    const char *env = getenv("NHP_EXP_SCAN");
    if (env && atoi(env) == 0) return 1;
    const bool force = env && atoi(env) == 1;
    if (ctx->kind != NHP_EXPONENTIAL || ev->n_halo != 0 || ev->index_base != 0) return 1;  // shards keep the horizon halo
    if (a.lam0ev) return 1;  // grid baseline: the window sweeps carry the per-event rate
    const int64_t K = ctx->K, n = ev->n;
    if (n < 2) return 1;
    const size_t smem_agg = (size_t)K * K * sizeof(double);
    if (smem_agg > (size_t)ctx->smem_optin - 4096) return 1;
    // chunk length: at least 256 events, and few enough chunks that the carried states stay below ~1 GB
    int64_t M = 256;
    while ((double)((n + M - 1) / M) * (double)(K * K) * 8.0 > 1.0e9 && M < 65536) M <<= 1;
    if ((double)((n + M - 1) / M) * (double)(K * K) * 8.0 > 1.0e9) return 1;
    // exponentials per event: M/2 (own chunk) + 2K (carry in, aggregate out) against the horizon window
    const double cost = 0.5 * (double)M + 2.0 * (double)K;
    if (!force && !(ev->mean_win > 2.0 * cost)) return 1;
    const size_t smem_ll = (size_t)M * 12 + 16;
    if (smem_ll > (size_t)ctx->smem_optin - 4096) return 1;
    ExpScanArgs x;
    x.s = a; x.M = (int)M; x.nchunks = (n + M - 1) / M;
    void *scratch;
    NHP_TRY(nhp_scratch(ctx, (size_t)x.nchunks * K * K * sizeof(double), &scratch));
    x.S = (double *)scratch;
    cudaStream_t s = ctx->stream;
    if (x.nchunks > 1) {
        NHP_CUDA(ctx, cudaFuncSetAttribute(k_exp_chunk_aggregate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_agg));
        k_exp_chunk_aggregate<<<(unsigned)(x.nchunks - 1), 256, smem_agg, s>>>(x);
        NHP_LAUNCHED(ctx);
    }
    k_exp_chunk_scan<<<(unsigned)((K * K + 127) / 128), 128, 0, s>>>(x);
    NHP_LAUNCHED(ctx);
    NHP_CUDA(ctx, cudaFuncSetAttribute(k_exp_chunk_loglik, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ll));
    const int grid = (int)std::min<int64_t>(x.nchunks, (int64_t)ctx->sm_count * 8);
    k_exp_chunk_loglik<<<grid, 256, smem_ll, s>>>(x);
    NHP_LAUNCHED(ctx);
    NHP_CUDA(ctx, cudaGetLastError());
    *grid_out = grid;
    return NHP_OK;
}
