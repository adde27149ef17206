// cont_adjacency.cu -- resample_adjacency_matrix!(process, data)  (continuous.jl:444-519).
//
// The reference evaluates, for every (p, c), two full passes over all N events (2 K^2 N work).  The
// conditional it samples from only depends on the child events on c whose window holds an event of
// node p:  with G_i[p] = W[p,c] sum_{j in win(i), c_j = p} h_pc(t_i - t_j)  and
// lambda_i = lambda0_c + sum_p A[p,c] G_i[p],
//     ll1 - ll0 = -W[p,c] n_p + sum_{i on c, G_i[p] > 0} [ log(lambda_i^{-p} + G_i[p]) - log(lambda_i^{-p}) ]
//                 + log rho - log(1 - rho),          A[p,c] ~ Bernoulli(exp(ll1 - logsumexp(ll0, ll1))),
// p = 1..K sequentially inside a column, columns independent (the reference's Threads.@threads axis).
// One CTA owns a column: (A) it walks the windows of the column's child events once and buckets the
// (event, G) entries by parent node in a per-CTA scratch area (counting sort, shared-memory
// histogram), (B) then runs the K sequential Bernoulli steps, each a block-wide reduction over one
// bucket.  Same conditional distribution, O(N w) work per sweep instead of O(K^2 N).
//
// Cached variant (default when it fits in memory): the bucketed structure -- which (child event, window predecessor) pairs
// exist, grouped by column and parent node, with their lags t_i - t_j -- depends on the data and the look-back horizon only.
// It is built once per events handle (k_adj_build: 12 B per pair) and every sweep then streams it (k_adj_sweep): the
// impulse value of an entry is evaluated on the fly with the (p, c) parameters held in registers for the whole bucket,
// there is no per-sweep bucketing, no scattered entry writes and no table gather.  (event, parent) pairs that occur once
// (almost all of them) skip the duplicate-aggregation pass through a flag set at build time.
#include "cont_sweep.cuh"
int nhp_cont_params_refresh(nhp_ctx *ctx);  // cont_conjugate.cu
#include <cub/cub.cuh>

struct AdjArgs {
    const double *t; const int *c; int64_t n;
    const int *order;        // child events sorted by (node, time): event indices
    const int *node_ptr;     // [K+1] offsets into order
    const double *Mn;        // [K] events per node
    int K; const void *table_w;  // table without the adjacency factor
    const double *lambda0; const double *W; double *A;  // A: [K*K] parent-major as passed, updated in place
    const double *rho; const double *u; uint64_t seed, counter;
    double D, horizon, duration;
    int64_t cap;             // scratch entries per CTA
    int *ent_i; double *ent_v; double *lam; double *gacc;  // per-CTA scratch regions
    int64_t max_col;         // max child events per column
    int *flag;
    int col_begin, col_stride;  // this call owns the columns c with c % col_stride == col_begin
};

__device__ __forceinline__ int lo_of_event(const double *t, int64_t i, double horizon) {
    // first j with t[j] > t[i] - horizon (galloping backwards, as k_tile_lo)
    double thr = t[i] - horizon;
    int64_t good = i, bad = -1, step = 32;
    while (true) {
        int64_t cand = i - step;
        if (cand <= 0) { if (t[0] > thr) good = 0; else bad = 0; break; }
        if (t[cand] > thr) { good = cand; step <<= 1; } else { bad = cand; break; }
    }
    while (good - bad > 1) { int64_t mid = (good + bad) >> 1; if (t[mid] > thr) good = mid; else bad = mid; }
    return (int)good;
}

template <int KIND> __global__ void __launch_bounds__(256) k_adjacency(const AdjArgs a) {
    typedef typename EntryOf<KIND>::type E;
    extern __shared__ int s_dyn[];  // [K+1] bucket offsets, [K] cursors
    __shared__ FastTables s_ft;
    __shared__ double s_red[8];
    __shared__ double s_anew;
    int *s_off = s_dyn, *s_cur = s_dyn + a.K + 1;
    fast_tables_load(&s_ft);
    int *ent_i = a.ent_i + (size_t)blockIdx.x * a.cap;
    double *ent_v = a.ent_v + (size_t)blockIdx.x * a.cap;
    double *lam = a.lam + (size_t)blockIdx.x * a.max_col;
    double *gacc = a.gacc + (size_t)blockIdx.x * a.max_col;
    for (int c = a.col_begin + blockIdx.x * a.col_stride; c < a.K; c += gridDim.x * a.col_stride) {
        const int e0 = a.node_ptr[c], e1 = a.node_ptr[c + 1];
        const E *col = reinterpret_cast<const E *>(a.table_w) + (size_t)c * a.K;
        const double lam0 = a.lambda0[c];
        // ---- A1: bucket sizes
        for (int k = threadIdx.x; k <= a.K; k += blockDim.x) s_off[k] = 0;
        __syncthreads();
        // one warp per child event, lanes striding its window (coalesced reads of the time-sorted stream)
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int e = e0 + wid; e < e1; e += nw) {
            int i = a.order[e];
            double thr = a.t[i] - a.horizon;
            for (int j = i - 1 - lane; j >= 0; j -= 32) {
                if (!(__ldg(a.t + j) > thr)) break;
                atomicAdd(&s_off[__ldg(a.c + j) + 1], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // exclusive scan (K is small next to the event work)
            int run = 0;
            for (int k = 0; k < a.K; k++) { int v = s_off[k + 1]; s_off[k] = run; s_cur[k] = run; run += v; }
            s_off[a.K] = run;
            if (run > a.cap) atomicOr(a.flag, 32);
        }
        __syncthreads();
        if (s_off[a.K] > a.cap) continue;  // cannot happen: cap is sized from the exact per-column totals
        // ---- A2: emit (event, G) entries and the current intensities
        for (int e = e0 + wid; e < e1; e += nw) {
            int i = a.order[e];
            double ti = a.t[i], thr = ti - a.horizon, s = 0.0;
            for (int j = i - 1 - lane; j >= 0; j -= 32) {
                double tj = __ldg(a.t + j);
                if (!(tj > thr)) break;
                int p = __ldg(a.c + j);
                double v = pair_value(load_entry(col + p), ti - tj, a.D, &s_ft);
                int pos = atomicAdd(&s_cur[p], 1);
                ent_i[pos] = e - e0;
                ent_v[pos] = v;
                s += a.A[p + (int64_t)a.K * c] * v;
            }
            s = warp_sum(s);
            if (lane == 0) { lam[e - e0] = lam0 + s; gacc[e - e0] = 0.0; }
        }
        __syncthreads();
        // ---- B: K sequential Bernoulli steps
        for (int p = 0; p < a.K; p++) {
            const int b0 = s_off[p], b1 = s_off[p + 1];
            const int64_t kk = p + (int64_t)a.K * c;
            const double a_old = a.A[kk];
            double part = 0.0;
            if (b1 > b0) {
                // aggregate G_i[p] over the (rare) events that hold node p more than once in their window:
                // fire-and-forget reductions, no ownership protocol
                for (int e = b0 + threadIdx.x; e < b1; e += blockDim.x) {
                    const double v = ent_v[e];
                    if (v > 0.0) red_add_f64(&gacc[ent_i[e]], v);
                }
                __syncthreads();
                // every entry carries its share v/g of the event's term log((base + g)/base)
                for (int e = b0 + threadIdx.x; e < b1; e += blockDim.x) {
                    const double v = ent_v[e];
                    if (v > 0.0) {
                        const int ii = ent_i[e];
                        const double g = gacc[ii], l = lam[ii];
                        const double base = a_old != 0.0 ? l - g : l;
                        const double term = log((base + g) / base);
                        part += (v == g) ? term : term * (v / g);
                    }
                }
                part = warp_sum(part);
                if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                double sum = 0.0;
                if (b1 > b0) for (int w = 0; w < 8; w++) sum += s_red[w];
                double rho = a.rho[kk];
                // ll1 - ll0 (continuous.jl:477-483): integrated-intensity difference, log-intensity difference, prior
                double delta = -a.W[kk] * a.Mn[p] + sum + (log(rho) - log(1.0 - rho));
                double p1 = delta >= 0.0 ? 1.0 / (1.0 + exp(-delta)) : exp(delta) / (1.0 + exp(delta));
                if (delta != delta) { atomicOr(a.flag, 64); p1 = 0.0; }
                double uu = a.u ? a.u[kk] : philox_uniform(a.seed, (uint64_t)kk, a.counter);
                double an = uu <= p1 ? 1.0 : 0.0;  // rand(Bernoulli(p)) = rand() <= p
                a.A[kk] = an;
                s_anew = an;
            }
            __syncthreads();
            if (b1 > b0) {
                const double an = s_anew;
                const double sgn = an - a_old;  // +1 link switched on, -1 switched off, 0 unchanged
                for (int e = b0 + threadIdx.x; e < b1; e += blockDim.x) {
                    const double v = ent_v[e];
                    if (v > 0.0) {
                        const int ii = ent_i[e];
                        if (sgn != 0.0) red_add_f64(&lam[ii], sgn * v);
                        gacc[ii] = 0.0;
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------
// cached structure: build once, sweep many times
// ---------------------------------------------------------------------------------------
constexpr int ADJ_WMAX = 256;  // window entries per event checked exactly for repeated parents (longer windows: all flagged)

struct AdjBuildArgs {
    const double *t; const int *c;
    const int *order, *node_ptr;
    int K; double horizon;
    const int64_t *col;      // [K+1] first entry of every column
    int *boff;               // [K][K+1]
    unsigned *ent_i; double *ent_dt;
    int col_begin, col_stride;
};

__global__ void __launch_bounds__(256) k_adj_build(const AdjBuildArgs a) {
    extern __shared__ int s_dyn[];  // [K+1] bucket offsets, [K] cursors, [8][ADJ_WMAX] window nodes per warp
    int *s_off = s_dyn, *s_cur = s_dyn + a.K + 1, *s_win = s_dyn + 2 * a.K + 2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int *win = s_win + wid * ADJ_WMAX;
    for (int c = a.col_begin + blockIdx.x * a.col_stride; c < a.K; c += gridDim.x * a.col_stride) {
        const int e0 = a.node_ptr[c], e1 = a.node_ptr[c + 1];
        unsigned *ent_i = a.ent_i + a.col[c];
        double *ent_dt = a.ent_dt + a.col[c];
        for (int k = threadIdx.x; k <= a.K; k += blockDim.x) s_off[k] = 0;
        __syncthreads();
        for (int e = e0 + wid; e < e1; e += nw) {
            const int i = a.order[e];
            const double thr = a.t[i] - a.horizon;
            for (int j = i - 1 - lane; j >= 0; j -= 32) {
                if (!(__ldg(a.t + j) > thr)) break;
                atomicAdd(&s_off[__ldg(a.c + j) + 1], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int k = 0; k < a.K; k++) { int v = s_off[k + 1]; s_off[k] = run; s_cur[k] = run; run += v; }
            s_off[a.K] = run;
        }
        __syncthreads();
        for (int k = threadIdx.x; k <= a.K; k += blockDim.x) a.boff[(int64_t)c * (a.K + 1) + k] = s_off[k];
        for (int e = e0 + wid; e < e1; e += nw) {
            const int i = a.order[e];
            const double ti = a.t[i], thr = ti - a.horizon;
            // the window's parent nodes into the warp's buffer (as far as it reaches) for the exact repeat check
            int wl = 0;
            for (int j0 = i - 1; j0 >= 0; j0 -= 32) {
                const int j = j0 - lane;
                const bool in = j >= 0 && __ldg(a.t + j) > thr;
                const int p = in ? __ldg(a.c + j) : -1;
                if (wl + lane < ADJ_WMAX) win[wl + lane] = p;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                wl += __popc(m);
                if (m != 0xffffffffu) break;
            }
            __syncwarp();
            const bool exact = wl <= ADJ_WMAX;
            for (int k = lane; k < wl; k += 32) {
                const int j = i - 1 - k;
                const int p = __ldg(a.c + j);
                bool dup = !exact;
                if (exact) for (int m = 0; m < wl; m++) dup |= (m != k) & (win[m] == p);
                const int pos = atomicAdd(&s_cur[p], 1);
                ent_i[pos] = (unsigned)(e - e0) | (dup ? 0x80000000u : 0u);
                ent_dt[pos] = ti - __ldg(a.t + j);
            }
            __syncwarp();
        }
        __syncthreads();
    }
}

struct AdjSweepArgs {
    const int *node_ptr; const double *Mn;
    int K; const void *table_w;
    const double *lambda0; const double *W; double *A;
    const double *rho; const double *u; uint64_t seed, counter;
    double D;
    const int64_t *col; const int *boff; const unsigned *ent_i; const double *ent_dt;
    double *lam, *gacc, *vbuf;          // per-CTA scratch: [max_col], [max_col], [max_bucket]
    int64_t max_col, max_bucket;
    int *flag;
    int col_begin, col_stride;
};

template <int KIND> __global__ void __launch_bounds__(1024) k_adj_sweep(const AdjSweepArgs a) {
    typedef typename EntryOf<KIND>::type E;
    __shared__ FastTables s_ft;
    __shared__ double s_red[32];
    __shared__ double s_anew;
    fast_tables_load(&s_ft);
    const FastTables *ft = &s_ft;
    double *lam = a.lam + (size_t)blockIdx.x * a.max_col;
    double *gacc = a.gacc + (size_t)blockIdx.x * a.max_col;
    double *vbuf = a.vbuf + (size_t)blockIdx.x * a.max_bucket;
    __syncthreads();
    for (int c = a.col_begin + blockIdx.x * a.col_stride; c < a.K; c += gridDim.x * a.col_stride) {
        const int ne = a.node_ptr[c + 1] - a.node_ptr[c];
        const E *col = reinterpret_cast<const E *>(a.table_w) + (size_t)c * a.K;
        const int *boff = a.boff + (int64_t)c * (a.K + 1);
        const unsigned *ent_i = a.ent_i + a.col[c];
        const double *ent_dt = a.ent_dt + a.col[c];
        const double lam0 = a.lambda0[c];
        for (int e = threadIdx.x; e < ne; e += blockDim.x) { lam[e] = lam0; gacc[e] = 0.0; }
        __syncthreads();
        // current intensities: contributions of the links that are on
        for (int p = 0; p < a.K; p++) {
            if (a.A[p + (int64_t)a.K * c] == 0.0) continue;  // block-uniform
            const E en = load_entry(col + p);
            for (int e = boff[p] + threadIdx.x; e < boff[p + 1]; e += blockDim.x) {
                const double v = pair_value(en, __ldg(ent_dt + e), a.D, ft);
                if (v > 0.0) red_add_f64(&lam[__ldg(ent_i + e) & 0x7fffffffu], v);
            }
        }
        __syncthreads();
        // K sequential Bernoulli steps
        for (int p = 0; p < a.K; p++) {
            const int b0 = boff[p], b1 = boff[p + 1];
            const int64_t kk = p + (int64_t)a.K * c;
            const double a_old = a.A[kk];
            double part = 0.0;
            int anydup = 0;
            if (b1 > b0) {
                const E en = load_entry(col + p);
                bool mydup = false;
                for (int e = b0 + threadIdx.x; e < b1; e += blockDim.x) {
                    const unsigned ii = __ldg(ent_i + e);
                    const double v = pair_value(en, __ldg(ent_dt + e), a.D, ft);
                    vbuf[e - b0] = v;
                    if ((ii & 0x80000000u) && v > 0.0) { red_add_f64(&gacc[ii & 0x7fffffffu], v); mydup = true; }
                }
                anydup = __syncthreads_or(mydup);  // the aggregated G_i[p] must be complete before it is read
                for (int e = b0 + threadIdx.x; e < b1; e += blockDim.x) {
                    const double v = vbuf[e - b0];
                    if (v > 0.0) {
                        const unsigned ii = __ldg(ent_i + e);
                        const bool dup = (ii & 0x80000000u) != 0u;
                        const int ie = (int)(ii & 0x7fffffffu);
                        const double g = dup ? gacc[ie] : v, l = lam[ie];
                        const double base = a_old != 0.0 ? l - g : l;
                        const double term = log((base + g) / base);
                        part += (v == g) ? term : term * (v / g);  // an entry carries its share v/g of the event's term
                    }
                }
                part = warp_sum(part);
                if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                double sum = 0.0;
                if (b1 > b0) for (int w = 0; w < (int)(blockDim.x >> 5); w++) sum += s_red[w];
                const double rho = a.rho[kk];
                // ll1 - ll0 (continuous.jl:477-483): integrated-intensity difference, log-intensity difference, prior
                const double delta = -a.W[kk] * a.Mn[p] + sum + (log(rho) - log(1.0 - rho));
                double p1 = delta >= 0.0 ? 1.0 / (1.0 + exp(-delta)) : exp(delta) / (1.0 + exp(delta));
                if (delta != delta) { atomicOr(a.flag, 64); p1 = 0.0; }
                const double uu = a.u ? a.u[kk] : philox_uniform(a.seed, (uint64_t)kk, a.counter);
                const double an = uu <= p1 ? 1.0 : 0.0;  // rand(Bernoulli(p)) = rand() <= p
                a.A[kk] = an;
                s_anew = an;
            }
            __syncthreads();
            if (b1 > b0) {
                const double sgn = s_anew - a_old;  // +1 link switched on, -1 switched off, 0 unchanged
                if (sgn != 0.0 || anydup) {
                    for (int e = b0 + threadIdx.x; e < b1; e += blockDim.x) {
                        const double v = vbuf[e - b0];
                        if (v > 0.0) {
                            const unsigned ii = __ldg(ent_i + e);
                            if (sgn != 0.0) red_add_f64(&lam[ii & 0x7fffffffu], sgn * v);
                            if (ii & 0x80000000u) gacc[ii & 0x7fffffffu] = 0.0;
                        }
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}

// per-column totals -> exclusive scan (K values; single thread) and the largest bucket of every column
__global__ void k_adj_scan(const unsigned long long *__restrict__ colcount, int K, int col_begin, int col_stride, int64_t *__restrict__ col) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int k = 0; k < K; k++) { col[k] = run; if (k % col_stride == col_begin) run += (int64_t)colcount[k]; }  // owned columns only
        col[K] = run;
    }
}
__global__ void k_adj_max_bucket(const int *__restrict__ boff, int K, unsigned long long *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * K) return;
    const int c = (int)(i / K), p = (int)(i % K);
    const int len = boff[(int64_t)c * (K + 1) + p + 1] - boff[(int64_t)c * (K + 1) + p];
    if (len > 0) atomicMax(out, (unsigned long long)len);
}

// per-column entry totals: colcount[c_i] += window length of event i
__global__ void k_adj_count(const double *__restrict__ t, const int *__restrict__ c, int64_t n, double horizon, unsigned long long *__restrict__ colcount) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = lo_of_event(t, i, horizon);
    if (i > lo) atomicAdd(&colcount[c[i]], (unsigned long long)(i - lo));
}

__global__ void k_iota(int *v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int)i;
}
__global__ void k_node_ptr(const double *__restrict__ Mn, int K, int *__restrict__ ptr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < K; k++) { ptr[k] = run; run += (int)Mn[k]; }
        ptr[K] = run;
    }
}

// table without the adjacency factor (the sampler needs W h for both values of A[p,c])
__global__ void k_table_noA_ln(int K, const double *__restrict__ W, const double *__restrict__ mu, const double *__restrict__ tau, double D, EntryLN *__restrict__ table) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)K * K) return;
    int cc = (int)(e / K), p = (int)(e % K);
    int64_t src = p + (int64_t)K * cc;
    double tt = tau[src];
    EntryLN en;
    en.cf = W[src] * sqrt(tt) * NHP_INVSQRT2PI * (D * D); en.mu = mu[src]; en.h = 0.5 * tt; en.pad = 0.0;
    table[e] = en;
}
__global__ void k_table_noA_ex(int K, const double *__restrict__ W, const double *__restrict__ theta, EntryEX *__restrict__ table) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)K * K) return;
    int cc = (int)(e / K), p = (int)(e % K);
    int64_t src = p + (int64_t)K * cc;
    EntryEX en;
    en.theta = theta[src]; en.wt = W[src] * en.theta;
    table[e] = en;
}

extern "C" int nhp_cont_resample_adjacency(nhp_ctx *ctx, nhp_events *ev, const double *rho, uint64_t seed, uint64_t counter, const double *u,
                                           double *A_inout) {
    return nhp_cont_resample_adjacency_cols(ctx, ev, rho, seed, counter, u, A_inout, 0, 1);
}

extern "C" int nhp_cont_resample_adjacency_cols(nhp_ctx *ctx, nhp_events *ev, const double *rho, uint64_t seed, uint64_t counter, const double *u,
                                                double *A_inout, int64_t col_begin, int64_t col_stride) {
    NHP_CHECK(ctx, ctx != nullptr, NHP_ERR_INVALID, "ctx is NULL");
    NHP_CHECK(ctx, col_stride >= 1 && col_begin >= 0 && col_begin < col_stride, NHP_ERR_INVALID, "nhp_cont_resample_adjacency_cols: need 0 <= col_begin < col_stride");
    NHP_CHECK(ctx, ctx->cont_set, NHP_ERR_STATE, "continuous parameters not set (call nhp_cont_params_set)");
    NHP_CHECK(ctx, ev != nullptr && ev->K == ctx->K, NHP_ERR_INVALID, "nhp_cont_resample_adjacency: bad events handle");
    NHP_CHECK(ctx, rho && A_inout, NHP_ERR_INVALID, "nhp_cont_resample_adjacency: NULL rho/A");
    NHP_CHECK(ctx, ev->n_halo == 0 && ev->index_base == 0, NHP_ERR_UNSUPPORTED,
              "nhp_cont_resample_adjacency works on unsharded data (multi-GPU partitions the columns, not the time axis)");
    NHP_CUDA(ctx, cudaSetDevice(ctx->device));
    NHP_CUDA(ctx, fast_tables_upload(ctx->stream));
    const int64_t K = ctx->K, KK = K * K, n = ev->n;
    cudaStream_t s = ctx->stream;
    double horizon = nhp_cont_horizon_value(ctx, n, 0);
    // ---- device buffers (freed at the end; the sampler is called once per Gibbs sweep)
    double *d_rho = nullptr, *d_u = nullptr, *d_A = nullptr;
    void *d_tw = nullptr, *d_sort = nullptr;
    int *d_keys = nullptr, *d_vals = nullptr, *d_order = nullptr, *d_ptr = nullptr, *d_ckeys = nullptr;
    unsigned long long *d_cc = nullptr;
    int *d_ent_i = nullptr; double *d_ent_v = nullptr, *d_lam = nullptr, *d_gacc = nullptr;
    auto fin = [&](int rc) {
        cudaStreamSynchronize(s);
        cudaFree(d_rho); cudaFree(d_u); cudaFree(d_A); cudaFree(d_tw); cudaFree(d_sort); cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_order);
        cudaFree(d_ptr); cudaFree(d_ckeys); cudaFree(d_cc); cudaFree(d_ent_i); cudaFree(d_ent_v); cudaFree(d_lam); cudaFree(d_gacc);
        return rc;
    };
#define ADJ_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fin(nhp_fail(ctx, NHP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__))); } while (0)
    ADJ_CUDA(cudaMalloc(&d_rho, KK * sizeof(double)));
    ADJ_CUDA(cudaMalloc(&d_A, KK * sizeof(double)));
    ADJ_CUDA(cudaMalloc(&d_tw, KK * sizeof(EntryLN)));
    ADJ_CUDA(cudaMalloc(&d_ptr, (K + 1) * sizeof(int)));
    ADJ_CUDA(cudaMalloc(&d_cc, K * sizeof(unsigned long long)));
    ADJ_CUDA(cudaMemcpyAsync(d_rho, rho, KK * sizeof(double), cudaMemcpyHostToDevice, s));
    ADJ_CUDA(cudaMemcpyAsync(d_A, A_inout, KK * sizeof(double), cudaMemcpyHostToDevice, s));
    if (u) { ADJ_CUDA(cudaMalloc(&d_u, KK * sizeof(double))); ADJ_CUDA(cudaMemcpyAsync(d_u, u, KK * sizeof(double), cudaMemcpyHostToDevice, s)); }
    unsigned kb = (unsigned)((KK + 255) / 256);
    if (ctx->kind == NHP_LOGITNORMAL) k_table_noA_ln<<<kb, 256, 0, s>>>((int)K, ctx->d_W, ctx->d_p1, ctx->d_p2, ctx->dtmax, (EntryLN *)d_tw);
    else k_table_noA_ex<<<kb, 256, 0, s>>>((int)K, ctx->d_W, ctx->d_p1, (EntryEX *)d_tw);
    NHP_LAUNCHED(ctx);
    ADJ_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), s));
    int64_t max_entries = 0, max_col = 0;
    // ---- child events grouped by node, time order kept: the cached by-node index of the events handle
    { int rc = nhp_events_build_node_index(ctx, ev); if (rc != NHP_OK) return fin(rc); }
    const char *envc = getenv("NHP_ADJ_CACHE");
    bool cached = n > 0 && !(envc && atoi(envc) == 0);
    // The cached structure stays valid for any horizon it covers: extra pairs beyond the requested cut-off are genuine
    // predecessors whose (tiny) contributions are simply included.  With a parameter-dependent horizon (Exponential cut-off,
    // which moves with every conjugate draw) it is built with a 25 % margin so that a chain does not rebuild it every sweep.
    const bool moving = ctx->kind == NHP_EXPONENTIAL && horizon < ctx->dtmax;
    const bool have_cache = cached && ev->d_adj_i && ev->adj_cb == (int)col_begin && ev->adj_cs == (int)col_stride &&
                            (moving ? (ev->adj_horizon >= horizon && ev->adj_horizon <= 2.0 * horizon) : ev->adj_horizon == horizon);
    if (cached && !have_cache && moving) horizon = std::min(ctx->dtmax, 1.25 * horizon);
    else if (have_cache) horizon = ev->adj_horizon;
    if (n > 0) {
        std::vector<double> mn(K);
        ADJ_CUDA(cudaMemcpyAsync(mn.data(), ev->d_Mn, K * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (!have_cache) {  // per-column window totals: only needed to size the buckets
            ADJ_CUDA(cudaMemsetAsync(d_cc, 0, K * sizeof(unsigned long long), s));
            k_adj_count<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ev->d_t, ev->d_c, n, horizon, d_cc);
            NHP_LAUNCHED(ctx);
            std::vector<unsigned long long> cc(K);
            ADJ_CUDA(cudaMemcpyAsync(cc.data(), d_cc, K * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            ADJ_CUDA(cudaStreamSynchronize(s));
            for (int64_t k = 0; k < K; k++) max_entries = std::max<int64_t>(max_entries, (int64_t)cc[k]);
            if (max_entries >= (int64_t)0x3fffffff) return fin(nhp_fail(ctx, NHP_ERR_UNSUPPORTED, "adjacency sampler: a column has %lld window entries (limit 2^30)", (long long)max_entries));
        }
        ADJ_CUDA(cudaStreamSynchronize(s));
        for (int64_t k = 0; k < K; k++) max_col = std::max<int64_t>(max_col, (int64_t)mn[k]);
    }
    int grid = (int)std::min<int64_t>(K, (int64_t)ctx->sm_count * 8);  // latency-bound phases: as many columns in flight as the scratch area allows
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    int64_t cap = std::max<int64_t>(max_entries, 1), mc = std::max<int64_t>(max_col, 1);
    // ---- cached structure: (re)build when the data handle has none for this horizon / column partition
    if (cached && !have_cache) {
        cudaFree(ev->d_adj_i); cudaFree(ev->d_adj_dt); cudaFree(ev->d_adj_boff); cudaFree(ev->d_adj_col);
        ev->d_adj_i = nullptr; ev->d_adj_dt = nullptr; ev->d_adj_boff = nullptr; ev->d_adj_col = nullptr; ev->adj_horizon = -1.0;
        cudaMemGetInfo(&free_b, &total_b);
        int64_t tot = 0;
        {
            std::vector<unsigned long long> cc(K);
            ADJ_CUDA(cudaMemcpy(cc.data(), d_cc, K * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            for (int64_t k = 0; k < K; k++) if (k % col_stride == col_begin) tot += (int64_t)cc[k];
        }
        const size_t need = (size_t)tot * 12 + (size_t)K * (K + 1) * sizeof(int) + (size_t)(K + 1) * sizeof(int64_t);
        const size_t bsmem = (size_t)(2 * K + 2 + 8 * ADJ_WMAX) * sizeof(int);
        if (need > free_b / 2 + free_b / 4 || bsmem > (size_t)ctx->smem_optin - 1024) cached = false;  // keep room for the sweep's scratch: fall back to the uncached kernel
        else {
            ADJ_CUDA(cudaMalloc(&ev->d_adj_i, std::max<size_t>((size_t)tot, 1) * sizeof(unsigned)));
            ADJ_CUDA(cudaMalloc(&ev->d_adj_dt, std::max<size_t>((size_t)tot, 1) * sizeof(double)));
            ADJ_CUDA(cudaMalloc(&ev->d_adj_boff, (size_t)K * (K + 1) * sizeof(int)));
            ADJ_CUDA(cudaMalloc(&ev->d_adj_col, (size_t)(K + 1) * sizeof(int64_t)));
            ADJ_CUDA(cudaMemsetAsync(ev->d_adj_boff, 0, (size_t)K * (K + 1) * sizeof(int), s));
            k_adj_scan<<<1, 32, 0, s>>>(d_cc, (int)K, (int)col_begin, (int)col_stride, ev->d_adj_col);
            NHP_LAUNCHED(ctx);
            AdjBuildArgs b;
            b.t = ev->d_t; b.c = ev->d_c; b.order = ev->d_order; b.node_ptr = ev->d_node_ptr; b.K = (int)K; b.horizon = horizon;
            b.col = ev->d_adj_col; b.boff = ev->d_adj_boff; b.ent_i = ev->d_adj_i; b.ent_dt = ev->d_adj_dt; b.col_begin = (int)col_begin; b.col_stride = (int)col_stride;
            if (bsmem > 48 * 1024) ADJ_CUDA(cudaFuncSetAttribute(k_adj_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
            k_adj_build<<<grid, 256, bsmem, s>>>(b);
            NHP_LAUNCHED(ctx);
            ADJ_CUDA(cudaMemsetAsync(d_cc, 0, sizeof(unsigned long long), s));
            k_adj_max_bucket<<<(unsigned)((KK + 255) / 256), 256, 0, s>>>(ev->d_adj_boff, (int)K, d_cc);
            NHP_LAUNCHED(ctx);
            unsigned long long mb = 0;
            ADJ_CUDA(cudaMemcpyAsync(&mb, d_cc, sizeof(mb), cudaMemcpyDeviceToHost, s));
            ADJ_CUDA(cudaStreamSynchronize(s));
            ADJ_CUDA(cudaGetLastError());
            ev->adj_total = tot; ev->adj_max_bucket = (int64_t)mb; ev->adj_max_col = mc; ev->adj_horizon = horizon;
            ev->adj_cb = (int)col_begin; ev->adj_cs = (int)col_stride;
            cudaMemGetInfo(&free_b, &total_b);
        }
    }
    AdjArgs a;
    a.t = ev->d_t; a.c = ev->d_c; a.n = n; a.order = ev->d_order; a.node_ptr = ev->d_node_ptr; a.Mn = ev->d_Mn; a.K = (int)K; a.table_w = d_tw;
    a.lambda0 = ctx->d_lambda0; a.W = ctx->d_W; a.A = d_A; a.rho = d_rho; a.u = d_u; a.seed = seed; a.counter = counter;
    a.D = ctx->dtmax; a.horizon = horizon; a.duration = ev->duration; a.flag = ctx->d_flag; a.col_begin = (int)col_begin; a.col_stride = (int)col_stride;
    int rc = NHP_OK;
    if (cached) {
        const int64_t mb = std::max<int64_t>(ev->adj_max_bucket, 1);
        const size_t per_cta = (size_t)(2 * mc + mb) * sizeof(double);
        // Threads per column: the per-column intensity array lam[] (8 B per child event) is gathered at random once per entry.
        // With 256-thread CTAs ~8 columns share an SM; once their arrays exceed L2 many times over (1e8 events at K = 1000:
        // 1.9 GB) every gather is a DRAM sector and fewer, larger CTAs win (measured: 163 vs 216 ms); below that the
        // 256-thread CTAs are faster (1e7 events: 17 vs 30 ms).
        int bs = 256;
        {
            const double lam_bytes = (double)mc * 16.0;  // lam + gacc
            if (lam_bytes * ctx->sm_count * 8 > 1e9) bs = 1024;
            const char *envb = getenv("NHP_ADJ_BLOCK");
            if (envb && (atoi(envb) == 256 || atoi(envb) == 512 || atoi(envb) == 1024)) bs = atoi(envb);
            grid = (int)std::min<int64_t>(grid, (int64_t)ctx->sm_count * (2048 / bs));
        }
        while (grid > 1 && (size_t)grid * per_cta > free_b / 2) grid = (grid + 1) / 2;
        // per-CTA work areas from the context's persistent scratch buffer (no allocation per sweep)
        void *sc = nullptr;
        { int rcs = nhp_scratch(ctx, (size_t)grid * per_cta, &sc); if (rcs != NHP_OK) return fin(rcs); }
        double *w_lam = (double *)sc, *w_gacc = w_lam + (size_t)grid * mc, *w_vbuf = w_gacc + (size_t)grid * mc;
        AdjSweepArgs w;
        w.node_ptr = ev->d_node_ptr; w.Mn = ev->d_Mn; w.K = (int)K; w.table_w = d_tw; w.lambda0 = ctx->d_lambda0; w.W = ctx->d_W; w.A = d_A; w.rho = d_rho; w.u = d_u;
        w.seed = seed; w.counter = counter; w.D = ctx->dtmax; w.col = ev->d_adj_col; w.boff = ev->d_adj_boff; w.ent_i = ev->d_adj_i; w.ent_dt = ev->d_adj_dt;
        w.lam = w_lam; w.gacc = w_gacc; w.vbuf = w_vbuf; w.max_col = mc; w.max_bucket = mb; w.flag = ctx->d_flag; w.col_begin = (int)col_begin; w.col_stride = (int)col_stride;
        rc = nhp_timer_begin(ctx);
        if (rc != NHP_OK) return fin(rc);
        if (ctx->kind == NHP_LOGITNORMAL) k_adj_sweep<NHP_LOGITNORMAL><<<grid, bs, 0, s>>>(w);
        else k_adj_sweep<NHP_EXPONENTIAL><<<grid, bs, 0, s>>>(w);
    } else {
    // bound the scratch area: entries cost 12 B per CTA slot
    while (grid > 1 && (size_t)grid * ((size_t)cap * 12 + (size_t)mc * 16) > free_b / 2) grid = (grid + 1) / 2;
    if ((size_t)grid * ((size_t)cap * 12 + (size_t)mc * 16) > free_b / 2)
        return fin(nhp_fail(ctx, NHP_ERR_UNSUPPORTED, "adjacency sampler: one column needs %lld window entries, more than the free device memory holds", (long long)cap));
    ADJ_CUDA(cudaMalloc(&d_ent_i, (size_t)grid * cap * sizeof(int)));
    ADJ_CUDA(cudaMalloc(&d_ent_v, (size_t)grid * cap * sizeof(double)));
    ADJ_CUDA(cudaMalloc(&d_lam, (size_t)grid * mc * sizeof(double)));
    ADJ_CUDA(cudaMalloc(&d_gacc, (size_t)grid * mc * sizeof(double)));
    a.cap = cap; a.ent_i = d_ent_i; a.ent_v = d_ent_v; a.lam = d_lam; a.gacc = d_gacc; a.max_col = mc;
    size_t smem = (size_t)(2 * K + 2) * sizeof(int);
    rc = nhp_timer_begin(ctx);
    if (rc != NHP_OK) return fin(rc);
    if (ctx->kind == NHP_LOGITNORMAL) {
        if (smem > 48 * 1024) ADJ_CUDA(cudaFuncSetAttribute(k_adjacency<NHP_LOGITNORMAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_adjacency<NHP_LOGITNORMAL><<<grid, 256, smem, s>>>(a);
    } else {
        if (smem > 48 * 1024) ADJ_CUDA(cudaFuncSetAttribute(k_adjacency<NHP_EXPONENTIAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_adjacency<NHP_EXPONENTIAL><<<grid, 256, smem, s>>>(a);
    }
    }
    NHP_LAUNCHED(ctx);
    ADJ_CUDA(cudaGetLastError());
    int flag = 0;
    ADJ_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    rc = nhp_timer_end(ctx);
    if (rc != NHP_OK) return fin(rc);
    if (flag & 32) return fin(nhp_fail(ctx, NHP_ERR_CUDA, "adjacency sampler: scratch overflow (internal error)"));
    if (flag & 64) return fin(nhp_fail(ctx, NHP_ERR_NUMERIC, "adjacency sampler: NaN log-likelihood difference"));
    ADJ_CUDA(cudaMemcpyAsync(A_inout, d_A, KK * sizeof(double), cudaMemcpyDeviceToHost, s));
    // the new adjacency becomes the context's A (device to device; the masked tables are rebuilt below)
    if (ctx->has_A) ADJ_CUDA(cudaMemcpyAsync(ctx->d_A, d_A, KK * sizeof(double), cudaMemcpyDeviceToDevice, s));
    ADJ_CUDA(cudaStreamSynchronize(s));
#undef ADJ_CUDA
    fin(NHP_OK);
    if (ctx->has_A) {
        ctx->cont_set = false;
        ctx->sweep_ll_valid = false;
        NHP_TRY(nhp_cont_params_refresh(ctx));
        ctx->cont_set = true;
    }
    return NHP_OK;
}
