// cont_sparse.cu -- window sweeps for sparse effective weights (ContinuousNetworkHawkesProcess with a
// Bernoulli adjacency, continuous.jl:315-321, 521-525): a pair contributes a*w*pdf, which is exactly 0
// wherever A[p,c]*W[p,c] == 0, so only the active pairs need the FP64 impulse evaluation.
//
// Per CTA (tile of STE = 256 / SG consecutive child events, window staged by TMA exactly as the dense sweep):
//   0. gather  -- thread tid owns child event tid % STE and window share tid / STE, so a warp holds 32 consecutive
//                 events; each lane copies its child's adjacency bit row (256-bit loads) into its OWN shared-memory
//                 bank (word w of lane l at [w * 32 + l]) while the TMA transfer is in flight; one mbarrier covers both.
//   1. filter  -- SG (2 or 4) threads per child event, each walking a contiguous share of the window most-recent-first,
//                 eight probes per trip (node id -> row word -> bit, all bank-conflict free); hits go to a 64-bit mask.
//   2. compact -- block exclusive scan of the hit counts -> contiguous, ordered segment per (event, share).
//   3. evaluate-- all threads stride over the dense hit list: table gather (L2) + FP64 impulse.
//   4. combine -- one thread per event folds its segments in window order: log-likelihood term, or the
//                 inverse-cdf parent draw of parents.jl:25-46 (zero-weight entries can never be drawn and
//                 do not move the cumulative sum, so skipping them is exact) + fused statistics.
// Events whose window share exceeds the 64-bit mask, tiles with more hits than the list holds, and tiles
// whose window does not fit the staging buffer take the direct path (same arithmetic, in place).
#include "cont_sweep.cuh"

enum { SP_LOGLIK = 0, SP_INTENSITY = 1, SP_PARENTS = 2 };

struct SparseArgs {
    SweepArgs s;
    const uint32_t *abits;  // [K * words] child-major bit rows (bit p of row c <=> A*W*... != 0)
    int words;
    int cape;               // capacity of the per-tile hit list
    int m0_smem;            // baseline-count histogram lives in shared memory (K small enough)
    const unsigned short *wlen;  // [n - first] cached window length per own event (65535 = saturated)
};

template <int KIND, bool ST>
__device__ __forceinline__ double direct_sum(const SweepArgs &a, const Tile &tl, const FastTables *ft, int64_t i, double ti, int ci, int64_t jlo) {
    typedef typename EntryOf<KIND>::type E;
    const E *col = reinterpret_cast<const E *>(a.table) + (size_t)ci * a.K;
    const double thr = ti - a.horizon;
    double acc = 0.0;
    for (int64_t j = i - 1; j >= jlo; j--) {
        double tj = tile_T<ST>(a, tl, j);
        if (!(tj > thr)) break;
        acc += pair_value(load_entry(col + tile_C<ST>(a, tl, j)), ti - tj, a.D, ft);
    }
    return acc;
}

// inverse-cdf walk over the full window (direct path): returns i - j of the chosen parent, 0 = baseline
template <int KIND, bool ST>
__device__ __forceinline__ int direct_pick(const SweepArgs &a, const Tile &tl, const FastTables *ft, int64_t i, double ti, int ci, int64_t jlo, double target) {
    typedef typename EntryOf<KIND>::type E;
    const E *col = reinterpret_cast<const E *>(a.table) + (size_t)ci * a.K;
    const double thr = ti - a.horizon;
    double cum = 0.0;
    for (int64_t j = i - 1; j >= jlo; j--) {
        double tj = tile_T<ST>(a, tl, j);
        if (!(tj > thr)) break;
        cum += pair_value(load_entry(col + tile_C<ST>(a, tl, j)), ti - tj, a.D, ft);
        if (cum > target) return (int)(i - j);
    }
    return 0;
}

template <int KIND, bool ST>
__device__ __forceinline__ void finish_parent(const SweepArgs &a, const Tile &tl, int64_t i, double ti, int ci, double S, int chosen, int *m0_hist, bool use_hist) {
    const StatsLayout sl{a.K};
    if (!(S > 0.0) || S > 1.7976931348623157e308) atomicOr(a.flag, 8);
    a.poff[i] = chosen;
    if (chosen == 0) {
        if (use_hist) atomicAdd(m0_hist + ci, 1);  // most events are baseline events: keep that counter in shared memory
        else red_add_f64(a.stats + sl.off_M0() + ci, 1.0);
    }
    else {
        int64_t jp = i - chosen;
        int cj = tile_C<ST>(a, tl, jp);
        double dt = ti - tile_T<ST>(a, tl, jp);
        int64_t k = cj + (int64_t)a.K * ci;
        red_add_f64(a.stats + sl.off_Mnm() + k, 1.0);
        red_add_f64(a.stats + sl.off_S1() + k, KIND == NHP_LOGITNORMAL ? log_duration_dev(dt, a.D) : dt);
    }
}

constexpr int SQMAX = 64;               // window entries one filter thread can flag (64-bit hit mask)

// Persistent CTAs: each loops over tiles of STE child events so the per-CTA set-up (log/exp tables,
// barrier init) is paid once.
// SG = filter threads per child event (2 for short windows: twice the events per tile, half the per-tile
// overhead per event; 4 otherwise); STE = NHP_BLOCK / SG child events per tile.
// Thread tid works on event e = tid % STE, share g = tid / STE: a warp holds 32 CONSECUTIVE events with the same
// share index, so (i) its staged node-id loads touch 32 consecutive words and (ii) each lane can keep its
// child's adjacency bit row in its own shared-memory bank ("vertical" rows: word w of lane l's row at
// [w * 32 + l]) -- the random word probes of the filter are bank-conflict free.  The profile that led here
// (profiles/r01_ncu_sweep_kernels.md) showed the horizontal layout at 84% of the LSU wavefront peak, 58% of the
// wavefronts being bank conflicts.
template <int KIND, int MODE, int SG>
__global__ void __launch_bounds__(NHP_BLOCK, 5) k_sweep_sparse(const SparseArgs sa, const int64_t ntiles) {
    constexpr int STE = NHP_BLOCK / SG;
    typedef typename EntryOf<KIND>::type E;
    const SweepArgs &a = sa.s;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double red[16];
    __shared__ FastTables s_ft;
    __shared__ int s_wsum[NHP_BLOCK / 32];
    __shared__ int s_off[NHP_BLOCK], s_cnt[NHP_BLOCK];  // per filter thread: list offset, hit count (-1 = direct path)
    fast_tables_load(&s_ft);
    const FastTables *ft = &s_ft;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    double *st = reinterpret_cast<double *>(smem + 16);
    int *sc = reinterpret_cast<int *>(smem + 16 + (size_t)a.cap * 8 + 32);  // 8 zero ints in front: the filter may walk that far past a window
    if (threadIdx.x < 8) sc[-1 - (int)threadIdx.x] = 0;
    // one arrival for the TMA transaction + one per warp for its share of the bit-row gather
    if (threadIdx.x == 0) { mbar_init(bar, 1 + NHP_BLOCK / 32); fence_mbar_init(); }
    // offsets are rounded as integers (not through a pointer cast) so every access below stays a 32-bit LDS/STS
    const int wp = sa.words;  // words per bit row (multiple of 4)
    const uint32_t off_rows = (16u + (uint32_t)a.cap * 12u + 32u + 127u) & ~127u;
    const uint32_t off_list = off_rows + (uint32_t)wp * STE * 4u;
    const uint32_t off_val = (off_list + (uint32_t)sa.cape * 4u + 7u) & ~7u;
    uint32_t *rows = reinterpret_cast<uint32_t *>(smem + off_rows);           // [STE / 32][wp][32 lanes]
    uint32_t *list = reinterpret_cast<uint32_t *>(smem + off_list);           // [sa.cape] (event << 16) | (i - j)
    double *val = reinterpret_cast<double *>(smem + off_val);                 // [sa.cape]
    int *m0_hist = reinterpret_cast<int *>(smem + off_val + (uint32_t)sa.cape * 8u);  // [K], parents mode with m0_smem only
    const bool use_hist = MODE == SP_PARENTS && sa.m0_smem;
    if (use_hist) for (int k = threadIdx.x; k < a.K; k += NHP_BLOCK) m0_hist[k] = 0;
    const int e = threadIdx.x % STE, g = threadIdx.x / STE;
    const int lane = threadIdx.x & 31;
    uint32_t *myrow = rows + (size_t)(e >> 5) * wp * 32 + lane;  // word w of this event's bit row: myrow[w * 32]
    double sum_log = 0.0, sum_row = 0.0;
    uint32_t parity = 0;
    __syncthreads();

    constexpr int TQ = STE / NHP_TQ;  // tile_lo is kept per NHP_TQ events
    int64_t lo_next = blockIdx.x < ntiles ? a.tile_lo[(int64_t)blockIdx.x * TQ] : 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile tl;
        tl.i0 = a.first + tile * STE;
        tl.i1 = min(a.n, tl.i0 + (int64_t)STE);
        tl.lo = lo_next;
        if (tile + gridDim.x < ntiles) lo_next = a.tile_lo[(tile + gridDim.x) * TQ];  // prefetched a whole tile ahead
        tl.base = tl.lo & ~(int64_t)3;
        tl.st = st; tl.sc = sc;
        const int64_t cnt_stage = (tl.i1 - tl.base + 3) & ~(int64_t)3;
        tl.staged = cnt_stage <= a.cap;
        const int64_t jlo = max(tl.lo, a.jmin);
        const int64_t i = tl.i0 + e;
        const bool live = i < tl.i1;
        const int ib = (int)(i - tl.base);
        // the child's node id and its cached window length come straight from global memory so that the adjacency
        // bit-row gather (below) overlaps the TMA transfer instead of waiting for it
        const int ci = live ? __ldg(a.c + i) : 0;
        const int wraw = live ? (int)__ldg(sa.wlen + (i - a.first)) : 0;
        if (!tl.staged) {  // window larger than the staging buffer: direct path from global memory
            if (live && g == 0) {
                const double ti = __ldg(a.t + i);
                const double S = direct_sum<KIND, false>(a, tl, ft, i, ti, ci, jlo) + base_rate(a, i, ci);
                if (MODE == SP_LOGLIK) sum_log += log(S);
                else if (MODE == SP_INTENSITY) a.lam_out[i - a.first] = S;
                else {
                    const int64_t gi = a.index_base + i;
                    const double u = a.u ? __ldg(a.u + (i - a.first)) : philox_uniform(a.seed, (uint64_t)gi, a.counter);
                    int chosen = gi == 0 ? 0 : direct_pick<KIND, false>(a, tl, ft, i, ti, ci, jlo, u * S);
                    finish_parent<KIND, false>(a, tl, i, ti, ci, S, chosen, m0_hist, use_hist);
                    if (a.want_ll) sum_log += log(S);
                }
            }
            continue;  // nothing in shared memory was touched
        }
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (uint32_t)(cnt_stage * 12));
            bulk_g2s(st, a.t + tl.base, (uint32_t)(cnt_stage * 8), bar);
            bulk_g2s(sc, a.c + tl.base, (uint32_t)(cnt_stage * 4), bar);
        }
        // adjacency bit row of the lane's child into the lane's bank; the SG warps that share these 32 events split
        // the row's 16-byte pieces between them
        {
            const uint32_t *row = sa.abits + (size_t)ci * sa.words;  // rows are 32-byte multiples
            for (int j = g; j < (wp >> 3); j += SG) {              // 256-bit pieces: a quarter of the L1 tag look-ups of 64-bit ones
                uint32_t v[8];
                ldg256_u32(row + 8 * j, v);
                uint32_t *d = myrow + j * 256;
#pragma unroll
                for (int r = 0; r < 8; r++) d[32 * r] = v[r];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
        mbar_wait(bar, parity);
        parity ^= 1u;
        // ---- 1. filter -------------------------------------------------------------------------------------
        unsigned long long hits = 0ull;
        int k0 = 0;
        bool over = false;
        if (live) {
            // window length from the per-event cache (k_win_len), clipped to the first admissible parent
            const int wlen = min(wraw, ib - (int)(jlo - tl.base));
            // this thread's contiguous share of the window, most recent first: positions k0+1 .. k1
            const int q = (wlen + SG - 1) / SG;
            k0 = g * q;
            const int k1 = min(wlen, k0 + q);
            over = q > SQMAX || wraw >= 65535;  // the same for every share of the event
            if (!over && k1 > k0) {
                // Eight probes per trip and no remainder loop: a share is walked past its end in whole trips (the extra
                // entries are older staged events, or the zero pad in front of sc) and the excess bits are masked off.
                const int *src = sc + ib - k0 - 1;  // position k0+1+m is src[-m]
                const int cntk = k1 - k0;
#pragma unroll 1
                for (int m = 0; m < cntk; m += 8) {
                    const int p0 = src[-m], p1 = src[-m - 1], p2 = src[-m - 2], p3 = src[-m - 3];
                    const int p4 = src[-m - 4], p5 = src[-m - 5], p6 = src[-m - 6], p7 = src[-m - 7];
                    const uint32_t w0 = myrow[p0 & ~31], w1 = myrow[p1 & ~31], w2 = myrow[p2 & ~31], w3 = myrow[p3 & ~31];
                    const uint32_t w4 = myrow[p4 & ~31], w5 = myrow[p5 & ~31], w6 = myrow[p6 & ~31], w7 = myrow[p7 & ~31];
                    const uint32_t b8 = ((w0 >> (p0 & 31)) & 1u) | (((w1 >> (p1 & 31)) & 1u) << 1) | (((w2 >> (p2 & 31)) & 1u) << 2) |
                                        (((w3 >> (p3 & 31)) & 1u) << 3) | (((w4 >> (p4 & 31)) & 1u) << 4) | (((w5 >> (p5 & 31)) & 1u) << 5) |
                                        (((w6 >> (p6 & 31)) & 1u) << 6) | (((w7 >> (p7 & 31)) & 1u) << 7);
                    hits |= (unsigned long long)b8 << m;
                }
                hits &= ~0ull >> (64 - cntk);  // 1 <= cntk <= SQMAX = 64
            }
        }
        const int cnt = __popcll(hits);
        // ---- 2. compact: exclusive scan of cnt over the block; an event's hits form SG segments (one per share),
        //         each contiguous and in window order ----------------------------------------------------------
        int x = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) s_wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        int off = x - cnt, total = 0;
#pragma unroll
        for (int w = 0; w < NHP_BLOCK / 32; w++) { int v = s_wsum[w]; if (w < (int)(threadIdx.x >> 5)) off += v; total += v; }
        const bool fits = total <= sa.cape;  // block-uniform; otherwise every event of the tile takes the direct path
        if (fits) {
            unsigned long long h = hits;
            int o = off;
            while (h) {
                int b = __ffsll((long long)h) - 1;
                h &= h - 1;
                list[o++] = ((uint32_t)e << 16) | (uint32_t)(k0 + 1 + b);
            }
        }
        s_off[threadIdx.x] = off;
        s_cnt[threadIdx.x] = (over || !fits) ? -1 : cnt;
        __syncthreads();
        // ---- 3. evaluate the active pairs ------------------------------------------------------------------
        if (fits) {
            for (int qd = threadIdx.x; qd < total; qd += NHP_BLOCK) {
                uint32_t en = list[qd];
                int ibe = (int)(tl.i0 - tl.base) + (int)(en >> 16), k = (int)(en & 0xffff);
                const E *col = reinterpret_cast<const E *>(a.table) + (size_t)sc[ibe] * a.K;
                val[qd] = pair_value(load_entry(col + sc[ibe - k]), st[ibe] - st[ibe - k], a.D, ft);
            }
        }
        __syncthreads();
        // ---- 4. combine: the share-0 thread of each event (STE / 32 full warps) ---------------------------------
        if (g == 0 && live) {
            const double ti = st[ib];
            const bool direct = s_cnt[e] < 0;  // over / !fits are event- and tile-wide: share 0 tells
            double S = 0.0;
            if (direct) S = direct_sum<KIND, true>(a, tl, ft, i, ti, ci, jlo);
            else {
#pragma unroll
                for (int h = 0; h < SG; h++) {
                    const int o = s_off[h * STE + e], c = s_cnt[h * STE + e];
                    for (int k = 0; k < c; k++) S += val[o + k];
                }
            }
            S += base_rate(a, i, ci);
            if (MODE == SP_LOGLIK) sum_log += log(S);
            else if (MODE == SP_INTENSITY) a.lam_out[i - a.first] = S;
            else {
                const int64_t gi = a.index_base + i;
                const double u = a.u ? __ldg(a.u + (i - a.first)) : philox_uniform(a.seed, (uint64_t)gi, a.counter);
                const double target = u * S;
                int chosen = 0;
                if (gi != 0) {
                    if (direct) chosen = direct_pick<KIND, true>(a, tl, ft, i, ti, ci, jlo, target);
                    else {
                        double cum = 0.0;
#pragma unroll
                        for (int h = 0; h < SG; h++) {
                            const int o = s_off[h * STE + e], c = s_cnt[h * STE + e];
                            for (int k = 0; k < c && chosen == 0; k++) {
                                cum += val[o + k];
                                if (cum > target) chosen = (int)(list[o + k] & 0xffff);
                            }
                        }
                    }
                }
                finish_parent<KIND, true>(a, tl, i, ti, ci, S, chosen, m0_hist, use_hist);
                if (a.want_ll) sum_log += log(S);  // log-likelihood terms for free
            }
        }
        __syncthreads();  // staging buffers, rows, list and val are reused by the next tile
    }
    if (MODE == SP_LOGLIK || MODE == SP_PARENTS) {
        block_sum2(sum_log, sum_row, red);
        if (threadIdx.x == 0) { a.partials[2 * (size_t)blockIdx.x] = sum_log; a.partials[2 * (size_t)blockIdx.x + 1] = sum_row; }
    }
    if (use_hist) {
        const StatsLayout sl{a.K};
        for (int k = threadIdx.x; k < a.K; k += NHP_BLOCK)
            if (m0_hist[k]) red_add_f64(a.stats + sl.off_M0() + k, (double)m0_hist[k]);
    }
}

// ---------------------------------------------------------------------------------------
// host side: decide whether the sparse path applies and launch it
// ---------------------------------------------------------------------------------------
size_t nhp_sparse_smem(int cap, int words, int lcap, int m0_entries, int ste) {
    size_t b = 16 + (size_t)cap * 12 + 32 + 128;
    b += (size_t)words * ste * 4;
    b += (size_t)lcap * 4 + 8;
    b += (size_t)lcap * 8;
    b += (size_t)m0_entries * 4;
    return b;
}

template <typename KernelT> static int launch_sparse(nhp_ctx *ctx, KernelT kernel, int *grid_out, size_t smem, const SparseArgs &sa, int64_t ntiles) {
    NHP_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));  // always: static + dynamic may exceed the 48 KB default even when the dynamic part is small
    NHP_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 1;
    NHP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NHP_BLOCK, smem));
    int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * std::max(per_sm, 1));  // persistent: exactly one resident wave
    *grid_out = grid;
    kernel<<<grid, NHP_BLOCK, smem, ctx->stream>>>(sa, ntiles);
    NHP_LAUNCHED(ctx);
    NHP_CUDA(ctx, cudaGetLastError());
    return NHP_OK;
}

// mode: 0 loglik, 1 intensity, 2 parents.  Returns NHP_OK after launching (grid size in *grid_out: the number of
// log-likelihood partials), 1 if the sparse path does not apply, < 0 on error.
int nhp_cont_try_sparse(nhp_ctx *ctx, const nhp_events *ev, SweepArgs &a, int mode, int *grid_out) {
    const char *env = getenv("NHP_SPARSE");
    if (env && atoi(env) == 0) return 1;
    bool force = env && atoi(env) == 1;
    if (!force && !(ctx->density <= 0.25)) return 1;
    int words = (int)ctx->abits_words;
    // two filter threads per event when the windows are short enough for their 64-bit hit masks (128 events per tile)
    const char *sge = getenv("NHP_SPARSE_SG");
    int sg = (ev->mean_win * 1.5 + 16.0 <= 2.0 * SQMAX) ? 2 : 4;
    if (sge && (atoi(sge) == 2 || atoi(sge) == 4)) sg = atoi(sge);
    const int ste = NHP_BLOCK / sg;
    // capacity of the per-tile hit list: 2x the expected number of active pairs of a tile + 256, in steps of 256
    double expect = ev->mean_win * ctx->density * ste;
    int lcap = (int)std::min(16384.0, 256.0 * std::ceil((2.0 * expect + 256.0) / 256.0));
    if (lcap < 512) lcap = 512;
    int64_t own = ev->n - ev->n_halo;
    int64_t need = ((ev->max_win + ste + 8) + 3) & ~(int64_t)3;
    int cap = (int)std::min<int64_t>(need, 8192);
    int m0_entries = (mode == 2 && ctx->K <= 4096) ? (int)ctx->K : 0;
    size_t smem = nhp_sparse_smem(cap, words, lcap, m0_entries, ste);
    if (smem > 100 * 1024) {
        cap = (int)std::min<int64_t>(need, 1024);
        smem = nhp_sparse_smem(cap, words, lcap, m0_entries, ste);
        if (smem > (size_t)ctx->smem_optin - 4096) return 1;
    }
    int64_t ntiles = (own + ste - 1) / ste;
    if (ntiles == 0) return 1;
    NHP_CUDA(ctx, fast_tables_upload(ctx->stream));  // this translation unit's copy of the log/exp tables
    SparseArgs sa;
    a.te = ste; a.cap = cap;
    sa.s = a; sa.abits = ctx->d_abits; sa.words = words; sa.cape = lcap; sa.m0_smem = m0_entries > 0; sa.wlen = ev->d_wlen;
    int *grid = grid_out;
#define NHP_SPARSE_LAUNCH(KIND, SGV)                                                                                      \
    do {                                                                                                                  \
        if (mode == 0) return launch_sparse(ctx, k_sweep_sparse<KIND, SP_LOGLIK, SGV>, grid, smem, sa, ntiles);           \
        if (mode == 1) return launch_sparse(ctx, k_sweep_sparse<KIND, SP_INTENSITY, SGV>, grid, smem, sa, ntiles);        \
        return launch_sparse(ctx, k_sweep_sparse<KIND, SP_PARENTS, SGV>, grid, smem, sa, ntiles);                         \
    } while (0)
    if (ctx->kind == NHP_LOGITNORMAL) {
        if (sg == 2) NHP_SPARSE_LAUNCH(NHP_LOGITNORMAL, 2);
        NHP_SPARSE_LAUNCH(NHP_LOGITNORMAL, 4);
    }
    if (sg == 2) NHP_SPARSE_LAUNCH(NHP_EXPONENTIAL, 2);
    NHP_SPARSE_LAUNCH(NHP_EXPONENTIAL, 4);
#undef NHP_SPARSE_LAUNCH
}
