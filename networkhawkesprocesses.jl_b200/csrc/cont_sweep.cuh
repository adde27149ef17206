// cont_sweep.cuh -- shared pieces of the continuous predecessor-window sweeps:
// tile descriptor, TMA (cp.async.bulk) staging of the sliding window into shared memory,
// table-entry loads and group (sub-warp) reductions.
#pragma once
#include "nhp_internal.cuh"

constexpr int NHP_TQ = 64;      // granularity of the cached tile window starts
constexpr int NHP_BLOCK = 256;  // threads per CTA of the sweep kernels

struct SweepArgs {
    const double *t;        // [n] ascending event times
    const int *c;           // [n] 0-based nodes
    int64_t n;              // events in the handle (halo + own)
    int64_t first;          // first child event (= n_halo)
    int64_t index_base;     // global index of local event 0
    int64_t jmin;           // parents with local index < jmin are excluded (quirk Q6)
    const int *tile_lo;     // window start of event first + 64 b
    int te;                 // child events per CTA (multiple of 64)
    int cap;                // staging capacity in entries (multiple of 4)
    int K;
    const void *table;      // EntryLN / EntryEX [K*K] child-major
    const double *lambda0;  // [K]
    const double *lam0ev;   // [n] baseline rate at every event (grid baseline), or NULL: lambda0[node]
    const double *rowsum;   // [K]
    double D;               // dtmax of the LogitNormal support
    double horizon;         // look-back horizon of the window (dtmax, or the Exponential cut-off; may be +Inf)
    double *partials;       // [2 * gridDim.x] (log-sum, row-sum) per CTA
    double *lam_out;        // [n - first] per-event intensities or NULL
    int *poff;              // [n] parent offsets
    const double *u;        // [n - first] uniforms or NULL (Philox)
    uint64_t seed, counter;
    double *stats;          // StatsLayout buffer
    int *flag;              // device error flag
    int want_ll;            // parent sweep also accumulates the log-likelihood terms (NHP_OPT_SWEEP_LOGLIK)
};

// baseline rate of event i on node ci: lambda0_ci(t_i) (baselines.jl:328-336) precomputed per event, or the homogeneous lambda0_ci
__device__ __forceinline__ double base_rate(const SweepArgs &a, int64_t i, int ci) { return a.lam0ev ? __ldg(a.lam0ev + i) : __ldg(a.lambda0 + ci); }

// ---------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine; SASS UBLKCP)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {  // plain arrival (release at CTA scope)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// tile of consecutive child events [i0, i1) and its staged predecessor range [base, i1)
// ---------------------------------------------------------------------------------------
struct Tile {
    int64_t i0, i1, lo, base;
    const double *st;  // staged times  (entry j at st[j - base])
    const int *sc;     // staged nodes
    bool staged;
};

// smem layout: [0,16) mbarrier | times cap*8 | nodes cap*4
__device__ __forceinline__ size_t sweep_smem_bytes_dev(int cap) { return 16 + (size_t)cap * 12; }

__device__ __forceinline__ Tile stage_tile(const SweepArgs &a, unsigned char *smem) {
    Tile tl;
    tl.i0 = a.first + (int64_t)blockIdx.x * a.te;
    tl.i1 = min(a.n, tl.i0 + (int64_t)a.te);
    tl.lo = a.tile_lo[(int64_t)blockIdx.x * (a.te / NHP_TQ)];
    tl.base = tl.lo & ~(int64_t)3;
    int64_t cnt = (tl.i1 - tl.base + 3) & ~(int64_t)3;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    double *st = reinterpret_cast<double *>(smem + 16);
    int *sc = reinterpret_cast<int *>(smem + 16 + (size_t)a.cap * 8);
    tl.staged = cnt <= a.cap;
    tl.st = st;
    tl.sc = sc;
    if (tl.staged) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (uint32_t)(cnt * 12));
            bulk_g2s(st, a.t + tl.base, (uint32_t)(cnt * 8), bar);
            bulk_g2s(sc, a.c + tl.base, (uint32_t)(cnt * 4), bar);
        }
        mbar_wait(bar, 0);
    }
    return tl;
}

template <bool ST> __device__ __forceinline__ double tile_T(const SweepArgs &a, const Tile &tl, int64_t j) {
    return ST ? tl.st[j - tl.base] : __ldg(a.t + j);
}
template <bool ST> __device__ __forceinline__ int tile_C(const SweepArgs &a, const Tile &tl, int64_t j) {
    return ST ? tl.sc[j - tl.base] : __ldg(a.c + j);
}

// ---------------------------------------------------------------------------------------
// window view with 32-bit local offsets: entry (tl.base + jl) of the stream.  Staged tiles read shared
// memory through explicit 32-bit shared addresses (plain LDS, no generic-address arithmetic); the
// fallback reads global memory through the read-only path.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <bool ST> struct WinView {
    uint32_t st_addr, sc_addr;
    const double *gt;
    const int *gc;
    __device__ __forceinline__ WinView(const SweepArgs &a, const Tile &tl)
        : st_addr(smem_u32(tl.st)), sc_addr(smem_u32(tl.sc)), gt(a.t + tl.base), gc(a.c + tl.base) {}
    __device__ __forceinline__ double T(int jl) const { return ST ? lds_f64(st_addr + 8u * (uint32_t)jl) : __ldg(gt + jl); }
    __device__ __forceinline__ int C(int jl) const { return ST ? lds_s32(sc_addr + 4u * (uint32_t)jl) : __ldg(gc + jl); }
};

// persistent-CTA staging: one mbarrier reused across tiles (phase parity tracked by the caller)
struct Stager {
    uint64_t *bar;
    double *st;
    int *sc;
    uint32_t parity;
};
__device__ __forceinline__ Stager stager_init(const SweepArgs &a, unsigned char *smem) {
    Stager sg;
    sg.bar = reinterpret_cast<uint64_t *>(smem);
    sg.st = reinterpret_cast<double *>(smem + 16);
    sg.sc = reinterpret_cast<int *>(smem + 16 + (size_t)a.cap * 8);
    sg.parity = 0;
    if (threadIdx.x == 0) { mbar_init(sg.bar, 1); fence_mbar_init(); }
    return sg;  // caller: __syncthreads() before the first stage_tile_at
}
// all threads call; the previous tile's readers must have passed a __syncthreads() before
__device__ __forceinline__ Tile stage_tile_at(const SweepArgs &a, Stager &sg, int64_t tile) {
    Tile tl;
    tl.i0 = a.first + tile * a.te;
    tl.i1 = min(a.n, tl.i0 + (int64_t)a.te);
    tl.lo = a.tile_lo[tile * (a.te / NHP_TQ)];
    tl.base = tl.lo & ~(int64_t)3;
    const int64_t cnt = (tl.i1 - tl.base + 3) & ~(int64_t)3;
    tl.staged = cnt <= a.cap;
    tl.st = sg.st;
    tl.sc = sg.sc;
    if (tl.staged) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(sg.bar, (uint32_t)(cnt * 12));
            bulk_g2s(sg.st, a.t + tl.base, (uint32_t)(cnt * 8), sg.bar);
            bulk_g2s(sg.sc, a.c + tl.base, (uint32_t)(cnt * 4), sg.bar);
        }
        mbar_wait(sg.bar, sg.parity);
        sg.parity ^= 1u;
    }
    return tl;
}

// ---------------------------------------------------------------------------------------
// table loads through the read-only path (L1-resident for small K, L2 for K ~ 1000)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ EntryLN load_entry(const EntryLN *p) {
    // one 256-bit read-only load per 32 B entry (sm_100: LDG.E.256): a divergent gather costs one request per lane
    EntryLN e;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(e.cf), "=d"(e.mu), "=d"(e.h), "=d"(e.pad) : "l"(p));
    return e;
}
// 32 bytes (8 words) through the read-only path in one request; p must be 32-byte aligned
__device__ __forceinline__ void ldg256_u32(const uint32_t *p, uint32_t (&w)[8]) {
    unsigned long long a, b, c, d;
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32); w[2] = (uint32_t)b; w[3] = (uint32_t)(b >> 32);
    w[4] = (uint32_t)c; w[5] = (uint32_t)(c >> 32); w[6] = (uint32_t)d; w[7] = (uint32_t)(d >> 32);
}
__device__ __forceinline__ EntryEX load_entry(const EntryEX *p) {
    double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    EntryEX e;
    e.wt = a.x; e.theta = a.y;
    return e;
}

// ---------------------------------------------------------------------------------------
// sub-warp groups of G lanes (G = 1,2,4,8,16,32); one group per child event
// ---------------------------------------------------------------------------------------
template <int G> __device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    unsigned lane = threadIdx.x & 31;
    return ((1u << G) - 1u) << (lane / G * G);
}
template <int G> __device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
    for (int d = G / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(mask, v, d, G);
    return v;
}
template <int G> __device__ __forceinline__ double group_incl_scan(double v, unsigned mask, int gl) {
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        double y = __shfl_up_sync(mask, v, d, G);
        if (gl >= d) v += y;
    }
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
// deterministic block reduction of two accumulators; result valid in thread 0
// the same for any CTA size up to 1024 threads (red: [64] smem)
__device__ __forceinline__ void block_sum2_any(double &x, double &y, double *red) {
    x = warp_sum(x);
    y = warp_sum(y);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) { red[w] = x; red[32 + w] = y; }
    __syncthreads();
    if (w == 0) {
        double xs = l < nw ? red[l] : 0.0, ys = l < nw ? red[32 + l] : 0.0;
        x = warp_sum(xs);
        y = warp_sum(ys);
    }
}
__device__ __forceinline__ void block_sum2(double &x, double &y, double *red /* [16] smem */) {
    x = warp_sum(x);
    y = warp_sum(y);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[w] = x; red[8 + w] = y; }
    __syncthreads();
    if (w == 0) {
        double xs = l < (NHP_BLOCK / 32) ? red[l] : 0.0, ys = l < (NHP_BLOCK / 32) ? red[8 + l] : 0.0;
        x = warp_sum(xs);
        y = warp_sum(ys);
    }
}
