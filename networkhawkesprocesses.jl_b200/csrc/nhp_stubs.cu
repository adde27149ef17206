// nhp_stubs.cu -- entry points declared in include/nhp.h whose kernels are not written yet.
// They fail loudly (NHP_ERR_UNSUPPORTED); nothing falls back to a CPU path.
#include "nhp_internal.cuh"

#define NHP_STUB(name, ...)                                                                          \
    extern "C" int name(__VA_ARGS__) { return nhp_fail(ctx, NHP_ERR_UNSUPPORTED, #name ": not implemented in this build"); }

NHP_STUB(nhp_disc_upload, nhp_ctx *ctx, const int64_t *, int64_t, int64_t, int64_t, nhp_disc **)
NHP_STUB(nhp_disc_free, nhp_ctx *ctx, nhp_disc *)
NHP_STUB(nhp_disc_convolve, nhp_ctx *ctx, nhp_disc *, const double *, int64_t, int64_t, double *)
NHP_STUB(nhp_disc_params_set, nhp_ctx *ctx, int64_t, int64_t, const double *, const double *, const double *, const double *, double)
NHP_STUB(nhp_disc_intensity, nhp_ctx *ctx, nhp_disc *, double *)
NHP_STUB(nhp_disc_loglik, nhp_ctx *ctx, nhp_disc *, double *)
NHP_STUB(nhp_disc_gibbs_counts, nhp_ctx *ctx, nhp_disc *, uint64_t, uint64_t, const double *, int64_t, double *)
NHP_STUB(nhp_disc_vb_stats, nhp_ctx *ctx, nhp_disc *, const double *, const double *, double *, double *, double *, double *)
NHP_STUB(nhp_disc_resample_adjacency, nhp_ctx *ctx, nhp_disc *, const double *, uint64_t, uint64_t, const double *, double *)
extern "C" int nhp_disc_basis(int64_t, int64_t, double, double *) { return NHP_ERR_UNSUPPORTED; }
