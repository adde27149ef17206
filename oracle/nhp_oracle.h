/*
 * nhp_oracle.h -- CPU oracle for the event-history hot path of NetworkHawkesProcesses.jl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference is pure Julia and no Julia toolchain exists in the build
 * container or on the GPU box, and the reference's own tests hold no golden vectors for the
 * hot path.  The oracle is pinned only by (i) the peripheral fixtures of
 * /root/reference/test/baselines.jl (node_counts, DiscreteHomogeneousProcess) and (ii)
 * known-answer vectors re-derived independently from the cited formulas with SciPy
 * (tests/golden/make_kat.py -> tests/golden/kat.json).
 *
 * Conventions (identical to the Julia side so the product C ABI and the oracle take the same
 * buffers): Float64 everywhere; matrices column-major X[parent + K*child] (0-based offsets);
 * nodes are 1-based Int64; parents are 1-based Int64 indices with 0 = baseline.
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src/).
 */
#ifndef NHP_ORACLE_H
#define NHP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EXPONENTIAL 0
#define ORC_LOGITNORMAL 1

/* continuous model parameters (Standard: A == NULL) */
typedef struct {
    int kind;              /* ORC_EXPONENTIAL | ORC_LOGITNORMAL */
    int64_t K;
    const double *lambda0; /* [K]   HomogeneousProcess.lambda            baselines.jl:27-39 */
    const double *W;       /* [K*K] weights.W[parent,child]              weights.jl:47-55   */
    const double *A;       /* [K*K] adjacency_matrix or NULL             continuous.jl:315-321 */
    const double *p1;      /* [K*K] theta (Exponential) or mu (LogitNormal) */
    const double *p2;      /* [K*K] tau (LogitNormal) or NULL */
    double dtmax;          /* impulses.Dtmax (Inf allowed for Exponential) */
} orc_cont_model;

int orc_set_threads(int nthreads); /* 0/1 = serial reference path; >1 = the Threads.@threads sites */
int orc_get_max_threads(void);

/* scalar impulse pdfs */
double orc_exponential_pdf(double theta, double dt);            /* impulses.jl:106-108 */
double orc_logitnormal_pdf(double mu, double tau, double x);    /* impulses.jl:174-178 */
double orc_impulse_response(const orc_cont_model *m, int64_t parentnode, int64_t childnode, double dt); /* continuous.jl:302-305, 521-525 */

/* continuous likelihood / intensity */
double orc_total_intensity(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t index1, double time, int64_t node); /* continuous.jl:286-300, 391-405 */
int orc_cont_event_intensity(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double *out);
int orc_cont_loglik(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double duration, int recursive, double *ll); /* continuous.jl:210-239, 360-389 */
/* extension: analytic gradient of orc_cont_loglik (pinned by finite differences of it) */
int orc_cont_loglik_grad(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double duration, int recursive, double *ll,
                         double *dl0, double *dW, double *dp1, double *dp2);
int orc_cont_recursive_loglik(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double duration, double *ll); /* continuous.jl:241-276, 407-442 */
int orc_cont_intensity(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, const double *times, int64_t nq, double *out /* [nq*K] col-major */); /* continuous.jl:76-96 */

/* Gibbs parents + sufficient statistics */
int orc_cont_resample_parents(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, const double *u, int64_t *parents, int64_t *parentnodes); /* parents.jl:1-46 */
void orc_node_counts(const int64_t *nodes, int64_t n, int64_t K, double *Mn);                                   /* parents.jl:61-68 */
void orc_parent_counts(const int64_t *nodes, const int64_t *parentnodes, int64_t n, int64_t K, double *Mnm);     /* parents.jl:70-79 */
void orc_baseline_counts(const int64_t *nodes, const int64_t *parentnodes, int64_t n, int64_t K, double *M0);    /* baselines.jl:87-96 */
void orc_duration_mean(const double *events, const int64_t *nodes, const int64_t *parents, int64_t n, int64_t K, double *Xnm); /* impulses.jl:84-96 */
void orc_log_duration_sum(const double *events, const int64_t *nodes, const int64_t *parents, int64_t n, int64_t K, double dtmax, double *Xsum); /* impulses.jl:230-240 */
void orc_log_duration_variation(const double *Xbar, const double *events, const int64_t *nodes, const int64_t *parents, int64_t n, int64_t K, double dtmax, double *V); /* impulses.jl:242-252 */

/* adjacency Gibbs: A is mutated in place, u[p + K*c] are the Bernoulli uniforms, rho[p + K*c] link probabilities */
int orc_cont_resample_adjacency(const orc_cont_model *m, double *A_inout, const double *rho, const double *events, const int64_t *nodes, int64_t n, double duration, const double *u); /* continuous.jl:444-519 */
int orc_cont_resample_adjacency_cols(const orc_cont_model *m, double *A_inout, const double *rho, const double *events, const int64_t *nodes, int64_t n, double duration, const double *u,
                                     int64_t col_begin, int64_t col_stride); /* the columns c % col_stride == col_begin of the same sweep (continuous.jl:462-464) */

/* discrete path */
void orc_disc_basis(int64_t L, int64_t B, double dt, double *phi /* [L*B] col-major phi[l + L*b] */);   /* impulses.jl:321-335 */
void orc_disc_convolve(const int64_t *data /* [N*T] data[n + N*t] */, int64_t N, int64_t T, const double *phi, int64_t L, int64_t B, double *conv /* [T*N*B] conv[t + T*(n + N*b)] */); /* discrete.jl:146-151 */
typedef struct {
    int64_t N, B;
    const double *lambda0; /* [N] */
    const double *W;       /* [N*N] */
    const double *A;       /* [N*N] or NULL */
    const double *theta;   /* [N*N*B] theta[p + N*(c + N*b)] */
    double dt;
} orc_disc_model;
void orc_disc_intensity(const orc_disc_model *m, const double *conv, int64_t T, double *lam /* [T*N] lam[t + T*c] */); /* discrete.jl:115-129 */
double orc_disc_loglik(const orc_disc_model *m, const int64_t *data, const double *conv, int64_t T);   /* discrete.jl:91-102 */
/* Gibbs parent counts reduced over t: counts[c + N*k], k = 0 baseline, k = 1 + p*B + b.
 * Multinomial(n, mu) is realised as n categorical inverse-cdf draws from the supplied uniforms
 * u (consumed in (t outer, c inner, draw) order) -- distributionally equal to parents.jl:103-117. */
int orc_disc_gibbs_counts(const orc_disc_model *m, const int64_t *data, const double *conv, int64_t T, const double *u, int64_t nu, double *counts);
/* VB: u[t,c,:] normalised (parents.jl:136-177) reduced on the fly into the statistics that
 * baselines.jl:444-452, weights.jl:70-91, impulses.jl:355-371 compute.  e0[c] = exp(E log lambda0_c),
 * E[p + N*(c + N*b)] = exp(Elog theta + Elog W). */
void orc_disc_vb_stats(int64_t N, int64_t B, int64_t T, const int64_t *data, const double *conv, const double *e0, const double *E,
                       double *alpha_sum /* [N] */, double *kappa_sum /* [N*N] */, double *nu_sum /* [N*N] */, double *gamma_sum /* [N*N*B] */);
int orc_disc_resample_adjacency(const orc_disc_model *m, double *A_inout, const double *rho, const int64_t *data, const double *conv, int64_t T, const double *u); /* discrete.jl:426-480 */

#ifdef __cplusplus
}
#endif
#endif
