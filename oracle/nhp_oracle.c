/*
 * nhp_oracle.c -- CPU restatement (plain C, FP64) of the reference's event-history hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see nhp_oracle.h).  PARITY UNPINNED beyond the KAT vectors in
 * tests/golden/kat.json and the peripheral fixtures of /root/reference/test/baselines.jl.
 *
 * Each function follows the reference statement by statement, including the quirks listed
 * in SURVEY.md section 9 (Q3, Q4, Q6, Q7, Q8, Q9).  Third-party arithmetic that is not vendored in
 * /root/reference (Distributions 0.25.76, StatsFuns 1.1.1, LogExpFunctions 0.3.19, DSP 0.7.7;
 * versions from Manifest.toml) is restated from its published definition at the call site.
 *
 * Threading mirrors the reference: OpenMP `parallel for` exactly at the Threads.@threads
 * sites (continuous.jl:225,375,462; parents.jl:8,141; discrete.jl:438), serial elsewhere.
 */
#include "nhp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

int orc_set_threads(int nthreads) { g_threads = nthreads < 1 ? 1 : nthreads; return g_threads; }
int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define INVSQRT2PI 0.3989422804014327 /* StatsFuns.invsqrt2pi */

/* Distributions.pdf(Exponential(1 ./ theta), dt): rate = inv(scale); 0 for dt < 0.  impulses.jl:106-108 */
double orc_exponential_pdf(double theta, double dt) {
    double scale = 1.0 / theta;
    double rate = 1.0 / scale;
    if (dt < 0.0) return 0.0;
    return rate * exp(-rate * dt);
}

/* Distributions.pdf(LogitNormal(mu, tau^(-1/2)), x) = normpdf(mu, sigma, logit(x)) / (x (1-x)) on 0<x<1.
 * impulses.jl:174-178; StatsFuns.normpdf(mu,sigma,z) = exp(-abs2((z-mu)/sigma)/2) * invsqrt2pi / sigma;
 * LogExpFunctions.logit(x) = log(x / (1 - x)).  No 1/dtmax Jacobian (Q8). */
double orc_logitnormal_pdf(double mu, double tau, double x) {
    if (!(0.0 < x && x < 1.0)) return 0.0;
    double sigma = pow(tau, -0.5);
    double lx = log(x / (1.0 - x));
    double z = (lx - mu) / sigma;
    double npdf = exp(-(z * z) / 2.0) * INVSQRT2PI / sigma;
    return npdf / (x * (1.0 - x));
}

/* continuous.jl:302-305 (Standard: w * pdf) and :521-525 (Network: a * w * pdf). 1-based nodes. */
double orc_impulse_response(const orc_cont_model *m, int64_t parentnode, int64_t childnode, double dt) {
    int64_t k = (parentnode - 1) + m->K * (childnode - 1);
    double w = m->W[k];
    double pdf;
    if (m->kind == ORC_EXPONENTIAL) pdf = orc_exponential_pdf(m->p1[k], dt);
    else pdf = orc_logitnormal_pdf(m->p1[k], m->p2[k], dt / m->dtmax);
    if (m->A) return m->A[k] * w * pdf;
    return w * pdf;
}

/* continuous.jl:286-300 / 391-405.  index1 is the 1-based event index. */
double orc_total_intensity(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t index1, double time, int64_t node) {
    double lam = m->lambda0[node - 1]; /* baselines.jl:115-118 */
    if (index1 == 1) return lam;
    int64_t parentindex = index1 - 1; /* 1-based */
    while (events[parentindex - 1] > time - m->dtmax) {
        double parenttime = events[parentindex - 1];
        int64_t parentnode = nodes[parentindex - 1];
        double dt = time - parenttime;
        lam += orc_impulse_response(m, parentnode, node, dt);
        parentindex -= 1;
        if (parentindex == 0) break;
    }
    return lam;
}

int orc_cont_event_intensity(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double *out) {
#pragma omp parallel for schedule(dynamic, 256) if (g_threads > 1) num_threads(g_threads)
    for (int64_t i = 0; i < n; i++) out[i] = orc_total_intensity(m, events, nodes, i + 1, events[i], nodes[i]);
    return 0;
}

/* continuous.jl:241-276 (Standard) / 407-442 (Network).  Q3: the integral term uses W without A;
 * Q6: parenttimes > 0.0 is the "has a previous event" sentinel; Q7: ignores a finite dtmax. */
int orc_cont_recursive_loglik(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double duration, double *llout) {
    int64_t K = m->K;
    double ll = 0.0;
    double s0 = 0.0;
    for (int64_t k = 0; k < K; k++) s0 += m->lambda0[k] * duration; /* baselines.jl:98-102 */
    ll -= s0;
    for (int64_t i = 0; i < n; i++) {
        int64_t p = nodes[i] - 1;
        double rs = 0.0;
        for (int64_t c = 0; c < K; c++) rs += m->W[p + K * c];
        ll -= rs;
    }
    double *partial = (double *)calloc((size_t)(K * K), sizeof(double));
    double *ptimes = (double *)calloc((size_t)K, sizeof(double));
    if (!partial || !ptimes) { free(partial); free(ptimes); return -2; }
    const double *theta = m->p1;
    for (int64_t i = 0; i < n; i++) {
        double childtime = events[i];
        int64_t c = nodes[i] - 1;
        double lam = m->lambda0[c];
        if (ptimes[c] > 0.0) {
            double dt = childtime - ptimes[c];
            for (int64_t j = 0; j < K; j++) { /* row partialsums[childnode, :] */
                double next = exp(-dt * theta[c + K * j]);
                partial[c + K * j] = next * (1.0 + partial[c + K * j]);
            }
        }
        for (int64_t p = 0; p < K; p++) {
            double parenttime = ptimes[p];
            if (parenttime > 0.0) {
                double r;
                if (p == c) r = partial[p + K * c];
                else {
                    double dt = childtime - parenttime;
                    double next = exp(-dt * theta[p + K * c]);
                    r = next * (1.0 + partial[p + K * c]);
                }
                double w = m->A ? m->A[p + K * c] * m->W[p + K * c] : m->W[p + K * c];
                lam += w * theta[p + K * c] * r;
            }
        }
        ll += log(lam);
        ptimes[c] = childtime;
    }
    free(partial); free(ptimes);
    *llout = ll;
    return 0;
}

/* continuous.jl:210-239 / 360-389 */
int orc_cont_loglik(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double duration, int recursive, double *llout) {
    if (m->kind == ORC_EXPONENTIAL && recursive) return orc_cont_recursive_loglik(m, events, nodes, n, duration, llout);
    int64_t K = m->K;
    double ll = 0.0;
    double s0 = 0.0;
    for (int64_t k = 0; k < K; k++) s0 += m->lambda0[k] * duration;
    ll -= s0;
    for (int64_t i = 0; i < n; i++) {
        int64_t p = nodes[i] - 1;
        double rs = 0.0;
        for (int64_t c = 0; c < K; c++) rs += m->A ? m->A[p + K * c] * m->W[p + K * c] : m->W[p + K * c];
        ll -= rs;
    }
    if (g_threads > 1) {
        double acc = 0.0; /* Threads.Atomic{Float64} accumulation: order-free sum */
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : acc) num_threads(g_threads)
        for (int64_t i = 0; i < n; i++) acc += log(orc_total_intensity(m, events, nodes, i + 1, events[i], nodes[i]));
        ll += acc;
    } else {
        for (int64_t i = 0; i < n; i++) ll += log(orc_total_intensity(m, events, nodes, i + 1, events[i], nodes[i]));
    }
    *llout = ll;
    return 0;
}

/* EXTENSION (no counterpart in the reference, which hands Optim a gradient-free objective, continuous.jl:144-198):
 * analytic gradient of orc_cont_loglik.  Plain O(n * window) double loop in the reference's window order; for the
 * recursive Exponential path the history is every earlier event with time > 0.0 (quirks Q6/Q7) and the compensator
 * ignores A (Q3).  Pinned by central finite differences of orc_cont_loglik in tests/test_oracle.py. */
int orc_cont_loglik_grad(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, double duration, int recursive, double *llout,
                         double *dl0, double *dW, double *dp1, double *dp2) {
    int64_t K = m->K;
    int rec = (m->kind == ORC_EXPONENTIAL && recursive);
    int rc = orc_cont_loglik(m, events, nodes, n, duration, recursive, llout);
    if (rc) return rc;
    for (int64_t k = 0; k < K; k++) dl0[k] = -duration;
    for (int64_t k = 0; k < K * K; k++) { dW[k] = 0.0; dp1[k] = 0.0; if (dp2) dp2[k] = 0.0; }
    for (int64_t i = 0; i < n; i++) { /* compensator: -sum_i rowsum(c_i) */
        int64_t p = nodes[i] - 1;
        for (int64_t c = 0; c < K; c++) dW[p + K * c] -= (m->A && !rec) ? m->A[p + K * c] : 1.0;
    }
    for (int64_t i = 0; i < n; i++) {
        int64_t c = nodes[i] - 1;
        double ti = events[i];
        /* pass 1: lambda_i */
        double lam = m->lambda0[c];
        int64_t jstop = i; /* exclusive lower end of the window, found in pass 1 */
        for (int64_t j = i - 1; j >= 0; j--) {
            if (rec) { if (!(events[j] > 0.0)) { jstop = j + 1; break; } }
            else if (!(events[j] > ti - m->dtmax)) { jstop = j + 1; break; }
            lam += orc_impulse_response(m, nodes[j], c + 1, ti - events[j]);
            jstop = j;
        }
        double inv = 1.0 / lam;
        dl0[c] += inv;
        /* pass 2: per-pair terms */
        for (int64_t j = i - 1; j >= jstop; j--) {
            int64_t p = nodes[j] - 1, k = p + K * c;
            double dt = ti - events[j];
            double a = m->A ? m->A[k] : 1.0, w = m->W[k];
            if (m->kind == ORC_EXPONENTIAL) {
                double th = m->p1[k];
                double ex = dt < 0.0 ? 0.0 : exp(-th * dt);
                dW[k] += a * th * ex * inv;
                dp1[k] += a * w * ex * (1.0 - th * dt) * inv;
            } else {
                double x = dt / m->dtmax;
                if (!(0.0 < x && x < 1.0)) continue;
                double mu = m->p1[k], tau = m->p2[k];
                double pdf = orc_logitnormal_pdf(mu, tau, x);
                double z = log(x / (1.0 - x));
                dW[k] += a * pdf * inv;
                dp1[k] += a * w * pdf * tau * (z - mu) * inv;
                if (dp2) dp2[k] += a * w * pdf * (0.5 / tau - 0.5 * (z - mu) * (z - mu)) * inv;
            }
        }
    }
    return 0;
}

/* continuous.jl:76-96: strict window  time - dtmax < events < time, all K children. */
int orc_cont_intensity(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, const double *times, int64_t nq, double *out) {
    int64_t K = m->K;
    for (int64_t q = 0; q < nq; q++) {
        double t0 = times[q];
        for (int64_t c = 0; c < K; c++) {
            double lam = 0.0;
            for (int64_t j = 0; j < n; j++) {
                if (t0 - m->dtmax < events[j] && events[j] < t0) lam += orc_impulse_response(m, nodes[j], c + 1, t0 - events[j]);
            }
            out[q + nq * c] = m->lambda0[c] + lam;
        }
    }
    return 0;
}

/* Base.sum over a Vector: sequential below 16 elements, otherwise mapreduce_impl pairwise with
 * block size 1024 (Julia 1.8 base/reduce.jl). */
static double julia_sum_impl(const double *a, int64_t ifirst, int64_t ilast) {
    if (ifirst == ilast) return a[ifirst];
    if (ilast - ifirst < 1024) {
        double v = a[ifirst] + a[ifirst + 1];
        for (int64_t i = ifirst + 2; i <= ilast; i++) v += a[i];
        return v;
    }
    int64_t imid = ifirst + ((ilast - ifirst) >> 1);
    return julia_sum_impl(a, ifirst, imid) + julia_sum_impl(a, imid + 1, ilast);
}
static double julia_sum(const double *a, int64_t n) {
    if (n == 0) return 0.0;
    if (n == 1) return a[0];
    return julia_sum_impl(a, 0, n - 1);
}

/* Distributions rand(Categorical(p)): one uniform, cp = p[1]; while cp <= u && i < n: cp += p[++i]. */
static int64_t categorical_draw(const double *p, int64_t n, double u) {
    double cp = p[0];
    int64_t i = 0;
    while (cp <= u && i < n - 1) { i++; cp += p[i]; }
    return i;
}

/* parents.jl:25-46.  Returns -1 when Categorical would reject the probability vector. */
static int resample_parent(const orc_cont_model *m, double event, int64_t node, int64_t index1, const double *events, const int64_t *nodes,
                           double u, double **buf, int64_t *cap, int64_t *parent, int64_t *parentnode) {
    if (index1 == 1) { *parent = 0; *parentnode = 0; return 0; }
    int64_t cnt = 0;
    int64_t parentindex = index1 - 1;
    while (events[parentindex - 1] > event - m->dtmax) {
        if (cnt + 2 > *cap) {
            *cap = (*cap) * 2 + 16;
            *buf = (double *)realloc(*buf, (size_t)(*cap) * sizeof(double));
            if (!*buf) return -2;
        }
        (*buf)[cnt++] = orc_impulse_response(m, nodes[parentindex - 1], node, event - events[parentindex - 1]);
        parentindex -= 1;
        if (parentindex == 0) break;
    }
    if (cnt + 2 > *cap) {
        *cap = (*cap) * 2 + 16;
        *buf = (double *)realloc(*buf, (size_t)(*cap) * sizeof(double));
        if (!*buf) return -2;
    }
    (*buf)[cnt++] = m->lambda0[node - 1]; /* baseline LAST */
    double s = julia_sum(*buf, cnt);
    double tot = 0.0;
    for (int64_t k = 0; k < cnt; k++) {
        (*buf)[k] /= s;
        if (!((*buf)[k] >= 0.0)) return -1;
        tot += (*buf)[k];
    }
    if (!(fabs(tot - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(tot), 1.0))) return -1; /* isprobvec */
    int64_t pick = categorical_draw(*buf, cnt, u);
    if (pick == cnt - 1) { *parent = 0; *parentnode = 0; }
    else { *parent = (index1 - 1) - pick; *parentnode = nodes[*parent - 1]; }
    return 0;
}

/* parents.jl:1-23; u[i] is the uniform consumed by event i (event 1 consumes none). */
int orc_cont_resample_parents(const orc_cont_model *m, const double *events, const int64_t *nodes, int64_t n, const double *u, int64_t *parents, int64_t *parentnodes) {
    int rc = 0;
#pragma omp parallel if (g_threads > 1) num_threads(g_threads)
    {
        double *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; i++) {
            int r = resample_parent(m, events[i], nodes[i], i + 1, events, nodes, u[i], &buf, &cap, &parents[i], &parentnodes[i]);
            if (r != 0) {
#pragma omp atomic write
                rc = r;
            }
        }
        free(buf);
    }
    return rc;
}

void orc_node_counts(const int64_t *nodes, int64_t n, int64_t K, double *Mn) {
    memset(Mn, 0, (size_t)K * sizeof(double));
    for (int64_t i = 0; i < n; i++) Mn[nodes[i] - 1] += 1.0;
}

void orc_parent_counts(const int64_t *nodes, const int64_t *parentnodes, int64_t n, int64_t K, double *Mnm) {
    memset(Mnm, 0, (size_t)(K * K) * sizeof(double));
    for (int64_t i = 0; i < n; i++)
        if (parentnodes[i] > 0) Mnm[(parentnodes[i] - 1) + K * (nodes[i] - 1)] += 1.0;
}

void orc_baseline_counts(const int64_t *nodes, const int64_t *parentnodes, int64_t n, int64_t K, double *M0) {
    memset(M0, 0, (size_t)K * sizeof(double));
    for (int64_t i = 0; i < n; i++)
        if (parentnodes[i] == 0) M0[nodes[i] - 1] += 1.0;
}

/* impulses.jl:84-96; fillna!(Xnm ./ Mnm, 0) (helpers.jl:18-25) */
void orc_duration_mean(const double *events, const int64_t *nodes, const int64_t *parents, int64_t n, int64_t K, double *Xnm) {
    double *M = (double *)calloc((size_t)(K * K), sizeof(double));
    memset(Xnm, 0, (size_t)(K * K) * sizeof(double));
    for (int64_t i = 0; i < n; i++) {
        int64_t par = parents[i];
        if (par > 0) {
            int64_t k = (nodes[par - 1] - 1) + K * (nodes[i] - 1);
            M[k] += 1.0;
            Xnm[k] += events[i] - events[par - 1];
        }
    }
    for (int64_t k = 0; k < K * K; k++) {
        double v = Xnm[k] / M[k];
        Xnm[k] = isnan(v) ? 0.0 : v;
    }
    free(M);
}

/* impulses.jl:228 */
static double log_duration(double parent, double child, double dtmax) { return log((child - parent) / (dtmax - (child - parent))); }

void orc_log_duration_sum(const double *events, const int64_t *nodes, const int64_t *parents, int64_t n, int64_t K, double dtmax, double *Xsum) {
    memset(Xsum, 0, (size_t)(K * K) * sizeof(double));
    for (int64_t i = 0; i < n; i++) {
        int64_t par = parents[i];
        if (par > 0) Xsum[(nodes[par - 1] - 1) + K * (nodes[i] - 1)] += log_duration(events[par - 1], events[i], dtmax);
    }
}

void orc_log_duration_variation(const double *Xbar, const double *events, const int64_t *nodes, const int64_t *parents, int64_t n, int64_t K, double dtmax, double *V) {
    memset(V, 0, (size_t)(K * K) * sizeof(double));
    for (int64_t i = 0; i < n; i++) {
        int64_t par = parents[i];
        if (par > 0) {
            int64_t k = (nodes[par - 1] - 1) + K * (nodes[i] - 1);
            double d = log_duration(events[par - 1], events[i], dtmax) - Xbar[k];
            V[k] += d * d;
        }
    }
}

/* helpers.jl:13-16 */
static double logsumexp2(double a, double b) {
    double mx = a > b ? a : b;
    return mx + log(exp(a - mx) + exp(b - mx));
}

/* continuous.jl:489-498 with A column taken from `A` */
static double adj_integrated_intensity(const orc_cont_model *m, const double *A, int64_t node0, const double *counts, double duration) {
    int64_t K = m->K;
    double I = m->lambda0[node0] * duration;
    for (int64_t p = 0; p < K; p++) I += A[p + K * node0] * m->W[p + K * node0] * counts[p];
    return I;
}

/* continuous.jl:500-519 (Q4: the first event never contributes its log term) */
static double adj_sum_log_intensity(const orc_cont_model *m, int64_t node1, const double *events, const int64_t *nodes, int64_t n) {
    double S = 0.0;
    for (int64_t index = 1; index <= n; index++) {
        if (nodes[index - 1] != node1) continue;
        if (index == 1) continue;
        S += log(orc_total_intensity(m, events, nodes, index, events[index - 1], node1));
    }
    return S;
}

/* continuous.jl:444-487.  Bernoulli draw: rand() <= p (Distributions). */
/* columns c with c % col_stride == col_begin only (continuous.jl:462-464: columns are independent, so a subset is the
 * same computation restricted to those columns; the full sweep is col_begin = 0, col_stride = 1) */
int orc_cont_resample_adjacency_cols(const orc_cont_model *m0, double *A, const double *rho, const double *events, const int64_t *nodes, int64_t n, double duration, const double *u,
                                     int64_t col_begin, int64_t col_stride) {
    int64_t K = m0->K;
    double *counts = (double *)malloc((size_t)K * sizeof(double));
    orc_node_counts(nodes, n, K, counts);
    orc_cont_model m = *m0;
    m.A = A;
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1) num_threads(g_threads)
    for (int64_t c = col_begin; c < K; c += col_stride) {
        for (int64_t p = 0; p < K; p++) {
            int64_t k = p + K * c;
            A[k] = 0.0;
            double ll0 = -adj_integrated_intensity(&m, A, c, counts, duration);
            ll0 += adj_sum_log_intensity(&m, c + 1, events, nodes, n);
            ll0 += log(1.0 - rho[k]);
            A[k] = 1.0;
            double ll1 = -adj_integrated_intensity(&m, A, c, counts, duration);
            ll1 += adj_sum_log_intensity(&m, c + 1, events, nodes, n);
            ll1 += log(rho[k]);
            double Z = logsumexp2(ll0, ll1);
            A[k] = (u[k] <= exp(ll1 - Z)) ? 1.0 : 0.0;
        }
    }
    free(counts);
    return 0;
}
int orc_cont_resample_adjacency(const orc_cont_model *m0, double *A, const double *rho, const double *events, const int64_t *nodes, int64_t n, double duration, const double *u) {
    return orc_cont_resample_adjacency_cols(m0, A, rho, events, nodes, n, duration, u, 0, 1);
}

/* ------------------------------------------------------------------ discrete path */

/* impulses.jl:321-335.  sigma = L/(B-1); means LinRange(1,L,B+2)[2:end-1] if B<L else LinRange(1,L,B);
 * phi = exp(-1/2 * sigma^-1 / 2 * (lag - mu)^2) (i.e. exp(-(l-mu)^2/(4 sigma))), each column / (sum * dt). */
static double linrange(double a, double b, int64_t len, int64_t j /* 0-based */) {
    if (len == 1) return a;
    double t = (double)j / (double)(len - 1);
    return (1.0 - t) * a + t * b;
}
void orc_disc_basis(int64_t L, int64_t B, double dt, double *phi) {
    double sigma = (double)L / (double)(B - 1);
    for (int64_t b = 0; b < B; b++) {
        double mu = (B < L) ? linrange(1.0, (double)L, B + 2, b + 1) : linrange(1.0, (double)L, B, b);
        double s = 0.0;
        for (int64_t l = 0; l < L; l++) {
            double d = (double)(l + 1) - mu;
            double v = exp(-1.0 / 2.0 * (1.0 / sigma) / 2.0 * (d * d));
            phi[l + L * b] = v;
            s += v;
        }
        for (int64_t l = 0; l < L; l++) phi[l + L * b] = phi[l + L * b] / (s * dt);
    }
}

/* discrete.jl:146-151: conv(data', [0; phi_b])[1:T, :] restated as the exact causal FIR
 * conv[t,n,b] = sum_{l=1..L, t-l>=1} phi_b[l] * data[n, t-l]; then max(., 0). (DSP.conv is FFT
 * based: agreement ~1e-13 relative.) */
void orc_disc_convolve(const int64_t *data, int64_t N, int64_t T, const double *phi, int64_t L, int64_t B, double *conv) {
    for (int64_t b = 0; b < B; b++)
        for (int64_t n = 0; n < N; n++)
            for (int64_t t = 0; t < T; t++) {
                double s = 0.0;
                for (int64_t l = 1; l <= L; l++) {
                    if (t - l < 0) break;
                    s += phi[(l - 1) + L * b] * (double)data[n + N * (t - l)];
                }
                conv[t + T * (n + N * b)] = s > 0.0 ? s : 0.0;
            }
}

/* discrete.jl:381-385 / 511-516 */
static double bump(const orc_disc_model *m, int64_t p, int64_t c, int64_t b) {
    int64_t N = m->N;
    double w = m->W[p + N * c];
    double th = m->theta[p + N * (c + N * b)];
    if (m->A) return m->A[p + N * c] * w * th * m->dt;
    return w * th * m->dt;
}

/* discrete.jl:115-129; baseline rows lambda0 * dt (baselines.jl:402-405) */
void orc_disc_intensity(const orc_disc_model *m, const double *conv, int64_t T, double *lam) {
    int64_t N = m->N, B = m->B;
    for (int64_t t = 0; t < T; t++)
        for (int64_t c = 0; c < N; c++) {
            double v = m->lambda0[c] * m->dt;
            for (int64_t p = 0; p < N; p++)
                for (int64_t b = 0; b < B; b++) v += conv[t + T * (p + N * b)] * bump(m, p, c, b);
            lam[t + T * c] = v;
        }
}

/* StatsFuns.poislogpdf(lambda, x) = xlogy(x, lambda) - lambda - loggamma(x+1);
 * the reference takes log(pdf(.)) = log(exp(logpdf)) (discrete.jl:98). */
static double log_poisson_pdf(double lambda, int64_t s) {
    double xl = (s == 0) ? 0.0 : (double)s * log(lambda);
    double lp = xl - lambda - lgamma((double)s + 1.0);
    return log(exp(lp));
}

double orc_disc_loglik(const orc_disc_model *m, const int64_t *data, const double *conv, int64_t T) {
    int64_t N = m->N;
    double *lam = (double *)malloc((size_t)(T * N) * sizeof(double));
    orc_disc_intensity(m, conv, T, lam);
    double ll = 0.0;
    for (int64_t t = 0; t < T; t++)
        for (int64_t n = 0; n < N; n++) ll += log_poisson_pdf(lam[t + T * n], data[n + N * t]);
    free(lam);
    return ll;
}

/* parents.jl:82-117 reduced over t (the only form downstream code uses: impulses.jl:341,
 * parents.jl:130, baselines.jl:414). */
int orc_disc_gibbs_counts(const orc_disc_model *m, const int64_t *data, const double *conv, int64_t T, const double *u, int64_t nu, double *counts) {
    int64_t N = m->N, B = m->B, NK = 1 + N * B;
    memset(counts, 0, (size_t)(N * NK) * sizeof(double));
    double *mu = (double *)malloc((size_t)NK * sizeof(double));
    int64_t iu = 0;
    for (int64_t t = 0; t < T; t++)
        for (int64_t c = 0; c < N; c++) {
            int64_t s = data[c + N * t];
            if (s == 0) continue; /* Multinomial(0, mu) is the zero vector */
            mu[0] = m->lambda0[c] * m->dt;
            double tot = mu[0];
            for (int64_t p = 0; p < N; p++)
                for (int64_t b = 0; b < B; b++) {
                    double v = conv[t + T * (p + N * b)] * bump(m, p, c, b);
                    mu[1 + p * B + b] = v;
                    tot += v;
                }
            for (int64_t k = 0; k < NK; k++) mu[k] /= tot;
            for (int64_t d = 0; d < s; d++) {
                if (iu >= nu) { free(mu); return -1; }
                int64_t k = categorical_draw(mu, NK, u[iu++]);
                counts[c + N * k] += 1.0;
            }
        }
    free(mu);
    return 0;
}

/* parents.jl:136-177 + baselines.jl:444-452 + weights.jl:70-91 + impulses.jl:355-371 */
void orc_disc_vb_stats(int64_t N, int64_t B, int64_t T, const int64_t *data, const double *conv, const double *e0, const double *E,
                       double *alpha_sum, double *kappa_sum, double *nu_sum, double *gamma_sum) {
    int64_t NK = 1 + N * B;
    memset(alpha_sum, 0, (size_t)N * sizeof(double));
    memset(kappa_sum, 0, (size_t)(N * N) * sizeof(double));
    memset(nu_sum, 0, (size_t)(N * N) * sizeof(double));
    memset(gamma_sum, 0, (size_t)(N * N * B) * sizeof(double));
    double *uu = (double *)malloc((size_t)NK * sizeof(double));
    for (int64_t t = 0; t < T; t++)
        for (int64_t c = 0; c < N; c++) {
            uu[0] = e0[c];
            double Z = uu[0];
            for (int64_t p = 0; p < N; p++)
                for (int64_t b = 0; b < B; b++) {
                    double v = conv[t + T * (p + N * b)] * E[p + N * (c + N * b)];
                    uu[1 + p * B + b] = v;
                    Z += v;
                }
            double sc = (double)data[c + N * t];
            alpha_sum[c] += (uu[0] / Z) * sc;
            for (int64_t p = 0; p < N; p++) {
                for (int64_t b = 0; b < B; b++) {
                    double un = uu[1 + p * B + b] / Z;
                    kappa_sum[p + N * c] += sc * un;
                    gamma_sum[p + N * (c + N * b)] += sc * un;
                }
                nu_sum[p + N * c] += (double)data[p + N * t];
            }
        }
    free(uu);
}

/* discrete.jl:462-480 */
static double disc_conditional_loglik(const orc_disc_model *m, const double *A, const int64_t *data, const double *conv, int64_t T, double value, int64_t pidx, int64_t cidx) {
    int64_t N = m->N, B = m->B;
    double ll = 0.0;
    for (int64_t t = 0; t < T; t++) {
        double lam = m->lambda0[cidx] * m->dt;
        for (int64_t p = 0; p < N; p++) {
            double w = m->W[p + N * cidx];
            double a = (p == pidx) ? value : A[p + N * cidx];
            for (int64_t b = 0; b < B; b++) lam += conv[t + T * (p + N * b)] * a * w * m->theta[p + N * (cidx + N * b)] * m->dt;
        }
        ll += log_poisson_pdf(lam, data[cidx + N * t]);
    }
    return ll;
}

/* discrete.jl:426-460 */
int orc_disc_resample_adjacency(const orc_disc_model *m, double *A, const double *rho, const int64_t *data, const double *conv, int64_t T, const double *u) {
    int64_t N = m->N;
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1) num_threads(g_threads)
    for (int64_t c = 0; c < N; c++)
        for (int64_t p = 0; p < N; p++) {
            int64_t k = p + N * c;
            double ll0 = disc_conditional_loglik(m, A, data, conv, T, 0.0, p, c) + log(1.0 - rho[k]);
            double ll1 = disc_conditional_loglik(m, A, data, conv, T, 1.0, p, c) + log(rho[k]);
            double lZ = logsumexp2(ll0, ll1);
            A[k] = (u[k] <= exp(ll1 - lZ)) ? 1.0 : 0.0;
        }
    return 0;
}
