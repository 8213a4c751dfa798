/* mpc_osqp.c -- plain-C CPU restatement of the reference's per-tick MPC path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product path may link or load this file: it is a checker and the CPU baseline
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and --impl reference legs).
 *
 * It restates, tick by tick, what /root/reference/MPC.py + the `osqp` package do on the CPU:
 *     MPC.__init__                 MPC.py:22-82      constants
 *     construct_gait / construct_S MPC.py:635-652, 611-633
 *     create_ML / update_ML        MPC.py:98-190, 316-360   sparse constraint matrix, fixed CSC pattern
 *     create_NK / update_NK        MPC.py:192-234, 362-378  bounds
 *     create_weight_matrices       MPC.py:236-288    diagonal P, q = 0
 *     call_solver                  MPC.py:380-430    shifted warm start, setup / update / warm_start / solve
 *     retrieve_result              MPC.py:432-458
 * and the OSQP algorithm behind the five calls MPC.py makes (MPC.py:73, 414-420, 427-428).  `osqp` is a
 * third-party package that is not vendored in the reference, is version-unpinned there (README:
 * `pip3 install --user osqp`; era => 0.6.x) and cannot be installed offline, so its PUBLISHED algorithm is
 * restated (Stellato, Banjac, Goulart, Bemporad, Boyd, Math. Prog. Comp. 2020; 0.6 defaults): modified Ruiz
 * equilibration (10 passes) + cost scaling, quasi-definite KKT system factorised by a sparse LDL' (up-looking,
 * elimination tree; Davis, ACM TOMS 2005, which is what OSQP's QDLDL implements) under a minimum-degree
 * ordering, sigma = 1e-6, rho = 0.1 (1e3 rho on equality rows), alpha = 1.6, termination every 25 iterations
 * on unscaled infinity-norm residuals, adaptive rho on a fixed interval, polish off.
 * It is the same algorithm, setting for setting, as oracle/osqp_port.py (the numpy/scipy restatement);
 * tests/test_oracle_c.py holds the two against each other and against tests/golden/.
 *
 * PARITY STATUS: build half pinned against the reference's own ML.data / NK / NK_inf (tests/golden, produced
 * by importing the unmodified MPC.py); solve half *unpinned by the reference* (no goldens, no osqp here) and
 * certified instead by KKT residuals (oracle/kkt.py): the QP is strictly convex, the optimum unique.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_TOL 1e-4
#define RHO_EQ_OVER_INEQ 1e3

/* ------------------------------------------------------------------------------------------------ constants */
typedef struct {
    int N;
    double dt, mass, mu, fz_max, gravity, w_force;
    double gI[9];
    double footholds[12];   /* 3 x 4 row major */
    double w_state[12];
} mpc_params;

static void params_default(mpc_params* p, int N, double dt) {
    static const double gI[9] = {3.09249e-2, -8.00101e-7, 1.865287e-5, -8.00101e-7, 5.106100e-2, 1.245813e-4,
                                 1.865287e-5, 1.245813e-4, 6.939757e-2};                 /* MPC.py:35-37 */
    static const double fh[12] = {0.19, 0.19, -0.19, -0.19, 0.15005, -0.15005, 0.15005, -0.15005, 0, 0, 0, 0};  /* :67-70 */
    p->N = N; p->dt = dt;
    p->mass = 2.50000279;       /* MPC.py:28 */
    p->mu = 0.9;                /* MPC.py:39 */
    p->fz_max = 25.0;           /* MPC.py:228 */
    p->gravity = 9.81;          /* MPC.py:201 */
    p->w_force = 1e-5;          /* MPC.py:282-284 */
    memcpy(p->gI, gI, sizeof gI);
    memcpy(p->footholds, fh, sizeof fh);
    double* w = p->w_state;     /* MPC.py:255-275 */
    w[0] = 0.1; w[1] = 0.1; w[2] = 1.0; w[3] = w[4] = w[5] = 0.11;
    for (int i = 0; i < 3; ++i) { w[6 + i] = 2.0 * sqrt(w[i]); w[9 + i] = 0.05 * sqrt(w[3 + i]); }
}

/* ------------------------------------------------------------------------------------------------ sparse LDL' */
typedef struct {
    int n;
    int *Up, *Ui;           /* permuted upper triangle, CSC, row indices sorted per column not required */
    double* Ux;
    int *Lp, *Li, *parent, *lnz, *flag, *pattern;
    double *Lx, *D, *Dinv, *Y;
    int* perm;              /* perm[new] = old */
    int* iperm;             /* iperm[old] = new */
    double* work;
} ldl_t;

/* minimum-degree ordering on the graph of a symmetric matrix given by its full adjacency (dense bitmap; n ~ 1e3) */
static void min_degree_order(int n, unsigned char* adj, int* perm) {
    int* deg = (int*)malloc(sizeof(int) * n);
    unsigned char* gone = (unsigned char*)calloc(n, 1);
    int* nb = (int*)malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) {
        int d = 0;
        for (int j = 0; j < n; ++j) d += adj[(size_t)i * n + j] && j != i;
        deg[i] = d;
    }
    for (int step = 0; step < n; ++step) {
        int best = -1;
        for (int i = 0; i < n; ++i)
            if (!gone[i] && (best < 0 || deg[i] < deg[best])) best = i;
        perm[step] = best;
        gone[best] = 1;
        int cnt = 0;
        for (int j = 0; j < n; ++j)
            if (!gone[j] && adj[(size_t)best * n + j]) nb[cnt++] = j;
        for (int a = 0; a < cnt; ++a) {
            unsigned char* ra = adj + (size_t)nb[a] * n;
            ra[best] = 0;
            for (int b = 0; b < cnt; ++b)
                if (a != b) ra[nb[b]] = 1;
        }
        for (int a = 0; a < cnt; ++a) {
            const unsigned char* ra = adj + (size_t)nb[a] * n;
            int d = 0;
            for (int j = 0; j < n; ++j) d += ra[j] && !gone[j] && j != nb[a];
            deg[nb[a]] = d;
        }
    }
    free(deg); free(gone); free(nb);
}

/* symbolic phase: elimination tree and column counts of L for the upper triangle (Up, Ui) */
static void ldl_symbolic(ldl_t* F) {
    const int n = F->n;
    for (int k = 0; k < n; ++k) {
        F->parent[k] = -1; F->flag[k] = k; F->lnz[k] = 0;
        for (int p = F->Up[k]; p < F->Up[k + 1]; ++p) {
            int i = F->Ui[p];
            if (i >= k) continue;
            for (; F->flag[i] != k; i = F->parent[i]) {
                if (F->parent[i] == -1) F->parent[i] = k;
                F->lnz[i]++;
                F->flag[i] = k;
            }
        }
    }
    F->Lp[0] = 0;
    for (int k = 0; k < n; ++k) F->Lp[k + 1] = F->Lp[k] + F->lnz[k];
}

/* numeric phase, up-looking: row k of L from a sparse triangular solve along the elimination tree */
static int ldl_numeric(ldl_t* F) {
    const int n = F->n;
    double* Y = F->Y;
    for (int k = 0; k < n; ++k) {
        Y[k] = 0.0;
        int top = n;
        F->flag[k] = k;
        F->lnz[k] = 0;
        for (int p = F->Up[k]; p < F->Up[k + 1]; ++p) {
            int i = F->Ui[p];
            if (i > k) continue;
            Y[i] += F->Ux[p];
            int len = 0;
            for (; F->flag[i] != k; i = F->parent[i]) { F->pattern[len++] = i; F->flag[i] = k; }
            while (len > 0) F->pattern[--top] = F->pattern[--len];
        }
        double dk = Y[k];
        Y[k] = 0.0;
        for (; top < n; ++top) {
            const int i = F->pattern[top];
            const double yi = Y[i];
            Y[i] = 0.0;
            const int p2 = F->Lp[i] + F->lnz[i];
            for (int p = F->Lp[i]; p < p2; ++p) Y[F->Li[p]] -= F->Lx[p] * yi;
            const double lki = yi * F->Dinv[i];
            dk -= lki * yi;
            F->Li[p2] = k;
            F->Lx[p2] = lki;
            F->lnz[i]++;
        }
        if (dk == 0.0) return -1;
        F->D[k] = dk;
        F->Dinv[k] = 1.0 / dk;
    }
    return 0;
}

/* x <- K^-1 b (b in original ordering) */
static void ldl_solve(const ldl_t* F, const double* b, double* x) {
    const int n = F->n;
    double* w = F->work;
    for (int i = 0; i < n; ++i) w[i] = b[F->perm[i]];
    for (int j = 0; j < n; ++j) {
        const double wj = w[j];
        for (int p = F->Lp[j]; p < F->Lp[j + 1]; ++p) w[F->Li[p]] -= F->Lx[p] * wj;
    }
    for (int j = 0; j < n; ++j) w[j] *= F->Dinv[j];
    for (int j = n - 1; j >= 0; --j) {
        double s = w[j];
        for (int p = F->Lp[j]; p < F->Lp[j + 1]; ++p) s -= F->Lx[p] * w[F->Li[p]];
        w[j] = s;
    }
    for (int i = 0; i < n; ++i) x[F->perm[i]] = w[i];
}

static void ldl_free(ldl_t* F) {
    free(F->Up); free(F->Ui); free(F->Ux); free(F->Lp); free(F->Li); free(F->parent); free(F->lnz); free(F->flag);
    free(F->pattern); free(F->Lx); free(F->D); free(F->Dinv); free(F->Y); free(F->perm); free(F->iperm); free(F->work);
    memset(F, 0, sizeof *F);
}

/* ------------------------------------------------------------------------------------------------ OSQP restatement */
typedef struct {
    int n, m, nnz;
    int *Ap, *Ai;               /* CSC pattern of A (fixed) */
    double *A0, *A;             /* unscaled / scaled values */
    double *P0, *P;             /* diagonal P */
    double *q0, *q, *l0, *u0, *l, *u;
    double *D, *E, c;
    double *x, *z, *y;          /* scaled iterates */
    double rho, *rho_vec;
    int* kind;
    int have_kind;
    ldl_t F;
    int *posA;                  /* A entry -> position in the permuted upper-triangular KKT */
    int *posDx, *posDz;         /* diagonal positions */
    double *rhs, *sol, *t_n, *t_m, *t_n2, *t_m2;
    /* settings */
    double sigma, alpha, eps_abs, eps_rel, rho0, adaptive_tol;
    int scaling, max_iter, check_termination, adaptive_rho, adaptive_interval;
    /* info of the last solve */
    int iter, status, rho_updates, factorizations;
    double pri_res, dua_res;
} osqp_t;

static double limit_scaling(double v) { v = v < MIN_SCALING ? 1.0 : v; return v > MAX_SCALING ? MAX_SCALING : v; }
/* 0 (default): OSQP's cost scaling as published -- ||q||_inf below 1e-4 counts as 1, so with the reference's q = 0 (MPC.py:286-288)
 * the cost scale stays c = 1 and ADMM needs ~1000 iterations at eps 1e-8.  1: a zero q is left out of the max, c = 1 / mean column
 * norm of P (~6e3 here) and ~50 iterations suffice -- what a cost-scaled OSQP WOULD do, reported next to the faithful arm. */
static int g_ignore_zero_q = 0;
void mpc_oracle_set_cost_scaling_variant(int ignore_zero_q) { g_ignore_zero_q = ignore_zero_q; }
static double scale_bound(double b, double e) { return fabs(b) < OSQP_INFTY ? e * b : b; }

static void osqp_scale(osqp_t* s) {
    const int n = s->n, m = s->m;
    memcpy(s->A, s->A0, sizeof(double) * s->nnz);
    memcpy(s->P, s->P0, sizeof(double) * n);
    memcpy(s->q, s->q0, sizeof(double) * n);
    for (int j = 0; j < n; ++j) s->D[j] = 1.0;
    for (int i = 0; i < m; ++i) s->E[i] = 1.0;
    s->c = 1.0;
    double* d = s->t_n; double* e = s->t_m;
    for (int pass = 0; pass < s->scaling; ++pass) {
        /* infinity norms of the columns of [[P, A'], [A, 0]] */
        for (int i = 0; i < m; ++i) e[i] = 0.0;
        for (int j = 0; j < n; ++j) {
            double cn = fabs(s->P[j]);
            for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) {
                const double a = fabs(s->A[p]);
                if (a > cn) cn = a;
                if (a > e[s->Ai[p]]) e[s->Ai[p]] = a;
            }
            d[j] = 1.0 / sqrt(limit_scaling(cn));
        }
        for (int i = 0; i < m; ++i) e[i] = 1.0 / sqrt(limit_scaling(e[i]));
        double mean_P = 0.0, norm_q = 0.0;
        for (int j = 0; j < n; ++j) {
            s->P[j] *= d[j] * d[j];
            for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) s->A[p] *= e[s->Ai[p]] * d[j];
            s->q[j] *= d[j];
            s->D[j] *= d[j];
            mean_P += fabs(s->P[j]);
            if (fabs(s->q[j]) > norm_q) norm_q = fabs(s->q[j]);
        }
        for (int i = 0; i < m; ++i) s->E[i] *= e[i];
        mean_P = limit_scaling(n ? mean_P / n : 0.0);
        norm_q = (g_ignore_zero_q && norm_q == 0.0) ? 0.0 : limit_scaling(norm_q);
        const double ci = 1.0 / (mean_P > norm_q ? mean_P : norm_q);
        for (int j = 0; j < n; ++j) { s->P[j] *= ci; s->q[j] *= ci; }
        s->c *= ci;
    }
    for (int i = 0; i < m; ++i) { s->l[i] = scale_bound(s->l0[i], s->E[i]); s->u[i] = scale_bound(s->u0[i], s->E[i]); }
}

static int osqp_set_rho_vec(osqp_t* s) {
    int changed = !s->have_kind;
    for (int i = 0; i < s->m; ++i) {
        const int lo_inf = s->l0[i] <= -OSQP_INFTY * MIN_SCALING, up_inf = s->u0[i] >= OSQP_INFTY * MIN_SCALING;
        const int kind = (lo_inf && up_inf) ? -1 : (fabs(s->u0[i] - s->l0[i]) < RHO_TOL ? 1 : 0);
        if (kind != s->kind[i]) changed = 1;
        s->kind[i] = kind;
        double r = kind == -1 ? RHO_MIN : (kind == 1 ? RHO_EQ_OVER_INEQ * s->rho : s->rho);
        if (r < RHO_MIN) r = RHO_MIN;
        if (r > RHO_MAX * RHO_EQ_OVER_INEQ) r = RHO_MAX * RHO_EQ_OVER_INEQ;
        s->rho_vec[i] = r;
    }
    s->have_kind = 1;
    return changed;
}

static int osqp_factor(osqp_t* s) {
    for (int j = 0; j < s->n; ++j) s->F.Ux[s->posDx[j]] = s->P[j] + s->sigma;
    for (int i = 0; i < s->m; ++i) s->F.Ux[s->posDz[i]] = -1.0 / s->rho_vec[i];
    for (int p = 0; p < s->nnz; ++p) s->F.Ux[s->posA[p]] = s->A[p];
    s->factorizations++;
    return ldl_numeric(&s->F);
}

/* KKT pattern, ordering, symbolic factorisation: once per problem (the pattern never changes) */
static void osqp_analyse(osqp_t* s) {
    const int n = s->n, m = s->m, nk = n + m;
    unsigned char* adj = (unsigned char*)calloc((size_t)nk * nk, 1);
    for (int j = 0; j < n; ++j)
        for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) {
            const int i = n + s->Ai[p];
            adj[(size_t)i * nk + j] = 1; adj[(size_t)j * nk + i] = 1;
        }
    ldl_t* F = &s->F;
    F->n = nk;
    F->perm = (int*)malloc(sizeof(int) * nk);
    F->iperm = (int*)malloc(sizeof(int) * nk);
    min_degree_order(nk, adj, F->perm);
    free(adj);
    for (int i = 0; i < nk; ++i) F->iperm[F->perm[i]] = i;
    /* permuted upper triangle: entry (r, c) of K goes to column max(r', c'), row min(r', c') */
    const int nnzU = nk + s->nnz;
    F->Up = (int*)calloc(nk + 1, sizeof(int));
    F->Ui = (int*)malloc(sizeof(int) * nnzU);
    F->Ux = (double*)calloc(nnzU, sizeof(double));
    int* cnt = (int*)calloc(nk, sizeof(int));
    for (int k = 0; k < nk; ++k) cnt[k] = 1;
    for (int j = 0; j < n; ++j)
        for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) {
            const int a = F->iperm[j], b = F->iperm[n + s->Ai[p]];
            cnt[a > b ? a : b]++;
        }
    for (int k = 0; k < nk; ++k) F->Up[k + 1] = F->Up[k] + cnt[k];
    int* fill = (int*)malloc(sizeof(int) * nk);
    for (int k = 0; k < nk; ++k) fill[k] = F->Up[k];
    s->posA = (int*)malloc(sizeof(int) * s->nnz);
    s->posDx = (int*)malloc(sizeof(int) * n);
    s->posDz = (int*)malloc(sizeof(int) * m);
    for (int j = 0; j < n; ++j) { const int k = F->iperm[j]; s->posDx[j] = fill[k]; F->Ui[fill[k]++] = k; }
    for (int i = 0; i < m; ++i) { const int k = F->iperm[n + i]; s->posDz[i] = fill[k]; F->Ui[fill[k]++] = k; }
    for (int j = 0; j < n; ++j)
        for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) {
            const int a = F->iperm[j], b = F->iperm[n + s->Ai[p]];
            const int col = a > b ? a : b, row = a > b ? b : a;
            s->posA[p] = fill[col];
            F->Ui[fill[col]++] = row;
        }
    free(cnt); free(fill);
    F->Lp = (int*)malloc(sizeof(int) * (nk + 1));
    F->parent = (int*)malloc(sizeof(int) * nk);
    F->lnz = (int*)malloc(sizeof(int) * nk);
    F->flag = (int*)malloc(sizeof(int) * nk);
    F->pattern = (int*)malloc(sizeof(int) * nk);
    F->D = (double*)malloc(sizeof(double) * nk);
    F->Dinv = (double*)malloc(sizeof(double) * nk);
    F->Y = (double*)calloc(nk, sizeof(double));
    F->work = (double*)malloc(sizeof(double) * nk);
    ldl_symbolic(F);
    const int nnzL = F->Lp[nk];
    F->Li = (int*)malloc(sizeof(int) * (nnzL > 0 ? nnzL : 1));
    F->Lx = (double*)malloc(sizeof(double) * (nnzL > 0 ? nnzL : 1));
}

static void osqp_defaults(osqp_t* s) {
    s->rho0 = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->scaling = 10; s->max_iter = 4000;
    s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->check_termination = 25;
    s->adaptive_rho = 1; s->adaptive_interval = 100; s->adaptive_tol = 5.0;
}

static double* dalloc(int n) { return (double*)calloc(n > 0 ? n : 1, sizeof(double)); }

/* prob.setup(P, q, A, l, u)   [MPC.py:414]  -- P diagonal (all MPC.py ever passes) */
static int osqp_setup(osqp_t* s, int n, int m, const int* Ap, const int* Ai, const double* Ax, const double* Pdiag,
                      const double* q, const double* l, const double* u) {
    const int nnz = Ap[n];
    s->n = n; s->m = m; s->nnz = nnz;
    s->Ap = (int*)malloc(sizeof(int) * (n + 1)); memcpy(s->Ap, Ap, sizeof(int) * (n + 1));
    s->Ai = (int*)malloc(sizeof(int) * nnz); memcpy(s->Ai, Ai, sizeof(int) * nnz);
    s->A0 = dalloc(nnz); s->A = dalloc(nnz); memcpy(s->A0, Ax, sizeof(double) * nnz);
    s->P0 = dalloc(n); s->P = dalloc(n); memcpy(s->P0, Pdiag, sizeof(double) * n);
    s->q0 = dalloc(n); s->q = dalloc(n); if (q) memcpy(s->q0, q, sizeof(double) * n);
    s->l0 = dalloc(m); s->u0 = dalloc(m); s->l = dalloc(m); s->u = dalloc(m);
    for (int i = 0; i < m; ++i) { s->l0[i] = l[i] < -OSQP_INFTY ? -OSQP_INFTY : l[i]; s->u0[i] = u[i] > OSQP_INFTY ? OSQP_INFTY : u[i]; }
    s->D = dalloc(n); s->E = dalloc(m);
    s->x = dalloc(n); s->z = dalloc(m); s->y = dalloc(m);
    s->rho_vec = dalloc(m); s->kind = (int*)calloc(m, sizeof(int)); s->have_kind = 0;
    s->rhs = dalloc(n + m); s->sol = dalloc(n + m);
    s->t_n = dalloc(n); s->t_m = dalloc(m); s->t_n2 = dalloc(n); s->t_m2 = dalloc(m);
    s->rho = s->rho0;
    s->factorizations = 0;
    osqp_analyse(s);
    osqp_scale(s);
    osqp_set_rho_vec(s);
    return osqp_factor(s);
}

/* prob.update(Ax=.., l=.., u=..)   [MPC.py:419]: unscaled copies are kept, so OSQP 0.6's "unscale, overwrite,
 * rescale" on a matrix update is "overwrite, rescale"; the stored (scaled) iterates follow the change of scaling */
static int osqp_update(osqp_t* s, const double* Ax, const double* l, const double* u) {
    const int n = s->n, m = s->m;
    if (l) for (int i = 0; i < m; ++i) s->l0[i] = l[i] < -OSQP_INFTY ? -OSQP_INFTY : l[i];
    if (u) for (int i = 0; i < m; ++i) s->u0[i] = u[i] > OSQP_INFTY ? OSQP_INFTY : u[i];
    for (int i = 0; i < m; ++i) if (s->l0[i] > s->u0[i]) return -2;
    if (Ax) {
        memcpy(s->A0, Ax, sizeof(double) * s->nnz);
        for (int j = 0; j < n; ++j) s->x[j] *= s->D[j];
        for (int i = 0; i < m; ++i) { s->z[i] /= s->E[i]; s->y[i] *= s->E[i] / s->c; }
        osqp_scale(s);
        for (int j = 0; j < n; ++j) s->x[j] /= s->D[j];
        for (int i = 0; i < m; ++i) { s->z[i] *= s->E[i]; s->y[i] *= s->c / s->E[i]; }
        osqp_set_rho_vec(s);
        return osqp_factor(s);
    }
    for (int i = 0; i < m; ++i) { s->l[i] = scale_bound(s->l0[i], s->E[i]); s->u[i] = scale_bound(s->u0[i], s->E[i]); }
    if (osqp_set_rho_vec(s)) return osqp_factor(s);
    return 0;
}

static void spmv_A(const osqp_t* s, const double* x, double* out) {        /* out = A x */
    for (int i = 0; i < s->m; ++i) out[i] = 0.0;
    for (int j = 0; j < s->n; ++j) {
        const double xj = x[j];
        for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) out[s->Ai[p]] += s->A[p] * xj;
    }
}
static void spmv_At(const osqp_t* s, const double* y, double* out) {       /* out = A' y */
    for (int j = 0; j < s->n; ++j) {
        double acc = 0.0;
        for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) acc += s->A[p] * y[s->Ai[p]];
        out[j] = acc;
    }
}

/* prob.warm_start(x=initx)   [MPC.py:420]: x, z = A x; the previous y is kept */
static void osqp_warm_start_x(osqp_t* s, const double* x) {
    for (int j = 0; j < s->n; ++j) s->x[j] = x[j] / s->D[j];
    spmv_A(s, s->x, s->z);
}

static void osqp_residuals(osqp_t* s, double* pri, double* dua, double* n_pri, double* n_dua) {
    const int n = s->n, m = s->m;
    double* Ax = s->t_m; double* Aty = s->t_n;
    spmv_A(s, s->x, Ax);
    spmv_At(s, s->y, Aty);
    double p = 0, nAx = 0, nz = 0, d = 0, nPx = 0, nAty = 0, nq = 0;
    for (int i = 0; i < m; ++i) {
        const double ei = 1.0 / s->E[i];
        const double r = fabs(ei * (Ax[i] - s->z[i])); if (r > p) p = r;
        const double a = fabs(ei * Ax[i]); if (a > nAx) nAx = a;
        const double b = fabs(ei * s->z[i]); if (b > nz) nz = b;
    }
    for (int j = 0; j < n; ++j) {
        const double dj = 1.0 / s->D[j], Px = s->P[j] * s->x[j];
        const double r = fabs(dj * (Px + s->q[j] + Aty[j])); if (r > d) d = r;
        const double a = fabs(dj * Px); if (a > nPx) nPx = a;
        const double b = fabs(dj * Aty[j]); if (b > nAty) nAty = b;
        const double c = fabs(dj * s->q[j]); if (c > nq) nq = c;
    }
    *pri = p; *dua = d / s->c;
    *n_pri = nAx > nz ? nAx : nz;
    double nd = nPx > nAty ? nPx : nAty; nd = nd > nq ? nd : nq;
    *n_dua = nd / s->c;
}

/* prob.solve()   [MPC.py:427] */
static int osqp_solve(osqp_t* s) {
    const int n = s->n, m = s->m;
    const double sigma = s->sigma, alpha = s->alpha;
    double pri = INFINITY, dua = INFINITY;
    s->status = 2; s->rho_updates = 0;
    int it;
    for (it = 1; it <= s->max_iter; ++it) {
        for (int j = 0; j < n; ++j) s->rhs[j] = sigma * s->x[j] - s->q[j];
        for (int i = 0; i < m; ++i) s->rhs[n + i] = s->z[i] - s->y[i] / s->rho_vec[i];
        ldl_solve(&s->F, s->rhs, s->sol);
        for (int j = 0; j < n; ++j) s->x[j] = alpha * s->sol[j] + (1.0 - alpha) * s->x[j];
        for (int i = 0; i < m; ++i) {
            const double rv = s->rho_vec[i];
            const double zt = s->z[i] + (s->sol[n + i] - s->y[i]) / rv;
            const double zr = alpha * zt + (1.0 - alpha) * s->z[i];
            double zn = zr + s->y[i] / rv;
            zn = zn < s->l[i] ? s->l[i] : (zn > s->u[i] ? s->u[i] : zn);
            s->y[i] += rv * (zr - zn);
            s->z[i] = zn;
        }
        const int check = s->check_termination && it % s->check_termination == 0;
        const int adapt = s->adaptive_rho && s->adaptive_interval && it % s->adaptive_interval == 0;
        if (check || adapt) {
            double n_pri, n_dua;
            osqp_residuals(s, &pri, &dua, &n_pri, &n_dua);
            if (check && pri <= s->eps_abs + s->eps_rel * n_pri && dua <= s->eps_abs + s->eps_rel * n_dua) { s->status = 1; break; }
            if (adapt) {
                const double pn = pri / (n_pri + 1e-10), dn = dua / (n_dua + 1e-10);
                double rn = s->rho * sqrt(pn / (dn + 1e-10));
                rn = rn < RHO_MIN ? RHO_MIN : (rn > RHO_MAX ? RHO_MAX : rn);
                if (rn > s->rho * s->adaptive_tol || rn < s->rho / s->adaptive_tol) {
                    s->rho = rn;
                    osqp_set_rho_vec(s);
                    if (osqp_factor(s)) return -1;
                    s->rho_updates++;
                }
            }
        }
    }
    s->iter = it > s->max_iter ? s->max_iter : it;
    s->pri_res = pri; s->dua_res = dua;
    return 0;
}

static void osqp_free(osqp_t* s) {
    free(s->Ap); free(s->Ai); free(s->A0); free(s->A); free(s->P0); free(s->P); free(s->q0); free(s->q);
    free(s->l0); free(s->u0); free(s->l); free(s->u); free(s->D); free(s->E); free(s->x); free(s->z); free(s->y);
    free(s->rho_vec); free(s->kind); free(s->posA); free(s->posDx); free(s->posDz); free(s->rhs); free(s->sol);
    free(s->t_n); free(s->t_m); free(s->t_n2); free(s->t_m2);
    ldl_free(&s->F);
}

/* ------------------------------------------------------------------------------------------------ MPC.py build */
typedef struct {
    mpc_params p;
    int n, m, nnz;
    int *Ap, *Ai;
    double *Ax, *l, *u, *Pd;
    double* x;          /* MPC.x of the last tick (24 N) */
    double* warm;
    osqp_t qp;
    int is_setup;
    double eps;
} mpc_oracle;

/* row indices / column pointers of ML (MPC.py:151; layout: SURVEY.md appendix A) */
static void build_pattern(mpc_oracle* o) {
    const int N = o->p.N;
    int pos = 0, col = 0;
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) {
            o->Ap[col++] = pos;
            o->Ai[pos++] = 12 * k + i;
            if (k < N - 1) {
                if (i >= 6) o->Ai[pos++] = 12 * (k + 1) + i - 6;
                o->Ai[pos++] = 12 * (k + 1) + i;
            }
        }
    for (int k = 0; k < N; ++k)
        for (int j = 0; j < 4; ++j)
            for (int c = 0; c < 3; ++c) {
                o->Ap[col++] = pos;
                o->Ai[pos++] = 12 * k + 6 + c;
                o->Ai[pos++] = 12 * k + 9; o->Ai[pos++] = 12 * k + 10; o->Ai[pos++] = 12 * k + 11;
                o->Ai[pos++] = 12 * N + 12 * k + 3 * j + c;
                const int base = 24 * N + 20 * k + 5 * j;
                if (c == 0) { o->Ai[pos++] = base; o->Ai[pos++] = base + 1; }
                else if (c == 1) { o->Ai[pos++] = base + 2; o->Ai[pos++] = base + 3; }
                else for (int r = 0; r < 5; ++r) o->Ai[pos++] = base + r;
            }
    o->Ap[col] = pos;
}

static void inv3(const double* a, double* o) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

/* xref: 12 x (N+1) row major, fsteps: 20 x 13 row major (NaN = swing).  Neither is written. */
static void build_values(mpc_oracle* o, const double* xref, const double* fsteps, int first_tick) {
    const mpc_params* p = &o->p;
    const int N = p->N, ld = N + 1;
    const double dt = p->dt, mu = p->mu;
    /* construct_gait (MPC.py:635-652): contact iff x is neither NaN nor 0; rows until the first count of 0 */
    int row_of_step[256]; double contact[256][4];
    for (int k = 0; k < N; ++k) { row_of_step[k] = -1; for (int j = 0; j < 4; ++j) contact[k][j] = 0.0; }
    int k0 = 0;
    for (int r = 0; r < 20; ++r) {
        const double cnt = fsteps[r * 13];
        if (cnt == 0.0) break;
        const int c = (int)cnt;
        for (int k = k0; k < k0 + c && k < N; ++k) {
            row_of_step[k] = r;
            for (int j = 0; j < 4; ++j) { const double x = fsteps[r * 13 + 1 + 3 * j]; contact[k][j] = (isnan(x) || x == 0.0) ? 0.0 : 1.0; }
        }
        k0 += c;
    }
    int pos = 0;
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) {
            o->Ax[pos++] = -1.0;
            if (k < N - 1) { if (i >= 6) o->Ax[pos++] = dt; o->Ax[pos++] = 1.0; }        /* MPC.py:110-111 */
        }
    for (int k = 0; k < N; ++k) {
        const double cs = cos(xref[5 * ld + k]), sn = sin(xref[5 * ld + k]);             /* MPC.py:330 */
        const double R[9] = {cs, -sn, 0, sn, cs, 0, 0, 0, 1};
        double RgI[9], Iinv[9];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) RgI[3 * a + b] = R[3 * a] * p->gI[b] + R[3 * a + 1] * p->gI[3 + b] + R[3 * a + 2] * p->gI[6 + b];
        inv3(RgI, Iinv);                                                                  /* MPC.py:339-340 */
        for (int j = 0; j < 4; ++j) {
            double foot[3] = {0, 0, 0};
            int have = 0;
            if (first_tick) { for (int c = 0; c < 3; ++c) foot[c] = p->footholds[c * 4 + j]; have = 1; }   /* MPC.py:176 */
            else if (row_of_step[k] >= 0) {
                for (int c = 0; c < 3; ++c) { const double v = fsteps[row_of_step[k] * 13 + 1 + 3 * j + c]; foot[c] = isnan(v) ? 0.0 : v; }  /* :327 */
                have = 1;
            }
            double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (have) {
                const double r0 = foot[0] - xref[0 * ld + k], r1 = foot[1] - xref[1 * ld + k], r2 = foot[2] - xref[2 * ld + k];   /* :343 */
                const double S[9] = {0, -r2, r1, r2, 0, -r0, -r1, r0, 0};                 /* utils.py:179-185 */
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) B[3 * a + b] = dt * (Iinv[3 * a] * S[b] + Iinv[3 * a + 1] * S[3 + b] + Iinv[3 * a + 2] * S[6 + b]);
            }
            for (int c = 0; c < 3; ++c) {
                o->Ax[pos++] = dt / p->mass;                                              /* MPC.py:119 */
                o->Ax[pos++] = B[c]; o->Ax[pos++] = B[3 + c]; o->Ax[pos++] = B[6 + c];
                o->Ax[pos++] = 1.0 - contact[k][j];                                       /* MPC.py:628-630 */
                if (c < 2) { o->Ax[pos++] = 1.0; o->Ax[pos++] = -1.0; }
                else { for (int r = 0; r < 4; ++r) o->Ax[pos++] = -mu; o->Ax[pos++] = -1.0; }
            }
        }
    }
    /* bounds (MPC.py:200-232, 362-378, 410) */
    for (int i = 0; i < o->m; ++i) { o->u[i] = 0.0; o->l[i] = 0.0; }
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) {
            double v = xref[i * ld + k + 1];
            if (i == 8) v += p->gravity * dt;
            double ax = xref[i * ld + k];
            if (i < 6) ax += dt * xref[(i + 6) * ld + k];
            o->u[12 * k + i] = v - ax;
            o->l[12 * k + i] = v - ax;
        }
    for (int i = 24 * N; i < 44 * N; ++i) o->l[i] = ((i - 24 * N) % 5 == 4) ? -p->fz_max : -INFINITY;
}

/* ------------------------------------------------------------------------------------------------ C ABI of the oracle */
void* mpc_oracle_create(int n_steps, double dt, double eps) {
    if (n_steps < 1 || n_steps > 256) return NULL;
    mpc_oracle* o = (mpc_oracle*)calloc(1, sizeof *o);
    params_default(&o->p, n_steps, dt);
    const int N = n_steps;
    o->n = 24 * N; o->m = 44 * N; o->nnz = 126 * N - 18;
    o->Ap = (int*)malloc(sizeof(int) * (o->n + 1));
    o->Ai = (int*)malloc(sizeof(int) * o->nnz);
    o->Ax = dalloc(o->nnz); o->l = dalloc(o->m); o->u = dalloc(o->m); o->Pd = dalloc(o->n);
    o->x = dalloc(o->n); o->warm = dalloc(o->n);
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) { o->Pd[12 * k + i] = o->p.w_state[i]; o->Pd[12 * N + 12 * k + i] = o->p.w_force; }
    build_pattern(o);
    o->eps = eps;
    return o;
}

void mpc_oracle_destroy(void* h) {
    mpc_oracle* o = (mpc_oracle*)h;
    if (!o) return;
    if (o->is_setup) osqp_free(&o->qp);
    free(o->Ap); free(o->Ai); free(o->Ax); free(o->l); free(o->u); free(o->Pd); free(o->x); free(o->warm);
    free(o);
}

int mpc_oracle_nnz(void* h) { return ((mpc_oracle*)h)->nnz; }

/* The per-tick coefficients alone (parity of the build half): Ax (nnz), l, u (44 N), pattern (Ap 24N+1, Ai nnz) */
int mpc_oracle_build(void* h, const double* xref, const double* fsteps, int first_tick, double* Ax, double* l, double* u,
                     int* Ap, int* Ai) {
    mpc_oracle* o = (mpc_oracle*)h;
    build_values(o, xref, fsteps, first_tick);
    if (Ax) memcpy(Ax, o->Ax, sizeof(double) * o->nnz);
    if (l) memcpy(l, o->l, sizeof(double) * o->m);
    if (u) memcpy(u, o->u, sizeof(double) * o->m);
    if (Ap) memcpy(Ap, o->Ap, sizeof(int) * (o->n + 1));
    if (Ai) memcpy(Ai, o->Ai, sizeof(int) * o->nnz);
    return 0;
}

/* One MPC.run(k, xref, fsteps) (MPC.py:460-514): build, shifted warm start, setup / update, solve, extraction.
 * out_x: 24 N (MPC.x), out_y: 44 N multipliers (unscaled) or NULL, out_f: 12 (f_applied) or NULL,
 * info[4] = {iterations, status (1 solved), rho updates, factorisations so far} or NULL. */
int mpc_oracle_run(void* h, int first_tick, const double* xref, const double* fsteps, double* out_x, double* out_y, double* out_f,
                   double* info) {
    mpc_oracle* o = (mpc_oracle*)h;
    const int N = o->p.N, n = o->n, m = o->m;
    build_values(o, xref, fsteps, first_tick);
    int rc;
    if (first_tick || !o->is_setup) {
        if (o->is_setup) { osqp_free(&o->qp); o->is_setup = 0; }
        memset(&o->qp, 0, sizeof o->qp);
        osqp_defaults(&o->qp);
        o->qp.eps_abs = o->eps; o->qp.eps_rel = o->eps;                       /* MPC.py:415-416 */
        rc = osqp_setup(&o->qp, n, m, o->Ap, o->Ai, o->Ax, o->Pd, NULL, o->l, o->u);      /* MPC.py:414 */
        o->is_setup = 1;
    } else {
        /* MPC.py:403-406: previous solution advanced by one stage, last state block zeroed, f_0 wraps round */
        for (int i = 0; i < 12 * N - 12; ++i) o->warm[i] = o->x[i + 12];
        for (int i = 12 * N - 12; i < 12 * N; ++i) o->warm[i] = 0.0;
        for (int i = 0; i < 12 * N; ++i) o->warm[12 * N + i] = o->x[12 * N + (i + 12) % (12 * N)];
        rc = osqp_update(&o->qp, o->Ax, o->l, o->u);                          /* MPC.py:419 */
        osqp_warm_start_x(&o->qp, o->warm);                                   /* MPC.py:420 */
    }
    if (rc) return rc;
    rc = osqp_solve(&o->qp);                                                  /* MPC.py:427 */
    if (rc) return rc;
    for (int j = 0; j < n; ++j) o->x[j] = o->qp.D[j] * o->qp.x[j];            /* MPC.py:428 */
    if (out_x) memcpy(out_x, o->x, sizeof(double) * n);
    if (out_y) for (int i = 0; i < m; ++i) out_y[i] = o->qp.E[i] * o->qp.y[i] / o->qp.c;
    if (out_f) memcpy(out_f, o->x + 12 * N, sizeof(double) * 12);             /* MPC.py:440 */
    if (info) { info[0] = o->qp.iter; info[1] = o->qp.status; info[2] = o->qp.rho_updates; info[3] = o->qp.factorizations; }
    return 0;
}

/* ---- CPU baseline: `threads` robots, one per host thread, each replaying its own recorded closed-loop input
 * sequence (T ticks of xref / fsteps) through mpc_oracle_run; ticks < warm are untimed.  Returns the wall time
 * in seconds of the slowest thread over its timed ticks; out_f (threads x T x 12) receives the forces. */
typedef struct { int N, T, warm; double dt, eps; const double *xref, *fsteps; double* out_f; double seconds; long iters; int rc; } replay_job;

static void* replay_thread(void* arg) {
    replay_job* j = (replay_job*)arg;
    void* h = mpc_oracle_create(j->N, j->dt, j->eps);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const size_t sx = (size_t)12 * (j->N + 1), sf = 260;
    j->iters = 0; j->rc = 0;
    for (int t = 0; t < j->T; ++t) {
        if (t == j->warm) clock_gettime(CLOCK_MONOTONIC, &t0);
        double info[4];
        const int rc = mpc_oracle_run(h, t == 0, j->xref + t * sx, j->fsteps + t * sf, NULL, NULL, j->out_f ? j->out_f + (size_t)t * 12 : NULL, info);
        if (rc) j->rc = rc;
        if (t >= j->warm) j->iters += (long)info[0];
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    j->seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    mpc_oracle_destroy(h);
    return NULL;
}

double mpc_oracle_replay_mt(int threads, int n_steps, double dt, double eps, int T, int warm, const double* xref, const double* fsteps,
                            double* out_f, double* mean_iters) {
    replay_job* jobs = (replay_job*)calloc(threads, sizeof *jobs);
    pthread_t* th = (pthread_t*)calloc(threads, sizeof *th);
    const size_t sx = (size_t)T * 12 * (n_steps + 1), sf = (size_t)T * 260;
    for (int i = 0; i < threads; ++i) {
        jobs[i].N = n_steps; jobs[i].T = T; jobs[i].warm = warm; jobs[i].dt = dt; jobs[i].eps = eps;
        jobs[i].xref = xref + i * sx; jobs[i].fsteps = fsteps + i * sf;
        jobs[i].out_f = out_f ? out_f + (size_t)i * T * 12 : NULL;
        pthread_create(&th[i], NULL, replay_thread, &jobs[i]);
    }
    double worst = 0.0; long iters = 0; int bad = 0;
    for (int i = 0; i < threads; ++i) {
        pthread_join(th[i], NULL);
        if (jobs[i].seconds > worst) worst = jobs[i].seconds;
        iters += jobs[i].iters;
        if (jobs[i].rc) bad = 1;
    }
    if (mean_iters) *mean_iters = (double)iters / ((double)threads * (T - warm > 0 ? T - warm : 1));
    free(jobs); free(th);
    return bad ? -1.0 : worst;
}

/* Generic QP entry (known-answer tests): diagonal P, CSC A.  x (n), y (m) out; returns iterations or < 0. */
int mpc_oracle_solve_qp(int n, int m, const int* Ap, const int* Ai, const double* Ax, const double* Pdiag, const double* q,
                        const double* l, const double* u, double eps, double* x, double* y) {
    osqp_t s;
    memset(&s, 0, sizeof s);
    osqp_defaults(&s);
    s.eps_abs = eps; s.eps_rel = eps;
    if (osqp_setup(&s, n, m, Ap, Ai, Ax, Pdiag, q, l, u)) { osqp_free(&s); return -1; }
    if (osqp_solve(&s)) { osqp_free(&s); return -1; }
    for (int j = 0; j < n; ++j) x[j] = s.D[j] * s.x[j];
    for (int i = 0; i < m; ++i) y[i] = s.E[i] * s.y[i] / s.c;
    const int it = s.status == 1 ? s.iter : -s.iter;
    osqp_free(&s);
    return it;
}
