/* ORACLE (test infrastructure) — Tikhonov-regularised NNLS and the regularisation-parameter
 * choosers (none / lcurve / gcv / chi2 / mdp).  Follows src/lsqnonneg.jl. */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "decaes_oracle.h"

#define NCACHE 8     /* Val(8), src/lsqnonneg.jl:396 */
#define LC_MAX 256   /* GrowableCache initial size 64 (src/lsqnonneg.jl:767-771); grows if needed */

/* NNLSTikhonovRegProblem  src/lsqnonneg.jl:250-269 */
typedef struct {
  orc_nnls_work *nnls; /* (m+n) x n workspace */
  double mu;           /* regparam, NaN = empty slot */
  double *tmp;         /* n */
} tikh_prob;

/* NNLSTikhonovRegProblemCache  src/lsqnonneg.jl:392-415 */
struct orc_tikh_cache {
  tikh_prob slot[NCACHE];
  int idx; /* 1-based current slot */
  double *null_soln;
};

typedef struct {
  double x;
  double P[2];
} lc_fval; /* CachedFunction entry */
typedef struct {
  double x;
  double P[2];
  double C;
} lc_point; /* LCurveCornerPoint keyed by x */
typedef struct {
  double key;
  double x[4];
  double P[4][2];
} lc_state; /* LCurveCornerState keyed by iteration */
typedef struct {
  lc_fval fc[LC_MAX];
  int nfc;
  lc_point pc[LC_MAX];
  int npc;
  lc_state sc[LC_MAX];
  int nsc;
} lc_caches;

orc_reg_work *orc_reg_alloc(int m, int n) {
  orc_reg_work *w = (orc_reg_work *)calloc(1, sizeof(*w));
  w->m = m, w->n = n;
  w->nnls = orc_nnls_alloc(m, n);
  w->cache = (orc_tikh_cache *)calloc(1, sizeof(orc_tikh_cache));
  for (int i = 0; i < NCACHE; i++) {
    w->cache->slot[i].nnls = orc_nnls_alloc(m + n, n);
    w->cache->slot[i].mu = NAN;
    w->cache->slot[i].tmp = (double *)calloc(n, sizeof(double));
  }
  w->cache->idx = 1;
  w->cache->null_soln = (double *)calloc(n, sizeof(double));
  int mn = m < n ? m : n;
  w->gamma = (double *)calloc(mn, sizeof(double));
  w->svd_work = (double *)calloc((size_t)m * n + m + n, sizeof(double));
  w->lcurve = calloc(1, sizeof(lc_caches));
  return w;
}

void orc_reg_free(orc_reg_work *w) {
  if (!w) return;
  orc_nnls_free(w->nnls);
  for (int i = 0; i < NCACHE; i++) {
    orc_nnls_free(w->cache->slot[i].nnls);
    free(w->cache->slot[i].tmp);
  }
  free(w->cache->null_soln), free(w->cache), free(w->gamma), free(w->svd_work), free(w->lcurve), free(w);
}

orc_nnls_work *orc_reg_slot(orc_reg_work *w, int i) { return w->cache->slot[i].nnls; }

void orc_reg_bind(orc_reg_work *w, const double *A, const double *b) { w->A = A, w->b = b; }

/* ---- NNLSTikhonovRegProblem accessors ---- */
static double tikh_loss(const tikh_prob *p) { return p->nnls->rnorm * p->nnls->rnorm; } /* :307 */
static double tikh_seminorm_sq(const tikh_prob *p) { /* :318  sum(abs2, positive_solution) */
  const orc_nnls_work *k = p->nnls;
  double s = 0.0;
  for (int i = 0; i < k->nsetp; i++) {
    double xi = k->x[k->idx[i] - 1];
    s = fma(xi, xi, s);
  }
  return s;
}
static double tikh_resnorm_sq(const tikh_prob *p) { /* :309,:313 */
  double reg = (p->mu * p->mu) * tikh_seminorm_sq(p);
  double r = tikh_loss(p) - reg;
  return r > 0 ? r : 0.0;
}

/* ---- cache  :401-444 ---- */
static void reset_cache(orc_tikh_cache *c) {
  for (int i = 0; i < NCACHE; i++) c->slot[i].mu = NAN;
}
static int mod1(int i, int n) {
  int r = i % n;
  if (r <= 0) r += n;
  return r;
}
static void next_cache_index(orc_tikh_cache *c) {
  for (int i = 0; i < NCACHE; i++) {
    if (isnan(c->slot[i].mu)) {
      c->idx = mod1(i + 1, NCACHE);
      return;
    }
  }
  c->idx = mod1(c->idx + 1, NCACHE);
}
static tikh_prob *cache_cur(orc_tikh_cache *c) { return &c->slot[c->idx - 1]; }

static const double *cache_solve(orc_reg_work *w, double mu) {
  orc_tikh_cache *c = w->cache;
  int empty = 1, imax = 0;
  double dmax = INFINITY;
  for (int i = 0; i < NCACHE; i++) {
    double mui = c->slot[i].mu;
    if (!isnan(mui)) {
      empty = 0;
      double d = (mu == mui) ? 0.0 : fabs(log1p((mu - mui) / mui));
      imax = i + 1;
      if (d < dmax) dmax = d;
      if (dmax == 0) break;
    }
  }
  if (empty || dmax > 0) {
    next_cache_index(c);
    tikh_prob *p = cache_cur(c);
    p->mu = mu; /* regparam!  :299 */
    orc_nnls_solve_tikh(p->nnls, w->A, w->m, w->b, w->m, w->n, mu);
    w->n_solves_tikh++;
  } else {
    c->idx = mod1(imax, NCACHE);
    w->n_cache_hits++;
  }
  return cache_cur(c)->nnls->x;
}

static void solve_unreg(orc_reg_work *w) {
  orc_nnls_solve(w->nnls, w->A, w->m, w->b, w->m, w->n);
  w->n_solves_unreg++;
}

/* lsqnonneg!  :189-192 */
const double *orc_lsqnonneg(orc_reg_work *w) {
  solve_unreg(w);
  return w->nnls->x;
}

/* lsqnonneg_tikh!  :290-302 (through the cache so the L-curve helpers can be tested) */
const double *orc_lsqnonneg_tikh(orc_reg_work *w, double mu, double *res2, double *semi2) {
  reset_cache(w->cache);
  const double *x = cache_solve(w, mu);
  if (res2) *res2 = tikh_resnorm_sq(cache_cur(w->cache));
  if (semi2) *semi2 = tikh_seminorm_sq(cache_cur(w->cache));
  return x;
}

/* =================== L-curve  :812-972 =================== */
static int isapprox(double x, double y) { /* Base.isapprox, rtol = sqrt(eps), atol = 0 */
  if (x == y) return 1;
  if (!isfinite(x) || !isfinite(y)) return 0;
  const double rtol = 1.4901161193847656e-08;
  return fabs(x - y) <= rtol * fmax(fabs(x), fabs(y));
}
/* Base.isless for floats: NaN is larger than everything, -0.0 < 0.0 */
static int isless_f(double a, double b) {
  if (isnan(a)) return 0;
  if (isnan(b)) return 1;
  if (a == 0 && b == 0) return signbit(a) && !signbit(b);
  return a < b;
}

typedef struct {
  orc_lcurve_fn f;
  void *ctx;
  lc_caches *c;
  int nfeval;
} lc_fun;

/* CachedFunction call: get!(f, cache, x)  src/utils.jl:207-217, 229 */
static void lc_eval(lc_fun *F, double x, double *P) {
  lc_caches *c = F->c;
  for (int i = 0; i < c->nfc; i++) {
    if (isapprox(x, c->fc[i].x)) {
      P[0] = c->fc[i].P[0], P[1] = c->fc[i].P[1];
      return;
    }
  }
  F->f(x, P, F->ctx);
  F->nfeval++;
  if (c->nfc < LC_MAX) {
    c->fc[c->nfc].x = x, c->fc[c->nfc].P[0] = P[0], c->fc[c->nfc].P[1] = P[1];
    c->nfc++;
  }
}
static int pc_find(const lc_caches *c, double x) {
  for (int i = 0; i < c->npc; i++)
    if (isapprox(x, c->pc[i].x)) return i;
  return -1;
}
static void pc_push(lc_caches *c, double x, const double *P, double C) {
  if (c->npc >= LC_MAX) return;
  c->pc[c->npc].x = x, c->pc[c->npc].P[0] = P[0], c->pc[c->npc].P[1] = P[1], c->pc[c->npc].C = C;
  c->npc++;
}
static void pc_set(lc_caches *c, double x, const double *P, double C) { /* setindex!  src/utils.jl:163-171 */
  int i = pc_find(c, x);
  if (i < 0)
    pc_push(c, x, P, C);
  else
    c->pc[i].P[0] = P[0], c->pc[i].P[1] = P[1], c->pc[i].C = C;
}
static double norm2(const double *P, const double *Q) {
  double dx = P[0] - Q[0], dy = P[1] - Q[1];
  return sqrt(dx * dx + dy * dy);
}
/* menger  :967-972 */
static double menger(const double *Pj, const double *Pk, const double *Pl) {
  double jk[2] = {Pj[0] - Pk[0], Pj[1] - Pk[1]};
  double kl[2] = {Pk[0] - Pl[0], Pk[1] - Pl[1]};
  double lj[2] = {Pl[0] - Pj[0], Pl[1] - Pj[1]};
  double d1 = jk[0] * jk[0] + jk[1] * jk[1];
  double d2 = kl[0] * kl[0] + kl[1] * kl[1];
  double d3 = lj[0] * lj[0] + lj[1] * lj[1];
  double cross = jk[0] * kl[1] - jk[1] * kl[0];
  return 2 * cross / sqrt(d1 * d2 * d3);
}
/* update_curvature!  :948-965 */
static void update_curvature(lc_fun *F, const lc_state *s, const double *Ptl, const double *Pbr, double Ctol) {
  lc_caches *c = F->c;
  for (int i = 0; i < 4; i++) {
    double x = s->x[i];
    const double *P = s->P[i];
    double C = -INFINITY;
    if (fmin(norm2(P, Ptl), norm2(P, Pbr)) > Ctol) {
      double xm = -INFINITY, xp = INFINITY;
      const double *Pm = P, *Pp = P;
      for (int k = 0; k < c->npc; k++) {
        double _x = c->pc[k].x;
        if (xm < _x && _x < x) xm = _x, Pm = c->pc[k].P;
        if (x < _x && _x < xp) xp = _x, Pp = c->pc[k].P;
      }
      C = menger(Pm, P, Pp);
    }
    pc_set(c, x, P, C);
  }
}
/* mapfindmax(C) over the point cache: first maximum, NaN is maximal  :894, :915 */
static double pc_argmax(const lc_caches *c) {
  int best = 0;
  for (int i = 1; i < c->npc; i++)
    if (isless_f(c->pc[best].C, c->pc[i].C)) best = i;
  return c->pc[best].x;
}

double orc_lcurve_corner(orc_lcurve_fn f, void *ctx, double xlow, double xhigh, double xtol, double Ptol,
                         double Ctol, int backtracking, int *n_feval) {
  static __thread lc_caches tl_caches;
  lc_fun F = {f, ctx, &tl_caches, 0};
  lc_caches *c = F.c;
  c->nfc = c->npc = c->nsc = 0;
  const double phi = 1.618033988749895;

  /* initial_state  :921-929 */
  lc_state st;
  {
    double x1 = xlow, x4 = xhigh;
    double x2 = (phi * x1 + x4) / (phi + 1);
    double x3 = x1 + (x4 - x2);
    st.x[0] = x1, st.x[1] = x2, st.x[2] = x3, st.x[3] = x4;
    for (int i = 0; i < 4; i++) lc_eval(&F, st.x[i], st.P[i]);
    for (int i = 0; i < 4; i++) pc_push(c, st.x[i], st.P[i], -INFINITY);
  }
  double Ptl[2] = {st.P[0][0], st.P[0][1]}, Pbr[2] = {st.P[3][0], st.P[3][1]};
  update_curvature(&F, &st, Ptl, Pbr, Ctol);

  int iter = 0;
  /* is_converged  :931 */
  while (!(fabs(st.x[3] - st.x[0]) < xtol || norm2(st.P[0], st.P[3]) < Ptol)) {
    iter += 1;
    if (backtracking) { /* :892-900 */
      double x = pc_argmax(c);
      for (int k = 0; k < c->nsc; k++) {
        const lc_state *s = &c->sc[k];
        if ((s->x[1] == x || s->x[2] == x) && fabs(s->x[3] - s->x[0]) <= fabs(st.x[3] - st.x[0])) st = *s;
      }
    }
    double C2 = c->pc[pc_find(c, st.x[1])].C, C3 = c->pc[pc_find(c, st.x[2])].C;
    lc_state nw;
    if (C2 > C3) { /* move_left  :933-939 */
      nw.x[0] = st.x[0], nw.x[1] = (phi * st.x[0] + st.x[2]) / (phi + 1), nw.x[2] = st.x[1], nw.x[3] = st.x[2];
      memcpy(nw.P[0], st.P[0], 16), memcpy(nw.P[2], st.P[1], 16), memcpy(nw.P[3], st.P[2], 16);
      lc_eval(&F, nw.x[1], nw.P[1]);
    } else { /* move_right  :941-946 */
      nw.x[0] = st.x[1], nw.x[1] = st.x[2], nw.x[2] = st.x[1] + (st.x[3] - st.x[2]), nw.x[3] = st.x[3];
      memcpy(nw.P[0], st.P[1], 16), memcpy(nw.P[1], st.P[2], 16), memcpy(nw.P[3], st.P[3], 16);
      lc_eval(&F, nw.x[2], nw.P[2]);
    }
    st = nw;
    update_curvature(&F, &st, Ptl, Pbr, Ctol);
    if (backtracking && c->nsc < LC_MAX) {
      st.key = (double)iter;
      c->sc[c->nsc++] = st;
    }
    if (iter > 10000) break; /* safety net only; the reference has no iteration cap */
  }
  if (n_feval) *n_feval = F.nfeval;
  return pc_argmax(c);
}

static void f_lcurve(double logmu, double *P, void *ctx) { /* :819-824 */
  orc_reg_work *w = (orc_reg_work *)ctx;
  cache_solve(w, exp(logmu));
  const tikh_prob *p = cache_cur(w->cache);
  P[0] = log(tikh_resnorm_sq(p));
  P[1] = log(tikh_seminorm_sq(p));
}

/* lsqnonneg_lcurve!  :812-840 */
const double *orc_lsqnonneg_lcurve(orc_reg_work *w, double *mu, double *chi2) {
  reset_cache(w->cache);
  double logmu = orc_lcurve_corner(f_lcurve, w, -8.0, 2.0, 1e-4, 1e-4, 1e-4, 1, NULL);
  double mu_final = exp(logmu);
  const double *x_final = cache_solve(w, mu_final);
  solve_unreg(w);
  double r2u = w->nnls->rnorm * w->nnls->rnorm;
  *mu = mu_final;
  *chi2 = tikh_resnorm_sq(cache_cur(w->cache)) / r2u;
  return x_final;
}

/* =================== chi2  :504-593 (method = :brent) =================== */
typedef struct {
  orc_reg_work *w;
  double target;
  int mode; /* 0: chi2 relative error, 1: mdp absolute error */
} root_ctx;

static double f_root(double logmu, void *ctx) {
  root_ctx *r = (root_ctx *)ctx;
  cache_solve(r->w, exp(logmu));
  double res2 = tikh_resnorm_sq(cache_cur(r->w->cache));
  if (r->mode == 0) return (res2 - r->target) / r->target; /* chi2_relerr!  :374-385 */
  return res2 - r->target;                                  /* :723-726 (target = delta^2) */
}

const double *orc_lsqnonneg_chi2(orc_reg_work *w, double chi2_target, double *mu, double *chi2, int *early) {
  solve_unreg(w);
  const double *x_unreg = w->nnls->x;
  double res2_min = w->nnls->rnorm * w->nnls->rnorm;
  if (early) *early = 0;
  if (res2_min == 0 || w->nnls->nsetp == 0) {
    /* :510-515.  NOTE: the reference's save_results! would read the (stale) cache slot here,
     * src/lsqnonneg.jl:465; the oracle hands back the function's own return value. */
    if (early) *early = 1;
    *mu = 0.0, *chi2 = 1.0;
    return x_unreg;
  }
  double res2_target = chi2_target * res2_min;
  reset_cache(w->cache);
  root_ctx rc = {w, res2_target, 0};
  double a, b, fa, fb;
  orc_bracket_root_monotonic(f_root, &rc, -4.0, 1.0, 1.5, +1, 6, &a, &b, &fa, &fb);
  double logmu_final, relerr_final;
  if (fa * fb < 0) {
    orc_brent_root(f_root, &rc, a, b, fa, fb, 0.0, 0.0, 1e-3 * (chi2_target - 1), 100, &logmu_final, &relerr_final);
  } else {
    if (!isfinite(fa))
      logmu_final = b, relerr_final = fb;
    else if (!isfinite(fb))
      logmu_final = a, relerr_final = fa;
    else if (fabs(fa) < fabs(fb))
      logmu_final = a, relerr_final = fa;
    else
      logmu_final = b, relerr_final = fb;
  }
  if (isfinite(relerr_final)) {
    double mu_final = exp(logmu_final);
    double res2_final = res2_target * (1 + relerr_final); /* chi2_relerr^-1  :386 */
    const double *x_final = cache_solve(w, mu_final);
    *mu = mu_final, *chi2 = res2_final / res2_min;
    return x_final;
  }
  *mu = 0.0, *chi2 = 1.0 / res2_min;
  return x_unreg;
}

/* lsqnonneg_chi2!(work, chi2_target, legacy = true)  :504-533: method = :legacy */
typedef struct {
  orc_reg_work *w;
  double res2_min;
} legacy_ctx;

static double f_res2_mu(double mu, void *ctx) { /* the do-block :523-527 */
  legacy_ctx *c = (legacy_ctx *)ctx;
  if (mu == 0) return c->res2_min;
  cache_solve(c->w, mu);
  return tikh_resnorm_sq(cache_cur(c->w->cache));
}

const double *orc_lsqnonneg_chi2_legacy(orc_reg_work *w, double chi2_target, double *mu, double *chi2, int *early) {
  solve_unreg(w);
  const double *x_unreg = w->nnls->x;
  double res2_min = w->nnls->rnorm * w->nnls->rnorm;
  if (early) *early = 0;
  if (res2_min == 0 || w->nnls->nsetp == 0) { /* :510-515, same stale-slot note as the :brent method */
    if (early) *early = 1;
    *mu = 0.0, *chi2 = 1.0;
    return x_unreg;
  }
  reset_cache(w->cache);
  legacy_ctx lc = {w, res2_min};
  double mu_final, res2_final;
  if (orc_chi2_search_legacy(f_res2_mu, &lc, res2_min, chi2_target, &mu_final, &res2_final)) {
    /* the doubling did not reach the target (the reference would not return): report NaN */
    if (early) *early = 4;
    *mu = NAN, *chi2 = NAN;
    return x_unreg;
  }
  *mu = mu_final, *chi2 = res2_final / res2_min;
  if (mu_final == 0) { /* :528-529; save_results! would read the stale cache slot (:465): counted */
    if (early) *early = 3;
    return x_unreg;
  }
  return cache_solve(w, mu_final); /* :531 — a cache hit, f(mu_final) was the last solve */
}

/* =================== MDP  :700-747 =================== */
const double *orc_lsqnonneg_mdp(orc_reg_work *w, double delta, double *mu, double *chi2, int *early) {
  solve_unreg(w);
  const double *x_unreg = w->nnls->x;
  double res2_min = w->nnls->rnorm * w->nnls->rnorm;
  if (early) *early = 0;
  if (delta <= sqrt(res2_min)) { /* :708-711 (same stale-slot note as chi2) */
    if (early) *early = 1;
    *mu = 0.0, *chi2 = 1.0;
    return x_unreg;
  }
  double res2_max = 0.0;
  for (int i = 0; i < w->m; i++) res2_max = fma(w->b[i], w->b[i], res2_max);
  if (delta >= sqrt(res2_max)) { /* :713-718 */
    if (early) *early = 2;
    *mu = INFINITY, *chi2 = res2_max / res2_min;
    return w->cache->null_soln;
  }
  reset_cache(w->cache);
  root_ctx rc = {w, delta * delta, 1};
  double a, b, fa, fb;
  orc_bracket_root_monotonic(f_root, &rc, -4.0, 1.0, 1.5, +1, 6, &a, &b, &fa, &fb);
  double logmu_final, err_final;
  if (fa * fb < 0) {
    orc_brent_root(f_root, &rc, a, b, fa, fb, 0.0, 0.0, 1e-3 * (delta * delta), 100, &logmu_final, &err_final);
  } else {
    if (!isfinite(fa))
      logmu_final = b, err_final = fb;
    else if (!isfinite(fb))
      logmu_final = a, err_final = fa;
    else if (fabs(fa) < fabs(fb))
      logmu_final = a, err_final = fa;
    else
      logmu_final = b, err_final = fb;
  }
  if (isfinite(err_final)) {
    double mu_final = exp(logmu_final);
    double res2_final = delta * delta + err_final;
    const double *x_final = cache_solve(w, mu_final);
    *mu = mu_final, *chi2 = res2_final / res2_min;
    return x_final;
  }
  *mu = 0.0, *chi2 = 1.0 / res2_min;
  return x_unreg;
}

/* =================== GCV  :1136-1205, 1213-1229, 1262-1263, 1321-1329 =================== */
double orc_gcv_dof(int m, int n, const double *gamma, double lambda) {
  double dof = (double)((m - n) > 0 ? (m - n) : 0);
  double l2 = lambda * lambda;
  int mn = m < n ? m : n;
  for (int i = 0; i < mn; i++) {
    double g2 = gamma[i] * gamma[i];
    dof += l2 / (g2 + l2);
  }
  return dof;
}

/* Singular values by one-sided (Hestenes) Jacobi.  The reference calls LAPACK dgesdd_ with
 * jobz = 'N' (src/utils.jl:103-134), a third-party routine that is not part of the
 * reference sources; any accurate singular-value algorithm restates it.  tests/ pins this
 * against LAPACK gesdd through numpy.linalg.svd. */
void orc_svdvals(int m, int n, const double *A, int lda, double *S, double *work) {
  /* operate on a tall matrix G (r x c, r >= c): G = A or A^T */
  int r = m >= n ? m : n, c = m >= n ? n : m;
  double *G = work;
  if (m >= n) {
    for (int j = 0; j < n; j++)
      for (int i = 0; i < m; i++) G[i + (size_t)j * r] = A[i + (size_t)j * lda];
  } else {
    for (int j = 0; j < n; j++)
      for (int i = 0; i < m; i++) G[j + (size_t)i * r] = A[i + (size_t)j * lda];
  }
  for (int sweep = 0; sweep < 60; sweep++) {
    int rotated = 0;
    for (int p = 0; p < c - 1; p++) {
      for (int q = p + 1; q < c; q++) {
        double *gp = G + (size_t)p * r, *gq = G + (size_t)q * r;
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < r; i++) al += gp[i] * gp[i], be += gq[i] * gq[i], ga += gp[i] * gq[i];
        if (ga == 0.0 || fabs(ga) <= DBL_EPSILON * sqrt(al * be)) continue;
        rotated = 1;
        double zeta = (be - al) / (2 * ga);
        double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1 + zeta * zeta));
        double cs = 1 / sqrt(1 + t * t), sn = cs * t;
        for (int i = 0; i < r; i++) {
          double u = gp[i], v = gq[i];
          gp[i] = cs * u - sn * v;
          gq[i] = sn * u + cs * v;
        }
      }
    }
    if (!rotated) break;
  }
  for (int j = 0; j < c; j++) {
    double s = 0;
    const double *g = G + (size_t)j * r;
    for (int i = 0; i < r; i++) s += g[i] * g[i];
    S[j] = sqrt(s);
  }
  for (int i = 1; i < c; i++) { /* descending insertion sort */
    double v = S[i];
    int k = i - 1;
    while (k >= 0 && S[k] < v) S[k + 1] = S[k], k--;
    S[k + 1] = v;
  }
}

typedef struct {
  orc_reg_work *w;
  double gcv_low;
} gcv_ctx;

static double f_loggcv(double logmu, void *ctx) { /* log𝒢  :1150-1154 with gcv!  :1213-1229 */
  gcv_ctx *g = (gcv_ctx *)ctx;
  orc_reg_work *w = g->w;
  double mu = exp(logmu);
  cache_solve(w, mu);
  double res2 = tikh_resnorm_sq(cache_cur(w->cache));
  double dof = orc_gcv_dof(w->m, w->n, w->gamma, mu);
  double gcv = res2 / (dof * dof);
  gcv = fmax(gcv, g->gcv_low);
  return log(gcv);
}

const double *orc_lsqnonneg_gcv(orc_reg_work *w, double *mu, double *chi2) {
  orc_svdvals(w->m, w->n, w->A, w->m, w->gamma, w->svd_work); /* svdvals!(work)  :1143 */
  gcv_ctx g = {w, (DBL_EPSILON * DBL_EPSILON) / w->m};        /* gcv_lower_bound  :1262-1263 */
  reset_cache(w->cache);
  double logmu_final, logG;
  orc_brent_minimize(f_loggcv, &g, -8.0, 2.0, 0.0, 1e-4, 20, &logmu_final, &logG);
  double mu_final = exp(logmu_final);
  const double *x_final = cache_solve(w, mu_final);
  solve_unreg(w);
  double r2u = w->nnls->rnorm * w->nnls->rnorm;
  *mu = mu_final;
  *chi2 = tikh_resnorm_sq(cache_cur(w->cache)) / r2u;
  return x_final;
}
