/* ORACLE (test infrastructure) — the voxel pipeline: tables, flip-angle fit, regularised
 * NNLS, save_results!, T2part.  Follows src/T2mapSEcorr.jl, src/T2partSEcorr.jl,
 * src/types.jl, src/utils.jl. */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "decaes_oracle.h"

/* ------------------------------------------------------------------ grids */

/* range(a, b; length = n): Julia evaluates start + i*step in twice-precision arithmetic,
 * i.e. each element is (nearly always) the correctly rounded exact value.  x87 long double
 * plays that role here.  src/types.jl:109 (flip_angles), src/utils.jl:7 (inside logrange). */
void orc_linrange(double a, double b, int n, double *out) {
  if (n == 1) {
    out[0] = a;
    return;
  }
  for (int i = 0; i < n; i++) {
    long double v = ((long double)a * (long double)(n - 1 - i) + (long double)b * (long double)i) /
                    (long double)(n - 1);
    out[i] = (double)v;
  }
  out[0] = a, out[n - 1] = b;
}

/* logrange  src/utils.jl:7: exp.(range(log a, log b; length)) with exact end points */
void orc_logrange(double a, double b, int n, double *out) {
  orc_linrange(log(a), log(b), n, out);
  for (int i = 0; i < n; i++) out[i] = exp(out[i]);
  out[0] = a, out[n - 1] = b;
}

/* ------------------------------------------------------------------ option checks */

#define FAIL(...)                        \
  do {                                   \
    if (msg) snprintf(msg, msglen, __VA_ARGS__); \
    return DECAES_EINVAL;                \
  } while (0)

/* assertions of T2mapOptions  src/types.jl:28-84 */
int orc_validate_t2map_opts(const decaes_t2map_opts *o, char *msg, int msglen) {
  if (!(o->nx >= 1 && o->ny >= 1 && o->nz >= 1)) FAIL("MatrixSize must be a tuple of 3 positive integers");
  if (!(o->nTE >= 4)) FAIL("At least four echoes are required for T2 mapping, but nTE = %d.", o->nTE);
  if (!(o->TE > 0.0)) FAIL("Echo spacing must be positive, but TE = %g.", o->TE);
  if (!(o->nT2 >= 2)) FAIL("At least two T2 components are required for T2 mapping, but nT2 = %d.", o->nT2);
  if (!(0.0 < o->T2min && o->T2min < o->T2max)) FAIL("T2Range must a sorted 2-tuple of positive values");
  if (!(o->T1 > 0.0)) FAIL("T1 must be positive, but T1 = %g.", o->T1);
  if (!(o->Threshold >= 0.0 || o->Threshold == -INFINITY))
    FAIL("First echo signal threshold must be non-negative or -Inf");
  if (!(0.0 <= o->MinRefAngle && o->MinRefAngle <= 180.0)) FAIL("Minimum refocusing angle must be in the range [0, 180]");
  if (!(o->nRefAngles >= 2)) FAIL("nRefAngles must be at least 2, but nRefAngles = %d.", o->nRefAngles);
  if (!(2 <= o->nRefAnglesMin && o->nRefAnglesMin <= o->nRefAngles))
    FAIL("nRefAnglesMin must be in the range [2, nRefAngles]");
  if (!(o->reg >= DECAES_REG_NONE && o->reg <= DECAES_REG_MDP)) FAIL("Unrecognized regularization method: %d", o->reg);
  if (o->reg == DECAES_REG_CHI2 && !(o->Chi2Factor > 1.0)) FAIL("Chi2Factor must be greater than 1.0");
  if (o->reg == DECAES_REG_MDP && !(o->NoiseLevel > 0.0)) FAIL("Noise level must be positive");
  if (!(0.0 <= o->RefConAngle && o->RefConAngle <= 180.0)) FAIL("Refocusing control angle must be in the range [0, 180]");
  if (!isnan(o->SetFlipAngle) && !(0.0 <= o->SetFlipAngle && o->SetFlipAngle <= 180.0))
    FAIL("Fixed flip angle must be in the range [0, 180]");
  if (o->legacy && o->nRefAngles > ORC_SPLINE_MAX) {
    if (msg) snprintf(msg, msglen, "legacy = true supports at most %d refocusing angles", ORC_SPLINE_MAX);
    return DECAES_EUNSUPPORTED;
  }
  return DECAES_OK;
}

/* assertions of T2partOptions  src/types.jl:148-168 */
int orc_validate_t2part_opts(const decaes_t2part_opts *o, char *msg, int msglen) {
  if (!(o->nx >= 1 && o->ny >= 1 && o->nz >= 1)) FAIL("MatrixSize must be positive");
  if (!(o->nT2 >= 2)) FAIL("nT2 must be at least 2");
  if (!(0.0 < o->T2min && o->T2min < o->T2max)) FAIL("T2Range must be sorted and positive");
  if (!(o->SPWin_lo < o->SPWin_hi)) FAIL("SPWin must be sorted");
  if (!(o->MPWin_lo < o->MPWin_hi)) FAIL("MPWin must be sorted");
  if (!isnan(o->Sigmoid) && !(o->Sigmoid > 0)) FAIL("Sigmoid must be positive");
  return DECAES_OK;
}

/* ------------------------------------------------------------------ tables */

/* T2Maps(opts) table fields, src/T2mapSEcorr.jl:24-33; basis + Jacobian per angle,
 * src/T2mapSEcorr.jl:339-373, 387-407. */
int orc_setup_tables(const decaes_t2map_opts *o, double *echotimes, double *t2times, double *refangleset,
                     double *basis, double *dbasis) {
  const int nTE = o->nTE, nT2 = o->nT2;
  double *T2 = (double *)malloc(sizeof(double) * nT2);
  orc_logrange(o->T2min, o->T2max, nT2, T2);
  if (echotimes)
    for (int i = 0; i < nTE; i++) echotimes[i] = o->TE * (double)(i + 1);
  if (t2times) memcpy(t2times, T2, sizeof(double) * nT2);
  const int fixed = !isnan(o->SetFlipAngle);
  const int nA = fixed ? 1 : o->nRefAngles;
  double *ang = (double *)malloc(sizeof(double) * nA);
  if (fixed)
    ang[0] = o->SetFlipAngle;
  else
    orc_linrange(o->MinRefAngle, 180.0, nA, ang);
  if (refangleset) memcpy(refangleset, ang, sizeof(double) * nA);
  if (basis || dbasis) {
    double *work = (double *)malloc(sizeof(double) * 12 * nTE);
    double *dc = (double *)malloc(sizeof(double) * nTE), *ddc = (double *)malloc(sizeof(double) * nTE);
    for (int k = 0; k < nA; k++)
      for (int j = 0; j < nT2; j++) {
        if (o->RefConAngle == 180.0) {
          orc_epg_decay_curve_jac(nTE, ang[k], o->TE, T2[j], o->T1, dc, ddc, work);
        } else {
          /* beta != 180: general kernel :722-818, differentiated in forward mode like the reference's
           * ForwardDiff pass (src/T2mapSEcorr.jl:299-308) */
          orc_epg_decay_curve_beta_jac(nTE, ang[k], o->TE, T2[j], o->T1, o->RefConAngle, dc, ddc, work);
        }
        size_t off = ((size_t)k * nT2 + j) * nTE;
        if (basis) memcpy(basis + off, dc, sizeof(double) * nTE);
        if (dbasis) memcpy(dbasis + off, ddc, sizeof(double) * nTE);
      }
    free(work), free(dc), free(ddc);
  }
  free(T2), free(ang);
  return DECAES_OK;
}

/* ------------------------------------------------------------------ T2 parts */

typedef struct {
  int nT2;
  int sp_lo, sp_hi, mp_lo, mp_hi; /* 1-based inclusive ranges; empty if hi < lo */
  double *logT2;
  double *weights; /* NULL unless Sigmoid */
} part_tables;

static double erfinv_newton(double y) {
  double x = 0.0;
  for (int it = 0; it < 100; it++) {
    double e = erf(x) - y;
    double dx = e / (1.1283791670955126 * exp(-x * x));
    x -= dx;
    if (fabs(dx) < 1e-16 * fmax(1.0, fabs(x))) break;
  }
  return x;
}

/* thread_buffer_maker(::T2partOptions)  src/T2partSEcorr.jl:143-164 */
static int part_tables_init(part_tables *t, const decaes_t2part_opts *o) {
  const int n = o->nT2;
  t->nT2 = n;
  double *T2 = (double *)malloc(sizeof(double) * n);
  orc_logrange(o->T2min, o->T2max, n, T2);
  t->logT2 = (double *)malloc(sizeof(double) * n);
  for (int j = 0; j < n; j++) t->logT2[j] = log(T2[j]);
  int f;
  /* findfirst(>=(lo)) : findlast(<=(hi)) */
  t->sp_lo = t->sp_hi = t->mp_lo = t->mp_hi = 0;
  for (f = 0; f < n && !(T2[f] >= o->SPWin_lo); f++) {}
  t->sp_lo = f < n ? f + 1 : 0;
  for (f = n - 1; f >= 0 && !(T2[f] <= o->SPWin_hi); f--) {}
  t->sp_hi = f >= 0 ? f + 1 : 0;
  for (f = 0; f < n && !(T2[f] >= o->MPWin_lo); f++) {}
  t->mp_lo = f < n ? f + 1 : 0;
  for (f = n - 1; f >= 0 && !(T2[f] <= o->MPWin_hi); f--) {}
  t->mp_hi = f >= 0 ? f + 1 : 0;
  int ok = t->sp_lo && t->sp_hi && t->mp_lo && t->mp_hi; /* `nothing` bound => Julia throws */
  t->weights = NULL;
  if (!isnan(o->Sigmoid)) { /* sigmoid_weights  :155-164 */
    t->weights = (double *)malloc(sizeof(double) * n);
    double k = 0.1, T2_kperc = o->Sigmoid, T2_50 = o->SPWin_hi;
    double sigma = fabs(T2_kperc / (sqrt(2.0) * erfinv_newton(2 * k - 1)));
    for (int j = 0; j < n; j++) {
      double xx = (T2[j] - T2_50) / sigma;
      double wv = erfc(xx / sqrt(2.0)) / 2; /* normccdf  src/utils.jl:10 */
      t->weights[j] = wv <= DBL_EPSILON ? 0.0 : wv;
    }
  }
  free(T2);
  return ok ? DECAES_OK : DECAES_EINVAL;
}
static void part_tables_free(part_tables *t) { free(t->logT2), free(t->weights); }

/* voxelwise_T2_parts!  src/T2partSEcorr.jl:95-138.  dist has stride `ds` between bins. */
static void voxel_parts(const part_tables *t, const double *dist, int64_t ds, double *sfr, double *sgm,
                        double *mfr, double *mgm) {
  const int n = t->nT2;
  for (int j = 0; j < n; j++)
    if (isnan(dist[j * ds])) return;
  double S = 0, Ssp = 0, Smp = 0, dsp = 0, dmp = 0;
  for (int j = 0; j < n; j++) S += dist[j * ds];
  for (int j = t->sp_lo; j <= t->sp_hi; j++) {
    dsp += dist[(j - 1) * ds] * t->logT2[j - 1];
    Ssp += dist[(j - 1) * ds];
  }
  for (int j = t->mp_lo; j <= t->mp_hi; j++) {
    dmp += dist[(j - 1) * ds] * t->logT2[j - 1];
    Smp += dist[(j - 1) * ds];
  }
  if (S > 0) {
    if (t->weights) {
      double d = 0;
      for (int j = 0; j < n; j++) d = fma(dist[j * ds], t->weights[j], d);
      *sfr = d / S;
    } else {
      *sfr = Ssp / S;
    }
    *mfr = Smp / S;
  }
  if (Ssp > 0) *sgm = exp(dsp / Ssp);
  if (Smp > 0) *mgm = exp(dmp / Smp);
}

int orc_t2part(const double *dist, int64_t nvox, int64_t stride, const decaes_t2part_opts *part, double *sfr,
               double *sgm, double *mfr, double *mgm) {
  char msg[256];
  int rc = orc_validate_t2part_opts(part, msg, sizeof msg);
  if (rc) return rc;
  part_tables t;
  rc = part_tables_init(&t, part);
  if (rc) {
    part_tables_free(&t);
    return rc;
  }
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < nvox; v++) voxel_parts(&t, dist + v, stride, sfr + v, sgm + v, mfr + v, mgm + v);
  part_tables_free(&t);
  return DECAES_OK;
}

/* ------------------------------------------------------------------ voxel pipeline */

typedef struct {
  int nTE, nT2, nA;
  const decaes_t2map_opts *o;
  const double *T2, *logT2, *angles, *basis_set, *dbasis_set;
  double *decay_basis, *decay_data, *fit, *resid, *epg_work, *dAx, *Axb;
  orc_nnls_work *fa_nnls; /* NNLSDiscreteSurrogateSearch.nnls_work  src/splines.jl:1004 */
  orc_reg_work *reg;
  double flops;
  int64_t n_fa_solves;
} voxel_buf;

/* loss_with_grad!  src/splines.jl:1010-1041 */
static void fa_loss_grad(int I, double *u, double *du, void *ctx) {
  voxel_buf *vb = (voxel_buf *)ctx;
  const int m = vb->nTE, n = vb->nT2;
  const double *Ak = vb->basis_set + (size_t)(I - 1) * m * n;
  const double *dAk = vb->dbasis_set + (size_t)(I - 1) * m * n;
  orc_nnls_solve(vb->fa_nnls, Ak, m, vb->decay_data, m, n);
  vb->n_fa_solves++;
  *u = vb->fa_nnls->rnorm * vb->fa_nnls->rnorm;
  const double *x = vb->fa_nnls->x;
  int npos = 0;
  for (int i = 0; i < m; i++) vb->Axb[i] = 0.0, vb->dAx[i] = 0.0;
  for (int j = 0; j < n; j++)
    if (x[j] > 0) {
      npos++;
      for (int i = 0; i < m; i++) vb->Axb[i] = fma(x[j], Ak[i + (size_t)j * m], vb->Axb[i]);
    }
  for (int i = 0; i < m; i++) vb->Axb[i] -= vb->decay_data[i];
  for (int j = 0; j < n; j++)
    if (x[j] > 0)
      for (int i = 0; i < m; i++) vb->dAx[i] = fma(x[j], dAk[i + (size_t)j * m], vb->dAx[i]);
  double d = 0.0;
  for (int i = 0; i < m; i++) d = fma(vb->dAx[i], vb->Axb[i], d);
  *du = 2 * d;
  vb->flops += 4.0 * m * npos;
}

static void epg_basis_at(voxel_buf *vb, double alpha) { /* epg_decay_basis!  src/T2mapSEcorr.jl:275-283 */
  const decaes_t2map_opts *o = vb->o;
  const int m = vb->nTE;
  for (int j = 0; j < vb->nT2; j++) {
    if (o->RefConAngle == 180.0)
      orc_epg_decay_curve(m, alpha, o->TE, vb->T2[j], o->T1, vb->decay_basis + (size_t)j * m, vb->epg_work);
    else
      orc_epg_decay_curve_beta(m, alpha, o->TE, vb->T2[j], o->T1, o->RefConAngle,
                               vb->decay_basis + (size_t)j * m, vb->epg_work);
  }
  int h = m / 2;
  vb->flops += vb->nT2 * (13.0 * (h * (h + 1) - 2) + 2.0 * m);
}

int orc_t2map(const double *image, int64_t nvox, int64_t stride, const decaes_t2map_opts *o,
              const decaes_t2part_opts *part, const decaes_t2map_out *out, int nthreads, orc_stats *stats) {
  char msg[256];
  int rc = orc_validate_t2map_opts(o, msg, sizeof msg);
  if (rc) return rc;
  const int nTE = o->nTE, nT2 = o->nT2;
  const int fixed = !isnan(o->SetFlipAngle);
  const int nA = fixed ? 1 : o->nRefAngles;
  part_tables pt;
  memset(&pt, 0, sizeof pt);
  if (part) {
    rc = orc_validate_t2part_opts(part, msg, sizeof msg);
    if (!rc && part->nT2 != nT2) rc = DECAES_EINVAL;
    if (!rc) rc = part_tables_init(&pt, part);
    if (rc) {
      part_tables_free(&pt);
      return rc;
    }
  }
  double *T2 = (double *)malloc(sizeof(double) * nT2), *logT2 = (double *)malloc(sizeof(double) * nT2);
  double *angles = (double *)malloc(sizeof(double) * nA);
  double *basis_set = (double *)malloc(sizeof(double) * (size_t)nA * nTE * nT2);
  double *dbasis_set = (double *)malloc(sizeof(double) * (size_t)nA * nTE * nT2);
  orc_setup_tables(o, NULL, T2, angles, basis_set, dbasis_set);
  for (int j = 0; j < nT2; j++) logT2[j] = log(T2[j]);

#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  orc_stats tot;
  memset(&tot, 0, sizeof tot);
  tot.threads = nthreads;
  double t0 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif

#pragma omp parallel num_threads(nthreads)
  {
    /* thread_buffer_maker  src/T2mapSEcorr.jl:596-614 */
    voxel_buf vb;
    memset(&vb, 0, sizeof vb);
    vb.nTE = nTE, vb.nT2 = nT2, vb.nA = nA, vb.o = o;
    vb.T2 = T2, vb.logT2 = logT2, vb.angles = angles, vb.basis_set = basis_set, vb.dbasis_set = dbasis_set;
    vb.decay_basis = (double *)calloc((size_t)nTE * nT2, sizeof(double));
    vb.decay_data = (double *)calloc(nTE, sizeof(double));
    vb.fit = (double *)calloc(nTE, sizeof(double));
    vb.resid = (double *)calloc(nTE, sizeof(double));
    vb.epg_work = (double *)calloc(12 * (size_t)nTE, sizeof(double));
    vb.dAx = (double *)calloc(nTE, sizeof(double));
    vb.Axb = (double *)calloc(nTE, sizeof(double));
    vb.fa_nnls = orc_nnls_alloc(nTE, nT2);
    vb.reg = orc_reg_alloc(nTE, nT2);
    double *xs = (double *)calloc(nT2, sizeof(double));
    orc_stats loc;
    memset(&loc, 0, sizeof loc);
    if (fixed) memcpy(vb.decay_basis, basis_set, sizeof(double) * (size_t)nTE * nT2); /* :392-396 */

#pragma omp for schedule(dynamic, 64) /* default_blocksize() = 64  src/utils.jl:383 */
    for (int64_t v = 0; v < nvox; v++) {
      if (!(image[v] > o->Threshold)) continue; /* src/T2mapSEcorr.jl:177 */
      loc.voxels_processed++;

      /* voxelwise_T2_distribution!  src/T2mapSEcorr.jl:201-238 */
      double max_signal = 0.0;
      for (int i = 0; i < nTE; i++) {
        double bi = image[v + (int64_t)i * stride];
        max_signal = bi > max_signal ? bi : max_signal;
        vb.decay_data[i] = bi;
      }
      if (max_signal > 0)
        for (int i = 0; i < nTE; i++) vb.decay_data[i] /= max_signal;

      double alpha;
      if (o->alpha_provided) { /* :220-225 */
        alpha = out->alpha[v];
        epg_basis_at(&vb, alpha);
      } else if (fixed) {
        alpha = o->SetFlipAngle;
      } else { /* optimize_flip_angle!  :409-423 */
        double u_opt;
        if (o->legacy) /* CubicSplineSurrogate(...; legacy = true)  :401-402 */
          orc_surrogate_search_legacy(fa_loss_grad, &vb, angles, nA, o->nRefAnglesMin, o->nRefAngles, &alpha, &u_opt,
                                      NULL, NULL);
        else
          orc_surrogate_search(fa_loss_grad, &vb, angles, nA, o->nRefAnglesMin, o->nRefAngles, &alpha, &u_opt,
                               NULL, NULL);
        epg_basis_at(&vb, alpha);
      }

      /* T2_distribution!  :475-505 */
      orc_reg_bind(vb.reg, vb.decay_basis, vb.decay_data);
      const double *x = NULL;
      double mu = NAN, chi2 = NAN;
      int early = 0;
      switch (o->reg) {
        case DECAES_REG_NONE:
          mu = 0.0, chi2 = 1.0;
          x = orc_lsqnonneg(vb.reg);
          break;
        case DECAES_REG_LCURVE:
          x = orc_lsqnonneg_lcurve(vb.reg, &mu, &chi2);
          break;
        case DECAES_REG_GCV:
          x = orc_lsqnonneg_gcv(vb.reg, &mu, &chi2);
          break;
        case DECAES_REG_CHI2:
          x = o->legacy ? orc_lsqnonneg_chi2_legacy(vb.reg, o->Chi2Factor, &mu, &chi2, &early) /* :445, :495 */
                        : orc_lsqnonneg_chi2(vb.reg, o->Chi2Factor, &mu, &chi2, &early);
          break;
        case DECAES_REG_MDP: {
          double sigma = o->NoiseLevel / max_signal; /* :501-502 */
          double delta = sqrt((double)nTE) * sigma;
          x = orc_lsqnonneg_mdp(vb.reg, delta, &mu, &chi2, &early);
        } break;
      }
      if (early) loc.early_returns++;

      /* save_results!  :512-591 */
      for (int i = 0; i < nTE; i++) vb.decay_data[i] *= max_signal;
      for (int j = 0; j < nT2; j++) xs[j] = x[j] * max_signal;
      for (int i = 0; i < nTE; i++) { /* mul!(decay_curvefit, decay_basis, T2_dist) */
        double s = 0.0;
        for (int j = 0; j < nT2; j++) s = fma(vb.decay_basis[i + (size_t)j * nTE], xs[j], s);
        vb.fit[i] = s;
        vb.resid[i] = s - vb.decay_data[i];
      }
      double S = 0, R2 = 0, mean = 0, var = 0, dotl = 0;
      for (int j = 0; j < nT2; j++) S += xs[j];
      for (int i = 0; i < nTE; i++) R2 = fma(vb.resid[i], vb.resid[i], R2);
      for (int i = 0; i < nTE; i++) mean += vb.resid[i];
      mean /= nTE;
      for (int i = 0; i < nTE; i++) var = fma(vb.resid[i] - mean, vb.resid[i] - mean, var);
      double sigma_res = sqrt(var / (nTE - 1)); /* std(residuals), corrected */
      for (int j = 0; j < nT2; j++) dotl = fma(xs[j], logT2[j], dotl);
      double log_ggm = dotl / S;
      double l1p = 0;
      for (int j = 0; j < nT2; j++) {
        double dlt = logT2[j] - log_ggm;
        l1p = fma(dlt * dlt, xs[j], l1p);
      }
      l1p /= S;

      out->gdn[v] = S;
      out->ggm[v] = exp(log_ggm);
      out->gva[v] = expm1(l1p);
      out->fnr[v] = S / sqrt(R2 / (nTE - 1));
      out->snr[v] = max_signal / sigma_res;
      out->alpha[v] = alpha;
      for (int j = 0; j < nT2; j++) out->dist[v + (int64_t)j * stride] = xs[j];
      if (out->mu && out->chi2factor) out->mu[v] = mu, out->chi2factor[v] = chi2;
      if (out->resnorm) out->resnorm[v] = sqrt(R2);
      if (out->decaycurve)
        for (int i = 0; i < nTE; i++) out->decaycurve[v + (int64_t)i * stride] = vb.fit[i];
      if (out->decaybasis && !fixed)
        for (int k = 0; k < nTE * nT2; k++) out->decaybasis[v + (int64_t)k * stride] = vb.decay_basis[k];
      vb.flops += 2.0 * nTE * nT2 + 8.0 * nT2 + 6.0 * nTE;

      if (part && out->sfr && out->sgm && out->mfr && out->mgm) {
        voxel_parts(&pt, out->dist + v, stride, out->sfr + v, out->sgm + v, out->mfr + v, out->mgm + v);
        vb.flops += 3.0 * nT2;
      }
    }

    /* gather instrumentation */
    loc.nnls_unreg = vb.n_fa_solves + vb.reg->n_solves_unreg;
    {
      orc_nnls_work *ws[2 + 8];
      int nw = 0;
      ws[nw++] = vb.fa_nnls;
      ws[nw++] = vb.reg->nnls;
      extern orc_nnls_work *orc_reg_slot(orc_reg_work *, int);
      for (int i = 0; i < 8; i++) ws[nw++] = orc_reg_slot(vb.reg, i);
      for (int i = 0; i < nw; i++) {
        loc.cols_entered += ws[i]->n_enter, loc.cols_exited += ws[i]->n_exit, loc.cols_rejected += ws[i]->n_reject;
        loc.flops += ws[i]->flops;
      }
      loc.flops += vb.flops;
      loc.nnls_tikh = vb.reg->n_solves_tikh;
      loc.cache_hits = vb.reg->n_cache_hits;
    }
#pragma omp critical
    {
      tot.voxels_processed += loc.voxels_processed;
      tot.nnls_unreg += loc.nnls_unreg;
      tot.nnls_tikh += loc.nnls_tikh, tot.cache_hits += loc.cache_hits;
      tot.cols_entered += loc.cols_entered, tot.cols_exited += loc.cols_exited, tot.cols_rejected += loc.cols_rejected;
      tot.early_returns += loc.early_returns;
      tot.flops += loc.flops;
    }
    free(vb.decay_basis), free(vb.decay_data), free(vb.fit), free(vb.resid), free(vb.epg_work), free(vb.dAx),
        free(vb.Axb), free(xs);
    orc_nnls_free(vb.fa_nnls), orc_reg_free(vb.reg);
  }
#ifdef _OPENMP
  tot.seconds = omp_get_wtime() - t0;
#endif
  if (stats) *stats = tot;
  free(T2), free(logT2), free(angles), free(basis_set), free(dbasis_set);
  if (part) part_tables_free(&pt);
  return DECAES_OK;
}

/* ------------------------------------------------------------------ synthetic volume */

static inline uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline double urand(uint64_t seed, uint64_t vox, uint64_t k) { /* (0,1) */
  uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL * (vox + 1));
  h = mix64(h + 0x9E3779B97F4A7C15ULL * (k + 1));
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* mock_image  src/utils.jl:623-658, with a per-voxel flip angle alpha ~ U(120,180) instead of
 * the fixed 165 deg so that the flip-angle fit is exercised (documented deviation). */
void orc_mock_image(double *image, int64_t nvox, int64_t stride, int64_t first_voxel, int nTE, double TE,
                    double T1, double SNR, uint64_t seed) {
  const double sigma = pow(10.0, -SNR / 20);
#pragma omp parallel
  {
    double *work = (double *)malloc(sizeof(double) * 6 * nTE);
    double *d1 = (double *)malloc(sizeof(double) * nTE), *d2 = (double *)malloc(sizeof(double) * nTE);
#pragma omp for schedule(static)
    for (int64_t v = 0; v < nvox; v++) {
      uint64_t g = (uint64_t)(first_voxel + v);
      double sfr = 0.05 + (0.25 - 0.05) * urand(seed, g, 0);
      double T21 = 10e-3 + (20e-3 - 10e-3) * urand(seed, g, 1);
      double T22 = 50e-3 + (100e-3 - 50e-3) * urand(seed, g, 2);
      double alpha = 120.0 + 60.0 * urand(seed, g, 3);
      orc_epg_decay_curve(nTE, alpha, TE, T21, T1, d1, work);
      orc_epg_decay_curve(nTE, alpha, TE, T22, T1, d2, work);
      for (int k = 0; k < nTE; k++) {
        double m = sfr * d1[k] + (1 - sfr) * d2[k];
        double u1 = urand(seed, g, 4 + 2 * (uint64_t)k), u2 = urand(seed, g, 5 + 2 * (uint64_t)k);
        double r = sqrt(-2.0 * log(u1));
        double zR = sigma * r * cos(6.283185307179586 * u2), zI = sigma * r * sin(6.283185307179586 * u2);
        image[v + (int64_t)k * stride] = sqrt((m + zR) * (m + zR) + zI * zI);
      }
    }
    free(work), free(d1), free(d2);
  }
}
