/* ORACLE (test infrastructure) — the `legacy = true` algorithms of DECAES.jl.
 *
 *   spline_opt_legacy / spline_root_legacy      src/splines.jl:419-446
 *   make_spline (Dierckx.Spline1D, s = 0)       src/splines.jl:311-314
 *   CubicSplineSurrogate                        src/splines.jl:456-500
 *   chi2_search_from_minimum(...; legacy=true)  src/lsqnonneg.jl:595-636
 *
 * Third-party arithmetic: the interpolating spline is Dierckx.jl (Project.toml:38, compat "0.4, 0.5"), a
 * wrapper of P. Dierckx's FITPACK (netlib ddierckx).  FITPACK is not under the reference checkout,
 * so its published algorithm is restated here routine by routine: curfit/fpcurf for iopt = 0,
 * s = 0 (knot placement of the interpolating spline, row-by-row Givens reduction of the
 * observation matrix, back substitution fpback), fpbspl (de Boor-Cox), splev.  tests/ pin it
 * against the same Fortran routines as shipped in scipy.interpolate (splrep / splev).
 *
 * The sample abscissae `knots[1]:0.001:knots[end]` are a Julia StepRangeLen; `orc_jl_range`
 * restates Base's `(:)(start, step, stop)` for Float64 (base/twiceprecision.jl: `rat`,
 * rational lifting, `floatrange`) so that the oracle samples the same doubles. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "decaes_oracle.h"

/* ---------------- FITPACK ---------------- */

/* fpbspl: the k+1 non-zero B-splines of degree k at t[l] <= x < t[l+1] (l 1-based) */
static void fpbspl(const double *t, int k, double x, int l, double *h) {
  double hh[5];
  h[0] = 1.0;
  for (int j = 1; j <= k; j++) {
    for (int i = 0; i < j; i++) hh[i] = h[i];
    h[0] = 0.0;
    for (int i = 1; i <= j; i++) {
      int li = l + i, lj = li - j;
      double f = hh[i - 1] / (t[li - 1] - t[lj - 1]);
      h[i - 1] = h[i - 1] + f * (t[li - 1] - x);
      h[i] = f * (x - t[lj - 1]);
    }
  }
}

static void fpgivs(double piv, double *ww, double *c, double *s) {
  double store = fabs(piv), dd;
  if (store >= *ww) {
    double r = *ww / piv;
    dd = store * sqrt(1.0 + r * r);
  } else {
    double r = piv / *ww;
    dd = *ww * sqrt(1.0 + r * r);
  }
  *c = *ww / dd, *s = piv / dd, *ww = dd;
}

static void fprota(double c, double s, double *a, double *b) {
  double stor1 = *a, stor2 = *b;
  *b = c * stor2 + s * stor1;
  *a = c * stor1 - s * stor2;
}

/* curfit(iopt = 0, w = 1, xb = x[0], xe = x[m-1], s = 0): t[m+k+1], c[m] */
int orc_fitpack_interp(const double *x, const double *y, int m, int k, double *t, double *c) {
  if (k < 1 || k > 3 || m <= k || m > ORC_SPLINE_MAX) return -1;
  for (int i = 1; i < m; i++)
    if (!(x[i] > x[i - 1])) return -1;
  const int k1 = k + 1, n = m + k1, nk1 = n - k1;
  /* interior knots of the interpolating spline (fpcurf, "find the position of the interior knots") */
  const int mk1 = m - k1, k3 = k / 2;
  int i = k1 + 1, j = k3 + 2; /* 1-based */
  if (k3 * 2 != k) {
    for (int l = 0; l < mk1; l++, i++, j++) t[i - 1] = x[j - 1];
  } else {
    for (int l = 0; l < mk1; l++, i++, j++) t[i - 1] = (x[j - 1] + x[j - 2]) * 0.5;
  }
  for (int q = 0; q < k1; q++) t[q] = x[0], t[n - 1 - q] = x[m - 1];
  /* observation matrix, rotated row by row into the band triangle a[nk1][k1] */
  double a[ORC_SPLINE_MAX][4], z[ORC_SPLINE_MAX], h[6];
  memset(a, 0, sizeof a), memset(z, 0, sizeof z);
  int l = k1;
  for (int it = 0; it < m; it++) {
    double xi = x[it], yi = y[it];
    while (!(xi < t[l] || l == nk1)) l++; /* t(l+1) is t[l] 0-based */
    fpbspl(t, k, xi, l, h);
    int jj = l - k1; /* row index, 1-based after the increment below */
    for (int ii = 1; ii <= k1; ii++) {
      jj++;
      double piv = h[ii - 1];
      if (piv == 0.0) continue;
      double cs, sn;
      fpgivs(piv, &a[jj - 1][0], &cs, &sn);
      fprota(cs, sn, &yi, &z[jj - 1]);
      if (ii == k1) break;
      int i2 = 0;
      for (int i1 = ii + 1; i1 <= k1; i1++) {
        i2++;
        fprota(cs, sn, &h[i1 - 1], &a[jj - 1][i2]);
      }
    }
  }
  /* fpback */
  c[nk1 - 1] = z[nk1 - 1] / a[nk1 - 1][0];
  int ib = nk1 - 1;
  for (int jb = 2; jb <= nk1; jb++) {
    double store = z[ib - 1];
    int i1 = (jb <= k1 - 1) ? jb - 1 : k1 - 1;
    int mm = ib;
    for (int ll = 1; ll <= i1; ll++) {
      mm++;
      store = store - c[mm - 1] * a[ib - 1][ll];
    }
    c[ib - 1] = store / a[ib - 1][0];
    ib--;
  }
  return 0;
}

/* splev (e = 0, "extrapolate") for one argument; *l is the running knot interval (start at k+1) */
static double splev1(const double *t, int n, const double *c, int k, double arg, int *l) {
  const int k1 = k + 1, nk1 = n - k1;
  double h[6];
  while (arg < t[*l - 1] && *l != k1) (*l)--;
  while (!(arg < t[*l] || *l == nk1)) (*l)++;
  fpbspl(t, k, arg, *l, h);
  double sp = 0.0;
  int ll = *l - k1;
  for (int j = 0; j < k1; j++) sp = sp + c[ll + j] * h[j];
  return sp;
}

void orc_fitpack_splev(const double *t, int n, const double *c, int k, const double *x, int m, double *y) {
  int l = k + 1;
  for (int i = 0; i < m; i++) y[i] = splev1(t, n, c, k, x[i], &l);
}

/* ---------------- Julia's start:step:stop for Float64 ---------------- */

/* Base.rat (base/twiceprecision.jl): continued-fraction rational with |num|, |den| <= 2^24 */
static void jl_rat(double x, int64_t *num, int64_t *den) {
  double y = x;
  int64_t a = 1, d = 1, b = 0, c = 0;
  const double m = 16777216.0; /* maxintfloat(Float32) */
  while (fabs(y) <= m) {
    int64_t f = (int64_t)trunc(y);
    y -= (double)f;
    int64_t a2 = f * a + c, b2 = f * b + d;
    c = a, a = a2, d = b, b = b2;
    int64_t mx = llabs(a) > llabs(b) ? llabs(a) : llabs(b);
    if (!(mx <= (int64_t)m)) {
      *num = c, *den = d;
      return;
    }
    if ((double)a / (double)b == x) break;
    y = 1.0 / y;
  }
  *num = a, *den = b;
}

static int64_t gcd64(int64_t a, int64_t b) {
  a = llabs(a), b = llabs(b);
  while (b) {
    int64_t r = a % b;
    a = b, b = r;
  }
  return a;
}

static int isbetween(double a, double x, double b) { return (a <= x && x <= b) || (b <= x && x <= a); }

void orc_jl_range(double start, double step, double stop, orc_jl_range_t *r) {
  memset(r, 0, sizeof *r);
  r->start = start, r->step = step;
  int64_t step_n, step_d;
  jl_rat(step, &step_n, &step_d);
  if (step_d != 0 && (double)step_n / (double)step_d == step) {
    int64_t start_n, start_d, stop_n, stop_d;
    jl_rat(start, &start_n, &start_d);
    jl_rat(stop, &stop_n, &stop_d);
    if (start_d != 0 && stop_d != 0 && (double)start_n / (double)start_d == start &&
        (double)stop_n / (double)stop_d == stop) {
      int64_t den = start_d / gcd64(start_d, step_d) * step_d; /* lcm_unchecked */
      const double mi = 9007199254740992.0;                    /* maxintfloat(Float64) */
      if (den != 0 && fabs(start * (double)den) <= mi && fabs(step * (double)den) <= mi && den % start_d == 0 &&
          den % step_d == 0) {
        start_n = llround(start * (double)den); /* round(Int, x): ties are impossible for these exact products */
        step_n = llround(step * (double)den);
        int64_t len = (den * stop_n - stop_d * start_n + step_n * stop_d) / (step_n * stop_d);
        if (len < 0) len = 0;
        if (isbetween(start, start + (double)(len - 1) * step, stop + step / 2) &&
            !isbetween(start, start + (double)len * step, stop)) {
          r->rational = 1, r->start_n = start_n, r->step_n = step_n, r->den = den, r->len = len;
          return;
        }
      }
    }
  }
  /* fallback: start and step taken literally */
  double lf = (stop - start) / step;
  int64_t len;
  if (lf < 0)
    len = 0;
  else if (lf == 0)
    len = 1;
  else {
    len = llrint(lf) + 1; /* round(Int, lf): ties to even */
    double stop2 = start + (double)(len - 1) * step;
    len -= (start < stop && stop < stop2) + (start > stop && stop > stop2);
  }
  r->len = len;
}

/* element i (0-based).  Julia evaluates ref + i*step in twice precision and rounds once: for the
 * rational form that is the correctly rounded (start_n + i*step_n)/den, otherwise fma(i, step, start). */
double orc_jl_range_at(const orc_jl_range_t *r, int64_t i) {
  if (r->rational) return (double)(r->start_n + i * r->step_n) / (double)r->den;
  return fma((double)i, r->step, r->start);
}

/* ---------------- legacy spline searches ---------------- */

/* spline_opt_legacy  src/splines.jl:419-430 */
int orc_spline_opt_legacy(const double *X, const double *Y, int m, double *xo, double *yo) {
  int k = m - 1 < 3 ? m - 1 : 3;
  double t[ORC_SPLINE_MAX + 4], c[ORC_SPLINE_MAX];
  if (orc_fitpack_interp(X, Y, m, k, t, c)) return -1;
  orc_jl_range_t r;
  orc_jl_range(X[0], 0.001, X[m - 1], &r);
  int l = k + 1;
  double x = orc_jl_range_at(&r, 0), y = splev1(t, m + k + 1, c, k, x, &l);
  for (int64_t i = 1; i < r.len; i++) {
    double xi = orc_jl_range_at(&r, i), yi = splev1(t, m + k + 1, c, k, xi, &l);
    if (yi < y) x = xi, y = yi;
  }
  *xo = x, *yo = y;
  return 0;
}

/* spline_root_legacy  src/splines.jl:435-446 */
int orc_spline_root_legacy(const double *X, const double *Y, int m, double value, double *xo) {
  int k = m - 1 < 3 ? m - 1 : 3;
  double t[ORC_SPLINE_MAX + 4], c[ORC_SPLINE_MAX];
  if (orc_fitpack_interp(X, Y, m, k, t, c)) return -1;
  orc_jl_range_t r;
  orc_jl_range(X[0], 0.001, X[m - 1], &r);
  int l = k + 1;
  double x = orc_jl_range_at(&r, 0), y = fabs(splev1(t, m + k + 1, c, k, x, &l) - value);
  for (int64_t i = 1; i < r.len; i++) {
    double xi = orc_jl_range_at(&r, i), yi = fabs(splev1(t, m + k + 1, c, k, xi, &l) - value);
    if (yi < y) x = xi, y = yi;
  }
  *xo = x;
  return 0;
}

/* chi2_search_from_minimum(f, res2min, chi2fact; legacy = true)  src/lsqnonneg.jl:595-636:
 * mu doubles from 1e-3 until res2(mu) >= chi2fact * res2min, then the sampled spline root through
 * every (mu, res2) seen, mu = 0 included.  Returns 0, or -1 when the doubling does not terminate
 * within ORC_CHI2_LEGACY_MAXPTS - 1 = 21 steps (mu ~ 1049; the reference would loop on; the GPU uses the same cut). */
int orc_chi2_search_legacy(orc_fn1 f, void *ctx, double res2min, double chi2fact, double *mu_out, double *res2_out) {
  double mus[ORC_CHI2_LEGACY_MAXPTS], rs[ORC_CHI2_LEGACY_MAXPTS];
  int n = 0;
  mus[n] = 0.0, rs[n] = res2min, n++;
  double munew = 1e-3;
  while (1) {
    if (n >= ORC_CHI2_LEGACY_MAXPTS) return -1;
    double r = f(munew, ctx);
    mus[n] = munew, rs[n] = r, n++;
    if (r >= chi2fact * res2min) break;
    munew *= 2.0;
  }
  double mu;
  if (orc_spline_root_legacy(mus, rs, n, chi2fact * res2min, &mu)) return -1;
  *mu_out = mu;
  *res2_out = f(mu, ctx);
  return 0;
}
