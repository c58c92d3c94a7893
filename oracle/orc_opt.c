/* ORACLE (test infrastructure) — 1-D root finding / minimisation used by the regularisation
 * choosers.  Follows src/optimization.jl. */
#include <math.h>
#include "decaes_oracle.h"

static inline double sgn(double x) { return (x > 0) - (x < 0); } /* Base.sign for finite x */

/* secant_step  src/optimization.jl:177-180 */
static double secant_step(double a, double b, double fa, double fb) { return a - fa * (b - a) / (fb - fa); }

/* inverse_quadratic_step  src/optimization.jl:182-189 */
static double inverse_quadratic_step(double a, double b, double c, double fa, double fb, double fc) {
  double s = 0.0;
  s += a * fb * fc / (fa - fb) / (fa - fc);
  s += b * fa * fc / (fb - fa) / (fb - fc);
  s += c * fa * fb / (fc - fa) / (fc - fb);
  return s;
}

/* brent_root  src/optimization.jl:71-128 */
void orc_brent_root(orc_fn1 f, void *ctx, double x0, double x1, double fx0, double fx1, double xatol,
                    double xrtol, double ftol, int maxiters, double *xo, double *fxo) {
  if (fx0 == 0) {
    *xo = x0, *fxo = fx0;
    return;
  }
  if (fx1 == 0) {
    *xo = x1, *fxo = fx1;
    return;
  }
  double a = x0, b = x1, fa = fx0, fb = fx1;
  if (fabs(fa) < fabs(fb)) {
    double t = a;
    a = b, b = t;
    t = fa, fa = fb, fb = t;
  }
  double c = x0, d = x0, fc = fx0;
  int mflag = 1;
  for (int iter = 1; iter <= maxiters; iter++) {
    if (fabs(b - a) <= 2 * (xatol + xrtol * fabs(b))) break;
    double s = inverse_quadratic_step(a, b, c, fa, fb, fc);
    if (isnan(s) || isinf(s)) s = secant_step(a, b, fa, fb);
    double u = (3 * a + b) / 4, v = b;
    if (u > v) {
      double t = u;
      u = v, v = t;
    }
    double tol = fmax(xatol, xrtol * fmax(fabs(b), fmax(fabs(c), fabs(d))));
    if (!(u < s && s < v) || (mflag && fabs(s - b) >= fabs(b - c) / 2) ||
        (!mflag && fabs(s - b) >= fabs(b - c) / 2) || (mflag && fabs(b - c) <= tol) ||
        (!mflag && fabs(c - d) <= tol)) {
      s = (a + b) / 2;
      mflag = 1;
    } else {
      mflag = 0;
    }
    double fs = f(s, ctx);
    if (fs == 0) {
      *xo = s, *fxo = fs;
      return;
    }
    if (isnan(fs) || isinf(fs)) break; /* return (b, fb) */
    if (fabs(fs) <= ftol) {
      *xo = s, *fxo = fs;
      return;
    }
    d = c; /* c, fc, d = b, fb, c */
    c = b, fc = fb;
    if (sgn(fa) * sgn(fs) < 0) {
      b = s, fb = fs;
    } else {
      a = s, fa = fs;
    }
    if (fabs(fa) < fabs(fb)) {
      double t = a;
      a = b, b = t;
      t = fa, fa = fb, fb = t;
    }
  }
  *xo = b, *fxo = fb;
}

/* bracket_root_monotonic  src/optimization.jl:191-219 */
void orc_bracket_root_monotonic(orc_fn1 f, void *ctx, double a, double delta, double dilate, int mono,
                                int maxiters, double *oa, double *ob, double *ofa, double *ofb) {
  double fa = f(a, ctx);
  if (!isfinite(fa)) {
    *oa = a, *ob = a, *ofa = NAN, *ofb = NAN;
    return;
  }
  if (fa == 0) {
    *oa = a, *ob = a, *ofa = fa, *ofb = fa;
    return;
  }
  double sgn_d = sgn((double)mono) * sgn(fa);
  double b = a - sgn_d * delta;
  double fb = f(b, ctx);
  if (!isfinite(fb)) {
    *oa = a, *ob = a, *ofa = fa, *ofb = fa;
    return;
  }
  if (fb == 0) {
    *oa = b, *ob = b, *ofa = fb, *ofb = fb;
    return;
  }
  delta *= dilate;
  int cnt = 0;
  while (fa * fb > 0 && cnt < maxiters) {
    a = b, fa = fb;
    b = a - sgn_d * delta;
    fb = f(b, ctx);
    if (!isfinite(fb)) {
      *oa = a, *ob = a, *ofa = fa, *ofb = fa;
      return;
    }
    if (fb == 0) {
      *oa = b, *ob = b, *ofa = fb, *ofb = fb;
      return;
    }
    delta *= dilate;
    cnt += 1;
  }
  if (a < b) {
    *oa = a, *ob = b, *ofa = fa, *ofb = fb;
  } else {
    *oa = b, *ob = a, *ofa = fb, *ofb = fa;
  }
}

/* brent_minimize  src/optimization.jl:319-413 */
void orc_brent_minimize(orc_fn1 f, void *ctx, double x1, double x2, double xrtol, double xatol,
                        int maxiters, double *xo, double *yo) {
  const double phi = 1.618033988749895; /* Float64(Base.MathConstants.golden) */
  const double alpha = 2 - phi;
  double x = x1 + alpha * (x2 - x1);
  double y = f(x, ctx);
  double dx_old = 0.0, dx = 0.0;
  double x_older = x, x_old = x;
  double y_older = y, y_old = y;
  int iter = 0;
  while (iter < maxiters) {
    double p = 0.0, q = 0.0;
    double xm = (x2 + x1) / 2;
    double dx_tol = xatol + xrtol * fabs(x);
    if (fabs(x - xm) + (x2 - x1) / 2 <= 2 * dx_tol) break;
    iter += 1;
    if (fabs(dx_old) > dx_tol) {
      double r = (x - x_old) * (y - y_older);
      q = (x - x_older) * (y - y_old);
      p = (x - x_older) * q - (x - x_old) * r;
      q = 2 * (q - r);
      if (q > 0)
        p = -p;
      else
        q = -q;
    }
    if (fabs(p) < fabs(q * dx_old / 2) && p < q * (x2 - x) && p < q * (x - x1)) {
      dx_old = dx;
      dx = p / q;
      double x_tmp = x + dx;
      if ((x_tmp - x1) < 2 * dx_tol || (x2 - x_tmp) < 2 * dx_tol) dx = (x < xm) ? dx_tol : -dx_tol;
    } else {
      dx_old = (x < xm) ? x2 - x : x1 - x;
      dx = alpha * dx_old;
    }
    double x_new;
    if (fabs(dx) >= dx_tol)
      x_new = x + dx;
    else
      x_new = x + ((dx > 0) ? dx_tol : -dx_tol);
    double y_new = f(x_new, ctx);
    if (y_new < y) {
      if (x_new < x)
        x2 = x;
      else
        x1 = x;
      x_older = x_old, x_old = x, x = x_new;
      y_older = y_old, y_old = y, y = y_new;
    } else {
      if (x_new < x)
        x1 = x_new;
      else
        x2 = x_new;
      if (y_new <= y_old || x_old == x) {
        x_older = x_old, x_old = x_new;
        y_older = y_old, y_old = y_new;
      } else if (y_new <= y_older || x_older == x || x_older == x_old) {
        x_older = x_new;
        y_older = y_new;
      }
    }
  }
  *xo = x, *yo = y;
}
