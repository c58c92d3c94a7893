/* ORACLE (test infrastructure) — discrete surrogate search for the refocusing flip angle.
 * Follows src/splines.jl (CubicHermiteInterpolator :53-110, CubicHermiteSplineSurrogate
 * :504-566, BoundingBox :664-699, DiscreteSurrogateSearcher :705-744, bisection_search
 * :750-843) specialised to D = 1. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "decaes_oracle.h"

/* CubicHermiteInterpolator constructor :62-69 and minimize :88-110 */
void orc_hermite_minimize(double a, double b, double u0, double u1, double m0, double m1, double *xo,
                          double *uo) {
  double r = (b - a) / 2;
  m0 = r * m0, m1 = r * m1;
  double du = u1 - u0, dm = m1 - m0;
  double su = u1 + u0, sm = m1 + m0;
  double c0 = su / 2 - dm / 4, c1 = (3 * du - sm) / 4, c2 = dm / 4, c3 = (sm - du) / 4;
  double xend = (u0 < u1) ? a : b, uend = (u0 < u1) ? u0 : u1;

  double D = 3 * (u0 - u1);
  double th = D / 2 + (m0 + m1);
  double g = th * th - m0 * m1;
  g = g > 0 ? -sqrt(g) : 0.0;
  double p = -(D + (m0 + m1));
  double q = 2 * g + (m0 - m1);
  if (fabs(p) < fabs(q)) {
    double t = p / q;
    double y = fma(t, fma(t, fma(t, c3, c2), c1), c0); /* evalpoly (Horner with muladd) */
    if (y < uend) {
      /* todomain(t, dom, Val(:nearest))  :81-86 */
      double c = (a + b) / 2, rr = (b - a) / 2;
      double x = fma(rr, t, c);
      x = x < a ? a : (x > b ? b : x);
      *xo = x, *uo = y;
      return;
    }
  }
  *xo = xend, *uo = uend;
}

typedef struct {
  orc_fg_fn fg;
  void *ctx;
  const double *grid;
  int n;
  /* CubicHermiteSplineSurrogate :504-512 */
  unsigned char *seen_s;
  double *u, *du;
  int *idx; /* sorted probed indices (1-based) */
  int npts;
  /* DiscreteSurrogateSearcher :705-709 */
  unsigned char *seen;
  int numeval;
  int *order;
  int norder;
  int legacy; /* CubicSplineSurrogate(...; legacy = true) :456-500 */
} search_t;

typedef struct {
  int lo, hi;
} box_t;

/* update!(surr, I) :526-534 with insertsorted! (src/utils.jl:42-49) */
static void surr_update(search_t *s, int I) {
  if (s->seen_s[I - 1]) return;
  double u, du;
  s->fg(I, &u, &du, s->ctx);
  s->seen_s[I - 1] = 1;
  s->u[I - 1] = u, s->du[I - 1] = du;
  s->npts += 1;
  int val = I;
  for (int i = 0; i < s->npts - 1; i++) {
    int xi = s->idx[i];
    int lo = xi < val ? xi : val, hi = xi < val ? val : xi;
    s->idx[i] = lo, val = hi;
  }
  s->idx[s->npts - 1] = val;
  if (s->order) s->order[s->norder] = I;
  s->norder++;
}

/* update!(surr, state, I; maxeval) :736-744 — returns 1 when the evaluation budget is spent */
static int state_update(search_t *s, int I, int maxeval) {
  if (s->numeval >= maxeval) return 1;
  if (s->seen[I - 1]) return 0;
  surr_update(s, I);
  s->seen[I - 1] = 1;
  s->numeval += 1;
  return s->numeval >= maxeval;
}

static int is_evaluated(const search_t *s, box_t b) { return s->seen[b.lo - 1] && s->seen[b.hi - 1]; } /* :817-820 */
static int box_width(box_t b) { return abs(b.hi - b.lo); }
static void bisect(box_t b, box_t *l, box_t *r) { /* :688-697 */
  int mid = (b.lo + b.hi) / 2;
  l->lo = b.lo, l->hi = mid;
  r->lo = mid, r->hi = b.hi;
}

/* evaluate_box! :802-815; x == NULL -> corners in (lo, hi) order (:676-681), otherwise sorted
 * by squared distance to x, stable (:833-837) */
static void evaluate_box(search_t *s, box_t b, const double *x, int maxeval) {
  int cs[2] = {b.lo, b.hi};
  if (x) {
    double d0 = (s->grid[b.lo - 1] - *x) * (s->grid[b.lo - 1] - *x);
    double d1 = (s->grid[b.hi - 1] - *x) * (s->grid[b.hi - 1] - *x);
    if (d1 < d0) cs[0] = b.hi, cs[1] = b.lo;
  }
  for (int k = 0; k < 2; k++) {
    if (is_evaluated(s, b)) break;
    if (state_update(s, cs[k], maxeval)) break;
  }
}

/* initialize!(surr, state, box, depth) :726-734 */
static void initialize_rec(search_t *s, box_t b, int depth, int mineval, int maxeval) {
  if (depth <= 0) return;
  evaluate_box(s, b, NULL, maxeval);
  if (s->numeval >= mineval) return;
  box_t l, r;
  bisect(b, &l, &r);
  initialize_rec(s, l, depth - 1, mineval, maxeval);
  initialize_rec(s, r, depth - 1, mineval, maxeval);
}

/* suggest_point :544-566 */
static void suggest_point(const search_t *s, double *xo, double *uo) {
  if (s->legacy) { /* CubicSplineSurrogate suggest_point :492-500 -> spline_opt_legacy :419-430 */
    double ps[ORC_SPLINE_MAX], us[ORC_SPLINE_MAX];
    for (int i = 0; i < s->npts; i++) ps[i] = s->grid[s->idx[i] - 1], us[i] = s->u[s->idx[i] - 1];
    if (orc_spline_opt_legacy(ps, us, s->npts, xo, uo)) *xo = NAN, *uo = NAN;
    return;
  }
  int I0 = s->idx[0];
  double plast = s->grid[I0 - 1], ulast = s->u[I0 - 1], dlast = s->du[I0 - 1];
  double p = plast, u = ulast;
  for (int i = 1; i < s->npts; i++) {
    int I = s->idx[i];
    double pc = s->grid[I - 1], uc = s->u[I - 1], dc = s->du[I - 1];
    double x_, u_;
    orc_hermite_minimize(plast, pc, ulast, uc, dlast, dc, &x_, &u_);
    if (u_ < u) p = x_, u = u_;
    plast = pc, ulast = uc, dlast = dc;
  }
  *xo = p, *uo = u;
}

/* minimal_bounding_box :778-800 */
static box_t minimal_bounding_box(const search_t *s, double x) {
  box_t b = {1, s->n};
  while (1) {
    box_t l, r;
    bisect(b, &l, &r);
    int in_left = (s->grid[l.lo - 1] <= x) && (x <= s->grid[l.hi - 1]); /* contains :839-843 */
    box_t pick = in_left ? l : r;
    if (!is_evaluated(s, pick) || !(box_width(pick) > 1)) return pick;
    b = pick;
  }
}

static void search_impl(orc_fg_fn fg, void *ctx, const double *grid, int ngrid, int mineval, int maxeval,
                        double *x_opt, double *u_opt, int *order, int *norder, int legacy) {
  search_t s;
  memset(&s, 0, sizeof(s));
  s.fg = fg, s.ctx = ctx, s.grid = grid, s.n = ngrid, s.order = order, s.legacy = legacy;
  s.seen_s = (unsigned char *)calloc(ngrid, 1);
  s.seen = (unsigned char *)calloc(ngrid, 1);
  s.u = (double *)malloc(sizeof(double) * ngrid);
  s.du = (double *)malloc(sizeof(double) * ngrid);
  s.idx = (int *)calloc(ngrid, sizeof(int));
  for (int i = 0; i < ngrid; i++) s.u[i] = NAN, s.du[i] = NAN;

  /* DiscreteSurrogateSearcher(...) -> initialize! :710-724 */
  box_t root = {1, ngrid};
  for (int depth = 1; depth <= mineval; depth++) {
    initialize_rec(&s, root, depth, mineval, maxeval);
    if (s.numeval >= mineval) break;
  }

  /* bisection_search :750-775 */
  double x, u;
  suggest_point(&s, &x, &u);
  while (1) {
    box_t b = minimal_bounding_box(&s, x);
    evaluate_box(&s, b, &x, maxeval);
    suggest_point(&s, &x, &u);
    if (s.numeval >= maxeval || box_width(b) <= 1) break; /* converged :822-825 */
  }
  *x_opt = x, *u_opt = u;
  if (norder) *norder = s.norder;
  free(s.seen_s), free(s.seen), free(s.u), free(s.du), free(s.idx);
}

void orc_surrogate_search(orc_fg_fn fg, void *ctx, const double *grid, int ngrid, int mineval,
                          int maxeval, double *x_opt, double *u_opt, int *order, int *norder) {
  search_impl(fg, ctx, grid, ngrid, mineval, maxeval, x_opt, u_opt, order, norder, 0);
}

void orc_surrogate_search_legacy(orc_fg_fn fg, void *ctx, const double *grid, int ngrid, int mineval,
                                 int maxeval, double *x_opt, double *u_opt, int *order, int *norder) {
  search_impl(fg, ctx, grid, ngrid, mineval, maxeval, x_opt, u_opt, order, norder, 1);
}
