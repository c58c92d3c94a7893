/* ORACLE (test infrastructure) — Lawson–Hanson NNLS, plain and Tikhonov-structured.
 * Follows src/NNLS.jl and the warm-started drivers in src/lsqnonneg.jl:30-164.
 * The whole NNLS module of the reference is wrapped in @muladd (src/NNLS.jl:47): every
 * `a + b*c` there is an fma here.  1-based indices of the reference are kept via macros. */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "decaes_oracle.h"

#define A_(i, j) A[((size_t)(i)-1) + ((size_t)(j)-1) * (size_t)lda]
#define B_(i) b[(i)-1]
#define X_(i) x[(i)-1]
#define W_(i) w[(i)-1]
#define ZZ_(i) zz[(i)-1]
#define IDX_(i) idx[(i)-1]

orc_nnls_work *orc_nnls_alloc(int M, int N) {
  orc_nnls_work *w = (orc_nnls_work *)calloc(1, sizeof(*w));
  w->M = M, w->N = N;
  w->A = (double *)calloc((size_t)M * N, sizeof(double));
  w->b = (double *)calloc(M, sizeof(double));
  w->x = (double *)calloc(N, sizeof(double));
  w->w = (double *)calloc(N, sizeof(double));
  w->zz = (double *)calloc(M, sizeof(double));
  w->idx = (int *)calloc(N, sizeof(int));
  w->invidx = (int *)calloc(N, sizeof(int));
  w->diag = (unsigned char *)calloc(N, 1);
  return w;
}

void orc_nnls_free(orc_nnls_work *w) {
  if (!w) return;
  free(w->A), free(w->b), free(w->x), free(w->w), free(w->zz), free(w->idx), free(w->invidx),
      free(w->diag), free(w);
}

/* Base.Math.hypot (Borges' corrected algorithm with hardware fma), used by
 * orthogonal_rotmat (src/NNLS.jl:486-491). */
double orc_hypot(double x, double y) {
  if (isinf(x) || isinf(y)) return INFINITY;
  if (isnan(x) || isnan(y)) return NAN;
  double ax = fabs(x), ay = fabs(y);
  if (ay > ax) {
    double t = ax;
    ax = ay, ay = t;
  }
  if (ay <= ax * sqrt(DBL_EPSILON / 2)) return ax;
  double scale = DBL_EPSILON * sqrt(DBL_MIN);
  if (ax > sqrt(DBL_MAX / 2)) {
    ax *= scale, ay *= scale, scale = 1.0 / scale;
  } else if (ay < sqrt(DBL_MIN)) {
    ax /= scale, ay /= scale;
  } else {
    scale = 1.0;
  }
  double h = sqrt(fma(ax, ax, ay * ay));
  double hsq = h * h, axsq = ax * ax;
  h -= (fma(-ay, ay, hsq - axsq) + fma(h, h, -hsq) - fma(ax, ax, -axsq)) / (2 * h);
  return h * scale;
}

/* construct_apply_householder!  src/NNLS.jl:259-336.  Returns tau >= 0 if the column is
 * accepted (A, b updated, column swapped into position ip), -1 if rejected (nothing touched). */
static double construct_apply_householder(double *A, int lda, double *b, int ip, int jp, int m) {
  if (ip > m) return 0.0;
  double alpha = A_(ip, jp);
  double xnorm = 0.0;
  ORC_SIMD_REDUCE(xnorm) /* @simd, src/NNLS.jl:278 */
  for (int i = ip; i <= m; i++) xnorm = fma(A_(i, jp), A_(i, jp), xnorm);
  xnorm = sqrt(xnorm);
  if (xnorm == 0.0) return -1.0;

  double beta = copysign(xnorm, alpha);
  alpha = alpha + beta;
  double tau = alpha / beta;

  double sm = B_(ip);
  ORC_SIMD_REDUCE(sm) /* @simd, src/NNLS.jl:292 */
  for (int i = ip + 1; i <= m; i++) sm = fma(B_(i), A_(i, jp) / alpha, sm);
  sm *= -tau;

  double A1 = -beta;
  double b1 = B_(ip) + sm;

  if (b1 / A1 > 0) {
    if (ip < m) {
      if (ip != jp) {
        for (int i = 1; i <= ip - 1; i++) {
          double t = A_(i, ip);
          A_(i, ip) = A_(i, jp), A_(i, jp) = t;
        }
        B_(ip) = b1;
        {
          double t = A_(ip, ip);
          A_(ip, ip) = A1, A_(ip, jp) = t;
        }
        for (int i = ip + 1; i <= m; i++) {
          double Aij = A_(i, jp) / alpha;
          double t = A_(i, ip);
          A_(i, ip) = Aij, A_(i, jp) = t;
          B_(i) = fma(sm, Aij, B_(i));
        }
      } else {
        B_(ip) = b1;
        A_(ip, ip) = A1;
        for (int i = ip + 1; i <= m; i++) {
          double Aii = A_(i, ip) / alpha;
          A_(i, ip) = Aii;
          B_(i) = fma(sm, Aii, B_(i));
        }
      }
    } else {
      tau = 0.0;
      if (ip != jp) {
        for (int i = 1; i <= m; i++) {
          double t = A_(i, ip);
          A_(i, ip) = A_(i, jp), A_(i, jp) = t;
        }
      }
    }
    return tau;
  }
  return -1.0;
}

/* apply_householder_dual!  src/NNLS.jl:379-436.  The reference unrolls four columns at a
 * time; per column the arithmetic is the one below. */
static void apply_householder_dual(double *A, int lda, int n, double *w, const double *b, double tau,
                                   int j1, int m1) {
  if (j1 >= m1) return;
  double aii = A_(j1, j1);
  A_(j1, j1) = 1.0;
  for (int j = j1 + 1; j <= n; j++) {
    double sm = 0.0;
    ORC_SIMD_REDUCE(sm) /* @simd, src/NNLS.jl:398, 418 */
    for (int i = j1; i <= m1; i++) sm = fma(A_(i, j), A_(i, j1), sm);
    sm *= -tau;
    double wj = 0.0;
    A_(j1, j) = A_(j1, j) + sm;
    ORC_SIMD_REDUCE(wj) /* @simd, src/NNLS.jl:405, 424 */
    for (int i = j1 + 1; i <= m1; i++) {
      double Aij = fma(sm, A_(i, j1), A_(i, j));
      wj = fma(Aij, B_(i), wj);
      A_(i, j) = Aij;
    }
    W_(j) = wj;
  }
  A_(j1, j1) = aii;
}

/* compute_dual!  src/NNLS.jl:441-470 */
static void compute_dual(double *w, const double *A, int lda, int n, const double *b, int j1, int m1) {
  for (int j = j1; j <= n; j++) {
    double sm = 0.0;
    ORC_SIMD_REDUCE(sm) /* @simd, src/NNLS.jl:452, 462 */
    for (int i = j1; i <= m1; i++) sm = fma(A_(i, j), B_(i), sm);
    W_(j) = sm;
  }
}

/* solve_triangular_system!  src/NNLS.jl:506-539 */
void orc_solve_triangular(double *z, const double *A, int lda, int n, int transp) {
  if (!transp) {
    for (int j = n; j >= 1; j--) {
      double zi = -z[j - 1] / A_(j, j);
      for (int i = 1; i <= j - 1; i++) z[i - 1] = fma(A_(i, j), zi, z[i - 1]);
      z[j - 1] = -zi;
    }
  } else {
    for (int j = 1; j <= n; j++) {
      double z1 = z[j - 1];
      for (int l = 1; l <= j - 1; l++) z1 = fma(-A_(l, j), z[l - 1], z1);
      z1 /= A_(j, j);
      z[j - 1] = z1;
    }
  }
}

/* largest_positive_dual  src/NNLS.jl:541-554 */
static double largest_positive_dual(const double *w, int n, int j1, int *jmax_out) {
  double wmax = 0.0;
  int jmax = 0;
  for (int j = j1; j <= n; j++) {
    if (W_(j) > wmax) {
      wmax = W_(j);
      jmax = j;
    }
  }
  *jmax_out = jmax;
  return wmax;
}

/* unsafe_nnls!(work)            src/NNLS.jl:605-825   (lambda < 0: plain problem)
 * unsafe_nnls!(work, lambda)    src/NNLS.jl:827-1061  (lambda >= 0: A = [A0; lambda I], rows
 *                                                      of lambda I introduced lazily)
 * The two reference functions differ only in the lines guarded by `tikh` below. */
static void unsafe_nnls(orc_nnls_work *wk, int Mrows, int mrows, int init_dual, int tikh, double lambda) {
  double *A = wk->A, *b = wk->b, *x = wk->x, *w = wk->w, *zz = wk->zz;
  int *idx = wk->idx, *invidx = wk->invidx;
  unsigned char *diag = wk->diag;
  const int lda = wk->M, n = wk->N;
  const int M = Mrows;
  int m = mrows;
  const int max_iter = 3 * n;

  if (init_dual) {
    for (int j = 1; j <= n; j++) W_(j) = 0.0;
    compute_dual(w, A, lda, n, b, 1, m);
  }

  int nsetp = 0, iter = 0, terminated = 0;
  wk->mode = 0;

  while (1) {
    if (tikh ? (nsetp >= n) : (nsetp >= n || nsetp >= m)) {
      terminated = 1;
      break;
    }
    int jmax = nsetp;
    double tau = 0.0;
    while (1) {
      double wmax = largest_positive_dual(w, n, nsetp + 1, &jmax);
      if (wmax <= 0) {
        terminated = 1;
        break;
      }
      if (tikh && !diag[IDX_(jmax) - 1]) A_(m + 1, jmax) = lambda; /* :869-871 */
      tau = construct_apply_householder(A, lda, b, nsetp + 1, jmax, tikh ? (m + 1 < M ? m + 1 : M) : m);
      if (tau >= 0) break;
      W_(jmax) = 0.0;
      wk->n_reject++;
      if (tikh && m < M) A_(m + 1, jmax) = 0.0; /* :885-887 */
    }
    if (terminated) break;

    if (tikh && !diag[IDX_(jmax) - 1]) { /* :900-903 */
      m = (m + 1 < M) ? m + 1 : M;
      diag[IDX_(jmax) - 1] = 1;
    }
    nsetp += 1;
    wk->n_enter++;
    {
      double r = (double)(m - nsetp + 1), k = (double)nsetp;
      wk->flops += 4 * r + (6 * r - 2) * (n - k) + k * k;
    }
    {
      int t = IDX_(nsetp);
      IDX_(nsetp) = IDX_(jmax), IDX_(jmax) = t;
    }
    if (nsetp < n) apply_householder_dual(A, lda, n, w, b, tau, nsetp, m);
    for (int i = nsetp + 1; i <= m; i++) A_(i, nsetp) = 0.0;
    W_(nsetp) = 0.0;

    for (int i = 1; i <= nsetp; i++) ZZ_(i) = B_(i);
    orc_solve_triangular(zz, A, lda, nsetp, 0);

    int dual_flag = 0;
    while (1) {
      iter += 1;
      if (iter > max_iter) {
        wk->mode = 1;
        terminated = 1;
        break;
      }
      int imv = nsetp;
      double alpha = 2.0;
      for (int i = 1; i <= nsetp; i++) {
        if (ZZ_(i) <= 0) {
          double xi = X_(IDX_(i));
          double t = -xi / (ZZ_(i) - xi);
          if (alpha > t) {
            imv = i;
            alpha = t;
          }
        }
      }
      if (alpha == 2.0) break;
      dual_flag = 1;

      for (int i = 1; i <= nsetp; i++) {
        int ix = IDX_(i);
        X_(ix) = fma(alpha, ZZ_(i) - X_(ix), X_(ix));
      }

      while (1) {
        X_(IDX_(imv)) = 0.0;
        wk->n_exit++;
        wk->flops += 6.0 * n * (nsetp - imv);
        if (imv != nsetp) {
          for (int i = imv + 1; i <= nsetp; i++) {
            /* orthogonal_rotmat :486-491 */
            double sig = orc_hypot(A_(i - 1, i), A_(i, i));
            double cc = A_(i - 1, i) / sig, ss = A_(i, i) / sig;
            A_(i - 1, i) = sig;
            A_(i, i) = 0.0;
            /* orthogonal_rotmatvec :493-497 under @muladd: x = c*a + s*b -> fma(c, a, s*b) */
            for (int j = 1; j <= n; j++) {
              if (j == i) continue;
              double p = A_(i - 1, j), q = A_(i, j);
              A_(i - 1, j) = fma(cc, p, ss * q);
              A_(i, j) = fma(-ss, p, cc * q);
            }
            {
              double p = B_(i - 1), q = B_(i);
              B_(i - 1) = fma(cc, p, ss * q);
              B_(i) = fma(-ss, p, cc * q);
            }
          }
          for (int j = imv; j <= nsetp - 1; j++) {
            for (int i = 1; i <= m; i++) {
              double t = A_(i, j);
              A_(i, j) = A_(i, j + 1), A_(i, j + 1) = t;
            }
            int t = IDX_(j);
            IDX_(j) = IDX_(j + 1), IDX_(j + 1) = t;
          }
        }
        nsetp -= 1;
        int allfeasible = 1;
        for (int i = 1; i <= nsetp; i++) {
          if (X_(IDX_(i)) <= 0) {
            allfeasible = 0;
            imv = i;
            break;
          }
        }
        if (allfeasible) break;
      }
      for (int i = 1; i <= nsetp; i++) ZZ_(i) = B_(i);
      orc_solve_triangular(zz, A, lda, nsetp, 0);
      wk->flops += (double)nsetp * nsetp;
    }
    if (terminated) break;

    if (dual_flag) {
      compute_dual(w, A, lda, n, b, nsetp + 1, m);
      wk->flops += 2.0 * (m - nsetp) * (n - nsetp);
    }
    for (int i = 1; i <= nsetp; i++) X_(IDX_(i)) = ZZ_(i);
  }

  for (int i = 1; i <= n; i++) invidx[IDX_(i) - 1] = i;

  /* residual norm :810-823 / :1046-1059 (rows up to M in the Tikhonov variant) */
  const int mres = tikh ? M : m;
  double sm = 0.0;
  if (nsetp < mres) {
    ORC_SIMD_REDUCE(sm) /* @simd, src/NNLS.jl:813, 1049 */
    for (int i = nsetp + 1; i <= mres; i++) {
      double bi = B_(i);
      ZZ_(i) = bi;
      sm = fma(bi, bi, sm);
    }
  } else {
    for (int j = 1; j <= n; j++) W_(j) = 0.0;
  }
  wk->rnorm = sqrt(sm);
  wk->nsetp = nsetp;
}

/* nnls!(work, A, b)  src/NNLS.jl:198-209.  A is M x N contiguous (lda = M). */
void orc_nnls(orc_nnls_work *wk, const double *Ain, const double *bin) {
  memcpy(wk->A, Ain, sizeof(double) * (size_t)wk->M * wk->N);
  memcpy(wk->b, bin, sizeof(double) * wk->M);
  for (int j = 0; j < wk->N; j++) wk->x[j] = 0.0, wk->idx[j] = j + 1, wk->invidx[j] = j + 1;
  unsafe_nnls(wk, wk->M, wk->M, 1, 0, -1.0);
}

/* nnls!(work, A, b, lambda)  src/NNLS.jl:211-224 + init_nnls!(work, lambda) :567-590.
 * Note the reference zeroes the padded rows of A and b and re-introduces lambda lazily. */
void orc_nnls_tikh_explicit(orc_nnls_work *wk, const double *Apad, const double *bpad, double lambda) {
  const int M = wk->M, N = wk->N, m = M - N;
  memcpy(wk->A, Apad, sizeof(double) * (size_t)M * N);
  memcpy(wk->b, bpad, sizeof(double) * M);
  for (int j = 0; j < N; j++)
    for (int i = m; i < M; i++) wk->A[i + (size_t)j * M] = 0.0;
  for (int i = 0; i < N; i++) {
    wk->x[i] = 0.0, wk->b[m + i] = 0.0, wk->idx[i] = i + 1, wk->invidx[i] = i + 1, wk->diag[i] = 0;
  }
  unsafe_nnls(wk, M, m, 1, 1, lambda);
}

/* Shared body of the two warm-started drivers, src/lsqnonneg.jl:30-84 (mu < 0) and :86-164.
 * The dual is initialised as if the LAST column were already active (:44-70); w[n] is then
 * 0, or 1 when every other dual is <= 0; nsetp still starts at 0.  The reductions at
 * :46-64 / :117-139 are @simd loops (contracted to fma). */
static void warm_start(orc_nnls_work *wk, const double *A0, int lda0, const double *b0, int m, int n,
                       double mu, int tikh) {
  double *C = wk->A, *f = wk->b, *x = wk->x, *w = wk->w, *z = wk->zz;
  const int ldc = wk->M;
#define A0_(i, j) A0[((size_t)(i)-1) + ((size_t)(j)-1) * (size_t)lda0]
#define C_(i, j) C[((size_t)(i)-1) + ((size_t)(j)-1) * (size_t)ldc]
  double den = 0.0;
  ORC_SIMD_REDUCE(den) /* @simd, src/lsqnonneg.jl:46, 117 */
  for (int i = 1; i <= m; i++) den = fma(A0_(i, n), A0_(i, n), den);
  if (tikh) den += mu * mu;
  double xj = 0.0;
  ORC_SIMD_REDUCE(xj) /* @simd, src/lsqnonneg.jl:51, 123 */
  for (int i = 1; i <= m; i++) xj = fma(A0_(i, n) / den, b0[i - 1], xj);
  for (int i = 1; i <= m; i++) z[i - 1] = b0[i - 1] - A0_(i, n) * xj;
  for (int j = 1; j <= n - 1; j++) {
    double wj = 0.0;
    ORC_SIMD_REDUCE(wj) /* @simd, src/lsqnonneg.jl:62, 134 */
    for (int i = 1; i <= m; i++) {
      double Aij = A0_(i, j);
      wj = fma(Aij, z[i - 1], wj);
      C_(i, j) = Aij;
    }
    w[j - 1] = wj;
  }
  w[n - 1] = 0.0;
  {
    int all_nonpos = 1;
    for (int j = 0; j < n; j++)
      if (!(w[j] <= 0)) all_nonpos = 0;
    w[n - 1] = all_nonpos ? 1.0 : 0.0;
  }
  for (int i = 1; i <= m; i++) {
    f[i - 1] = b0[i - 1];
    C_(i, n) = A0_(i, n);
  }
  if (tikh) {
    for (int j = 1; j <= n; j++)
      for (int i = m + 1; i <= wk->M; i++) C_(i, j) = 0.0;
  }
  for (int j = 1; j <= n; j++) {
    x[j - 1] = 0.0;
    wk->idx[j - 1] = j;
    if (tikh) f[m + j - 1] = 0.0, wk->diag[j - 1] = 0;
  }
#undef A0_
#undef C_
}

void orc_nnls_solve(orc_nnls_work *wk, const double *A, int lda, const double *b, int m, int n) {
  wk->flops += 4.0 * m * n;
  warm_start(wk, A, lda, b, m, n, -1.0, 0);
  unsafe_nnls(wk, m, m, 0, 0, -1.0);
}

void orc_nnls_solve_tikh(orc_nnls_work *wk, const double *A0, int lda, const double *b0, int m, int n,
                         double mu) {
  wk->flops += 4.0 * m * n;
  warm_start(wk, A0, lda, b0, m, n, mu, 1);
  unsafe_nnls(wk, m + n, m, 0, 1, mu);
}
