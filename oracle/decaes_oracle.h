/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
 *
 * Plain-C, double-precision, CPU restatement of the DECAES.jl voxelwise T2-distribution
 * path (T2mapSEcorr + T2partSEcorr).  It exists to check libdecaes_cuda and to be timed as
 * the CPU baseline ("port") in bench.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Pinning status: DECAES.jl cannot be executed in this image (no Julia runtime) and its
 * test-suite holds no golden vectors.  The restatement is pinned by the docstring
 * known-answers (src/T2mapSEcorr.jl:126,134), the analytic alpha=180 curve, independent
 * implementations (numpy EPG spec, scipy.optimize.nnls, LAPACK gesdd through numpy) and the
 * invariants ported from test/{nnls,epg,splines,optimization}.jl — see tests/.  Anything
 * beyond that is "parity unpinned" against real Julia output.
 *
 * Arithmetic model: explicit fma() exactly where the reference has muladd / @muladd /
 * @simd-contracted reductions, evaluated in plain sequential order; everything else is
 * unfused (compiled with -ffp-contract=off).  Bitwise equality with Julia is not a goal
 * (LLVM reassociates the @simd loops); following the same algorithmic path is.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout).
 */
#ifndef DECAES_ORACLE_H
#define DECAES_ORACLE_H

#include <stddef.h>
#include <stdint.h>

/* Two builds of the same sources (oracle/Makefile):
 *   liborc.so       the CHECKER: reductions in plain sequential order (ORC_SIMD undefined).
 *   liborc_simd.so  -DORC_SIMD: the loops that carry @simd in the reference (src/NNLS.jl:278-462, 813, 1049;
 *                   src/lsqnonneg.jl:46-64, 117-139) are vectorised with reassociated partial sums, which is
 *                   what LLVM does to them in Julia.  It is (a) the CPU baseline that bench.py times and
 *                   (b) the "second faithful CPU build" against which the L-curve mu-flip rate of the GPU is
 *                   judged (tools/lcurve_ab.py): two builds of the same algorithm that differ only in the
 *                   summation order of those loops already choose a different mu for a few percent of voxels. */
#define ORC_STR_(x) #x
#define ORC_STR(x) ORC_STR_(x)
#ifdef ORC_SIMD
#define ORC_SIMD_REDUCE(v) _Pragma(ORC_STR(omp simd reduction(+ : v)))
#else
#define ORC_SIMD_REDUCE(v)
#endif
#include "../include/decaes_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------- grids (src/utils.jl:7, src/types.jl:108-109, src/T2mapSEcorr.jl:28) ---------- */
void orc_logrange(double a, double b, int n, double *out);
void orc_linrange(double a, double b, int n, double *out);

/* ---------- EPG (src/EPGdecaycurve.jl) ---------- */
/* default kernel for RefConAngle == 180: :936-1028.  work: 6*ETL doubles */
void orc_epg_decay_curve(int ETL, double alpha_deg, double TE, double T2, double T1, double *dc,
                         double *work);
/* value + d/d(alpha in degrees): hand forward-mode of the same recursion, replaces :224-248 */
void orc_epg_decay_curve_jac(int ETL, double alpha_deg, double TE, double T2, double T1,
                             double *dc, double *ddc, double *work /* 12*ETL */);
/* general RefConAngle variant: :722-818.  work: 6*ETL doubles */
void orc_epg_decay_curve_beta(int ETL, double alpha_deg, double TE, double T2, double T1,
                              double beta_deg, double *dc, double *work);
/* value + d/dalpha (per degree) of the general variant, forward mode.  work: 12*ETL doubles */
void orc_epg_decay_curve_beta_jac(int ETL, double alpha_deg, double TE, double T2, double T1,
                                  double beta_deg, double *dc, double *ddc, double *work);
double orc_sind(double x);

/* ---------- NNLS (src/NNLS.jl, src/lsqnonneg.jl:5-164) ---------- */
typedef struct {
  int M, N;       /* allocated rows (m or m+n), columns */
  double *A;      /* M x N column-major, lda = M */
  double *b;      /* M */
  double *x, *w;  /* N */
  double *zz;     /* M */
  int *idx, *invidx;
  unsigned char *diag;
  double rnorm;
  int mode, nsetp;
  /* instrumentation */
  int64_t n_enter, n_exit, n_reject;
  double flops; /* algorithmic FLOPs, cost table in DESIGN.md / SURVEY 8(d) */
} orc_nnls_work;

orc_nnls_work *orc_nnls_alloc(int M, int N);
void orc_nnls_free(orc_nnls_work *w);
/* NNLS.jl:198-209 (load + init + unsafe_nnls!) */
void orc_nnls(orc_nnls_work *w, const double *A, const double *b);
/* NNLS.jl:211-224 with A = [A0; lambda I] supplied explicitly as (m+n) x n */
void orc_nnls_tikh_explicit(orc_nnls_work *w, const double *Apad, const double *bpad, double lambda);
/* lsqnonneg.jl:30-84: warm-started dual, then unsafe_nnls!(init_dual=false) */
void orc_nnls_solve(orc_nnls_work *w, const double *A, int lda, const double *b, int m, int n);
/* lsqnonneg.jl:86-164 */
void orc_nnls_solve_tikh(orc_nnls_work *w, const double *A0, int lda, const double *b0, int m, int n,
                         double mu);
/* NNLS.jl:506-539 */
void orc_solve_triangular(double *z, const double *A, int lda, int n, int transp);
double orc_hypot(double a, double b);

/* ---------- Tikhonov problem + choosers (src/lsqnonneg.jl) ---------- */
typedef struct orc_tikh_cache orc_tikh_cache; /* 8-slot NNLSTikhonovRegProblemCache :392-444 */

typedef struct {
  int m, n;
  const double *A; /* m x n column-major, lda = m */
  const double *b;
  orc_nnls_work *nnls;   /* unregularised */
  orc_tikh_cache *cache; /* regularised */
  double *gamma;         /* singular values (GCV) */
  double *svd_work;
  /* L-curve caches */
  void *lcurve;
  /* instrumentation */
  int64_t n_solves_unreg, n_solves_tikh, n_cache_hits;
} orc_reg_work;

orc_reg_work *orc_reg_alloc(int m, int n);
void orc_reg_free(orc_reg_work *w);
void orc_reg_bind(orc_reg_work *w, const double *A, const double *b);

/* each returns the solution pointer (length n) that save_results! would read, see
 * solution(work) definitions :166,:465,:657,:775,:1088 */
const double *orc_lsqnonneg(orc_reg_work *w);
const double *orc_lsqnonneg_tikh(orc_reg_work *w, double mu, double *res2, double *seminorm2);
const double *orc_lsqnonneg_lcurve(orc_reg_work *w, double *mu, double *chi2);
const double *orc_lsqnonneg_gcv(orc_reg_work *w, double *mu, double *chi2);
const double *orc_lsqnonneg_chi2(orc_reg_work *w, double chi2_target, double *mu, double *chi2,
                                 int *early);
const double *orc_lsqnonneg_mdp(orc_reg_work *w, double delta, double *mu, double *chi2, int *early);

/* singular values by one-sided Jacobi; stands in for LAPACK dgesdd_ (src/utils.jl:103-134) */
void orc_svdvals(int m, int n, const double *A, int lda, double *S /* min(m,n), descending */,
                 double *work /* m*n + n */);
/* lsqnonneg.jl:1321-1329 */
double orc_gcv_dof(int m, int n, const double *gamma, double lambda);

/* lcurve_corner with an arbitrary f: logmu -> (xi, eta)   (src/lsqnonneg.jl:872-972) */
typedef void (*orc_lcurve_fn)(double t, double *P, void *ctx);
double orc_lcurve_corner(orc_lcurve_fn f, void *ctx, double xlow, double xhigh, double xtol,
                         double Ptol, double Ctol, int backtracking, int *n_feval);

/* ---------- 1-D optimisers (src/optimization.jl) ---------- */
typedef double (*orc_fn1)(double x, void *ctx);
void orc_brent_root(orc_fn1 f, void *ctx, double x0, double x1, double fx0, double fx1, double xatol,
                    double xrtol, double ftol, int maxiters, double *x, double *fx);
void orc_bracket_root_monotonic(orc_fn1 f, void *ctx, double a, double delta, double dilate, int mono,
                                int maxiters, double *oa, double *ob, double *ofa, double *ofb);
void orc_brent_minimize(orc_fn1 f, void *ctx, double x1, double x2, double xrtol, double xatol,
                        int maxiters, double *x, double *y);

/* ---------- flip-angle search (src/splines.jl) ---------- */
/* CubicHermiteInterpolator + minimize: splines.jl:62-110 */
void orc_hermite_minimize(double a, double b, double u0, double u1, double m0, double m1, double *x,
                          double *u);
typedef void (*orc_fg_fn)(int I /* 1-based grid index */, double *u, double *du, void *ctx);
/* DiscreteSurrogateSearcher + bisection_search with CubicHermiteSplineSurrogate:
 * splines.jl:504-566, 705-850.  order[] receives the 1-based probe order. */
void orc_surrogate_search(orc_fg_fn fg, void *ctx, const double *grid, int ngrid, int mineval,
                          int maxeval, double *x_opt, double *u_opt, int *order, int *norder);

/* same search with CubicSplineSurrogate(...; legacy = true)  splines.jl:456-500: suggest_point is the
 * sampled minimum of the FITPACK interpolating spline through the probed points (du is ignored) */
void orc_surrogate_search_legacy(orc_fg_fn fg, void *ctx, const double *grid, int ngrid, int mineval,
                                 int maxeval, double *x_opt, double *u_opt, int *order, int *norder);

/* ---------- legacy = true algorithms (oracle/orc_spline.c) ---------- */
#define ORC_SPLINE_MAX 64
#define ORC_CHI2_LEGACY_MAXPTS 22 /* mu = 0 plus at most 21 doublings of 1e-3; same cut as csrc/legacy.cuh */
/* FITPACK curfit(iopt=0, s=0) + splev as wrapped by Dierckx.Spline1D (src/splines.jl:311-314) */
int orc_fitpack_interp(const double *x, const double *y, int m, int k, double *t /* m+k+1 */, double *c /* m */);
void orc_fitpack_splev(const double *t, int n, const double *c, int k, const double *x, int m, double *y);
/* Julia's start:step:stop for Float64 (base/twiceprecision.jl) */
typedef struct {
  int rational;
  int64_t start_n, step_n, den, len;
  double start, step;
} orc_jl_range_t;
void orc_jl_range(double start, double step, double stop, orc_jl_range_t *r);
double orc_jl_range_at(const orc_jl_range_t *r, int64_t i /* 0-based */);
/* src/splines.jl:419-446 */
int orc_spline_opt_legacy(const double *X, const double *Y, int m, double *x, double *y);
int orc_spline_root_legacy(const double *X, const double *Y, int m, double value, double *x);
/* src/lsqnonneg.jl:595-636 with legacy = true; f(mu) -> res2(mu) */
int orc_chi2_search_legacy(orc_fn1 f, void *ctx, double res2min, double chi2fact, double *mu, double *res2);
/* lsqnonneg_chi2!(work, chi2_target, legacy = true)  src/lsqnonneg.jl:504-533; *early = 3 when mu_final == 0 */
const double *orc_lsqnonneg_chi2_legacy(orc_reg_work *w, double chi2_target, double *mu, double *chi2, int *early);

/* ---------- pipeline (src/T2mapSEcorr.jl, src/T2partSEcorr.jl) ---------- */
typedef struct {
  int64_t voxels_processed;
  int64_t nnls_unreg, nnls_tikh, cols_entered, cols_exited, cols_rejected, cache_hits;
  int64_t early_returns; /* chi2/mdp early-return branches (stale-slot quirk) */
  double flops;          /* algorithmic FLOPs, SURVEY 8(d) cost table */
  double seconds;
  int threads;
} orc_stats;

int orc_t2map(const double *image, int64_t nvox, int64_t stride, const decaes_t2map_opts *opts,
              const decaes_t2part_opts *part, const decaes_t2map_out *out, int nthreads,
              orc_stats *stats);
int orc_t2part(const double *dist, int64_t nvox, int64_t stride, const decaes_t2part_opts *part,
               double *sfr, double *sgm, double *mfr, double *mgm);
int orc_setup_tables(const decaes_t2map_opts *opts, double *echotimes, double *t2times,
                     double *refangleset, double *decaybasisset, double *ddecaybasisset);
int orc_validate_t2map_opts(const decaes_t2map_opts *o, char *msg, int msglen);
int orc_validate_t2part_opts(const decaes_t2part_opts *o, char *msg, int msglen);

/* synthetic volume (mock_image recipe, src/utils.jl:623-658) on the CPU; same RNG keying
 * as decaes_mock_image_device but NOT required to be bit-identical to it. */
void orc_mock_image(double *image, int64_t nvox, int64_t stride, int64_t first_voxel, int nTE,
                    double TE, double T1, double SNR, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
