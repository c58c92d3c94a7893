// `legacy = true` algorithms of DECAES.jl on the device (warp-per-voxel, called from voxel.cuh).
//
//   spline_opt_legacy / spline_root_legacy      src/splines.jl:419-446
//   make_spline (Dierckx.Spline1D, s = 0)       src/splines.jl:311-314
//   chi2_search_from_minimum(...; legacy=true)  src/lsqnonneg.jl:595-636
//
// The reference fits an interpolating spline with FITPACK (Dierckx.jl) and then scans it on the grid
// knots[1]:0.001:knots[end] (130,001 points for the default 50..180 degree range) — brute force kept from
// the MATLAB toolbox.  Here lane 0 builds the spline exactly as FITPACK's curfit does for s = 0 (row-by-row
// Givens reduction of the banded observation matrix, back substitution) and the 32 lanes share the scan:
// lane l evaluates samples l, l+32, ... with the de Boor-Cox recursion of fpbspl, operation by operation
// (unfused: the library is built with --fmad=false), and the warp reduces (value, sample index) with
// "first strictly smaller wins", which is what the sequential scan of the reference returns.
//
// Only the LEGACY instantiation of the pipeline kernel (Warp<true, true>) references this code: the default kernel
// is compiled without it, so its register allocation and code footprint are untouched.
#pragma once
#include "common.cuh"

namespace decaes {

#define DECAES_LEGACY_CHI2_MAXPTS 22  // mu = 0 plus at most 21 doublings of 1e-3 (2^20 + 1 samples)

// Julia's start:step:stop for Float64 (base/twiceprecision.jl) resolved on the host: either the rational form
// (start_n + i*step_n)/den or the literal start + i*step, both rounded once.
struct LegacyRange {
  int rational;
  long long start_n, step_n, den, len;
  double start, step;
};

__device__ __forceinline__ double legacy_range_at(const LegacyRange &r, long long i) {
  if (r.rational) return (double)(r.start_n + i * r.step_n) / (double)r.den;
  return fma((double)i, r.step, r.start);
}

// ---- FITPACK: curfit(iopt = 0, w = 1, s = 0) for m points, degree k; lane 0 only.
// ws: t[m+k+1] | c[m] | z[m] | a[m][4]   (X, Y may not alias ws)
__device__ __noinline__ void fitpack_interp_dev(const double *X, const double *Y, int m, int k, double *t, double *c,
                                                double *z, double *a) {
  const int k1 = k + 1, n = m + k1, nk1 = n - k1;
  const int mk1 = m - k1, k3 = k / 2;
  {
    int i = k1 + 1, j = k3 + 2;
    if (k3 * 2 != k) {
      for (int l = 0; l < mk1; l++, i++, j++) t[i - 1] = X[j - 1];
    } else {
      for (int l = 0; l < mk1; l++, i++, j++) t[i - 1] = (X[j - 1] + X[j - 2]) * 0.5;
    }
    for (int q = 0; q < k1; q++) t[q] = X[0], t[n - 1 - q] = X[m - 1];
  }
  for (int i = 0; i < nk1; i++) {
    z[i] = 0.0;
    for (int q = 0; q < 4; q++) a[4 * i + q] = 0.0;
  }
  int l = k1;
  for (int it = 0; it < m; it++) {
    const double xi = X[it];
    double yi = Y[it];
    while (!(xi < t[l] || l == nk1)) l++;
    // fpbspl
    double h[4] = {1.0, 0.0, 0.0, 0.0}, hh[3];
    for (int j = 1; j <= k; j++) {
      for (int i = 0; i < j; i++) hh[i] = h[i];
      h[0] = 0.0;
      for (int i = 1; i <= j; i++) {
        const int li = l + i, lj = li - j;
        const double f = hh[i - 1] / (t[li - 1] - t[lj - 1]);
        h[i - 1] = h[i - 1] + f * (t[li - 1] - xi);
        h[i] = f * (xi - t[lj - 1]);
      }
    }
    // rotate the new row into the band triangle (fpgivs / fprota)
    int jj = l - k1;
    for (int ii = 1; ii <= k1; ii++) {
      jj++;
      const double piv = h[ii - 1];
      if (piv == 0.0) continue;
      double ww = a[4 * (jj - 1)];
      const double store = fabs(piv);
      double dd;
      if (store >= ww) {
        const double r = ww / piv;
        dd = store * sqrt(1.0 + r * r);
      } else {
        const double r = piv / ww;
        dd = ww * sqrt(1.0 + r * r);
      }
      const double cs = ww / dd, sn = piv / dd;
      a[4 * (jj - 1)] = dd;
      {
        const double s1 = yi, s2 = z[jj - 1];
        z[jj - 1] = cs * s2 + sn * s1;
        yi = cs * s1 - sn * s2;
      }
      if (ii == k1) break;
      int i2 = 0;
      for (int i1 = ii + 1; i1 <= k1; i1++) {
        i2++;
        const double s1 = h[i1 - 1], s2 = a[4 * (jj - 1) + i2];
        a[4 * (jj - 1) + i2] = cs * s2 + sn * s1;
        h[i1 - 1] = cs * s1 - sn * s2;
      }
    }
  }
  // fpback
  c[nk1 - 1] = z[nk1 - 1] / a[4 * (nk1 - 1)];
  int ib = nk1 - 1;
  for (int jb = 2; jb <= nk1; jb++) {
    double store = z[ib - 1];
    const int i1 = (jb <= k1 - 1) ? jb - 1 : k1 - 1;
    int mm = ib;
    for (int ll = 1; ll <= i1; ll++) {
      mm++;
      store = store - c[mm - 1] * a[4 * (ib - 1) + ll];
    }
    c[ib - 1] = store / a[4 * (ib - 1)];
    ib--;
  }
}

// ---- the scan: samples i = lane, lane + 32, ... of range r; mode 0: minimise spl(x), mode 1: minimise |spl(x) - value|.
// Returns (x, y) of the first strict minimum of the sequential scan on every lane.
__device__ __noinline__ void legacy_spline_scan(const double *t, const double *c, int m, int k, LegacyRange r, int mode,
                                                double value, double &x_out, double &y_out) {
  const int lane = lane_id();
  const int k1 = k + 1, n = m + k1, nk1 = n - k1;
  int l = k1, lcur = -1;
  double tt[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, cc[4] = {0.0, 0.0, 0.0, 0.0};
  double best = CUDART_INF;
  int besti = 0x7fffffff;
  _Pragma("unroll 1") for (long long i = lane; i < r.len; i += 32) {
    const double x = legacy_range_at(r, i);
    // splev: knot interval t(l) <= x < t(l+1).  The scan is monotone, so the cached interval (tt[2] = t(l),
    // tt[3] = t(l+1)) almost always still holds; otherwise search from it and reload the 2k knots / k+1 coefficients.
    if (!(lcur == l && !(x < tt[2]) && (x < tt[3] || l == nk1))) {
      while (x < t[l - 1] && l != k1) l--;
      while (!(x < t[l] || l == nk1)) l++;
      lcur = l;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        int p = l - 3 + q;  // 0-based index of t(l-2+q)
        p = p < 0 ? 0 : (p > n - 1 ? n - 1 : p);
        tt[q] = t[p];
      }
#pragma unroll
      for (int q = 0; q < 4; q++) cc[q] = (q <= k) ? c[l - k1 + q] : 0.0;
    }
    // fpbspl, unrolled for k <= 3: t(l+i) = tt[i+2], t(l+i-j) = tt[i-j+2]
    double h0, h1 = 0.0, h2 = 0.0, h3 = 0.0;
    {
      const double f = 1.0 / (tt[3] - tt[2]);
      h0 = 0.0 + f * (tt[3] - x);
      h1 = f * (x - tt[2]);
    }
    if (k >= 2) {
      const double g0 = h0, g1 = h1;
      double f = g0 / (tt[3] - tt[1]);
      h0 = 0.0 + f * (tt[3] - x);
      h1 = f * (x - tt[1]);
      f = g1 / (tt[4] - tt[2]);
      h1 = h1 + f * (tt[4] - x);
      h2 = f * (x - tt[2]);
    }
    if (k >= 3) {
      const double g0 = h0, g1 = h1, g2 = h2;
      double f = g0 / (tt[3] - tt[0]);
      h0 = 0.0 + f * (tt[3] - x);
      h1 = f * (x - tt[0]);
      f = g1 / (tt[4] - tt[1]);
      h1 = h1 + f * (tt[4] - x);
      h2 = f * (x - tt[1]);
      f = g2 / (tt[5] - tt[2]);
      h2 = h2 + f * (tt[5] - x);
      h3 = f * (x - tt[2]);
    }
    double sp = 0.0;
    sp = sp + cc[0] * h0;
    sp = sp + cc[1] * h1;
    if (k >= 2) sp = sp + cc[2] * h2;
    if (k >= 3) sp = sp + cc[3] * h3;
    const double y = mode == 0 ? sp : fabs(sp - value);
    if (y < best) best = y, besti = (int)i;
  }
  warp_argmin_first(best, besti);  // smaller value wins, ties -> smaller sample index
  if (besti == 0x7fffffff) besti = 0, best = CUDART_NAN;  // nothing compared below +Inf: the scan keeps its first sample
  x_out = legacy_range_at(r, besti);
  y_out = best;
}

}  // namespace decaes
