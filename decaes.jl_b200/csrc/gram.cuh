// Warp-cooperative Lawson–Hanson NNLS on the normal equations ("Gram form"), one voxel per warp.
//
// Same active-set algorithm as the reference (src/NNLS.jl:605-1061 through the warm-started
// drivers src/lsqnonneg.jl:30-164): identical pivot rule (largest positive dual, first on ties),
// identical accept/reject test on the sign of the entering coefficient (b1/A1 > 0, src/NNLS.jl:300),
// identical feasibility step / removal rule and 3n iteration cap.  What changes is the linear
// algebra underneath: instead of reflecting the whole nTE x nT2 matrix for every pivot (a BLAS-2
// sweep over ~18 KB of shared memory per pivot), the warp keeps
//     G = A'A (+ mu^2 I)   n x n symmetric
//     c = A'b              n
//     M = L^-1             inverse Cholesky factor of G_PP in pivot order
// so that appending a column is two O(k) fma chains per lane, the solution is s = M'(M c_P)
// maintained incrementally, and the dual is w = c - G_P x_P.
// The minimiser of the Tikhonov problems (mu > 0) is unique, so those solves may be warm-started
// from the active set of the nearest mu already solved; unregularised solves follow the
// reference's cold-start path pivot by pivot and are polished with one step of iterative
// refinement on the explicit residual (done by the caller, which owns A).
//
// The active sets are small (mean 5-7 columns on the benchmark volumes), so every operation is
// latency bound and the kernel as a whole is instruction-cache bound (L1.5 I-cache = 32 KB).  The
// code is therefore written for SIZE: one out-of-line copy of each primitive, plain loops, all
// shared-memory vectors at compile-time offsets from one base pointer, integer REDUX for the
// arg-max / arg-min reductions.  Warm starts and removals refactor the (tiny) active block with a
// lane-parallel symmetric elimination (gram_factor) instead of k sequential appends.
//
// Shared-memory layout of one warp (doubles, relative to V):
//     c[VS] y[VS] s[VS] t1[VS] t2[VS] x[VS] w[VS] P[VS ints]  T[n x ld]      (VS = 40 or 64 >= n)
// T (ld = (n+1)|1, odd => both access directions are bank-conflict free) holds the lower triangle
// of G and, in the strictly-upper part, M:
//     G(p,q) = T[max(p,q)*ld + min(p,q)]          M(t,u) = T[u*ld + t + 1]   (u <= t)
#pragma once
#include "common.cuh"

namespace decaes {

#ifdef DECAES_PROFILE
// [0..3] cycles append/factor/dual/nnls, [4..7] calls, [8] sum k at append, [9] sum k at factor,
// [11] warm calls, [12] cold calls, [13] sum final k, [14] sum inner iterations, [15] factor fallbacks
__device__ unsigned long long g_prof[16];
__device__ unsigned long long g_khist[2][5];
__device__ unsigned long long g_solve_hist[3][48];  // per solve index within a voxel (0-27 Tikhonov, 28+ unregularised): calls, appends, inner iterations  // k at append / factor: <=4, <=8, <=12, <=16, >16
#define GP_BEGIN() long long gp_t0 = clock64()
#define GP_END(id) if (lane_id() == 0) { atomicAdd(&g_prof[id], (unsigned long long)(clock64() - gp_t0)); atomicAdd(&g_prof[4 + id], 1ull); }
#define GP_ADD(id, v) if (lane_id() == 0) atomicAdd(&g_prof[id], (unsigned long long)(v))
#define GP_HIST(w, k) if (lane_id() == 0) atomicAdd(&g_khist[w][(k) <= 4 ? 0 : (k) <= 8 ? 1 : (k) <= 12 ? 2 : (k) <= 16 ? 3 : 4], 1ull)
#else
#define GP_HIST(w, k)
#define GP_BEGIN()
#define GP_END(id)
#define GP_ADD(id, v)
#endif

// VS = vector stride: 40 when nT2 <= 40 (the reference's default grid; 1.4 KB less shared memory per warp than a
// stride of 64, which is what lets a 12th warp fit on the SM for the benchmark configs), 64 otherwise.  It is a
// template parameter of the solver so that every offset stays an immediate.
#define GV_LAYOUT(VS)                                                                                             \
  enum { GV_C = 0, GV_Y = (VS), GV_S = 2 * (VS), GV_T1 = 3 * (VS), GV_T2 = 4 * (VS), GV_X = 5 * (VS), GV_W = 6 * (VS), \
         GV_P = 7 * (VS), GV_T = 7 * (VS) + (VS) / 2 }
__host__ __device__ constexpr int gv_block_doubles(int vs) { return 7 * vs + vs / 2; }  // vectors + pivot list, in front of G / M

struct GramOut {
  int k;                     // number of active columns
  unsigned long long mask;   // active set as a bit mask
  double xnorm_sq;           // sum of squares of the solution
  int iters, nappend;        // diagnostics (DECAES_PROFILE histograms)
  bool capped;               // stopped by the 3n iteration cap (mode = 1, src/NNLS.jl:693-698)
};

#define GM_(t, u) T[(u) * ld + (t) + 1]
__device__ __forceinline__ double gram_G(const double *T, int ld, int a, int b) {
  int hi = a > b ? a : b, lo = a > b ? b : a;
  return T[hi * ld + lo];
}

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
  unsigned lo = __reduce_or_sync(DECAES_FULL_MASK, (unsigned)v);
  unsigned hi = __reduce_or_sync(DECAES_FULL_MASK, (unsigned)(v >> 32));
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __noinline__ unsigned long long mask_of(const int *P, int k) {  // cold path (k <= 64: two slots per lane)
  const int lane = lane_id();
  unsigned long long m = 0ull;
  if (lane < k) m = 1ull << P[lane];
  if (lane + 32 < k) m |= 1ull << P[lane + 32];
  return warp_or64(m);
}

// Arg-max / arg-min of NON-NEGATIVE doubles through their (order-preserving) bit patterns with the
// integer REDUX unit: three reductions instead of a five-round shuffle butterfly.  `key` is the
// lane's best candidate (0 / ~0 = none), `idx` its index; ties resolve to the smallest index.
__device__ __forceinline__ int warp_argmax_bits(unsigned long long key, int idx, unsigned long long &best) {
  const unsigned hi = __reduce_max_sync(DECAES_FULL_MASK, (unsigned)(key >> 32));
  const unsigned lo = __reduce_max_sync(DECAES_FULL_MASK, ((unsigned)(key >> 32) == hi) ? (unsigned)key : 0u);
  best = ((unsigned long long)hi << 32) | lo;
  return (int)__reduce_min_sync(DECAES_FULL_MASK, (key == best) ? (unsigned)idx : 0x7fffffffu);
}
__device__ __forceinline__ int warp_argmin_bits(unsigned long long key, int idx, unsigned long long &best) {
  const unsigned hi = __reduce_min_sync(DECAES_FULL_MASK, (unsigned)(key >> 32));
  const unsigned lo = __reduce_min_sync(DECAES_FULL_MASK, ((unsigned)(key >> 32) == hi) ? (unsigned)key : 0xffffffffu);
  best = ((unsigned long long)hi << 32) | lo;
  return (int)__reduce_min_sync(DECAES_FULL_MASK, (key == best) ? (unsigned)idx : 0x7fffffffu);
}

// Order-preserving map double -> uint64 (total order -inf < ... < -0 < +0 < ... < +inf; NaNs are the
// caller's business), so that value reductions can run on the integer REDUX unit.
__device__ __forceinline__ unsigned long long dkey(double x) {
  const long long b = __double_as_longlong(x);
  return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}

// Append column j to the factorisation at pivot position k.  Returns false (and changes nothing)
// when the column is numerically dependent (d^2 <= 0) or, with `need_positive`, when its
// coefficient would not be positive (the reference's b1/A1 > 0 test).  The caller increments k.
template <int VS>
__device__ __noinline__ bool gram_append(double *V, int ld, int k, int j, double mu2, bool need_positive) {
  GV_LAYOUT(VS);
  __builtin_assume(__isShared(V));
  const int lane = lane_id();
  double *T = V + GV_T;
  const int *P = (const int *)(V + GV_P);
  GP_BEGIN();
  GP_ADD(8, k);
  GP_HIST(0, k);
  _Pragma("unroll 1") for (int t = lane; t < k; t += 32) V[GV_T1 + t] = gram_G(T, ld, P[t], j);
  __syncwarp();
  // l = M g  (lane <-> row t; for fixed u the lanes read consecutive words)
  _Pragma("unroll 1") for (int t = lane; t < k; t += 32) {
    double a = 0.0;
    _Pragma("unroll 4") for (int u = 0; u <= t; u++) a = fma(GM_(t, u), V[GV_T1 + u], a);
    V[GV_T2 + t] = a;
  }
  __syncwarp();
#ifndef DECAES_OLD_APPEND
  // |l|^2 and l'y: every lane sums the k terms in the same order (k is small; cheaper than a butterfly).  The same
  // pass accumulates, for lane u < k, the raw entry of the new row of M, a_u = sum_{t >= u} l_t M(t,u): it does not
  // depend on the accept test, so it overlaps with the two chains instead of forming a third stage after the rsqrt.
  double ll = 0.0, ly = 0.0, au = 0.0;
  const bool mine = lane < k;
  _Pragma("unroll 4") for (int t = 0; t < k; t++) {
    const double lt = V[GV_T2 + t];
    double mt = 0.0;
    if (mine && t >= lane) mt = GM_(t, lane);
    ll = fma(lt, lt, ll);
    ly = fma(lt, V[GV_Y + t], ly);
    au = fma(lt, mt, au);
  }
  const double d2 = (T[j * ld + j] + mu2) - ll;
  bool ok = d2 > 0.0;
  double dinv = 0.0, ynew = 0.0;
  if (ok) {
    dinv = rsqrt(d2);
    ynew = (V[GV_C + j] - ly) * dinv;
    ok = !need_positive || ynew > 0.0;
  }
  if (ok) {
    // new row of M: M(k,u) = -dinv * a_u (lane <-> column u); columns 32.. (k > 32 only) take the separate loop
    if (mine) {
      const double m = -dinv * au;
      GM_(k, lane) = m;
      V[GV_S + lane] = fma(ynew, m, V[GV_S + lane]);
    }
    _Pragma("unroll 1") for (int u = lane + 32; u < k; u += 32) {
      double a = 0.0;
      _Pragma("unroll 4") for (int t = u; t < k; t++) a = fma(V[GV_T2 + t], GM_(t, u), a);
      const double m = -dinv * a;
      GM_(k, u) = m;
      V[GV_S + u] = fma(ynew, m, V[GV_S + u]);
    }
    if (lane == 0) {
      GM_(k, k) = dinv;
      V[GV_Y + k] = ynew;
      V[GV_S + k] = ynew * dinv;
      ((int *)(V + GV_P))[k] = j;
    }
    __syncwarp();
  }
#else
  // |l|^2 and l'y: every lane sums the k terms in the same order (k is small; cheaper than a butterfly)
  double ll = 0.0, ly = 0.0;
  _Pragma("unroll 4") for (int t = 0; t < k; t++) {
    const double lt = V[GV_T2 + t];
    ll = fma(lt, lt, ll);
    ly = fma(lt, V[GV_Y + t], ly);
  }
  const double d2 = (T[j * ld + j] + mu2) - ll;
  bool ok = d2 > 0.0;
  double dinv = 0.0, ynew = 0.0;
  if (ok) {
    dinv = rsqrt(d2);
    ynew = (V[GV_C + j] - ly) * dinv;
    ok = !need_positive || ynew > 0.0;
  }
  if (ok) {
    // new row of M: M(k,u) = -dinv * sum_{t >= u} l_t M(t,u)   (lane <-> column u, contiguous in t)
    _Pragma("unroll 1") for (int u = lane; u < k; u += 32) {
      double a = 0.0;
      _Pragma("unroll 4") for (int t = u; t < k; t++) a = fma(V[GV_T2 + t], GM_(t, u), a);
      const double m = -dinv * a;
      GM_(k, u) = m;
      V[GV_S + u] = fma(ynew, m, V[GV_S + u]);
    }
    if (lane == 0) {
      GM_(k, k) = dinv;
      V[GV_Y + k] = ynew;
      V[GV_S + k] = ynew * dinv;
      ((int *)(V + GV_P))[k] = j;
    }
    __syncwarp();
  }
#endif
  GP_END(0);
  return ok;
}

// Factor the whole active block at once: M = L^-1 with G_PP + mu2 I = L L', y = M c_P, s = M'y for
// the k columns listed in P.  Symmetric Gaussian elimination on [K | c | I] done in place in M's
// storage: after step p, row i holds the multipliers R(i, 0..p) of the unit lower triangular
// R = Ltilde^-1 and the still-to-be-eliminated K entries (i, p+1..i).  M = D^-1/2 R.
// Work split: lane -> (row i = lane / 4 + 8a, entries q = lane % 4 + 4e), so that the usual block
// (k <= 8) takes one row and at most two entries per lane and the loops below run once or twice.
// (A register-resident, fully unrolled variant is 2-3x faster in isolation and SLOWER inside the
// pipeline: the kernel is instruction-fetch bound and straight-line code has no reuse.)
// Returns false when a pivot is not positive (numerically dependent set): the caller falls back
// to sequential appends, which drop the offending column.
template <int VS>
__device__ __noinline__ bool gram_factor(double *V, int ld, int k, double mu2) {
  GV_LAYOUT(VS);
  __builtin_assume(__isShared(V));
  const int lane = lane_id();
  const int r0 = lane >> 2, c0 = lane & 3;
  double *T = V + GV_T;
  const int *P = (const int *)(V + GV_P);
  GP_BEGIN();
  GP_ADD(9, k);
  GP_HIST(1, k);
  _Pragma("unroll 1") for (int i = r0; i < k; i += 8) {
    const int pi = P[i];
    _Pragma("unroll 1") for (int q = c0; q <= i; q += 4) GM_(i, q) = gram_G(T, ld, pi, P[q]) + (q == i ? mu2 : 0.0);
    if (c0 == 0) V[GV_Y + i] = V[GV_C + pi];
  }
  bool ok = true;
  _Pragma("unroll 1") for (int p = 0; p < k; p++) {
    __syncwarp();
    // snapshot of column p (rows >= p), double-buffered in t1 / t2 so one barrier per step suffices
    double *col = V + ((p & 1) ? GV_T2 : GV_T1);
    _Pragma("unroll 1") for (int i = p + lane; i < k; i += 32) col[i] = GM_(i, p);
    __syncwarp();
    const double d = col[p];
    if (!(d > 0.0)) {
      ok = false;
      break;
    }
    const double rinv = __drcp_rn(d);
    _Pragma("unroll 1") for (int i = r0; i < k; i += 8) {
      if (i <= p) continue;
      const double f = col[i] * rinv;
      // row_i -= f * (row p of R | column p of K) on the entries q in [0, i] \ {p}; entry p becomes -f
      _Pragma("unroll 1") for (int q = c0; q <= i; q += 4) {
        const double o = (q < p) ? GM_(p, q) : col[q];
        GM_(i, q) = (q == p) ? -f : fma(-f, o, GM_(i, q));
      }
      if (c0 == 0) V[GV_Y + i] = fma(-f, V[GV_Y + p], V[GV_Y + i]);
    }
  }
  __syncwarp();
  if (ok) {
    _Pragma("unroll 1") for (int i = r0; i < k; i += 8) {
      const double dinv = rsqrt(GM_(i, i));
      __syncwarp(0xfu << (lane & 28));  // the four lanes of a row read the pivot before it is overwritten
      _Pragma("unroll 1") for (int q = c0; q < i; q += 4) GM_(i, q) *= dinv;
      if (c0 == (i & 3)) GM_(i, i) = dinv;
      if (c0 == 0) V[GV_Y + i] *= dinv;
    }
    __syncwarp();
    _Pragma("unroll 1") for (int u = lane; u < k; u += 32) {
      double a = 0.0;
      _Pragma("unroll 4") for (int t = u; t < k; t++) a = fma(GM_(t, u), V[GV_Y + t], a);
      V[GV_S + u] = a;
    }
    __syncwarp();
  }
  GP_END(1);
  return ok;
}

#ifndef DECAES_NO_FACTOR_SMALL
// Factor a small active block (k <= 8) in REGISTERS: the same symmetric elimination on [K | c | I] as gram_factor,
// with lane (i, c0) = (lane >> 2, lane & 3) holding entries q = c0 and c0 + 4 of row i (for q <= p the multipliers
// R(i, q), beyond them what is left of K(i, q)) and a replica of y_i.  Pivot row, pivot column and y_p travel by warp
// shuffles: no shared-memory round trips and no barriers inside the elimination (gram_factor needs two per step).
// M, y and s = M'y are written to shared memory at the end, exactly where gram_factor leaves them.  73 % of the
// (re)factorisations of the benchmark volumes have k <= 8 (DECAES_PROFILE histogram).
template <int VS>
__device__ __noinline__ bool gram_factor_small(double *V, int ld, int k, double mu2) {
  GV_LAYOUT(VS);
  __builtin_assume(__isShared(V));
  const int lane = lane_id();
  const int i = lane >> 2, c0 = lane & 3, c1 = c0 + 4;
  double *T = V + GV_T;
  const int *P = (const int *)(V + GV_P);
  GP_BEGIN();
  GP_ADD(9, k);
  GP_HIST(1, k);
  const bool rv = i < k, v0 = rv && c0 <= i, v1 = rv && c1 <= i;
  const int pi = P[rv ? i : 0];
  double e0 = 0.0, e1 = 0.0;
  if (v0) e0 = gram_G(T, ld, pi, P[c0]) + (c0 == i ? mu2 : 0.0);
  if (v1) e1 = gram_G(T, ld, pi, P[c1]) + (c1 == i ? mu2 : 0.0);
  double y = rv ? V[GV_C + pi] : 0.0;
  const int rowbase = lane & ~3;
  bool ok = true;
  _Pragma("unroll 1") for (int p = 0; p < k; p++) {
    const int pc = p & 3;
    const double v = (p >> 2) ? e1 : e0;  // for the lanes with c0 == pc: their row's entry in column p
    const double d = __shfl_sync(DECAES_FULL_MASK, v, 4 * p + pc);       // K(p, p)
    const double colv = __shfl_sync(DECAES_FULL_MASK, v, rowbase | pc);  // K(i, p)
    const double a0 = __shfl_sync(DECAES_FULL_MASK, e0, 4 * p + c0);     // R(p, c0)
    const double a1 = __shfl_sync(DECAES_FULL_MASK, e1, 4 * p + c0);     // R(p, c1)
    const double b0 = __shfl_sync(DECAES_FULL_MASK, v, 4 * c0 + pc);     // K(c0, p)
    const double b1 = __shfl_sync(DECAES_FULL_MASK, v, 4 * c1 + pc);     // K(c1, p)
    const double yp = __shfl_sync(DECAES_FULL_MASK, y, 4 * p);
    if (!(d > 0.0)) {
      ok = false;
      break;
    }
    const double f = colv * __drcp_rn(d);
    if (rv && i > p) {
      // row_i -= f * (row p of R | column p of K) on the entries q in [0, i] \ {p}; entry p becomes -f
      if (v0) e0 = (c0 == p) ? -f : fma(-f, c0 < p ? a0 : b0, e0);
      if (v1) e1 = (c1 == p) ? -f : fma(-f, c1 < p ? a1 : b1, e1);
      y = fma(-f, yp, y);
    }
  }
  if (ok) {
    const double dii = __shfl_sync(DECAES_FULL_MASK, (i >> 2) ? e1 : e0, rowbase | (i & 3));
    const double dinv = rsqrt(rv ? dii : 1.0);
    e0 = (c0 == i) ? dinv : e0 * dinv;
    e1 = (c1 == i) ? dinv : e1 * dinv;
    y *= dinv;
    if (v0) GM_(i, c0) = e0;
    if (v1) GM_(i, c1) = e1;
    if (rv && c0 == 0) V[GV_Y + i] = y;
    // s_u = sum_{t >= u} M(t, u) y_t: one product per entry, summed over the rows (lane bits 2..4)
    double s0 = v0 ? e0 * y : 0.0, s1 = v1 ? e1 * y : 0.0;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      s0 += __shfl_xor_sync(DECAES_FULL_MASK, s0, o);
      s1 += __shfl_xor_sync(DECAES_FULL_MASK, s1, o);
    }
    if (i == 0) {
      if (c0 < k) V[GV_S + c0] = s0;
      if (c1 < k) V[GV_S + c1] = s1;
    }
  }
  __syncwarp();
  GP_END(1);
  return ok;
}
#endif

// (Re)build the factorisation of the columns listed in P[0:k).  Returns the number of columns kept.
template <int VS>
__device__ __noinline__ int gram_refactor(double *V, int ld, int k, double mu2) {
  GV_LAYOUT(VS);
#ifndef DECAES_NO_FACTOR_SMALL
  if (k == 0 || (k <= 8 ? gram_factor_small<VS>(V, ld, k, mu2) : gram_factor<VS>(V, ld, k, mu2))) return k;
#else
  if (k == 0 || gram_factor<VS>(V, ld, k, mu2)) return k;
#endif
  // numerically dependent set (rare): sequential appends, dropping the offending columns
  GP_ADD(15, 1);
  int *P = (int *)(V + GV_P);
  int kk = 0;
  for (int t = 0; t < k; t++) {
    const int j = P[t];
    __syncwarp();
    if (gram_append<VS>(V, ld, kk, j, mu2, false)) {
      kk++;
    } else {
      if (lane_id() == 0) V[GV_X + j] = 0.0;
      __syncwarp();
    }
  }
  return kk;
}

// Remove pivot position r from the factorisation by Givens rotations instead of a refactorisation:
// rotating rows (i, i+1), i = r..k-2, of M pushes the mass of column r into the last row; dropping that
// row and column r leaves the inverse Cholesky factor of the reduced block (up to row signs, which
// neither the append formulas nor s = M'y care about).  lane <-> columns lane and lane + 32 of M (the
// second one only when k > 32, warp-uniform): a rotation touches only the lane's own entries, the
// rotation coefficients follow from column r alone and are computed redundantly by every lane; the lane
// whose column disappears carries y.  Removing the LAST pivot costs nothing (no rotation).
// The caller compacts P and recomputes s (gram_solve_s) once all removals are done.
template <int VS>
__device__ __noinline__ void gram_downdate(double *V, int ld, int k, int r) {
  GV_LAYOUT(VS);
  __builtin_assume(__isShared(V));
  const int lane = lane_id();
  double *T = V + GV_T;
  const bool wide = k > 32;
  const int u0 = lane, u1 = lane + 32;          // this lane's columns
  const int ud0 = u0 - (u0 > r ? 1 : 0), ud1 = u1 - (u1 > r ? 1 : 0);  // ... and where they end up
  const bool isy0 = (u0 == r), isy1 = (u1 == r);
  double a = GM_(r, r);                         // running (i, r) entry
  double carry0 = isy0 ? V[GV_Y + r] : (u0 < r ? GM_(r, u0) : 0.0), carry1 = 0.0;
  if (wide) carry1 = isy1 ? V[GV_Y + r] : (u1 < r ? GM_(r, u1) : 0.0);
  _Pragma("unroll 1") for (int i = r; i < k - 1; i++) {
    const double b = GM_(i + 1, r);
    const double h2 = fma(a, a, b * b);
    const double hinv = rsqrt(h2);
    const double c = b * hinv, sn = a * hinv;
    a = h2 * hinv;
    double e0 = 0.0, e1 = 0.0, n1 = 0.0;
    if (isy0) e0 = V[GV_Y + i + 1];
    else if (u0 < k && u0 <= i + 1) e0 = GM_(i + 1, u0);
    const double n0 = c * carry0 - sn * e0;
    carry0 = fma(sn, carry0, c * e0);
    if (wide) {
      if (isy1) e1 = V[GV_Y + i + 1];
      else if (u1 < k && u1 <= i + 1) e1 = GM_(i + 1, u1);
      n1 = c * carry1 - sn * e1;
      carry1 = fma(sn, carry1, c * e1);
    }
    // one barrier per rotation: row i was last read in the previous iteration (and M(r,r) before the
    // loop); it is overwritten now (a lane writes into its left neighbour's column)
    __syncwarp();
    if (isy0) V[GV_Y + i] = n0;
    else if (u0 < k && u0 <= i + 1) GM_(i, ud0) = n0;
    if (wide) {
      if (isy1) V[GV_Y + i] = n1;
      else if (u1 < k && u1 <= i + 1) GM_(i, ud1) = n1;
    }
  }
  __syncwarp();
}

// s = M'y for the current factorisation
template <int VS>
__device__ __forceinline__ void gram_solve_s(double *V, int ld, int k) {
  GV_LAYOUT(VS);
  const int lane = lane_id();
  double *T = V + GV_T;
  _Pragma("unroll 1") for (int u = lane; u < k; u += 32) {
    double a = 0.0;
    _Pragma("unroll 4") for (int t = u; t < k; t++) a = fma(GM_(t, u), V[GV_Y + t], a);
    V[GV_S + u] = a;
  }
  __syncwarp();
}

// Remove pivot position imv, then any other non-positive coefficient (first found), compacting P
// (src/NNLS.jl:735-778); the factorisation follows by Givens downdates and the removed columns leave `mask`.
// Returns the new number of active columns; s is up to date on return.
template <int VS>
__device__ __noinline__ int gram_remove(double *V, int ld, int k, int imv, unsigned long long &mask) {
  GV_LAYOUT(VS);
  __builtin_assume(__isShared(V));
  const int lane = lane_id();
  int *P = (int *)(V + GV_P);
  while (true) {
    gram_downdate<VS>(V, ld, k, imv);
    const bool m0 = lane >= imv && lane < k - 1, m1 = lane + 32 >= imv && lane + 32 < k - 1;
    int pn0 = 0, pn1 = 0;
    if (m0) pn0 = P[lane + 1];
    if (m1) pn1 = P[lane + 33];
    const int jrem = P[imv];
    mask &= ~(1ull << jrem);
    if (lane == 0) V[GV_X + jrem] = 0.0;
    __syncwarp();
    if (m0) P[lane] = pn0;
    if (m1) P[lane + 32] = pn1;
    __syncwarp();
    k -= 1;
    unsigned bad = 0x7fffffffu;
    _Pragma("unroll 1") for (int t = lane; t < k; t += 32)
      if (V[GV_X + P[t]] <= 0.0 && (unsigned)t < bad) bad = t;
    bad = __reduce_min_sync(DECAES_FULL_MASK, bad);
    if (bad == 0x7fffffffu) break;
    imv = (int)bad;
  }
  gram_solve_s<VS>(V, ld, k);
  return k;
}

// Heavily regularised solves (mu = e^2, the right end of the L-curve) keep nearly every column: instead of ~25 pivots
// one at a time, solve on the full set F directly and drop what comes out non-positive.  Left-looking Cholesky of
// K = G_FF + mu2 I in M's storage (L(t,u) where M(t,u) would be, 1/L(p,p) on the diagonal), the right-hand side riding
// along as an extra row (y = L^-1 c_F in GV_Y); lane <-> row, two rows per lane while more than 32 are left.  Back
// substitution in registers.  Positions with s <= 0 leave F and the factorisation resumes at the first of them:
// dropping the LAST pivots (the long-T2 columns, which is what happens) costs only the back substitution.  The result
// is accepted when the duals of the excluded columns are all <= 0 - a KKT point of a strictly convex problem is its
// minimiser; otherwise (or after 4 rounds, or on a non-positive pivot) the caller falls back to the active-set
// iteration, warm-started from what is left of F.
// Returns k (P[0:k), s in GV_S, x in GV_X, `mask` = F, xnorm_sq) or -1.
template <int VS>
__device__ __noinline__ int gram_dense_solve(double *V, int n, int ld, double mu2, unsigned long long &mask, double &xnorm_sq) {
  GV_LAYOUT(VS);
  __builtin_assume(__isShared(V));
  const int lane = lane_id();
  double *T = V + GV_T;
  int *P = (int *)(V + GV_P);
  const unsigned long long full = mask;
  int k = __popcll(mask), p0 = 0;
  double s0 = 0.0, s1 = 0.0;
  _Pragma("unroll 1") for (int round = 0;; round++) {
    if (round == 4 || k == 0) return -1;
    // pivot list in ascending column order (positions below p0 are unchanged)
    if (lane < n && ((mask >> lane) & 1ull)) P[__popcll(mask & ((1ull << lane) - 1ull))] = lane;
    if (lane + 32 < n && ((mask >> (lane + 32)) & 1ull)) P[__popcll(mask & ((1ull << (lane + 32)) - 1ull))] = lane + 32;
    __syncwarp();
    _Pragma("unroll 1") for (int p = p0; p < k; p++) {
      const int jp = P[p];
      const int i0 = p + lane, i1 = i0 + 32;  // rows of this lane; row k is the right-hand side
      double a0 = 0.0, a1 = 0.0;
      const double *r0 = V + GV_Y, *r1 = V + GV_Y;
      int st0 = 1, st1 = 1;
      if (i0 < k) a0 = gram_G(T, ld, P[i0], jp), r0 = T + i0 + 1, st0 = ld;
      else if (i0 == k) a0 = V[GV_C + jp];
      if (i1 < k) a1 = gram_G(T, ld, P[i1], jp), r1 = T + i1 + 1, st1 = ld;
      else if (i1 == k) a1 = V[GV_C + jp];
      if (lane == 0) a0 += mu2;
      const double *rp = T + p + 1;
      if (k - p >= 32) {  // warp-uniform: more than 32 rows (with the right-hand side) are left
        _Pragma("unroll 2") for (int q = 0; q < p; q++) {
          const double lp = rp[q * ld];
          a0 = fma(-r0[q * st0], lp, a0), a1 = fma(-r1[q * st1], lp, a1);
        }
      } else {
        _Pragma("unroll 2") for (int q = 0; q < p; q++) a0 = fma(-r0[q * st0], rp[q * ld], a0);
      }
      const double d = __shfl_sync(DECAES_FULL_MASK, a0, 0);
      if (!(d > 0.0)) return -1;
      const double dinv = rsqrt(d);
      // column p (and y_p) is written now; nobody reads it before the barrier below
      if (lane == 0) GM_(p, p) = dinv;
      else if (i0 < k) GM_(i0, p) = a0 * dinv;
      else if (i0 == k) V[GV_Y + p] = a0 * dinv;
      if (i1 < k) GM_(i1, p) = a1 * dinv;
      else if (i1 == k) V[GV_Y + p] = a1 * dinv;
      __syncwarp();
    }
    // back substitution s = L^-T y, y in registers (lane <-> positions lane, lane + 32)
    double y0 = lane < k ? V[GV_Y + lane] : 0.0, y1 = lane + 32 < k ? V[GV_Y + lane + 32] : 0.0;
    _Pragma("unroll 1") for (int p = k - 1; p >= 0; p--) {
      const double yp = __shfl_sync(DECAES_FULL_MASK, p < 32 ? y0 : y1, p & 31);
      const double sp = yp * GM_(p, p);
      if (lane == (p & 31)) {
        if (p < 32) s0 = sp;
        else s1 = sp;
      }
      if (lane < p) y0 = fma(-GM_(p, lane), sp, y0);
      if (lane + 32 < p) y1 = fma(-GM_(p, lane + 32), sp, y1);
    }
    const unsigned b0 = __ballot_sync(DECAES_FULL_MASK, lane < k && !(s0 > 0.0));
    const unsigned b1 = __ballot_sync(DECAES_FULL_MASK, lane + 32 < k && !(s1 > 0.0));
    const unsigned long long bad = ((unsigned long long)b1 << 32) | b0;
    if (bad == 0ull) break;
    unsigned long long drop = 0ull;
    if (lane < k && !(s0 > 0.0)) drop |= 1ull << P[lane];
    if (lane + 32 < k && !(s1 > 0.0)) drop |= 1ull << P[lane + 32];
    mask &= ~warp_or64(drop);
    p0 = __ffsll((long long)bad) - 1;
    k = __popcll(mask);
    __syncwarp();
  }
  // duals of the excluded columns: w_j = c_j - sum_t G(P[t], j) s_t must not be positive
  const int pj0 = lane < k ? P[lane] : 0, pj1 = lane + 32 < k ? P[lane + 32] : 0;
  _Pragma("unroll 1") for (unsigned long long ex = full & ~mask; ex; ex &= ex - 1ull) {
    const int j = __ffsll((long long)ex) - 1;
    double a = 0.0;
    if (lane < k) a = gram_G(T, ld, pj0, j) * s0;
    if (lane + 32 < k) a = fma(gram_G(T, ld, pj1, j), s1, a);
    if (V[GV_C + j] - warp_sum(a) > 0.0) return -1;
  }
  if (lane < n) V[GV_X + lane] = 0.0;
  if (lane + 32 < n) V[GV_X + lane + 32] = 0.0;
  __syncwarp();
  if (lane < k) V[GV_S + lane] = s0, V[GV_X + pj0] = s0;
  if (lane + 32 < k) V[GV_S + lane + 32] = s1, V[GV_X + pj1] = s1;
  double q2 = 0.0;
  if (lane < k) q2 = s0 * s0;
  if (lane + 32 < k) q2 = fma(s1, s1, q2);
  xnorm_sq = warp_sum(q2);
  __syncwarp();
  return k;
}

// Lawson–Hanson main loop.  cold: start from the empty set with the reference's warm dual
// (src/lsqnonneg.jl:44-70).  warm: start from the feasible point x supported on `mask`.
template <int VS>
__device__ __noinline__ GramOut gram_nnls(double *V, int n, int ld, double mu2, int max_set, bool warm, unsigned long long mask) {
  GV_LAYOUT(VS);
  __builtin_assume(__isShared(V));
  const int lane = lane_id();
  double *T = V + GV_T;
  int *P = (int *)(V + GV_P);
  GP_BEGIN();
  GP_ADD(warm ? 11 : 12, 1);
  int k = 0, iter = 0, nappend = 0;
  bool hit_cap = false;
  const int max_iter = 3 * n;
  bool check_first = false;
  // the two columns of this lane (n <= 64); out-of-range ones are clamped and never win
  const int j0 = lane < n ? lane : 0, j1 = lane + 32 < n ? lane + 32 : 0;
  const bool v0 = lane < n, v1 = lane + 32 < n;
  unsigned long long key = 0ull;  // this lane's best positive dual (bit pattern) and its column
  int bj = 0x7fffffff;
  bool have_pick = false;

  if (!warm) {
    mask = 0ull;
    // dual as if the last column were active; w[n-1] = 0, or 1 if every other dual is <= 0
    const double xj = ddiv(V[GV_C + n - 1], T[(n - 1) * ld + (n - 1)] + mu2);
    bool anypos = false;
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) {
      const double wj = (j < n - 1) ? fma(-T[(n - 1) * ld + j], xj, V[GV_C + j]) : 0.0;
      V[GV_W + j] = wj;
      V[GV_X + j] = 0.0;
      anypos |= !(wj <= 0.0);
    }
    if (!__any_sync(DECAES_FULL_MASK, anypos) && lane == 0) V[GV_W + n - 1] = 1.0;
    __syncwarp();
  } else {
    // factor the inherited set (ascending column order); x outside the set is zero
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32)
      if ((mask >> j) & 1ull) {
        P[__popcll(mask & ((1ull << j) - 1ull))] = j;
      } else {
        V[GV_X + j] = 0.0;
      }
    __syncwarp();
    const int kk = __popcll(mask);
    k = gram_refactor<VS>(V, ld, kk, mu2);
    if (k != kk) mask = mask_of(P, k);  // a column was dropped as dependent
    check_first = (k > 0);
    if (k == 0) {
      _Pragma("unroll 1") for (int j = lane; j < n; j += 32) V[GV_W + j] = V[GV_C + j];
      __syncwarp();
    }
  }

  while (true) {
    if (!check_first) {
      if (k >= max_set) break;
      // ---- entering column: largest positive dual, first on ties; test its coefficient ----
      if (!have_pick) {  // duals in shared memory (initial dual, or after a rejection)
        key = 0ull, bj = 0x7fffffff;
        _Pragma("unroll 1") for (int j = lane; j < n; j += 32) {
          const double v = V[GV_W + j];
          const unsigned long long kb = (unsigned long long)__double_as_longlong(v);
          if (!((mask >> j) & 1ull) && v > 0.0 && kb > key) key = kb, bj = j;
        }
      }
      have_pick = false;
      unsigned long long best;
      bj = warp_argmax_bits(key, bj, best);
      if (best == 0ull) break;  // no positive dual left: KKT point
      if (!gram_append<VS>(V, ld, k, bj, mu2, true)) {
        if (lane == 0) V[GV_W + bj] = 0.0;  // rejected (src/NNLS.jl:652-657)
        __syncwarp();
        continue;
      }
      k += 1;
      nappend += 1;
      mask |= 1ull << bj;
    }
    check_first = false;

    // ---- secondary loop: keep the iterate feasible (src/NNLS.jl:692-786) ----
    bool capped = false;
    while (true) {
      iter += 1;
      if (iter > max_iter) {
        capped = true;
        break;
      }
      unsigned long long fkey = ~0ull, best;
      int imv = 0x7fffffff;
      _Pragma("unroll 1") for (int t = lane; t < k; t += 32) {
        const double st = V[GV_S + t];
        if (st <= 0.0) {
          const double xi = V[GV_X + P[t]];
          const double tt = ddiv(-xi, st - xi);
          const unsigned long long kb = (unsigned long long)__double_as_longlong(tt);
          if (tt < 2.0 && kb < fkey) fkey = kb, imv = t;  // tt >= 0 here, so the bit pattern orders like the value
        }
      }
      if (__all_sync(DECAES_FULL_MASK, fkey == ~0ull)) break;  // all coefficients feasible
      imv = warp_argmin_bits(fkey, imv, best);
      const double al = __longlong_as_double((long long)best);
      _Pragma("unroll 1") for (int t = lane; t < k; t += 32) {
        const int jx = P[t];
        V[GV_X + jx] = fma(al, V[GV_S + t] - V[GV_X + jx], V[GV_X + jx]);
      }
      __syncwarp();
      k = gram_remove<VS>(V, ld, k, imv, mask);
    }
    if (capped) {
      hit_cap = true;
      break;
    }

    _Pragma("unroll 1") for (int t = lane; t < k; t += 32) V[GV_X + P[t]] = V[GV_S + t];
    __syncwarp();
    // ---- dual: w_j = c_j - sum_t G(P[t], j) s_t, zero on the active set; both columns of the lane
    //      advance together and the lane's candidate for the next pivot is picked on the fly ----
    {
      GP_BEGIN();
      double a0 = V[GV_C + j0], a1 = V[GV_C + j1];
      _Pragma("unroll 4") for (int t = 0; t < k; t++) {
        const int pt = P[t];
        const double st = V[GV_S + t];
        a0 = fma(-gram_G(T, ld, pt, j0), st, a0);
        a1 = fma(-gram_G(T, ld, pt, j1), st, a1);
      }
      if (!v0 || ((mask >> j0) & 1ull)) a0 = 0.0;
      if (!v1 || ((mask >> j1) & 1ull)) a1 = 0.0;
      if (v0) V[GV_W + j0] = a0;
      if (v1) V[GV_W + j1] = a1;
      const unsigned long long k0 = a0 > 0.0 ? (unsigned long long)__double_as_longlong(a0) : 0ull;
      const unsigned long long k1 = a1 > 0.0 ? (unsigned long long)__double_as_longlong(a1) : 0ull;
      key = k1 > k0 ? k1 : k0, bj = k1 > k0 ? j1 : (k0 ? j0 : 0x7fffffff);
      have_pick = true;
      __syncwarp();
      GP_END(2);
    }
  }

  GramOut o;
  o.k = k;
  o.mask = mask;
  double acc = 0.0;
  _Pragma("unroll 1") for (int j = lane; j < n; j += 32) acc = fma(V[GV_X + j], V[GV_X + j], acc);
  o.xnorm_sq = warp_sum(acc);
  o.iters = iter, o.nappend = nappend, o.capped = hit_cap;
  GP_ADD(13, k);
  GP_ADD(14, iter);
  GP_END(3);
  return o;
}

}  // namespace decaes
