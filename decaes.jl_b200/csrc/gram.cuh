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
// so that appending a column is two O(k) fma chains per lane plus two warp reductions, the
// solution is s = M'(M c_P) maintained incrementally, and the dual is w = c - G_P x_P.
// The minimiser of the Tikhonov problems (mu > 0) is unique, so those solves may be warm-started
// from the active set of the nearest mu already solved; unregularised solves follow the
// reference's cold-start path pivot by pivot and are polished with one step of iterative
// refinement on the explicit residual (done by the caller, which owns A).
//
// Shared-memory layout: ONE n x ld array T (ld = (n+1)|1, odd => both access directions are
// bank-conflict free) holds the lower triangle of G and, in the strictly-upper part, M:
//     G(p,q) = T[max(p,q)*ld + min(p,q)]          M(t,u) = T[u*ld + t + 1]   (u <= t)
#pragma once
#include "common.cuh"

namespace decaes {

#ifdef DECAES_PROFILE
__device__ unsigned long long g_prof[16];  // [0..3] cycles append/rebuild/dual/nnls, [4..7] calls, [8] sum k at append
#define GP_BEGIN() long long gp_t0 = clock64()
#define GP_END(id) if (lane_id() == 0) { atomicAdd(&g_prof[id], (unsigned long long)(clock64() - gp_t0)); atomicAdd(&g_prof[4 + id], 1ull); }
#else
#define GP_BEGIN()
#define GP_END(id)
#endif

struct GramProb {
  double *T;        // combined G / M array in shared memory
  int ld;
  const double *c;  // [n] shared memory
  double mu2;       // mu^2 (0 for the plain problem)
  int n;
  int max_set;      // min(m, n) for the plain problem, n for Tikhonov (src/NNLS.jl:627, :851)
};

struct GramWs {  // shared-memory scratch of one warp
  double *y;     // [n] y = M c_P
  double *s;     // [n] s = M' y   (solution on P, in pivot order)
  double *x;     // [n] current feasible solution, indexed by column
  double *w;     // [n] dual, indexed by column
  double *t1;    // [n] scratch
  double *t2;    // [n] scratch
  int *P;        // [n] active columns in pivot order
};

struct GramOut {
  int k;                     // number of active columns
  unsigned long long mask;   // active set as a bit mask
  double xnorm_sq;           // sum of squares of the solution
};

__device__ __forceinline__ double gram_G(const GramProb &p, int a, int b) {
  int hi = a > b ? a : b, lo = a > b ? b : a;
  return p.T[hi * p.ld + lo];
}
#define GM_(t, u) T[(u) * ld + (t) + 1]

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
  unsigned lo = __reduce_or_sync(DECAES_FULL_MASK, (unsigned)v);
  unsigned hi = __reduce_or_sync(DECAES_FULL_MASK, (unsigned)(v >> 32));
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long mask_of(const int *P, int k) {
  const int lane = lane_id();
  unsigned long long m = 0ull;
  if (lane < k) m |= 1ull << P[lane];
  if (lane + 32 < k) m |= 1ull << P[lane + 32];
  return warp_or64(m);
}

// Append column j to the factorisation (pivot position k).  Returns false (and changes nothing)
// when the column is numerically dependent (d^2 <= 0) or, with `need_positive`, when its
// coefficient would not be positive (the reference's b1/A1 > 0 test).
__device__ __forceinline__ bool gram_append(const GramProb &p, const GramWs &ws, int &k, int j, bool need_positive) {
  const int lane = lane_id();
  double *T = p.T;
  const int ld = p.ld;
  _Pragma("unroll 1") for (int t = lane; t < k; t += 32) ws.t1[t] = gram_G(p, ws.P[t], j);
  __syncwarp();
  // l = M g  (lane <-> row t; for fixed u the lanes read consecutive words)
  double ll = 0.0, ly = 0.0;
  _Pragma("unroll 1") for (int t = lane; t < k; t += 32) {
    double a0 = 0.0, a1 = 0.0;
    int u = 0;
#pragma unroll 1
    for (; u + 1 <= t; u += 2) {
      a0 = fma(GM_(t, u), ws.t1[u], a0);
      a1 = fma(GM_(t, u + 1), ws.t1[u + 1], a1);
    }
    if (u <= t) a0 = fma(GM_(t, u), ws.t1[u], a0);
    double lt = a0 + a1;
    ws.t2[t] = lt;
    ll = fma(lt, lt, ll);
    ly = fma(lt, ws.y[t], ly);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {  // both sums ride the same butterfly
    ll += __shfl_xor_sync(DECAES_FULL_MASK, ll, o);
    ly += __shfl_xor_sync(DECAES_FULL_MASK, ly, o);
  }
  const double d2 = (T[j * ld + j] + p.mu2) - ll;
  if (!(d2 > 0.0)) return false;
  const double dinv = rsqrt(d2);
  const double ynew = (p.c[j] - ly) * dinv;
  if (need_positive && !(ynew > 0.0)) return false;
  __syncwarp();
  // new row of M: M(k,u) = -dinv * sum_{t >= u} l_t M(t,u)   (lane <-> column u, own row contiguous in t)
  _Pragma("unroll 1") for (int u = lane; u < k; u += 32) {
    double a0 = 0.0, a1 = 0.0;
    int t = u;
#pragma unroll 1
    for (; t + 1 < k; t += 2) {
      a0 = fma(ws.t2[t], GM_(t, u), a0);
      a1 = fma(ws.t2[t + 1], GM_(t + 1, u), a1);
    }
    if (t < k) a0 = fma(ws.t2[t], GM_(t, u), a0);
    double mu_ = -dinv * (a0 + a1);
    GM_(k, u) = mu_;
    ws.s[u] = fma(ynew, mu_, ws.s[u]);
  }
  if (lane == 0) {
    GM_(k, k) = dinv;
    ws.y[k] = ynew;
    ws.s[k] = ynew * dinv;
    ws.P[k] = j;
  }
  __syncwarp();
  k += 1;
  return true;
}

// Rebuild rows [from, k) of M (after a removal or for a warm start); P[0:k) already holds the columns.
__device__ __forceinline__ void gram_rebuild(const GramProb &p, const GramWs &ws, int &k, int from) {
  const int lane = lane_id();
  const int kold = k, ld = p.ld;
  double *T = p.T;
  // s = M[0:from]' y[0:from]
  _Pragma("unroll 1") for (int u = lane; u < kold; u += 32) {
    double a = 0.0;
    for (int t = u; t < from; t++) a = fma(GM_(t, u), ws.y[t], a);
    ws.s[u] = a;
  }
  __syncwarp();
  k = from;
  for (int t = from; t < kold; t++) {
    int j = ws.P[t];
    __syncwarp();
    if (!gram_append(p, ws, k, j, false)) {
      // numerically dependent column inside a set that was independent a moment ago: drop it
      if (lane == 0) ws.x[j] = 0.0;
      __syncwarp();
    }
  }
}


// w_j = c_j - sum_t G(P[t], j) * s_t for every column; entries of active columns are forced to 0.
__device__ __forceinline__ void gram_dual(const GramProb &p, const GramWs &ws, int k, unsigned long long mask) {
  const int lane = lane_id();
  _Pragma("unroll 1") for (int j = lane; j < p.n; j += 32) {
    double a0 = p.c[j], a1 = 0.0;
    int t = 0;
#pragma unroll 1
    for (; t + 1 < k; t += 2) {
      a0 = fma(-gram_G(p, ws.P[t], j), ws.s[t], a0);
      a1 = fma(-gram_G(p, ws.P[t + 1], j), ws.s[t + 1], a1);
    }
    if (t < k) a0 = fma(-gram_G(p, ws.P[t], j), ws.s[t], a0);
    ws.w[j] = ((mask >> j) & 1ull) ? 0.0 : a0 + a1;
  }
  __syncwarp();
}

// Lawson–Hanson main loop.  cold: start from the empty set with the reference's warm dual
// (src/lsqnonneg.jl:44-70).  warm: start from the feasible point ws.x supported on `mask`.
__device__ __noinline__ GramOut gram_nnls(const GramProb &p_in, const GramWs &ws_in, bool warm, unsigned long long mask) {
  const GramProb p = p_in;  // copies: keep every pointer in registers instead of reloading it from the caller's frame
  const GramWs ws = ws_in;
  // every one of these lives in shared memory: without the hint the compiler emits generic LD/ST
  // (plus uniform-register descriptor shuffling) instead of LDS/STS in the hottest loops
  __builtin_assume(__isShared(p.T));
  __builtin_assume(__isShared(p.c));
  __builtin_assume(__isShared(ws.y));
  __builtin_assume(__isShared(ws.s));
  __builtin_assume(__isShared(ws.x));
  __builtin_assume(__isShared(ws.w));
  __builtin_assume(__isShared(ws.t1));
  __builtin_assume(__isShared(ws.t2));
  __builtin_assume(__isShared(ws.P));
  const int lane = lane_id();
  const int n = p.n;
  int k = 0, iter = 0;
  const int max_iter = 3 * n;
  bool need_solve_check = false;

  if (!warm) {
    mask = 0ull;
    // dual as if the last column were active; w[n-1] = 0, or 1 if every other dual is <= 0
    const double xj = ddiv(p.c[n - 1], p.T[(n - 1) * p.ld + (n - 1)] + p.mu2);
    bool anypos = false;
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) {
      double wj = (j < n - 1) ? fma(-gram_G(p, n - 1, j), xj, p.c[j]) : 0.0;
      ws.w[j] = wj;
      ws.x[j] = 0.0;
      anypos |= !(wj <= 0.0);
    }
    if (!__any_sync(DECAES_FULL_MASK, anypos) && lane == 0) ws.w[n - 1] = 1.0;
    __syncwarp();
  } else {
    // factor the inherited set (ascending column order)
    unsigned long long m2 = mask;
    int kk = 0;
    while (m2) {
      int j = __ffsll((long long)m2) - 1;
      m2 &= m2 - 1;
      if (lane == 0) ws.P[kk] = j;
      kk++;
    }
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32)
      if (!((mask >> j) & 1ull)) ws.x[j] = 0.0;
    __syncwarp();
    k = kk;
    gram_rebuild(p, ws, k, 0);
    if (k != kk) mask = mask_of(ws.P, k);  // a column was dropped as dependent
    need_solve_check = (k > 0);
    if (k == 0) {
      _Pragma("unroll 1") for (int j = lane; j < n; j += 32) ws.w[j] = p.c[j];
      __syncwarp();
    }
  }

  bool terminated = false;
  while (true) {
    if (!need_solve_check) {
      if (k >= p.max_set) break;
      // ---- entering column: largest positive dual, first on ties; test its coefficient ----
      bool accepted = false;
      while (true) {
        double best = 0.0;
        int bj = 0x7fffffff;
        _Pragma("unroll 1") for (int j = lane; j < n; j += 32) {
          double v = ws.w[j];
          if (!((mask >> j) & 1ull) && v > best) best = v, bj = j;
        }
        warp_argmax_first(best, bj);
        if (!(best > 0.0)) {
          terminated = true;
          break;
        }
        if (gram_append(p, ws, k, bj, true)) {
          mask |= 1ull << bj;
          accepted = true;
          break;
        }
        if (lane == 0) ws.w[bj] = 0.0;  // rejected (src/NNLS.jl:652-657)
        __syncwarp();
      }
      if (terminated || !accepted) break;
    }
    need_solve_check = false;

    // ---- secondary loop: keep the iterate feasible (src/NNLS.jl:692-786) ----
    while (true) {
      iter += 1;
      if (iter > max_iter) {
        terminated = true;
        break;
      }
      double al = 2.0;
      int imv = 0x7fffffff;
      _Pragma("unroll 1") for (int t = lane; t < k; t += 32) {
        double st = ws.s[t];
        if (st <= 0.0) {
          double xi = ws.x[ws.P[t]];
          double tt = ddiv(-xi, st - xi);
          if (al > tt) al = tt, imv = t;
        }
      }
      warp_argmin_first(al, imv);
      if (!(al < 2.0)) break;  // all coefficients feasible
      _Pragma("unroll 1") for (int t = lane; t < k; t += 32) {
        int jx = ws.P[t];
        ws.x[jx] = fma(al, ws.s[t] - ws.x[jx], ws.x[jx]);
      }
      __syncwarp();
      // remove imv, then any other non-positive coefficient (first found), compacting P
      int first_removed = imv;
      while (true) {
        if (lane == 0) {
          int jr = ws.P[imv];
          ws.x[jr] = 0.0;
          for (int t = imv; t < k - 1; t++) ws.P[t] = ws.P[t + 1];
        }
        __syncwarp();
        k -= 1;
        if (imv < first_removed) first_removed = imv;
        unsigned bad0 = __ballot_sync(DECAES_FULL_MASK, lane < k && ws.x[ws.P[lane]] <= 0.0);
        unsigned bad1 = __ballot_sync(DECAES_FULL_MASK, lane + 32 < k && ws.x[ws.P[lane + 32]] <= 0.0);
        if (bad0) imv = __ffs(bad0) - 1;
        else if (bad1) imv = 32 + __ffs(bad1) - 1;
        else break;
      }
      gram_rebuild(p, ws, k, first_removed);
      mask = mask_of(ws.P, k);
    }
    if (terminated) break;

    _Pragma("unroll 1") for (int t = lane; t < k; t += 32) ws.x[ws.P[t]] = ws.s[t];
    __syncwarp();
    gram_dual(p, ws, k, mask);
  }

  GramOut o;
  o.k = k;
  o.mask = mask;
  double acc = 0.0;
  _Pragma("unroll 1") for (int j = lane; j < n; j += 32) acc = fma(ws.x[j], ws.x[j], acc);
  o.xnorm_sq = warp_sum(acc);
  return o;
}

}  // namespace decaes
