// Warp-cooperative Lawson–Hanson NNLS (plain and Tikhonov-structured) for one voxel per warp.
//
// Algorithm: src/NNLS.jl:605-1061 (unsafe_nnls!) with the helpers :259-336, :379-436, :441-470,
// :486-554 and the warm-started drivers src/lsqnonneg.jl:30-164 — same pivot rule, same
// accept/reject test, same lazy introduction of the lambda rows, same Givens down-dating.
//
// B200 mapping: the working matrix lives in shared memory ROW-major with an odd leading
// dimension, so both access patterns are bank-conflict free:
//   lane <-> column j  (Householder application, dual recomputation, Givens): A[i*ld + j], j contiguous
//   lane <-> row i     (column norms, swaps, back-substitution):              A[i*ld + j], stride ld (odd)
// Reductions over rows are warp-shuffle butterflies; scalars (nsetp, m, iter) stay in registers
// and control flow is warp-uniform.  All indices are 0-based here.
#pragma once
#include "common.cuh"

namespace decaes {

struct NnlsWs {
  double *A;    // [(m0 + n) * ld]  (plain problems only use m0 rows)
  double *b;    // [m0 + n]
  double *u;    // [m0 + n]  Householder vector scratch / zz
  double *x;    // [n]
  double *w;    // [n]
  int *idx;     // [n] original column of each position
  int ld, n, m0;
};

struct NnlsOut {
  double rnorm_sq;  // ||b[nsetp:M]||^2
  double xnorm_sq;  // sum of squares of the positive solution
  int nsetp;
  int rows_used;    // final m (rows touched) — lets the caller zero only what is dirty
};

// dual of the trailing columns from the current rotated system: w[j] = sum_{i=r0}^{m1-1} A[i][j] b[i]
// (compute_dual!, src/NNLS.jl:441-470)
__device__ __noinline__ void nnls_compute_dual(const NnlsWs &s, int j0, int r0, int m1) {
  const int lane = lane_id();
  for (int j = j0 + lane; j < s.n; j += 32) {
    double s0 = 0.0, s1 = 0.0;
    int i = r0;
    for (; i + 1 < m1; i += 2) {
      s0 = fma(s.A[i * s.ld + j], s.b[i], s0);
      s1 = fma(s.A[(i + 1) * s.ld + j], s.b[i + 1], s1);
    }
    if (i < m1) s0 = fma(s.A[i * s.ld + j], s.b[i], s0);
    s.w[j] = s0 + s1;
  }
}

// back-substitution R z = b[0:k]  (solve_triangular_system!, src/NNLS.jl:517-523).
// lane <-> row; z is returned in registers: z0 = z[lane], z1 = z[lane + 32].
__device__ __noinline__ void nnls_backsolve(const NnlsWs &s, int k, double &z0, double &z1) {
  const int lane = lane_id();
  z0 = (lane < k) ? s.b[lane] : 0.0;
  z1 = (lane + 32 < k) ? s.b[lane + 32] : 0.0;
  for (int j = k - 1; j >= 0; j--) {
    double zj = (j < 32) ? warp_bcast(z0, j) : warp_bcast(z1, j - 32);
    double q = zj / s.A[j * s.ld + j];
    if (lane < j) z0 = fma(-s.A[lane * s.ld + j], q, z0);
    if (lane + 32 < j) z1 = fma(-s.A[(lane + 32) * s.ld + j], q, z1);
    if (lane == (j & 31)) {
      if (j < 32) z0 = q; else z1 = q;
    }
  }
}

// The solver.  On entry: rows [0, m0) of s.A hold the matrix, s.b[0:m0) the right-hand side,
// s.w the initial dual, s.x = 0, s.idx = identity; for tikh, rows [m0, m0+n) of A and b are 0.
template <bool TIKH>
__device__ __noinline__ NnlsOut nnls_core(const NnlsWs &s, double lambda) {
  const int lane = lane_id();
  const int n = s.n, ld = s.ld;
  const int M = TIKH ? s.m0 + n : s.m0;
  int m = s.m0;
  int nsetp = 0, iter = 0;
  const int max_iter = 3 * n;
  unsigned long long diag = 0ull;  // bit c set <=> lambda row of original column c is in the system
  bool terminated = false;
  double *A = s.A, *b = s.b, *u = s.u, *x = s.x, *w = s.w;
  int *idx = s.idx;

  while (true) {
    if (TIKH ? (nsetp >= n) : (nsetp >= n || nsetp >= m)) break;

    int jmax = -1;
    double tau = 0.0;
    // ---- pick the entering column (largest positive dual, first on ties) and test it ----
    while (true) {
      double best = 0.0;
      int bj = 0x7fffffff;
      for (int j = nsetp + lane; j < n; j += 32) {
        double v = w[j];
        if (v > best) best = v, bj = j;
      }
      warp_argmax_first(best, bj);
      if (!(best > 0.0)) {
        terminated = true;
        break;
      }
      jmax = bj;
      const int ip = nsetp;
      const bool fresh = TIKH && !((diag >> idx[jmax]) & 1ull);
      int m1 = m;  // number of rows taking part in the reflection
      if (TIKH) {
        m1 = (m + 1 < M) ? m + 1 : M;
        if (fresh && lane == 0) A[m * ld + jmax] = lambda;  // src/NNLS.jl:869-871
        __syncwarp();
      }
      // construct_apply_householder!  src/NNLS.jl:259-336
      double acc = 0.0;
      for (int i = ip + lane; i < m1; i += 32) {
        double a = A[i * ld + jmax];
        acc = fma(a, a, acc);
      }
      double xnorm = sqrt(warp_sum(acc));
      bool good = false;
      double sm = 0.0, A1 = 0.0, b1 = 0.0;
      if (ip >= m1) {
        tau = 0.0;  // (ip > m) branch: nothing to do, treated as accepted with tau = 0
        good = true;
      } else if (xnorm != 0.0) {
        double alpha = A[ip * ld + jmax];
        double beta = copysign(xnorm, alpha);
        alpha = alpha + beta;
        tau = alpha / beta;
        acc = 0.0;
        for (int i = ip + 1 + lane; i < m1; i += 32) {
          double ui = A[i * ld + jmax] / alpha;
          u[i] = ui;
          acc = fma(b[i], ui, acc);
        }
        sm = b[ip] + warp_sum(acc);
        sm *= -tau;
        A1 = -beta;
        b1 = b[ip] + sm;
        good = (b1 / A1 > 0.0);
      }
      __syncwarp();
      if (good) {
        if (ip < m1) {
          if (ip + 1 < m1) {
            // swap columns ip <-> jmax, install the reflected column and update b
            for (int i = lane; i < m1; i += 32) {
              double old_ip = A[i * ld + ip];
              if (i < ip) {
                if (ip != jmax) {
                  A[i * ld + ip] = A[i * ld + jmax];
                  A[i * ld + jmax] = old_ip;
                }
              } else if (i == ip) {
                A[i * ld + ip] = A1;
                if (ip != jmax) A[i * ld + jmax] = old_ip;
                b[i] = b1;
              } else {
                // below the diagonal the reference stores u and zeroes it after the reflection
                // (src/NNLS.jl:676-678); u is kept in scratch here so the zero is written at once
                A[i * ld + ip] = 0.0;
                if (ip != jmax) A[i * ld + jmax] = old_ip;
                b[i] = fma(sm, u[i], b[i]);
              }
            }
          } else {
            tau = 0.0;  // ip == m: plain column swap (src/NNLS.jl:324-331)
            if (ip != jmax)
              for (int i = lane; i < m1; i += 32) {
                double t = A[i * ld + ip];
                A[i * ld + ip] = A[i * ld + jmax];
                A[i * ld + jmax] = t;
              }
          }
        }
        __syncwarp();
        break;
      }
      // rejected: w[j] = 0, undo the lambda entry (src/NNLS.jl:880-888)
      if (lane == 0) {
        w[jmax] = 0.0;
        if (TIKH && m < M) A[m * ld + jmax] = 0.0;
      }
      __syncwarp();
    }
    if (terminated) break;

    // ---- move the column into set P ----
    if (TIKH) {
      int c = idx[jmax];
      if (!((diag >> c) & 1ull)) {
        m = (m + 1 < M) ? m + 1 : M;
        diag |= (1ull << c);
      }
    }
    __syncwarp();
    if (lane == 0) {
      int t = idx[nsetp];
      idx[nsetp] = idx[jmax];
      idx[jmax] = t;
    }
    nsetp += 1;
    const int j1 = nsetp - 1;

    // apply_householder_dual!  src/NNLS.jl:379-436: reflect the trailing columns and
    // recompute their duals in the same sweep.  lane <-> column.
    if (nsetp < n && j1 + 1 < m) {
      const double ntau = -tau;
      for (int j = nsetp + lane; j < n; j += 32) {
        double *col = A + j;
        double s0 = col[j1 * ld], s1 = 0.0;
        int i = j1 + 1;
        for (; i + 1 < m; i += 2) {
          s0 = fma(col[i * ld], u[i], s0);
          s1 = fma(col[(i + 1) * ld], u[i + 1], s1);
        }
        if (i < m) s0 = fma(col[i * ld], u[i], s0);
        double smj = (s0 + s1) * ntau;
        col[j1 * ld] += smj;
        double w0 = 0.0, w1 = 0.0;
        i = j1 + 1;
        for (; i + 1 < m; i += 2) {
          double a0 = fma(smj, u[i], col[i * ld]);
          double a1 = fma(smj, u[i + 1], col[(i + 1) * ld]);
          w0 = fma(a0, b[i], w0);
          w1 = fma(a1, b[i + 1], w1);
          col[i * ld] = a0;
          col[(i + 1) * ld] = a1;
        }
        if (i < m) {
          double a0 = fma(smj, u[i], col[i * ld]);
          w0 = fma(a0, b[i], w0);
          col[i * ld] = a0;
        }
        w[j] = w0 + w1;
      }
    }
    if (lane == 0) w[j1] = 0.0;
    __syncwarp();

    // ---- solve the triangular system; secondary loop ----
    double z0, z1;
    nnls_backsolve(s, nsetp, z0, z1);
    bool dual_flag = false;
    while (true) {
      iter += 1;
      if (iter > max_iter) {
        terminated = true;  // mode = 1 in the reference; the current x is returned
        break;
      }
      // feasibility: alpha = min over zz[i] <= 0 of -x/(zz - x), first minimiser
      double al = 2.0;
      int imv = 0x7fffffff;
      if (lane < nsetp && z0 <= 0.0) {
        double xi = x[idx[lane]];
        al = -xi / (z0 - xi), imv = lane;
      }
      if (lane + 32 < nsetp && z1 <= 0.0) {
        double xi = x[idx[lane + 32]];
        double t = -xi / (z1 - xi);
        if (al > t) al = t, imv = lane + 32;
      }
      // note: a lane-local NaN/>=2 candidate never wins, as in the sequential scan
      if (!(al < 2.0)) al = 2.0, imv = 0x7fffffff;
      warp_argmin_first(al, imv);
      if (al == 2.0) break;
      dual_flag = true;

      if (lane < nsetp) {
        int ix = idx[lane];
        x[ix] = fma(al, z0 - x[ix], x[ix]);
      }
      if (lane + 32 < nsetp) {
        int ix = idx[lane + 32];
        x[ix] = fma(al, z1 - x[ix], x[ix]);
      }
      __syncwarp();

      // move coefficient imv from set P to set Z (src/NNLS.jl:731-779)
      while (true) {
        if (lane == 0) x[idx[imv]] = 0.0;
        if (imv != nsetp - 1) {
          for (int i = imv + 1; i < nsetp; i++) {
            double p = A[(i - 1) * ld + i], q = A[i * ld + i];
            double sig = hypot_julia(p, q);
            double cc = p / sig, ss = q / sig;
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
              double a0 = A[(i - 1) * ld + j], a1 = A[i * ld + j];
              if (j == i) {
                A[(i - 1) * ld + j] = sig;
                A[i * ld + j] = 0.0;
              } else {
                A[(i - 1) * ld + j] = fma(cc, a0, __dmul_rn(ss, a1));
                A[i * ld + j] = fma(-ss, a0, __dmul_rn(cc, a1));
              }
            }
            if (lane == 0) {
              double b0 = b[i - 1], b1v = b[i];
              b[i - 1] = fma(cc, b0, __dmul_rn(ss, b1v));
              b[i] = fma(-ss, b0, __dmul_rn(cc, b1v));
            }
            __syncwarp();
          }
          // cyclic shift of columns imv..nsetp-1 (equivalent to the adjacent swaps :753-758)
          for (int i = lane; i < m; i += 32) {
            double *row = A + i * ld;
            double t = row[imv];
            for (int j = imv; j < nsetp - 1; j++) row[j] = row[j + 1];
            row[nsetp - 1] = t;
          }
          if (lane == 0) {
            int t = idx[imv];
            for (int j = imv; j < nsetp - 1; j++) idx[j] = idx[j + 1];
            idx[nsetp - 1] = t;
          }
        }
        __syncwarp();
        nsetp -= 1;
        // any remaining non-positive coefficient is removed as well (first one found)
        unsigned bad0 = __ballot_sync(DECAES_FULL_MASK, lane < nsetp && x[idx[lane]] <= 0.0);
        unsigned bad1 = __ballot_sync(DECAES_FULL_MASK, lane + 32 < nsetp && x[idx[lane + 32]] <= 0.0);
        if (bad0) imv = __ffs(bad0) - 1;
        else if (bad1) imv = 32 + __ffs(bad1) - 1;
        else break;
      }
      nnls_backsolve(s, nsetp, z0, z1);
    }
    if (terminated) break;

    if (dual_flag) nnls_compute_dual(s, nsetp, nsetp, m);
    if (lane < nsetp) x[idx[lane]] = z0;
    if (lane + 32 < nsetp) x[idx[lane + 32]] = z1;
    __syncwarp();
  }

  // residual norm over rows nsetp..M-1 (rows >= m are zero)  src/NNLS.jl:810-823 / :1046-1059
  NnlsOut out;
  double acc = 0.0;
  for (int i = nsetp + lane; i < m; i += 32) acc = fma(b[i], b[i], acc);
  out.rnorm_sq = warp_sum(acc);
  acc = 0.0;
  for (int i = lane; i < nsetp; i += 32) {
    double xi = x[idx[i]];
    acc = fma(xi, xi, acc);
  }
  out.xnorm_sq = warp_sum(acc);
  out.nsetp = nsetp;
  out.rows_used = m;
  return out;
}

// Warm-started dual of src/lsqnonneg.jl:44-83 / :115-161: computed as if the LAST column were
// already active; w[n-1] is 0, or 1 when every other dual is <= 0.  The matrix is already in
// s.A (rows [0,m0)), the data in `bsrc`.  Also resets x, idx, b and the lambda rows
// [m0, m0 + zero_rows).
template <bool TIKH>
__device__ __noinline__ void nnls_warm_start(const NnlsWs &s, const double *bsrc, double mu, int zero_rows) {
  const int lane = lane_id();
  const int n = s.n, ld = s.ld, m0 = s.m0;
  double acc = 0.0;
  for (int i = lane; i < m0; i += 32) {
    double a = s.A[i * ld + n - 1];
    acc = fma(a, a, acc);
  }
  double den = warp_sum(acc);
  if (TIKH) den += mu * mu;
  acc = 0.0;
  for (int i = lane; i < m0; i += 32) acc = fma(s.A[i * ld + n - 1] / den, bsrc[i], acc);
  double xj = warp_sum(acc);
  for (int i = lane; i < m0; i += 32) {
    double bi = bsrc[i];
    s.u[i] = bi - __dmul_rn(s.A[i * ld + n - 1], xj);
    s.b[i] = bi;
  }
  __syncwarp();
  bool anypos = false;
  for (int j = lane; j < n; j += 32) {
    double s0 = 0.0, s1 = 0.0;
    if (j < n - 1) {
      int i = 0;
      for (; i + 1 < m0; i += 2) {
        s0 = fma(s.A[i * ld + j], s.u[i], s0);
        s1 = fma(s.A[(i + 1) * ld + j], s.u[i + 1], s1);
      }
      if (i < m0) s0 = fma(s.A[i * ld + j], s.u[i], s0);
    }
    double wj = s0 + s1;
    s.w[j] = wj;
    anypos |= !(wj <= 0.0);
    s.x[j] = 0.0;
    s.idx[j] = j;
  }
  if (!__any_sync(DECAES_FULL_MASK, anypos)) {
    if (lane == 0) s.w[n - 1] = 1.0;
  }
  if (TIKH) {
    for (int k = lane; k < zero_rows * ld; k += 32) s.A[m0 * ld + k] = 0.0;
    for (int k = lane; k < n; k += 32) s.b[m0 + k] = 0.0;
  }
  __syncwarp();
}

}  // namespace decaes
