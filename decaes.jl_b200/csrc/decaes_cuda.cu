// libdecaes_cuda: kernels + C ABI (include/decaes_cuda.h).  sm_100a only, no CPU fallback.
//
// Kernels
//   basis_setup_kernel   EPG decay basis set and its flip-angle Jacobian for the whole angle grid,
//                        once per call (thread_buffer_maker / EPGBasisSetEnsemble,
//                        src/T2mapSEcorr.jl:339-407, 596-614; EPGJacobianFunctor src/EPGdecaycurve.jl:224-248)
//   voxel_pipeline_kernel persistent one-warp CTAs; each warp pulls groups of 4 voxels and runs the
//                        whole per-voxel chain (voxel.cuh) with its NNLS system in shared memory
//   t2part_kernel        standalone T2part (src/T2partSEcorr.jl:95-138), one thread per voxel
//   mock_image_kernel    synthetic MSE volume (src/utils.jl:623-658)
//   dfma_peak_kernel     measured FP64 FMA peak for the roofline denominator
#include <algorithm>
#include <numeric>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/decaes_cuda.h"
#include "voxel.cuh"

using namespace decaes;

// ============================================================================ device code

// EPG value + d/dalpha (per degree) for one (angle, T2) pair; hand forward-mode through
// epg_impulse_response! (src/EPGdecaycurve.jl:948-1028) and |sind(alpha/2) * .| (:936-946).
#define EPG_MAXK 52  // phase states kept per curve: supports ETL <= DECAES_MAX_NTE (ETL - ETL / 2 + 1 = 49 at 96 echoes)
struct Dual {
  double v, d;
};
__device__ static void epg_curve_jac(int ETL, double alpha_deg, double E1, double E2, double *dc, double *ddc,
                                     int out_stride) {
  const double kk = 0.017453292519943295;
  double sina, cosa;
  sincos(alpha_deg * kk, &sina, &cosa);
  const double dsina = cosa * kk, dcosa = -sina * kk;
  const double E2h = (E2 * E2) / 2, E1E2 = E1 * E2, E1sq = E1 * E1;
  const double a = E2h, b = E2h * cosa, c = E1E2 * sina, d = E1sq * cosa, cp = -c / 2;
  const double db = E2h * dcosa, dcc = E1E2 * dsina, dd = E1sq * dcosa, dcp = -dcc / 2;
  const double m0 = sind_0_180(alpha_deg / 2);
  double shalf, chalf;
  {
    double lo, hi = deg2rad_dd(alpha_deg / 2, lo);
    sincos(hi, &shalf, &chalf);
    chalf = fma(-shalf, lo, chalf);
  }
  const double dm0 = chalf * kk / 2;
  Dual F[EPG_MAXK], Fb[EPG_MAXK], Z[EPG_MAXK];
  auto emit = [&](int i, double v, double dv) {
    double val = m0 * v, dval = dm0 * v + m0 * dv;
    dc[i * out_stride] = fabs(val);
    ddc[i * out_stride] = signbit(val) ? -dval : dval;
  };
  emit(0, a - b, -db);
  F[1] = {a - b, -db}, Fb[1] = {0, 0}, Z[1] = {cp, dcp};
  F[2] = {a + b, db}, Fb[2] = {0, 0}, Z[2] = {0, 0};
  double C, S, Cp, Sp, dC, dS, dCp, dSp;
  Dual vF, vFb, vZ;
#define LOADK(k)                                                          \
  C = F[k].v + Fb[k].v, S = F[k].v - Fb[k].v, dC = F[k].d + Fb[k].d, dS = F[k].d - Fb[k].d; \
  Cp = a * C, Sp = b * S, dCp = a * dC, dSp = db * S + b * dS;            \
  vFb.v = fma(-c, Z[k].v, Cp - Sp), vFb.d = (-dcc) * Z[k].v + (-c) * Z[k].d + (dCp - dSp); \
  vF.v = fma(c, Z[k].v, Cp + Sp), vF.d = dcc * Z[k].v + c * Z[k].d + (dCp + dSp); \
  vZ.v = fma(cp, S, d * Z[k].v), vZ.d = dcp * S + cp * dS + (dd * Z[k].v + d * Z[k].d)
  for (int i = 2; i <= ETL - 1; i++) {
    const bool first_half = (i <= ETL / 2);
    const int kmax = first_half ? i : ETL - i + 1;
    LOADK(1);
    emit(i - 1, vFb.v, vFb.d);
    F[1] = vFb, Z[1] = vZ;
    Dual pend = vF;
    for (int k = 2; k <= kmax; k++) {
      LOADK(k);
      F[k] = pend;
      pend = vF;
      Fb[k - 1] = vFb;
      Z[k] = vZ;
    }
    F[kmax + 1] = pend;
    if (first_half) Fb[i] = {0, 0}, Fb[i + 1] = {0, 0}, Z[i + 1] = {0, 0};
  }
  C = F[1].v + Fb[1].v, S = F[1].v - Fb[1].v, dC = F[1].d + Fb[1].d, dS = F[1].d - Fb[1].d;
  double v = fma(-c, Z[1].v, fma(a, C, (-b) * S));
  double dv = (-dcc) * Z[1].v + (-c) * Z[1].d + (a * dC - (db * S + b * dS));
  emit(ETL - 1, v, dv);
#undef LOADK
}

// General refocusing-control-angle variant (RefConAngle != 180): first refocusing pulse A*180 = alpha,
// later pulses A*beta with A = alpha/180 (src/EPGdecaycurve.jl:722-818).  Same in-place state update as
// above with the three dot products F.M, Fbar.M, Z.M per state; value + d/dalpha (per degree).
__device__ static void epg_curve_beta_jac(int ETL, double alpha_deg, double beta_deg, double E1, double E2, double *dc,
                                          double *ddc, int out_stride) {
  const double kk = 0.017453292519943295;
  const double A = alpha_deg / 180;
  const double a1r = (A * 180) * kk, air = (A * beta_deg) * kk;
  const double kb = (beta_deg / 180) * kk;  // d(air)/d(alpha_deg)
  double sh, ch, sini, cosi;
  sincos(a1r / 2, &sh, &ch);
  sincos(air, &sini, &cosi);
  const double dsh = ch * (kk / 2), dch = -sh * (kk / 2);
  const double s2h = sh * sh, c2h = ch * ch, ds2h = 2 * sh * dsh, dc2h = 2 * ch * dch;
  const double sin1 = 2 * sh * ch, dsin1 = 2 * (dsh * ch + sh * dch);
  const double dsini = cosi * kb, dcosi = -sini * kb;
  const double c2hi = (1 + cosi) / 2, s2hi = 1 - c2hi, dc2hi = dcosi / 2, ds2hi = -dc2hi;
  const double E2sq = E2 * E2, E1E2 = E1 * E2, E1sq = E1 * E1;
  const double a1 = E2sq * c2h, b1 = E2sq * s2h, c1 = E1E2 * sin1;
  const double da1 = E2sq * dc2h, db1 = E2sq * ds2h, dc1 = E1E2 * dsin1;
  const double ai = E2sq * c2hi, bi = E2sq * s2hi, ci = E1E2 * sini, di = E1sq * cosi;
  const double dai = E2sq * dc2hi, dbi = E2sq * ds2hi, dci = E1E2 * dsini, ddi = E1sq * dcosi;
  const double Fv[3] = {ai, bi, ci}, Fbv[3] = {bi, ai, -ci}, Zv[3] = {-ci / 2, ci / 2, di};
  const double dFv[3] = {dai, dbi, dci}, dFbv[3] = {dbi, dai, -dci}, dZv[3] = {-dci / 2, dci / 2, ddi};
  Dual F[EPG_MAXK], Fb[EPG_MAXK], Z[EPG_MAXK];
  auto dot = [](const double *u, const double *du, const Dual &x, const Dual &y, const Dual &z) -> Dual {
    Dual r;
    r.v = u[0] * x.v + u[1] * y.v + u[2] * z.v;
    r.d = (du[0] * x.v + du[1] * y.v + du[2] * z.v) + (u[0] * x.d + u[1] * y.d + u[2] * z.d);
    return r;
  };
  auto emit = [&](int i, const Dual &x) {
    dc[i * out_stride] = fabs(x.v);
    ddc[i * out_stride] = signbit(x.v) ? -x.d : x.d;
  };
  const double m0 = sh, dm0 = dsh;
  F[1] = {b1 * m0, db1 * m0 + b1 * dm0}, Fb[1] = {0, 0}, Z[1] = {-c1 * m0 / 2, -(dc1 * m0 + c1 * dm0) / 2};
  F[2] = {a1 * m0, da1 * m0 + a1 * dm0}, Fb[2] = {0, 0}, Z[2] = {0, 0};
  emit(0, F[1]);
  for (int i = 2; i <= ETL - 1; i++) {
    const bool first_half = (i <= ETL / 2);
    const int nproc = first_half ? i : ETL - i + 1;
    Dual FM = dot(Fv, dFv, F[1], Fb[1], Z[1]), FbM = dot(Fbv, dFbv, F[1], Fb[1], Z[1]), ZM = dot(Zv, dZv, F[1], Fb[1], Z[1]);
    emit(i - 1, FbM);
    F[1] = FbM, Z[1] = ZM;
    Dual pend = FM;
    for (int j = 2; j <= nproc; j++) {
      FM = dot(Fv, dFv, F[j], Fb[j], Z[j]), FbM = dot(Fbv, dFbv, F[j], Fb[j], Z[j]), ZM = dot(Zv, dZv, F[j], Fb[j], Z[j]);
      F[j] = pend;
      pend = FM;
      Fb[j - 1] = FbM;
      Z[j] = ZM;
    }
    if (first_half) F[nproc + 1] = pend, Fb[nproc] = {0, 0}, Fb[nproc + 1] = {0, 0}, Z[nproc + 1] = {0, 0};
  }
  emit(ETL - 1, dot(Fbv, dFbv, F[1], Fb[1], Z[1]));
}

struct SetupParams {
  int nTE, nT2, nA, ld, copy_elems;
  double E1, refcon;
  double angles[DECAES_MAX_ANGLES];
  double E2[DECAES_MAX_NT2];
  double *basis_rm, *basis_cm, *dbasis_cm;
};

__global__ void basis_setup_kernel(const __grid_constant__ SetupParams S) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S.nA * S.nT2) return;
  int k = t / S.nT2, j = t % S.nT2;
  double *dc = S.basis_cm + ((size_t)k * S.nT2 + j) * S.nTE;
  double *ddc = S.dbasis_cm + ((size_t)k * S.nT2 + j) * S.nTE;
  if (S.refcon == 180.0) epg_curve_jac(S.nTE, S.angles[k], S.E1, S.E2[j], dc, ddc, 1);  // dispatch of src/T2mapSEcorr.jl:616-620
  else epg_curve_beta_jac(S.nTE, S.angles[k], S.refcon, S.E1, S.E2[j], dc, ddc, 1);
  double *rm = S.basis_rm + (size_t)k * S.copy_elems;
  for (int i = 0; i < S.nTE; i++) rm[i * S.ld + j] = dc[i];
  if (j == 0) {  // zero the padding so the bulk copy never moves uninitialised bytes
    for (int i = 0; i < S.nTE; i++)
      for (int jj = S.nT2; jj < S.ld; jj++) rm[i * S.ld + jj] = 0.0;
    for (int e = S.nTE * S.ld; e < S.copy_elems; e++) rm[e] = 0.0;
  }
}

// G_k = A_k' A_k for every grid angle (Gram solver): one thread per (k, p, q)
__global__ void gram_setup_kernel(const __grid_constant__ SetupParams S, double *gram_set, int ldg, int gstride) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = S.nT2;
  if (t >= (long long)S.nA * n * n) return;
  int k = (int)(t / (n * n)), p = (int)((t / n) % n), q = (int)(t % n);
  const double *rm = S.basis_rm + (size_t)k * S.copy_elems;
  double a0 = 0.0, a1 = 0.0;
  int i = 0;
  for (; i + 1 < S.nTE; i += 2) {
    a0 = fma(rm[i * S.ld + p], rm[i * S.ld + q], a0);
    a1 = fma(rm[(i + 1) * S.ld + p], rm[(i + 1) * S.ld + q], a1);
  }
  if (i < S.nTE) a0 = fma(rm[i * S.ld + p], rm[i * S.ld + q], a0);
  gram_set[(size_t)k * gstride + p * ldg + q] = a0 + a1;
}

// Persistent CTAs of P.warps_per_cta warps, one CTA per SM.  Every warp owns a slice of the dynamic
// shared memory and processes its own voxels; CTA barriers keep the warps in the same PHASE of the
// per-voxel chain (flip-angle fit | EPG basis | regularised solve + outputs).  The kernel is
// instruction-fetch bound (ncu: stall_no_inst dominates when warps wander through the ~100 KB of hot
// code independently; every phase runs 10-20 % slower when phases overlap), and warps that execute
// the same phase share its cache lines.
template <bool GRAM, bool LEGACY = false, int VS = 64>
__global__ void __launch_bounds__(DECAES_MAX_WARPS * 32, 1) voxel_pipeline_kernel(const __grid_constant__ PipeParams P) {
  extern __shared__ __align__(128) double smem[];
  const int wid = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + wid;
  double *gscratch = P.scratch + (size_t)gwarp * P.scratch_per_warp;
  Warp<GRAM, LEGACY, VS> W(P, smem + (size_t)wid * (P.smem_per_warp / 8), gscratch);
  const int lane = lane_id();
  if (lane == 0) {
    mbar_init(W.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncwarp();

  // Phase barriers: CTA-wide (sync_groups = 1) or per group of warps that share an SM sub-partition
  // (warp id mod 4; sync_groups = 4), so that fewer warps wait for the slowest one of their group.
  const int nwarps = blockDim.x >> 5;
  const int grp = (P.sync_groups > 1) ? (wid % P.sync_groups) : 0;
  const int gthreads = 32 * ((nwarps - grp + (P.sync_groups > 1 ? P.sync_groups : 1) - 1) / (P.sync_groups > 1 ? P.sync_groups : 1));
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(gthreads) : "memory"); };
  auto group_or = [&](bool pred) -> bool {
    int r;
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.b32 p, %3, 0;\nbarrier.red.or.pred q, %1, %2, p;\nselp.b32 %0, 1, 0, q;\n}\n"
        : "=r"(r)
        : "r"(grp + 1), "r"(gthreads), "r"((int)pred)
        : "memory");
    return r != 0;
  };
  const long long ngroups = (P.nvox + DECAES_GROUP - 1) / DECAES_GROUP;
  unsigned long long processed = 0;
  long long cyc[4] = {0, 0, 0, 0};  // per-warp cycles: barrier wait, flip angle, basis, solve+save
  const long long t_begin = clock64();
  long long v0 = 0;
  int q = DECAES_GROUP;  // next voxel of the current group (DECAES_GROUP = group exhausted)
  bool more_groups = true;
  while (true) {
    // ---- find this warp's next voxel above threshold (NaN-fill the skipped ones on the way) ----
    bool have = false;
    long long v = 0;
    const double *signal = nullptr;
    while (!have) {
      if (q >= DECAES_GROUP) {
        if (!more_groups) break;
        unsigned long long gidx = 0;
        if (lane == 0) gidx = atomicAdd(&P.counters[0], 1ull);
        gidx = __shfl_sync(DECAES_FULL_MASK, gidx, 0);
        if ((long long)gidx >= ngroups) {
          more_groups = false;
          break;
        }
        v0 = (long long)gidx * DECAES_GROUP;  // 4 consecutive voxels share one 32-byte sector per echo (L2 serves the re-reads)
        q = 0;
      }
      v = v0 + q;
      if (v >= P.nvox) {
        q = DECAES_GROUP;
        continue;
      }
      signal = P.image + v;
      q++;
      if (__ldg(signal) > P.Threshold) {  // src/T2mapSEcorr.jl:177
        have = true;
      } else {
        // skipped voxel: outputs are NaN (the reference pre-fills NaN, src/T2mapSEcorr.jl:36-52)
        const double nanv = CUDART_NAN;
        if (lane == 0) {
          P.gdn[v] = nanv, P.ggm[v] = nanv, P.gva[v] = nanv, P.fnr[v] = nanv, P.snr[v] = nanv;
          if (!P.alpha_provided) P.alpha[v] = nanv;
          if (P.mu && P.chi2factor) P.mu[v] = nanv, P.chi2factor[v] = nanv;
          if (P.resnorm) P.resnorm[v] = nanv;
          if (P.has_part) P.sfr[v] = nanv, P.sgm[v] = nanv, P.mfr[v] = nanv, P.mgm[v] = nanv;
        }
        for (int j = lane; j < P.nT2; j += 32) P.dist[v + (long long)j * P.stride] = nanv;
        if (P.decaycurve)
          for (int i = lane; i < P.nTE; i += 32) P.decaycurve[v + (long long)i * P.stride] = nanv;
        if (P.decaybasis && !P.fixed_alpha)
          for (int k = lane; k < P.nTE * P.nT2; k += 32) P.decaybasis[v + (long long)k * P.stride] = nanv;
      }
    }
    // Barriers per voxel round (sync_mask): bit 2 = before the flip-angle fit (the strict three-phase lock step),
    // bit 0 = after it (mandatory: it doubles as the "every warp is out of work" vote), bit 1 = after the basis.
    // Without bit 2 a warp that finishes its regularised solve early starts the next voxel's flip-angle fit at
    // once: one barrier fewer to wait at, at the price of some overlap between the two code regions.
    long long t0 = clock64();
    if (P.sync_mask & 4) group_sync();
    long long t1 = clock64();
    if (have) W.phase_flip_angle(v, signal);
    else if (P.step_sync & 1) cta_drain();  // warps without a voxel take part in the votes
    long long t2 = clock64();
    if (!group_or(have)) break;  // every warp of the group is out of work
    long long t3 = clock64();
    if (have) W.phase_basis();
    long long t4 = clock64();
    if (P.sync_mask & 2) group_sync();
    long long t5 = clock64();
    if (have) {
      W.phase_solve_and_save();
      processed++;
    } else if (P.step_sync & 14) {
      cta_drain();
    }
    long long t6 = clock64();
    cyc[0] += (t1 - t0) + (t3 - t2) + (t5 - t4), cyc[1] += t2 - t1, cyc[2] += t4 - t3, cyc[3] += t6 - t5;
  }
  if (lane == 0) {
    for (int c = 0; c < 4; c++) atomicAdd(&P.counters[4 + c], (unsigned long long)cyc[c]);
    atomicAdd(&P.counters[24], (unsigned long long)(clock64() - t_begin));
#ifdef DECAES_PROFILE
    for (int c = 0; c < PF_COUNT; c++) atomicAdd(&P.counters[8 + c], (unsigned long long)W.prof_cyc[c]);
#endif
  }
  if (lane == 0) {
    if (processed) atomicAdd(&P.counters[1], processed);
    if (W.n_early) atomicAdd(&P.counters[2], W.n_early);
    if (W.n_overflow) atomicAdd(&P.counters[3], W.n_overflow);
    if (W.n_itercap) atomicAdd(&P.counters[26], W.n_itercap);
    if (W.n_polish) atomicAdd(&P.counters[27], W.n_polish);
  }
}

struct PartParams {
  int nT2, sp_lo, sp_hi, mp_lo, mp_hi, has_sigmoid;
  long long nvox, stride;
  const double *dist;
  double *sfr, *sgm, *mfr, *mgm;
  double logT2[DECAES_MAX_NT2], weights[DECAES_MAX_NT2];
};

// voxelwise_T2_parts!  src/T2partSEcorr.jl:95-138 — one thread per voxel, bins strided by `stride`
// so that a warp reads 32 consecutive voxels of each bin (coalesced).
__global__ void t2part_kernel(const __grid_constant__ PartParams Q) {
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= Q.nvox) return;
  double S = 0, Ssp = 0, Smp = 0, dsp = 0, dmp = 0, dw = 0;
  bool isn = false;
  for (int j = 0; j < Q.nT2; j++) {
    double dj = Q.dist[v + (long long)j * Q.stride];
    isn |= isnan(dj);
    S += dj;
    if (j >= Q.sp_lo && j <= Q.sp_hi) dsp += dj * Q.logT2[j], Ssp += dj;
    if (j >= Q.mp_lo && j <= Q.mp_hi) dmp += dj * Q.logT2[j], Smp += dj;
    if (Q.has_sigmoid) dw = fma(dj, Q.weights[j], dw);
  }
  if (isn) return;
  if (S > 0) {
    Q.sfr[v] = Q.has_sigmoid ? dw / S : Ssp / S;
    Q.mfr[v] = Smp / S;
  }
  if (Ssp > 0) Q.sgm[v] = exp(dsp / Ssp);
  if (Smp > 0) Q.mgm[v] = exp(dmp / Smp);
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double urand(uint64_t seed, uint64_t vox, uint64_t k) {
  uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL * (vox + 1));
  h = mix64(h + 0x9E3779B97F4A7C15ULL * (k + 1));
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// mock_image  src/utils.jl:623-658 with a per-voxel flip angle ~ U(120,180)
__global__ void mock_image_kernel(double *image, long long nvox, long long stride, long long first_voxel, int nTE,
                                  double TE, double T1, double sigma, uint64_t seed) {
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  uint64_t gid = (uint64_t)(first_voxel + v);
  double sfr = 0.05 + (0.25 - 0.05) * urand(seed, gid, 0);
  double T21 = 10e-3 + (20e-3 - 10e-3) * urand(seed, gid, 1);
  double T22 = 50e-3 + (100e-3 - 50e-3) * urand(seed, gid, 2);
  double alpha = 120.0 + 60.0 * urand(seed, gid, 3);
  double E1 = exp(-(TE / 2) / T1);
  double d1[DECAES_MAX_NTE], d2[DECAES_MAX_NTE], dd[DECAES_MAX_NTE];
  epg_curve_jac(nTE, alpha, E1, exp(-(TE / 2) / T21), d1, dd, 1);
  epg_curve_jac(nTE, alpha, E1, exp(-(TE / 2) / T22), d2, dd, 1);
  for (int k = 0; k < nTE; k++) {
    double m = sfr * d1[k] + (1 - sfr) * d2[k];
    double u1 = urand(seed, gid, 4 + 2 * (uint64_t)k), u2 = urand(seed, gid, 5 + 2 * (uint64_t)k);
    double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    double zR = sigma * r * cs, zI = sigma * r * sn;
    image[v + (long long)k * stride] = sqrt((m + zR) * (m + zR) + zI * zI);
  }
}

// independent DFMA chains, 8 per thread
__global__ void dfma_peak_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
    a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ============================================================================ host code

static thread_local char g_err[512] = "";
static thread_local decaes_run_stats g_stats;

static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(DECAES_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ---- option checks: the assertions of T2mapOptions / T2partOptions (src/types.jl:28-84, 148-168) ----
static int validate_map(const decaes_t2map_opts *o) {
  if (!o) return fail(DECAES_EINVAL, "opts is NULL");
  if (!(o->nx >= 1 && o->ny >= 1 && o->nz >= 1)) return fail(DECAES_EINVAL, "MatrixSize must be a tuple of 3 positive integers");
  if (!(o->nTE >= 4)) return fail(DECAES_EINVAL, "At least four echoes are required for T2 mapping, but nTE = %d.", o->nTE);
  if (!(o->TE > 0.0)) return fail(DECAES_EINVAL, "Echo spacing must be positive, but TE = %g.", o->TE);
  if (!(o->nT2 >= 2)) return fail(DECAES_EINVAL, "At least two T2 components are required for T2 mapping, but nT2 = %d.", o->nT2);
  if (!(0.0 < o->T2min && o->T2min < o->T2max)) return fail(DECAES_EINVAL, "T2Range must a sorted 2-tuple of positive values");
  if (!(o->T1 > 0.0)) return fail(DECAES_EINVAL, "T1 must be positive, but T1 = %g.", o->T1);
  if (!(o->Threshold >= 0.0 || o->Threshold == -INFINITY))
    return fail(DECAES_EINVAL, "First echo signal threshold must be non-negative or -Inf");
  if (!(0.0 <= o->MinRefAngle && o->MinRefAngle <= 180.0)) return fail(DECAES_EINVAL, "Minimum refocusing angle must be in the range [0, 180]");
  if (!(o->nRefAngles >= 2)) return fail(DECAES_EINVAL, "nRefAngles must be at least 2, but nRefAngles = %d.", o->nRefAngles);
  if (!(2 <= o->nRefAnglesMin && o->nRefAnglesMin <= o->nRefAngles))
    return fail(DECAES_EINVAL, "nRefAnglesMin must be in the range [2, nRefAngles]");
  if (!(o->reg >= DECAES_REG_NONE && o->reg <= DECAES_REG_MDP)) return fail(DECAES_EINVAL, "Unrecognized regularization method: %d", o->reg);
  if (o->reg == DECAES_REG_CHI2 && !(o->Chi2Factor > 1.0)) return fail(DECAES_EINVAL, "Chi2Factor must be greater than 1.0");
  if (o->reg == DECAES_REG_MDP && !(o->NoiseLevel > 0.0)) return fail(DECAES_EINVAL, "Noise level must be positive");
  if (!(0.0 <= o->RefConAngle && o->RefConAngle <= 180.0)) return fail(DECAES_EINVAL, "Refocusing control angle must be in the range [0, 180]");
  if (!std::isnan(o->SetFlipAngle) && !(0.0 <= o->SetFlipAngle && o->SetFlipAngle <= 180.0))
    return fail(DECAES_EINVAL, "Fixed flip angle must be in the range [0, 180]");
  if (o->nT2 > DECAES_MAX_NT2) return fail(DECAES_EUNSUPPORTED, "nT2 > %d is not supported", DECAES_MAX_NT2);
  if (o->nRefAngles > DECAES_MAX_ANGLES) return fail(DECAES_EUNSUPPORTED, "nRefAngles > %d is not supported", DECAES_MAX_ANGLES);
  if (o->nTE > DECAES_MAX_NTE) return fail(DECAES_EUNSUPPORTED, "nTE > %d is not supported", DECAES_MAX_NTE);
  return DECAES_OK;
}
static int validate_part(const decaes_t2part_opts *o) {
  if (!o) return fail(DECAES_EINVAL, "part opts is NULL");
  if (!(o->nx >= 1 && o->ny >= 1 && o->nz >= 1)) return fail(DECAES_EINVAL, "MatrixSize must be positive");
  if (!(o->nT2 >= 2)) return fail(DECAES_EINVAL, "nT2 must be at least 2");
  if (!(0.0 < o->T2min && o->T2min < o->T2max)) return fail(DECAES_EINVAL, "T2Range must be sorted and positive");
  if (!(o->SPWin_lo < o->SPWin_hi)) return fail(DECAES_EINVAL, "SPWin must be sorted");
  if (!(o->MPWin_lo < o->MPWin_hi)) return fail(DECAES_EINVAL, "MPWin must be sorted");
  if (!std::isnan(o->Sigmoid) && !(o->Sigmoid > 0)) return fail(DECAES_EINVAL, "Sigmoid must be positive");
  if (o->nT2 > DECAES_MAX_NT2) return fail(DECAES_EUNSUPPORTED, "nT2 > %d is not supported", DECAES_MAX_NT2);
  return DECAES_OK;
}

// ---- host tables ----
// range(a, b; length = n) is evaluated in twice-precision by Julia; long double stands in.
static void linrange(double a, double b, int n, double *out) {
  if (n == 1) {
    out[0] = a;
    return;
  }
  for (int i = 0; i < n; i++)
    out[i] = (double)(((long double)a * (long double)(n - 1 - i) + (long double)b * (long double)i) / (long double)(n - 1));
  out[0] = a, out[n - 1] = b;
}
// start:step:stop for Float64 as Julia's Base builds it (base/twiceprecision.jl: `rat`, the lift to rationals,
// `floatrange`): the legacy searches scan knots[1]:0.001:knots[end] (src/splines.jl:422, 438) and every sample must be
// the double Julia would produce.  When start, step and stop are small rationals the elements are
// (start_n + i*step_n)/den rounded once; otherwise start + i*step taken literally (also rounded once).
static void julia_rat(double x, long long *num, long long *den) {
  double y = x;
  long long a = 1, d = 1, b = 0, c = 0;
  const double m = 16777216.0;  // maxintfloat(Float32): Base narrows Float64 before the continued fraction
  while (std::fabs(y) <= m) {
    const long long f = (long long)std::trunc(y);
    y -= (double)f;
    const long long a2 = f * a + c, b2 = f * b + d;
    c = a, a = a2, d = b, b = b2;
    if (!(std::max(std::llabs(a), std::llabs(b)) <= (long long)m)) {
      *num = c, *den = d;
      return;
    }
    if ((double)a / (double)b == x) break;
    y = 1.0 / y;
  }
  *num = a, *den = b;
}
static LegacyRange julia_colon(double start, double step, double stop) {
  LegacyRange r;
  memset(&r, 0, sizeof r);
  r.start = start, r.step = step;
  auto between = [](double lo, double x, double hi) { return (lo <= x && x <= hi) || (hi <= x && x <= lo); };
  long long step_n, step_d, start_n, start_d, stop_n, stop_d;
  julia_rat(step, &step_n, &step_d);
  if (step_d != 0 && (double)step_n / (double)step_d == step) {
    julia_rat(start, &start_n, &start_d);
    julia_rat(stop, &stop_n, &stop_d);
    if (start_d != 0 && stop_d != 0 && (double)start_n / (double)start_d == start && (double)stop_n / (double)stop_d == stop) {
      const long long den = start_d / std::gcd(start_d, step_d) * step_d;
      const double mi = 9007199254740992.0;
      if (den != 0 && std::fabs(start * (double)den) <= mi && std::fabs(step * (double)den) <= mi && den % start_d == 0 &&
          den % step_d == 0) {
        start_n = std::llround(start * (double)den), step_n = std::llround(step * (double)den);
        long long len = std::max(0ll, (den * stop_n - stop_d * start_n + step_n * stop_d) / (step_n * stop_d));
        if (between(start, start + (double)(len - 1) * step, stop + step / 2) && !between(start, start + (double)len * step, stop)) {
          r.rational = 1, r.start_n = start_n, r.step_n = step_n, r.den = den, r.len = len;
          return r;
        }
      }
    }
  }
  const double lf = (stop - start) / step;
  long long len = lf < 0 ? 0 : (lf == 0 ? 1 : std::llrint(lf) + 1);
  if (lf > 0) {
    const double stop2 = start + (double)(len - 1) * step;
    len -= (start < stop && stop < stop2) + (start > stop && stop > stop2);
  }
  r.len = len;
  return r;
}
static void logrange(double a, double b, int n, double *out) {  // src/utils.jl:7
  linrange(std::log(a), std::log(b), n, out);
  for (int i = 0; i < n; i++) out[i] = std::exp(out[i]);
  out[0] = a, out[n - 1] = b;
}

// probe order of initialize! (src/splines.jl:716-734): depends only on the grid size and budgets
struct SeedState {
  std::vector<char> seen;
  std::vector<int> order;
  int numeval = 0;
};
static void seed_eval_box(SeedState &s, int lo, int hi, int maxeval) {
  int cs[2] = {lo, hi};
  for (int k = 0; k < 2; k++) {
    if (s.seen[lo] && s.seen[hi]) break;
    if (s.numeval >= maxeval) break;
    if (s.seen[cs[k]]) continue;
    s.seen[cs[k]] = 1, s.order.push_back(cs[k]), s.numeval++;
    if (s.numeval >= maxeval) break;
  }
}
static void seed_rec(SeedState &s, int lo, int hi, int depth, int mineval, int maxeval) {
  if (depth <= 0) return;
  seed_eval_box(s, lo, hi, maxeval);
  if (s.numeval >= mineval) return;
  int mid = (lo + hi) / 2;  // 0-based midpoint of 1-based (lo+hi)/2: ((lo+1)+(hi+1))/2 - 1
  mid = ((lo + 1) + (hi + 1)) / 2 - 1;
  seed_rec(s, lo, mid, depth - 1, mineval, maxeval);
  seed_rec(s, mid, hi, depth - 1, mineval, maxeval);
}
static std::vector<int> seed_order(int n, int mineval, int maxeval) {
  SeedState s;
  s.seen.assign(n, 0);
  for (int depth = 1; depth <= mineval; depth++) {
    seed_rec(s, 0, n - 1, depth, mineval, maxeval);
    if (s.numeval >= mineval) break;
  }
  return s.order;
}

static double erfinv_newton(double y) {
  double x = 0.0;
  for (int it = 0; it < 100; it++) {
    double dx = (std::erf(x) - y) / (1.1283791670955126 * std::exp(-x * x));
    x -= dx;
    if (std::fabs(dx) < 1e-16 * std::max(1.0, std::fabs(x))) break;
  }
  return x;
}

struct PartTables {
  int sp_lo, sp_hi, mp_lo, mp_hi, has_sigmoid;  // 0-based inclusive
  double logT2[DECAES_MAX_NT2], weights[DECAES_MAX_NT2];
};
// thread_buffer_maker(::T2partOptions)  src/T2partSEcorr.jl:143-164
static int make_part_tables(const decaes_t2part_opts *o, PartTables *t) {
  const int n = o->nT2;
  double T2[DECAES_MAX_NT2];
  logrange(o->T2min, o->T2max, n, T2);
  for (int j = 0; j < n; j++) t->logT2[j] = std::log(T2[j]), t->weights[j] = 0.0;
  int f;
  for (f = 0; f < n && !(T2[f] >= o->SPWin_lo); f++) {}
  int sp_lo = f < n ? f : -1;
  for (f = n - 1; f >= 0 && !(T2[f] <= o->SPWin_hi); f--) {}
  int sp_hi = f;
  for (f = 0; f < n && !(T2[f] >= o->MPWin_lo); f++) {}
  int mp_lo = f < n ? f : -1;
  for (f = n - 1; f >= 0 && !(T2[f] <= o->MPWin_hi); f--) {}
  int mp_hi = f;
  if (sp_lo < 0 || sp_hi < 0 || mp_lo < 0 || mp_hi < 0)
    return fail(DECAES_EINVAL, "SPWin/MPWin do not intersect the T2 grid (findfirst/findlast returned nothing)");
  t->sp_lo = sp_lo, t->sp_hi = sp_hi, t->mp_lo = mp_lo, t->mp_hi = mp_hi;
  t->has_sigmoid = !std::isnan(o->Sigmoid);
  if (t->has_sigmoid) {  // sigmoid_weights  :155-164
    double sigma = std::fabs(o->Sigmoid / (std::sqrt(2.0) * erfinv_newton(2 * 0.1 - 1)));
    for (int j = 0; j < n; j++) {
      double w = std::erfc(((T2[j] - o->SPWin_hi) / sigma) / std::sqrt(2.0)) / 2;
      t->weights[j] = w <= 2.220446049250313e-16 ? 0.0 : w;
    }
  }
  return DECAES_OK;
}

// the instantiation that serves a parameter set: solver variant x legacy searches x vector stride of the Gram solver
typedef void (*pipeline_kernel_t)(const PipeParams);
static pipeline_kernel_t pipeline_kernel_for(const PipeParams &P) {
  if (!P.gram) return voxel_pipeline_kernel<false, false, 64>;
  if (P.legacy) return P.gv_stride == 40 ? voxel_pipeline_kernel<true, true, 40> : voxel_pipeline_kernel<true, true, 64>;
  return P.gv_stride == 40 ? voxel_pipeline_kernel<true, false, 40> : voxel_pipeline_kernel<true, false, 64>;
}

// ---- per-device workspace (grow-only, cached across calls) ----
struct DeviceWs {
  double *basis_rm = nullptr, *basis_cm = nullptr, *dbasis_cm = nullptr, *scratch = nullptr, *gram_set = nullptr;
  double *slab = nullptr;  // host API: device copy of this device's voxel slab (image + outputs), kept between calls
  unsigned long long *counters = nullptr;
  size_t basis_rm_cap = 0, basis_cm_cap = 0, scratch_cap = 0, gram_cap = 0, slab_cap = 0;
  size_t l2_persist_max = 0, l2_window_max = 0;  // persisting-L2 carve-out set aside for the scratch, largest access-policy window
  bool props_known = false;
  cudaDeviceProp prop;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool ev_pending = false;
  bool ev2_recorded = false;  // ev[2] marks the end of the last pipeline kernel enqueued on this device (any stream)
  int chunks_pending = 0;  // pipeline launches since the last collect_device_stats (one counter block each)
  int sm_count = 0;
  void *ring[4] = {nullptr, nullptr, nullptr, nullptr};  // pinned staging ring (host API, pageable caller buffers)
  cudaEvent_t ring_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t ring_cap = 0;  // bytes per slot
};
#define DECAES_NSTAGE 4
#define DECAES_MAX_CHUNKS 8  // sub-slabs of one device's slab whose copies overlap the neighbours' kernels
static std::mutex g_ws_mutex;
static DeviceWs g_ws[64];
// One set of tables / scratch / counters / kernel parameters (the __constant__ cP) exists per device, so everything
// that plans or launches on a device holds that device's mutex, and a new launch waits for the previous pipeline
// kernel of the device on the GPU side too (launch_pipeline: cudaStreamWaitEvent on ev[2]).
static std::recursive_mutex g_dev_mutex[64];

static int ensure(double **p, size_t *cap, size_t need_doubles) {
  if (*cap >= need_doubles) return DECAES_OK;
  if (*p) CUDA_TRY(cudaFree(*p));
  *p = nullptr, *cap = 0;
  CUDA_TRY(cudaMalloc(p, need_doubles * sizeof(double)));
  *cap = need_doubles;
  return DECAES_OK;
}

struct Plan {
  PipeParams P;
  SetupParams S;
  int smem_bytes, grid, warps_per_cta, cta_smem;
};

// Build kernel parameters (everything except volume pointers) and make sure the device workspace fits.
static int make_plan(const decaes_t2map_opts *o, const decaes_t2part_opts *part, int dev, Plan *plan) {
  PipeParams &P = plan->P;
  SetupParams &S = plan->S;
  memset(&P, 0, sizeof P);
  memset(&S, 0, sizeof S);
  const int nTE = o->nTE, nT2 = o->nT2;
  const bool fixed = !std::isnan(o->SetFlipAngle);
  const int nA = fixed ? 1 : o->nRefAngles;
  P.nTE = nTE, P.nT2 = nT2, P.nA = nA, P.reg = o->reg;
  P.ld = nT2 | 1;
  P.copy_elems = (nTE * P.ld + 1) & ~1;
  P.rows_alloc = (o->reg == DECAES_REG_NONE) ? nTE : nTE + nT2;
  P.epg_kmax = nTE - nTE / 2 + 1;  // phase states 1..K touched by the truncated recursion (stored at 0..K-1)
  // solver variant: normal-equation active set (default) or the QR port (DECAES_SOLVER=qr, kept for A/B checks)
  const char *sv = getenv("DECAES_SOLVER");
  P.gram = !(sv && strcmp(sv, "qr") == 0);
  // Components per pass of the shared-memory EPG: ceil(nT2 / 32) passes; among the splits with that many passes, the
  // one whose accesses need the fewest 128-byte shared-memory wavefronts (16 doubles each) - the phase is bound by the
  // shared-memory pipe.  nT2 = 40: 24 + 16 lanes (2 + 1 wavefronts per access) instead of 20 + 20 (2 + 2).
  const int epg_npass = (nT2 + 31) / 32, epg_lanes_even = (nT2 + epg_npass - 1) / epg_npass;
  {
    int best = epg_lanes_even, best_w = 1 << 30;
    for (int lw = epg_lanes_even; lw <= 32; lw++) {
      int w = 0;
      for (int j0 = 0; j0 < nT2; j0 += lw) w += (std::min(lw, nT2 - j0) + 15) / 16;
      if (w < best_w) best_w = w, best = lw;
    }
    P.epg_lanes = P.gram ? best : 32;  // the QR port keeps 32 lanes per pass
    if (const char *e = getenv("DECAES_EPG_LANES")) P.epg_lanes = std::max(epg_lanes_even, std::min(32, atoi(e)));
  }
  int epg_elems = 3 * P.epg_kmax * 32;
  P.a_elems = std::max(std::max(P.rows_alloc * P.ld, P.copy_elems), epg_elems);
  // Brent-based choosers (gcv / chi2 / mdp) are numerically stable searches: keep their inputs at
  // reference-level accuracy.  The L-curve search flips on 1-ulp noise anyway (tests/test_oracle_sensitivity.py).
  P.refine_tikh = (o->reg != DECAES_REG_LCURVE);
  if (const char *e = getenv("DECAES_REFINE")) P.refine_tikh = atoi(e);
  P.sync_mask = 7;  // strict lock step of the three phases (measured best: 2.54 M warp-cycles per voxel against 2.68 M without the round barrier)
  P.sync_groups = 1;
  if (const char *e = getenv("DECAES_SYNC_GROUPS")) P.sync_groups = std::max(1, atoi(e));
  P.fa_warm = 4;
  if (const char *e = getenv("DECAES_FA_WARM")) P.fa_warm = atoi(e);
  P.fa_polish = 1, P.fa_refine = 1;
  if (const char *e = getenv("DECAES_FA_POLISH")) P.fa_polish = atoi(e);
  // seed probes without refinement / polish: +1 % (cfg3) to +4 % (cfg4, cfg5), but at SNR 15 the cruder loss slopes steer 3 of
  // 2,048 voxels into another bracket of a flat loss (flip angle off by 0.07 degrees): off (profiles/r02_s2_ab_rough_seed_probes.txt)
  P.fa_rough_seeds = 0;
  if (const char *e = getenv("DECAES_FA_ROUGH_SEEDS")) P.fa_rough_seeds = atoi(e) != 0;
  P.kkt_tau = 1e-8;  // two orders of magnitude above the noise of the normal-equation duals (1e-6: three candidates per solve instead of one)
  if (const char *e = getenv("DECAES_KKT_TAU")) P.kkt_tau = atof(e);
  if (const char *e = getenv("DECAES_FA_REFINE")) P.fa_refine = atoi(e);
  P.warm_ones = 1;
  if (const char *e = getenv("DECAES_WARM_ONES")) P.warm_ones = atoi(e) != 0;
  P.lc_hints = 15;  // bit 2: the full-set start is a direct Cholesky solve (gram_dense_solve); bit 3: dilated sets for points 2 and 3
  if (const char *e = getenv("DECAES_LC_HINTS")) P.lc_hints = atoi(e) & 15;  // bit 3: dilated sets for the second / third point
  // In-phase votes (voxel.cuh: cta_or): flip-angle probes and L-curve steps are uniform enough that keeping the warps
  // in step pays (cfg3: +20 %, cfg1: +13 %); the four initial L-curve points and the Brent searches vary too much between
  // voxels (-1 % / -6 %).  With few warps per SM (nT2 = 60: six) there is little instruction-cache pressure to relieve
  // and the votes inside the searches only cost: decided below, once the CTA shape is known.
  P.step_sync = 7;  // (bit 2, the four initial L-curve points: -1 % when first tried, +0.5 % on cfg3 / +1.0 % on cfg2 with the final kernel,
                    //  profiles/r02_s4_ab_votes_initial_points.txt)
  if (const char *e = getenv("DECAES_SYNC_MASK")) P.sync_mask = atoi(e);
  P.ldg = (nT2 + 1) | 1;  // one array holds the lower triangle of G and, above it, M = L^-1 (gram.cuh)
  if (P.gram) P.a_elems = nT2 * P.ldg;
  P.a_elems = (P.a_elems + 1) & ~1;
  P.fixed_alpha = fixed, P.alpha_provided = o->alpha_provided;
  P.maxeval = o->nRefAngles;
  P.TE = o->TE, P.T1 = o->T1, P.Threshold = o->Threshold, P.Chi2Factor = o->Chi2Factor;
  P.NoiseLevel = o->NoiseLevel, P.SetFlipAngle = o->SetFlipAngle;
  P.E1 = std::exp(-(o->TE / 2) / o->T1);
  double T2[DECAES_MAX_NT2];
  logrange(o->T2min, o->T2max, nT2, T2);
  for (int j = 0; j < nT2; j++) P.logT2[j] = std::log(T2[j]), P.E2[j] = std::exp(-(o->TE / 2) / T2[j]);
  if (fixed)
    P.angles[0] = o->SetFlipAngle;
  else
    linrange(o->MinRefAngle, 180.0, nA, P.angles);
  if (o->legacy) {
    // legacy = true: CubicSplineSurrogate with the sampled spline minimum for the flip angle, doubling search +
    // sampled spline root for Reg = chi2 (csrc/legacy.cuh); everything else is shared with the modern path
    if (!P.gram) return fail(DECAES_EUNSUPPORTED, "legacy = true needs the default solver (unset DECAES_SOLVER)");
    P.legacy = 1;
    if (!fixed) {
      P.lg_range = julia_colon(P.angles[0], 0.001, P.angles[nA - 1]);
      if (P.lg_range.len < 1 || P.lg_range.len > (1ll << 24))
        return fail(DECAES_EUNSUPPORTED, "legacy flip-angle scan of %lld samples is not supported", P.lg_range.len);
    }
  }
  if (!fixed && !o->alpha_provided) {
    std::vector<int> seeds = seed_order(nA, o->nRefAnglesMin, o->nRefAngles);
    P.nseed = (int)seeds.size();
    for (int i = 0; i < P.nseed; i++) P.seeds[i] = (int8_t)seeds[i];
  }
  if (part) {
    PartTables pt;
    int rc = make_part_tables(part, &pt);
    if (rc) return rc;
    if (part->nT2 != nT2) return fail(DECAES_EINVAL, "T2part nT2 (%d) != T2map nT2 (%d)", part->nT2, nT2);
    // the fused epilogue shares the T2 grid of the map (log T2 table, window indices): a different T2Range would
    // silently disagree with the standalone T2partSEcorr on the same distribution
    if (part->T2min != o->T2min || part->T2max != o->T2max)
      return fail(DECAES_EINVAL, "fused T2part needs the T2Range of the T2map options ((%g, %g) != (%g, %g)); call decaes_t2part on the distribution instead",
                  part->T2min, part->T2max, o->T2min, o->T2max);
    P.has_part = 1, P.sp_lo = pt.sp_lo, P.sp_hi = pt.sp_hi, P.mp_lo = pt.mp_lo, P.mp_hi = pt.mp_hi;
    P.has_sigmoid = pt.has_sigmoid;
    memcpy(P.weights, pt.weights, sizeof pt.weights);
  }

  // Shared memory per warp decides how many voxels an SM holds.  Moving the two coldest per-voxel tables (cached
  // solutions, L-curve state records) to the warp's global scratch costs ~1 % each and is done when it buys a warp.
  // stride of the solver's vectors: 40 covers the reference's default nT2 = 40 and buys a 12th warp on configs 1-3
  P.gv_stride = (P.gram && nT2 <= 40 && !getenv("DECAES_GV_STRIDE_64")) ? 40 : 64;
  P.spill = 0;
  if (P.gram) {
    int optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int best_w = 0;
    bool best_epg = false;
    for (int sp : {0, 1, 3}) {
      SmemLayout Ls(nTE, nT2, P.rows_alloc, P.a_elems, P.gram, sp, P.gv_stride);
      int w = (int)std::min<size_t>(DECAES_MAX_WARPS, ((size_t)optin - 1024) / Ls.total_bytes);
      const bool epg_ok = 3 * P.epg_kmax * P.epg_lanes <= Ls.bd;  // the shared-memory EPG must still fit in front of the signal
      if ((epg_ok && !best_epg && w >= 1) || (epg_ok == best_epg && w > best_w)) best_w = w, best_epg = epg_ok, P.spill = sp;
    }
    if (const char *e = getenv("DECAES_SPILL")) P.spill = atoi(e) & 3;
  }
  SmemLayout L(nTE, nT2, P.rows_alloc, P.a_elems, P.gram, P.spill, P.gv_stride);
  plan->smem_bytes = L.total_bytes;
  // the solver block + search caches (everything in front of the voxel's signal) are idle while the basis is built
  if (P.gram && 3 * P.epg_kmax * P.epg_lanes > L.bd && 3 * P.epg_kmax * epg_lanes_even <= L.bd) P.epg_lanes = epg_lanes_even;
  P.epg_smem = P.gram && 3 * P.epg_kmax * P.epg_lanes <= L.bd;
  if (const char *e = getenv("DECAES_EPG_SMEM")) P.epg_smem = P.epg_smem && atoi(e);
  P.gcv_smem = P.gram && o->reg == DECAES_REG_GCV && nTE * nT2 <= L.bd && std::min(nTE, nT2) <= 64;
  // 2 = bidiagonalisation + multisection (needs the tall copy with an odd leading dimension), 1 = parallel Jacobi (A/B: DECAES_GCV_SMEM=1)
  if (P.gcv_smem && std::max(nTE, nT2) * (std::min(nTE, nT2) | 1) <= L.bd) P.gcv_smem = 2;
  if (const char *e = getenv("DECAES_GCV_SMEM")) P.gcv_smem = std::min(P.gcv_smem, atoi(e));
  // one copy of the voxel's basis in the global scratch (column-major) unless someone needs the row-major one too: the QR
  // port (TMA source), the shuffle EPG (no on-the-fly right-hand side), the global-memory SVD of Reg = gcv
  P.need_rm = !P.gram || !P.epg_smem || (o->reg == DECAES_REG_GCV && !P.gcv_smem);
  if (const char *e = getenv("DECAES_NEED_RM")) P.need_rm = P.need_rm || atoi(e);
  P.epg_fuse = 1;
  if (const char *e = getenv("DECAES_EPG_FUSE")) P.epg_fuse = atoi(e) != 0;
  P.refcon = o->RefConAngle;
  if (P.refcon != 180.0 && !fixed && !(P.gram && 3 * P.epg_kmax * P.epg_lanes <= L.bd))
    return fail(DECAES_EUNSUPPORTED, "RefConAngle != 180 needs the Gram solver and %d bytes of shared EPG scratch per warp", 3 * P.epg_kmax * P.epg_lanes * 8);
  if (P.refcon != 180.0) P.epg_smem = 1;
  {
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    DeviceWs &w0 = g_ws[dev];
    if (!w0.props_known) {  // cudaGetDeviceProperties is slow (milliseconds): once per device
      CUDA_TRY(cudaGetDeviceProperties(&w0.prop, dev));
      w0.props_known = true;
      if (w0.prop.persistingL2CacheMaxSize > 0 &&
          cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)w0.prop.persistingL2CacheMaxSize) == cudaSuccess) {
        w0.l2_persist_max = (size_t)w0.prop.persistingL2CacheMaxSize;
        w0.l2_window_max = (size_t)w0.prop.accessPolicyMaxWindowSize;
      } else {
        cudaGetLastError();
      }
    }
  }
  const cudaDeviceProp &prop = g_ws[dev].prop;
  if ((size_t)plan->smem_bytes > prop.sharedMemPerBlockOptin)
    return fail(DECAES_EUNSUPPORTED, "problem needs %d bytes of shared memory per warp (> %zu)", plan->smem_bytes,
                prop.sharedMemPerBlockOptin);
  // one persistent CTA per SM; as many warps as shared memory allows (registers: launch bounds)
  int wpc = (int)std::min<size_t>(DECAES_MAX_WARPS, (prop.sharedMemPerBlockOptin - 1024) / plan->smem_bytes);
  if (const char *e = getenv("DECAES_WARPS_PER_CTA")) wpc = std::max(1, std::min(wpc, atoi(e)));
  if (wpc < 1) return fail(DECAES_EUNSUPPORTED, "not enough shared memory for one warp");
  P.warps_per_cta = wpc, P.smem_per_warp = plan->smem_bytes;
  // few warps per SM (nT2 = 60: six): the vote before every flip-angle probe still pays with today's code (cfg4-chi2 +2.3 %,
  // cfg5 +0.5 %; Reg = gcv -1.4 %: off there), the votes inside the regularised searches do not
  // (profiles/r02_s4_ab_votes_six_warps.txt)
  if (wpc < 9) P.step_sync &= (o->reg == DECAES_REG_GCV) ? 0 : 1;
  if (const char *e = getenv("DECAES_STEP_SYNC")) P.step_sync = atoi(e) & 15;
  if (fixed || o->alpha_provided) P.step_sync &= ~1;
  if (o->reg != DECAES_REG_LCURVE) P.step_sync &= ~6;
  if (o->reg == DECAES_REG_LCURVE || o->reg == DECAES_REG_NONE) P.step_sync &= ~8;
  plan->warps_per_cta = wpc;
  plan->cta_smem = wpc * plan->smem_bytes;
  CUDA_TRY(cudaFuncSetAttribute(pipeline_kernel_for(P), cudaFuncAttributeMaxDynamicSharedMemorySize, plan->cta_smem));
  plan->grid = prop.multiProcessorCount;

  ScratchLayout sl(nTE, nT2, P.copy_elems, o->reg == DECAES_REG_GCV, P.need_rm != 0);
  P.scratch_per_warp = sl.total;

  std::lock_guard<std::mutex> lk(g_ws_mutex);
  DeviceWs &ws = g_ws[dev];
  ws.sm_count = prop.multiProcessorCount;
  int rc;
  if ((rc = ensure(&ws.basis_rm, &ws.basis_rm_cap, (size_t)nA * P.copy_elems))) return rc;
  size_t cm = (size_t)nA * nTE * nT2;
  if (ws.basis_cm_cap < cm) {
    if (ws.basis_cm) cudaFree(ws.basis_cm), cudaFree(ws.dbasis_cm);
    ws.basis_cm = ws.dbasis_cm = nullptr, ws.basis_cm_cap = 0;
    CUDA_TRY(cudaMalloc(&ws.basis_cm, cm * sizeof(double)));
    CUDA_TRY(cudaMalloc(&ws.dbasis_cm, cm * sizeof(double)));
    ws.basis_cm_cap = cm;
  }
  if ((rc = ensure(&ws.scratch, &ws.scratch_cap, (size_t)plan->grid * wpc * sl.total))) return rc;
  if ((rc = ensure(&ws.gram_set, &ws.gram_cap, (size_t)nA * P.a_elems))) return rc;
  P.gram_set = ws.gram_set;
  if (!ws.counters) CUDA_TRY(cudaMalloc(&ws.counters, DECAES_MAX_CHUNKS * 32 * sizeof(unsigned long long)));
  for (int i = 0; i < 4; i++)
    if (!ws.ev[i]) CUDA_TRY(cudaEventCreate(&ws.ev[i]));
  P.basis_rm = ws.basis_rm, P.basis_cm = ws.basis_cm, P.dbasis_cm = ws.dbasis_cm;
  P.scratch = ws.scratch, P.counters = ws.counters;

  S.nTE = nTE, S.nT2 = nT2, S.nA = nA, S.ld = P.ld, S.copy_elems = P.copy_elems, S.E1 = P.E1, S.refcon = o->RefConAngle;
  memcpy(S.angles, P.angles, sizeof S.angles);
  memcpy(S.E2, P.E2, sizeof S.E2);
  S.basis_rm = ws.basis_rm, S.basis_cm = ws.basis_cm, S.dbasis_cm = ws.dbasis_cm;
  return DECAES_OK;
}

static int check_out(const decaes_t2map_out *out, const decaes_t2part_opts *part) {
  if (!out) return fail(DECAES_EINVAL, "out is NULL");
  if (!out->gdn || !out->ggm || !out->gva || !out->fnr || !out->snr || !out->alpha || !out->dist)
    return fail(DECAES_EINVAL, "gdn, ggm, gva, fnr, snr, alpha and dist outputs are required");
  if (part && !(out->sfr && out->sgm && out->mfr && out->mgm))
    return fail(DECAES_EINVAL, "fused T2part needs sfr, sgm, mfr and mgm outputs");
  return DECAES_OK;
}

// enqueue setup + pipeline kernels on `stream` for device-resident data
static int launch_pipeline(Plan &plan, int dev, const double *d_image, int64_t nvox, int64_t stride,
                           const decaes_t2map_out *o, cudaStream_t stream, int chunk = 0) {
  PipeParams &P = plan.P;
  if (chunk < 0 || chunk >= DECAES_MAX_CHUNKS) return fail(DECAES_EINVAL, "bad chunk index");
  P.image = d_image, P.nvox = nvox, P.stride = stride;
  P.gdn = o->gdn, P.ggm = o->ggm, P.gva = o->gva, P.fnr = o->fnr, P.snr = o->snr, P.alpha = o->alpha, P.dist = o->dist;
  P.resnorm = o->resnorm, P.decaycurve = o->decaycurve, P.mu = o->mu, P.chi2factor = o->chi2factor;
  P.decaybasis = o->decaybasis;
  P.sfr = o->sfr, P.sgm = o->sgm, P.mfr = o->mfr, P.mgm = o->mgm;
  DeviceWs &ws = g_ws[dev];
  // the previous launch on this device may still be running on another stream: it owns cP, the scratch and (chunk 0)
  // the basis tables until its kernel has finished
  if (ws.ev2_recorded) CUDA_TRY(cudaStreamWaitEvent(stream, ws.ev[2], 0));
  P.counters = ws.counters + 32 * chunk;
  CUDA_TRY(cudaMemsetAsync(P.counters, 0, 32 * sizeof(unsigned long long), stream));
  if (chunk == 0) {  // the per-run tables are shared by every chunk of the slab
    CUDA_TRY(cudaEventRecord(ws.ev[0], stream));
    int nt = plan.S.nA * plan.S.nT2;
    basis_setup_kernel<<<(nt + 63) / 64, 64, 0, stream>>>(plan.S);
    CUDA_TRY(cudaGetLastError());
    if (P.gram) {
      long long ng = (long long)plan.S.nA * plan.S.nT2 * plan.S.nT2;
      CUDA_TRY(cudaMemsetAsync(ws.gram_set, 0, sizeof(double) * (size_t)plan.S.nA * P.a_elems, stream));
      gram_setup_kernel<<<(unsigned)((ng + 127) / 128), 128, 0, stream>>>(plan.S, ws.gram_set, P.ldg, P.a_elems);
      CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(ws.ev[1], stream));
    ws.chunks_pending = 0;
  }
  int64_t ngroups = (nvox + DECAES_GROUP - 1) / DECAES_GROUP;
  int grid = (int)std::min<int64_t>(plan.grid, std::max<int64_t>((ngroups + plan.warps_per_cta - 1) / plan.warps_per_cta, 1));
  CUDA_TRY(cudaMemcpyToSymbolAsync(cP, &P, sizeof(PipeParams), 0, cudaMemcpyHostToDevice, stream));
  // Keep the per-warp scratch (the voxel's basis, 27 KB per warp - 46 KB with the row-major copy - rewritten for every voxel)
  // resident in L2: without the window the streaming image / output traffic evicts it and every basis is
  // written back to and re-read from HBM (ncu: 19 KB of DRAM writes per voxel).
  const size_t scratch_bytes = (size_t)plan.grid * plan.warps_per_cta * P.scratch_per_warp * sizeof(double);
  bool window = ws.l2_persist_max > 0 && ws.l2_window_max > 0 && !getenv("DECAES_NO_L2_WINDOW");
  if (window) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = (void *)ws.scratch;
    attr.accessPolicyWindow.num_bytes = std::min(scratch_bytes, ws.l2_window_max);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ws.l2_persist_max / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = getenv("DECAES_L2_MISS_NORMAL") ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
    if (const char *e = getenv("DECAES_L2_HIT_RATIO")) attr.accessPolicyWindow.hitRatio = (float)atof(e);
    window = cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
    if (!window) cudaGetLastError();
    if (getenv("DECAES_PHASE_CYCLES"))
      fprintf(stderr, "[decaes] L2 window: scratch %.1f MB, persisting carve-out %.1f MB, largest window %.1f MB, hit ratio %.3f, set: %d\n",
              scratch_bytes / 1048576.0, ws.l2_persist_max / 1048576.0, ws.l2_window_max / 1048576.0,
              (double)attr.accessPolicyWindow.hitRatio, (int)window);
  }
  pipeline_kernel_for(P)<<<grid, 32 * plan.warps_per_cta, plan.cta_smem, stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  if (window) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);  // later work on this stream: default policy
  }
  CUDA_TRY(cudaEventRecord(ws.ev[2], stream));  // re-recorded by every chunk: ev[1] -> ev[2] spans all of them
  ws.ev_pending = true;
  ws.ev2_recorded = true;
  ws.chunks_pending = chunk + 1;
  return DECAES_OK;
}

static int collect_device_stats(int dev, decaes_run_stats *st) {
  DeviceWs &ws = g_ws[dev];
  if (!ws.ev_pending) return DECAES_OK;
  CUDA_TRY(cudaEventSynchronize(ws.ev[2]));
  float a = 0, b = 0;
  CUDA_TRY(cudaEventElapsedTime(&a, ws.ev[0], ws.ev[1]));
  CUDA_TRY(cudaEventElapsedTime(&b, ws.ev[1], ws.ev[2]));
  unsigned long long c[32] = {0}, call[DECAES_MAX_CHUNKS * 32];
  const int nch = std::max(1, std::min(ws.chunks_pending, DECAES_MAX_CHUNKS));
  CUDA_TRY(cudaMemcpy(call, ws.counters, (size_t)nch * 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (int k = 0; k < nch; k++)
    for (int i = 0; i < 32; i++) c[i] += call[32 * k + i];
  st->setup_ms = std::max(st->setup_ms, (double)a);
  st->pipeline_ms = std::max(st->pipeline_ms, (double)b);
  st->voxels_processed += (int64_t)c[1];
  st->early_returns += (int64_t)c[2], st->lcurve_overflow += (int64_t)c[3], st->nnls_itercap += (int64_t)c[26];
  if (getenv("DECAES_PHASE_CYCLES"))
    fprintf(stderr, "[decaes] KKT-polish resumptions per voxel: %.4f\n", (double)c[27] / std::max<double>(1.0, (double)c[1]));
  if (getenv("DECAES_PHASE_CYCLES"))
    fprintf(stderr, "[decaes] warp-cycles per voxel: barrier %.0f  flip-angle %.0f  basis %.0f  solve+save %.0f  total %.0f\n",
            (double)c[4] / std::max<double>(1.0, (double)c[1]), (double)c[5] / std::max<double>(1.0, (double)c[1]),
            (double)c[6] / std::max<double>(1.0, (double)c[1]), (double)c[7] / std::max<double>(1.0, (double)c[1]),
            (double)c[24] / std::max<double>(1.0, (double)c[1]));
#ifdef DECAES_PROFILE
  if (getenv("DECAES_PHASE_CYCLES")) {
    const char *nm[PF_COUNT] = {"rhs", "nnls_unreg", "resid", "refine", "grad", "stage", "suggest", "epg", "build", "nnls_tikh", "lc_book", "save"};
    for (int i = 0; i < PF_COUNT; i++) fprintf(stderr, "  %-10s %9.0f\n", nm[i], (double)c[8 + i] / std::max<double>(1.0, (double)c[1]));
    unsigned long long gp[16];
    cudaMemcpyFromSymbol(gp, g_prof, sizeof gp);
    const char *gn[4] = {"append", "factor", "dual", "nnls"};
    double nv = std::max<double>(1.0, (double)c[1]);
    for (int i = 0; i < 4; i++)
      fprintf(stderr, "  gram %-8s cycles/voxel %9.0f  calls/voxel %7.1f  cycles/call %7.0f\n", gn[i], gp[i] / nv, gp[4 + i] / nv,
              gp[4 + i] ? (double)gp[i] / gp[4 + i] : 0.0);
    fprintf(stderr, "  mean k at append %.1f, at factor %.1f; nnls warm %.1f cold %.1f per voxel, mean final k %.1f, mean inner iters %.2f, factor fallbacks/voxel %.3f\n",
            gp[4] ? (double)gp[8] / gp[4] : 0.0, gp[5] ? (double)gp[9] / gp[5] : 0.0, gp[11] / nv, gp[12] / nv,
            gp[7] ? (double)gp[13] / gp[7] : 0.0, gp[7] ? (double)gp[14] / gp[7] : 0.0, gp[15] / nv);
    fprintf(stderr, "  KKT-polish candidate columns per voxel %.2f\n", gp[10] / nv);
    unsigned long long kh[10];
    cudaMemcpyFromSymbol(kh, g_khist, sizeof kh);
    for (int w = 0; w < 2; w++)
      fprintf(stderr, "  k at %s: <=4 %.1f%%  <=8 %.1f%%  <=12 %.1f%%  <=16 %.1f%%  >16 %.1f%%\n", w ? "factor" : "append",
              100.0 * kh[5 * w] / std::max<double>(1, (double)gp[4 + w]), 100.0 * kh[5 * w + 1] / std::max<double>(1, (double)gp[4 + w]),
              100.0 * kh[5 * w + 2] / std::max<double>(1, (double)gp[4 + w]), 100.0 * kh[5 * w + 3] / std::max<double>(1, (double)gp[4 + w]),
              100.0 * kh[5 * w + 4] / std::max<double>(1, (double)gp[4 + w]));
    unsigned long long sh[3][48];
    cudaMemcpyFromSymbol(sh, g_solve_hist, sizeof sh);
    fprintf(stderr, "  solve # (0-27 Tikhonov, 28+ unregularised): calls/voxel, appends/solve, inner iterations/solve\n");
    for (int i = 0; i < 48; i++)
      if (sh[0][i]) fprintf(stderr, "   %2d: %.3f  %.2f  %.2f\n", i, sh[0][i] / nv, (double)sh[1][i] / sh[0][i], (double)sh[2][i] / sh[0][i]);
    static unsigned long long zz[3][48];
    cudaMemcpyToSymbol(g_solve_hist, zz, sizeof zz);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_prof, z, sizeof z);
    cudaMemcpyToSymbol(g_khist, z, sizeof kh);
  }
#endif
  st->kernel_launches += 2 + nch;
  ws.ev_pending = false;
  ws.chunks_pending = 0;
  return DECAES_OK;
}

extern "C" {

const char *decaes_last_error(void) { return g_err; }

int decaes_slab_bounds(int64_t nvox, int32_t nshards, int32_t index, int64_t *v0, int64_t *v1) {
  if (nvox < 0 || nshards < 1 || index < 0 || index >= nshards || !v0 || !v1)
    return fail(DECAES_EINVAL, "bad slab arguments");
  const int64_t g = DECAES_GROUP;
  *v0 = (nvox * index / nshards) & ~(g - 1);
  *v1 = (index == nshards - 1) ? nvox : ((nvox * (index + 1) / nshards) & ~(g - 1));
  return DECAES_OK;
}
int decaes_abi_version(void) { return DECAES_ABI_VERSION; }

int decaes_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void decaes_get_stats(decaes_run_stats *stats) {
  // device-API calls leave their events pending; resolve them now
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
    std::lock_guard<std::recursive_mutex> dl(g_dev_mutex[dev]);
    if (g_ws[dev].ev_pending) {
      g_stats.ngpus_used = 1;
      collect_device_stats(dev, &g_stats);
    }
  }
  if (stats) *stats = g_stats;
}

int decaes_t2map_device(const double *d_image, int64_t nvox, int64_t stride, const decaes_t2map_opts *opts,
                        const decaes_t2part_opts *part, const decaes_t2map_out *d_out, void *stream) {
  int rc = validate_map(opts);
  if (rc) return rc;
  if (part && (rc = validate_part(part))) return rc;
  if ((rc = check_out(d_out, part))) return rc;
  if (!d_image || nvox < 0 || stride < nvox) return fail(DECAES_EINVAL, "bad image pointer / nvox / stride");
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(DECAES_EUNSUPPORTED, "device index %d out of range", dev);
  std::lock_guard<std::recursive_mutex> dl(g_dev_mutex[dev]);
  Plan plan;
  if ((rc = make_plan(opts, part, dev, &plan))) return rc;
  memset(&g_stats, 0, sizeof g_stats);
  g_stats.voxels_total = nvox;
  if (nvox == 0) return DECAES_OK;
  return launch_pipeline(plan, dev, d_image, nvox, stride, d_out, (cudaStream_t)stream);
}

static int launch_part(const decaes_t2part_opts *part, const double *d_dist, int64_t nvox, int64_t stride, double *sfr,
                       double *sgm, double *mfr, double *mgm, cudaStream_t stream) {
  PartTables pt;
  int rc = make_part_tables(part, &pt);
  if (rc) return rc;
  PartParams Q;
  memset(&Q, 0, sizeof Q);
  Q.nT2 = part->nT2, Q.sp_lo = pt.sp_lo, Q.sp_hi = pt.sp_hi, Q.mp_lo = pt.mp_lo, Q.mp_hi = pt.mp_hi;
  Q.has_sigmoid = pt.has_sigmoid, Q.nvox = nvox, Q.stride = stride, Q.dist = d_dist;
  Q.sfr = sfr, Q.sgm = sgm, Q.mfr = mfr, Q.mgm = mgm;
  memcpy(Q.logT2, pt.logT2, sizeof pt.logT2);
  memcpy(Q.weights, pt.weights, sizeof pt.weights);
  if (nvox > 0) {
    t2part_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, stream>>>(Q);
    CUDA_TRY(cudaGetLastError());
  }
  return DECAES_OK;
}

int decaes_t2part_device(const double *d_dist, int64_t nvox, int64_t stride, const decaes_t2part_opts *part,
                         double *d_sfr, double *d_sgm, double *d_mfr, double *d_mgm, void *stream) {
  int rc = validate_part(part);
  if (rc) return rc;
  if (!d_dist || !d_sfr || !d_sgm || !d_mfr || !d_mgm || nvox < 0 || stride < nvox)
    return fail(DECAES_EINVAL, "bad pointers / nvox / stride");
  return launch_part(part, d_dist, nvox, stride, d_sfr, d_sgm, d_mfr, d_mgm, (cudaStream_t)stream);
}

int decaes_mock_image_device(double *d_image, int64_t nvox, int64_t stride, int64_t first_voxel, int32_t nTE,
                             double TE, double T1, double SNR, uint64_t seed, void *stream) {
  if (!d_image || nvox < 0 || stride < nvox || nTE < 4 || nTE > DECAES_MAX_NTE) return fail(DECAES_EINVAL, "bad arguments");
  if (nvox == 0) return DECAES_OK;
  double sigma = std::pow(10.0, -SNR / 20);
  mock_image_kernel<<<(unsigned)((nvox + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_image, nvox, stride, first_voxel,
                                                                                     nTE, TE, T1, sigma, seed);
  CUDA_TRY(cudaGetLastError());
  return DECAES_OK;
}

int decaes_measure_fp64_peak(double *flops_per_s) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
  double *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double) * threads * blocks));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    dfma_peak_kernel<<<blocks, threads>>>(d, iters);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3);
    if (rep > 0) best = std::max(best, fl);
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(d);
  if (flops_per_s) *flops_per_s = best;
  return DECAES_OK;
}

int decaes_setup_tables(const decaes_t2map_opts *o, double *echotimes, double *t2times, double *refangleset,
                        double *decaybasisset) {
  int rc = validate_map(o);
  if (rc) return rc;
  const int nTE = o->nTE, nT2 = o->nT2;
  const bool fixed = !std::isnan(o->SetFlipAngle);
  const int nA = fixed ? 1 : o->nRefAngles;
  if (echotimes)
    for (int i = 0; i < nTE; i++) echotimes[i] = o->TE * (double)(i + 1);  // src/T2mapSEcorr.jl:28
  if (t2times) logrange(o->T2min, o->T2max, nT2, t2times);
  if (refangleset) {
    if (fixed) refangleset[0] = o->SetFlipAngle; else linrange(o->MinRefAngle, 180.0, nA, refangleset);
  }
  if (decaybasisset) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(DECAES_EUNSUPPORTED, "device index %d out of range", dev);
    std::lock_guard<std::recursive_mutex> dl(g_dev_mutex[dev]);
    Plan plan;
    if ((rc = make_plan(o, nullptr, dev, &plan))) return rc;
    if (g_ws[dev].ev2_recorded) CUDA_TRY(cudaEventSynchronize(g_ws[dev].ev[2]));  // a running pipeline reads these tables
    basis_setup_kernel<<<(nA * nT2 + 63) / 64, 64>>>(plan.S);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(decaybasisset, plan.S.basis_cm, sizeof(double) * (size_t)nA * nTE * nT2, cudaMemcpyDeviceToHost));
  }
  return DECAES_OK;
}

// ---------------------------------------------------------------------------- host-pointer API
static std::mutex g_call_mutex;  // re-entrant host calls are serialised
static std::mutex &g_call_mutex_fwd() { return g_call_mutex; }
struct SlabJob {
  int dev;
  int64_t v0, v1;
  int rc;
  char err[512];
  decaes_run_stats st;
};

// Float32 volume -> the Float64 image the pipeline reads (copyto!(Array{Float64,4}(undef, sz), data), src/main.jl:612-617,
// done on the device: the conversion is exact and the host-to-device traffic halves)
__global__ void f32_to_f64_kernel(const float *__restrict__ src, double *__restrict__ dst, long long n, long long pitch, int rows) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  for (int r = 0; r < rows; r++) dst[v + r * pitch] = (double)src[v + r * pitch];
}

// Pinned staging ring of one device (host API, pageable caller buffers): DECAES_NSTAGE slots that alternate between
// "being filled by the host / drained by the device" (H2D) and "being filled by the device / drained by the host" (D2H).
static int ring_ensure(DeviceWs &ws, size_t bytes) {
  if (ws.ring_cap >= bytes) return DECAES_OK;
  for (int s = 0; s < DECAES_NSTAGE; s++) {
    if (ws.ring[s]) cudaFreeHost(ws.ring[s]);
    ws.ring[s] = nullptr;
  }
  ws.ring_cap = 0;
  for (int s = 0; s < DECAES_NSTAGE; s++) {
    if (cudaHostAlloc(&ws.ring[s], bytes, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return fail(DECAES_ENOMEM, "cannot allocate %zu bytes of pinned staging memory", bytes);
    }
    if (!ws.ring_ev[s]) CUDA_TRY(cudaEventCreateWithFlags(&ws.ring_ev[s], cudaEventDisableTiming));
  }
  ws.ring_cap = bytes;
  return DECAES_OK;
}

static bool host_ptr_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// One device's share of a host-API call: voxels [job.v0, job.v1) of the caller's arrays.
//   * the device copy of the slab (image + every requested output) is one cached allocation;
//   * the slab is processed as sub-slabs ("chunks") so that transfers overlap the neighbouring chunks' kernels;
//   * page-locked caller buffers are copied directly (cudaMemcpy2DAsync, strided by Nvox); pageable ones - what Julia
//     hands over (src/T2mapSEcorr.jl:24-54 allocates ordinary Arrays) - go through the device's pinned staging ring:
//     the host thread packs rows into a slot, the copy engine moves the slot, and results come back the same way, so
//     the copies stay asynchronous and at full PCIe rate (a cudaMemcpy from pageable memory would block the host and
//     run at a fraction of it).  cudaHostRegister on the caller's buffers was ruled out: pinning 5.5 GB costs more
//     than the whole 8-GPU run.
static int run_slab(SlabJob &job, const void *image_any, bool img_f32, int64_t Nvox, const decaes_t2map_opts *opts,
                    const decaes_t2part_opts *part, const decaes_t2map_out *out) {
  CUDA_TRY(cudaSetDevice(job.dev));
  if (job.dev < 0 || job.dev >= 64) return fail(DECAES_EUNSUPPORTED, "device index %d out of range", job.dev);
  std::lock_guard<std::recursive_mutex> dl(g_dev_mutex[job.dev]);
  const int nTE = opts->nTE, nT2 = opts->nT2;
  const int64_t nv = job.v1 - job.v0;
  memset(&job.st, 0, sizeof job.st);
  if (nv <= 0) return DECAES_OK;
  Plan plan;
  int rc = make_plan(opts, part, job.dev, &plan);
  if (rc) return rc;
  double *hostp[16] = {out->gdn, out->ggm, out->gva, out->fnr, out->snr, out->alpha, out->dist, out->resnorm,
                       out->decaycurve, out->mu, out->chi2factor, out->decaybasis, out->sfr, out->sgm, out->mfr, out->mgm};
  int64_t mult[16] = {1, 1, 1, 1, 1, 1, nT2, 1, nTE, 1, 1, (int64_t)nTE * nT2, 1, 1, 1, 1};
  if (!std::isnan(opts->SetFlipAngle)) hostp[11] = nullptr;  // shared basis is not a per-voxel output (:581-587)
  if (!part) hostp[12] = hostp[13] = hostp[14] = hostp[15] = nullptr;
  // one allocation for the slab: image + every requested output (+ the Float32 copy of the image), all with stride nv
  size_t total = (size_t)nv * nTE, off[16];
  for (int f = 0; f < 16; f++) {
    off[f] = total;
    if (hostp[f]) total += (size_t)nv * mult[f];
  }
  const size_t off_f32 = total;
  if (img_f32) total += ((size_t)nv * nTE + 1) / 2;
  // The slab buffer is cached per device (grow-only, like the other workspaces): cudaMalloc + cudaFree of the 5.5 GB
  // of a full cfg3 volume cost more than the exposed copies.  decaes_release() or DECAES_SLAB_CACHE=0 give it back.
  static const bool cache_slab = !(getenv("DECAES_SLAB_CACHE") && atoi(getenv("DECAES_SLAB_CACHE")) == 0);
  double *dbuf = nullptr;
  if (cache_slab) {
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    DeviceWs &ws = g_ws[job.dev];
    if ((rc = ensure(&ws.slab, &ws.slab_cap, total))) return rc;
    dbuf = ws.slab;
  } else {
    CUDA_TRY(cudaMalloc(&dbuf, total * sizeof(double)));
  }
  decaes_t2map_out dout;
  double **dptr = (double **)&dout;
  for (int f = 0; f < 16; f++) dptr[f] = hostp[f] ? dbuf + off[f] : nullptr;
  float *dimg32 = img_f32 ? (float *)(dbuf + off_f32) : nullptr;
  const size_t elem = img_f32 ? sizeof(float) : sizeof(double);
  const char *image = (const char *)image_any;

  // pageable caller buffers -> pinned staging ring (DECAES_STAGING=0/1 overrides the detection)
  bool staged = !host_ptr_is_pinned(image_any);
  for (int f = 0; f < 16 && !staged; f++)
    if (hostp[f] && !host_ptr_is_pinned(hostp[f])) staged = true;
  if (const char *e = getenv("DECAES_STAGING")) staged = atoi(e) != 0;
  job.st.pinned_staging = staged;

  // Sub-slabs.  Exposed transfer time = first chunk in + last chunk out, so in staged mode (where the host packs and
  // unpacks every byte) the first and the last chunk are made small.
  int nchunks = (int)std::min<int64_t>(staged ? 8 : 4, std::max<int64_t>(1, nv / 65536));
  if (const char *e = getenv("DECAES_CHUNKS")) nchunks = std::max(1, std::min(DECAES_MAX_CHUNKS, atoi(e)));
  nchunks = (int)std::min<int64_t>(nchunks, nv);
  int64_t cb[DECAES_MAX_CHUNKS + 1];
  {
    double wsum = 0, acc = 0, wt[DECAES_MAX_CHUNKS];
    for (int c = 0; c < nchunks; c++) wt[c] = (staged && nchunks >= 4 && (c == 0 || c + 1 == nchunks)) ? 0.25 : 1.0, wsum += wt[c];
    cb[0] = 0;
    for (int c = 0; c < nchunks; c++) acc += wt[c], cb[c + 1] = ((int64_t)((double)nv * acc / wsum) / DECAES_GROUP) * DECAES_GROUP;
    cb[nchunks] = nv;
    for (int c = 0; c < nchunks; c++) cb[c + 1] = std::max(cb[c + 1], cb[c]);
  }
  int64_t wmax = 0;
  for (int c = 0; c < nchunks; c++) wmax = std::max(wmax, cb[c + 1] - cb[c]);

  DeviceWs &dws = g_ws[job.dev];
  if (staged) {
    size_t slot = (size_t)32 << 20;
    if (const char *e = getenv("DECAES_STAGE_MB")) slot = (size_t)std::max(1, atoi(e)) << 20;
    slot = std::max(slot, (size_t)wmax * sizeof(double));
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    if ((rc = ring_ensure(dws, slot))) return rc;
  }
  cudaStream_t st, sc, sd;
  CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));  // kernels
  CUDA_TRY(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));  // host -> device
  CUDA_TRY(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));  // device -> host
  cudaEvent_t e0, e1, e2, e3, evin[DECAES_MAX_CHUNKS], evdone[DECAES_MAX_CHUNKS];
  cudaEventCreate(&e0), cudaEventCreate(&e1), cudaEventCreate(&e2), cudaEventCreate(&e3);
  for (int c = 0; c < nchunks; c++)
    cudaEventCreateWithFlags(&evin[c], cudaEventDisableTiming), cudaEventCreateWithFlags(&evdone[c], cudaEventDisableTiming);

  // ---- staging ring bookkeeping ----
  struct Pending {
    double *host;   // first destination row in the caller's array
    int64_t w;      // doubles per row
    int rows;
    bool active;
  } pend[DECAES_NSTAGE];
  for (int s = 0; s < DECAES_NSTAGE; s++) pend[s].active = false;
  int ring_next = 0;
  auto acquire = [&]() -> int {  // next slot, free of device work and with its results handed to the caller
    const int s = ring_next;
    ring_next = (ring_next + 1) % DECAES_NSTAGE;
    cudaEventSynchronize(dws.ring_ev[s]);
    if (pend[s].active) {
      const double *src = (const double *)dws.ring[s];
      for (int r = 0; r < pend[s].rows; r++) memcpy(pend[s].host + (size_t)r * Nvox, src + (size_t)r * pend[s].w, (size_t)pend[s].w * sizeof(double));
      pend[s].active = false;
    }
    return s;
  };
  auto t_host0 = std::chrono::steady_clock::now();
  double staged_h2d_first_ms = 0, staged_d2h_last_ms = 0;

  CUDA_TRY(cudaEventRecord(e0, sc));
  for (int c = 0; c < nchunks; c++) {
    const int64_t a = cb[c], b = cb[c + 1], w = b - a;
    if (w > 0) {
      char *ddst = img_f32 ? (char *)(dimg32 + a) : (char *)(dbuf + a);
      if (!staged) {
        CUDA_TRY(cudaMemcpy2DAsync(ddst, nv * elem, image + (size_t)(job.v0 + a) * elem, Nvox * elem, w * elem, nTE,
                                   cudaMemcpyHostToDevice, sc));
        if (opts->alpha_provided)
          CUDA_TRY(cudaMemcpyAsync(dout.alpha + a, out->alpha + job.v0 + a, w * sizeof(double), cudaMemcpyHostToDevice, sc));
      } else {
        const int rows_per_slot = (int)std::max<size_t>(1, dws.ring_cap / ((size_t)w * elem));
        for (int r0 = 0; r0 < nTE; r0 += rows_per_slot) {
          const int nr = std::min(rows_per_slot, nTE - r0);
          const int sidx = acquire();
          char *slotp = (char *)dws.ring[sidx];
          for (int r = 0; r < nr; r++)
            memcpy(slotp + (size_t)r * w * elem, image + ((size_t)(r0 + r) * Nvox + job.v0 + a) * elem, (size_t)w * elem);
          CUDA_TRY(cudaMemcpy2DAsync(ddst + (size_t)r0 * nv * elem, nv * elem, slotp, w * elem, w * elem, nr, cudaMemcpyHostToDevice, sc));
          CUDA_TRY(cudaEventRecord(dws.ring_ev[sidx], sc));
        }
        if (opts->alpha_provided) {
          const int sidx = acquire();
          memcpy(dws.ring[sidx], out->alpha + job.v0 + a, (size_t)w * sizeof(double));
          CUDA_TRY(cudaMemcpyAsync(dout.alpha + a, dws.ring[sidx], w * sizeof(double), cudaMemcpyHostToDevice, sc));
          CUDA_TRY(cudaEventRecord(dws.ring_ev[sidx], sc));
        }
      }
    }
    CUDA_TRY(cudaEventRecord(evin[c], sc));
    if (c == 0) {
      CUDA_TRY(cudaEventRecord(e1, sc));
      staged_h2d_first_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    }
    // the kernel of this chunk is enqueued as soon as its input is on its way
    decaes_t2map_out dc = dout;
    double **dcp = (double **)&dc;
    for (int f = 0; f < 16; f++)
      if (dcp[f]) dcp[f] += a;
    CUDA_TRY(cudaStreamWaitEvent(st, evin[c], 0));
    if (w > 0) {
      if (img_f32) {
        f32_to_f64_kernel<<<(unsigned)((w + 255) / 256), 256, 0, st>>>(dimg32 + a, dbuf + a, w, nv, nTE);
        CUDA_TRY(cudaGetLastError());
        job.st.kernel_launches += 1;
      }
      rc = launch_pipeline(plan, job.dev, dbuf + a, w, nv, &dc, st, c);
      if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(evdone[c], st));
    if (!staged) {
      CUDA_TRY(cudaStreamWaitEvent(sd, evdone[c], 0));
      if (c + 1 == nchunks) CUDA_TRY(cudaEventRecord(e2, sd));
      for (int f = 0; f < 16; f++)
        if (hostp[f] && w > 0)
          CUDA_TRY(cudaMemcpy2DAsync(hostp[f] + job.v0 + a, Nvox * sizeof(double), dptr[f] + a, nv * sizeof(double),
                                     w * sizeof(double), mult[f], cudaMemcpyDeviceToHost, sd));
    }
  }
  if (staged) {
    // results: chunk by chunk as the kernels finish, device -> slot -> caller's arrays, DECAES_NSTAGE - 1 slots in flight
    for (int c = 0; c < nchunks; c++) {
      const int64_t a = cb[c], b = cb[c + 1], w = b - a;
      CUDA_TRY(cudaStreamWaitEvent(sd, evdone[c], 0));
      auto t_last0 = std::chrono::steady_clock::now();
      if (c + 1 == nchunks) {
        CUDA_TRY(cudaEventRecord(e2, sd));
        cudaEventSynchronize(evdone[c]);
        t_last0 = std::chrono::steady_clock::now();
      }
      if (w > 0) {
        const int rows_per_slot = (int)std::max<size_t>(1, dws.ring_cap / ((size_t)w * sizeof(double)));
        for (int f = 0; f < 16; f++) {
          if (!hostp[f]) continue;
          for (int64_t r0 = 0; r0 < mult[f]; r0 += rows_per_slot) {
            const int nr = (int)std::min<int64_t>(rows_per_slot, mult[f] - r0);
            const int sidx = acquire();
            CUDA_TRY(cudaMemcpy2DAsync(dws.ring[sidx], w * sizeof(double), dptr[f] + a + (size_t)r0 * nv, nv * sizeof(double),
                                       w * sizeof(double), nr, cudaMemcpyDeviceToHost, sd));
            CUDA_TRY(cudaEventRecord(dws.ring_ev[sidx], sd));
            pend[sidx].host = hostp[f] + job.v0 + a + (size_t)r0 * Nvox, pend[sidx].w = w, pend[sidx].rows = nr, pend[sidx].active = true;
          }
        }
      }
      if (c + 1 == nchunks) {
        for (int s = 0; s < DECAES_NSTAGE; s++) acquire();  // drain what is still in flight
        staged_d2h_last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_last0).count();
      }
    }
  }
  CUDA_TRY(cudaEventRecord(e3, sd));
  CUDA_TRY(cudaStreamSynchronize(sc));
  CUDA_TRY(cudaStreamSynchronize(sd));
  CUDA_TRY(cudaStreamSynchronize(st));
  float h2d = 0, d2h = 0;
  cudaEventElapsedTime(&h2d, e0, e1), cudaEventElapsedTime(&d2h, e2, e3);
  // exposed parts: first sub-slab in, last sub-slab out (staged mode: host wall time of the same two steps)
  job.st.h2d_ms = staged ? staged_h2d_first_ms : h2d, job.st.d2h_ms = staged ? staged_d2h_last_ms : d2h;
  const int extra_launches = job.st.kernel_launches;
  rc = collect_device_stats(job.dev, &job.st);
  (void)extra_launches;
  cudaEventDestroy(e0), cudaEventDestroy(e1), cudaEventDestroy(e2), cudaEventDestroy(e3);
  for (int c = 0; c < nchunks; c++) cudaEventDestroy(evin[c]), cudaEventDestroy(evdone[c]);
  if (!cache_slab) cudaFree(dbuf);
  cudaStreamDestroy(sc);
  cudaStreamDestroy(sd);
  cudaStreamDestroy(st);
  return rc;
}

void decaes_release(void) {
  std::lock_guard<std::mutex> lk(g_call_mutex_fwd());
  std::lock_guard<std::mutex> lk2(g_ws_mutex);
  int cur = 0, ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  cudaGetDevice(&cur);
  for (int d = 0; d < ndev && d < 64; d++) {
    DeviceWs &ws = g_ws[d];
    if (!(ws.slab || ws.scratch || ws.basis_rm || ws.basis_cm || ws.gram_set || ws.ring[0])) continue;
    cudaSetDevice(d);
    cudaDeviceSynchronize();
    cudaFree(ws.slab), cudaFree(ws.scratch), cudaFree(ws.basis_rm), cudaFree(ws.basis_cm), cudaFree(ws.dbasis_cm), cudaFree(ws.gram_set);
    ws.slab = ws.scratch = ws.basis_rm = ws.basis_cm = ws.dbasis_cm = ws.gram_set = nullptr;
    for (int r = 0; r < DECAES_NSTAGE; r++) {
      if (ws.ring[r]) cudaFreeHost(ws.ring[r]);
      ws.ring[r] = nullptr;
    }
    ws.ring_cap = 0;
    ws.slab_cap = ws.scratch_cap = ws.basis_rm_cap = ws.basis_cm_cap = ws.gram_cap = 0;
  }
  cudaSetDevice(cur);
}


// Slab boundaries for `ng` devices.  Work is proportional to the number of voxels above Threshold
// (src/T2mapSEcorr.jl:177), not to the number of voxels: a brain-masked volume is 60-70 % background and the
// background is not spread evenly over the slowest dimension, so equal-length slabs would leave the devices that own
// the top and bottom slices idle.  One pass over the first echo counts foreground voxels per block of 1024 and the
// cuts are placed on the block boundaries that split the COUNT evenly (an unmasked volume gives equal lengths).
extern "C++" {
template <typename T>
static void balanced_bounds(const T *first_echo, int64_t Nvox, double threshold, int ng, std::vector<int64_t> &cut) {
  const int64_t B = 1024, nb = (Nvox + B - 1) / B;
  std::vector<int64_t> cnt((size_t)nb + 1, 0);
  const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(8, nb / 256));
  auto work = [&](int t) {
    for (int64_t k = nb * t / nthreads; k < nb * (t + 1) / nthreads; k++) {
      int64_t c = 0;
      const int64_t e = std::min(Nvox, (k + 1) * B);
      for (int64_t v = k * B; v < e; v++) c += ((double)first_echo[v] > threshold);
      cnt[(size_t)k + 1] = c;
    }
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto &t : th) t.join();
  }
  for (int64_t k = 0; k < nb; k++) cnt[(size_t)k + 1] += cnt[(size_t)k];
  const int64_t totalfg = cnt[(size_t)nb];
  cut.assign((size_t)ng + 1, 0);
  cut[(size_t)ng] = Nvox;
  for (int d = 1; d < ng; d++) {
    if (totalfg == 0) {
      cut[(size_t)d] = ((Nvox * d / ng) / DECAES_GROUP) * DECAES_GROUP;
      continue;
    }
    const int64_t want = totalfg * d / ng;
    const int64_t k = std::lower_bound(cnt.begin(), cnt.end(), want) - cnt.begin();
    cut[(size_t)d] = std::min(Nvox, k * B);  // B is a multiple of the 4-voxel work group
  }
  for (int d = 1; d <= ng; d++) cut[(size_t)d] = std::max(cut[(size_t)d], cut[(size_t)d - 1]);
}
}  // extern "C++"

static int t2map_host(const void *image, bool img_f32, const decaes_t2map_opts *opts, const decaes_t2part_opts *part,
                      const decaes_t2map_out *out) {
  std::lock_guard<std::mutex> lk(g_call_mutex);
  auto t0 = std::chrono::steady_clock::now();
  int rc = validate_map(opts);
  if (rc) return rc;
  if (part && (rc = validate_part(part))) return rc;
  if ((rc = check_out(out, part))) return rc;
  if (!image) return fail(DECAES_EINVAL, "image is NULL");
  int ndev = decaes_device_count();
  if (ndev < 1) return fail(DECAES_ECUDA, "no CUDA device available (libdecaes_cuda has no CPU fallback)");
  int ng = opts->ngpus > 0 ? std::min(opts->ngpus, ndev) : ndev;
  const int64_t Nvox = (int64_t)opts->nx * opts->ny * opts->nz;
  int cur = 0;
  cudaGetDevice(&cur);
  std::vector<SlabJob> jobs(ng);
  // contiguous voxel slabs, boundaries aligned to the 4-voxel work group; equal foreground counts when a
  // threshold is in force and more than one device takes part
  std::vector<int64_t> cut;
  const bool balance = ng > 1 && opts->Threshold > -INFINITY && !(getenv("DECAES_BALANCE") && atoi(getenv("DECAES_BALANCE")) == 0);
  if (balance) {
    if (img_f32) balanced_bounds((const float *)image, Nvox, opts->Threshold, ng, cut);
    else balanced_bounds((const double *)image, Nvox, opts->Threshold, ng, cut);
  }
  for (int d = 0; d < ng; d++) {
    int64_t a = 0, b = 0;
    if (balance) a = cut[(size_t)d], b = cut[(size_t)d + 1];
    else decaes_slab_bounds(Nvox, ng, d, &a, &b);
    jobs[d].dev = (ng == 1) ? cur : d, jobs[d].v0 = a, jobs[d].v1 = b, jobs[d].rc = 0, jobs[d].err[0] = 0;
  }
  auto worker = [&](SlabJob &j) {
    j.rc = run_slab(j, image, img_f32, Nvox, opts, part, out);
    if (j.rc) snprintf(j.err, sizeof j.err, "%s", g_err);
  };
  if (ng == 1) {
    worker(jobs[0]);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < ng; d++) th.emplace_back(worker, std::ref(jobs[d]));
    for (auto &t : th) t.join();
  }
  cudaSetDevice(cur);
  memset(&g_stats, 0, sizeof g_stats);
  g_stats.voxels_total = Nvox, g_stats.ngpus_used = ng;
  for (auto &j : jobs) {
    if (j.rc) return fail(j.rc, "device %d: %s", j.dev, j.err);
    g_stats.voxels_processed += j.st.voxels_processed;
    g_stats.early_returns += j.st.early_returns, g_stats.lcurve_overflow += j.st.lcurve_overflow;
    g_stats.nnls_itercap += j.st.nnls_itercap, g_stats.pinned_staging |= j.st.pinned_staging;
    g_stats.kernel_launches += j.st.kernel_launches;
    g_stats.setup_ms = std::max(g_stats.setup_ms, j.st.setup_ms);
    g_stats.pipeline_ms = std::max(g_stats.pipeline_ms, j.st.pipeline_ms);
    g_stats.h2d_ms = std::max(g_stats.h2d_ms, j.st.h2d_ms);
    g_stats.d2h_ms = std::max(g_stats.d2h_ms, j.st.d2h_ms);
  }
  g_stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return DECAES_OK;
}

int decaes_slab_bounds_masked(const double *first_echo, int64_t nvox, double threshold, int32_t nshards, int64_t *cuts) {
  if (!first_echo || !cuts || nvox < 0 || nshards < 1) return fail(DECAES_EINVAL, "bad slab arguments");
  std::vector<int64_t> cut;
  balanced_bounds(first_echo, nvox, threshold, nshards, cut);
  for (int d = 0; d <= nshards; d++) cuts[d] = cut[(size_t)d];
  return DECAES_OK;
}

int decaes_t2map(const double *image, const decaes_t2map_opts *opts, const decaes_t2part_opts *part,
                 const decaes_t2map_out *out) {
  return t2map_host(image, false, opts, part, out);
}

int decaes_t2map_f32(const float *image, const decaes_t2map_opts *opts, const decaes_t2part_opts *part,
                     const decaes_t2map_out *out) {
  return t2map_host(image, true, opts, part, out);
}

void *decaes_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    fail(DECAES_ENOMEM, "cudaHostAlloc of %zu bytes failed", bytes);
    return nullptr;
  }
  return p;
}
void decaes_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int decaes_t2part(const double *dist, const decaes_t2part_opts *part, double *sfr, double *sgm, double *mfr,
                  double *mgm) {
  std::lock_guard<std::mutex> lk(g_call_mutex);
  int rc = validate_part(part);
  if (rc) return rc;
  if (!dist || !sfr || !sgm || !mfr || !mgm) return fail(DECAES_EINVAL, "NULL pointer");
  if (decaes_device_count() < 1) return fail(DECAES_ECUDA, "no CUDA device available (libdecaes_cuda has no CPU fallback)");
  const int64_t Nvox = (int64_t)part->nx * part->ny * part->nz;
  const int nT2 = part->nT2;
  double *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double) * (size_t)Nvox * (nT2 + 4)));
  double *o[4] = {d + (size_t)Nvox * nT2, d + (size_t)Nvox * (nT2 + 1), d + (size_t)Nvox * (nT2 + 2), d + (size_t)Nvox * (nT2 + 3)};
  double *h[4] = {sfr, sgm, mfr, mgm};
  CUDA_TRY(cudaMemcpy(d, dist, sizeof(double) * (size_t)Nvox * nT2, cudaMemcpyHostToDevice));
  // outputs keep the caller's pre-fill (NaN) wherever the reference would not write
  for (int k = 0; k < 4; k++) CUDA_TRY(cudaMemcpy(o[k], h[k], sizeof(double) * Nvox, cudaMemcpyHostToDevice));
  rc = launch_part(part, d, Nvox, Nvox, o[0], o[1], o[2], o[3], 0);
  if (rc) {
    cudaFree(d);
    return rc;
  }
  for (int k = 0; k < 4; k++) CUDA_TRY(cudaMemcpy(h[k], o[k], sizeof(double) * Nvox, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaFree(d));
  memset(&g_stats, 0, sizeof g_stats);
  g_stats.voxels_total = Nvox, g_stats.ngpus_used = 1, g_stats.kernel_launches = 1;
  return DECAES_OK;
}

}  // extern "C"
