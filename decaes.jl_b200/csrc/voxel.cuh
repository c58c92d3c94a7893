// Per-voxel pipeline executed by one warp: normalise -> flip-angle fit -> EPG basis at the fitted
// angle -> regularised NNLS (none / lcurve / gcv / chi2 / mdp) -> output maps + T2part epilogue.
//
// Reference: voxelwise_T2_distribution! src/T2mapSEcorr.jl:201-238, optimize_flip_angle! :409-423,
// T2_distribution! :475-505, save_results! :512-591, voxelwise_T2_parts! src/T2partSEcorr.jl:95-138,
// surrogate search src/splines.jl:504-566, 705-850, 975-1041, choosers src/lsqnonneg.jl,
// 1-D optimisers src/optimization.jl:71-128, 191-219, 319-413.
#pragma once
#include "nnls.cuh"
#include "gram.cuh"
#include "legacy.cuh"

namespace decaes {

struct PipeParams {
  // sizes
  int nTE, nT2, ld, nA, reg;
  int legacy;           // legacy = true (src/types.jl:20-21): sampled-spline flip-angle and chi2 searches (legacy.cuh)
  LegacyRange lg_range; // angles[0]:0.001:angles[nA-1] as Julia builds it
  int rows_alloc;       // rows of the shared working matrix (nTE or nTE + nT2)
  int a_elems;          // doubles reserved for the working matrix / EPG scratch per warp
  int copy_elems;       // doubles per TMA bulk copy of one nTE x nT2 matrix (even)
  int nseed, maxeval;
  int fixed_alpha;      // SetFlipAngle given
  int alpha_provided;   // B1 map in out.alpha
  int epg_kmax;         // phase states kept by the shared-memory EPG
  int epg_lanes;        // components per pass of the shared-memory EPG (n split evenly over ceil(n/32) passes)
  int has_part, sp_lo, sp_hi, mp_lo, mp_hi, has_sigmoid;
  int8_t seeds[DECAES_MAX_ANGLES];
  double TE, T1, Threshold, Chi2Factor, NoiseLevel, SetFlipAngle, E1, refcon;
  // volume
  const double *image;
  long long nvox, stride;
  // outputs (device pointers, may be null)
  double *gdn, *ggm, *gva, *fnr, *snr, *alpha, *dist, *resnorm, *decaycurve, *mu, *chi2factor, *decaybasis;
  double *sfr, *sgm, *mfr, *mgm;
  // tables
  const double *basis_rm;   // [nA][copy_elems]  row-major, leading dimension ld (TMA source)
  const double *basis_cm;   // [nA][nT2][nTE]
  const double *dbasis_cm;  // [nA][nT2][nTE]   d/d(alpha in degrees)
  const double *gram_set;   // [nA][nT2*ldg]     A_k'A_k per grid angle (Gram solver)
  int gram, ldg;            // solver variant (1 = normal-equation active set), leading dimension of G
  int gv_stride;            // stride of the Gram solver's vectors in shared memory (gram.cuh: 40 or 64, >= nT2)
  int warps_per_cta, smem_per_warp;  // CTA shape: bytes of dynamic shared memory owned by each warp
  int fa_warm;              // warm-start flip-angle probes from a probed angle at most this many grid steps away (0 = never)
  int spill;                // which per-voxel tables live in global scratch instead of shared memory (SmemLayout)
  int sync_groups;          // 1: CTA-wide phase barriers; g > 1: one barrier per group of warps (warp id mod g)
  int sync_mask;            // which intra-round CTA barriers are active (bit 0: after flip angle, bit 1: after basis)
  int epg_smem;             // EPG at the fitted angle keeps its states in shared memory (lane <-> component)
  int epg_fuse;             // shared-memory EPG: two echoes per sweep over the states
  int need_rm;              // the voxel's basis is also kept row-major in the global scratch (0: column-major only, c = A'b comes out of the EPG)
  int gcv_smem;             // Reg = gcv: the singular values are computed in shared memory during the basis phase (2: bidiagonalisation +
                            // multisection, gcv_svdvals_bidiag; 1: parallel one-sided Jacobi, gcv_svdvals_smem; 0: global-memory Jacobi)
  int warm_ones;            // warm starts of the Tikhonov solves begin at x = 1 on the inherited set (0: at the cached solution of the nearest mu)
  int refine_tikh;          // polish every Tikhonov solve with one refinement step (Gram solver)
  int fa_polish;            // KKT polish (explicit duals) on the flip-angle probes too: 1 = all probes, 2 = all but the seed probes
  double kkt_tau;           // screening threshold of the polish: duals above -kkt_tau * max|c| are recomputed explicitly
  int fa_refine;            // iterative-refinement step on the flip-angle probes
  int fa_rough_seeds;       // seed probes without refinement / polish; a seed that ends up in the final bracket is probed again precisely
  int step_sync;            // CTA-wide votes inside the phases (cta_or): bit 0 = before every flip-angle probe, bit 1 = at every step of the
                            // L-curve search, bit 2 = at its four initial points too, bit 3 = at every other regularised solve (Brent searches)
  int lc_hints;             // L-curve start hints: bit 0 = full column set at mu = e^2, bit 1 = flip-angle fit's set at mu = e^-8
  double angles[DECAES_MAX_ANGLES];                            // flip-angle grid (degrees)
  double logT2[DECAES_MAX_NT2], E2[DECAES_MAX_NT2];            // log(T2_j), exp(-(TE/2)/T2_j)
  double weights[DECAES_MAX_NT2];                              // sigmoid weights (has_sigmoid)
  // scratch + scheduling
  double *scratch;          // per-warp global scratch
  long long scratch_per_warp;
  unsigned long long *counters;  // [0] work counter, [1] voxels processed, [2] early returns, [3] lcurve overflow, [26] NNLS iteration caps
};

// layout of the per-warp global scratch (in doubles)
struct ScratchLayout {
  int pristine, pristine_cm, slots_x, lc_pts, lc_states, fa_u, fa_du, fa_mask, gcv_gamma, gcv_mat, total;
  // rm: keep a row-major copy of the voxel's basis next to the column-major one (QR port, shuffle EPG, global-memory SVD)
  __host__ __device__ ScratchLayout(int nTE, int nT2, int copy_elems, bool gcv, bool rm = true) {
    int o = 0;
    pristine = o, o += rm ? copy_elems : 0;
    pristine_cm = o, o += nTE * nT2 + (nTE * nT2 & 1);
    slots_x = o, o += DECAES_NCACHE * nT2;
    lc_pts = o, o += DECAES_LC_MAX * 4;
    lc_states = o, o += DECAES_LC_MAX * 5;
    fa_u = o, o += DECAES_MAX_ANGLES;
    fa_du = o, o += DECAES_MAX_ANGLES;
    fa_mask = o, o += DECAES_MAX_ANGLES;
    gcv_gamma = o, o += (nTE < nT2 ? nTE : nT2);
    gcv_mat = o, o += gcv ? nTE * nT2 : 0;
    total = (o + 1) & ~1;
  }
};

// per-warp shared memory layout (in doubles, then ints)
struct SmemLayout {
  int A, b, u, x, w, bd, sig, fit, slot_mu, slot_lmu, slot_r2, slot_x2, slot_mask, idx, bar, total_bytes;
  int M, c, y, s, t1, t2, lc_pts, lc_states, slots_x, fa_u, fa_du, fa_mask;  // Gram solver only
  // spill (Gram solver): bit 0 = the cached solutions (slots_x), bit 1 = the L-curve state records live in the warp's
  // global scratch instead (one L2 round trip per solve / per L-curve step, 2.5 KB each of shared memory back)
  __host__ __device__ SmemLayout(int nTE, int nT2, int rows_alloc, int a_elems, int gram, int spill = 0, int vs = 64) {
    int o = 0;
    b = u = M = c = y = s = t1 = t2 = lc_pts = lc_states = slots_x = fa_u = fa_du = fa_mask = 0;
    if (gram) {
      // solver block first (gram.cuh: vectors at fixed offsets from V, then the combined G / M array)
      c = 0, y = vs, s = 2 * vs, t1 = 3 * vs, t2 = 4 * vs, x = 5 * vs, w = 6 * vs, idx = 7 * vs;  // = GV_LAYOUT(vs)
      A = gv_block_doubles(vs), o = A + a_elems;
      lc_pts = o, o += 4 * DECAES_LC_MAX;
      if (!(spill & 2)) lc_states = o, o += 5 * DECAES_LC_MAX;
      if (!(spill & 1)) slots_x = o, o += DECAES_NCACHE * nT2;
      // the flip-angle tables are dead once the angle is fitted: they alias the L-curve caches
      fa_u = lc_pts, fa_du = lc_pts + DECAES_MAX_ANGLES, fa_mask = lc_pts + 2 * DECAES_MAX_ANGLES;
      static_assert(3 * DECAES_MAX_ANGLES <= 9 * DECAES_LC_MAX, "flip-angle tables must fit in the L-curve caches");
    } else {
      A = o, o += a_elems;  // working matrix / EPG scratch
      b = o, o += rows_alloc;
      u = o, o += rows_alloc;
      x = o, o += nT2;
      w = o, o += nT2;
    }
    // everything in front of `bd` is dead while the EPG basis is built (the EPG keeps its phase states there);
    // the voxel's signal and the mbarrier must survive it
    fit = o, o += nTE;
    slot_mu = o, o += DECAES_NCACHE;
    slot_lmu = o, o += DECAES_NCACHE;
    slot_r2 = o, o += DECAES_NCACHE;
    slot_x2 = o, o += DECAES_NCACHE;
    slot_mask = o, o += DECAES_NCACHE;
    bd = o, o += nTE;
    sig = 0;
    bar = o, o += 2;
    if (!gram) idx = o, o += (nT2 + 1) / 2;
    total_bytes = ((o * 8) + 15) & ~15;
  }
};

struct GramWs {  // views into the solver block of one warp (gram.cuh layout)
  double *y, *s, *x, *w, *t1, *t2;
  int *P;
};

struct Src {               // where the current basis of the Gram solver lives
  const double *G;         // n x n Gram matrix (shared or global)
  int ldg;
  const double *Arm;       // row-major nTE x ld
  const double *Acm;       // column-major [col * nTE + i]
};

#define DECAES_PRAGMA_(x) _Pragma(#x)
#define DECAES_PRAGMA(x) DECAES_PRAGMA_(x)
#ifndef DECAES_RESID_CHUNK
#define DECAES_RESID_CHUNK 4  // columns per L2 round trip of the explicit residual (8: -0.9 % on cfg3 with today's code, -3 % on the nT2 = 60 configs)
#endif
#ifndef DECAES_EPG_UNROLL
#define DECAES_EPG_UNROLL 2  // state loop of the shared-memory EPG (independent iterations: unrolling buys ILP, costs code; +1.3 % with two echoes per sweep)
#endif
#ifdef DECAES_PROFILE
#define PROF_BEGIN(id) long long prof_t0_##id = clock64()
#define PROF_END(id) prof_cyc[id] += clock64() - prof_t0_##id
#else
#define PROF_BEGIN(id)
#define PROF_END(id)
#endif
enum { PF_RHS = 0, PF_NNLS_UNREG, PF_RESID, PF_REFINE, PF_GRAD, PF_STAGE, PF_SUGGEST, PF_EPG, PF_BUILD, PF_NNLS_TIKH,
       PF_LC_BOOK, PF_SAVE, PF_COUNT };

// The kernel parameters as seen by the per-warp code: a __constant__ copy (written by the host right
// before each launch, stream-ordered), so that every field is one LDC with an immediate offset instead
// of a pointer chase through *this.  One launch per device at a time (the host API serialises calls).
__constant__ PipeParams cP;
#define GL(p) __builtin_assume(__isGlobal(p))

// CTA-wide OR on a barrier of its own (id 8): keeps the warps of a CTA in step INSIDE a phase.  The kernel is bound by
// instruction fetch (ncu: the GPC-level instruction cache serves requests at 93-95 % of its peak rate when the twelve
// warps of an SM wander through ~40 KB of hot code independently; the SM's own instruction cache hits 67-70 %).  Warps
// that start every NNLS solve at the same time execute the same lines at about the same time and share them: +20 %
// throughput for a vote per solve.  Protocol: a warp that has work votes true at each of its sync points; a warp that
// is done with the phase (or has no voxel) keeps voting false until a vote comes back false (cta_drain).
// Measured and rejected (profiles/r02_s2_ab_votes.txt): votes within groups of 6 or 4 warps (-6 % / -27 %: two or three
// instruction streams per SM again), a vote every second or third L-curve step (-3 % / -4 %), votes at every
// active-set iteration (-47 %), at the four initial L-curve points (-1 %), at every solve of the Brent searches (-6 %).
__device__ __forceinline__ bool cta_or(bool pred) {
  int r;
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %2, 0;\nbarrier.red.or.pred q, 8, %1, p;\nselp.b32 %0, 1, 0, q;\n}\n"
      : "=r"(r)
      : "r"((int)blockDim.x), "r"((int)pred)
      : "memory");
  return r != 0;
}
__device__ __noinline__ void cta_drain() {
  while (cta_or(false)) {
  }
}

// Member functions copy the pointers they use into locals and tell the compiler that they point to
// shared memory: otherwise every access is a generic LD/ST, and every store forces a reload of the
// members of *this (which lives in local memory) because it might alias them.
#define SH(p) __builtin_assume(__isShared(p))
#define SHG(p) do { if constexpr (GRAM) __builtin_assume(__isShared(p)); } while (0)
#define VIEW(T, name) T *const name = this->name; SH(name)
#define VIEWG(T, name) T *const name = this->name; SHG(name)
#define VIEW_GWS() const GramWs gws = this->gws; SH(gws.y); SH(gws.s); SH(gws.x); SH(gws.w); SH(gws.t1); SH(gws.t2); SH(gws.P)

// LEGACY instantiates the legacy = true searches (legacy.cuh); the default instantiation carries none of that code
template <bool GRAM, bool LEGACY = false, int VS = 64>
struct Warp {
  long long prof_cyc[PF_COUNT] = {0};
  Src cursrc;
  NnlsWs ws;
  GramWs gws;                    // Gram solver scratch
  double *V;                     // base of the Gram solver block in shared memory (gram.cuh layout)
  double *Gs, *cvec;             // per-voxel G = A'A and c = A'b in shared memory (Gram solver)
  unsigned long long *slot_mask; // active set of each cache slot
  // small control tables: shared memory for the Gram solver (an L2 round trip per access would
  // dominate the latency of the scalar search code), global scratch for the QR port
  double *lc_pts_p, *lc_states_p, *slots_x_p, *fa_u_p, *fa_du_p;
  unsigned long long *fa_mask_p;
  double *bd, *sig, *fit, *slot_mu, *slot_lmu, *slot_r2, *slot_x2;
  unsigned long long fa_rough = 0ull; // grid angles probed without refinement / polish so far
  double c_epg[2];                  // c = A'b of the fitted-angle basis, accumulated by the shared-memory EPG (one entry per pass)
  int solve_vote = 0;               // step_sync bit that makes the regularised solves vote right now (0: they do not)
  unsigned long long fa_mask_best;  // active set of the probed grid angle nearest to the fitted one (0 = none)
  uint64_t *bar;
  unsigned phase;
  double *g;  // global scratch of this warp
  ScratchLayout sl;
  int dirty_rows;   // lambda rows of the working matrix that may be non-zero
  int cur_slot;     // 0-based current cache slot (persists across voxels like work.idx[])
  int lane;
  unsigned long long n_early, n_overflow, n_itercap, n_polish;
  bool gcv_fixed_done = false;             // Reg = gcv with SetFlipAngle: singular values of the one basis already computed by this warp
  int nsolve_voxel = 0, nunreg_voxel = 0;  // Tikhonov / unregularised solves of the current voxel (DECAES_PROFILE)

  __device__ Warp(const PipeParams &p, double *smem, double *gscratch)
      : sl(p.nTE, p.nT2, p.copy_elems, p.reg == 2, p.need_rm != 0) {
    SmemLayout L(p.nTE, p.nT2, p.rows_alloc, p.a_elems, p.gram, p.spill, p.gv_stride);
    ws.A = smem + L.A, ws.b = smem + L.b, ws.u = smem + L.u, ws.x = smem + L.x, ws.w = smem + L.w;
    ws.idx = (int *)(smem + L.idx);
    ws.ld = p.ld, ws.n = p.nT2, ws.m0 = p.nTE;
    V = smem, Gs = smem + L.A, cvec = smem + L.c;
    gws.y = smem + L.y, gws.s = smem + L.s, gws.x = smem + L.x, gws.w = smem + L.w;
    gws.t1 = smem + L.t1, gws.t2 = smem + L.t2, gws.P = (int *)(smem + L.idx);
    slot_mask = (unsigned long long *)(smem + L.slot_mask);
    if (p.gram) {
      lc_pts_p = smem + L.lc_pts;
      lc_states_p = (p.spill & 2) ? gscratch + sl.lc_states : smem + L.lc_states;
      slots_x_p = (p.spill & 1) ? gscratch + sl.slots_x : smem + L.slots_x;
      fa_u_p = smem + L.fa_u, fa_du_p = smem + L.fa_du, fa_mask_p = (unsigned long long *)(smem + L.fa_mask);
    } else {
      lc_pts_p = gscratch + sl.lc_pts, lc_states_p = gscratch + sl.lc_states, slots_x_p = gscratch + sl.slots_x;
      fa_u_p = gscratch + sl.fa_u, fa_du_p = gscratch + sl.fa_du, fa_mask_p = (unsigned long long *)(gscratch + sl.fa_mask);
    }
    bd = smem + L.bd, sig = nullptr, fit = smem + L.fit;
    slot_mu = smem + L.slot_mu, slot_lmu = smem + L.slot_lmu, slot_r2 = smem + L.slot_r2, slot_x2 = smem + L.slot_x2;
    fa_mask_best = 0ull;
    bar = (uint64_t *)(smem + L.bar);
    phase = 0;
    g = gscratch;
    dirty_rows = p.rows_alloc - p.nTE;
    cur_slot = 0;
    lane = lane_id();
    n_early = n_overflow = n_itercap = n_polish = 0;
  }

  // ---- TMA: stage one nTE x nT2 matrix (row-major, ld) from global into the working matrix ----
  __device__ __noinline__ void stage_matrix(const double *src) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_expect_tx(bar, (unsigned)(cP.copy_elems * 8));
      tma_bulk_g2s(ws.A, src, (unsigned)(cP.copy_elems * 8), bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
  }

  // TMA bulk copy of `bytes` (multiple of 16) from global to shared memory, whole warp waits
  __device__ __noinline__ void stage_bulk(void *dst, const void *src, unsigned bytes) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_expect_tx(bar, bytes);
      tma_bulk_g2s(dst, src, bytes, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
  }

  // ================= flip-angle fit =================
  // loss_with_grad!  src/splines.jl:1010-1041
  __device__ __noinline__ void fa_eval(int k, double &u, double &du) {
    stage_matrix(cP.basis_rm + (size_t)k * cP.copy_elems);
    nnls_warm_start<false>(ws, bd, 0.0, 0);
    NnlsOut o = nnls_core<false>(ws, 0.0);
    u = o.rnorm_sq;
    const double *Ak = cP.basis_cm + (size_t)k * cP.nTE * cP.nT2;
    const double *dAk = cP.dbasis_cm + (size_t)k * cP.nTE * cP.nT2;
    double acc = 0.0;
    for (int i = lane; i < cP.nTE; i += 32) {
      double ax = 0.0, dax = 0.0;
      for (int j = 0; j < cP.nT2; j++) {
        double xj = ws.x[j];
        if (xj > 0.0) {
          ax = fma(xj, Ak[j * cP.nTE + i], ax);
          dax = fma(xj, dAk[j * cP.nTE + i], dax);
        }
      }
      acc = fma(dax, ax - bd[i], acc);
    }
    du = 2.0 * warp_sum(acc);
  }

  // CubicHermiteInterpolator + minimize  src/splines.jl:62-110
  __device__ __noinline__ static void hermite_minimize(double a, double b, double u0, double u1, double m0, double m1,
                                          double &xo, double &uo) {
    double r = (b - a) / 2;
    m0 = __dmul_rn(r, m0), m1 = __dmul_rn(r, m1);
    double du = u1 - u0, dm = m1 - m0, su = u1 + u0, sm = m1 + m0;
    double c0 = __dsub_rn(su / 2, dm / 4), c1 = __dsub_rn(__dmul_rn(3.0, du), sm) / 4, c2 = dm / 4, c3 = (sm - du) / 4;
    double xend = (u0 < u1) ? a : b, uend = (u0 < u1) ? u0 : u1;
    double D = __dmul_rn(3.0, u0 - u1);
    double th = __dadd_rn(D / 2, m0 + m1);
    double gg = __dsub_rn(__dmul_rn(th, th), __dmul_rn(m0, m1));
    gg = gg > 0 ? -sqrt(gg) : 0.0;
    double p = -__dadd_rn(D, m0 + m1);
    double q = __dadd_rn(__dmul_rn(2.0, gg), m0 - m1);
    xo = xend, uo = uend;
    if (fabs(p) < fabs(q)) {
      double t = p / q;
      double y = fma(t, fma(t, fma(t, c3, c2), c1), c0);
      if (y < uend) {
        double c = (a + b) / 2, rr = (b - a) / 2;
        double x = fma(rr, t, c);
        x = x < a ? a : (x > b ? b : x);
        xo = x, uo = y;
      }
    }
  }

  // suggest_point  src/splines.jl:544-566.  Pieces are evaluated in parallel (lane <-> piece);
  // "first strictly smaller wins" of the sequential scan = lexicographic min over (value, order).
  __device__ __noinline__ void suggest_point(unsigned long long seen, double &xs, double &us) {
    const int lane = this->lane;
    VIEWG(double, fa_u_p);
    VIEWG(double, fa_du_p);
    const double *fu = fa_u_p, *fdu = fa_du_p;
    int npts = __popcll(seen);
    double bestu = CUDART_INF, bestx = 0.0;
    int besto = 0x7fffffff;
    // piece t connects the t-th and (t+1)-th probed node; order -1 is the first node itself
    _Pragma("unroll 1") for (int t = lane - 1; t < npts - 1; t += 32) {
      double x_, u_;
      if (t < 0) {
        int I0 = __ffsll((long long)seen) - 1;
        x_ = cP.angles[I0], u_ = fu[I0];
      } else {
        // indices of the t-th and (t+1)-th set bits
        unsigned long long msk = seen;
        _Pragma("unroll 1") for (int s = 0; s < t; s++) msk &= msk - 1;
        int Ia = __ffsll((long long)msk) - 1;
        msk &= msk - 1;
        int Ib = __ffsll((long long)msk) - 1;
        hermite_minimize(cP.angles[Ia], cP.angles[Ib], fu[Ia], fu[Ib], fdu[Ia], fdu[Ib], x_, u_);
      }
      int ord = t + 1;
      // sequential semantics: the seed node always starts the scan; later candidates replace the
      // current best only when strictly smaller (NaN never replaces)
      bool better = (besto == 0x7fffffff) ? (ord == 0 || u_ == u_) : (u_ < bestu);
      if (better) bestu = u_, bestx = x_, besto = ord;
    }
    // warp reduction on the REDUX unit: smaller value wins, ties -> smaller order; NaN loses
    const unsigned long long key = (besto == 0x7fffffff) ? ~0ull : (bestu != bestu ? ~0ull - 1ull : dkey(bestu));
    unsigned long long best;
    const int ord = warp_argmin_bits(key, besto, best);
    const int winner = __ffs(__ballot_sync(DECAES_FULL_MASK, key == best && besto == ord)) - 1;
    bestx = __shfl_sync(DECAES_FULL_MASK, bestx, winner), bestu = __shfl_sync(DECAES_FULL_MASK, bestu, winner);
    xs = warp_bcast(bestx, 0), us = warp_bcast(bestu, 0);
  }

  // Workspace of the legacy searches in the warp's global scratch: the L-curve / flip-angle tables of the QR port
  // (ScratchLayout lc_pts .. fa_mask, 768 doubles) are idle whenever these run.  X | Y | t | c | z | a
  struct LegacyWs {
    double *X, *Y, *t, *c, *z, *a;
  };
  __device__ __forceinline__ LegacyWs legacy_ws() const {
    double *lw = g + sl.lc_pts;
    GL(lw);
    LegacyWs w;
    w.X = lw, w.Y = lw + 64, w.t = lw + 128, w.c = lw + 196, w.z = lw + 260, w.a = lw + 324;  // 324 + 4*64 = 580 <= 768
    return w;
  }

  // suggest_point of CubicSplineSurrogate(...; legacy = true)  src/splines.jl:492-500 -> spline_opt_legacy :419-430
  __device__ __noinline__ void suggest_point_legacy(unsigned long long seen, double &xs, double &us) {
    VIEWG(double, fa_u_p);
    const LegacyWs w = legacy_ws();
    const int m = __popcll(seen), k = m - 1 < 3 ? m - 1 : 3;
    if (lane == 0) {
      int q = 0;
      for (unsigned long long msk = seen; msk; msk &= msk - 1, q++) {
        const int I = __ffsll((long long)msk) - 1;
        w.X[q] = cP.angles[I], w.Y[q] = fa_u_p[I];
      }
      fitpack_interp_dev(w.X, w.Y, m, k, w.t, w.c, w.z, w.a);
    }
    __syncwarp();
    legacy_spline_scan(w.t, w.c, m, k, cP.lg_range, 0, 0.0, xs, us);
  }

  // lsqnonneg_chi2!(...; method = :legacy)  src/lsqnonneg.jl:521-533 with chi2_search_from_minimum :595-636:
  // mu doubles from 1e-3 until res2(mu) >= Chi2Factor * res2_min, then the sampled root of the spline through every
  // (mu, res2) seen (mu = 0 included) on the grid 0:0.001:mu_last.  Returns -1 when the doubling does not end
  // within DECAES_LEGACY_CHI2_MAXPTS points (the reference would keep doubling).
  __device__ __noinline__ double chi2_legacy(double res2_min, const double *Asrc) {
    const LegacyWs w = legacy_ws();
    const double target = __dmul_rn(cP.Chi2Factor, res2_min);
    int m = 1;
    if (lane == 0) w.X[0] = 0.0, w.Y[0] = res2_min;
    double munew = 1e-3;
    bool ok = false;
    _Pragma("unroll 1") while (m < DECAES_LEGACY_CHI2_MAXPTS) {
      cache_solve(munew, Asrc);
      const double r = cur_resnorm_sq();
      if (lane == 0) w.X[m] = munew, w.Y[m] = r;
      m++;
      if (r >= target) {
        ok = true;
        break;
      }
      munew *= 2.0;
    }
    if (!ok) return -1.0;
    const int k = m - 1 < 3 ? m - 1 : 3;
    if (lane == 0) fitpack_interp_dev(w.X, w.Y, m, k, w.t, w.c, w.z, w.a);
    __syncwarp();
    LegacyRange r;  // 0.0:0.001:(1e-3 * 2^(m-2)) lifts to the rationals i/1000, i = 0 .. 2^(m-2)
    r.rational = 1, r.start_n = 0, r.step_n = 1, r.den = 1000, r.len = (1ll << (m - 2)) + 1, r.start = 0.0, r.step = 0.001;
    double mu, yr;
    legacy_spline_scan(w.t, w.c, m, k, r, 1, target, mu, yr);
    return mu;
  }

  // EPG basis at `alpha` for this voxel (+ Gram matrix and right-hand side for the Gram solver)
  __device__ void basis_at(double alpha, long long v) {
    if constexpr (GRAM) {
      PROF_BEGIN(7);
      if (cP.refcon != 180.0) epg_basis_beta(alpha, v, V);
      else if (cP.epg_smem) epg_basis(alpha, v, V);  // lane <-> T2 component, states in the (idle) solver block
      else if (cP.nTE <= 63) epg_basis_shfl<false>(alpha, v);
      else epg_basis_shfl<true>(alpha, v);
      PROF_END(7);
      // the basis lives column-major in the scratch; a row-major copy (element (i, j) at i * ld + j) only when need_rm
      const bool rm = cP.need_rm != 0;
      const double *Ab = rm ? g + sl.pristine : g + sl.pristine_cm;
      const int rs = rm ? cP.ld : 1, cs = rm ? 1 : cP.nTE;
      if (cP.reg == 2 && cP.gcv_smem) gcv_svd_shared(Ab, rs, cs, V);
      if (!rm && cP.epg_smem) {  // right-hand side c = A'b straight from the EPG's registers
        const int LW = cP.epg_lanes;
        if (lane < LW && lane < cP.nT2) cvec[lane] = c_epg[0];
        if (lane < LW && LW + lane < cP.nT2) cvec[LW + lane] = c_epg[1];
        __syncwarp();
      }
      PROF_BEGIN(8);
      gram_build(Ab, rs, cs, rm || !cP.epg_smem);
      PROF_END(8);
      cursrc.G = Gs, cursrc.ldg = cP.ldg, cursrc.Arm = g + sl.pristine, cursrc.Acm = g + sl.pristine_cm;
    } else {
      epg_basis(alpha, v, ws.A);
    }
  }

  // rough: solve without refinement and KKT polish (the loss is second-order accurate, its slope to ~cond^2 eps: good
  // enough to steer the search, not to define the fitted angle)
  __device__ __noinline__ void fa_probe(int I, unsigned long long &seen, int &numeval, bool rough = false) {
    double u, du;
    if (cP.step_sync & 1) cta_or(true);
    if constexpr (GRAM) fa_eval_gram(I, u, du, seen, rough);
    else fa_eval(I, u, du);
    if (lane == 0) fa_u_p[I] = u, fa_du_p[I] = du;
    __syncwarp();
    if (rough) fa_rough |= 1ull << I;
    else fa_rough &= ~(1ull << I);
    if (!((seen >> I) & 1ull)) numeval++;
    seen |= (1ull << I);
  }

  // DiscreteSurrogateSearcher + bisection_search  src/splines.jl:705-850 (D = 1).  The seed order
  // produced by initialize! depends only on (nRefAngles, nRefAnglesMin) and is computed on the host.
  __device__ __noinline__ double optimize_flip_angle() {
    unsigned long long seen = 0ull;
    int numeval = 0;
    const int maxeval = cP.maxeval, nA = cP.nA;
    fa_rough = 0ull;
#ifdef DECAES_FA_ROUGH
    for (int s = 0; s < cP.nseed; s++) fa_probe(cP.seeds[s], seen, numeval, GRAM && !LEGACY && cP.fa_rough_seeds);
#else
    _Pragma("unroll 1") for (int s = 0; s < cP.nseed; s++) fa_probe(cP.seeds[s], seen, numeval);
#endif
    double x, u;
    unsigned long long seen_sugg = 0ull;  // legacy: the scan is expensive, skip it when nothing new was probed
    if constexpr (LEGACY) suggest_point_legacy(seen, x, u), seen_sugg = seen;
    else suggest_point(seen, x, u);
    while (true) {
      // minimal_bounding_box :778-800
      int lo = 0, hi = nA - 1;
      while (true) {
        int mid = (lo + hi) / 2;
        bool in_left = (cP.angles[lo] <= x) && (x <= cP.angles[mid]);
        int plo = in_left ? lo : mid, phi = in_left ? mid : hi;
        bool evaluated = ((seen >> plo) & 1ull) && ((seen >> phi) & 1ull);
        lo = plo, hi = phi;
        if (!evaluated || !(hi - lo > 1)) break;
      }
      // evaluate_box! :802-815 with corners sorted by distance to x (stable)
      {
        double d0 = __dmul_rn(cP.angles[lo] - x, cP.angles[lo] - x), d1 = __dmul_rn(cP.angles[hi] - x, cP.angles[hi] - x);
        int c0 = lo, c1 = hi;
        if (d1 < d0) c0 = hi, c1 = lo;
        int cs[2] = {c0, c1};
        for (int k = 0; k < 2; k++) {
          if (((seen >> lo) & 1ull) && ((seen >> hi) & 1ull)) break;
          if (numeval >= maxeval) break;
          if ((seen >> cs[k]) & 1ull) continue;
          fa_probe(cs[k], seen, numeval);
          if (numeval >= maxeval) break;
        }
      }
      if constexpr (!LEGACY) suggest_point(seen, x, u);
      else if (seen != seen_sugg) suggest_point_legacy(seen, x, u), seen_sugg = seen;
      if (numeval >= maxeval || (hi - lo) <= 1) {
#ifdef DECAES_FA_ROUGH  // rejected experiment (see PipeParams::fa_rough_seeds), kept out of the shipped kernel
        if constexpr (GRAM && !LEGACY) {
          // the fitted angle is the minimum of the Hermite piece over [lo, hi]: both ends must be precise.  A seed that
          // was probed roughly is probed again (warm start from its own active set) and the search resumes from there.
          const unsigned long long need = fa_rough & seen & ((1ull << lo) | (1ull << hi));
          if (need) {
            if ((need >> lo) & 1ull) fa_probe(lo, seen, numeval);
            if (hi != lo && ((need >> hi) & 1ull)) fa_probe(hi, seen, numeval);
            suggest_point(seen, x, u);
            continue;
          }
        }
#endif
        if constexpr (GRAM) {
          // active set of the probed node nearest to the fitted angle: warm start of the first regularised solve
          int In = (fabs(cP.angles[lo] - x) <= fabs(cP.angles[hi] - x)) ? lo : hi;
          if (!((seen >> In) & 1ull)) In = lo + hi - In;
          fa_mask_best = ((seen >> In) & 1ull) ? fa_mask_p[In] : 0ull;
        }
        break;
      }
    }
    return x;
  }

  // ================= EPG basis at the fitted angle =================
  // lane <-> T2 component; phase states of each lane's curve live in shared memory (the working
  // matrix region is free at this point), updated in place.  Arithmetic follows
  // epg_impulse_response! src/EPGdecaycurve.jl:948-1028 operation by operation.
  __device__ __noinline__ void epg_basis(double alpha_deg, long long v, double *S /* [3][K][32] shared scratch */) {
    const int lane = this->lane;
    SH(S);
    const int ETL = cP.nTE, n = cP.nT2, ld = cP.ld;
    const int K = cP.epg_kmax;
    double *pr = g + sl.pristine;
    double *pc = GRAM ? g + sl.pristine_cm : pr;
    GL(pr);
    GL(pc);
    VIEW(double, bd);
    const bool rm = !GRAM || cP.need_rm != 0;  // row-major copy wanted (otherwise column-major only, and c = A'b on the fly)
    double sina, cosa;
    sincos(alpha_deg * 0.017453292519943295, &sina, &cosa);
    const double m0 = sind_0_180(alpha_deg / 2);
    const double E1 = cP.E1;
    // the three state arrays never alias: lets the compiler overlap the loads of one state with the stores of the previous one
    // LW lanes per pass (PipeParams::epg_lanes).  The phase is bound by the shared-memory pipe (all warps of the CTA run
    // it at the same time) and a 64-bit access costs one wavefront per HALF-warp that has an active lane, so the host
    // picks the split with the fewest active half-warps (nT2 = 40: 24 + 16 lanes = 2 + 1 wavefronts per access instead
    // of 2 + 2 for 20 + 20) and the idle lanes of a pass sit the whole recursion out (no loads, no stores).
    const int LW = cP.epg_lanes, LS = LW;
    for (int j0 = 0; j0 < n; j0 += LW) {
      const int j = j0 + lane;
      if (lane < LW && j < n) {
        double *const sF = S + lane, *const sB = S + K * LS + lane, *const sZ = S + 2 * K * LS + lane;
#define ST(c, k) ((c) == 0 ? sF : (c) == 1 ? sB : sZ)[((k)-1) * LS]
        const double E2 = cP.E2[j];
        const double E2h = __dmul_rn(E2, E2) / 2, E1E2 = __dmul_rn(E1, E2), E1sq = __dmul_rn(E1, E1);
        const double a = E2h, b = __dmul_rn(E2h, cosa), c = __dmul_rn(E1E2, sina), d = __dmul_rn(E1sq, cosa);
        const double cp = -c / 2;
        double F, Fb, Z, C, Sd, Cp, Sp, vF, vFb, vZ;
#define UPD()                                   \
  C = __dadd_rn(F, Fb), Sd = __dsub_rn(F, Fb);  \
  Cp = __dmul_rn(a, C), Sp = __dmul_rn(b, Sd);  \
  vFb = fma(-c, Z, __dsub_rn(Cp, Sp));          \
  vF = fma(c, Z, __dadd_rn(Cp, Sp));            \
  vZ = fma(cp, Sd, __dmul_rn(d, Z))
        double dc = __dsub_rn(a, b);
        double cacc;  // c_j = sum_i A(i, j) b_i in gram_atv's order (i ascending, fma)
        {
          const double val = fabs(__dmul_rn(m0, dc));
          if (rm) pr[0 * ld + j] = val;
          if (GRAM) pc[j * ETL] = val;
          cacc = fma(val, bd[0], 0.0);
        }
        ST(0, 1) = __dsub_rn(a, b), ST(1, 1) = 0.0, ST(2, 1) = cp;
        ST(0, 2) = __dadd_rn(a, b), ST(1, 2) = 0.0, ST(2, 2) = 0.0;
        // Two echoes per sweep over the states when both lie in the same half of the train (the state count grows by
        // one per echo in the first half and shrinks by one in the second): echo i at state k + 1 feeds echo i + 1 at
        // state k (new F_k = F'_{k-1}, new Fbar_k = Fbar'_{k+1}, new Z_k = Z'_k), so one pass of 3 loads + 3 stores per
        // state carries two updates - the phase is bound by the shared-memory pipe.  Same operations on the same
        // values as one echo at a time (bit-identical basis; DECAES_EPG_FUSE=0 for the A/B).
        int i = 2;
        while (i <= ETL - 1) {
          const bool first_half = (i <= ETL / 2);
          const int kmax = first_half ? i : ETL - i + 1;
          F = ST(0, 1), Fb = ST(1, 1), Z = ST(2, 1);
          UPD();
          {
            const double val = fabs(__dmul_rn(m0, vFb));
            if (rm) pr[(i - 1) * ld + j] = val;
            if (GRAM) pc[j * ETL + i - 1] = val;
            cacc = fma(val, bd[i - 1], cacc);
          }
          if (cP.epg_fuse && i + 1 <= ETL - 1 && ((i + 1 <= ETL / 2) == first_half)) {
            const int ka = kmax, kb = first_half ? ka + 1 : ka - 1;
            double f2 = vFb, f1 = vF, z1 = vZ;  // new F_1 (= Fbar'_1), F'_1 (-> new F_2), new Z_1
            DECAES_PRAGMA(unroll DECAES_EPG_UNROLL) for (int k = 1; k <= kb; k++) {
              double nFb = 0.0, nf1 = 0.0, nz1 = 0.0;  // first half: the states beyond ka are zero after echo i
              if (k + 1 <= ka) {
                F = ST(0, k + 1), Fb = ST(1, k + 1), Z = ST(2, k + 1);  // echo i at state k + 1
                UPD();
                nFb = vFb, nf1 = vF, nz1 = vZ;
              }
              F = f2, Fb = nFb, Z = z1;  // echo i + 1 at state k
              UPD();
              if (k == 1) {
                const double val = fabs(__dmul_rn(m0, vFb));
                if (rm) pr[i * ld + j] = val;
                if (GRAM) pc[j * ETL + i] = val;
                cacc = fma(val, bd[i], cacc);
                ST(0, 1) = vFb;
              } else {
                ST(1, k - 1) = vFb;
              }
              ST(0, k + 1) = vF, ST(2, k) = vZ;
              f2 = f1, f1 = nf1, z1 = nz1;
            }
            if (first_half) ST(1, i + 1) = 0.0, ST(1, i + 2) = 0.0, ST(2, i + 2) = 0.0;
            i += 2;
            continue;
          }
          ST(0, 1) = vFb, ST(2, 1) = vZ;
          double pend = vF;
          DECAES_PRAGMA(unroll DECAES_EPG_UNROLL) for (int k = 2; k <= kmax; k++) {
            F = ST(0, k), Fb = ST(1, k), Z = ST(2, k);
            ST(0, k) = pend;
            UPD();
            pend = vF;
            ST(1, k - 1) = vFb;
            ST(2, k) = vZ;
          }
          ST(0, kmax + 1) = pend;
          if (first_half) ST(1, i) = 0.0, ST(1, i + 1) = 0.0, ST(2, i + 1) = 0.0;
          i += 1;
        }
        F = ST(0, 1), Fb = ST(1, 1), Z = ST(2, 1);
        C = __dadd_rn(F, Fb), Sd = __dsub_rn(F, Fb);
        dc = fma(-c, Z, fma(a, C, __dmul_rn(-b, Sd)));
        {
          const double val = fabs(__dmul_rn(m0, dc));
          if (rm) pr[(ETL - 1) * ld + j] = val;
          if (GRAM) pc[j * ETL + ETL - 1] = val;
          cacc = fma(val, bd[ETL - 1], cacc);
        }
        c_epg[j0 != 0] = cacc;  // at most two passes (LW >= nT2 / ceil(nT2 / 32))
      }
      __syncwarp();
    }
#undef ST
#undef UPD
    dirty_rows = cP.rows_alloc - cP.nTE;  // the scratch overlapped the lambda rows
    if (cP.decaybasis && !cP.fixed_alpha) {
      __syncwarp();
      for (int k = lane; k < ETL * n; k += 32) {
        int i = k % ETL, jj = k / ETL;
        cP.decaybasis[v + (long long)k * cP.stride] = GRAM ? pc[k] : pr[i * ld + jj];
      }
    }
    __threadfence_block();
    __syncwarp();
  }

  // EPG basis at `alpha` for RefConAngle != 180 (src/EPGdecaycurve.jl:722-818): first refocusing pulse
  // alpha, later pulses alpha * beta / 180.  lane <-> T2 component, phase states in shared memory,
  // updated in place (each old state M_j yields F.M_j -> new F_{j+1}, Fbar.M_j -> new Fbar_{j-1},
  // Z.M_j -> new Z_j); dot products in the reference's order.
  __device__ __noinline__ void epg_basis_beta(double alpha_deg, long long v, double *S /* [3][K][32] shared scratch */) {
    const int lane = this->lane;
    SH(S);
    const int ETL = cP.nTE, n = cP.nT2, ld = cP.ld;
    const int K = cP.epg_kmax;
    double *pr = g + sl.pristine, *pc = g + sl.pristine_cm;
    GL(pr);
    GL(pc);
    VIEW(double, bd);
    const bool rm = cP.need_rm != 0;
    const double kk = 0.017453292519943295;
    const double A = alpha_deg / 180;
    double sh, ch, sini, cosi;
    sincos(__dmul_rn(__dmul_rn(A, 180.0), kk) / 2, &sh, &ch);
    sincos(__dmul_rn(__dmul_rn(A, cP.refcon), kk), &sini, &cosi);
    const double s2h = __dmul_rn(sh, sh), c2h = __dmul_rn(ch, ch), sin1 = __dmul_rn(__dmul_rn(2.0, sh), ch);
    const double c2hi = (1 + cosi) / 2, s2hi = 1 - c2hi;
    const double E1 = cP.E1, m0 = sh;
    const int LW = cP.epg_lanes, LS = LW;  // rows of LW doubles; the idle lanes of a pass sit it out (see epg_basis)
#define ST(c, k) S[((c)*K + (k)-1) * LS + lane]
#define DOT3(u0, u1, u2) __dadd_rn(__dadd_rn(__dmul_rn(u0, mF), __dmul_rn(u1, mFb)), __dmul_rn(u2, mZ))
    _Pragma("unroll 1") for (int j0 = 0; j0 < n; j0 += LW) {
      const int j = j0 + lane;
      if (lane < LW && j < n) {
        const double E2 = cP.E2[j];
        const double E2sq = __dmul_rn(E2, E2), E1E2 = __dmul_rn(E1, E2);
        const double a1 = __dmul_rn(E2sq, c2h), b1 = __dmul_rn(E2sq, s2h), c1 = __dmul_rn(E1E2, sin1);
        const double ai = __dmul_rn(E2sq, c2hi), bi = __dmul_rn(E2sq, s2hi), ci = __dmul_rn(E1E2, sini);
        const double di = __dmul_rn(__dmul_rn(E1, E1), cosi), hci = ci / 2;
        double mF, mFb, mZ, FM, FbM, ZM, cacc;
        ST(0, 1) = __dmul_rn(b1, m0), ST(1, 1) = 0.0, ST(2, 1) = __dmul_rn(-c1, m0) / 2;
        ST(0, 2) = __dmul_rn(a1, m0), ST(1, 2) = 0.0, ST(2, 2) = 0.0;
        {
          const double val = fabs(__dmul_rn(b1, m0));
          if (rm) pr[j] = val;
          pc[j * ETL] = val;
          cacc = fma(val, bd[0], 0.0);
        }
        _Pragma("unroll 1") for (int i = 2; i <= ETL - 1; i++) {
          const bool first_half = (i <= ETL / 2);
          const int nproc = first_half ? i : ETL - i + 1;
          mF = ST(0, 1), mFb = ST(1, 1), mZ = ST(2, 1);
          FM = DOT3(ai, bi, ci), FbM = DOT3(bi, ai, -ci), ZM = DOT3(-hci, hci, di);
          {
            const double val = fabs(FbM);
            if (rm) pr[(i - 1) * ld + j] = val;
            pc[j * ETL + i - 1] = val;
            cacc = fma(val, bd[i - 1], cacc);
          }
          ST(0, 1) = FbM, ST(2, 1) = ZM;
          double pend = FM;
          _Pragma("unroll 1") for (int k = 2; k <= nproc; k++) {
            mF = ST(0, k), mFb = ST(1, k), mZ = ST(2, k);
            FM = DOT3(ai, bi, ci), FbM = DOT3(bi, ai, -ci), ZM = DOT3(-hci, hci, di);
            ST(0, k) = pend;
            pend = FM;
            ST(1, k - 1) = FbM;
            ST(2, k) = ZM;
          }
          if (first_half) ST(0, nproc + 1) = pend, ST(1, nproc) = 0.0, ST(1, nproc + 1) = 0.0, ST(2, nproc + 1) = 0.0;
        }
        mF = ST(0, 1), mFb = ST(1, 1), mZ = ST(2, 1);
        {
          const double val = fabs(DOT3(bi, ai, -ci));
          if (rm) pr[(ETL - 1) * ld + j] = val;
          pc[j * ETL + ETL - 1] = val;
          cacc = fma(val, bd[ETL - 1], cacc);
        }
        c_epg[j0 != 0] = cacc;
      }
      __syncwarp();
    }
#undef ST
#undef DOT3
    __syncwarp();
    if (cP.decaybasis && !cP.fixed_alpha) {
      _Pragma("unroll 1") for (int k = lane; k < ETL * n; k += 32) cP.decaybasis[v + (long long)k * cP.stride] = pc[k];
    }
    __threadfence_block();
    __syncwarp();
  }

  // ================= regularised solves =================
  __device__ __noinline__ NnlsOut solve_unreg(const double *Asrc) {
    if constexpr (GRAM) {
      GramOut go;
      NnlsOut o;
      o.rnorm_sq = gram_solve_unreg(cursrc, go);
      o.xnorm_sq = go.xnorm_sq, o.nsetp = go.k, o.rows_used = 0;
      return o;
    } else {
      stage_matrix(Asrc);
      nnls_warm_start<false>(ws, bd, 0.0, 0);
      return nnls_core<false>(ws, 0.0);
    }
  }

  // solve!(cache, mu)  src/lsqnonneg.jl:417-444 — exact-mu hit or solve into the next slot
  __device__ void cache_reset() {
    nsolve_voxel = 0;
    if (lane < DECAES_NCACHE) slot_mu[lane] = CUDART_NAN;
    __syncwarp();
  }
  // lmu = log(mu) when the caller has it (NaN: computed on a miss); hint (Gram solver): 1 = start from the full
  // column set (heavily regularised solves keep nearly every column), 2 = start from the flip-angle fit's active set
  __device__ __noinline__ void cache_solve(double mu, const double *Asrc, double lmu = NAN, int hint = 0) {
    if constexpr (GRAM) {
      cache_solve_gram(mu, cursrc, lmu, hint);
      return;
    }
    int hit = -1, firstnan = -1;
    for (int i = 0; i < DECAES_NCACHE; i++) {
      double mui = slot_mu[i];
      if (isnan(mui)) {
        if (firstnan < 0) firstnan = i;
      } else if (mu == mui) {
        hit = i;
        break;
      }
    }
    if (hit >= 0) {
      cur_slot = hit;
      return;
    }
    cur_slot = (firstnan >= 0) ? firstnan : (cur_slot + 1) % DECAES_NCACHE;
    stage_matrix(Asrc);
    nnls_warm_start<true>(ws, bd, mu, dirty_rows);
    NnlsOut o = nnls_core<true>(ws, mu);
    dirty_rows = o.rows_used - cP.nTE;
    double *sx = slots_x_p + cur_slot * cP.nT2;
    for (int j = lane; j < cP.nT2; j += 32) sx[j] = ws.x[j];
    if (lane == 0) slot_mu[cur_slot] = mu, slot_r2[cur_slot] = o.rnorm_sq, slot_x2[cur_slot] = o.xnorm_sq;
    __syncwarp();
  }
  __device__ __forceinline__ double cur_seminorm_sq() const { return slot_x2[cur_slot]; }
  __device__ __forceinline__ double cur_resnorm_sq() const {  // src/lsqnonneg.jl:309-313
    if constexpr (GRAM) return slot_r2[cur_slot];  // stored as the explicit ||Ax - b||^2
    double mu = slot_mu[cur_slot];
    double r = slot_r2[cur_slot] - __dmul_rn(__dmul_rn(mu, mu), slot_x2[cur_slot]);
    return r > 0 ? r : 0.0;
  }

  // ---- L-curve  src/lsqnonneg.jl:812-972 ----
  static __device__ __forceinline__ bool isapprox(double x, double y) {
    if (x == y) return true;
    if (!isfinite(x) || !isfinite(y)) return false;
    return fabs(x - y) <= 1.4901161193847656e-08 * fmax(fabs(x), fabs(y));
  }
  static __device__ __forceinline__ bool isless_f(double a, double b) {
    if (isnan(a)) return false;
    if (isnan(b)) return true;
    if (a == 0 && b == 0) return signbit(a) && !signbit(b);
    return a < b;
  }
  static __device__ __forceinline__ double norm2(double ax, double ay, double bx, double by) {
    double dx = ax - bx, dy = ay - by;
    return dsqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  }
  static __device__ double menger(double jx, double jy, double kx, double ky, double lx, double ly) {
    double jk0 = jx - kx, jk1 = jy - ky, kl0 = kx - lx, kl1 = ky - ly, lj0 = lx - jx, lj1 = ly - jy;
    double d1 = __dadd_rn(__dmul_rn(jk0, jk0), __dmul_rn(jk1, jk1));
    double d2 = __dadd_rn(__dmul_rn(kl0, kl0), __dmul_rn(kl1, kl1));
    double d3 = __dadd_rn(__dmul_rn(lj0, lj0), __dmul_rn(lj1, lj1));
    double cr = __dsub_rn(__dmul_rn(jk0, kl1), __dmul_rn(jk1, kl0));
    return ddiv(__dmul_rn(2.0, cr), dsqrt(__dmul_rn(__dmul_rn(d1, d2), d3)));
  }

  // The L-curve bookkeeping scans (point cache, state cache) are lane-parallel: lane <-> cache entry,
  // with warp reductions that reproduce the sequential "first match / first maximum" semantics.

  // first index i with isapprox(t, key_i), or INT_MAX
  __device__ __noinline__ int lc_find(double t, int npts) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    const double *pts = lc_pts_p;
    int found = 0x7fffffff;
    for (int i = lane; i < npts; i += 32)
      if (isapprox(t, pts[4 * i])) {
        found = i;
        break;
      }
    return (int)__reduce_min_sync(DECAES_FULL_MASK, (unsigned)found);
  }

  // cached evaluation of P(t) = (log ||Ax-b||^2, log ||x||^2); returns the point-cache index
  __device__ __noinline__ int lc_eval(double t, int &npts, const double *Asrc, int hint = 0) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    double *pts = lc_pts_p;
    int i = lc_find(t, npts);
    if (i != 0x7fffffff) return i;
    cache_solve(dexp(t), Asrc, t, hint);
    // the two logarithms (a software sequence each) side by side on the odd and the even lanes
    const double lv = dlog((lane & 1) ? cur_seminorm_sq() : cur_resnorm_sq());
    const double xi = __shfl_sync(DECAES_FULL_MASK, lv, 0), eta = __shfl_sync(DECAES_FULL_MASK, lv, 1);
    i = npts;
    if (npts < DECAES_LC_MAX) {
      if (lane == 0) pts[4 * i] = t, pts[4 * i + 1] = xi, pts[4 * i + 2] = eta, pts[4 * i + 3] = -CUDART_INF;
      npts++;
    } else {
      i = DECAES_LC_MAX - 1;
      n_overflow++;
    }
    __syncwarp();
    return i;
  }

#ifndef DECAES_LC_CURV_SEQ
  // The four points of the state are handled in PARALLEL: lanes 8 q .. 8 q + 7 own point q, scan the cached abscissae with a
  // stride of eight and reduce inside their group by shuffles; the Menger curvature (a division and a square root, both
  // software sequences) is then evaluated once by every lane for its own point instead of four times in a row by the whole
  // warp.  Same arithmetic per point as the sequential scan (first index wins on ties): bit-identical curvatures.
  __device__ __noinline__ void lc_update_curvature(const double *sx, const int *si, int npts, double tlx, double tly, double brx,
                                      double bry, double Ctol) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    double *pts = lc_pts_p;
    const int q = lane >> 3, sub = lane & 7;
    const int pi = si[q];
    const double x = sx[q], px = pts[4 * pi + 1], py = pts[4 * pi + 2];
    // nearest cached abscissae on either side of x (src/lsqnonneg.jl:954-959): first index on ties
    unsigned long long km = 0ull, kp = ~0ull;
    int im = 0x7fffffff, ip = 0x7fffffff;
    _Pragma("unroll 1") for (int k = sub; k < npts; k += 8) {
      const double _x = pts[4 * k];
      const unsigned long long kk = dkey(_x);
      if (_x < x && kk > km) km = kk, im = k;
      if (x < _x && kk < kp) kp = kk, ip = k;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned long long okm = __shfl_xor_sync(DECAES_FULL_MASK, km, o), okp = __shfl_xor_sync(DECAES_FULL_MASK, kp, o);
      const int oim = __shfl_xor_sync(DECAES_FULL_MASK, im, o), oip = __shfl_xor_sync(DECAES_FULL_MASK, ip, o);
      if (okm > km || (okm == km && oim < im)) km = okm, im = oim;
      if (okp < kp || (okp == kp && oip < ip)) kp = okp, ip = oip;
    }
    double C = -CUDART_INF;
    if (fmin(norm2(px, py, tlx, tly), norm2(px, py, brx, bry)) > Ctol) {
      double mx = px, my = py, qx = px, qy = py;
      if (km != 0ull) mx = pts[4 * im + 1], my = pts[4 * im + 2];
      if (kp != ~0ull) qx = pts[4 * ip + 1], qy = pts[4 * ip + 2];
      C = menger(mx, my, px, py, qx, qy);
    }
    __syncwarp();
    if (sub == 0) pts[4 * pi + 3] = C;
    __syncwarp();
  }
#else
  __device__ __noinline__ void lc_update_curvature(const double *sx, const int *si, int npts, double tlx, double tly, double brx,
                                      double bry, double Ctol) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    double *pts = lc_pts_p;
    _Pragma("unroll 1") for (int q = 0; q < 4; q++) {
      int pi = si[q];
      double x = sx[q], px = pts[4 * pi + 1], py = pts[4 * pi + 2];
      double C = -CUDART_INF;
      if (fmin(norm2(px, py, tlx, tly), norm2(px, py, brx, bry)) > Ctol) {
        // nearest cached abscissae on either side of x (src/lsqnonneg.jl:954-959): first index on ties
        unsigned long long km = 0ull, kp = ~0ull, best;
        int im = 0x7fffffff, ip = 0x7fffffff;
        _Pragma("unroll 1") for (int k = lane; k < npts; k += 32) {
          const double _x = pts[4 * k];
          const unsigned long long kk = dkey(_x);
          if (_x < x && kk > km) km = kk, im = k;
          if (x < _x && kk < kp) kp = kk, ip = k;
        }
        im = warp_argmax_bits(km, im, best);
        if (best == 0ull) im = 0x7fffffff;
        ip = warp_argmin_bits(kp, ip, best);
        if (best == ~0ull) ip = 0x7fffffff;
        double mx = px, my = py, qx = px, qy = py;
        if (im != 0x7fffffff) mx = pts[4 * im + 1], my = pts[4 * im + 2];
        if (ip != 0x7fffffff) qx = pts[4 * ip + 1], qy = pts[4 * ip + 2];
        C = menger(mx, my, px, py, qx, qy);
      }
      __syncwarp();
      if (lane == 0) pts[4 * pi + 3] = C;
      __syncwarp();
    }
  }

#endif

  // mapfindmax over the curvatures: first maximum under Base.isless (NaN is maximal)
  __device__ __noinline__ int lc_argmax(int npts) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    const double *pts = lc_pts_p;
    unsigned long long key = 0ull, best;
    int bi = 0x7fffffff;
    _Pragma("unroll 1") for (int i = lane; i < npts; i += 32) {
      const double c = pts[4 * i + 3];
      // isless order: -Inf < ... < -0 < +0 < ... < +Inf < NaN; +1 keeps "no candidate" (0) below -Inf... (dkey(-Inf) > 0 already)
      const unsigned long long kk = isnan(c) ? ~0ull : dkey(c);
      if (bi == 0x7fffffff || kk > key) key = kk, bi = i;
    }
    if (bi == 0x7fffffff) key = 0ull;
    return warp_argmax_bits(key, bi, best);
  }

  // backtracking (src/lsqnonneg.jl:892-900): among the stored states that have the arg-max point as
  // an interior point, the sequential scan ends on the LAST one of minimal width, provided that width
  // does not exceed the current state's.  Returns its index or -1.
  __device__ __noinline__ int lc_backtrack(double xb, double wcur, int nst) {
    const int lane = this->lane;
    double *const lc_states_p = this->lc_states_p;  // shared or global (PipeParams::spill)
    const double *sts = lc_states_p;
    unsigned long long key = ~0ull;  // widths are >= 0: their bit patterns order like the values
    int bk = -1;
    _Pragma("unroll 1") for (int k = lane; k < nst; k += 32) {
      const double *s = sts + 5 * k;
      if (s[1] == xb || s[2] == xb) {
        const unsigned long long kk = (unsigned long long)__double_as_longlong(fabs(s[3] - s[0]));
        if (kk <= key) key = kk, bk = k;
      }
    }
    const unsigned hi = __reduce_min_sync(DECAES_FULL_MASK, (unsigned)(key >> 32));
    const unsigned lo = __reduce_min_sync(DECAES_FULL_MASK, ((unsigned)(key >> 32) == hi) ? (unsigned)key : 0xffffffffu);
    const unsigned long long best = ((unsigned long long)hi << 32) | lo;
    bk = (int)__reduce_max_sync(DECAES_FULL_MASK, (key == best && bk >= 0) ? (unsigned)(bk + 1) : 0u) - 1;  // LAST state of minimal width
    return (bk >= 0 && __longlong_as_double((long long)best) <= wcur) ? bk : -1;
  }

  __device__ __noinline__ double lcurve_corner(const double *Asrc) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    double *const lc_states_p = this->lc_states_p;  // shared or global (PipeParams::spill)
    const double phi = 1.618033988749895, xtol = 1e-4, Ptol = 1e-4, Ctol = 1e-4;
    double *pts = lc_pts_p, *sts = lc_states_p;
    int npts = 0, nst = 0;
    double sx[4];
    int si[4];
    sx[0] = -8.0, sx[3] = 2.0;
    sx[1] = __dadd_rn(__dmul_rn(phi, sx[0]), sx[3]) / (phi + 1);
    sx[2] = sx[0] + (sx[3] - sx[1]);
    // first point (mu = e^-8: practically the unregularised problem) starts from the flip-angle fit's active set,
    // the last one (mu = e^2) from the full column set
    solve_vote = 4;
    for (int q = 0; q < 4; q++) si[q] = lc_eval(sx[q], npts, Asrc, q == 0 ? (cP.lc_hints & 2) : (q == 3 ? (cP.lc_hints & 1) : ((cP.lc_hints & 8) ? q + 2 : 0)));
    const double tlx = pts[4 * si[0] + 1], tly = pts[4 * si[0] + 2], brx = pts[4 * si[3] + 1], bry = pts[4 * si[3] + 2];
    lc_update_curvature(sx, si, npts, tlx, tly, brx, bry, Ctol);
    int iter = 0;
    solve_vote = 0;  // the steps below vote at the top of the loop (one vote per step, cache hit or not)
    while (true) {
      double p1x = pts[4 * si[0] + 1], p1y = pts[4 * si[0] + 2], p4x = pts[4 * si[3] + 1], p4y = pts[4 * si[3] + 2];
      if (fabs(sx[3] - sx[0]) < xtol || norm2(p1x, p1y, p4x, p4y) < Ptol) break;
      iter++;
      if (cP.step_sync & 2) cta_or(true);
      {  // backtracking  :892-900
        double xb = pts[4 * lc_argmax(npts)];
        int kb = lc_backtrack(xb, fabs(sx[3] - sx[0]), nst);
        if (kb >= 0) {
          const double *s = sts + 5 * kb;
          unsigned long long packed = (unsigned long long)__double_as_longlong(s[4]);
          for (int q = 0; q < 4; q++) sx[q] = s[q], si[q] = (int)((packed >> (16 * q)) & 0xffff);
        }
      }
      double C2 = pts[4 * si[1] + 3], C3 = pts[4 * si[2] + 3];
      if (C2 > C3) {  // move_left  :933-939
        double nx = __dadd_rn(__dmul_rn(phi, sx[0]), sx[2]) / (phi + 1);
        sx[3] = sx[2], si[3] = si[2];
        sx[2] = sx[1], si[2] = si[1];
        sx[1] = nx;
        si[1] = lc_eval(nx, npts, Asrc);
      } else {  // move_right  :941-946
        double nx = sx[1] + (sx[3] - sx[2]);
        sx[0] = sx[1], si[0] = si[1];
        sx[1] = sx[2], si[1] = si[2];
        sx[2] = nx;
        si[2] = lc_eval(nx, npts, Asrc);
      }
      lc_update_curvature(sx, si, npts, tlx, tly, brx, bry, Ctol);
      if (nst < DECAES_LC_MAX) {
        if (lane == 0) {
          double *s = sts + 5 * nst;
          unsigned long long packed = 0ull;
          for (int q = 0; q < 4; q++) s[q] = sx[q], packed |= ((unsigned long long)si[q]) << (16 * q);
          s[4] = __longlong_as_double((long long)packed);
        }
        nst++;
        __syncwarp();
      } else {
        n_overflow++;
        break;
      }
    }
    return pts[4 * lc_argmax(npts)];
  }

  // ---- Brent root / bracket (src/optimization.jl:71-128, 177-219) on f(log mu) ----
  // mode 0: chi2 relative error (src/lsqnonneg.jl:374-385); mode 1: res^2 - delta^2 (:723-726)
  __device__ double root_fun(double logmu, double target, int mode, const double *Asrc) {
    cache_solve(exp(logmu), Asrc, logmu);
    double r2 = cur_resnorm_sq();
    return mode == 0 ? (r2 - target) / target : r2 - target;
  }
  static __device__ __forceinline__ double sgn(double x) { return (double)((x > 0) - (x < 0)); }

  __device__ __noinline__ void bracket_and_brent(double target, int mode, double ftol, const double *Asrc, double &x_final,
                                    double &f_final) {
    // bracket_root_monotonic(f, -4, 1; dilate = 1.5, mono = +1, maxiters = 6)
    double a = -4.0, delta = 1.0, b, fa, fb;
    bool bracket_done = false;
    fa = root_fun(a, target, mode, Asrc);
    if (!isfinite(fa)) {
      b = a, fa = CUDART_NAN, fb = CUDART_NAN, bracket_done = true;
    } else if (fa == 0) {
      b = a, fb = fa, bracket_done = true;
    }
    if (!bracket_done) {
      double sd = sgn(fa);  // sign(mono) = +1
      b = a - sd * delta;
      fb = root_fun(b, target, mode, Asrc);
      if (!isfinite(fb)) {
        b = a, fb = fa, bracket_done = true;
      } else if (fb == 0) {
        a = b, fa = fb, bracket_done = true;
      }
      if (!bracket_done) {
        delta *= 1.5;
        int cnt = 0;
        while (fa * fb > 0 && cnt < 6) {
          a = b, fa = fb;
          b = a - sd * delta;
          fb = root_fun(b, target, mode, Asrc);
          if (!isfinite(fb)) {
            b = a, fb = fa, bracket_done = true;
            break;
          }
          if (fb == 0) {
            a = b, fa = fb, bracket_done = true;
            break;
          }
          delta *= 1.5;
          cnt++;
        }
        if (!bracket_done && !(a < b)) {
          double t = a;
          a = b, b = t, t = fa, fa = fb, fb = t;
        }
      }
    }
    if (fa * fb < 0) {
      // brent_root(f, a, b, fa, fb; xatol = 0, xrtol = 0, ftol, maxiters = 100)
      double x0 = a, x1 = b, fx0 = fa, fx1 = fb;
      double A_ = x0, B_ = x1, fA = fx0, fB = fx1;
      if (fabs(fA) < fabs(fB)) {
        double t = A_;
        A_ = B_, B_ = t, t = fA, fA = fB, fB = t;
      }
      double c = x0, d = x0, fc = fx0;
      bool mflag = true;
      x_final = B_, f_final = fB;
      bool done = false;
      for (int it = 1; it <= 100 && !done; it++) {
        if (fabs(B_ - A_) <= 0.0) break;
        double s = 0.0;
        s += A_ * fB * fc / (fA - fB) / (fA - fc);
        s += B_ * fA * fc / (fB - fA) / (fB - fc);
        s += c * fA * fB / (fc - fA) / (fc - fB);
        if (isnan(s) || isinf(s)) s = A_ - fA * (B_ - A_) / (fB - fA);
        double uu = (3 * A_ + B_) / 4, vv = B_;
        if (uu > vv) {
          double t = uu;
          uu = vv, vv = t;
        }
        double tol = 0.0;  // max(xatol, xrtol * ...) with both zero
        if (!(uu < s && s < vv) || (mflag && fabs(s - B_) >= fabs(B_ - c) / 2) ||
            (!mflag && fabs(s - B_) >= fabs(B_ - c) / 2) || (mflag && fabs(B_ - c) <= tol) ||
            (!mflag && fabs(c - d) <= tol)) {
          s = (A_ + B_) / 2;
          mflag = true;
        } else {
          mflag = false;
        }
        double fs = root_fun(s, target, mode, Asrc);
        if (fs == 0) {
          x_final = s, f_final = fs, done = true;
          break;
        }
        if (isnan(fs) || isinf(fs)) break;
        if (fabs(fs) <= ftol) {
          x_final = s, f_final = fs, done = true;
          break;
        }
        d = c, c = B_, fc = fB;
        if (sgn(fA) * sgn(fs) < 0)
          B_ = s, fB = fs;
        else
          A_ = s, fA = fs;
        if (fabs(fA) < fabs(fB)) {
          double t = A_;
          A_ = B_, B_ = t, t = fA, fA = fB, fB = t;
        }
        x_final = B_, f_final = fB;
      }
    } else {
      if (!isfinite(fa))
        x_final = b, f_final = fb;
      else if (!isfinite(fb))
        x_final = a, f_final = fa;
      else if (fabs(fa) < fabs(fb))
        x_final = a, f_final = fa;
      else
        x_final = b, f_final = fb;
    }
  }

  // ---- GCV  src/lsqnonneg.jl:1136-1205 ----
  // singular values of the voxel's basis by one-sided Jacobi on the smaller Gram-free side
  // (stands in for LAPACK dgesdd_, src/utils.jl:103-134).  lane <-> element of a column pair.
  __device__ __noinline__ void gcv_svdvals(const double *Asrc) {
    const int m = cP.nTE, n = cP.nT2, ld = cP.ld;
    double *Gm = g + sl.gcv_mat;  // tall r x c, column-major
    const int r = m >= n ? m : n, c = m >= n ? n : m;
    for (int k = lane; k < m * n; k += 32) {
      int i = k % m, j = k / m;
      double v = Asrc[i * ld + j];
      if (m >= n) Gm[i + j * r] = v; else Gm[j + i * r] = v;
    }
    __syncwarp();
    for (int sweep = 0; sweep < 60; sweep++) {
      bool rotated = false;
      for (int p = 0; p < c - 1; p++)
        for (int q = p + 1; q < c; q++) {
          double *gp = Gm + p * r, *gq = Gm + q * r;
          double al = 0, be = 0, ga = 0;
          for (int i = lane; i < r; i += 32) {
            double up = gp[i], uq = gq[i];
            al = fma(up, up, al), be = fma(uq, uq, be), ga = fma(up, uq, ga);
          }
          al = warp_sum(al), be = warp_sum(be), ga = warp_sum(ga);
          if (ga == 0.0 || fabs(ga) <= DBL_EPSILON * sqrt(al * be)) continue;
          rotated = true;
          double zeta = (be - al) / (2 * ga);
          double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1 + zeta * zeta));
          double cs = 1 / sqrt(1 + t * t), sn = cs * t;
          for (int i = lane; i < r; i += 32) {
            double up = gp[i], uq = gq[i];
            gp[i] = cs * up - sn * uq;
            gq[i] = sn * up + cs * uq;
          }
          __syncwarp();
        }
      if (!rotated) break;
    }
    double *gam = g + sl.gcv_gamma;
    for (int j = 0; j < c; j++) {
      double s = 0;
      for (int i = lane; i < r; i += 32) s = fma(Gm[j * r + i], Gm[j * r + i], s);
      s = warp_sum(s);
      if (lane == 0) gam[j] = sqrt(s);
    }
    __syncwarp();
  }
  // Singular values of the voxel's basis for Reg = gcv, in SHARED memory and with the column pairs of a one-sided
  // Jacobi sweep handled in PARALLEL: lane <-> pair of a round-robin round (c / 2 disjoint pairs, c - 1 rounds per
  // sweep), every lane runs its own dot products and its own rotation - no warp reductions, no barriers inside a
  // round.  (The first version kept the matrix in global scratch and swept the pairs one after the other with
  // lane <-> row: 56 M warp-cycles per voxel.)  Runs in the basis phase, when everything in front of the voxel's
  // signal is free: the r x c matrix (r = max(nTE, nT2) >= c) sits at the start of the warp's shared memory, row-major.
  __device__ __noinline__ void gcv_svdvals_smem(const double *Asrc, int rs, int cs, double *B) {  // element (i, j) at i * rs + j * cs
    SH(B);
    GL(Asrc);
    const int lane = this->lane;
    const int m = cP.nTE, n = cP.nT2;
    const int r = m >= n ? m : n, c = m >= n ? n : m;
    _Pragma("unroll 1") for (int k = lane; k < m * n; k += 32) {
      const int i = k / n, j = k - i * n;
      const double v = Asrc[i * rs + j * cs];
      if (m >= n) B[i * c + j] = v;
      else B[j * c + i] = v;
    }
    __syncwarp();
    const int ce = (c + 1) & ~1, npair = ce >> 1;  // odd c: one dummy column (index c), its pair sits the round out
    _Pragma("unroll 1") for (int sweep = 0; sweep < 60; sweep++) {
      bool rotated = false;
      _Pragma("unroll 1") for (int rd = 0; rd < ce - 1; rd++) {
        // round-robin tournament: column ce - 1 stays, the others rotate
        int p = ce - 1, q = rd;
        if (lane > 0) p = (rd + lane) % (ce - 1), q = (rd + (ce - 1) - lane) % (ce - 1);
        if (lane < npair && p < c && q < c) {
          if (p > q) {
            const int t = p;
            p = q, q = t;
          }
          double al = 0.0, be = 0.0, ga = 0.0;
          _Pragma("unroll 4") for (int i = 0; i < r; i++) {
            const double up = B[i * c + p], uq = B[i * c + q];
            al = fma(up, up, al), be = fma(uq, uq, be), ga = fma(up, uq, ga);
          }
          if (!(ga == 0.0 || fabs(ga) <= DBL_EPSILON * sqrt(al * be))) {
            rotated = true;
            const double zeta = (be - al) / (2 * ga);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1 + zeta * zeta));
            const double cs = 1 / sqrt(1 + t * t), sn = cs * t;
            _Pragma("unroll 4") for (int i = 0; i < r; i++) {
              const double up = B[i * c + p], uq = B[i * c + q];
              B[i * c + p] = cs * up - sn * uq;
              B[i * c + q] = sn * up + cs * uq;
            }
          }
        }
        __syncwarp();
      }
      if (!__any_sync(DECAES_FULL_MASK, rotated)) break;
    }
    double *gam = g + sl.gcv_gamma;
    GL(gam);
    _Pragma("unroll 1") for (int j = lane; j < c; j += 32) {
      double s2 = 0.0;
      _Pragma("unroll 4") for (int i = 0; i < r; i++) s2 = fma(B[i * c + j], B[i * c + j], s2);
      gam[j] = sqrt(s2);
    }
    __threadfence_block();
    __syncwarp();
  }

  // Singular values for Reg = gcv, third version (PipeParams::gcv_smem == 2): Golub-Kahan Householder bidiagonalisation of the
  // tall r x c copy in shared memory (4 r c^2 / 3 flops instead of ~10 Jacobi sweeps of 6 r c^2 each), then multisection on the
  // Sturm counts of the Golub-Kahan tridiagonal form of the bidiagonal (order 2c, zero diagonal, off-diagonals d1 e1 d2 e2 ...:
  // its eigenvalues are -sigma_c .. -sigma_1, sigma_1 .. sigma_c).  This is the route LAPACK takes as well (dgebrd + a bidiagonal
  // solver; the reference's dgesdd_, src/utils.jl:103-134): backward stable, every sigma to a few eps * sigma_max.
  //  * left reflector k (column k, rows k..r-1): lane <-> column j > k, the reflector is read as a broadcast;
  //  * right reflector k (row k, columns k+1..c-1): lane <-> row i > k; the leading dimension c | 1 is odd, so both
  //    access directions are free of bank conflicts; two columns / rows per lane advance together;
  //  * multisection: lane <-> singular values lane and lane + 32, all lanes take the same 34 trisections of [0, ||B||_F]
  //    (interval 6e-17 ||B||_F), two Sturm pivots per reciprocal, four independent chains per lane.
  // One lane's share of a Householder update x -= g (v'x) v for one (TWO: two) vectors x: element t of x at o[t * os], of v at
  // vec[t * vs] for t in [lo, hi), the leading element (index head) of v is v0.  Chunks of four elements are loaded before they
  // are stored (the compiler cannot see that v and x never overlap).  TWO is warp-uniform (does ANY lane have a second vector:
  // when none has, its loads would only cost shared-memory wavefronts), `two` is the lane's own answer.
  template <bool TWO>
  static __device__ __noinline__ void reflect_apply(double *o0, double *o1, bool two, const double *vec, int os, int vs, int head, int lo,
                                                    int hi, double v0, double g) {
    SH(o0);
    SH(o1);
    SH(vec);
    double w0 = v0 * o0[head * os], w1 = TWO ? v0 * o1[head * os] : 0.0;
    int t = lo;
    _Pragma("unroll 1") for (; t + 4 <= hi; t += 4) {
      double vv[4], a0[4], a1[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        vv[u] = vec[(t + u) * vs], a0[u] = o0[(t + u) * os];
        if (TWO) a1[u] = o1[(t + u) * os];
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        w0 = fma(vv[u], a0[u], w0);
        if (TWO) w1 = fma(vv[u], a1[u], w1);
      }
    }
    _Pragma("unroll 1") for (; t < hi; t++) {
      w0 = fma(vec[t * vs], o0[t * os], w0);
      if (TWO) w1 = fma(vec[t * vs], o1[t * os], w1);
    }
    w0 *= g, w1 *= g;
    o0[head * os] = fma(-v0, w0, o0[head * os]);
    if (TWO && two) o1[head * os] = fma(-v0, w1, o1[head * os]);
    t = lo;
    _Pragma("unroll 1") for (; t + 4 <= hi; t += 4) {
      double vv[4], a0[4], a1[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        vv[u] = vec[(t + u) * vs], a0[u] = o0[(t + u) * os];
        if (TWO) a1[u] = o1[(t + u) * os];
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        o0[(t + u) * os] = fma(-vv[u], w0, a0[u]);
        if (TWO && two) o1[(t + u) * os] = fma(-vv[u], w1, a1[u]);
      }
    }
    _Pragma("unroll 1") for (; t < hi; t++) {
      const double vt = vec[t * vs];
      o0[t * os] = fma(-vt, w0, o0[t * os]);
      if (TWO && two) o1[t * os] = fma(-vt, w1, o1[t * os]);
    }
  }
  // 1 / a for finite, normal |a| in [1e-150, 1e150] (the clamped Sturm pivots): the hardware seed (MUFU.RCP64H, ~20 bits) and
  // one cubic Newton step, 2^-60 relative - no range checks and no slow-path branch, so independent chains interleave
  static __device__ __forceinline__ double rcp_nr(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    e = fma(e, e, e);
    return fma(x, e, x);
  }
  __device__ __noinline__ void gcv_svdvals_bidiag(const double *Asrc, int rs, int cs, double *B) {  // element (i, j) at i * rs + j * cs
    SH(B);
    GL(Asrc);
    const int lane = this->lane;
    const int m = cP.nTE, n = cP.nT2;
    const int R = m >= n ? m : n, C = m >= n ? n : m, ld = C | 1;
    {
      // copy in the order the source is laid out (inner extent w: column-major m, row-major n), four loads in flight;
      // the (outer, inner) position advances by 32 without a division
      const bool cmaj = rs == 1;
      const int w = cmaj ? m : n, tot = m * n;
      int in = lane % w, out = lane / w;
      _Pragma("unroll 1") for (int k0 = lane; k0 < tot; k0 += 128) {
        double v[4];
        int dst[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const bool ok = k0 + 32 * u < tot;
          const int i = cmaj ? in : out, j = cmaj ? out : in;
          v[u] = ok ? Asrc[i * rs + j * cs] : 0.0;
          dst[u] = ok ? (m >= n ? i * ld + j : j * ld + i) : -1;
          in += 32;
          while (in >= w) in -= w, out++;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (dst[u] >= 0) B[dst[u]] = v[u];
      }
    }
    __syncwarp();
    // Reflectors are kept unnormalised: v = x - beta e_1 with beta = -sign(x_0) ||x||, H = I - g v v', g = 1 / (||x|| (||x|| + |x_0|));
    // v_0 lives in a register, the rest of v stays where x was (reflect_apply).
    _Pragma("unroll 1") for (int k = 0; k < C; k++) {
      {  // left reflector: annihilate B[k+1:R, k]; lane <-> columns k + 1 + lane (+ 32)
        double *const colk = B + k;
        double acc = 0.0;
        _Pragma("unroll 1") for (int i = k + 1 + lane; i < R; i += 32) acc = fma(colk[i * ld], colk[i * ld], acc);
        const double xn2 = warp_sum(acc), x0 = colk[k * ld];
        if (xn2 != 0.0) {  // (warp-uniform: the butterfly sum is bitwise identical on every lane)
          const double nrm = sqrt(fma(x0, x0, xn2));
          const double v0 = x0 + copysign(nrm, x0), g = 1.0 / (nrm * (nrm + fabs(x0)));
          __syncwarp();
          if (lane == 0) colk[k * ld] = -copysign(nrm, x0);
          const int j0 = k + 1 + lane;
          if (j0 < C) {
            const bool two = j0 + 32 < C;
            if (C - k - 1 > 32) reflect_apply<true>(B + j0, B + (two ? j0 + 32 : j0), two, colk, ld, ld, k, k + 1, R, v0, g);
            else reflect_apply<false>(B + j0, B + j0, false, colk, ld, ld, k, k + 1, R, v0, g);
          }
          __syncwarp();
        }
      }
      if (k < C - 2) {  // right reflector: annihilate B[k, k+2:C]; lane <-> rows k + 1 + lane (+ 32)
        double *const rowk = B + k * ld;
        double acc = 0.0;
        _Pragma("unroll 1") for (int j = k + 2 + lane; j < C; j += 32) acc = fma(rowk[j], rowk[j], acc);
        const double xn2 = warp_sum(acc), x0 = rowk[k + 1];
        if (xn2 != 0.0) {
          const double nrm = sqrt(fma(x0, x0, xn2));
          const double u0 = x0 + copysign(nrm, x0), g = 1.0 / (nrm * (nrm + fabs(x0)));
          __syncwarp();
          if (lane == 0) rowk[k + 1] = -copysign(nrm, x0);
          _Pragma("unroll 1") for (int i0 = k + 1 + lane; i0 < R; i0 += 64) {
            const bool two = i0 + 32 < R;
            if (R - k - 1 > 32) reflect_apply<true>(B + i0 * ld, B + (two ? i0 + 32 : i0) * ld, two, rowk, 1, 1, k + 1, k + 2, C, u0, g);
            else reflect_apply<false>(B + i0 * ld, B + i0 * ld, false, rowk, 1, 1, k + 1, k + 2, C, u0, g);
          }
          __syncwarp();
        }
      }
    }
    // squares of the off-diagonals of the Golub-Kahan form, b = (d_0, e_0, d_1, ..., d_{C-1}), at the start of B
    const int nb = 2 * C - 1;
    double bq[4], fro = 0.0, mx = 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int t = lane + 32 * u, kk = t >> 1;
      double bt = 0.0;
      if (t < nb) bt = B[kk * ld + kk + (t & 1)];
      bq[u] = __dmul_rn(bt, bt);
      fro += bq[u], mx = fmax(mx, bq[u]);
    }
    fro = warp_sum(fro), mx = warp_max(mx);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (lane + 32 * u < nb) B[lane + 32 * u] = bq[u];
    __syncwarp();
    // pivots below pivmin in magnitude are replaced by -pivmin (LAPACK dlaebz's safeguard; an absolute perturbation of 1e-150)
    const double bound = sqrt(fro) * 1.0000001, pivmin = 1e-150 * fmax(1.0, mx);
    // lane <-> singular values lane and lane + 32; every step cuts both intervals in THREE (Sturm counts at the two interior
    // points: four independent chains per lane, the recursion is bound by the latency of its reciprocal): 34 steps shrink
    // [0, ||B||_F] by 3^34 = 1.7e16
    double lo[2] = {0.0, 0.0}, hi[2] = {bound, bound};
    const int tg[2] = {C + lane, C + lane + 32};  // sigma_j (ascending, 0-based) < x  <=>  #(eigenvalues < x) > C + j
    _Pragma("unroll 1") for (int it = 0; it < 34; it++) {
      double x[4], q[4];
      int c[4];
#pragma unroll
      for (int v = 0; v < 2; v++) {
        const double third = (hi[v] - lo[v]) * (1.0 / 3.0);
        x[2 * v] = lo[v] + third, x[2 * v + 1] = hi[v] - third;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) q[u] = (x[u] < pivmin) ? -pivmin : -x[u], c[u] = 1;  // first pivot: negative (x > 0)
      // two pivots per reciprocal: with num = q_{i-1} q_i = -x q_{i-1} - b_i,  q_{i+1} = -x - b_{i+1} q_{i-1} / num and
      // sign(q_i) = sign(num) sign(q_{i-1}); nb is odd, so the last entry takes a single step
      _Pragma("unroll 1") for (int i = 0; i + 1 < nb; i += 2) {
        const double ba = B[i], bb = B[i + 1];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          double nu = fma(-x[u], q[u], -ba);
          if (fabs(nu) < pivmin * fabs(q[u])) nu = -pivmin * q[u];  // q_i := -pivmin
          c[u] += (nu < 0.0) != (q[u] < 0.0);
          double qn = fma(-__dmul_rn(bb, q[u]), rcp_nr(nu), -x[u]);
          if (fabs(qn) < pivmin) qn = -pivmin;
          c[u] += qn < 0.0;
          q[u] = qn;
        }
      }
      const double bl = B[nb - 1];
#pragma unroll
      for (int u = 0; u < 4; u++) c[u] += fma(-bl, rcp_nr(q[u]), -x[u]) < 0.0;
#pragma unroll
      for (int v = 0; v < 2; v++) {
        if (c[2 * v] > tg[v]) hi[v] = x[2 * v];
        else if (c[2 * v + 1] > tg[v]) lo[v] = x[2 * v], hi[v] = x[2 * v + 1];
        else lo[v] = x[2 * v + 1];
      }
    }
    double *gam = g + sl.gcv_gamma;
    GL(gam);
    if (lane < C) gam[lane] = (lo[0] + hi[0]) / 2;
    if (lane + 32 < C) gam[lane + 32] = (lo[1] + hi[1]) / 2;
    __threadfence_block();
    __syncwarp();
  }
  __device__ __forceinline__ void gcv_svd_shared(const double *Asrc, int rs, int cs, double *B) {
    if (cP.gcv_smem == 2) gcv_svdvals_bidiag(Asrc, rs, cs, B);
    else gcv_svdvals_smem(Asrc, rs, cs, B);
  }

  __device__ __noinline__ double gcv_fun(double logmu, const double *Asrc) {  // log(max(gcv, eps^2/m))  :1150-1154, 1213-1229
    const int m = cP.nTE, n = cP.nT2;
    double mu = exp(logmu);
    cache_solve(mu, Asrc, logmu);
    double r2 = cur_resnorm_sq();
    double dof = (double)((m - n) > 0 ? (m - n) : 0);
    double l2 = __dmul_rn(mu, mu);
    const double *gam = g + sl.gcv_gamma;
    GL(gam);
    int mn = m < n ? m : n;
    // sum_i mu^2 / (gamma_i^2 + mu^2): one division per lane and pass instead of min(m, n) in a row on every lane (the sum is
    // taken in butterfly order: ~1 ulp from the sequential one)
    double part = 0.0;
    _Pragma("unroll 1") for (int i = lane; i < mn; i += 32) {
      double g2 = __dmul_rn(gam[i], gam[i]);
      part += l2 / (g2 + l2);
    }
    dof += warp_sum(part);
    double gcv = r2 / __dmul_rn(dof, dof);
    gcv = fmax(gcv, (DBL_EPSILON * DBL_EPSILON) / m);
    return log(gcv);
  }
  __device__ __noinline__ double gcv_minimize(const double *Asrc) {  // brent_minimize  src/optimization.jl:319-413
    const double phi = 1.618033988749895, alpha = 2 - phi, xatol = 1e-4;
    double x1 = -8.0, x2 = 2.0;
    double x = x1 + alpha * (x2 - x1);
    double y = gcv_fun(x, Asrc);
    double dx_old = 0.0, dx = 0.0, x_older = x, x_old = x, y_older = y, y_old = y;
    int iter = 0;
    while (iter < 20) {
      double p = 0.0, q = 0.0, xm = (x2 + x1) / 2;
      double dx_tol = xatol;  // + xrtol * |x| with xrtol = 0
      if (fabs(x - xm) + (x2 - x1) / 2 <= 2 * dx_tol) break;
      iter++;
      if (fabs(dx_old) > dx_tol) {
        double r = __dmul_rn(x - x_old, y - y_older);
        q = __dmul_rn(x - x_older, y - y_old);
        p = __dsub_rn(__dmul_rn(x - x_older, q), __dmul_rn(x - x_old, r));
        q = 2 * (q - r);
        if (q > 0) p = -p; else q = -q;
      }
      if (fabs(p) < fabs(q * dx_old / 2) && p < q * (x2 - x) && p < q * (x - x1)) {
        dx_old = dx;
        dx = p / q;
        double x_tmp = x + dx;
        if ((x_tmp - x1) < 2 * dx_tol || (x2 - x_tmp) < 2 * dx_tol) dx = (x < xm) ? dx_tol : -dx_tol;
      } else {
        dx_old = (x < xm) ? x2 - x : x1 - x;
        dx = __dmul_rn(alpha, dx_old);
      }
      double x_new = (fabs(dx) >= dx_tol) ? x + dx : x + ((dx > 0) ? dx_tol : -dx_tol);
      double y_new = gcv_fun(x_new, Asrc);
      if (y_new < y) {
        if (x_new < x) x2 = x; else x1 = x;
        x_older = x_old, x_old = x, x = x_new;
        y_older = y_old, y_old = y, y = y_new;
      } else {
        if (x_new < x) x1 = x_new; else x2 = x_new;
        if (y_new <= y_old || x_old == x) {
          x_older = x_old, x_old = x_new, y_older = y_old, y_old = y_new;
        } else if (y_new <= y_older || x_older == x || x_older == x_old) {
          x_older = x_new, y_older = y_new;
        }
      }
    }
    return x;
  }

  // =====================================================================================
  // Gram-form solver path (gram.cuh).  Per voxel the warp keeps G = A'A and c = A'b in shared
  // memory; A itself stays in the warp's global scratch (L2) in both layouts and is only read for
  // explicit residuals (||Ax - b||^2, iterative refinement, fitted curve).
  // =====================================================================================
  // out = A' vec  (lane <-> column, coalesced rows of the row-major matrix); gram_rhs: c = A' bd
  __device__ __forceinline__ void gram_rhs(const double *Arm) { gram_atv(Arm, this->bd, this->cvec); }
  __device__ __noinline__ void gram_atv(const double *Arm, const double *vec, double *outv) {
    GL(Arm);
    const int lane = this->lane;
    const double *const bd = vec;
    double *const cvec = outv;
    SH(bd);
    SH(cvec);
    const int nTE = cP.nTE, ld = cP.ld, n = cP.nT2;
    // lane <-> columns lane and lane + 32: both advance together, 16 loads in flight (the matrix lives in L2).  Deeper
    // chunks (32 loads in flight) buy 0.3 % on cfg3 and cost 3 % on the nT2 = 60 configs (code size), so: unroll 8.
    const double *c0 = Arm + (lane < n ? lane : 0), *c1 = Arm + (lane + 32 < n ? lane + 32 : 0);
    double a0 = 0.0, a1 = 0.0;
    _Pragma("unroll 8") for (int i = 0; i < nTE; i++) {
      const double bi = bd[i];
      a0 = fma(c0[i * ld], bi, a0), a1 = fma(c1[i * ld], bi, a1);
    }
    if (lane < n) cvec[lane] = a0;
    if (lane + 32 < n) cvec[lane + 32] = a1;
    __syncwarp();
  }

  // explicit residual r = bd - A_P s (stored in `fit`), returns ||r||^2.  lane <-> echo.
  __device__ __noinline__ double gram_residual(const double *Acm, int k) {
    GL(Acm);
    const int lane = this->lane;
    VIEW(double, bd);
    VIEW(double, fit);
    VIEW_GWS();
    const int nTE = cP.nTE;
    // lane <-> echoes lane, lane + 32 (, lane + 64): all of them advance together so that one L2 round
    // trip serves up to 12 loads per lane (A lives in L2; the loop is latency bound)
    const int i0 = lane < nTE ? lane : 0, i1 = lane + 32 < nTE ? lane + 32 : i0, i2 = lane + 64 < nTE ? lane + 64 : i0;
    double r0 = bd[i0], r1 = bd[i1], r2 = bd[i2];
    // columns in chunks of DECAES_RESID_CHUNK, every load of a chunk issued before its first fma (A lives in L2 and the loop
    // is bound by its latency); slots beyond k repeat the chunk's first column with a zero coefficient: no branches, same sums
    constexpr int RC = DECAES_RESID_CHUNK;
    _Pragma("unroll 1") for (int tb = 0; tb < k; tb += RC) {
      double v0[RC], v1[RC];
#pragma unroll
      for (int u = 0; u < RC; u++) {
        const double *col = Acm + gws.P[tb + u < k ? tb + u : tb] * nTE;
        v0[u] = col[i0], v1[u] = col[i1];
      }
#pragma unroll
      for (int u = 0; u < RC; u++) {
        const double st = tb + u < k ? gws.s[tb + u < k ? tb + u : tb] : 0.0;
        r0 = fma(-v0[u], st, r0), r1 = fma(-v1[u], st, r1);
      }
    }
    if (nTE > 64) {
      _Pragma("unroll 2") for (int t = 0; t < k; t++) r2 = fma(-Acm[gws.P[t] * nTE + i2], gws.s[t], r2);
    }
    double acc = 0.0;
    if (lane < nTE) fit[i0] = r0, acc = __dmul_rn(r0, r0);
    if (lane + 32 < nTE) fit[i1] = r1, acc = fma(r1, r1, acc);
    if (lane + 64 < nTE) fit[i2] = r2, acc = fma(r2, r2, acc);
    acc = warp_sum(acc);
    __syncwarp();
    return acc;
  }

  // one step of iterative refinement on the active set: s += (G_PP + mu2 I)^-1 (A_P' r - mu2 s),
  // with r = bd - A_P s already in `fit`.  Restores QR-level accuracy of the normal-equation solve.
  __device__ __noinline__ void gram_refine(const double *Acm, int k, double mu2) {
    GL(Acm);
    const int lane = this->lane;
    VIEW(double, fit);
    VIEW(double, Gs);
    VIEW_GWS();
    const int nTE = cP.nTE;
    double *T = Gs;
    const int ld = cP.ldg;
    // g_t = A[:,P[t]]' r - mu2 s_t: lane <-> echo (coalesced L2 reads), four columns per L2 round trip, sums through the
    // shared out-of-line butterfly (the kernel pays for code size: 463 -> ~300 instructions, same arithmetic)
    const int i0 = lane < nTE ? lane : 0, i1 = lane + 32 < nTE ? lane + 32 : i0, i2 = lane + 64 < nTE ? lane + 64 : i0;
    const double r0 = lane < nTE ? fit[i0] : 0.0, r1 = lane + 32 < nTE ? fit[i1] : 0.0, r2 = lane + 64 < nTE ? fit[i2] : 0.0;
    _Pragma("unroll 1") for (int tb = 0; tb < k; tb += 4) {
      double v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double *col = Acm + gws.P[tb + u < k ? tb + u : tb] * nTE;
        v0[u] = col[i0], v1[u] = col[i1];
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        double a = fma(v1[u], r1, __dmul_rn(v0[u], r0));
        if (nTE > 64) a = fma(Acm[gws.P[tb + u < k ? tb + u : tb] * nTE + i2], r2, a);
        a = warp_sum(a);
        if (lane == 0 && tb + u < k) gws.t1[tb + u] = fma(-mu2, gws.s[tb + u], a);
      }
    }
    __syncwarp();
    // v = M g
    _Pragma("unroll 1") for (int t = lane; t < k; t += 32) {
      double a = 0.0;
      _Pragma("unroll 4") for (int u = 0; u <= t; u++) a = fma(GM_(t, u), gws.t1[u], a);
      gws.t2[t] = a;
    }
    __syncwarp();
    // s += M' v
    _Pragma("unroll 1") for (int u = lane; u < k; u += 32) {
      double a = 0.0;
      _Pragma("unroll 4") for (int t = u; t < k; t++) a = fma(GM_(t, u), gws.t2[t], a);
      double sn = gws.s[u] + a;
      gws.s[u] = sn;
      gws.x[gws.P[u]] = sn;
    }
    __syncwarp();
  }

  // Explicit duals of the columns whose normal-equation dual (left in gws.w by gram_nnls) is not clearly negative:
  // w_j = A_j' r with r = b - A_P s in `fit`.  Returns the column with the largest positive explicit dual (first on
  // ties, like largest_positive_dual, src/NNLS.jl:541-554) or -1.  "Clearly negative" = below -kkt_tau max|c| (kkt_tau = 1e-8; the
  // maximum is taken on the upper halves of the doubles, one REDUX: the threshold is a screening heuristic, two
  // orders of magnitude above the noise the normal-equation dual was seen to carry).  lane <-> echo; two candidate
  // columns per L2 round trip.  Written for size: almost every call finds zero to two candidates.
  __device__ __noinline__ int kkt_pick(const double *Acm, unsigned long long excl) {
    GL(Acm);
    const int lane = this->lane;
    VIEW(double, fit);
    VIEW(double, cvec);
    VIEW_GWS();
    const int nTE = cP.nTE, n = cP.nT2;
    const int j0 = lane < n ? lane : 0, j1 = lane + 32 < n ? lane + 32 : 0;
    const double cm = fmax(lane < n ? fabs(cvec[j0]) : 0.0, lane + 32 < n ? fabs(cvec[j1]) : 0.0);
    const unsigned hi = __reduce_max_sync(DECAES_FULL_MASK, (unsigned)__double2hiint(cm));
    const double tau = -cP.kkt_tau * __hiloint2double((int)hi, 0);
    const unsigned c0 = __ballot_sync(DECAES_FULL_MASK, lane < n && !((excl >> j0) & 1ull) && gws.w[j0] > tau);
    const unsigned c1 = __ballot_sync(DECAES_FULL_MASK, lane + 32 < n && !((excl >> j1) & 1ull) && gws.w[j1] > tau);
    unsigned long long cand = ((unsigned long long)c1 << 32) | c0;
    GP_ADD(10, __popcll(cand));
    const int i0 = lane < nTE ? lane : 0, i1 = lane + 32 < nTE ? lane + 32 : i0, i2 = lane + 64 < nTE ? lane + 64 : i0;
    const double r0 = lane < nTE ? fit[i0] : 0.0, r1 = lane + 32 < nTE ? fit[i1] : 0.0, r2 = lane + 64 < nTE ? fit[i2] : 0.0;
    double best = 0.0;
    int bj = -1;
    _Pragma("unroll 1") while (cand) {
      const int ja = __ffsll((long long)cand) - 1;
      cand &= cand - 1;
      const int jb = cand ? __ffsll((long long)cand) - 1 : ja;
      cand &= cand - 1;  // (0 & -1 = 0)
      const double *ca = Acm + ja * nTE, *cb = Acm + jb * nTE;
      double a = fma(ca[i0], r0, fma(ca[i1], r1, __dmul_rn(ca[i2], r2)));
      double b = fma(cb[i0], r0, fma(cb[i1], r1, __dmul_rn(cb[i2], r2)));
      a = warp_sum(a), b = warp_sum(b);
      if (a > best) best = a, bj = ja;
      if (b > best) best = b, bj = jb;
    }
    return bj;
  }

  // unregularised NNLS following the reference's cold-start path, polished by one refinement step;
  // returns ||A x - b||^2 (explicit) and leaves r in `fit`, x in gws.x, the active set in gws.cP.
  // `warm_mask` != 0: start from that active set instead (flip-angle probes only: the loss and its
  // gradient depend on the minimiser, which is unique, not on the pivoting path).
  __device__ __noinline__ double gram_solve_unreg(const Src &src, GramOut &o, unsigned long long warm_mask = 0ull, bool refine = true,
                                                  bool polish = true) {
    const int lane = this->lane;
    VIEW(double, V);
    VIEW_GWS();
    const int max_set = cP.nTE < cP.nT2 ? cP.nTE : cP.nT2;
    if (warm_mask) {
      _Pragma("unroll 1") for (int j = lane; j < cP.nT2; j += 32) gws.x[j] = ((warm_mask >> j) & 1ull) ? 1.0 : 0.0;
      __syncwarp();
    }
    PROF_BEGIN(1);
    o = gram_nnls<VS>(V, cP.nT2, cP.ldg, 0.0, max_set, warm_mask != 0ull, warm_mask);
    n_itercap += o.capped;
#ifdef DECAES_PROFILE
    if (lane == 0) {
      const int si = 28 + (nunreg_voxel < 19 ? nunreg_voxel : 19);
      atomicAdd(&g_solve_hist[0][si], 1ull), atomicAdd(&g_solve_hist[1][si], (unsigned long long)o.nappend), atomicAdd(&g_solve_hist[2][si], (unsigned long long)o.iters);
    }
    nunreg_voxel++;
#endif
    PROF_END(1);
    PROF_BEGIN(2);
    double r2 = gram_residual(src.Acm, o.k);
    PROF_END(2);
    if (o.k > 0 && refine) {
      PROF_BEGIN(3);
      gram_refine(src.Acm, o.k, 0.0);
      PROF_END(3);
      PROF_BEGIN(2);
      r2 = gram_residual(src.Acm, o.k);
      PROF_END(2);
    }
    if (!polish) return r2;
    // KKT polish.  The active-set decisions above were taken on normal-equation quantities: the dual w = c - G_P s
    // carries ~cond(A_P)^2 eps of noise (1e-10 on long-T2 pools), so a column whose true dual is a small positive
    // number can be left out (or a coefficient that should be clamped kept in) - a slightly worse stationary point
    // where the reference's QR iteration, whose duals are good to ~1e-15, goes on.  With the refined solution and
    // its explicit residual in hand, the duals that are not clearly negative are recomputed EXPLICITLY (w_j = A_j'r)
    // and the iteration resumes from here until the reference's own termination test - no positive dual, all
    // coefficients positive - holds at that accuracy.  Almost every solve passes at once (kkt_pick: one L2 round trip).
    unsigned long long tried = 0ull;
    _Pragma("unroll 1") for (int round = 0; round < 6; round++) {
      PROF_BEGIN(0);
      const int bj = (o.k < max_set) ? kkt_pick(src.Acm, o.mask | tried) : -1;
      unsigned long long neg = 0ull;
      _Pragma("unroll 1") for (int t = lane; t < o.k; t += 32)
        if (!(gws.s[t] > 0.0)) neg |= 1ull << gws.P[t];
      neg = warp_or64(neg);
      PROF_END(0);
      if (bj < 0 && neg == 0ull) break;
      unsigned long long m2 = o.mask & ~neg;
      if (bj >= 0) m2 |= 1ull << bj, tried |= 1ull << bj;
      _Pragma("unroll 1") for (int j = lane; j < cP.nT2; j += 32) {
        const double xj = gws.x[j];
        gws.x[j] = (((m2 >> j) & 1ull) && xj > 0.0) ? xj : 0.0;
      }
      __syncwarp();
      o = gram_nnls<VS>(V, cP.nT2, cP.ldg, 0.0, max_set, m2 != 0ull, m2);
      n_itercap += o.capped;
      r2 = gram_residual(src.Acm, o.k);
      if (o.k > 0) {
        gram_refine(src.Acm, o.k, 0.0);
        r2 = gram_residual(src.Acm, o.k);
      }
      n_polish++;
    }
    return r2;
  }

  // loss_with_grad!  src/splines.jl:1010-1041 on grid angle k
  __device__ __noinline__ void fa_eval_gram(int kang, double &u, double &du, unsigned long long seen, bool rough = false) {
    const int lane = this->lane;
    VIEW(double, fit);
    VIEW(double, Gs);
    VIEW(unsigned long long, fa_mask_p);
    VIEW_GWS();
    const int nTE = cP.nTE, n = cP.nT2;
    Src src;
    src.G = cP.gram_set + (size_t)kang * cP.a_elems, src.ldg = cP.ldg;
    src.Arm = cP.basis_rm + (size_t)kang * cP.copy_elems;
    src.Acm = cP.basis_cm + (size_t)kang * nTE * n;
    PROF_BEGIN(5);
    stage_bulk(Gs, src.G, (unsigned)(cP.a_elems * 8));  // TMA: G_k (lower triangle valid) -> shared memory (overlapping it with the
    PROF_END(5);                                         // right-hand side below buys nothing: profiles/r02_s4_ab_tma_overlap_vote_late_atv16.txt)
    PROF_BEGIN(0);
    gram_rhs(src.Arm);
    PROF_END(0);
    // warm start from the active set found at the nearest angle already probed, if it is close enough:
    // from a distant angle the inherited set is mostly wrong and costs more removals than the reference's
    // cold start needs pivots (measured: 17 inner iterations from 63 grid steps away, 7.5 cold, 4.4 from 1 step)
    unsigned long long warm = 0ull;
    if (seen && cP.fa_warm) {
      unsigned long long below = seen & ((1ull << kang) - 1ull), above = seen >> kang;  // (bit kang set: a second, precise probe of a seed - its own set)
      int jb = below ? 63 - __clzll((long long)below) : -1000;
      int ja = above ? kang + __ffsll((long long)above) - 1 : 1000;
      int jn = (kang - jb <= ja - kang) ? jb : ja;
      if (abs(kang - jn) <= cP.fa_warm) warm = fa_mask_p[jn];
    }
    GramOut o;
    // fa_polish = 0 leaves the KKT polish to the solves whose x is an output (+1.2 % throughput; one voxel in 2,048 of the
    // three-pool stress family then misses the flip-angle tolerance: 1.4e-6 instead of 3.4e-7)
    u = gram_solve_unreg(src, o, warm, cP.fa_refine != 0 && !rough,
                         !rough && (cP.fa_polish == 1 || (cP.fa_polish == 2 && __popcll(seen) >= cP.nseed)));
    if (lane == 0) fa_mask_p[kang] = o.mask;
    const double *dAk = cP.dbasis_cm + (size_t)kang * nTE * n;
    GL(dAk);
    double acc = 0.0;
    PROF_BEGIN(4);
    _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) {
      double dax = 0.0;
      _Pragma("unroll 4") for (int t = 0; t < o.k; t++) {
        const double st = gws.s[t], dv = dAk[gws.P[t] * nTE + i];
        if (st > 0.0) dax = fma(st, dv, dax);
      }
      acc = fma(dax, -fit[i], acc);  // A x - b = -r
    }
    du = 2.0 * warp_sum(acc);
    PROF_END(4);
  }

  // EPG basis at the fitted angle, phase states in registers: lane <-> state index, shifts are warp
  // shuffles, four T2 components advance together for instruction-level parallelism.  Arithmetic per
  // state update follows epg_impulse_response! (src/EPGdecaycurve.jl:948-1028) exactly; states outside
  // the reference's truncated range are simply carried along (they never feed back for ETL <= 63,
  // resp. <= 127 with the second register set).
  template <bool TWO>
  __device__ __noinline__ void epg_basis_shfl(double alpha_deg, long long v) {
    const int ETL = cP.nTE, n = cP.nT2, ld = cP.ld;
    double *prm = g + sl.pristine, *pcm = g + sl.pristine_cm;
    double sina, cosa;
    sincos(alpha_deg * 0.017453292519943295, &sina, &cosa);
    const double m0 = sind_0_180(alpha_deg / 2);
    const double E1 = cP.E1;
    constexpr int NS = TWO ? 2 : 1;
    for (int j0 = 0; j0 < n; j0 += 4) {
      double a[4], b[4], c[4], d[4], cp[4];
      double F[NS][4], Fb[NS][4], Z[NS][4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int j = (j0 + q < n) ? j0 + q : n - 1;
        const double E2 = cP.E2[j];
        const double E2h = __dmul_rn(E2, E2) / 2, E1E2 = __dmul_rn(E1, E2), E1sq = __dmul_rn(E1, E1);
        a[q] = E2h, b[q] = __dmul_rn(E2h, cosa), c[q] = __dmul_rn(E1E2, sina), d[q] = __dmul_rn(E1sq, cosa);
        cp[q] = -c[q] / 2;
#pragma unroll
        for (int sidx = 0; sidx < NS; sidx++) F[sidx][q] = 0.0, Fb[sidx][q] = 0.0, Z[sidx][q] = 0.0;
        // state after the first echo (:966-968): state 1 = (a-b, 0, c'), state 2 = (a+b, 0, 0)
        if (lane == 0) F[0][q] = __dsub_rn(a[q], b[q]), Z[0][q] = cp[q];
        if (lane == 1) F[0][q] = __dadd_rn(a[q], b[q]);
        if (lane == 0 && j0 + q < n) {
          double val = fabs(__dmul_rn(m0, __dsub_rn(a[q], b[q])));
          prm[0 * ld + j] = val, pcm[j * ETL + 0] = val;
        }
      }
      for (int i = 2; i <= ETL; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          double vF[NS], vFb[NS], vZ[NS];
#pragma unroll
          for (int sidx = 0; sidx < NS; sidx++) {
            double C = __dadd_rn(F[sidx][q], Fb[sidx][q]), Sd = __dsub_rn(F[sidx][q], Fb[sidx][q]);
            double Cp = __dmul_rn(a[q], C), Sp = __dmul_rn(b[q], Sd);
            if (i < ETL || sidx > 0) {
              vFb[sidx] = fma(-c[q], Z[sidx][q], __dsub_rn(Cp, Sp));
            } else {
              vFb[sidx] = fma(-c[q], Z[sidx][q], fma(a[q], C, __dmul_rn(-b[q], Sd)));  // last echo, :1024
            }
            vF[sidx] = fma(c[q], Z[sidx][q], __dadd_rn(Cp, Sp));
            vZ[sidx] = fma(cp[q], Sd, __dmul_rn(d[q], Z[sidx][q]));
          }
          // echo amplitude = new F of state 1
          if (lane == 0 && j0 + q < n) {
            double val = fabs(__dmul_rn(m0, vFb[0]));
            prm[(i - 1) * ld + j0 + q] = val, pcm[(j0 + q) * ETL + (i - 1)] = val;
          }
          // shifts: F moves up one state, Fbar moves down one state, Z stays
          double upF0 = __shfl_up_sync(DECAES_FULL_MASK, vF[0], 1);
          double dnFb0 = __shfl_down_sync(DECAES_FULL_MASK, vFb[0], 1);
          if (TWO) {
            double upF1 = __shfl_up_sync(DECAES_FULL_MASK, vF[NS - 1], 1);
            double wrapF = __shfl_sync(DECAES_FULL_MASK, vF[0], 31);        // state 32 -> state 33
            double dnFb1 = __shfl_down_sync(DECAES_FULL_MASK, vFb[NS - 1], 1);
            double wrapFb = __shfl_sync(DECAES_FULL_MASK, vFb[NS - 1], 0);  // state 33 -> state 32
            F[NS - 1][q] = (lane == 0) ? wrapF : upF1;
            Fb[NS - 1][q] = (lane == 31) ? 0.0 : dnFb1;
            Z[NS - 1][q] = vZ[NS - 1];
            Fb[0][q] = (lane == 31) ? wrapFb : dnFb0;
          } else {
            Fb[0][q] = (lane == 31) ? 0.0 : dnFb0;
          }
          F[0][q] = (lane == 0) ? vFb[0] : upF0;
          Z[0][q] = vZ[0];
        }
      }
    }
    __syncwarp();
    if (cP.decaybasis && !cP.fixed_alpha) {
      _Pragma("unroll 1") for (int k = lane; k < ETL * n; k += 32) cP.decaybasis[v + (long long)k * cP.stride] = pcm[k];
    }
  }

  // G = A'A (lower triangle) and, with `rhs`, c = A'bd from the basis in global scratch (element (i, j) at i * rs + j * cs:
  // row-major rs = ld, cs = 1; column-major rs = 1, cs = nTE) into shared memory.
  // The Gram contraction runs on the FP64 tensor cores: mma.sync m8n8k4 with D(p, q) += sum over four
  // echoes of A(i, p) A(i, q); both operand fragments are the SAME load pattern (lane (g, t) holds
  // A[i0 + t][8 c + g]), one 8-column tile row of G at a time.
  __device__ __noinline__ void gram_build(const double *Arm, int rs, int cs, bool rhs) {
    GL(Arm);
    const int lane = this->lane;
    VIEW(double, Gs);
    const int nTE = cP.nTE, n = cP.nT2, ldg = cP.ldg;
    const int g = lane >> 2, t = lane & 3;
    const int NT = (n + 7) >> 3;
    _Pragma("unroll 1") for (int cp = 0; cp < NT; cp++) {
      double acc[8][2];
#pragma unroll
      for (int cq = 0; cq < 8; cq++) acc[cq][0] = acc[cq][1] = 0.0;
      const bool pv = 8 * cp + g < n;
      _Pragma("unroll 1") for (int i0 = 0; i0 < nTE; i0 += 4) {  // (unroll 4: four L2 round trips per tile row instead of fourteen, but 4.6x the code: -4 % on the nT2 = 60 configs, nothing on cfg3)
        const bool iv = i0 + t < nTE;
        const double *row = Arm + (iv ? i0 + t : 0) * rs + g * cs;
        const int c8 = 8 * cs;
        const double fa = (iv && pv) ? row[cp * c8] : 0.0;
        double f[8];
#pragma unroll
        for (int cq = 0; cq < 8; cq++)
          if (cq <= cp) f[cq] = (iv && 8 * cq + g < n) ? row[cq * c8] : 0.0;
#pragma unroll
        for (int cq = 0; cq < 8; cq++)
          if (cq <= cp)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(acc[cq][0]), "+d"(acc[cq][1])
                         : "d"(fa), "d"(f[cq]));
      }
      const int p = 8 * cp + g;
#pragma unroll
      for (int cq = 0; cq < 8; cq++)
        if (cq <= cp && p < n) {
          const int q = 8 * cq + 2 * t;
          if (q <= p) Gs[p * ldg + q] = acc[cq][0];
          if (q + 1 <= p) Gs[p * ldg + q + 1] = acc[cq][1];
        }
    }
    __syncwarp();
    if (rhs) gram_rhs(Arm);  // (row-major only; the shared-memory EPG accumulates c on the fly)
  }

  // solve!(cache, mu) for the Gram solver: exact-mu hit, else warm-start from the nearest cached mu (lane <-> slot)
  __device__ __noinline__ void cache_solve_gram(double mu, const Src &src, double lmu, int hint) {
    const int lane = this->lane;
    VIEW(double, slot_mu);
    VIEW(double, slot_lmu);
    VIEW(double, slot_r2);
    VIEW(double, slot_x2);
    VIEW(unsigned long long, slot_mask);
    double *const slots_x_p = this->slots_x_p;  // shared or global (PipeParams::spill)
    VIEW(double, V);
    VIEW_GWS();
    const int n = cP.nT2;
    const bool isl = lane < DECAES_NCACHE;
    const double mui = isl ? slot_mu[lane] : 0.0;
    const unsigned hitm = __ballot_sync(DECAES_FULL_MASK, isl && mu == mui);  // NaN (empty slot) never compares equal
    if (hitm) {
      cur_slot = __ffs(hitm) - 1;
      return;
    }
    const unsigned nanm = __ballot_sync(DECAES_FULL_MASK, isl && isnan(mui));
    cur_slot = nanm ? __ffs(nanm) - 1 : (cur_slot + 1) % DECAES_NCACHE;
    if (cP.step_sync & solve_vote) cta_or(true);
    const double mu2 = __dmul_rn(mu, mu);
    if (isnan(lmu)) lmu = dlog(mu);
    unsigned long long wmask = 0ull;
    if (hint == 1) {
      wmask = n >= 64 ? ~0ull : (1ull << n) - 1ull;
    } else if (hint == 2) {
      wmask = fa_mask_best;
    }
    if (wmask) {
      // a feasible interior point on the hinted set: only the final minimiser (unique for mu > 0) matters
      _Pragma("unroll 1") for (int j = lane; j < n; j += 32) gws.x[j] = ((wmask >> j) & 1ull) ? 1.0 : 0.0;
      __syncwarp();
    } else {
      // nearest cached mu in log distance (first slot on ties) that has a non-empty active set
      unsigned long long key = ~0ull, best;
      if (isl && !isnan(mui) && slot_mask[lane] != 0ull) key = (unsigned long long)__double_as_longlong(fabs(lmu - slot_lmu[lane]));
      const int nearest = warp_argmin_bits(key, lane, best);
      if (best != ~0ull) {
        const unsigned long long m0 = slot_mask[nearest];
        wmask = m0;
        if (hint >= 3) {
          // second / third L-curve point (mu = e^-4.2, e^-1.8): the active set widens around the peaks it already has; start
          // from the inherited set dilated by hint - 2 columns (oracle: 2.7 instead of 3.8, 4.0 instead of 8.3 set changes)
          _Pragma("unroll 1") for (int d = 0; d < hint - 2; d++) wmask |= (wmask << 1) | (wmask >> 1);
          if (n < 64) wmask &= (1ull << n) - 1ull;
        }
        // A feasible interior point on the inherited set: only the final minimiser (unique for mu > 0) matters, so the warm start
        // does not fetch the cached solution of that mu from the spilled table (one L2 round trip per solve; DECAES_WARM_ONES=0:
        // start from the cached solution - same parity figures, -0.3 %, profiles/r02_s4_ab_warm_start_from_ones.txt).  New
        // columns of a dilated set start at a small positive value, so that a wrong guess leaves at once.
        if (cP.warm_ones) {
          _Pragma("unroll 1") for (int j = lane; j < n; j += 32) gws.x[j] = ((m0 >> j) & 1ull) ? 1.0 : (((wmask >> j) & 1ull) ? 1e-3 : 0.0);
        } else {
          const double *sx = slots_x_p + nearest * n;
          _Pragma("unroll 1") for (int j = lane; j < n; j += 32) gws.x[j] = ((m0 >> j) & 1ull) ? sx[j] : (((wmask >> j) & 1ull) ? 1e-3 : 0.0);
        }
        __syncwarp();
      }
    }
    GramOut o;
    PROF_BEGIN(9);
    bool solved = false;
    if (hint == 1 && (cP.lc_hints & 4)) {
      unsigned long long m = wmask;
      double xn;
      const int kd = gram_dense_solve<VS>(V, n, cP.ldg, mu2, m, xn);
      if (kd >= 0) {
        o.k = kd, o.mask = m, o.xnorm_sq = xn, o.iters = 1, o.nappend = 0, o.capped = false;
        solved = true;
      } else {
        if (m) wmask = m;
        _Pragma("unroll 1") for (int j = lane; j < n; j += 32) gws.x[j] = ((wmask >> j) & 1ull) ? 1.0 : 0.0;
        __syncwarp();
      }
    }
    if (!solved) o = gram_nnls<VS>(V, n, cP.ldg, mu2, n, wmask != 0ull, wmask);
    n_itercap += o.capped;
#ifdef DECAES_PROFILE
    if (lane == 0) {
      const int si = nsolve_voxel < 47 ? nsolve_voxel : 47;
      atomicAdd(&g_solve_hist[0][si], 1ull), atomicAdd(&g_solve_hist[1][si], (unsigned long long)o.nappend), atomicAdd(&g_solve_hist[2][si], (unsigned long long)o.iters);
    }
    nsolve_voxel++;
#endif
    PROF_END(9);
    PROF_BEGIN(2);
    double r2 = gram_residual(src.Acm, o.k);
    PROF_END(2);
    if (o.k > 0 && cP.refine_tikh) {
      // one refinement step on the explicit residual: x(mu) accurate to ~cond([A; mu I]) * eps, so
      // that ||Ax - b||^2 and ||x||^2 (the inputs of the mu searches) carry reference-level noise
      gram_refine(src.Acm, o.k, mu2);
      r2 = gram_residual(src.Acm, o.k);
      double acc = 0.0;
      _Pragma("unroll 1") for (int t = lane; t < o.k; t += 32) acc = fma(gws.s[t], gws.s[t], acc);
      o.xnorm_sq = warp_sum(acc);
    }
    double *sx = slots_x_p + cur_slot * n;
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) sx[j] = gws.x[j];
    if (lane == 0)
      slot_mu[cur_slot] = mu, slot_lmu[cur_slot] = lmu, slot_r2[cur_slot] = r2, slot_x2[cur_slot] = o.xnorm_sq, slot_mask[cur_slot] = o.mask;
    __syncwarp();
  }

  // ================= one voxel =================
  // The per-voxel chain is split into three phases so that the warps of a CTA can be kept in the
  // same phase (CTA barriers in the kernel): the pipeline is instruction-cache bound and warps that
  // run the same code at the same time share its lines.
  long long v_cur;
  double max_signal_cur, alpha_cur;

  // phase 1: normalise (src/T2mapSEcorr.jl:205-218) and fit the flip angle (:409-423)
  __device__ __noinline__ void phase_flip_angle(long long v, const double *signal /* global: image + v, echo stride cP.stride */) {
    const int lane = this->lane;
    VIEW(double, bd);
    const int nTE = cP.nTE;
    v_cur = v;
    nunreg_voxel = 0;
    fa_mask_best = 0ull;
    double mx = 0.0;
    _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) {
      double bi = __ldg(signal + (long long)i * cP.stride);
      bd[i] = bi;
      mx = bi > mx ? bi : mx;
    }
    const double max_signal = warp_max(mx);
    max_signal_cur = max_signal;
    if (max_signal > 0)
      _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) bd[i] = ddiv(bd[i], max_signal);
    __syncwarp();
    if (cP.alpha_provided) alpha_cur = cP.alpha[v];
    else if (cP.fixed_alpha) alpha_cur = cP.SetFlipAngle;
    else alpha_cur = optimize_flip_angle();
    if (cP.step_sync & 1) cta_drain();
  }

  // phase 2: EPG basis at the fitted angle (+ Gram matrix / right-hand side)
  __device__ __noinline__ void phase_basis() {
    if (cP.fixed_alpha && !cP.alpha_provided) {
      if constexpr (GRAM) {
        cursrc.G = cP.gram_set, cursrc.ldg = cP.ldg, cursrc.Arm = cP.basis_rm, cursrc.Acm = cP.basis_cm;
        // one basis for the whole run: its singular values stay in the warp's scratch after the first voxel
        if (cP.reg == 2 && cP.gcv_smem && !gcv_fixed_done) gcv_svd_shared(cP.basis_rm, cP.ld, 1, V), gcv_fixed_done = true;
        stage_bulk(Gs, cP.gram_set, (unsigned)(cP.a_elems * 8));
        gram_rhs(cursrc.Arm);
      }
    } else {
      basis_at(alpha_cur, v_cur);
    }
  }

  // phase 3: regularised NNLS, output maps, T2part epilogue
  __device__ __noinline__ void phase_solve_and_save() {
    const int lane = this->lane;
    VIEW(double, bd);
    VIEW(double, fit);
    double *const slots_x_p = this->slots_x_p;  // shared or global (PipeParams::spill)
    VIEWG(double, lc_pts_p);
    const NnlsWs ws = this->ws;
    SH(ws.x);
    SH(ws.w);
    const int nTE = cP.nTE, n = cP.nT2;
    const long long v = v_cur;
    const double max_signal = max_signal_cur, alpha = alpha_cur;
    const double *Asrc = (cP.fixed_alpha && !cP.alpha_provided) ? cP.basis_rm : g + sl.pristine;

    // T2_distribution!  src/T2mapSEcorr.jl:475-505
    double mu = CUDART_NAN, chi2 = CUDART_NAN;
    int src_kind = 0;  // 0: ws.x (unregularised solve just done), 1: cache slot, 2: zeros
    const bool want_chi2 = (cP.chi2factor != nullptr);
    switch (cP.reg) {
      case 0: {
        mu = 0.0, chi2 = 1.0;
        solve_unreg(Asrc);
      } break;
      case 1: {  // lsqnonneg_lcurve!  src/lsqnonneg.jl:812-840
        cache_reset();
        double logmu = lcurve_corner(Asrc);
        mu = dexp(logmu);
        cache_solve(mu, Asrc, logmu);
        src_kind = 1;
        if (want_chi2) {  // the unregularised solve only feeds chi2factor
          double r2 = cur_resnorm_sq();
          NnlsOut o = solve_unreg(Asrc);
          chi2 = r2 / o.rnorm_sq;
        }
      } break;
      case 2: {  // lsqnonneg_gcv!  src/lsqnonneg.jl:1136-1205
        solve_vote = 8;
        if (!(GRAM && cP.gcv_smem)) gcv_svdvals(Asrc);
        cache_reset();
        double logmu = gcv_minimize(Asrc);
        mu = exp(logmu);
        cache_solve(mu, Asrc, logmu);
        src_kind = 1;
        if (want_chi2) {
          double r2 = cur_resnorm_sq();
          NnlsOut o = solve_unreg(Asrc);
          chi2 = r2 / o.rnorm_sq;
        }
      } break;
      case 3:    // lsqnonneg_chi2!  src/lsqnonneg.jl:504-593
      case 4: {  // lsqnonneg_mdp!   src/lsqnonneg.jl:700-747
        solve_vote = 8;
        NnlsOut o = solve_unreg(Asrc);
        double res2_min = o.rnorm_sq;
        bool early = false;
        double target, ftol;
        int mode;
        if (cP.reg == 3) {
          early = (res2_min == 0 || o.nsetp == 0);
          target = __dmul_rn(cP.Chi2Factor, res2_min), ftol = 1e-3 * (cP.Chi2Factor - 1), mode = 0;
          if (early) mu = 0.0, chi2 = 1.0;
        } else {
          double sigma = cP.NoiseLevel / max_signal;
          double delta = __dmul_rn(sqrt((double)nTE), sigma);
          double acc = 0.0;
          _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) acc = fma(bd[i], bd[i], acc);
          double res2_max = warp_sum(acc);
          target = __dmul_rn(delta, delta), ftol = 1e-3 * target, mode = 1;
          if (delta <= sqrt(res2_min)) {
            early = true, mu = 0.0, chi2 = 1.0;
          } else if (delta >= sqrt(res2_max)) {
            early = true, mu = CUDART_INF, chi2 = res2_max / res2_min, src_kind = 2;
          }
        }
        if (early) {
          // the reference's save_results! reads a stale cache slot here (src/lsqnonneg.jl:465, 657);
          // we return the value the chooser itself returns and count the voxel
          n_early++;
        } else if (LEGACY && cP.reg == 3) {
          if constexpr (LEGACY) {
            cache_reset();
            mu = chi2_legacy(res2_min, Asrc);
            if (!(mu > 0.0)) {
              // mu == 0: x_final = x_unreg (:528-529) and f(0) = res2_min, so chi2 = 1; mu < 0 flags a doubling
              // search that did not end (the reference would keep doubling): reported as NaN.  Both are counted.
              chi2 = (mu == 0.0) ? 1.0 : CUDART_NAN;
              if (mu < 0.0) mu = CUDART_NAN;
              n_early++;
              solve_unreg(Asrc);
            } else {
              cache_solve(mu, Asrc);  // a cache hit: f(mu_final) was the last solve
              chi2 = cur_resnorm_sq() / res2_min;
              src_kind = 1;
            }
          }
        } else {
          cache_reset();
          double xf, ff;
          bracket_and_brent(target, mode, ftol, Asrc, xf, ff);
          if (isfinite(ff)) {
            mu = exp(xf);
            double res2_final = (mode == 0) ? __dmul_rn(target, 1 + ff) : target + ff;
            cache_solve(mu, Asrc, xf);
            chi2 = res2_final / res2_min;
            src_kind = 1;
          } else {
            mu = 0.0, chi2 = 1.0 / res2_min;
            solve_unreg(Asrc);
          }
        }
      } break;
    }

    solve_vote = 0;
    if (cP.step_sync & 14) cta_drain();  // the one drain of this phase (warps without a voxel: the kernel's main loop)
    // save_results!  src/T2mapSEcorr.jl:512-591
    double *xs = ws.w;  // dual no longer needed
    {
      const double *sx = slots_x_p + cur_slot * n;
      _Pragma("unroll 1") for (int j = lane; j < n; j += 32) {
        double xv = (src_kind == 0) ? ws.x[j] : (src_kind == 1 ? sx[j] : 0.0);
        xs[j] = __dmul_rn(xv, max_signal);
      }
    }
    double r2 = 0.0, rs = 0.0;
    // residual scratch: the QR path has the Householder vector buffer; the Gram path borrows the
    // L-curve point cache in global scratch, which is free by now (4 * DECAES_LC_MAX >= nTE doubles)
    double *resv = GRAM ? (lc_pts_p) : ws.u;
    if constexpr (GRAM) {
      const double *Acm = cursrc.Acm;
      _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) {
        double s0 = 0.0;
        _Pragma("unroll 1") for (int j = 0; j < n; j++) {
          double xj = xs[j];
          if (xj != 0.0) s0 = fma(Acm[j * nTE + i], xj, s0);
        }
        fit[i] = s0;
        double res = s0 - __dmul_rn(bd[i], max_signal);
        resv[i] = res;
        r2 = fma(res, res, r2);
        rs += res;
      }
    } else {
      stage_matrix(Asrc);  // pristine basis back into shared memory for fit = A x
      _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) {
        const double *row = ws.A + i * cP.ld;
        double s0 = 0.0;
        _Pragma("unroll 1") for (int j = 0; j < n; j++) s0 = fma(row[j], xs[j], s0);
        fit[i] = s0;
        double res = s0 - __dmul_rn(bd[i], max_signal);
        resv[i] = res;
        r2 = fma(res, res, r2);
        rs += res;
      }
    }
    __syncwarp();
    r2 = warp_sum(r2);
    double mean = warp_sum(rs) / nTE;
    double var = 0.0;
    _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) {
      double dlt = resv[i] - mean;
      var = fma(dlt, dlt, var);
    }
    var = warp_sum(var);
    double S = 0.0, dotl = 0.0;
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) S += xs[j], dotl = fma(xs[j], cP.logT2[j], dotl);
    S = warp_sum(S), dotl = warp_sum(dotl);
    double log_ggm = dotl / S;
    double l1p = 0.0;
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) {
      double dlt = cP.logT2[j] - log_ggm;
      l1p = fma(__dmul_rn(dlt, dlt), xs[j], l1p);
    }
    l1p = warp_sum(l1p) / S;

    if (lane == 0) {
      cP.gdn[v] = S;
      cP.ggm[v] = exp(log_ggm);
      cP.gva[v] = expm1(l1p);
      cP.fnr[v] = S / sqrt(r2 / (nTE - 1));
      cP.snr[v] = max_signal / sqrt(var / (nTE - 1));
      cP.alpha[v] = alpha;
      if (cP.mu && cP.chi2factor) cP.mu[v] = mu, cP.chi2factor[v] = chi2;
      if (cP.resnorm) cP.resnorm[v] = sqrt(r2);
    }
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) cP.dist[v + (long long)j * cP.stride] = xs[j];
    if (cP.decaycurve)
      _Pragma("unroll 1") for (int i = lane; i < nTE; i += 32) cP.decaycurve[v + (long long)i * cP.stride] = fit[i];

    // fused T2part epilogue  src/T2partSEcorr.jl:95-138
    if (cP.has_part) {
      bool isn = false;
      double Ssp = 0, Smp = 0, dsp = 0, dmp = 0, dw = 0;
      _Pragma("unroll 1") for (int j = lane; j < n; j += 32) {
        double dj = xs[j];
        isn |= isnan(dj);
        if (j >= cP.sp_lo && j <= cP.sp_hi) dsp += __dmul_rn(dj, cP.logT2[j]), Ssp += dj;
        if (j >= cP.mp_lo && j <= cP.mp_hi) dmp += __dmul_rn(dj, cP.logT2[j]), Smp += dj;
        if (cP.has_sigmoid) dw = fma(dj, cP.weights[j], dw);
      }
      Ssp = warp_sum(Ssp), Smp = warp_sum(Smp), dsp = warp_sum(dsp), dmp = warp_sum(dmp), dw = warp_sum(dw);
      const bool anynan = __any_sync(DECAES_FULL_MASK, isn);
      if (lane == 0) {
        // entries the reference leaves untouched keep its NaN pre-fill (src/T2partSEcorr.jl:83-90)
        double sfr = CUDART_NAN, mfr = CUDART_NAN, sgm = CUDART_NAN, mgm = CUDART_NAN;
        if (!anynan) {
          if (S > 0) sfr = cP.has_sigmoid ? dw / S : Ssp / S, mfr = Smp / S;
          if (Ssp > 0) sgm = exp(dsp / Ssp);
          if (Smp > 0) mgm = exp(dmp / Smp);
        }
        cP.sfr[v] = sfr, cP.mfr[v] = mfr, cP.sgm[v] = sgm, cP.mgm[v] = mgm;
      }
    }
    __syncwarp();
  }
};

}  // namespace decaes
