// Device-side helpers shared by the kernels of libdecaes_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <float.h>
#include <math.h>

#define DECAES_FULL_MASK 0xffffffffu
#define DECAES_MAX_ANGLES 64   // flip-angle grid is tracked in one 64-bit mask
#define DECAES_MAX_NT2 64      // per-column flags live in one 64-bit mask
#define DECAES_MAX_NTE 96      // lane <-> echoes lane, lane + 32, lane + 64 in the explicit residuals
#define DECAES_LC_MAX 64       // L-curve point / state cache capacity per voxel
#define DECAES_NCACHE 8        // NNLSTikhonovRegProblemCache slots (src/lsqnonneg.jl:396)
#define DECAES_GROUP 4         // voxels fetched per work item = one 32-byte sector per echo
#ifndef DECAES_MAX_WARPS
#define DECAES_MAX_WARPS 12    // warps per persistent CTA (register file: 65536 / (12*32) = 170 regs/thread)
#endif

namespace decaes {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Out-of-line math: the pipeline kernel is I-cache bound (ncu: stall_no_inst), so the software
// sequences behind double-precision div / sqrt / log / exp are kept as single shared copies.
__device__ __noinline__ double ddiv(double a, double b) { return a / b; }
__device__ __noinline__ double dsqrt(double a) { return sqrt(a); }
__device__ __noinline__ double drsqrt(double a) { return rsqrt(a); }
__device__ __noinline__ double dlog(double a) { return log(a); }
__device__ __noinline__ double dexp(double a) { return exp(a); }

// Butterfly sums: every lane ends with the bitwise-identical total (addition commutes).
__device__ __noinline__ double warp_sum(double v) {  // out of line: ~40 call sites, the kernel is instruction-fetch bound
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DECAES_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(DECAES_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ double warp_bcast(double v, int src) { return __shfl_sync(DECAES_FULL_MASK, v, src); }

// (value, index) reductions with "first index wins on ties", matching sequential scans
// that replace only on strict comparison.
__device__ __noinline__ void warp_argmax_first(double &v, int &i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(DECAES_FULL_MASK, v, o);
    int oi = __shfl_xor_sync(DECAES_FULL_MASK, i, o);
    if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
  }
}
__device__ __noinline__ void warp_argmin_first(double &v, int &i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(DECAES_FULL_MASK, v, o);
    int oi = __shfl_xor_sync(DECAES_FULL_MASK, i, o);
    if (ov < v || (ov == v && oi < i)) v = ov, i = oi;
  }
}

// Base.Math.hypot (Borges' fma-corrected algorithm) as used by orthogonal_rotmat,
// src/NNLS.jl:486-491.
__device__ __forceinline__ double hypot_julia(double x, double y) {
  if (isinf(x) || isinf(y)) return CUDART_INF;
  if (isnan(x) || isnan(y)) return CUDART_NAN;
  double ax = fabs(x), ay = fabs(y);
  if (ay > ax) {
    double t = ax;
    ax = ay, ay = t;
  }
  if (ay <= ax * 1.0536712127723509e-08 /* sqrt(eps/2) */) return ax;
  double scale = DBL_EPSILON * 1.4916681462400413e-154 /* sqrt(floatmin) */;
  if (ax > 9.480751908109176e+153 /* sqrt(floatmax/2) */) {
    ax *= scale, ay *= scale, scale = 1.0 / scale;
  } else if (ay < 1.4916681462400413e-154) {
    ax /= scale, ay /= scale;
  } else {
    scale = 1.0;
  }
  double h = sqrt(fma(ax, ax, __dmul_rn(ay, ay)));
  double hsq = __dmul_rn(h, h), axsq = __dmul_rn(ax, ax);
  h -= (fma(-ay, ay, hsq - axsq) + fma(h, h, -hsq) - fma(ax, ax, -axsq)) / (2 * h);
  return h * scale;
}

// sind(x) for x in [0, 180] with the reduction of Base.Math.sind (src/EPGdecaycurve.jl:940 uses
// sind(alpha/2)): sin below 45 deg, cos(90 - x) up to 135 deg, sin(180 - x) above; the radian
// argument is formed in double-double so the result matches the extended-precision reference
// to the last bit or so.
__device__ __forceinline__ double deg2rad_dd(double d, double &lo) {
  const double khi = 0.017453292519943295, klo = 2.9486522708701687e-19; // pi/180 split
  double hi = d * khi;
  lo = fma(d, khi, -hi) + d * klo;
  return hi;
}
__device__ __forceinline__ double sind_0_180(double x) {
  double lo, hi, s, c;
  if (x < 45.0) {
    hi = deg2rad_dd(x, lo);
    sincos(hi, &s, &c);
    return fma(c, lo, s);
  } else if (x <= 135.0) {
    hi = deg2rad_dd(90.0 - x, lo);
    sincos(hi, &s, &c);
    return fma(-s, lo, c);
  } else {
    hi = deg2rad_dd(180.0 - x, lo);
    sincos(hi, &s, &c);
    return fma(c, lo, s);
  }
}

// ---- mbarrier + TMA bulk copy (global -> shared), one elected lane issues ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
// dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace decaes
