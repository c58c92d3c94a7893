// Microbenchmark: latency of TMEM as a per-lane scratchpad (tcgen05.ld / tcgen05.st, SASS LDTM / STTM) next to shared
// memory, for one warp per SM sub-partition.  Question behind it: can the 256 KB of tensor memory, idle
// in this FP64 vector kernel, hold per-voxel state (EPG phase states, Gram rows) that shared memory has no room for?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_lat tmem_lat.cu && ./tmem_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld2(uint32_t addr, uint32_t &a, uint32_t &b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_st2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 1) tmem_lat(long long *out, int iters, int active_warps) {
  __shared__ uint32_t base_s;
  __shared__ double sm[4][32 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = base_s + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 512; c += 2) tm_st2(taddr + c, 0u, 0u);
  tm_wait_st();
  for (int i = lane; i < 32 * 8; i += 32) sm[warp][i] = 0.0;
  __syncthreads();
  long long t[8] = {0};
  if (warp < active_warps) {
    uint32_t a = 0, b = 0, off = lane & 0;  // off stays 0, but only at run time
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {  // dependent chain of loads
      tm_ld2(taddr + ((off + a) & 0x1fe), a, b);
      tm_wait_ld();
    }
    t[0] = clock64() - t0;
    t0 = clock64();
    for (int i = 0; i < iters; i++) {  // store -> load of the same cell (in-place state update)
      tm_st2(taddr + ((off + a) & 0x1fe), a, b + 1);
      tm_wait_st();
      tm_ld2(taddr + ((off + a) & 0x1fe), a, b);
      tm_wait_ld();
      a &= 0;
    }
    t[1] = clock64() - t0;
    t0 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {  // 8 independent loads in flight, one wait
      uint32_t r[16];
#pragma unroll
      for (int q = 0; q < 8; q++) tm_ld2(taddr + ((off + a + 2 * q) & 0x1fe), r[2 * q], r[2 * q + 1]);
      tm_wait_ld();
#pragma unroll
      for (int q = 0; q < 16; q++) acc |= r[q];
      a = acc & 0;
    }
    t[2] = clock64() - t0;
    // shared-memory references: dependent LDS.64 chain, STS -> LDS round trip
    volatile double *s = sm[warp];
    double v = 0.0;
    t0 = clock64();
    for (int i = 0; i < iters; i++) v = s[lane + ((int)v & 7) * 32];
    t[3] = clock64() - t0;
    t0 = clock64();
    for (int i = 0; i < iters; i++) {
      s[lane + ((int)v & 7) * 32] = v;
      v = s[lane + ((int)v & 7) * 32];
    }
    t[4] = clock64() - t0;
    if (lane == 0)
      for (int q = 0; q < 5; q++) out[(blockIdx.x * 4 + warp) * 8 + q] = t[q] + (long long)(v + a + b) * 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "n"(512));
}

int main() {
  long long *d, h[148 * 4 * 8];
  cudaMalloc(&d, sizeof h);
  const int iters = 4096;
  for (int aw : {1, 4}) {
    cudaMemset(d, 0, sizeof h);
    tmem_lat<<<148, 128>>>(d, iters, aw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const char *nm[5] = {"LDTM dependent chain", "STTM->LDTM same cell", "8 LDTM in flight + wait (per group)", "LDS.64 dependent chain", "STS->LDS same cell"};
    for (int q = 0; q < 5; q++) printf("active warps/SM %d  %-36s %8.1f cycles\n", aw, nm[q], (double)h[q] / iters);
  }
  return 0;
}
