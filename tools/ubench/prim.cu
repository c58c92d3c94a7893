// Isolated cost of the Gram-solver primitives (one voxel per warp, W warps per CTA, 148 CTAs).
#include <cstdio>
#include "../../decaes.jl_b200/csrc/gram.cuh"
using namespace decaes;
__global__ void kern(long long *cyc, double *sink, int n, int ld, int k, int reps) {
  extern __shared__ double smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *V = smem + (size_t)wid * (GV_T + n * ld + 8);
  double *T = V + GV_T;
  // SPD matrix G = B'B + I, lower triangle
  for (int e = lane; e < n * ld; e += 32) T[e] = 0.0;
  __syncwarp();
  for (int p = lane; p < n; p += 32)
    for (int q = 0; q <= p; q++) T[p * ld + q] = (p == q ? 2.0 : 0.0) + 1.0 / (1.0 + p + q) + 0.9 * exp(-0.3 * (p - q));
  for (int j = lane; j < n; j += 32) V[GV_C + j] = 1.0 + 0.1 * j, V[GV_X + j] = 0.0;
  int *P = (int *)(V + GV_P);
  if (lane < k) P[lane] = 3 * lane + 1;
  __syncwarp();
  long long t0, t1, acc[4] = {0, 0, 0, 0};
  for (int r = 0; r < reps; r++) {
    t0 = clock64();
    bool ok = gram_factor(V, ld, k, 1e-3);
    t1 = clock64();
    acc[0] += t1 - t0;

    t0 = clock64();
    ok &= gram_append(V, ld, k, 3 * k + 2, 1e-3, false);
    t1 = clock64();
    acc[1] += t1 - t0;
    if (!ok) sink[0] = -1;
    // dual
    unsigned long long mask = mask_of(P, k + 1);
    t0 = clock64();
    for (int j = lane; j < n; j += 32) {
      double a = V[GV_C + j];
#pragma unroll 4
      for (int t = 0; t < k + 1; t++) a = fma(-gram_G(T, ld, P[t], j), V[GV_S + t], a);
      V[GV_W + j] = ((mask >> j) & 1ull) ? 0.0 : a;
    }
    __syncwarp();
    t1 = clock64();
    acc[2] += t1 - t0;
    sink[1] = V[GV_S];

  }
  if (lane == 0 && blockIdx.x == 0 && wid == 0)
    for (int c = 0; c < 4; c++) cyc[c] = acc[c] / reps;
}
int main() {
  long long *cyc; double *sink;
  cudaMalloc(&cyc, 64); cudaMalloc(&sink, 64);
  int n = 40, ld = 41;
  for (int W = 1; W <= 8; W *= 8)
    for (int k = 2; k <= 8; k += 1) {
      size_t sm = (size_t)W * (GV_T + n * ld + 8) * 8;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      kern<<<148, 32 * W, sm>>>(cyc, sink, n, ld, k, 200);
      cudaDeviceSynchronize();
      long long h[4]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      printf("W=%d k=%2d: factor %lld  append(at k) %lld  dual(k+1) %lld  factor_reg %lld   %s\n", W, k, h[0], h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
    }
}
