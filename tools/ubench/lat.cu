// Dependent-issue latencies on B200 (one warp per SM, clock64 around N dependent ops).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k(double *out, long long *cyc, int nw) {
  __shared__ double sm[1024];
  __shared__ int smi[1024];
  int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i, smi[i] = (i * 7 + 1) & 1023;
  __syncthreads();
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0, t1;
  int idx = lane;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = fma(a, b, c);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = a + c;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = a * b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // LDS pointer chase (32-bit)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) idx = smi[idx];
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // LDS.64 -> DFMA chain (address independent)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = fma(a, sm[(i + lane) & 1023], c);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // SHFL 64-bit chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = __shfl_xor_sync(0xffffffffu, a, 1) + c;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // REDUX chain
  unsigned r = idx;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) r = __reduce_max_sync(0xffffffffu, r + lane) ;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // syncwarp + STS/LDS round trip
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31] + c; __syncwarp(); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  // independent DFMA throughput (8 chains)
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c); x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  a += x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  // ddiv chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 1024; i++) a = b / a + c;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[9] = t1 - t0;
  // sqrt chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 1024; i++) a = sqrt(a) + b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[10] = t1 - t0;
  // global (L2) pointer chase through out[]
  long long p = 0;
  const long long *chase = (const long long *)(out + 1024);
  t0 = clock64();
  for (int i = 0; i < 256; i++) p = __ldcg(chase + p);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[11] = t1 - t0;
  // local/L1 hit: ld.global.ca same address chain
  t0 = clock64();
  for (int i = 0; i < 256; i++) p = __ldca(chase + p);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[12] = t1 - t0;
  out[threadIdx.x + blockIdx.x * blockDim.x] = a + idx + r + p;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 24); cudaMemset(out, 0, 1 << 24);
  // pointer-chase ring with stride 4 KB inside a 8 MB region (L2 resident after the first lap)
  { int n = 256; long long *h = new long long[1 << 20]; for (int i = 0; i < (1 << 20); i++) h[i] = 0; for (int i = 0; i < n; i++) h[(size_t)i * 512] = (long long)((i + 1) % n) * 512; cudaMemcpy(out + 1024, h, sizeof(long long) << 20, cudaMemcpyHostToDevice); }
  cudaMalloc(&cyc, 64 * 8);
  const char *nm[] = {"DFMA dep", "DADD dep", "DMUL dep", "LDS chase", "LDS.64->DFMA", "SHFL64+DADD", "REDUX", "STS+sync+LDS+sync", "DFMA x8 indep (per 8)", "DDIV+DADD (per op)", "DSQRT+DADD", "L2 chase (ldcg)", "L1 chase (ldca)"};
  int div[] = {N, N, N, N, N, N, N, N, N, 1024, 1024, 256, 256};
  for (int warps = 1; warps <= 8; warps *= 2) {
    k<<<1, 32 * warps>>>(out, cyc, warps); k<<<1, 32 * warps>>>(out, cyc, warps);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("warps/CTA=%d:", warps);
    for (int i = 0; i < 13; i++) printf("  %s=%.1f", nm[i], (double)h[i] / div[i]);
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
