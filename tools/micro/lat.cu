// Micro-benchmarks for the latency model of the Jacobi inner solver (B200).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[1024];
  const int t = threadIdx.x;
  sm[t] = seed + t; sm[t + 512] = seed * 2 + t;
  __syncthreads();
  double x = seed + 1e-3 * t, y = 0.999999;
  long long t0, t1;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (t == 0) cyc[0] = (t1 - t0);
  // 2. dependent rsqrt chain
  double z = x + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) z = rsqrt(z) + 1.5;
  t1 = clock64();
  if (t == 0) cyc[1] = (t1 - t0);
  // 3. barrier (all threads of the block)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64();
  if (t == 0) cyc[2] = (t1 - t0);
  // 4. dependent LDS chain
  int idx = t;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) idx = ((int)sm[idx & 1023] + i) & 1023;
  t1 = clock64();
  if (t == 0) cyc[3] = (t1 - t0);
  // 5. STS -> BAR -> LDS round trip
  double w = z;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { sm[t] = w; __syncthreads(); w = sm[(t + 1) % blockDim.x] + 1.0; __syncthreads(); }
  t1 = clock64();
  if (t == 0) cyc[4] = (t1 - t0);
  // 6. divergent: only lane 0 of each warp runs a DFMA chain, then barrier
  double v = w;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if ((t & 31) == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v = fma(v, y, 1e-9);
    }
    __syncthreads();
  }
  t1 = clock64();
  if (t == 0) cyc[5] = (t1 - t0);
  // 7. independent DFMA throughput (8 chains)
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
    a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
  }
  t1 = clock64();
  if (t == 0) cyc[6] = (t1 - t0);
  out[blockIdx.x * blockDim.x + t] = x + z + idx + w + v + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 64);
  const char* names[] = {"dep DFMA x256", "dep rsqrt+add x64", "syncthreads x64", "dep LDS x64",
                         "STS-BAR-LDS-BAR x64", "divergent 16xDFMA + BAR x16", "8 indep DFMA chains x64"};
  const int per[] = {256, 64, 64, 64, 64, 16, 512};
  for (int nt : {32, 256, 512}) {
    k<<<1, nt>>>(out, cyc, 1.25);
    cudaDeviceSynchronize();
    k<<<1, nt>>>(out, cyc, 1.25);
    cudaDeviceSynchronize();
    printf("threads %d\n", nt);
    for (int i = 0; i < 7; ++i) printf("  %-32s %8lld cyc  -> %.1f per op\n", names[i], cyc[i], (double)cyc[i] / per[i]);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
