// Shared helpers for the oqupy_b200 C-ABI library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/oqupy_b200.h"

typedef double2 cplx;

namespace b200 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// roofline profiling of the dominant kernel (see b200_profile_enable)
bool profile_on();
// kind: 0 Jacobi stage, 1 rank-revealing QR, 2 back-transformation + emit, 3 other
void profile_begin(cudaStream_t s, int kind);
void profile_end(cudaStream_t s, int kind, double flops, const int* sweeps_dev);

#define B200_CUDA_CHECK(expr)                                                   \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                      cudaGetErrorString(_e));                                  \
      return B200_ECUDA;                                                        \
    }                                                                           \
  } while (0)

#define B200_LAUNCH_CHECK()                                                     \
  do {                                                                          \
    b200::count_launch();                                                       \
    B200_CUDA_CHECK(cudaGetLastError());                                        \
  } while (0)

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {
  c.x = fma(a.x, b.x, c.x);
  c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y);
  c.y = fma(a.y, b.x, c.y);
  return c;
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }

// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor-core path (SASS: DMMA.8x8x4).
// Fragment layout (PTX ISA, mma.m8n8k4 .f64):  g = lane>>2, t = lane&3
//   a : A[g][t]     b : B[t][g]     d0,d1 : D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a,
                                        double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, "
      "{%0,%1};\n"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

}  // namespace b200
