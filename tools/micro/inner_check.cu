// Unit check of the 32x32 inner Jacobi sweep: J unitary, J^H G J == rotated G.

#include "../../oqupy_b200/csrc/svd.cu"
#include "../../oqupy_b200/csrc/common.cu"
#include <cstdio>
#include <vector>
#include <random>
#include <complex>
namespace {
__global__ void __launch_bounds__(JT, 1) inner_test(const cplx* g, cplx* jout, cplx* gout, long long* cyc, unsigned mask) {
  __shared__ InnerShared S;
  __shared__ cplx sj[PB * PB];
  const int t = threadIdx.x;
  for (int e = t; e < PB * PB; e += JT) { S.gr[e >> 5][e & 31] = g[e].x; S.gi[e >> 5][e & 31] = g[e].y; }
  __syncthreads();
  long long t0 = clock64();
  inner_sweep(S, mask, 1e-60, 0.0, sj);
  long long t1 = clock64();
  if (t == 0) cyc[0] = t1 - t0;
  for (int e = t; e < PB * PB; e += JT) { jout[e] = sj[e]; gout[e] = make_double2(S.gr[e >> 5][e & 31], S.gi[e >> 5][e & 31]); }
}
}
typedef std::complex<double> Z;
int main() {
  std::mt19937_64 rng(1);
  std::normal_distribution<double> nd;
  const int m = 64;
  std::vector<Z> a(m * 32), g(32 * 32);
  for (auto& v : a) v = Z(nd(rng), nd(rng));
  for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) { Z s = 0; for (int r = 0; r < m; ++r) s += std::conj(a[r * 32 + i]) * a[r * 32 + j]; g[i * 32 + j] = s; }
  cplx *dg, *dj, *dgo; long long* cyc;
  cudaMalloc(&dg, 16 * 1024); cudaMalloc(&dj, 16 * 1024); cudaMalloc(&dgo, 16 * 1024); cudaMallocManaged(&cyc, 64);
  std::vector<Z> jtot(1024, 0.0), J(1024), Go(1024);
  for (int i = 0; i < 32; ++i) jtot[i * 32 + i] = 1.0;
  std::vector<Z> gcur = g;
  for (int sweep = 0; sweep < 7; ++sweep) {
    const unsigned masks[7] = {1u, 1u << 16, 3u, 0xffffu, 0x7fffffffu, 0x7fffffffu, 0x7fffffffu};
    cudaMemcpy(dg, gcur.data(), 16 * 1024, cudaMemcpyHostToDevice);
    inner_test<<<1, JT>>>(dg, dj, dgo, cyc, masks[sweep]);
    cudaDeviceSynchronize();
    cudaMemcpy(J.data(), dj, 16 * 1024, cudaMemcpyDeviceToHost);
    cudaMemcpy(Go.data(), dgo, 16 * 1024, cudaMemcpyDeviceToHost);
    // host: G' = J^H G J
    std::vector<Z> tmp(1024), gn(1024);
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) { Z s = 0; for (int k = 0; k < 32; ++k) s += gcur[i * 32 + k] * J[k * 32 + j]; tmp[i * 32 + j] = s; }
    double uerr = 0, off = 0, diag = 0;
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) {
      Z s = 0, u = 0;
      for (int k = 0; k < 32; ++k) { s += std::conj(J[k * 32 + i]) * tmp[k * 32 + j]; u += std::conj(J[k * 32 + i]) * J[k * 32 + j]; }
      gn[i * 32 + j] = s;
      uerr = std::max(uerr, std::abs(u - (i == j ? 1.0 : 0.0)));
      if (i != j) off = std::max(off, std::abs(s)); else diag = std::max(diag, std::abs(s));
    }
        printf("sweep %d: %lld cycles, unitarity err %.2e, max offdiag of J^H G J %.3e (diag max %.3e), sorted: %d\n", sweep, cyc[0], uerr, off, diag,
           (int)(gn[0].real() >= gn[33].real() && gn[33].real() >= gn[66].real()));
    { double dmax = 0; int nan = 0; for (int e = 0; e < 1024; ++e) { if (Go[e] != Go[e]) ++nan; dmax = std::max(dmax, std::abs(Go[e] - gn[e])); }
      printf("    mask %08x: device G vs host J^H G J: max diff %.3e, NaNs %d; J[0][0..3] = %.3f %.3f %.3f %.3f; col norms:", masks[sweep], dmax, nan, J[0].real(), J[1].real(), J[2].real(), J[3].real());
      for (int j = 0; j < 32; j += 5) { double nn = 0; for (int k = 0; k < 32; ++k) nn += std::norm(J[k * 32 + j]); printf(" %.3f", nn); } printf("\n"); }
    gcur = gn;
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
