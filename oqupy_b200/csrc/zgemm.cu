// Strided doubly-batched complex128 GEMM with per-batch scale (C-ABI: b200_zgemm_strided).
//
// One kernel serves every contraction of the TEMPO / PT-TEMPO path: the influence
// MPO sites are delta-structured (B[l,x,y,r] = d_lr d_xy infl[l,x]), so the
// site contraction is d2*d2 independent (k x chi)(chi x chi') products whose
// operands are strided views of the carry and of the MPS site, scaled by one
// influence-matrix entry each.  Addressing everything with element strides lets the
// result land directly in the (rows=(k,y), cols=(l,l')) layout the SVD consumes.
//
// Two kernels behind the one entry point:
//   * zgemm_dmma_kernel (m, n >= 32 and k >= 16: the dense contractions of the zip-up and of
//     the absorb step once the bond dimension is a few dozen): 64x64 output tile per CTA,
//     K-tile 16 staged through shared memory (register prefetch of the next K-tile), eight
//     warps of 16x32 outputs each on the FP64 TENSOR CORES: mma.sync.m8n8k4.f64 (SASS
//     DMMA.8x8x4), four real products per complex tile
//         Cr += Ar Br - Ai Bi,   Ci += Ar Bi + Ai Br.
//     The staged tiles are k-major with a row stride of 66 complex numbers (= 2 mod 8), so
//     the (g, t) fragment loads of a quarter warp hit eight different 16-byte bank groups.
//   * zgemm_strided_kernel (everything else: the tiny d2 x d2 legs, vectors): 64x64 tile,
//     4x4 complex128 accumulators per thread on the FP64 FMA pipe, K-tile 8.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 8, NT = 256;

struct Operand {
  const cplx* p;
  long long row, col, b1, b2;
  int conj;
};

struct GemmArgs {
  int m, n, k, nb1, nb2;
  Operand a, b;
  cplx* c;
  long long c_row, c_col, c_b1, c_b2;
  const cplx* scale;
  long long s_b1, s_b2;
  int accumulate;
};

__global__ void __launch_bounds__(NT)
zgemm_strided_kernel(const GemmArgs g) {
  __shared__ cplx As[TK][TM + 1];
  __shared__ cplx Bs[TK][TN + 1];

  const int bz = blockIdx.z;
  const int ib1 = bz / g.nb2, ib2 = bz % g.nb2;
  const cplx* __restrict__ A = g.a.p + ib1 * g.a.b1 + ib2 * g.a.b2;
  const cplx* __restrict__ B = g.b.p + ib1 * g.b.b1 + ib2 * g.b.b2;
  cplx* __restrict__ C = g.c + ib1 * g.c_b1 + ib2 * g.c_b2;

  const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // loader mappings: 512 elements per tile, 2 per thread; the fast thread index
  // follows the operand's smaller stride so global loads coalesce.
  const bool a_kfast = g.a.col <= g.a.row;
  const bool b_kfast = g.b.row <= g.b.col;
  int a_r[2], a_k[2], b_k[2], b_c[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int e = tid + NT * r;
    if (a_kfast) { a_k[r] = e % TK; a_r[r] = e / TK; }
    else         { a_r[r] = e % TM; a_k[r] = e / TM; }
    if (b_kfast) { b_k[r] = e % TK; b_c[r] = e / TK; }
    else         { b_c[r] = e % TN; b_k[r] = e / TN; }
  }

  cplx acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_double2(0.0, 0.0);

  cplx ra[2], rb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int gr = row0 + a_r[r], gk = k0 + a_k[r];
      cplx v = make_double2(0.0, 0.0);
      if (gr < g.m && gk < g.k) {
        v = A[gr * g.a.row + gk * g.a.col];
        if (g.a.conj) v.y = -v.y;
      }
      ra[r] = v;
      const int gc = col0 + b_c[r], gk2 = k0 + b_k[r];
      cplx w = make_double2(0.0, 0.0);
      if (gc < g.n && gk2 < g.k) {
        w = B[gk2 * g.b.row + gc * g.b.col];
        if (g.b.conj) w.y = -w.y;
      }
      rb[r] = w;
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < g.k; k0 += TK) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      As[a_k[r]][a_r[r]] = ra[r];
      Bs[b_k[r]][b_c[r]] = rb[r];
    }
    __syncthreads();
    if (k0 + TK < g.k) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      cplx av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = b200::cfma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  cplx sc = make_double2(1.0, 0.0);
  if (g.scale) sc = g.scale[ib1 * g.s_b1 + ib2 * g.s_b2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = row0 + ty + 16 * i;
    if (gr >= g.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = col0 + tx + 16 * j;
      if (gc >= g.n) continue;
      cplx v = b200::cmul(sc, acc[i][j]);
      cplx* dst = C + gr * g.c_row + gc * g.c_col;
      if (g.accumulate) {
        cplx o = *dst;
        v.x += o.x;
        v.y += o.y;
      }
      *dst = v;
    }
  }
}

// ------------------------------------------------------------------ FP64 tensor-core path
constexpr int DK = 16;            // K-tile
constexpr int DS = TM + 2;        // row stride of the staged tiles (complex numbers)

__global__ void __launch_bounds__(NT)
zgemm_dmma_kernel(const GemmArgs g) {
  __shared__ cplx As[DK][DS];
  __shared__ cplx Bs[DK][DS];

  const int bz = blockIdx.z;
  const int ib1 = bz / g.nb2, ib2 = bz % g.nb2;
  const cplx* __restrict__ A = g.a.p + ib1 * g.a.b1 + ib2 * g.a.b2;
  const cplx* __restrict__ B = g.b.p + ib1 * g.b.b1 + ib2 * g.b.b2;
  cplx* __restrict__ C = g.c + ib1 * g.c_b1 + ib2 * g.c_b2;

  const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wy = warp >> 1, wx = warp & 1;       // 4 x 2 warps: 16 rows x 32 columns each
  const int fg = lane >> 2, ft = lane & 3;        // fragment coordinates

  // loaders: 64 x 16 elements per operand tile, four per thread; the fast thread index
  // follows the operand's smaller stride
  const bool a_kfast = g.a.col <= g.a.row;
  const bool b_kfast = g.b.row <= g.b.col;
  int a_r[4], a_k[4], b_k[4], b_c[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = tid + NT * r;
    if (a_kfast) { a_k[r] = e % DK; a_r[r] = e / DK; }
    else         { a_r[r] = e % TM; a_k[r] = e / TM; }
    if (b_kfast) { b_k[r] = e % DK; b_c[r] = e / DK; }
    else         { b_c[r] = e % TN; b_k[r] = e / TN; }
  }
  // accumulators: [row block 0..1][column block 0..3] x (re, im) x 2 doubles
  double cr[2][4][2], ci[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }

  cplx ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int gr = row0 + a_r[r], gk = k0 + a_k[r];
      cplx v = make_double2(0.0, 0.0);
      if (gr < g.m && gk < g.k) {
        v = A[gr * g.a.row + gk * g.a.col];
        if (g.a.conj) v.y = -v.y;
      }
      ra[r] = v;
      const int gc = col0 + b_c[r], gk2 = k0 + b_k[r];
      cplx w = make_double2(0.0, 0.0);
      if (gc < g.n && gk2 < g.k) {
        w = B[gk2 * g.b.row + gc * g.b.col];
        if (g.b.conj) w.y = -w.y;
      }
      rb[r] = w;
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < g.k; k0 += DK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      As[a_k[r]][a_r[r]] = ra[r];
      Bs[b_k[r]][b_c[r]] = rb[r];
    }
    __syncthreads();
    if (k0 + DK < g.k) fetch(k0 + DK);
#pragma unroll
    for (int kk = 0; kk < DK; kk += 4) {
      cplx av[2], bv[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) av[i] = As[kk + ft][wy * 16 + i * 8 + fg];   // A[g][t]
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk + ft][wx * 32 + j * 8 + fg];   // B[t][g]
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double nai = -av[i].y;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          b200::dmma884(cr[i][j][0], cr[i][j][1], av[i].x, bv[j].x);
          b200::dmma884(cr[i][j][0], cr[i][j][1], nai, bv[j].y);
          b200::dmma884(ci[i][j][0], ci[i][j][1], av[i].x, bv[j].y);
          b200::dmma884(ci[i][j][0], ci[i][j][1], av[i].y, bv[j].x);
        }
      }
    }
    __syncthreads();
  }

  cplx sc = make_double2(1.0, 0.0);
  if (g.scale) sc = g.scale[ib1 * g.s_b1 + ib2 * g.s_b2];
  // D[g][2t], D[g][2t+1]
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int gr = row0 + wy * 16 + i * 8 + fg;
    if (gr >= g.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gc = col0 + wx * 32 + j * 8 + 2 * ft + h;
        if (gc >= g.n) continue;
        cplx v = b200::cmul(sc, make_double2(cr[i][j][h], ci[i][j][h]));
        cplx* dst = C + gr * g.c_row + gc * g.c_col;
        if (g.accumulate) {
          const cplx o = *dst;
          v.x += o.x;
          v.y += o.y;
        }
        *dst = v;
      }
    }
  }
}

}  // namespace

extern "C" int b200_zgemm_strided(void* stream, int m, int n, int k, int nb1,
                                  int nb2, const b200_operand* a,
                                  const b200_operand* b, void* c, int64_t c_row,
                                  int64_t c_col, int64_t c_b1, int64_t c_b2,
                                  const void* scale, int64_t s_b1, int64_t s_b2,
                                  int accumulate) {
  if (m < 0 || n < 0 || k < 0 || nb1 < 1 || nb2 < 1 || !a || !b || !c) {
    b200::set_error("b200_zgemm_strided: invalid argument");
    return B200_EINVAL;
  }
  if (m == 0 || n == 0) return B200_OK;
  if ((long long)nb1 * nb2 > 65535) {
    b200::set_error("b200_zgemm_strided: too many batches (%d x %d)", nb1, nb2);
    return B200_ESIZE;
  }
  GemmArgs g;
  g.m = m; g.n = n; g.k = k; g.nb1 = nb1; g.nb2 = nb2;
  g.a = Operand{(const cplx*)a->ptr, a->row, a->col, a->b1, a->b2, a->conj};
  g.b = Operand{(const cplx*)b->ptr, b->row, b->col, b->b1, b->b2, b->conj};
  g.c = (cplx*)c;
  g.c_row = c_row; g.c_col = c_col; g.c_b1 = c_b1; g.c_b2 = c_b2;
  g.scale = (const cplx*)scale;
  g.s_b1 = s_b1; g.s_b2 = s_b2;
  g.accumulate = accumulate;
  dim3 grid((n + TN - 1) / TN, (m + TM - 1) / TM, nb1 * nb2);
  static const int dmma_on = [] { const char* e = getenv("B200_ZGEMM_DMMA"); return e ? atoi(e) : 1; }();
  if (dmma_on && m >= 32 && n >= 32 && k >= 16)
    zgemm_dmma_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(g);
  else
    zgemm_strided_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(g);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
