// eps-truncated complex128 SVD on the device: one-sided BLOCK Jacobi, row-sliced.
//
// Replaces tn.split_node_full_svd(max_truncation_err=eps, relative=True)
// (reference call sites oqupy/backends/node_array.py:262,285,541).
//
// Algorithm (B200-first, not LAPACK's bidiagonalisation).  The SVDs of a TEMPO chain
// are strictly sequential, so what matters is the LATENCY of one factorisation:
//   * X = theta (m >= n) or theta^H (m < n), p x q with p >= q.  Y = [X ; W] stacks X
//     on top of the q x q rotation accumulator W (initially I); T = p + q rows.  Y is
//     stored in BLOCKS of 16 columns, block j = Y[:, 16j:16j+16] as [row][16]
//     complex128 (256 B per row, fully coalesced).
//   * Hestenes one-sided Jacobi on block pairs (32 columns), round-robin tournament.
//     The grid is a 2-D decomposition: CTA (s, r) owns pair-slot s and ROW SLICE r of
//     the stacked matrix.  Per stage it
//       1. stages its slice of the two blocks in shared memory,
//       2. forms the partial 32x32 Gram matrix of its X rows and publishes it,
//       3. (slice 0 = leader) sums the partials in a fixed order (deterministic),
//          runs ONE cyclic sweep of a two-sided Jacobi eigensolver on the 32x32
//          Hermitian Gram matrix (cross-block pairs first, 2x2-block ownership: no
//          buffer hazards), sorts the columns by descending norm and publishes the
//          32x32 rotation J,
//       4. applies J to its slice of [X ; W] and writes it back.
//     Hand-over between stages is point to point through L2 (st.release / ld.acquire
//     per (block, slice)); only the end of a sweep is a grid-wide barrier (the
//     convergence vote).
//   * Convergence: a full sweep in which no pair violated
//       |g_ij|^2 <= max(a,b) * (tol^2 * min(a,b) + floor^2).
//   * In the same launch: column norms -> sort -> the reference tail-norm rule picks
//     `keep`, written straight to pinned host memory.  emit_kernel then writes
//     U[:, :keep] and S*Vh[:keep] in the layouts the MPS engine asks for.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int BC = 16;            // columns per block
constexpr int PB = 2 * BC;        // columns per block pair (the inner problem)
constexpr int JT = 512;           // threads of the Jacobi kernel
constexpr int MAX_SWEEPS = 120;   // <= NFLAGS
constexpr int NFLAGS = 128;
constexpr int FLOOR_GROW_AFTER = 90;
constexpr int CHUNK_ROWS = 256;   // rows of a slice staged in shared memory at a time
constexpr int MIN_SLICE_ROWS = 32;
constexpr int GP = PB + 1;        // padded leading dimension of the 32x32 work matrices

struct Header {        // lives at the start of the workspace (device)
  int m, n, p, q, nb, transposed, keep, sweeps;
  int status, rotations, R, RS;
  double eps, s0, fro2, pad2;
  long long phase_cycles[8];  // CTA 0: wait, load, gram, solve/J-wait, apply+store, vote, -, stages
};

struct Layout {
  size_t header, ctrl, ctrl_bytes, sig2, sval, perm, gpart, jbuf, y, total;
  int p, q, T, nb, S, R, RS, SE, transposed;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int device_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0) {
      (void)cudaGetLastError();
      sms = 148;   // B200
    }
  }
  return sms;
}

// ctrl region (ints, zeroed by one memset per factorisation):
//   [0]            grid barrier counter
//   [1]            spare
//   [2 .. 2+NFLAGS)           rotated-stage count per sweep
//   ready[nb*R]               stages completed per (block, slice)
//   cnt[2*S]                  slices that published their partial Gram (double-buffered)
//   flagj[S]                  2*(stage+1) + need, per slot
__host__ Layout make_layout(int m, int n) {
  Layout L;
  L.transposed = (m < n) ? 1 : 0;
  L.p = L.transposed ? n : m;
  L.q = L.transposed ? m : n;
  L.T = L.p + L.q;
  int nblk = (L.q + BC - 1) / BC;
  L.nb = (nblk & 1) ? nblk + 1 : nblk;   // even number of blocks
  if (L.nb < 2) L.nb = 2;
  L.S = L.nb / 2;
  const int sms = device_sms();
  int max_r = sms / L.S;
  if (max_r < 1) max_r = 1;
  int r = (L.T + MIN_SLICE_ROWS - 1) / MIN_SLICE_ROWS;
  if (r > max_r) r = max_r;
  int rs = (L.T + r - 1) / r;
  rs = (rs + 3) & ~3;
  L.RS = rs;
  L.R = (L.T + rs - 1) / rs;
  L.SE = L.S;                       // slots resident at once
  if (L.SE * L.R > sms) L.SE = sms / L.R;
  if (L.SE < 1) L.SE = 1;
  size_t off = 0;
  L.header = off; off = align256(off + sizeof(Header));
  L.ctrl = off;
  L.ctrl_bytes = sizeof(int) * (size_t)(2 + NFLAGS + (size_t)L.nb * L.R + 3 * (size_t)L.S);
  off = align256(off + L.ctrl_bytes);
  L.sig2 = off; off = align256(off + (size_t)L.nb * BC * sizeof(double));
  L.sval = off; off = align256(off + (size_t)L.nb * BC * sizeof(double));
  L.perm = off; off = align256(off + (size_t)L.nb * BC * sizeof(int));
  L.gpart = off; off = align256(off + (size_t)2 * L.S * L.R * PB * PB * sizeof(cplx));
  L.jbuf = off; off = align256(off + (size_t)2 * L.S * PB * PB * sizeof(cplx));
  // W accumulates the right rotations.  (Forming S*Vh as U^H*theta instead is NOT an
  // option: columns of U are orthogonal only down to the absolute rounding floor, and
  // the projection would amplify that by sigma_0/sigma_j.)
  L.y = off; off = align256(off + (size_t)L.nb * L.T * BC * sizeof(cplx));
  L.total = off;
  return L;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ cplx ldcg(const cplx* p) {
  return __ldcg(reinterpret_cast<const double2*>(p));
}

// grid-wide barrier on a monotonically increasing counter (cooperative launch
// guarantees co-residency).  `epoch` counts the barriers this CTA has passed.
__device__ __forceinline__ void grid_barrier(int* counter, int& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    __threadfence();
    red_release_add(counter, 1);
    const int target = epoch * (int)gridDim.x;
    while (ld_acquire(counter) < target) {}
  }
  __syncthreads();
}

// A pair (i, j) counts as orthogonal when
//   |g_ij| <= sqrt(max(a,b)) * (tol*sqrt(min(a,b)) + floor),  floor = kappa*eps*||X||_F:
// the accumulated rotations are accurate in the ABSOLUTE sense (eps*||X||), so columns
// that have sunk to the rounding floor are left alone.  (threshold combined in
// quadrature: no sqrt on the critical path.)  Pairs whose two columns are both
// NEGLIGIBLE (norm^2 < neg2 = (1e-2*eps*||X||_F)^2) are only orthogonalised loosely:
// such columns are discarded by the truncation rule whatever their mutual angles, and
// the tail norm only needs the Frobenius norm of their block, which is rotation
// invariant.  (neg2 = 0 when no truncation is requested.)
__device__ __forceinline__ bool pair_converged(double a, double b, double g2,
                                               double tol2, double floor2, double neg2) {
  const double big = fmax(a, b), small = fmin(a, b);
  if (big <= 0.0) return true;
  if (big < neg2) return g2 <= 1e-4 * big * fmax(small, 0.0) + big * floor2;
  return g2 <= big * (tol2 * fmax(small, 0.0) + floor2);
}

// Round-robin partner tables ---------------------------------------------------------
// outer tournament over nb blocks (nb even), stage in [0, nb-1)
__device__ __forceinline__ void outer_pair(int idx, int stage, int nb, int& pa, int& pb) {
  const int ka = idx, kb = nb - 1 - idx;
  pa = (ka == 0) ? 0 : 1 + ((ka - 1 + stage) % (nb - 1));
  pb = 1 + ((kb - 1 + stage) % (nb - 1));
  if (pa > pb) { const int t = pa; pa = pb; pb = t; }
}
// inner ordering over 32 columns, 31 rounds of 16 disjoint pairs: rounds 0..15 pair
// every column of the first block with every column of the second (the pairs that have
// never met), rounds 16..30 are two simultaneous 16-player tournaments inside the blocks.
__device__ __forceinline__ void inner_pair(int round, int k, int& i, int& j) {
  if (round < BC) {
    i = k;
    j = BC + ((k + round) & (BC - 1));
  } else {
    const int st = round - BC;          // 0..14
    const int half = k >> 3, kk = k & 7;
    const int ka = kk, kb = BC - 1 - kk;
    int a = (ka == 0) ? 0 : 1 + ((ka - 1 + st) % (BC - 1));
    int b = 1 + ((kb - 1 + st) % (BC - 1));
    if (a > b) { const int t = a; a = b; b = t; }
    i = half * BC + a;
    j = half * BC + b;
  }
}

struct InnerShared {
  double gr[PB][GP], gi[PB][GP];   // Hermitian working matrix
  double jr[PB][GP], ji[PB][GP];   // accumulated rotations
  double rc[BC], rsr[BC], rsi[BC]; // per pair of the round: c, s e^{i phi}
  int rp[BC], rq[BC], ract[BC];
  int order[PB];
  int need, flag;
};

// One cyclic sweep of two-sided Jacobi on the 32x32 Hermitian matrix S.gr + i S.gi,
// accumulating J.  All JT threads participate.  Thread (a, b), a, b in [0,16), owns the
// 2x2 block (rows of pair a) x (columns of pair b): G' = R_a^H G R_b touches only its
// own four entries, so the update is in place.  Threads 256.. own two rows of J each.
// R = [[c, se], [-conj(se), c]] acting on columns (p, q).
__device__ void inner_sweep(InnerShared& S, double floor2, double neg2) {
  const int t = threadIdx.x;
  for (int round = 0; round < PB - 1; ++round) {
    int act = 0;
    if (t < BC) {
      int i, j;
      inner_pair(round, t, i, j);
      const double a = S.gr[i][i], b = S.gr[j][j];
      const double xr = S.gr[i][j], xi = S.gi[i][j];
      const double mag2 = xr * xr + xi * xi;
      double c = 1.0, sr_ = 0.0, si_ = 0.0;
      if (mag2 > 0.0 && !pair_converged(a, b, mag2, 1e-28, 0.0625 * floor2, neg2)) {
        // cos(2t) = |h|/r, c = sqrt((1+cos 2t)/2), s e^{i phi} = sign(h) g/(2 r c)
        const double h = 0.5 * (b - a);
        const double inv_r = rsqrt(h * h + mag2);
        const double w = 0.5 + 0.5 * fabs(h) * inv_r;
        const double ic = rsqrt(w);            // two rsqrt: no sqrt, no division
        c = w * ic;
        const double k = ((h >= 0.0) ? 0.5 : -0.5) * inv_r * ic;
        sr_ = xr * k;
        si_ = xi * k;
        act = 1;
      }
      S.rp[t] = i; S.rq[t] = j; S.ract[t] = act;
      S.rc[t] = c; S.rsr[t] = sr_; S.rsi[t] = si_;
    }
    if (!__syncthreads_or(act)) continue;
    if (t < 256) {
      const int a = t >> 4, b = t & 15;
      const int aa = S.ract[a], ab = S.ract[b];
      if (aa | ab) {
        const int pa = S.rp[a], qa = S.rq[a], pb = S.rp[b], qb = S.rq[b];
        const double ca = S.rc[a], sar = S.rsr[a], sai = S.rsi[a];
        const double cb = S.rc[b], sbr = S.rsr[b], sbi = S.rsi[b];
        const double g00r = S.gr[pa][pb], g00i = S.gi[pa][pb];
        const double g01r = S.gr[pa][qb], g01i = S.gi[pa][qb];
        const double g10r = S.gr[qa][pb], g10i = S.gi[qa][pb];
        const double g11r = S.gr[qa][qb], g11i = S.gi[qa][qb];
        // column op: T[:,0] = cb g[:,0] - conj(seb) g[:,1] ; T[:,1] = seb g[:,0] + cb g[:,1]
        const double t00r = cb * g00r - (sbr * g01r + sbi * g01i);
        const double t00i = cb * g00i - (sbr * g01i - sbi * g01r);
        const double t01r = cb * g01r + (sbr * g00r - sbi * g00i);
        const double t01i = cb * g01i + (sbr * g00i + sbi * g00r);
        const double t10r = cb * g10r - (sbr * g11r + sbi * g11i);
        const double t10i = cb * g10i - (sbr * g11i - sbi * g11r);
        const double t11r = cb * g11r + (sbr * g10r - sbi * g10i);
        const double t11i = cb * g11i + (sbr * g10i + sbi * g10r);
        // row op: G'[0,:] = ca T[0,:] - sea T[1,:] ; G'[1,:] = conj(sea) T[0,:] + ca T[1,:]
        double n00r = ca * t00r - (sar * t10r - sai * t10i);
        double n00i = ca * t00i - (sar * t10i + sai * t10r);
        double n01r = ca * t01r - (sar * t11r - sai * t11i);
        double n01i = ca * t01i - (sar * t11i + sai * t11r);
        double n10r = ca * t10r + (sar * t00r + sai * t00i);
        double n10i = ca * t10i + (sar * t00i - sai * t00r);
        double n11r = ca * t11r + (sar * t01r + sai * t01i);
        double n11i = ca * t11i + (sar * t01i - sai * t01r);
        if (a == b) {            // diagonal block: real diagonal, annihilated off-diagonal
          n00i = 0.0; n11i = 0.0;
          n01r = 0.0; n01i = 0.0; n10r = 0.0; n10i = 0.0;
        }
        S.gr[pa][pb] = n00r; S.gi[pa][pb] = n00i;
        S.gr[pa][qb] = n01r; S.gi[pa][qb] = n01i;
        S.gr[qa][pb] = n10r; S.gi[qa][pb] = n10i;
        S.gr[qa][qb] = n11r; S.gi[qa][qb] = n11i;
      }
    } else {
      const int u = t - 256;
      const int b = u & 15, r0 = (u >> 4) * 2;
      if (S.ract[b]) {
        const int pb = S.rp[b], qb = S.rq[b];
        const double cb = S.rc[b], sbr = S.rsr[b], sbi = S.rsi[b];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = r0 + rr;
          const double v0r = S.jr[r][pb], v0i = S.ji[r][pb];
          const double v1r = S.jr[r][qb], v1i = S.ji[r][qb];
          S.jr[r][pb] = cb * v0r - (sbr * v1r + sbi * v1i);
          S.ji[r][pb] = cb * v0i - (sbr * v1i - sbi * v1r);
          S.jr[r][qb] = cb * v1r + (sbr * v0r - sbi * v0i);
          S.ji[r][qb] = cb * v1i + (sbr * v0i + sbi * v0r);
        }
      }
    }
    __syncthreads();
  }
}

// Partial Gram matrix of `rows` rows of the staged tile ([row][32] complex) into
// registers: thread (kh, oi, oj) accumulates the 2x2 block rows (2oi, 2oi+1), columns
// (2oj, 2oj+1) of T^H T over the rows r = kh (mod 2).
__device__ __forceinline__ void gram_accumulate(const cplx* tile, int rows, double acc[8]) {
  const int t = threadIdx.x;
  const int kh = t >> 8, oi = (t >> 4) & 15, oj = t & 15;
  for (int r = kh; r < rows; r += 2) {
    const cplx* row = tile + (size_t)r * PB;
    const cplx a0 = row[2 * oi], a1 = row[2 * oi + 1];
    const cplx b0 = row[2 * oj], b1 = row[2 * oj + 1];
    // conj(a) * b
    acc[0] = fma(a0.x, b0.x, acc[0]); acc[0] = fma(a0.y, b0.y, acc[0]);
    acc[1] = fma(a0.x, b0.y, acc[1]); acc[1] = fma(-a0.y, b0.x, acc[1]);
    acc[2] = fma(a0.x, b1.x, acc[2]); acc[2] = fma(a0.y, b1.y, acc[2]);
    acc[3] = fma(a0.x, b1.y, acc[3]); acc[3] = fma(-a0.y, b1.x, acc[3]);
    acc[4] = fma(a1.x, b0.x, acc[4]); acc[4] = fma(a1.y, b0.y, acc[4]);
    acc[5] = fma(a1.x, b0.y, acc[5]); acc[5] = fma(-a1.y, b0.x, acc[5]);
    acc[6] = fma(a1.x, b1.x, acc[6]); acc[6] = fma(a1.y, b1.y, acc[6]);
    acc[7] = fma(a1.x, b1.y, acc[7]); acc[7] = fma(-a1.y, b1.x, acc[7]);
  }
}

// out[row][:] = tile[row][:] * J  for `rows` rows; NR rows per thread.  Output columns
// 0..15 go to block A, 16..31 to block B (both [row][16] in global memory).
template <int NR>
__device__ __forceinline__ void apply_tile(const cplx* tile, int rows, const cplx* sj,
                                           cplx* outA, cplx* outB) {
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;     // 16 column pairs x 32 row groups
  for (int rbase = ty * NR; rbase < rows; rbase += 32 * NR) {
    double ar[NR][2], ai[NR][2];
#pragma unroll
    for (int x = 0; x < NR; ++x) { ar[x][0] = ar[x][1] = ai[x][0] = ai[x][1] = 0.0; }
#pragma unroll 4
    for (int k = 0; k < PB; ++k) {
      const cplx j0 = sj[k * PB + 2 * tx], j1 = sj[k * PB + 2 * tx + 1];
#pragma unroll
      for (int x = 0; x < NR; ++x) {
        const int r = (rbase + x < rows) ? rbase + x : rows - 1;
        const cplx v = tile[(size_t)r * PB + k];
        ar[x][0] = fma(v.x, j0.x, ar[x][0]); ar[x][0] = fma(-v.y, j0.y, ar[x][0]);
        ai[x][0] = fma(v.x, j0.y, ai[x][0]); ai[x][0] = fma(v.y, j0.x, ai[x][0]);
        ar[x][1] = fma(v.x, j1.x, ar[x][1]); ar[x][1] = fma(-v.y, j1.y, ar[x][1]);
        ai[x][1] = fma(v.x, j1.y, ai[x][1]); ai[x][1] = fma(v.y, j1.x, ai[x][1]);
      }
    }
#pragma unroll
    for (int x = 0; x < NR; ++x) {
      const int r = rbase + x;
      if (r < rows) {
        cplx* dst = (tx < 8) ? (outA + (size_t)r * BC + 2 * tx)
                             : (outB + (size_t)r * BC + 2 * (tx - 8));
        dst[0] = make_double2(ar[x][0], ai[x][0]);
        dst[1] = make_double2(ar[x][1], ai[x][1]);
      }
    }
  }
}

// global -> shared: rows [r0, r0+rows) of blocks A and B into tile[row][32]
__device__ __forceinline__ void load_tile(cplx* tile, const cplx* gA, const cplx* gB,
                                          int rows) {
  const int total = rows * PB;
  for (int e = threadIdx.x; e < total; e += JT) {
    const int r = e >> 5, c = e & 31;
    const cplx* src = (c < BC) ? (gA + (size_t)r * BC + c) : (gB + (size_t)r * BC + (c - BC));
    tile[e] = ldcg(src);
  }
}

// ------------------------------------------------------------------ Jacobi kernel
// Persistent cooperative kernel: load -> sweeps -> column norms -> rank rule.
__global__ void __launch_bounds__(JT, 1)
jacobi_kernel(const cplx* __restrict__ theta, long long rs, long long cs,
              cplx* __restrict__ y, cplx* __restrict__ gpart, cplx* __restrict__ jbuf,
              int* __restrict__ ctrl, double* __restrict__ sig2,
              double* __restrict__ sval, int* __restrict__ perm,
              Header* __restrict__ hdr, int32_t* __restrict__ info, int p, int q, int nb,
              int R, int RS, int SE, int transposed, int minmn, double tol, double eps,
              double neg_rel) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ InnerShared S;
  __shared__ double s_red[JT / 32];
  cplx* tile = reinterpret_cast<cplx*>(dyn_smem);          // [CHUNK][32]
  cplx* sj = tile + (size_t)CHUNK_ROWS * PB;               // [32][32] rotation to apply
  cplx* sg = sj + PB * PB;                                 // [32][32] scratch (Gram halves)

  const int T = p + q;
  const int S_slots = nb / 2;
  int* bar = ctrl;
  int* flags = ctrl + 2;
  int* ready = flags + NFLAGS;
  int* cnt = ready + (size_t)nb * R;
  int* flagj = cnt + 2 * S_slots;
  int epoch = 0;
  const int t = threadIdx.x;
  const size_t blk_elems = (size_t)T * BC;

  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tq = clock64();
#define PHASE(k) { const long long tn_ = clock64(); pc[k] += tn_ - tq; tq = tn_; }

  // ---- load: Y = [X ; I], ||X||_F^2
  {
    double local = 0.0;
    const long long total = (long long)nb * T * BC;
    for (long long e = blockIdx.x * (long long)JT + t; e < total;
         e += (long long)gridDim.x * JT) {
      const int c16 = (int)(e % BC);
      const long long u = e / BC;
      const int row = (int)(u % T);
      const int blk = (int)(u / T);
      const int col = blk * BC + c16;
      cplx v = make_double2(0.0, 0.0);
      if (row < p) {
        if (col < q) {
          if (!transposed) {
            v = theta[row * rs + col * cs];
          } else {           // X = theta^H : X[row][col] = conj(theta[col][row])
            v = theta[col * rs + row * cs];
            v.y = -v.y;
          }
        }
        local = fma(v.x, v.x, local);
        local = fma(v.y, v.y, local);
      } else if (row - p == col) {
        v.x = 1.0;
      }
      y[e] = v;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((t & 31) == 0) s_red[t >> 5] = local;
    __syncthreads();
    if (t == 0) {
      double tot = 0.0;
      for (int w = 0; w < JT / 32; ++w) tot += s_red[w];
      if (tot != 0.0) atomicAdd(&hdr->fro2, tot);
    }
    __threadfence();
  }
  grid_barrier(bar, epoch);
  const double fro = sqrt(__ldcg(&hdr->fro2));
  PHASE(1)

  const int r_slice = blockIdx.x % R;
  const int s_first = blockIdx.x / R;
  const int row0 = r_slice * RS;
  const int nrows = min(RS, T - row0);            // rows of this slice (> 0 by layout)
  const int xrows = max(0, min(nrows, p - row0)); // of which X rows (the Gram part)
  const bool single_chunk = nrows <= CHUNK_ROWS;
  const bool leader = (r_slice == 0);

  int sweeps_done = 0, total_rot = 0, status = 1;
  const double tol2 = tol * tol;
  const double neg2 = (neg_rel * fro) * (neg_rel * fro);
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    int my_rot = 0;
    // absolute floor: 8 eps ||X||_F, doubled every sweep after FLOOR_GROW_AFTER so that
    // the iteration always terminates
    double kappa = 8.0;
    for (int k = FLOOR_GROW_AFTER; k < sweep; ++k) kappa *= 2.0;
    const double floor_ = kappa * 2.220446049250313e-16 * fro;
    const double floor2 = floor_ * floor_;
    for (int stage = 0; stage < nb - 1; ++stage) {
      const int g = sweep * (nb - 1) + stage;
      const int par = g & 1;
      for (int s = s_first; s < S_slots; s += SE) {
        int pa, pb;
        outer_pair(s, stage, nb, pa, pb);
        cplx* gA = y + pa * blk_elems + (size_t)row0 * BC;
        cplx* gB = y + pb * blk_elems + (size_t)row0 * BC;
        // 1. wait until both input blocks (this slice) have finished the previous stage
        if (t == 0) {
          while (ld_acquire(ready + pa * R + r_slice) < g) {}
          while (ld_acquire(ready + pb * R + r_slice) < g) {}
        }
        __syncthreads();
        PHASE(0)
        // 2./3. stage the slice, partial Gram of the X rows
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c0 = 0; c0 < nrows; c0 += CHUNK_ROWS) {
          const int crow = min(CHUNK_ROWS, nrows - c0);
          if (c0 > 0) __syncthreads();
          load_tile(tile, gA + (size_t)c0 * BC, gB + (size_t)c0 * BC, crow);
          __syncthreads();
          const int xr = max(0, min(crow, xrows - c0));
          if (xr > 0) gram_accumulate(tile, xr, acc);
        }
        PHASE(1)
        cplx* my_part = gpart + ((size_t)(par * S_slots + s) * R + r_slice) * (PB * PB);
        if (xrows > 0) {
          const int kh = t >> 8, oi = (t >> 4) & 15, oj = t & 15;
          if (kh == 1) {
            cplx* d = sg + (2 * oi) * PB + 2 * oj;
            d[0] = make_double2(acc[0], acc[1]); d[1] = make_double2(acc[2], acc[3]);
            d[PB] = make_double2(acc[4], acc[5]); d[PB + 1] = make_double2(acc[6], acc[7]);
          }
          __syncthreads();
          if (kh == 0) {
            const cplx* d = sg + (2 * oi) * PB + 2 * oj;
            cplx* o = my_part + (2 * oi) * PB + 2 * oj;
            o[0] = make_double2(acc[0] + d[0].x, acc[1] + d[0].y);
            o[1] = make_double2(acc[2] + d[1].x, acc[3] + d[1].y);
            o[PB] = make_double2(acc[4] + d[PB].x, acc[5] + d[PB].y);
            o[PB + 1] = make_double2(acc[6] + d[PB + 1].x, acc[7] + d[PB + 1].y);
          }
          __threadfence();
        }
        __syncthreads();
        if (t == 0) red_release_add(cnt + par * S_slots + s, 1);
        PHASE(2)
        // 4. leader: reduce, decide, solve, publish J
        int need;
        cplx* my_j = jbuf + (size_t)(par * S_slots + s) * (PB * PB);
        if (leader) {
          if (t == 0) {
            const int want = (g / 2 + 1) * R;     // cnt is cumulative per parity
            while (ld_acquire(cnt + par * S_slots + s) < want) {}
          }
          __syncthreads();
          const int rx = (p + RS - 1) / RS;       // slices that hold X rows
          const cplx* base = gpart + (size_t)(par * S_slots + s) * R * (PB * PB);
          for (int e = t; e < PB * PB; e += JT) {
            double sr_ = 0.0, si_ = 0.0;
            for (int rr = 0; rr < rx; ++rr) {
              const cplx v = ldcg(base + (size_t)rr * (PB * PB) + e);
              sr_ += v.x; si_ += v.y;
            }
            const int i = e >> 5, j = e & 31;
            S.gr[i][j] = sr_; S.gi[i][j] = si_;
            S.jr[i][j] = (i == j) ? 1.0 : 0.0; S.ji[i][j] = 0.0;
          }
          __syncthreads();
          // exact Hermitian symmetry (upper triangle wins), convergence test
          int viol = 0;
          for (int e = t; e < PB * PB; e += JT) {
            const int i = e >> 5, j = e & 31;
            if (i < j) {
              const double a = S.gr[i][i], b = S.gr[j][j];
              const double xr = S.gr[i][j], xi = S.gi[i][j];
              if (!pair_converged(a, b, xr * xr + xi * xi, tol2, floor2, neg2)) viol = 1;
            }
          }
          need = __syncthreads_or(viol);
          if (need) {
            for (int e = t; e < PB * PB; e += JT) {
              const int i = e >> 5, j = e & 31;
              if (i > j) { S.gr[i][j] = S.gr[j][i]; S.gi[i][j] = -S.gi[j][i]; }
              if (i == j) S.gi[i][j] = 0.0;
            }
            __syncthreads();
            inner_sweep(S, floor2, neg2);
            // sort columns by descending norm^2 (the diagonal of the rotated Gram matrix)
            if (t < PB) {
              const double lam = S.gr[t][t];
              int rank = 0;
#pragma unroll 8
              for (int o = 0; o < PB; ++o) {
                const double lo = S.gr[o][o];
                if (lo > lam || (lo == lam && o < t)) ++rank;
              }
              S.order[rank] = t;
            }
            __syncthreads();
            for (int e = t; e < PB * PB; e += JT) {
              const int i = e >> 5, j = e & 31;
              const int src = S.order[j];
              const cplx v = make_double2(S.jr[i][src], S.ji[i][src]);
              sj[e] = v;
              my_j[e] = v;
            }
            __threadfence();
            ++my_rot;
          }
          __syncthreads();
          if (t == 0) st_release(flagj + s, 2 * (g + 1) + (need ? 1 : 0));
        } else {
          if (t == 0) {
            int f;
            while ((f = ld_acquire(flagj + s)) < 2 * (g + 1)) {}
            S.flag = f;
          }
          __syncthreads();
          need = S.flag & 1;
          if (need) {
            for (int e = t; e < PB * PB; e += JT) sj[e] = ldcg(my_j + e);
            __syncthreads();
          }
        }
        PHASE(3)
        // 5. apply J to this slice of [X ; W]
        if (need) {
          for (int c0 = 0; c0 < nrows; c0 += CHUNK_ROWS) {
            const int crow = min(CHUNK_ROWS, nrows - c0);
            if (!single_chunk) {
              __syncthreads();
              load_tile(tile, gA + (size_t)c0 * BC, gB + (size_t)c0 * BC, crow);
              __syncthreads();
            }
            cplx* oA = gA + (size_t)c0 * BC;
            cplx* oB = gB + (size_t)c0 * BC;
            if (crow <= 32) apply_tile<1>(tile, crow, sj, oA, oB);
            else if (crow <= 64) apply_tile<2>(tile, crow, sj, oA, oB);
            else apply_tile<4>(tile, crow, sj, oA, oB);
          }
          __threadfence();
        }
        __syncthreads();
        if (t == 0) {   // release is cumulative over the CTA barrier above
          st_release(ready + pa * R + r_slice, g + 1);
          st_release(ready + pb * R + r_slice, g + 1);
        }
        PHASE(4)
        ++pc[7];
      }
    }
    // convergence vote: flags[sweep] counts the stages rotated in this sweep
    if (t == 0 && my_rot) atomicAdd(&flags[sweep], my_rot);
    grid_barrier(bar, epoch);
    const int rot = __ldcg(&flags[sweep]);
    PHASE(5)
    total_rot += rot;
    sweeps_done = sweep + 1;
    if (rot == 0) { status = 0; break; }
  }

  // ---- sigma_j^2 = ||X[:, j]||^2 : one warp per column block, all CTAs
  {
    const int lane = t & 31, warp = t >> 5;
    for (int blk = blockIdx.x * (JT / 32) + warp; blk < nb; blk += gridDim.x * (JT / 32)) {
      const cplx* X = y + blk * blk_elems;
      const int c = lane & 15, half = lane >> 4;
      double acc = 0.0;
      for (int row = half; row < p; row += 2) {
        const cplx v = ldcg(X + (size_t)row * BC + c);
        acc = fma(v.x, v.x, acc);
        acc = fma(v.y, v.y, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      if (lane < 16) sig2[blk * BC + c] = acc;
    }
    __threadfence();
  }
  grid_barrier(bar, epoch);
  if (blockIdx.x != 0) return;

  // ---- CTA 0: sort (descending), tail-norm rule, publish keep
  {
    const int ncols = nb * BC;
    int npow = 1;
    while (npow < ncols) npow <<= 1;
    double* key = reinterpret_cast<double*>(dyn_smem);
    int* idx = reinterpret_cast<int*>(key + npow);
    for (int e = t; e < npow; e += JT) {
      key[e] = (e < ncols) ? __ldcg(sig2 + e) : -1.0;   // padding sorts last
      idx[e] = e;
    }
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int e = t; e < npow; e += JT) {
          const int o = e ^ j;
          if (o > e) {
            const bool desc = ((e & k) == 0);
            const double a = key[e], b = key[o];
            const int ia = idx[e], ib = idx[o];
            const bool a_first = (a > b) || (a == b && ia < ib);
            if (desc ? !a_first : a_first) {
              key[e] = b; key[o] = a;
              idx[e] = ib; idx[o] = ia;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int e = t; e < q; e += JT) {
      sval[e] = sqrt(fmax(key[e], 0.0));
      perm[e] = idx[e];
    }
    __syncthreads();
    if (t == 0) {
      // keep = #{ j : sqrt(sum_{i>=j} s_i^2) > eps*s_0 }, accumulated from the
      // smallest value upwards exactly like numpy.cumsum(s[::-1]**2).
      const int r = minmn;            // number of genuine singular values
      int keep = r;
      const double s0 = (r > 0) ? sqrt(fmax(key[0], 0.0)) : 0.0;
      if (eps >= 0.0) {
        const double thr = eps * s0;
        double tail = 0.0;
        keep = 0;
        for (int j = r - 1; j >= 0; --j) {
          const double s = sqrt(fmax(key[j], 0.0));
          tail += s * s;
          if (sqrt(tail) > thr) ++keep;
        }
      }
      PHASE(6)
      hdr->keep = keep;
      hdr->s0 = s0;
      hdr->sweeps = sweeps_done;
      hdr->status = status;
      hdr->rotations = total_rot;
      for (int k = 0; k < 8; ++k) hdr->phase_cycles[k] = pc[k];
      info[0] = keep;
      info[1] = sweeps_done;
      info[2] = status;
      info[3] = total_rot;
    }
  }
#undef PHASE
}

// ------------------------------------------------------------------ emit
__global__ void emit_kernel(const cplx* __restrict__ y, const double* __restrict__ sval,
                            const int* __restrict__ perm, int m, int n, int p, int q,
                            int transposed, int keep, cplx* __restrict__ u, int u_na,
                            long long u_so, long long u_sa, long long u_sj,
                            cplx* __restrict__ svh) {
  // element space: [0, m*keep) -> U ; [m*keep, (m+n)*keep) -> SVh
  const long long nu = (long long)m * keep, nv = (long long)n * keep;
  const size_t T = (size_t)p + q;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nu + nv;
       e += (long long)gridDim.x * blockDim.x) {
    if (e < nu) {
      if (!u) continue;
      const int j = (int)(e % keep);
      const int i = (int)(e / keep);
      const int c = perm[j];
      const size_t blk = c / BC, c16 = c % BC;
      cplx v;
      if (!transposed) {         // U = Y / sigma
        v = y[(blk * T + i) * BC + c16];
        const double s = sval[j];
        const double inv = (s > 0.0) ? 1.0 / s : 0.0;
        v.x *= inv; v.y *= inv;
      } else {                   // U = W
        v = y[(blk * T + p + i) * BC + c16];
      }
      u[(long long)(i / u_na) * u_so + (long long)(i % u_na) * u_sa + j * u_sj] = v;
    } else {
      if (!svh) continue;
      const long long f = e - nu;
      const int col = (int)(f % n);
      const int j = (int)(f / n);
      const int c = perm[j];
      const size_t blk = c / BC, c16 = c % BC;
      cplx v;
      if (!transposed) {         // S Vh[j, col] = sigma_j * conj(W[col, c])
        v = y[(blk * T + p + col) * BC + c16];
        const double s = sval[j];
        v = make_double2(v.x * s, -v.y * s);
      } else {                   // S Vh[j, col] = conj(Y[col, c])
        v = y[(blk * T + col) * BC + c16];
        v.y = -v.y;
      }
      svh[(long long)j * n + col] = v;
    }
  }
}

constexpr size_t kDynSmem = (size_t)CHUNK_ROWS * PB * sizeof(cplx) + 2 * PB * PB * sizeof(cplx);

}  // namespace

// ============================================================================ C-ABI
extern "C" size_t b200_svd_workspace_bytes(int m, int n) {
  if (m <= 0 || n <= 0) return 0;
  return make_layout(m, n).total;
}

extern "C" int b200_svd_factor(void* stream_, const void* theta, int m, int n,
                               int64_t rs, int64_t cs, double eps, void* work,
                               int32_t* info_host) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!theta || !work || !info_host || m <= 0 || n <= 0) {
    b200::set_error("b200_svd_factor: invalid argument");
    return B200_EINVAL;
  }
  const Layout L = make_layout(m, n);
  if (L.nb * BC > 16384) {
    b200::set_error("b200_svd_factor: min(m,n)=%d exceeds 16384", L.q);
    return B200_ESIZE;
  }
  unsigned char* base = (unsigned char*)work;
  Header* hdr = (Header*)(base + L.header);
  int* ctrl = (int*)(base + L.ctrl);
  double* sig2 = (double*)(base + L.sig2);
  double* sval = (double*)(base + L.sval);
  int* perm = (int*)(base + L.perm);
  cplx* gpart = (cplx*)(base + L.gpart);
  cplx* jbuf = (cplx*)(base + L.jbuf);
  cplx* y = (cplx*)(base + L.y);

  // header + control words are contiguous: one memset clears both (fro2 = 0, counters = 0)
  B200_CUDA_CHECK(cudaMemsetAsync(base, 0, L.ctrl + L.ctrl_bytes, stream));

  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(
        jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem));
    attr_set = true;
  }
  const int ncols = L.nb * BC;
  int npow = 1;
  while (npow < ncols) npow <<= 1;
  if ((size_t)npow * (sizeof(double) + sizeof(int)) > kDynSmem) {
    b200::set_error("b200_svd_factor: sort buffer exceeds shared memory");
    return B200_ESIZE;
  }
  const cplx* th = (const cplx*)theta;
  long long rs_ = rs, cs_ = cs;
  int p = L.p, q = L.q, nb = L.nb, R = L.R, RS = L.RS, SE = L.SE, tr = L.transposed;
  int minmn = (m < n) ? m : n;
  // relative orthogonality target |cos| <= 1e-11 (singular values are second order in
  // it); never tighter than the rounding level of a length-p dot product
  double tol = 2.0 * sqrt((double)L.p) * 2.220446049250313e-16;
  if (tol < 1e-11) tol = 1e-11;
  // columns below 1e-2*eps*||X||_F can never be kept nor change the rank decision
  double neg_rel = (eps > 0.0) ? 1e-2 * eps : 0.0;
  void* args[] = {&th, &rs_, &cs_, &y, &gpart, &jbuf, &ctrl, &sig2, &sval, &perm, &hdr,
                  &info_host, &p, &q, &nb, &R, &RS, &SE, &tr, &minmn, &tol, &eps, &neg_rel};
  const int grid = L.SE * L.R;
  b200::profile_begin(stream);
  B200_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)jacobi_kernel, dim3(grid), dim3(JT),
                                              args, kDynSmem, stream));
  b200::count_launch();
  {   // SURVEY 8d convention: 4*(14 m n^2 + 8 n^3) with m >= n
    const double mm = (double)L.p, nn = (double)L.q;
    b200::profile_end(stream, 4.0 * (14.0 * mm * nn * nn + 8.0 * nn * nn * nn),
                      &hdr->sweeps);
  }
  return B200_OK;
}

extern "C" int b200_svd_emit(void* stream_, const void* work, const void* theta, int m,
                             int n, int64_t rs, int64_t cs, int keep, void* u, int u_na,
                             int64_t u_so, int64_t u_sa, int64_t u_sj, void* svh) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!work || m <= 0 || n <= 0 || keep < 0 || u_na < 1) {
    b200::set_error("b200_svd_emit: invalid argument");
    return B200_EINVAL;
  }
  if (keep == 0) return B200_OK;
  const Layout L = make_layout(m, n);
  const unsigned char* base = (const unsigned char*)work;
  const long long total = (long long)(m + n) * keep;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  (void)theta; (void)rs; (void)cs;
  emit_kernel<<<blocks, 256, 0, stream>>>(
      (const cplx*)(base + L.y), (const double*)(base + L.sval),
      (const int*)(base + L.perm), m, n, L.p, L.q, L.transposed, keep, (cplx*)u, u_na,
      u_so, u_sa, u_sj, (cplx*)svh);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_svd_phase_cycles(void* stream_, const void* work, long long* out8) {
  if (!work || !out8) { b200::set_error("b200_svd_phase_cycles: invalid argument"); return B200_EINVAL; }
  Header h;
  B200_CUDA_CHECK(cudaMemcpyAsync(&h, work, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  B200_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream_));
  for (int k = 0; k < 8; ++k) out8[k] = h.phase_cycles[k];
  return B200_OK;
}

extern "C" int b200_svd_values(void* stream_, const void* work, int m, int n,
                               double* s_out) {
  if (!work || !s_out || m <= 0 || n <= 0) {
    b200::set_error("b200_svd_values: invalid argument");
    return B200_EINVAL;
  }
  const Layout L = make_layout(m, n);
  const int minmn = (m < n) ? m : n;
  B200_CUDA_CHECK(cudaMemcpyAsync(
      s_out, (const unsigned char*)work + L.sval, sizeof(double) * minmn,
      cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
  return B200_OK;
}
