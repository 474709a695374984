// eps-truncated complex128 SVD on the device: one-sided BLOCK Jacobi.
//
// Replaces tn.split_node_full_svd(max_truncation_err=eps, relative=True)
// (reference call sites oqupy/backends/node_array.py:262,285,541).
//
// Algorithm (B200-first, not LAPACK's bidiagonalisation):
//   * X = theta (m >= n) or theta^H (m < n), p x q with p >= q, stored as PANELS of
//     4 columns: panel j holds X[:, 4j:4j+4] as [row][4] complex128 (64 B per row),
//     so a panel streams through L2/shared memory fully coalesced.  A q x q matrix W
//     (initially I) accumulates the right rotations in the same panel layout.
//   * Hestenes one-sided Jacobi on panel pairs, round-robin tournament ordering:
//     per stage every CTA owns a pair (I, J) = 8 columns, stages them in shared
//     memory, forms their 8x8 Gram matrix on the fp64 tensor cores (DMMA.8x8x4),
//     diagonalises it with a warp-parallel two-sided Jacobi eigensolver, and applies
//     the 8x8 rotation to the 8 columns of X and of W, again with DMMA.
//     grid.sync() separates stages (cooperative persistent kernel).
//   * Convergence: a full sweep in which no pair had |g_ij| > tol*sqrt(g_ii g_jj).
//   * sigma_j = ||y_j||; sort; reference tail-norm rule picks `keep`; emit kernel
//     writes U[:, :keep] and S*Vh[:keep] in the layouts the MPS engine asks for.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;
using b200::dmma884;

namespace {

constexpr int PC = 4;            // columns per panel
constexpr int JT = 512;          // threads of the Jacobi kernel
constexpr int JW = JT / 32;      // warps
constexpr int MAX_SWEEPS = 120;    // <= NFLAGS
constexpr int NFLAGS = 128;
constexpr int FLOOR_GROW_AFTER = 90;
constexpr int INNER_SWEEPS = 1;    // upper bound; the inner solver stops early
constexpr int SMEM_STAGE_LIMIT = 200 * 1024;  // bytes of panel data staged per CTA

struct Header {        // lives at the start of the workspace (device)
  int m, n, p, q, npan, nb, transposed, keep;
  int sweeps, status, rotations, pad;
  double eps, s0, fro2, pad2;
  const void* theta;          // original matrix (reserved)
  long long rs, cs;
  long long phase_cycles[8];  // CTA 0: wait, loadX, gram, eig(+loadW), apply, store, vote, stages
};

struct Layout {
  size_t header, xp, wp, ucont, sig2, sval, perm, flags, ready, pn, total;
  int p, q, npan, nb, transposed;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ Layout make_layout(int m, int n) {
  Layout L;
  L.transposed = (m < n) ? 1 : 0;
  L.p = L.transposed ? n : m;
  L.q = L.transposed ? m : n;
  L.npan = (L.q + PC - 1) / PC;
  L.nb = (L.npan & 1) ? L.npan + 1 : L.npan;   // even number of panels
  if (L.nb < 2) L.nb = 2;
  size_t off = 0;
  L.header = off; off = align256(off + sizeof(Header));
  L.xp = off;   off = align256(off + (size_t)L.nb * L.p * PC * sizeof(cplx));
  // W accumulates the right rotations.  (Forming S*Vh as U^H*theta instead is NOT an
  // option: columns of U are orthogonal only down to the absolute rounding floor, and
  // the projection would amplify that by sigma_0/sigma_j.)
  L.wp = off;   off = align256(off + (size_t)L.nb * L.q * PC * sizeof(cplx));
  L.ucont = off;
  L.sig2 = off; off = align256(off + (size_t)L.nb * PC * sizeof(double));
  L.sval = off; off = align256(off + (size_t)L.nb * PC * sizeof(double));
  L.perm = off; off = align256(off + (size_t)L.nb * PC * sizeof(int));
  L.flags = off; off = align256(off + NFLAGS * sizeof(int));
  L.ready = off; off = align256(off + (size_t)L.nb * sizeof(int));
  L.pn = off; off = align256(off + (size_t)L.nb * sizeof(double));
  L.total = off;
  return L;
}

// ------------------------------------------------------------------ load / init
__global__ void svd_load_kernel(const cplx* __restrict__ theta, long long rs,
                                long long cs, int m, int n, int p, int q, int nb,
                                int transposed, cplx* __restrict__ xp,
                                cplx* __restrict__ wp, double* __restrict__ fro2,
                                double* __restrict__ pn) {
  double local = 0.0;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < nb; i += blockDim.x) pn[i] = 1e300;
  const long long total_x = (long long)nb * p * PC;
  const long long total_w = wp ? (long long)nb * q * PC : 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
       e < total_x + total_w; e += (long long)gridDim.x * blockDim.x) {
    if (e < total_x) {
      const int c4 = (int)(e % PC);
      const long long t = e / PC;
      const int row = (int)(t % p);
      const int pan = (int)(t / p);
      const int col = pan * PC + c4;
      cplx v = make_double2(0.0, 0.0);
      if (col < q) {
        if (!transposed) {
          v = theta[row * rs + col * cs];
        } else {           // X = theta^H : X[row][col] = conj(theta[col][row])
          v = theta[col * rs + row * cs];
          v.y = -v.y;
        }
      }
      xp[e] = v;
      local = fma(v.x, v.x, local);
      local = fma(v.y, v.y, local);
    } else {
      const long long f = e - total_x;
      const int c4 = (int)(f % PC);
      const long long t = f / PC;
      const int row = (int)(t % q);
      const int pan = (int)(t / q);
      const int col = pan * PC + c4;
      wp[f] = make_double2((col == row) ? 1.0 : 0.0, 0.0);
    }
  }
  // ||X||_F^2: the scale of the absolute noise floor used by the convergence test
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
    if (tot != 0.0) atomicAdd(fro2, tot);
  }
}

// ------------------------------------------------------------------ 8x8 helpers
// Shared scratch of the Jacobi kernel (static part).
struct JacobiShared {
  double red[JW][128];   // per-warp partial Gram fragments (re: 0..63, im: 64..127)
  double gr[2][8][8], gi[2][8][8];   // Gram / working Hermitian matrix (double-buffered)
  double vr[2][8][8], vi[2][8][8];   // accumulated eigenvectors (double-buffered)
  double sr[8][8], si[8][8];   // sorted eigenvectors (the 8x8 rotation to apply)
  double al[8], ber[8], bei[8];      // per index: alpha (real), beta (complex)
  int partner[8], rot[8];
  int need;                    // pair needs a rotation
  int skip;                    // both panels negligible: nothing to do this stage
  double pmax[2];              // max column norm^2 left in panel I / J
};

// Gram matrix of the 8 columns [XI | XJ] (each [rows][4]) over `rows` rows.
// Loads of panel data that lives in global memory bypass L1 (ld.global.cg): panels are
// handed from CTA to CTA through L2 with release/acquire flags, and L1 is not coherent.
template <bool GLOBAL>
__device__ __forceinline__ cplx ld_panel(const cplx* p) {
  if (GLOBAL) return __ldcg(reinterpret_cast<const double2*>(p));
  return *p;
}

template <bool GLOBAL>
__device__ void gram8(const cplx* XI, const cplx* XJ, int rows, JacobiShared& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const cplx* src = (g < 4) ? (XI + g) : (XJ + (g - 4));
  double r0 = 0.0, r1 = 0.0, i0 = 0.0, i1 = 0.0;
  for (int base = warp * 4; base < rows; base += JW * 4) {
    const int row = base + t;
    cplx x = make_double2(0.0, 0.0);
    if (row < rows) x = ld_panel<GLOBAL>(src + (size_t)row * PC);
    // G = X^H X :  Gr = Xr^T Xr + Xi^T Xi ;  Gi = Xr^T Xi - Xi^T Xr
    dmma884(r0, r1, x.x, x.x);
    dmma884(r0, r1, x.y, x.y);
    dmma884(i0, i1, x.x, x.y);
    dmma884(i0, i1, -x.y, x.x);
  }
  S.red[warp][g * 8 + 2 * t] = r0;
  S.red[warp][g * 8 + 2 * t + 1] = r1;
  S.red[warp][64 + g * 8 + 2 * t] = i0;
  S.red[warp][64 + g * 8 + 2 * t + 1] = i1;
  __syncthreads();
  if (threadIdx.x < 128) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < JW; ++w) acc += S.red[w][threadIdx.x];
    const int e = threadIdx.x & 63;
    if (threadIdx.x < 64) S.gr[0][e >> 3][e & 7] = acc;
    else S.gi[0][e >> 3][e & 7] = acc;
  }
  __syncthreads();
}

// Warp 0: decide whether the pair needs work; if so diagonalise the 8x8 Hermitian
// Gram matrix (cyclic two-sided Jacobi, 4 disjoint rotations per round) and leave
// the eigenvector matrix, columns sorted by DESCENDING eigenvalue, in S.sr/S.si.
// A pair (i, j) counts as orthogonal when
//   |g_ij| <= sqrt(max(a,b)) * (tol*sqrt(min(a,b)) + floor),  floor = kappa*eps*||X||_F:
// the accumulated 8x8 rotations are accurate in the ABSOLUTE sense (eps*||X||), so
// columns that have sunk to the rounding floor are left alone.
// (threshold combined in quadrature: no sqrt on the critical path.)  Pairs whose two
// columns are both NEGLIGIBLE (norm^2 < neg2 = (1e-2*eps*||X||_F)^2) are only
// orthogonalised loosely: such columns are discarded by the truncation rule whatever
// their mutual angles, and the tail norm only needs the Frobenius norm of their block,
// which is rotation invariant.  (neg2 = 0 when no truncation is requested.)
__device__ __forceinline__ bool pair_converged(double a, double b, double g2,
                                               double tol2, double floor2, double neg2) {
  const double big = fmax(a, b), small = fmin(a, b);
  if (big <= 0.0) return true;
  // negligible-negligible: only a loose |cos| <= 1e-2 (keeps the cleaning of the
  // relevant columns against them a contraction)
  if (big < neg2) return g2 <= 1e-4 * big * fmax(small, 0.0) + big * floor2;
  return g2 <= big * (tol2 * fmax(small, 0.0) + floor2);
}

__device__ void eig8_warp0(JacobiShared& S, double tol2, double floor2, double neg2) {
  const int lane = threadIdx.x;   // caller guarantees threadIdx.x < 32
  // --- convergence test on the raw Gram matrix
  double worst = 0.0;
  for (int e = lane; e < 64; e += 32) {
    const int i = e >> 3, j = e & 7;
    if (i < j) {
      const double a = S.gr[0][i][i], b = S.gr[0][j][j];
      const double g2 = S.gr[0][i][j] * S.gr[0][i][j] + S.gi[0][i][j] * S.gi[0][i][j];
      if (!pair_converged(a, b, g2, tol2, floor2, neg2)) worst = 1.0;
    }
  }
  const unsigned any = __ballot_sync(0xffffffffu, worst > 0.0);
  if (lane == 0) S.need = any ? 1 : 0;
  if (!any) {
    if (lane < 2) {
      const int o = 4 * lane;
      S.pmax[lane] = fmax(fmax(S.gr[0][o][o], S.gr[0][o + 1][o + 1]),
                          fmax(S.gr[0][o + 2][o + 2], S.gr[0][o + 3][o + 3]));
    }
    __syncwarp();
    return;
  }

  // exact Hermitian symmetry + V = I   (buffer 0)
  for (int e = lane; e < 64; e += 32) {
    const int i = e >> 3, j = e & 7;
    S.vr[0][i][j] = (i == j) ? 1.0 : 0.0;
    S.vi[0][i][j] = 0.0;
  }
  __syncwarp();
  for (int e = lane; e < 64; e += 32) {
    const int i = e >> 3, j = e & 7;
    if (i > j) { S.gr[0][i][j] = S.gr[0][j][i]; S.gi[0][i][j] = -S.gi[0][j][i]; }
    if (i == j) S.gi[0][i][j] = 0.0;
  }
  __syncwarp();

  // Cyclic two-sided Jacobi, 4 disjoint rotations per round, all 64 entries of
  // G' = R^H G R and V' = V R recomputed in ONE pass from the previous buffer
  // (double buffering: two __syncwarp per round).  With pi(x) the partner of index x:
  //   column op  T(x,y)  = al_y G[x][y] + be_y G[x][pi y]
  //   row op     G'[x][y] = al_x T(x,y) + conj(be_x) T(pi x, y)
  // where (al, be) = (c, -conj(se)) for the first index of a pair, (c, se) for the
  // second, R = [[c, se], [-conj(se), c]].
  int cur = 0;
  for (int sweep = 0; sweep < INNER_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int round = 0; round < 7; ++round) {
      if (lane < 4) {
        const int ka = lane, kb = 7 - lane;
        int i = (ka == 0) ? 0 : 1 + ((ka - 1 + round) % 7);
        int j = 1 + ((kb - 1 + round) % 7);
        if (i > j) { const int t = i; i = j; j = t; }
        const double a = S.gr[cur][i][i], b = S.gr[cur][j][j];
        const double xr = S.gr[cur][i][j], xi = S.gi[cur][i][j];
        const double mag2 = xr * xr + xi * xi;
        double c = 1.0, sr_ = 0.0, si_ = 0.0;
        int rot = 0;
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
          rot = 1;
          rotated = 1;
        }
        S.partner[i] = j; S.partner[j] = i;
        S.al[i] = c; S.al[j] = c;
        S.ber[i] = -sr_; S.bei[i] = si_;     // -conj(se)
        S.ber[j] = sr_;  S.bei[j] = si_;     //  se
        S.rot[i] = rot; S.rot[j] = rot;
      }
      __syncwarp();
      const int nxt = cur ^ 1;
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const int e = lane + 32 * rep;
        const int x = e >> 3, y = e & 7;
        const int px = S.partner[x], py = S.partner[y];
        const double aly = S.al[y], byr = S.ber[y], byi = S.bei[y];
        const double alx = S.al[x], bxr = S.ber[x], bxi = -S.bei[x];   // conj(be_x)
        // T(x,y)
        const double g00r = S.gr[cur][x][y], g00i = S.gi[cur][x][y];
        const double g01r = S.gr[cur][x][py], g01i = S.gi[cur][x][py];
        const double t0r = aly * g00r + (byr * g01r - byi * g01i);
        const double t0i = aly * g00i + (byr * g01i + byi * g01r);
        // T(pi x, y)
        const double g10r = S.gr[cur][px][y], g10i = S.gi[cur][px][y];
        const double g11r = S.gr[cur][px][py], g11i = S.gi[cur][px][py];
        const double t1r = aly * g10r + (byr * g11r - byi * g11i);
        const double t1i = aly * g10i + (byr * g11i + byi * g11r);
        double nr = alx * t0r + (bxr * t1r - bxi * t1i);
        double ni = alx * t0i + (bxr * t1i + bxi * t1r);
        if (x == y) ni = 0.0;
        if (y == px && S.rot[x]) { nr = 0.0; ni = 0.0; }   // annihilated entry
        S.gr[nxt][x][y] = nr;
        S.gi[nxt][x][y] = ni;
        // V' = V R
        const double v0r = S.vr[cur][x][y], v0i = S.vi[cur][x][y];
        const double v1r = S.vr[cur][x][py], v1i = S.vi[cur][x][py];
        S.vr[nxt][x][y] = aly * v0r + (byr * v1r - byi * v1i);
        S.vi[nxt][x][y] = aly * v0i + (byr * v1i + byi * v1r);
      }
      __syncwarp();
      cur = nxt;
    }
    if (__ballot_sync(0xffffffffu, rotated != 0) == 0u) break;
  }
  // --- sort eigenvectors by descending eigenvalue
  if (lane < 8) {
    const double lam = S.gr[cur][lane][lane];
    int rank = 0;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const double lo = S.gr[cur][o][o];
      if (lo > lam || (lo == lam && o < lane)) ++rank;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      S.sr[r][rank] = S.vr[cur][r][lane];
      S.si[r][rank] = S.vi[cur][r][lane];
    }
    if (rank == 0) S.pmax[0] = lam;     // largest -> panel I
    if (rank == 4) S.pmax[1] = lam;     // fifth largest -> panel J
  }
  __syncwarp();
}

// [XI | XJ] <- [XI | XJ] * R  over `rows` rows, R = S.sr + i S.si (8x8), in place.
template <bool GLOBAL>
__device__ void apply8(cplx* XI, cplx* XJ, int rows, const JacobiShared& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  // B fragments: B[k][col] with k = kstep*4 + t, col = g
  const double b0r = S.sr[t][g], b0i = S.si[t][g];
  const double b1r = S.sr[4 + t][g], b1i = S.si[4 + t][g];
  for (int base = warp * 8; base < rows; base += JW * 8) {
    const int row = base + g;
    cplx a0 = make_double2(0.0, 0.0), a1 = a0;
    if (row < rows) {
      a0 = ld_panel<GLOBAL>(XI + (size_t)row * PC + t);
      a1 = ld_panel<GLOBAL>(XJ + (size_t)row * PC + t);
    }
    double dr0 = 0.0, dr1 = 0.0, di0 = 0.0, di1 = 0.0;
    dmma884(dr0, dr1, a0.x, b0r);
    dmma884(dr0, dr1, -a0.y, b0i);
    dmma884(di0, di1, a0.x, b0i);
    dmma884(di0, di1, a0.y, b0r);
    dmma884(dr0, dr1, a1.x, b1r);
    dmma884(dr0, dr1, -a1.y, b1i);
    dmma884(di0, di1, a1.x, b1i);
    dmma884(di0, di1, a1.y, b1r);
    __syncwarp();
    if (row < rows) {
      // D[g][2t], D[g][2t+1]: columns 0..3 -> XI, 4..7 -> XJ
      cplx* dst = (t < 2) ? (XI + (size_t)row * PC + 2 * t)
                          : (XJ + (size_t)row * PC + 2 * (t - 2));
      dst[0] = make_double2(dr0, di0);
      dst[1] = make_double2(dr1, di1);
    }
  }
}

// global -> shared (L2 loads) by the threads [t0, t0+nt) of the CTA
__device__ __forceinline__ void load_panel(cplx* dst, const cplx* src, int rows, int t0,
                                           int nt) {
  const int total = rows * PC;
  for (int e = (int)threadIdx.x - t0; e < total; e += nt)
    if (e >= 0) dst[e] = __ldcg(reinterpret_cast<const double2*>(src) + e);
}
__device__ __forceinline__ void store_panel(cplx* dst, const cplx* src, int rows) {
  const int total = rows * PC;
  for (int e = threadIdx.x; e < total; e += JT) dst[e] = src[e];
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------ Jacobi kernel
// Persistent cooperative kernel.  Stage-to-stage hand-over of panels is point to point:
// ready[panel] counts the stages completed on that panel (release/acquire through L2),
// so a CTA only waits for the two CTAs that produced its inputs, not for the grid.
// One grid.sync per SWEEP carries the convergence vote.
__global__ void __launch_bounds__(JT, 1)
jacobi_kernel(cplx* __restrict__ xp, cplx* __restrict__ wp, int p, int q, int nb,
              int stage_x, int stage_w, double tol, double neg_rel,
              int* __restrict__ flags, int* __restrict__ ready,
              double* __restrict__ pn, Header* __restrict__ hdr) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ JacobiShared S;
  cg::grid_group grid = cg::this_grid();

  cplx* sXI = reinterpret_cast<cplx*>(dyn_smem);
  cplx* sXJ = sXI + (stage_x ? (size_t)p * PC : 0);
  cplx* sWI = sXJ + (stage_x ? (size_t)p * PC : 0);
  cplx* sWJ = sWI + (stage_w ? (size_t)q * PC : 0);
  const int npairs = nb / 2;
  const size_t xpan = (size_t)p * PC, wpan = (size_t)q * PC;

  int sweeps_done = 0;
  int total_rot = 0;
  int status = 1;   // 1 = not converged
  const double fro = sqrt(hdr->fro2);
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tq = clock64();
#define PHASE(k) { const long long tn_ = clock64(); pc[k] += tn_ - tq; tq = tn_; }
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    int my_rot = 0;
    // absolute floor: 8 eps ||X||_F, doubled every sweep after FLOOR_GROW_AFTER so that the
    // iteration always terminates
    double kappa = 8.0;
    for (int k = FLOOR_GROW_AFTER; k < sweep; ++k) kappa *= 2.0;
    const double floor_ = kappa * 2.220446049250313e-16 * fro;
    const double floor2 = floor_ * floor_, tol2 = tol * tol;
    const double neg2 = (neg_rel * fro) * (neg_rel * fro);
    for (int stage = 0; stage < nb - 1; ++stage) {
      const int gstage = sweep * (nb - 1) + stage;
      for (int idx = blockIdx.x; idx < npairs; idx += gridDim.x) {
        const int ka = idx, kb = nb - 1 - idx;
        int pa = (ka == 0) ? 0 : 1 + ((ka - 1 + stage) % (nb - 1));
        int pb = 1 + ((kb - 1 + stage) % (nb - 1));
        if (pa > pb) { const int t = pa; pa = pb; pb = t; }
        cplx* gXI = xp + pa * xpan;
        cplx* gXJ = xp + pb * xpan;
        cplx* gWI = wp + pa * wpan;
        cplx* gWJ = wp + pb * wpan;
        // wait until both input panels have finished the previous stage
        if (threadIdx.x == 0) {
          while (ld_acquire(ready + pa) < gstage) {}
          while (ld_acquire(ready + pb) < gstage) {}
          // both panels already below the negligible level: skip the stage for them
          S.skip = 0;   // (panel-level skipping disabled: see pair_converged)
        }
        __syncthreads();
        PHASE(0)
        if (S.skip) {
          __syncthreads();
          if (threadIdx.x == 0) {
            st_release(ready + pa, gstage + 1);
            st_release(ready + pb, gstage + 1);
          }
          ++pc[7];
          continue;
        }
        if (stage_x) {
          load_panel(sXI, gXI, p, 0, JT);
          load_panel(sXJ, gXJ, p, 0, JT);
          __syncthreads();
          PHASE(1)
          gram8<false>(sXI, sXJ, p, S);
        } else {
          gram8<true>(gXI, gXJ, p, S);
        }
        PHASE(2)
        if (threadIdx.x < 32) {
          eig8_warp0(S, tol2, floor2, neg2);
        } else if (stage_w) {      // overlap: the other warps fetch the W panels
          load_panel(sWI, gWI, q, 32, JT - 32);
          load_panel(sWJ, gWJ, q, 32, JT - 32);
        }
        __syncthreads();
        PHASE(3)
        if (S.need) {
          ++my_rot;
          if (stage_x) apply8<false>(sXI, sXJ, p, S); else apply8<true>(gXI, gXJ, p, S);
          if (stage_w) apply8<false>(sWI, sWJ, q, S); else apply8<true>(gWI, gWJ, q, S);
          __syncthreads();
          PHASE(4)
          if (stage_x) { store_panel(gXI, sXI, p); store_panel(gXJ, sXJ, p); }
          if (stage_w) { store_panel(gWI, sWI, q); store_panel(gWJ, sWJ, q); }
        }
        __syncthreads();
        if (threadIdx.x == 0) {   // release is cumulative over the CTA barrier above
          pn[pa] = S.pmax[0];
          pn[pb] = S.pmax[1];
          st_release(ready + pa, gstage + 1);
          st_release(ready + pb, gstage + 1);
        }
        PHASE(5)
        ++pc[7];
      }
    }
    // convergence vote: flags[sweep] counts pairs rotated in this sweep
    if (threadIdx.x == 0 && my_rot) atomicAdd(&flags[sweep], my_rot);
    grid.sync();
    const int rot = *((volatile int*)&flags[sweep]);
    PHASE(6)
    total_rot += rot;
    sweeps_done = sweep + 1;
    if (rot == 0) { status = 0; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->sweeps = sweeps_done;
    hdr->status = status;
    hdr->rotations = total_rot;
    for (int k = 0; k < 8; ++k) hdr->phase_cycles[k] = pc[k];
  }
#undef PHASE
}

// ------------------------------------------------------------------ finalize
// sigma_j^2 = ||X[:, j]||^2 ; one warp per panel.
__global__ void colnorm_kernel(const cplx* __restrict__ xp, int p, int nb,
                               double* __restrict__ sig2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nb) return;
  const cplx* X = xp + (size_t)warp * p * PC;
  double acc[PC] = {0.0, 0.0, 0.0, 0.0};
  for (int row = lane; row < p; row += 32) {
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      const cplx v = X[(size_t)row * PC + c];
      acc[c] = fma(v.x, v.x, acc[c]);
      acc[c] = fma(v.y, v.y, acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < PC; ++c) {
    double v = acc[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sig2[warp * PC + c] = v;
  }
}

// Sort (descending), apply the tail-norm rule, publish keep.  One CTA.
__global__ void __launch_bounds__(1024)
rank_kernel(const double* __restrict__ sig2, int ncols, int q, int minmn,
            double eps, double* __restrict__ sval, int* __restrict__ perm,
            Header* __restrict__ hdr, int32_t* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  int npow = 1;
  while (npow < ncols) npow <<= 1;
  double* key = reinterpret_cast<double*>(dyn_smem);
  int* idx = reinterpret_cast<int*>(key + npow);
  for (int e = threadIdx.x; e < npow; e += blockDim.x) {
    // padded / dummy columns sort to the end
    key[e] = (e < ncols) ? sig2[e] : -1.0;   // zero padding columns sort last
    idx[e] = e;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = threadIdx.x; e < npow; e += blockDim.x) {
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
  for (int e = threadIdx.x; e < q; e += blockDim.x) {
    sval[e] = sqrt(fmax(key[e], 0.0));
    perm[e] = idx[e];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
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
    hdr->keep = keep;
    hdr->s0 = s0;
    info[0] = keep;
    info[1] = hdr->sweeps;
    info[2] = hdr->status;
    info[3] = hdr->rotations;
  }
}

// ------------------------------------------------------------------ emit
__global__ void emit_kernel(const cplx* __restrict__ xp, const cplx* __restrict__ wp,
                            const double* __restrict__ sval,
                            const int* __restrict__ perm, int m, int n, int p, int q,
                            int transposed, int keep, cplx* __restrict__ u, int u_na,
                            long long u_so, long long u_sa, long long u_sj,
                            cplx* __restrict__ svh, cplx* __restrict__ ucont) {
  // element space: [0, m*keep) -> U ; [m*keep, (m+n)*keep) -> SVh
  const long long nu = (long long)m * keep, nv = (long long)n * keep;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nu + nv;
       e += (long long)gridDim.x * blockDim.x) {
    if (e < nu) {
      if (!u && !ucont) continue;
      const int j = (int)(e % keep);
      const int i = (int)(e / keep);
      const int c = perm[j];
      const size_t pan = c / PC, c4 = c % PC;
      cplx v;
      if (!transposed) {         // U = Y / sigma
        v = xp[(pan * p + i) * PC + c4];
        const double s = sval[j];
        const double inv = (s > 0.0) ? 1.0 / s : 0.0;
        v.x *= inv; v.y *= inv;
      } else {                   // U = W
        v = wp[(pan * q + i) * PC + c4];
      }
      if (u) u[(long long)(i / u_na) * u_so + (long long)(i % u_na) * u_sa + j * u_sj] = v;
      if (ucont) ucont[(long long)i * keep + j] = v;
    } else {
      if (!svh) continue;
      const long long f = e - nu;
      const int col = (int)(f % n);
      const int j = (int)(f / n);
      const int c = perm[j];
      const size_t pan = c / PC, c4 = c % PC;
      cplx v;
      if (!transposed) {         // S Vh[j, col] = sigma_j * conj(W[col, c])
        v = wp[(pan * q + col) * PC + c4];
        const double s = sval[j];
        v = make_double2(v.x * s, -v.y * s);
      } else {                   // S Vh[j, col] = conj(Y[col, c])
        v = xp[(pan * p + col) * PC + c4];
        v.y = -v.y;
      }
      svh[(long long)j * n + col] = v;
    }
  }
}

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
  if (L.nb * PC > 16384) {
    b200::set_error("b200_svd_factor: min(m,n)=%d exceeds 16384", L.q);
    return B200_ESIZE;
  }
  unsigned char* base = (unsigned char*)work;
  Header* hdr = (Header*)(base + L.header);
  cplx* xp = (cplx*)(base + L.xp);
  cplx* wp = (cplx*)(base + L.wp);
  double* sig2 = (double*)(base + L.sig2);
  double* sval = (double*)(base + L.sval);
  int* perm = (int*)(base + L.perm);
  int* flags = (int*)(base + L.flags);
  int* ready = (int*)(base + L.ready);
  double* pn = (double*)(base + L.pn);

  Header h;
  h.m = m; h.n = n; h.p = L.p; h.q = L.q; h.npan = L.npan; h.nb = L.nb;
  h.transposed = L.transposed; h.keep = 0; h.sweeps = 0; h.status = 1;
  h.rotations = 0; h.pad = 0; h.eps = eps; h.s0 = 0.0; h.fro2 = 0.0; h.pad2 = 0.0;
  h.theta = theta; h.rs = rs; h.cs = cs;
  B200_CUDA_CHECK(cudaMemcpyAsync(hdr, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
  B200_CUDA_CHECK(cudaMemsetAsync(flags, 0, NFLAGS * sizeof(int), stream));
  B200_CUDA_CHECK(cudaMemsetAsync(ready, 0, (size_t)L.nb * sizeof(int), stream));

  {
    const long long total = (long long)L.nb * (L.p + (wp ? L.q : 0)) * PC;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    svd_load_kernel<<<blocks, 256, 0, stream>>>((const cplx*)theta, rs, cs, m, n,
                                                L.p, L.q, L.nb, L.transposed, xp, wp,
                                                &hdr->fro2, pn);
    B200_LAUNCH_CHECK();
  }
  {
    static int dev_sms = 0;
    if (!dev_sms) {
      int dev = 0;
      B200_CUDA_CHECK(cudaGetDevice(&dev));
      B200_CUDA_CHECK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev));
      B200_CUDA_CHECK(cudaFuncSetAttribute(
          jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
          SMEM_STAGE_LIMIT));
    }
    const size_t x_bytes = (size_t)2 * L.p * PC * sizeof(cplx);
    const size_t w_bytes = (size_t)2 * L.q * PC * sizeof(cplx);
    int stage_x = x_bytes <= (size_t)SMEM_STAGE_LIMIT ? 1 : 0;
    int stage_w = (stage_x && x_bytes + w_bytes <= (size_t)SMEM_STAGE_LIMIT) ? 1 : 0;
    size_t dyn = (stage_x ? x_bytes : 0) + (stage_w ? w_bytes : 0);
    int per_sm = 0;
    B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, jacobi_kernel, JT, dyn));
    if (per_sm < 1) {
      b200::set_error("b200_svd_factor: Jacobi kernel cannot be resident (dyn smem %zu)", dyn);
      return B200_ECUDA;
    }
    int grid = L.nb / 2;
    const int cap = per_sm * dev_sms;
    if (grid > cap) grid = cap;
    int p = L.p, q = L.q, nb = L.nb;
    // relative orthogonality target |cos| <= 1e-11 (singular values are second order in
    // it); never tighter than the rounding level of a length-p dot product
    double tol = 2.0 * sqrt((double)L.p) * 2.220446049250313e-16;
    if (tol < 1e-11) tol = 1e-11;
    // columns below 1e-2*eps*||X||_F can never be kept nor change the rank decision
    double neg_rel = (eps > 0.0) ? 1e-2 * eps : 0.0;
    void* args[] = {&xp, &wp, &p, &q, &nb, &stage_x, &stage_w, &tol, &neg_rel,
                    &flags, &ready, &pn, &hdr};
    b200::profile_begin(stream);
    B200_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)jacobi_kernel, dim3(grid),
                                                dim3(JT), args, dyn, stream));
    b200::count_launch();
    {   // SURVEY 8d convention: 4*(14 m n^2 + 8 n^3) with m >= n
      const double mm = (double)L.p, nn = (double)L.q;
      b200::profile_end(stream, 4.0 * (14.0 * mm * nn * nn + 8.0 * nn * nn * nn),
                        &hdr->sweeps);
    }
  }
  {
    const int warps = L.nb;
    const int blocks = (warps * 32 + 255) / 256;
    colnorm_kernel<<<blocks, 256, 0, stream>>>(xp, L.p, L.nb, sig2);
    B200_LAUNCH_CHECK();
  }
  {
    static bool attr_set = false;
    if (!attr_set) {
      B200_CUDA_CHECK(cudaFuncSetAttribute(
          rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 12));
      attr_set = true;
    }
    const int ncols = L.nb * PC;
    int npow = 1;
    while (npow < ncols) npow <<= 1;
    const size_t dyn = (size_t)npow * (sizeof(double) + sizeof(int));
    const int minmn = (m < n) ? m : n;
    // info is written to device-visible pinned host memory directly
    rank_kernel<<<1, 1024, dyn, stream>>>(sig2, ncols, L.q, minmn, eps, sval, perm,
                                          hdr, info_host);
    B200_LAUNCH_CHECK();
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
  cplx* ucont = nullptr;
  (void)theta; (void)rs; (void)cs;
  emit_kernel<<<blocks, 256, 0, stream>>>(
      (const cplx*)(base + L.xp), (const cplx*)(base + L.wp),
      (const double*)(base + L.sval), (const int*)(base + L.perm), m, n, L.p, L.q,
      L.transposed, keep, (cplx*)u, u_na, u_so, u_sa, u_sj, (cplx*)svh, ucont);
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
