// DRAFT (round-2 work item) -- NOT part of liboqupy_b200.so, NOT
// yet run on a GPU.  Cross-compiles for sm_100a (nvcc -c) so that round 2 starts from code
// that builds; nothing in the product path, the tests or bench.py references this file.
//
// Column-pivoted Householder QR with early termination, as the preconditioner of the block
// Jacobi SVD (replaces nothing in the reference by itself: it is an internal stage of the
// kernel that stands in for tn.split_node_full_svd, oqupy/backends/node_array.py:262,285,541).
// numpy statement of what it must compute: tools/study_precond.py::pipeline().
//
// Layout: a is an m x n column-major work copy of Theta (m >= n).  PHYSICAL columns never
// move: column c belongs to CTA (c mod gridDim.x) for the whole factorisation, the pivot
// order lives in perm[] (position -> physical column).  On exit, for position p < k:
//   a[i, perm[p]], i <  p : R[i, p]            a[p, perm[p]] : R[p, p] (real)
//   a[i, perm[p]], i >  p : Householder vector v_p (v_p[p] = 1 implied), tau[p]
// and for positions p >= k the rows i < k hold R12, the rows i >= k the discarded block R22
// whose squared Frobenius norm is returned in *tail2 (it joins the tail of the rank rule).
//
// One pivot step = two grid barriers:
//   (1) every CTA publishes the largest remaining column norm among ITS columns; barrier;
//       every CTA picks the same pivot from the gridDim.x candidates (lowest physical index
//       on ties: deterministic).  Stop when the pivot norm <= stop (= 1e-5*eps*||X||_F).
//       The owner of the pivot column forms v, tau, R[p,p] and broadcasts v through L2.
//   (2) barrier; every CTA applies H = I - tau v v^H to its own remaining columns (one warp
//       per column: the dot product is a warp reduction, no cross-CTA reduction anywhere),
//       down-dates their norms (LAPACK xGEQP3 safeguard: recompute when cancellation has
//       eaten half the digits) and goes back to (1).
// Critical path: k x (2 barriers + one column pass through L2).  For operands <= 256 x 256
// the same scheme fits one 4-CTA cluster with the columns in distributed shared memory.

#include <cooperative_groups.h>

#include "../common.cuh"

namespace b200 {
namespace draft {

constexpr int QT = 256;  // threads per CTA (8 warps: 8 columns in flight)

__device__ __forceinline__ int ld_acquire_i(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void grid_barrier(int* counter, int& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(counter), "r"(1) : "memory");
    const int target = epoch * (int)gridDim.x;
    while (ld_acquire_i(counter) < target) {
    }
  }
  __syncthreads();
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

struct QrcpArgs {
  cplx* a;          // m x n column-major, in place
  int m, n;
  double stop;      // absolute pivot-norm threshold
  cplx* vbuf;       // 2 x m broadcast buffer (double buffered by step parity)
  double* tau;      // n   (real part of the complex tau is enough: beta is chosen real)
  cplx* tauc;       // n   complex tau
  int* perm;        // n   position -> physical column
  double* vn;       // n   running column norms (by physical column)
  double* vn_ref;   // n   norms at the last exact recomputation
  double* cand_val; // gridDim.x
  int* cand_idx;    // gridDim.x
  int* bar;         // zeroed by the host
  int* k_out;
  double* tail2;    // zeroed by the host; atomically accumulated (fixed order not needed:
                    // it only enters a comparison 5 decades away from its own value)
};

__global__ void __launch_bounds__(QT) qrcp_kernel(QrcpArgs q) {
  extern __shared__ unsigned char smem_raw[];
  int* done = reinterpret_cast<int*>(smem_raw);  // n flags: physical column already a pivot
  __shared__ double red_val[QT / 32];
  __shared__ int red_idx[QT / 32];
  __shared__ int s_pivot;
  __shared__ double s_pnorm;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = QT / 32;
  const int G = gridDim.x, me = blockIdx.x;
  const int m = q.m, n = q.n;
  int epoch = 0;

  for (int c = tid; c < n; c += QT) done[c] = 0;
  // exact initial norms of my columns
  for (int c = me + warp * G; c < n; c += nwarp * G) {
    double s = 0.0;
    for (int i = lane; i < m; i += 32) {
      cplx x = q.a[(size_t)c * m + i];
      s = fma(x.x, x.x, fma(x.y, x.y, s));
    }
    s = warp_sum(s);
    if (lane == 0) {
      q.vn[c] = sqrt(s);
      q.vn_ref[c] = sqrt(s);
    }
  }
  __syncthreads();

  const int kmax = min(m, n);
  int k = 0;
  for (; k < kmax; ++k) {
    // ---- (1) candidates
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int c = me + tid * G; c < n; c += QT * G)
      if (!done[c]) {
        double v = q.vn[c];
        if (v > best || (v == best && c < bidx)) best = v, bidx = c;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) best = ov, bidx = oi;
    }
    if (lane == 0) red_val[warp] = best, red_idx[warp] = bidx;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nwarp; ++w)
        if (red_val[w] > best || (red_val[w] == best && red_idx[w] < bidx))
          best = red_val[w], bidx = red_idx[w];
      q.cand_val[me] = best;
      q.cand_idx[me] = bidx;
    }
    grid_barrier(q.bar, epoch);
    if (warp == 0) {
      double b = -1.0;
      int bi = 0x7fffffff;
      for (int g = lane; g < G; g += 32) {
        double v = __ldcg(q.cand_val + g);
        int vi = __ldcg(q.cand_idx + g);
        if (v > b || (v == b && vi < bi)) b = v, bi = vi;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, b, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > b || (ov == b && oi < bi)) b = ov, bi = oi;
      }
      if (lane == 0) s_pivot = bi, s_pnorm = b;
    }
    __syncthreads();
    const int p = s_pivot;
    if (!(s_pnorm > q.stop)) break;  // every CTA sees the same candidates: uniform exit
    if (tid == 0) done[p] = 1;
    if (me == 0 && tid == 0) q.perm[k] = p;
    cplx* vb = q.vbuf + (size_t)(k & 1) * m;

    // ---- owner: Householder vector of column p, rows k..m-1 (LAPACK zlarfg, beta real)
    if (p % G == me) {
      cplx* col = q.a + (size_t)p * m;
      double s = 0.0;
      for (int i = k + 1 + tid; i < m; i += QT) {
        cplx x = col[i];
        s = fma(x.x, x.x, fma(x.y, x.y, s));
      }
      s = warp_sum(s);
      if (lane == 0) red_val[warp] = s;
      __syncthreads();
      double xnorm2 = 0.0;
      for (int w = 0; w < nwarp; ++w) xnorm2 += red_val[w];
      __syncthreads();
      const cplx alpha = col[k];
      const double an = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
      const double beta = alpha.x >= 0.0 ? -an : an;
      cplx tau = make_double2(0.0, 0.0), scale = make_double2(0.0, 0.0);
      if (xnorm2 > 0.0 || alpha.y != 0.0) {
        tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
        const cplx d = make_double2(alpha.x - beta, alpha.y);  // scale = 1/(alpha - beta)
        const double dn = d.x * d.x + d.y * d.y;
        scale = make_double2(d.x / dn, -d.y / dn);
      }
      for (int i = k + 1 + tid; i < m; i += QT) {
        cplx v = cmul(col[i], scale);
        col[i] = v;
        vb[i] = v;
      }
      if (tid == 0) {
        col[k] = make_double2(xnorm2 > 0.0 || alpha.y != 0.0 ? beta : alpha.x, 0.0);
        vb[k] = make_double2(1.0, 0.0);
        q.tauc[k] = tau;
      }
    }
    grid_barrier(q.bar, epoch);

    // ---- (2) apply H^H = I - conj(tau) v v^H to my remaining columns, one warp each
    const cplx tau = __ldcg(reinterpret_cast<const double2*>(q.tauc + k));
    const cplx ctau = cconj(tau);
    for (int c = me + warp * G; c < n; c += nwarp * G) {
      if (done[c]) continue;
      cplx* col = q.a + (size_t)c * m;
      cplx w = make_double2(0.0, 0.0);
      for (int i = k + lane; i < m; i += 32) {
        cplx v = __ldcg(reinterpret_cast<const double2*>(vb + i));
        w = cfma(cconj(v), col[i], w);
      }
      w.x = warp_sum(w.x);
      w.y = warp_sum(w.y);
      const cplx f = cmul(ctau, w);
      cplx top = make_double2(0.0, 0.0);
      for (int i = k + lane; i < m; i += 32) {
        cplx v = __ldcg(reinterpret_cast<const double2*>(vb + i));
        cplx x = col[i];
        x.x -= f.x * v.x - f.y * v.y;
        x.y -= f.x * v.y + f.y * v.x;
        col[i] = x;
        if (i == k) top = x;
      }
      // norm down-date (xGEQP3): vn^2 <- vn^2 - |r_kc|^2, recompute when unreliable
      top.x = __shfl_sync(0xffffffffu, top.x, 0);
      top.y = __shfl_sync(0xffffffffu, top.y, 0);
      double vn = q.vn[c];
      if (vn > 0.0) {
        const double r = sqrt(top.x * top.x + top.y * top.y) / vn;
        double t = fmax(0.0, (1.0 + r) * (1.0 - r));
        const double ratio = vn / q.vn_ref[c];
        if (t * ratio * ratio <= 1.4901161193847656e-08) {  // sqrt(eps_mach)
          double s = 0.0;
          for (int i = k + 1 + lane; i < m; i += 32) {
            cplx x = col[i];
            s = fma(x.x, x.x, fma(x.y, x.y, s));
          }
          s = warp_sum(s);
          vn = sqrt(s);
          if (lane == 0) q.vn_ref[c] = vn;
        } else {
          vn *= sqrt(t);
        }
        if (lane == 0) q.vn[c] = vn;
      }
    }
    __syncthreads();
  }

  // ---- discarded block: rows >= k of the columns that never became a pivot
  double t2 = 0.0;
  for (int c = me + warp * G; c < n; c += nwarp * G) {
    if (done[c]) continue;
    for (int i = k + lane; i < m; i += 32) {
      cplx x = q.a[(size_t)c * m + i];
      t2 = fma(x.x, x.x, fma(x.y, x.y, t2));
    }
  }
  t2 = warp_sum(t2);
  if (lane == 0 && t2 != 0.0) atomicAdd(q.tail2, t2);
  if (me == 0 && tid == 0) *q.k_out = k;
  // positions k..n-1 of perm: remaining physical columns in index order (CTA 0)
  if (me == 0 && tid == 0) {
    int pos = k;
    for (int c = 0; c < n; ++c)
      if (!done[c]) q.perm[pos++] = c;
  }
}

// U = Q[:, :k] * J for `ncol` columns of J (k x ncol, column-major, ld = k): every CTA takes
// columns of its own (no grid-wide step), holds one m-vector per warp in shared memory and
// applies the reflectors k-1 ... 0.   y <- (I - tau_i v_i v_i^H) y.
__global__ void __launch_bounds__(QT) apply_q_kernel(const cplx* a, int m, int k,
                                                     const int* perm, const cplx* tauc,
                                                     const cplx* j, int ncol, cplx* u,
                                                     int64_t u_ld) {
  extern __shared__ unsigned char smem_raw[];
  cplx* ybase = reinterpret_cast<cplx*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = QT / 32;
  cplx* y = ybase + (size_t)warp * m;
  for (int c = blockIdx.x * nwarp + warp; c < ncol; c += gridDim.x * nwarp) {
    for (int i = lane; i < m; i += 32)
      y[i] = i < k ? j[(size_t)c * k + i] : make_double2(0.0, 0.0);
    __syncwarp();
    for (int r = k - 1; r >= 0; --r) {
      const cplx* v = a + (size_t)perm[r] * m;
      const cplx tau = tauc[r];
      cplx w = make_double2(0.0, 0.0);
      for (int i = r + lane; i < m; i += 32) {
        cplx vi = i == r ? make_double2(1.0, 0.0) : v[i];
        w = cfma(cconj(vi), y[i], w);
      }
      w.x = warp_sum(w.x);
      w.y = warp_sum(w.y);
      const cplx f = cmul(tau, w);
      for (int i = r + lane; i < m; i += 32) {
        cplx vi = i == r ? make_double2(1.0, 0.0) : v[i];
        cplx x = y[i];
        x.x -= f.x * vi.x - f.y * vi.y;
        x.y -= f.x * vi.y + f.y * vi.x;
        y[i] = x;
      }
      __syncwarp();
    }
    for (int i = lane; i < m; i += 32) u[(size_t)c * u_ld + i] = y[i];
    __syncwarp();
  }
}

// L = [R11 R12]^H (n x k, column-major, ld = n) gathered through perm: the Jacobi operand.
__global__ void extract_l_kernel(const cplx* a, int m, int n, int k, const int* perm, cplx* l) {
  const int pos = blockIdx.x;  // row of L = pivot position
  const cplx* col = a + (size_t)perm[pos] * m;
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    l[(size_t)i * n + pos] = i <= pos ? cconj(col[i]) : make_double2(0.0, 0.0);
}

}  // namespace draft
}  // namespace b200
