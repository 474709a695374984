// Rank-revealing front end of the truncated SVD: column-pivoted Householder QR that STOPS at
// the deflation level, and the back-transformation by the stored reflectors.
//
//   X P = Q [R11 R12; 0 R22],  stop at the first pivot whose column norm is <= stop
//   (= 1e-5 * eps * ||X||_F, the level below which the Jacobi iteration already drops columns).
//   The block-Jacobi kernel then runs on the k columns of L = [R11 R12]^H (q x k) instead of
//   the q columns of X; ||R22||_F^2 joins the tail norm of the rank rule.  With L J = Y
//   (orthogonal columns of norm sigma_j):
//       X = Q [J; 0] Y^H P^T      =>   left factor Q[:, :k] J  (apply_q_kernel),
//                                      sigma_j * right factor^H = conj(Y) scattered through P.
//
// Internal stage of what replaces tn.split_node_full_svd (reference call sites
// oqupy/backends/node_array.py:262,285,541); numpy statement of the whole pipeline:
// tools/study_precond.py::pipeline().
//
// Parallel layout (B200): physical column c lives with CTA (c mod G) for the whole
// factorisation -- in SHARED MEMORY when the operand fits (148 x ~200 KB: up to ~1350 x 1350),
// the overflow in L2.  The pivot order lives in perm[].  ONE grid-wide hand-shake per pivot:
// every CTA speculatively publishes its best remaining column together with its norm; the
// record's release-store is the arrival flag, so that after one poll over the G records every
// CTA knows the pivot AND already has its data in L2.  All CTAs form the (bit-identical)
// reflector redundantly and apply it to their own columns; the new column norms are
// accumulated exactly in the same pass (no down-dating, no drift).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace b200 {
namespace qr {

constexpr int QT = 512;               // threads of the QRCP kernel
constexpr int QW = QT / 32;
constexpr int QR_SMEM_BYTES = 226 * 1024;  // dynamic shared memory requested for the kernel
constexpr int QR_MAX_P = 8192;        // rows (apply_q keeps p / threads <= 8 rows per thread)

struct QrHeader {
  int k, status;
  double fro2, stop2;
};

struct QrLayout {
  size_t header, cand_val, cand_tag, fro_part, tail_part, ctrl_bytes;   // zeroed region first
  size_t perm, tau, cbuf, a, jac, total;
  int p, q, G, NCmax, nc_res, transposed;
  size_t smem;
};

__host__ __device__ inline size_t qalign(size_t x) { return (x + 255) & ~(size_t)255; }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ cplx ldcg_c(const cplx* p) {
  return __ldcg(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

struct QrArgs {
  const cplx* theta;
  long long rs, cs, rsi, csi;
  int rin, cin, transposed;
  int p, q;
  cplx* a;                    // p x q column-major work copy / result (lda = p)
  double* cand_val;           // [2][G]
  unsigned long long* cand_tag;  // [2][G]: (epoch << 32) | physical column
  double* fro_part;           // [G]
  double* tail_part;          // [G]
  cplx* cbuf;                 // [2][G][p] speculatively published candidate columns
  int* perm;                  // [q] position -> physical column
  cplx* tauc;                 // [q]
  QrHeader* hdr;
  int32_t* info_host;         // pinned; info_host[4] <- k
  double stop_rel;            // stop = stop_rel * ||X||_F
  int nc_res;                 // local columns resident in shared memory
  int ncmax;
};

// theta element (i, j) of the ORIGINAL m x n operand
__device__ __forceinline__ cplx theta_at(const QrArgs& A, int i, int j) {
  return A.theta[(long long)(i / A.rin) * A.rs + (long long)(i % A.rin) * A.rsi +
                 (long long)(j / A.cin) * A.cs + (long long)(j % A.cin) * A.csi];
}

__global__ void __launch_bounds__(QT, 1) qrcp_kernel(const QrArgs A) {
  extern __shared__ __align__(16) unsigned char qsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, me = blockIdx.x;
  const int p = A.p, q = A.q;
  const int NC = (q - me + G - 1) / G;          // my local columns: c = me + lc * G
  cplx* vb = reinterpret_cast<cplx*>(qsm);                       // [p] reflector
  cplx* scols = vb + p;                                          // [nc_res][p]
  double* vn2 = reinterpret_cast<double*>(scols + (size_t)A.nc_res * p);   // [ncmax]
  unsigned char* done = reinterpret_cast<unsigned char*>(vn2 + A.ncmax);   // [q]
  __shared__ cplx s_dot[QW];
  __shared__ double s_nrm[QW];
  __shared__ double s_red[QW];
  __shared__ int s_bl, s_widx, s_wcta;
  __shared__ double s_best, s_wval, s_stop2;

  auto colptr = [&](int lc) -> cplx* {
    return (lc < A.nc_res) ? scols + (size_t)lc * p : A.a + (size_t)(me + lc * G) * p;
  };

  for (int c = tid; c < q; c += QT) done[c] = 0;
  // ---- load my columns (X = theta or theta^H), exact norms
  double my_fro = 0.0;
  for (int lc = warp; lc < NC; lc += QW) {
    const int c = me + lc * G;
    cplx* col = colptr(lc);
    double s = 0.0;
    for (int i = lane; i < p; i += 32) {
      cplx v = A.transposed ? theta_at(A, c, i) : theta_at(A, i, c);
      if (A.transposed) v.y = -v.y;
      col[i] = v;
      s = fma(v.x, v.x, s);
      s = fma(v.y, v.y, s);
    }
    s = warp_sum(s);
    if (lane == 0) vn2[lc] = s;
  }
  __syncthreads();
  if (tid == 0) {
    for (int lc = 0; lc < NC; ++lc) my_fro += vn2[lc];     // fixed order
    A.fro_part[me] = my_fro;
    s_stop2 = 0.0;
  }

  // warps per column in the apply pass
  int W = 1;
  if (NC <= 1) W = 16; else if (NC <= 2) W = 8; else if (NC <= 4) W = 4; else if (NC <= 8) W = 2;
  const int slots = QW / W;
  const int my_slot = warp / W, my_part = warp % W;

  int j = 0;
  for (; j < q; ++j) {
    const int par = j & 1;
    const unsigned epoch = (unsigned)(j + 1);
    // ---- A. my best remaining column
    if (warp == 0) {
      double best = -1.0;
      int bl = 0x7fffffff;
      for (int lc = lane; lc < NC; lc += 32)
        if (!done[me + lc * G]) {
          const double v = vn2[lc];
          if (v > best || (v == best && lc < bl)) { best = v; bl = lc; }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (ov > best || (ov == best && ol < bl)) { best = ov; bl = ol; }
      }
      if (lane == 0) { s_bl = (best >= 0.0) ? bl : -1; s_best = best; }
    }
    __syncthreads();
    // ---- B. publish it (rows j..p-1) and announce: the release-store of the tag is the
    //         arrival flag of this CTA for pivot j
    const int bl = s_bl;
    if (bl >= 0) {
      const cplx* col = colptr(bl);
      cplx* dst = A.cbuf + ((size_t)par * G + me) * p;
      for (int i = j + tid; i < p; i += QT) dst[i] = col[i];
    }
    __syncthreads();
    if (tid == 0) {
      A.cand_val[par * G + me] = s_best;
      __threadfence();
      const unsigned idx = (bl >= 0) ? (unsigned)(me + bl * G) : 0x7fffffffu;
      st_release_u64(A.cand_tag + par * G + me, ((unsigned long long)epoch << 32) | idx);
    }
    // ---- C. poll all G records, pick the pivot (largest norm, lowest column on ties)
    if (warp == 0) {
      double b = -2.0;
      unsigned bi = 0xffffffffu;
      int bc = 0;
      for (int g0 = 0; g0 < G; g0 += 128) {
        unsigned long long tag[4];
        bool ok;
        do {
          ok = true;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int g = g0 + u * 32 + lane;
            tag[u] = (g < G) ? ld_acquire_u64(A.cand_tag + par * G + g)
                             : ((unsigned long long)epoch << 32);
            ok = ok && ((unsigned)(tag[u] >> 32) == epoch);
          }
        } while (!ok);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int g = g0 + u * 32 + lane;
          if (g < G) {
            const double v = __ldcg(A.cand_val + par * G + g);
            const unsigned vi = (unsigned)(tag[u] & 0xffffffffu);
            if (v > b || (v == b && vi < bi)) { b = v; bi = vi; bc = g; }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, b, o);
        const unsigned oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (ov > b || (ov == b && oi < bi)) { b = ov; bi = oi; bc = oc; }
      }
      if (j == 0) {             // ||X||_F^2: every CTA sums the G partials in the same order
        double f = 0.0;
        for (int g = lane; g < G; g += 32) f += __ldcg(A.fro_part + g);
        f = warp_sum(f);
        if (lane == 0) {
          s_stop2 = (A.stop_rel * A.stop_rel) * f;
          if (me == 0) { A.hdr->fro2 = f; A.hdr->stop2 = s_stop2; }
        }
      }
      if (lane == 0) { s_wval = b; s_widx = (int)bi; s_wcta = bc; }
    }
    __syncthreads();
    if (!(s_wval > s_stop2)) break;       // same records everywhere: uniform exit
    const int widx = s_widx, wcta = s_wcta;
    // ---- D. the pivot column -> shared memory, its Householder reflector (zlarfg)
    {
      const cplx* src = A.cbuf + ((size_t)par * G + wcta) * p;
      double acc = 0.0;
      for (int i = j + tid; i < p; i += QT) {
        const cplx x = ldcg_c(src + i);
        vb[i] = x;
        if (i > j) { acc = fma(x.x, x.x, acc); acc = fma(x.y, x.y, acc); }
      }
      acc = warp_sum(acc);
      if (lane == 0) s_red[warp] = acc;
    }
    __syncthreads();
    double xnorm2 = 0.0;
#pragma unroll
    for (int w = 0; w < QW; ++w) xnorm2 += s_red[w];
    const cplx alpha = vb[j];
    double beta = alpha.x;
    cplx tau = make_double2(0.0, 0.0), scale = make_double2(0.0, 0.0);
    if (xnorm2 > 0.0 || alpha.y != 0.0) {
      const double an = sqrt(fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xnorm2)));
      beta = (alpha.x >= 0.0) ? -an : an;
      tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
      const double dx = alpha.x - beta, dy = alpha.y;
      const double dn = 1.0 / fma(dx, dx, dy * dy);
      scale = make_double2(dx * dn, -dy * dn);
    }
    __syncthreads();                      // everybody has read vb[j]
    for (int i = j + 1 + tid; i < p; i += QT) vb[i] = cmul(vb[i], scale);
    if (tid == 0) { vb[j] = make_double2(1.0, 0.0); done[widx] = 1; }
    __syncthreads();
    // ---- E. the owner files the pivot column: R[0..j, j], then the reflector
    if (me == wcta) {
      const int lcw = (widx - me) / G;
      cplx* gcol = A.a + (size_t)widx * p;
      if (lcw < A.nc_res) {
        const cplx* col = colptr(lcw);
        for (int i = tid; i < j; i += QT) gcol[i] = col[i];
      }
      for (int i = j + 1 + tid; i < p; i += QT) gcol[i] = vb[i];
      if (tid == 0) {
        gcol[j] = make_double2(beta, 0.0);
        A.perm[j] = widx;
        A.tauc[j] = tau;
      }
    }
    // ---- F. y <- (I - conj(tau) v v^H) y on my remaining columns; exact new norms
    const cplx ctau = cconj(tau);
    for (int base = 0; base < NC; base += slots) {
      const int lc = base + my_slot;
      const bool act = (lc < NC) && !done[me + lc * G];
      cplx* col = act ? colptr(lc) : nullptr;
      cplx w = make_double2(0.0, 0.0);
      if (act) {
        for (int i = j + my_part * 32 + lane; i < p; i += W * 32) w = cfma(cconj(vb[i]), col[i], w);
        w.x = warp_sum(w.x);
        w.y = warp_sum(w.y);
      }
      if (lane == 0) s_dot[warp] = w;
      __syncthreads();
      double nn = 0.0;
      if (act) {
        cplx tot = make_double2(0.0, 0.0);
        for (int u = 0; u < W; ++u) {           // fixed order
          const cplx d = s_dot[my_slot * W + u];
          tot.x += d.x; tot.y += d.y;
        }
        const cplx f = cmul(ctau, tot);
        for (int i = j + my_part * 32 + lane; i < p; i += W * 32) {
          const cplx v = vb[i];
          cplx x = col[i];
          x.x -= f.x * v.x - f.y * v.y;
          x.y -= f.x * v.y + f.y * v.x;
          col[i] = x;
          if (i > j) { nn = fma(x.x, x.x, nn); nn = fma(x.y, x.y, nn); }
        }
        nn = warp_sum(nn);
      }
      if (lane == 0) s_nrm[warp] = nn;
      __syncthreads();
      if (act && my_part == 0 && lane == 0) {
        double tot = 0.0;
        for (int u = 0; u < W; ++u) tot += s_nrm[my_slot * W + u];
        vn2[lc] = tot;
      }
    }
    __syncthreads();
  }
  const int k = j;

  // ---- discarded block: its Frobenius mass (the norms are exact as of the last pivot)
  if (tid == 0) {
    double t2 = 0.0;
    for (int lc = 0; lc < NC; ++lc)
      if (!done[me + lc * G]) t2 += vn2[lc];
    A.tail_part[me] = t2;
  }
  // ---- R12: rows 0..k-1 of my columns that never became a pivot (resident ones)
  for (int lc = 0; lc < NC && lc < A.nc_res; ++lc) {
    const int c = me + lc * G;
    if (done[c]) continue;
    const cplx* col = colptr(lc);
    cplx* gcol = A.a + (size_t)c * p;
    for (int i = tid; i < k; i += QT) gcol[i] = col[i];
  }
  // ---- positions k..q-1: the remaining physical columns in index order (CTA 0)
  if (me == 0) {
    __shared__ int s_cnt[QW];
    const int chunk = (q + QT - 1) / QT;
    const int c0 = tid * chunk, c1 = min(q, c0 + chunk);
    int cnt = 0;
    for (int c = c0; c < c1; ++c) cnt += done[c] ? 0 : 1;
    int incl = cnt;                       // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_cnt[warp] = incl;
    __syncthreads();
    int off = incl - cnt;
    for (int w = 0; w < warp; ++w) off += s_cnt[w];
    int pos = k + off;
    for (int c = c0; c < c1; ++c)
      if (!done[c]) A.perm[pos++] = c;
    if (tid == 0) {
      A.hdr->k = k;
      A.info_host[5] = G;
      __threadfence_system();             // info_host[4] is the word the host polls
      A.info_host[4] = k;
    }
  }
}

// ------------------------------------------------------------------ back-transformation
// Z = Q[:, :k] * J[:, sel]  for `keep` selected columns of the accumulated rotation J (k x k,
// held in the Jacobi kernel's block layout: rows rowW0.. of yjac, 16-column blocks of Tj rows).
// A CTA takes C columns, its threads own rows i = tid + t * blockDim (RPT rows each, in
// registers) and apply the reflectors k-1 ... 0:  z <- z - tau_r v_r (v_r^H z).  v_r is read
// from L2 once per CTA and reflector (prefetched one reflector ahead).
//   out_mode 0: left factor  U[i, jc]  at (i / u_na) u_so + (i % u_na) u_sa + jc u_sj
//   out_mode 1: right factor rows: out[jc * ld + i] = conj(z_i) * (scale ? sigma_jc : 1)
struct ApplyQArgs {
  const cplx* a;
  const int* qperm;
  const cplx* tauc;
  int p, k, keep;
  const cplx* yjac;
  long long Tj;
  int rowW0;
  const int* permJ;
  const double* sval;
  cplx* out;
  int out_mode, u_na, scale_sigma;
  long long u_so, u_sa, u_sj, ld;
};

template <int RPT, int C>
__global__ void __launch_bounds__(512) apply_q_kernel(const ApplyQArgs A) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nw = nt >> 5;
  const int j0 = blockIdx.x * C;
  __shared__ cplx s_part[2][32][C];
  cplx z[C][RPT];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int jc = j0 + c;
    const bool valid = jc < A.keep;
    const int col = valid ? A.permJ[jc] : 0;
    const size_t blk = col >> 4, c16 = col & 15;
#pragma unroll
    for (int t = 0; t < RPT; ++t) {
      const int i = tid + t * nt;
      z[c][t] = (valid && i < A.k) ? A.yjac[(blk * A.Tj + A.rowW0 + i) * 16 + c16]
                                   : make_double2(0.0, 0.0);
    }
  }
  cplx vnext[RPT];
  auto load_v = [&](int r, cplx* v) {
    const cplx* col = A.a + (size_t)A.qperm[r] * A.p;
#pragma unroll
    for (int t = 0; t < RPT; ++t) {
      const int i = tid + t * nt;
      v[t] = (i > r && i < A.p) ? ldcg_c(col + i)
                                : make_double2(i == r ? 1.0 : 0.0, 0.0);
    }
  };
  if (A.k > 0) load_v(A.k - 1, vnext);
  int buf = 0;
  for (int r = A.k - 1; r >= 0; --r) {
    cplx v[RPT];
#pragma unroll
    for (int t = 0; t < RPT; ++t) v[t] = vnext[t];
    if (r > 0) load_v(r - 1, vnext);
    const cplx tau = A.tauc[r];
    cplx w[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      w[c] = make_double2(0.0, 0.0);
#pragma unroll
      for (int t = 0; t < RPT; ++t) w[c] = cfma(cconj(v[t]), z[c][t], w[c]);
      w[c].x = warp_sum(w[c].x);
      w[c].y = warp_sum(w[c].y);
    }
    if (nw > 1) {
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) s_part[buf][warp][c] = w[c];
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < C; ++c) {
        cplx x = (lane < nw) ? s_part[buf][lane][c] : make_double2(0.0, 0.0);
        w[c].x = warp_sum(x.x);
        w[c].y = warp_sum(x.y);
      }
      buf ^= 1;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const cplx f = cmul(tau, w[c]);
#pragma unroll
      for (int t = 0; t < RPT; ++t) {
        z[c][t].x -= f.x * v[t].x - f.y * v[t].y;
        z[c][t].y -= f.x * v[t].y + f.y * v[t].x;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int jc = j0 + c;
    if (jc >= A.keep) continue;
    const double s = (A.out_mode == 1 && A.scale_sigma) ? A.sval[jc] : 1.0;
#pragma unroll
    for (int t = 0; t < RPT; ++t) {
      const int i = tid + t * nt;
      if (i >= A.p) continue;
      if (A.out_mode == 0) {
        A.out[(long long)(i / A.u_na) * A.u_so + (long long)(i % A.u_na) * A.u_sa +
              (long long)jc * A.u_sj] = z[c][t];
      } else {
        A.out[(long long)jc * A.ld + i] = make_double2(z[c][t].x * s, -z[c][t].y * s);
      }
    }
  }
}

// The other factor comes straight out of the iterated matrix Y = L J (q rows of the Jacobi
// block layout), scattered through the pivot permutation:
//   out_mode 1 (X = theta):    S Vh[jc, perm[i]] = conj(Y[i, c])  (/ sigma when unscaled)
//   out_mode 0 (X = theta^H):  U[perm[i], jc]    = Y[i, c] / sigma_jc
__global__ void emit_l_kernel(const cplx* __restrict__ yjac, long long Tj,
                              const int* __restrict__ permJ, const double* __restrict__ sval,
                              const int* __restrict__ qperm, int q, int keep,
                              cplx* __restrict__ out, int out_mode, int unscaled, int u_na,
                              long long u_so, long long u_sa, long long u_sj, long long ld,
                              cplx* __restrict__ lam, cplx* __restrict__ inv_lam) {
  if (blockIdx.x == 0) {
    for (int j = threadIdx.x; j < keep; j += blockDim.x) {
      const double s = sval[j];
      if (lam) lam[j] = make_double2(s, 0.0);
      if (inv_lam) inv_lam[j] = make_double2(1.0 / s, 0.0);
    }
  }
  if (!out) return;
  const long long total = (long long)q * keep;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int i, jc;
    if (out_mode == 1) { i = (int)(e % q); jc = (int)(e / q); }
    else { jc = (int)(e % keep); i = (int)(e / keep); }
    const int col = permJ[jc];
    const size_t blk = col >> 4, c16 = col & 15;
    cplx v = yjac[(blk * Tj + i) * 16 + c16];
    const double s = sval[jc];
    const double inv = (s > 0.0) ? 1.0 / s : 0.0;
    const int pi = qperm[i];
    if (out_mode == 1) {
      const double f = unscaled ? inv : 1.0;
      out[(long long)jc * ld + pi] = make_double2(v.x * f, -v.y * f);
    } else {
      out[(long long)(pi / u_na) * u_so + (long long)(pi % u_na) * u_sa +
          (long long)jc * u_sj] = make_double2(v.x * inv, v.y * inv);
    }
  }
}

}  // namespace qr
}  // namespace b200
