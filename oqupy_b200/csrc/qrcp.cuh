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
constexpr int QR_SMEM_BYTES = 218 * 1024;  // dynamic shared memory requested for the kernel
                                           // (+ ~8 KB static = the 227 KB a CTA can have)
constexpr int QR_MAX_P = 8192;        // rows (apply_q keeps p / threads <= 8 rows per thread)

struct QrHeader {
  int k, status;
  double fro2, stop2;
  long long phase_cycles[8];  // CTA 0 (PROF): 0 pick own best, 1 poll wait, 2 select, 3 fetch,
                              // 4 panel, 5 file pivots, 6 apply, 7 publish the offer
};

struct QrLayout {
  size_t header, cand_tag, fro_part, tail_part, ctrl_bytes;   // zeroed region first
  size_t perm, tau, cbuf, a, jac, total;
  int p, q, G, NCmax, nc_res, transposed, panel;
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
  unsigned long long* cand_tag;  // [2][G] x TAG_STRIDE words (one 128-byte line per offer: G
                                 // CTAs poll G words each, on one line that is an L2 hot spot):
                                 // norm^2 (top 35 bits) | hand-shake (15) | column (14)
  double* fro_part;           // [G]
  double* tail_part;          // [G]
  cplx* cbuf;                 // [2][G][p] speculatively published candidate columns
  int* perm;                  // [q] position -> physical column
  cplx* tauc;                 // [q]
  QrHeader* hdr;
  int32_t* info_host;         // pinned; info_host[4] <- k
  double stop_rel;            // stop = stop_rel * ||X||_F
  double theta2;              // panel acceptance: remaining norm^2 > theta2 * best offer outside
  int nc_res;                 // local columns resident in shared memory
  int ncmax;
  int panel;                  // pivots taken per hand-shake, at most (<= PBMAX)
  int fast_panel;             // 1: one warp per panel column, one CTA barrier per pivot (short
                              // columns); 0: two warps per column
};

constexpr int PBMAX = 8;
constexpr int TAG_STRIDE = 16;   // 64-bit words between two offers

// theta element (i, j) of the ORIGINAL m x n operand
__device__ __forceinline__ cplx theta_at(const QrArgs& A, int i, int j) {
  return A.theta[(long long)(i / A.rin) * A.rs + (long long)(i % A.rin) * A.rsi +
                 (long long)(j / A.cin) * A.cs + (long long)(j % A.cin) * A.csi];
}

// One grid-wide hand-shake selects a PANEL of up to `panel` pivots: every CTA offers its best
// remaining column; the largest offers (different CTAs) are fetched by everybody and
// factorised redundantly -- greedy pivoting inside the panel on the exact remaining norms, a
// panel column is only taken while its remaining norm^2 stays above theta2 x the best offer
// left outside the panel (and above the stop level); the others simply stay ordinary columns
// of their CTAs.  (CPU study tools/study_panel.py: theta = 0.5 keeps the sweep count of the
// Jacobi stage at that of strict column pivoting with 3-4 pivots per hand-shake; theta = 0
// doubles it.)  Then every CTA applies the taken reflectors to its own columns, one warp per
// column, no CTA barrier inside, and accumulates their exact remaining norms.
template <bool PROF>
__global__ void __launch_bounds__(QT, 1) qrcp_kernel(const QrArgs A) {
  extern __shared__ __align__(16) unsigned char qsm[];
  long long pc[PROF ? 8 : 1];
  long long tq = 0;
  if constexpr (PROF) {
#pragma unroll
    for (int k_ = 0; k_ < 8; ++k_) pc[k_] = 0;
    tq = clock64();
  }
#define QPHASE(k) if constexpr (PROF) { const long long tn_ = clock64(); pc[k] += tn_ - tq; tq = tn_; }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, me = blockIdx.x;
  const int p = A.p, q = A.q;
  const int NC = (q - me + G - 1) / G;          // my local columns: c = me + lc * G
  const int PB = A.panel;
  cplx* pan = reinterpret_cast<cplx*>(qsm);                      // [PB][p] panel / reflectors
  cplx* scols = pan + (size_t)PB * p;                            // [nc_res][p]
  double* vn2 = reinterpret_cast<double*>(scols + (size_t)A.nc_res * p);   // [ncmax]
  unsigned char* done = reinterpret_cast<unsigned char*>(vn2 + A.ncmax);   // [q]
  __shared__ cplx s_dot[QW];
  __shared__ double s_pn[QW];
  __shared__ int s_bl;
  __shared__ double s_best, s_stop2, s_outside;
  __shared__ int s_np, s_pidx[PBMAX], s_pcta[PBMAX];
  __shared__ int s_tslot[PBMAX];          // panel slot taken as pivot t
  __shared__ cplx s_tau[PBMAX], s_scale[PBMAX];
  __shared__ double s_beta[PBMAX];
  __shared__ cplx s_vv[PBMAX][PBMAX];     // v_t^H v_s (s < t) of the reflectors of a panel
  __shared__ cplx s_fd[QW][PBMAX];        // partial v_t^H y per warp (columns split over warps)
  __shared__ cplx s_ff[QW][PBMAX];        // the coefficients f_t of a column (per column slot)
  __shared__ cplx s_fs[QW][PBMAX];        // f_t * scale_t
  __shared__ double s_fn[QW];
  __shared__ double s_full[PBMAX];        // remaining norm^2 of a panel column incl. its diagonal
  __shared__ int s_eready;                // reflectors whose scalars are published (fast panel)

  auto colptr = [&](int lc) -> cplx* {
    return (lc < A.nc_res) ? scols + (size_t)lc * p : A.a + (size_t)(me + lc * G) * p;
  };

  for (int c = tid; c < q; c += QT) done[c] = 0;
  // ---- load my columns (X = theta or theta^H), exact norms
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
    double my_fro = 0.0;
    for (int lc = 0; lc < NC; ++lc) my_fro += vn2[lc];     // fixed order
    A.fro_part[me] = my_fro;
    s_stop2 = 0.0;
  }

  int j = 0;
  for (unsigned shake = 1; j < q; ++shake) {
    const int par = shake & 1;
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
      if (lane == 0) { s_bl = (best >= 0.0) ? bl : -1; s_best = (best >= 0.0) ? best : 0.0; }
    }
    __syncthreads();
    QPHASE(0)
    // ---- B. offer it (rows j..p-1); the release-store of the tag is this CTA's arrival
    const int bl = s_bl;
    if (bl >= 0) {
      const cplx* col = colptr(bl);
      cplx* dst = A.cbuf + ((size_t)par * G + me) * p;
      for (int i = j + tid; i < p; i += QT) dst[i] = col[i];
    }
    __syncthreads();
    if (tid == 0) {
      const unsigned long long vb35 =
          ((unsigned long long)__double_as_longlong(s_best) >> 29) << 29;
      const unsigned long long idx = (bl >= 0) ? (unsigned long long)(me + bl * G) : 0x3fffull;
      st_release_u64(A.cand_tag + (size_t)(par * G + me) * TAG_STRIDE,
                     vb35 | ((unsigned long long)(shake & 0x7fffu) << 14) | idx);
    }
    QPHASE(7)
    // ---- C. poll the G offers; the PB largest above the stop level form the panel
    if (warp == 0) {
      unsigned long long key[5];       // norm^2 (35 bits) | 0x3fff - column (14) | CTA (8)
      unsigned pending = 0u;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        key[u] = 0ull;
        if (u * 32 + lane < G) pending |= 1u << u;
      }
      while (pending) {                // re-read only the offers that have not arrived yet
        unsigned long long tag[5];
#pragma unroll
        for (int u = 0; u < 5; ++u)
          if (pending & (1u << u))
            tag[u] = ld_acquire_u64(A.cand_tag + (size_t)(par * G + u * 32 + lane) * TAG_STRIDE);
#pragma unroll
        for (int u = 0; u < 5; ++u)
          if ((pending & (1u << u)) &&
              (((unsigned)(tag[u] >> 14) & 0x7fffu) == (shake & 0x7fffu))) {
            const unsigned long long idx = tag[u] & 0x3fffull;
            key[u] = ((tag[u] >> 29) << 22) | ((0x3fffull - idx) << 8) |
                     (unsigned long long)(u * 32 + lane);
            pending &= ~(1u << u);
          }
      }
      QPHASE(1)
      if (shake == 1) {         // ||X||_F^2: every CTA sums the G partials in the same order
        double f = 0.0;
        for (int g = lane; g < G; g += 32) f += __ldcg(A.fro_part + g);
        f = warp_sum(f);
        if (lane == 0) {
          s_stop2 = (A.stop_rel * A.stop_rel) * f;
          if (me == 0) { A.hdr->fro2 = f; A.hdr->stop2 = s_stop2; }
        }
      }
      __syncwarp();
      const double stop2 = s_stop2;
      int np = 0;
      double outside = 0.0;
      for (int t = 0; t <= PB; ++t) {
        unsigned long long m = 0ull;
#pragma unroll
        for (int u = 0; u < 5; ++u) m = (key[u] > m) ? key[u] : m;
        {   // warp maximum of a 57-bit key: two 32-bit hardware reductions
          const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(m >> 32));
          const unsigned lo = __reduce_max_sync(
              0xffffffffu, ((unsigned)(m >> 32) == hi) ? (unsigned)(m & 0xffffffffull) : 0u);
          m = ((unsigned long long)hi << 32) | lo;
        }
        const double val = __longlong_as_double((long long)((m >> 22) << 29));
        if (!(val > stop2)) break;
        if (t == PB) { outside = val; break; }
#pragma unroll
        for (int u = 0; u < 5; ++u)
          if (key[u] == m) key[u] = 0ull;          // keys are unique (CTA field)
        if (lane == 0) {
          s_pidx[t] = (int)(0x3fffull - ((m >> 8) & 0x3fffull));
          s_pcta[t] = (int)(m & 0xffull);
        }
        np = t + 1;
      }
      if (lane == 0) { s_np = np; s_outside = outside; }
    }
    __syncthreads();
    const int np = s_np;
    QPHASE(2)
    if (np == 0) break;                   // same offers everywhere: uniform exit
    // ---- D. fetch the panel columns (rows j..p-1): all loads of a batch in flight at once
    {
      const int len = p - j, total = np * len;
      for (int e0 = tid; e0 < total; e0 += 4 * QT) {
        cplx v[4];
        int dsto[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * QT;
          dsto[u] = -1;
          if (e < total) {
            const int t = e / len, i = j + (e - t * len);
            v[u] = ldcg_c(A.cbuf + ((size_t)par * G + s_pcta[t]) * p + i);
            dsto[u] = t * p + i;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (dsto[u] >= 0) pan[dsto[u]] = v[u];
      }
    }
    __syncthreads();
    QPHASE(3)
    // ---- E. factorise the panel (every CTA, bit-identical): greedy pivoting inside it
    const double thr = fmax(s_stop2, A.theta2 * s_outside);
    unsigned rem = (1u << np) - 1u;       // panel slots not yet taken
    int bt = 0;
    if (A.fast_panel) {
      // SHORT columns.  The fp64 pipe issues one warp instruction per 4 cycles and SM
      // sub-partition whatever the number of active lanes, so work repeated in every WARP is
      // what costs here:
      //  * warp c < np owns panel column c: dot product and update of a step without any
      //    cross-warp reduction, ONE CTA barrier per pivot;
      //  * the reflector's scalars (sqrt, two divisions) are formed by the pivot's own warp
      //    only -- it has nothing else to do -- while the others already run their dot
      //    products (which do not need them) and pick the scalars up through a flag;
      //  * the pivot choice is an integer arg-max over the bit patterns of the remaining
      //    norms (no fp64 instruction);
      //  * the spare warps form the inner products x_t^H x_s of the reflectors (phase F).
      // The new diagonal entry y[row] is written one step late: row `row` of every column is
      // still being read for the pivot choice.
      const int c = warp;
      const bool owner = c < np;
      cplx* y = pan + (size_t)(owner ? c : 0) * p;
      if (owner) {
        double nn = 0.0, al2 = 0.0;
        for (int i = j + lane; i < p; i += 32) {
          const cplx v = y[i];
          const double a2 = fma(v.x, v.x, v.y * v.y);
          if (i == j) al2 = a2; else nn += a2;
        }
        nn = warp_sum(nn);
        if (lane == 0) { s_pn[c] = nn; s_full[c] = nn + al2; }
      }
      if (tid == 0) s_eready = 0;
      __syncthreads();
      cplx pend = make_double2(0.0, 0.0);
      int pend_row = -1;
      for (int t = 0; t < np; ++t) {
        const int row = j + t;
        if (pend_row >= 0 && lane == 0) y[pend_row] = pend;
        pend_row = -1;
        // arg-max of the remaining full norms (non-negative doubles order like integers);
        // ties go to the lowest slot
        int bc;
        double bfull;
        {
          const bool cand = (lane < np) && ((rem >> lane) & 1u);
          const unsigned long long key =
              cand ? (unsigned long long)__double_as_longlong(s_full[lane]) + 1ull : 0ull;
          const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(key >> 32));
          const unsigned lo = __reduce_max_sync(
              0xffffffffu, ((unsigned)(key >> 32) == hi) ? (unsigned)(key & 0xffffffffull) : 0u);
          const unsigned long long m = ((unsigned long long)hi << 32) | lo;
          const unsigned who = __ballot_sync(0xffffffffu, cand && key == m);
          bc = __ffs(who) - 1;
          bfull = __longlong_as_double((long long)(m - 1ull));
        }
        if (t > 0 && !(bfull > thr)) break;
        rem &= ~(1u << bc);
        const cplx* x0 = pan + (size_t)bc * p;
        const bool more = !(rem == 0u || t + 1 >= np);
        const bool act = more && owner && ((rem >> c) & 1u);
        if (c == bc) {
          // the pivot's warp: scalars of the reflector (unscaled: v = [1 ; scale * x])
          const cplx alpha = x0[row];
          const double xnorm2 = s_pn[bc];
          double beta = alpha.x;
          cplx tau = make_double2(0.0, 0.0), scale = make_double2(0.0, 0.0);
          if (xnorm2 > 0.0 || alpha.y != 0.0) {
            const double an = sqrt(fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xnorm2)));
            beta = (alpha.x >= 0.0) ? -an : an;
            const double ib = 1.0 / beta;
            tau = make_double2((beta - alpha.x) * ib, -alpha.y * ib);
            const double dx = alpha.x - beta, dy = alpha.y;
            const double dn = 1.0 / fma(dx, dx, dy * dy);
            scale = make_double2(dx * dn, -dy * dn);
          }
          if (lane == 0) {
            s_tslot[t] = bc; s_tau[t] = tau; s_beta[t] = beta; s_scale[t] = scale;
            __threadfence_block();
            *(volatile int*)&s_eready = t + 1;
          }
        } else if (!owner) {
          // spare warps: x_t^H x_s over rows > row for the earlier reflectors s < t
          for (int s_ = c - np; s_ < t; s_ += QW - np) {
            const cplx* xs = pan + (size_t)s_tslot[s_] * p;
            cplx acc = make_double2(0.0, 0.0);
            for (int i = row + 1 + lane; i < p; i += 32) acc = cfma(cconj(x0[i]), xs[i], acc);
            acc.x = warp_sum(acc.x);
            acc.y = warp_sum(acc.y);
            if (lane == 0) s_vv[t][s_] = acc;
          }
        }
        bt = t + 1;
        if (!more) break;
        if (act) {
          const cplx yr = y[row];
          cplx w = make_double2(0.0, 0.0);
          for (int i = row + 1 + lane; i < p; i += 32) w = cfma(cconj(x0[i]), y[i], w);
          w.x = warp_sum(w.x);
          w.y = warp_sum(w.y);
          while (*(volatile int*)&s_eready < t + 1) {}
          __threadfence_block();
          const volatile double* vt_ = reinterpret_cast<const volatile double*>(&s_tau[t]);
          const volatile double* vs_ = reinterpret_cast<const volatile double*>(&s_scale[t]);
          const cplx tau = make_double2(vt_[0], vt_[1]);
          const cplx scale = make_double2(vs_[0], vs_[1]);
          cplx tot = cmul(cconj(scale), w);
          tot.x += yr.x; tot.y += yr.y;
          const cplx f = cmul(cconj(tau), tot);          // y -= f v
          const cplx fs = cmul(f, scale);
          double nn = 0.0, al2 = 0.0;
          for (int i = row + 1 + lane; i < p; i += 32) {
            const cplx xi = x0[i];
            cplx yy = y[i];
            yy.x -= fs.x * xi.x - fs.y * xi.y;
            yy.y -= fs.x * xi.y + fs.y * xi.x;
            y[i] = yy;
            const double a2 = fma(yy.x, yy.x, yy.y * yy.y);
            if (i == row + 1) al2 = a2; else nn += a2;
          }
          nn = warp_sum(nn);
          if (lane == 0) { s_pn[c] = nn; s_full[c] = nn + al2; }
          pend = make_double2(yr.x - f.x, yr.y - f.y);
          pend_row = row;
        }
        __syncthreads();
      }
      if (pend_row >= 0 && lane == 0) y[pend_row] = pend;
    } else {
    {
      // remaining norms^2 over rows > j of every panel column: two warps per column
      const int c = warp >> 1, half = warp & 1;
      double nn = 0.0;
      if (c < np) {
        const cplx* x = pan + (size_t)c * p;
        for (int i = j + 1 + half * 32 + lane; i < p; i += 64) {
          const cplx v = x[i];
          nn = fma(v.x, v.x, nn); nn = fma(v.y, v.y, nn);
        }
        nn = warp_sum(nn);
      }
      if (lane == 0) s_pn[warp] = nn;
    }
    __syncthreads();
    for (int t = 0; t < np; ++t) {
      const int row = j + t;
      // best remaining panel column by its full remaining norm (rows >= row)
      int bc = -1;
      double bfull = -1.0;
      for (int c = 0; c < np; ++c)
        if (rem & (1u << c)) {
          const cplx al = pan[(size_t)c * p + row];
          const double full = s_pn[2 * c] + s_pn[2 * c + 1] + fma(al.x, al.x, al.y * al.y);
          if (full > bfull) { bfull = full; bc = c; }
        }
      if (t > 0 && !(bfull > thr)) break;
      rem &= ~(1u << bc);
      // the reflector stays UNSCALED in shared memory: v = [1 ; scale * x], H = I - tau v v^H
      const cplx* x0 = pan + (size_t)bc * p;
      const cplx alpha = x0[row];
      const double xnorm2 = s_pn[2 * bc] + s_pn[2 * bc + 1];
      double beta = alpha.x;
      cplx tau = make_double2(0.0, 0.0), scale = make_double2(0.0, 0.0);
      if (xnorm2 > 0.0 || alpha.y != 0.0) {
        const double an = sqrt(fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xnorm2)));
        beta = (alpha.x >= 0.0) ? -an : an;
        const double ib = 1.0 / beta;
        tau = make_double2((beta - alpha.x) * ib, -alpha.y * ib);
        const double dx = alpha.x - beta, dy = alpha.y;
        const double dn = 1.0 / fma(dx, dx, dy * dy);
        scale = make_double2(dx * dn, -dy * dn);
      }
      if (tid == 0) {
        s_tslot[t] = bc; s_tau[t] = tau; s_beta[t] = beta; s_scale[t] = scale;
      }
      bt = t + 1;
      if (rem == 0u || t + 1 >= np) break;
      // apply H^H to the remaining panel columns (two warps per column), new norms over
      // rows > row + 1:  w = v^H y = y_row + conj(scale) sum_{i>row} conj(x_i) y_i
      const int c = warp >> 1, half = warp & 1;
      const bool act = (c < np) && (rem & (1u << c));
      cplx* y = pan + (size_t)c * p;
      cplx w = make_double2(0.0, 0.0);
      cplx yr = make_double2(0.0, 0.0);
      if (act) {
        yr = y[row];                    // read by both warps BEFORE the barrier, rewritten after
        for (int i = row + 1 + half * 32 + lane; i < p; i += 64) w = cfma(cconj(x0[i]), y[i], w);
        w.x = warp_sum(w.x);
        w.y = warp_sum(w.y);
      }
      if (lane == 0) s_dot[warp] = w;
      __syncthreads();
      double nn = 0.0;
      if (act) {
        const cplx d0 = s_dot[2 * c], d1 = s_dot[2 * c + 1];
        cplx tot = cmul(cconj(scale), make_double2(d0.x + d1.x, d0.y + d1.y));
        tot.x += yr.x; tot.y += yr.y;
        const cplx f = cmul(cconj(tau), tot);          // y -= f v
        const cplx fs = cmul(f, scale);
        for (int i = row + 1 + half * 32 + lane; i < p; i += 64) {
          const cplx xi = x0[i];
          cplx yy = y[i];
          yy.x -= fs.x * xi.x - fs.y * xi.y;
          yy.y -= fs.x * xi.y + fs.y * xi.x;
          y[i] = yy;
          if (i > row + 1) { nn = fma(yy.x, yy.x, nn); nn = fma(yy.y, yy.y, nn); }
        }
        if (half == 0 && lane == 0) y[row] = make_double2(yr.x - f.x, yr.y - f.y);
        nn = warp_sum(nn);
      }
      if (lane == 0) s_pn[warp] = nn;
      __syncthreads();
    }
    }
    // (the break conditions above are thread-uniform: they only read shared memory)
    __syncthreads();
    if (tid < bt) done[s_pidx[s_tslot[tid]]] = 1;
    __syncthreads();
    QPHASE(4)
    // ---- G. the owners file the pivot columns: R above the diagonal, then the reflector
    for (int t = 0; t < bt; ++t) {
      const int slot = s_tslot[t];
      if (s_pcta[slot] != me) continue;
      const int widx = s_pidx[slot];
      const int lcw = (widx - me) / G;
      const int row = j + t;
      cplx* gcol = A.a + (size_t)widx * p;
      const cplx* x = pan + (size_t)slot * p;
      const cplx sc = s_scale[t];
      if (lcw < A.nc_res) {
        const cplx* col = colptr(lcw);
        for (int i = tid; i < j; i += QT) gcol[i] = col[i];
      }
      for (int i = j + tid; i < p; i += QT) {
        if (i < row) gcol[i] = x[i];                       // R entries of this panel's rows
        else if (i > row) gcol[i] = cmul(x[i], sc);        // the reflector
      }
      if (tid == 0) gcol[row] = make_double2(s_beta[t], 0.0);
    }
    if (me == 0 && tid < bt) {
      A.perm[j + tid] = s_pidx[s_tslot[tid]];
      A.tauc[j + tid] = s_tau[tid];
    }
    QPHASE(5)
    if (!A.fast_panel) {
    // ---- F. y <- H_bt^H ... H_1^H y on my remaining columns, one warp per column; exact
    //         remaining norms (rows >= j + bt) in the last pass
    for (int lc = warp; lc < NC; lc += QW) {
      if (done[me + lc * G]) continue;
      cplx* col = colptr(lc);
      double nn = 0.0;
      for (int t = 0; t < bt; ++t) {
        const int row = j + t;
        const cplx* x = pan + (size_t)s_tslot[t] * p;
        const cplx sc = s_scale[t];
        cplx w = make_double2(0.0, 0.0);
        const cplx yr = col[row];
        for (int i = row + 1 + lane; i < p; i += 32) w = cfma(cconj(x[i]), col[i], w);
        w.x = warp_sum(w.x);
        w.y = warp_sum(w.y);
        cplx tot = cmul(cconj(sc), w);
        tot.x += yr.x; tot.y += yr.y;
        const cplx f = cmul(cconj(s_tau[t]), tot);
        const cplx fs = cmul(f, sc);
        const bool last = (t == bt - 1);
        for (int i = row + 1 + lane; i < p; i += 32) {
          const cplx xi = x[i];
          cplx yy = col[i];
          yy.x -= fs.x * xi.x - fs.y * xi.y;
          yy.y -= fs.x * xi.y + fs.y * xi.x;
          col[i] = yy;
          if (last) { nn = fma(yy.x, yy.x, nn); nn = fma(yy.y, yy.y, nn); }
        }
        __syncwarp();
        if (lane == 0) col[row] = make_double2(yr.x - f.x, yr.y - f.y);
        __syncwarp();
      }
      nn = warp_sum(nn);
      if (lane == 0) vn2[lc] = nn;
    }
    } else {
    // ---- F (short columns). y <- H_bt^H ... H_1^H y on my remaining columns in TWO passes
    //         over the rows instead of bt dependent dot/update pairs:  with d_t = v_t^H y on
    //         the column as it stands and C[t][s] = v_t^H v_s, the coefficients follow from
    //         f_t = conj(tau_t) (d_t - sum_{s<t} f_s C[t][s]), then y -= sum_t f_t v_t.
    //         Rows of a column are split over `wpc` warps; the bt partial dot products of a warp
    //         are reduced in one transposing butterfly (16 adds instead of 80); ONE warp per
    //         column solves the recurrence, one reflector per lane.  Exact remaining norms
    //         (rows >= j + bt) come out of the second pass.
    {
      const cplx* xp[PBMAX];
#pragma unroll
      for (int t = 0; t < PBMAX; ++t) xp[t] = pan + (size_t)s_tslot[(t < bt) ? t : 0] * p;
      if (tid < bt * (bt - 1) / 2) {      // C[t][s] from the raw inner products x_t^H x_s
        int t = 1, s_ = tid;
        while (s_ >= t) { s_ -= t; ++t; }
        const cplx ss = s_scale[s_];
        cplx v = cmul(cconj(s_scale[t]), s_vv[t][s_]);
        const cplx h = pan[(size_t)s_tslot[s_] * p + j + t];
        v.x += h.x; v.y += h.y;
        s_vv[t][s_] = cmul(v, ss);
      }
      const int wpc_ = (NC <= 4) ? 4 : ((NC <= 8) ? 2 : 1);
      if (wpc_ == 1) __syncthreads();     // (otherwise the barrier after the first pass does it)
      const int wpc = (NC <= 4) ? 4 : ((NC <= 8) ? 2 : 1);     // warps per column
      const int cpb = QW / wpc;                                 // columns per batch
      const int part = warp % wpc, cs = warp / wpc;
      for (int lc0 = 0; lc0 < NC; lc0 += cpb) {
        const int lc = lc0 + cs;
        const bool live = (lc < NC) && !done[me + lc * G];
        cplx* col = live ? colptr(lc) : nullptr;
        if (live) {
          double v[2 * PBMAX];
#pragma unroll
          for (int t = 0; t < 2 * PBMAX; ++t) v[t] = 0.0;
          for (int i = j + part * 32 + lane; i < p; i += 32 * wpc) {
            const cplx yi = col[i];
#pragma unroll
            for (int t = 0; t < PBMAX; ++t)
              if (t < bt && i > j + t) {
                const cplx xi = xp[t][i];
                v[2 * t] = fma(xi.x, yi.x, v[2 * t]); v[2 * t] = fma(xi.y, yi.y, v[2 * t]);
                v[2 * t + 1] = fma(xi.x, yi.y, v[2 * t + 1]);
                v[2 * t + 1] = fma(-xi.y, yi.x, v[2 * t + 1]);
              }
          }
          // transposing butterfly: lane l ends up with entry (l >> 1) & 15 summed over the warp
#pragma unroll
          for (int h = PBMAX, off = 16; h >= 1; h >>= 1, off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int k_ = 0; k_ < h; ++k_) {
              const double keep = upper ? v[k_ + h] : v[k_];
              const double send = upper ? v[k_] : v[k_ + h];
              v[k_] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
          if ((lane & 1) == 0) reinterpret_cast<double*>(&s_fd[warp][0])[lane >> 1] = v[0];
        }
        if (wpc > 1) __syncthreads(); else __syncwarp();
        // the column's solver warp (spread over the SM sub-partitions): lane t owns reflector t
        if (live && part == cs % wpc) {
          cplx tot = make_double2(0.0, 0.0), sct = make_double2(0.0, 0.0);
          cplx tau = make_double2(0.0, 0.0);
          if (lane < bt) {
            cplx d = s_fd[cs * wpc][lane];
            for (int u = 1; u < wpc; ++u) { const cplx o = s_fd[cs * wpc + u][lane]; d.x += o.x; d.y += o.y; }
            sct = s_scale[lane];
            tau = s_tau[lane];
            const cplx yr = col[j + lane];
            tot = cmul(cconj(sct), d);
            tot.x += yr.x; tot.y += yr.y;
          }
          cplx f = make_double2(0.0, 0.0);
          for (int s_ = 0; s_ < bt; ++s_) {
            const cplx fc = cmul(cconj(tau), tot);
            const double fx = __shfl_sync(0xffffffffu, fc.x, s_);
            const double fy = __shfl_sync(0xffffffffu, fc.y, s_);
            if (lane == s_) f = make_double2(fx, fy);
            if (lane > s_ && lane < bt) {
              const cplx cv = s_vv[lane][s_];
              tot.x -= fx * cv.x - fy * cv.y;
              tot.y -= fx * cv.y + fy * cv.x;
            }
          }
          if (lane < PBMAX) {
            s_ff[cs][lane] = f;
            s_fs[cs][lane] = cmul(f, sct);
          }
        }
        if (wpc > 1) __syncthreads(); else __syncwarp();
        double nn = 0.0;
        if (live) {
          cplx fs[PBMAX];
#pragma unroll
          for (int t = 0; t < PBMAX; ++t) fs[t] = s_fs[cs][t];
          for (int i = j + part * 32 + lane; i < p; i += 32 * wpc) {
            cplx yy = col[i];
            if (i >= j + bt) {
#pragma unroll
              for (int t = 0; t < PBMAX; ++t)
                if (t < bt) {
                  const cplx xi = xp[t][i];
                  yy.x -= fs[t].x * xi.x - fs[t].y * xi.y;
                  yy.y -= fs[t].x * xi.y + fs[t].y * xi.x;
                }
              nn = fma(yy.x, yy.x, nn); nn = fma(yy.y, yy.y, nn);
            } else {
#pragma unroll
              for (int t = 0; t < PBMAX; ++t)
                if (t < bt) {
                  if (i > j + t) {
                    const cplx xi = xp[t][i];
                    yy.x -= fs[t].x * xi.x - fs[t].y * xi.y;
                    yy.y -= fs[t].x * xi.y + fs[t].y * xi.x;
                  } else if (i == j + t) {
                    const cplx ft = s_ff[cs][t];
                    yy.x -= ft.x; yy.y -= ft.y;
                  }
                }
            }
            col[i] = yy;
          }
          nn = warp_sum(nn);
        }
        if (wpc > 1) { if (lane == 0) s_fn[warp] = nn; }
        else if (live && lane == 0) vn2[lc] = nn;
        if (wpc == 1) __syncwarp();
      }
      if (wpc > 1) {      // (a single batch) the partial norms meet after the CTA barrier
        __syncthreads();
        if (tid < NC && !done[me + tid * G]) {
          double tot = 0.0;
          for (int u = 0; u < wpc; ++u) tot += s_fn[tid * wpc + u];
          vn2[tid] = tot;
        }
      }
    }
    }
    j += bt;
    __syncthreads();
    QPHASE(6)
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
      if constexpr (PROF)
        for (int k_ = 0; k_ < 8; ++k_) A.hdr->phase_cycles[k_] = pc[k_];
      A.info_host[5] = G;
      __threadfence_system();             // info_host[4] is the word the host polls
      A.info_host[4] = k;
    }
  }
#undef QPHASE
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
