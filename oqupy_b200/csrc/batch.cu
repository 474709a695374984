// Lock-step ensemble engine: E independent TEMPO runs (same d2, dkmax, epsrel; different
// influence matrices, propagators, initial states) advance one time step per kernel launch.
//
// BASELINE configs[4] / SURVEY 8e: thousands of small TEMPO runs (chi <= ~24, truncated SVDs
// <= 96 x 96).  Each run is a strictly sequential chain of ~40 small factorisations per step,
// far too small for the grid-wide kernels of svd.cu and, driven one run at a time, bound by
// the host.  Here ONE CTA owns ONE member for the whole step: its chain lives in global
// memory (capacity-padded slots, shapes on the device), every SVD operand is built directly
// in shared memory and factorised there, nothing returns to the host between the sites.  148+
// members are in flight per GPU; the host issues one launch per step for all of them.
//
// One member step = BaseTempoBackend.compute_system_step (oqupy/backends/tempo_backend.py:
// 439-575): first half propagator (:521-529), sum out the oldest leg (:531-537), zip-up with
// the implicit influence MPO (:539-547 -> node_array.py:482-552), svd_sweep right-to-left
// (:549-553), append the second half propagator (:555-558), read-out (:560-573).
//
// The in-CTA truncated SVD (replaces tn.split_node_full_svd, node_array.py:262,285,541) is
// the small-matrix form of the pipeline in svd.cu / qrcp.cuh: column-pivoted Householder QR
// stopped at the deflation level (columns physically swapped, reflectors parked in global
// memory), cyclic one-sided Jacobi on the ROWS of R with the rotation accumulated in J, the
// reference's tail-norm rule, U = Q J / S Vh = R' P^T written straight into the chain.
// X is stored column-major with an ODD leading dimension: both its columns (QR) and its rows
// (Jacobi) are then free of shared-memory bank conflicts for 16-byte elements.
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

constexpr int BT = 512;
constexpr int BW = BT / 32;
constexpr int BATCH_SMEM = 223 * 1024;   // + ~3.7 KB static = the 227 KB a CTA can have
constexpr int MAXD = 128;          // largest operand dimension (chi * d2); an operand must also
                                   // fit shared memory: (p | 1) * q <= 14272 complex numbers

struct BatchDev {
  int E, d2, n_mpo, ns_slots, cap_chi, slot_elems;
  double eps;
  cplx* slots;        // [E][ns_slots][slot_elems]
  int* dims;          // [E][ns_slots][3]  (chi_l, a, chi_r)
  int* hdr;           // [E][8]: 0 n_sites, 1 head slot, 2 status, 3 svds, 4 sweeps, 5 max chi
  const cplx* mid;    // [E][n_infl][d2*d2]   B[w,n,s,e] = d_we d_ns mat[s,e]
  const cplx* start;  // [E][n_infl][d2*d2]   first aligned site: mat[s,e] * sum_west[e]
  const cplx* dense0; // [E][d2*d2][d2*d2]    dk = 0 site incl. the unitary transform
  const cplx* dense0w;// [E][d2][d2*d2]       the same with the west leg summed
  const cplx* sn;     // [d2] sum_north
  const cplx* p1;     // [E][d2*d2] prop_1 (row-major)
  const cplx* p2t;    // [E][d2*d2] prop_2^T
  int n_infl;
  cplx* carry;        // [E][2][cap_chi * cap_chi * d2]
  cplx* vg;           // [E][MAXD*(MAXD+1)] reflectors, then their taus
  cplx* jg;           // [E][MAXD*MAXD] rotation accumulator when it does not fit shared memory
  cplx* tmp;          // [E][slot_elems]
  cplx* states;       // [E][d2]
  int first_member;
  const int* order;   // [E] i-th member drawn from the work queue (nullptr: i): longest first
  int* counter;       // work queue head (zeroed before every launch)
};

__device__ __forceinline__ double wsum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// tournament partner table (circle method) over n (even) players, round in [0, n-1)
__device__ __forceinline__ void rr_pair(int idx, int round, int n, int& a, int& b) {
  const int ka = idx, kb = n - 1 - idx;
  a = (ka == 0) ? 0 : 1 + ((ka - 1 + round) % (n - 1));
  b = 1 + ((kb - 1 + round) % (n - 1));
  if (a > b) { const int t = a; a = b; b = t; }
}

struct SvdShared {
  double vn2[MAXD];
  double sig2[MAXD];
  int perm[MAXD];
  int order[MAXD];
  double red[BW];
  cplx cred[BW];
  int pivot, k, keep, flag, sweeps;
  double pval, stop2, fro2, tail2, beta;
  cplx tau, scale;
};

// ---------------------------------------------------------------- the in-CTA truncated SVD
// X (p x q, p >= q, column-major, leading dimension ld odd) holds theta (m >= n) or theta^H.
// Outputs: keep (returned); U[i, j] at u[i * u_ri + j * u_cj] (u_cj < 0: u_ri = keep, u_cj = 1,
// i.e. row-major m x keep); S Vh row-major (keep x n) in svh.
__device__ int cta_svd(SvdShared& S, cplx* X, int ld, cplx* Jsm, int jsm_elems, cplx* jglob,
                       cplx* vg, int m, int n, double eps, cplx* u, long long u_ri,
                       long long u_cj, cplx* svh, int* hdr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tr = m < n;
  const int p = tr ? n : m, q = tr ? m : n;

  // ---- column norms, ||X||_F
  for (int c = warp; c < q; c += BW) {
    const cplx* x = X + (size_t)c * ld;
    double s = 0.0;
    for (int i = lane; i < p; i += 32) { const cplx v = x[i]; s = fma(v.x, v.x, s); s = fma(v.y, v.y, s); }
    s = wsum(s);
    if (lane == 0) { S.vn2[c] = s; S.perm[c] = c; }
  }
  __syncthreads();
  if (tid == 0) {
    double f = 0.0;
    for (int c = 0; c < q; ++c) f += S.vn2[c];
    S.fro2 = f;
    const double sr = (eps > 0.0) ? 1e-5 * eps : 0.0;
    S.stop2 = sr * sr * f;
  }
  __syncthreads();

  // ---- stopped column-pivoted Householder QR
  int k = 0;
  for (int j = 0; j < q; ++j) {
    if (warp == 0) {
      double best = -1.0;
      int bi = 0x7fffffff;
      for (int c = j + lane; c < q; c += 32) {
        const double v = S.vn2[c];
        if (v > best || (v == best && c < bi)) { best = v; bi = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { S.pivot = bi; S.pval = best; }
    }
    __syncthreads();
    if (!(S.pval > S.stop2)) break;
    const int pv = S.pivot;
    if (pv != j) {
      cplx* a = X + (size_t)j * ld;
      cplx* b = X + (size_t)pv * ld;
      for (int i = tid; i < p; i += BT) { const cplx t = a[i]; a[i] = b[i]; b[i] = t; }
      if (tid == 0) {
        const double t = S.vn2[j]; S.vn2[j] = S.vn2[pv]; S.vn2[pv] = t;
        const int ti = S.perm[j]; S.perm[j] = S.perm[pv]; S.perm[pv] = ti;
      }
    }
    __syncthreads();
    cplx* xj = X + (size_t)j * ld;
    if (warp == 0) {                         // zlarfg: H^H x = beta e_j, beta real
      double s = 0.0;
      for (int i = j + 1 + lane; i < p; i += 32) { const cplx v = xj[i]; s = fma(v.x, v.x, s); s = fma(v.y, v.y, s); }
      s = wsum(s);
      if (lane == 0) {
        const cplx alpha = xj[j];
        double beta = alpha.x;
        cplx tau = make_double2(0.0, 0.0), scale = make_double2(0.0, 0.0);
        if (s > 0.0 || alpha.y != 0.0) {
          const double an = sqrt(fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, s)));
          beta = (alpha.x >= 0.0) ? -an : an;
          const double ib = 1.0 / beta;
          tau = make_double2((beta - alpha.x) * ib, -alpha.y * ib);
          const double dx = alpha.x - beta, dy = alpha.y;
          const double dn = 1.0 / fma(dx, dx, dy * dy);
          scale = make_double2(dx * dn, -dy * dn);
        }
        S.beta = beta; S.tau = tau; S.scale = scale;
      }
    }
    __syncthreads();
    {   // v = [1 ; scale * x]: kept in the column for the update, parked in global memory
      const cplx sc = S.scale;
      cplx* vcol = vg + (size_t)j * p;
      for (int i = j + tid; i < p; i += BT) {
        cplx v = (i == j) ? make_double2(1.0, 0.0) : b200::cmul(xj[i], sc);
        xj[i] = v;
        vcol[i] = v;
      }
    }
    __syncthreads();
    const cplx ctau = b200::cconj(S.tau);
    for (int c = j + 1 + warp; c < q; c += BW) {     // y <- (I - conj(tau) v v^H) y
      cplx* y = X + (size_t)c * ld;
      cplx w = make_double2(0.0, 0.0);
      for (int i = j + lane; i < p; i += 32) w = b200::cfma(b200::cconj(xj[i]), y[i], w);
      w.x = wsum(w.x);
      w.y = wsum(w.y);
      const cplx f = b200::cmul(ctau, w);
      double nn = 0.0;
      for (int i = j + lane; i < p; i += 32) {
        const cplx v = xj[i];
        cplx t = y[i];
        t.x -= f.x * v.x - f.y * v.y;
        t.y -= f.x * v.y + f.y * v.x;
        y[i] = t;
        if (i > j) { nn = fma(t.x, t.x, nn); nn = fma(t.y, t.y, nn); }
      }
      nn = wsum(nn);
      if (lane == 0) S.vn2[c] = nn;
    }
    __syncthreads();
    // the pivot column now holds R[0..j-1, j] above, beta on, zeros below the diagonal
    for (int i = j + tid; i < p; i += BT) xj[i] = (i == j) ? make_double2(S.beta, 0.0) : make_double2(0.0, 0.0);
    if (tid == 0) { S.vn2[j] = -1.0; vg[(size_t)MAXD * MAXD + j] = S.tau; }   // taus behind the reflectors
    k = j + 1;
    __syncthreads();
  }
  if (tid == 0) {
    double t2 = 0.0;
    for (int c = k; c < q; ++c) t2 += fmax(S.vn2[c], 0.0);
    S.tail2 = t2;
    S.k = k;
  }
  __syncthreads();
  if (k == 0) {                     // a zero operand
    if (tid == 0) S.keep = 0;
    __syncthreads();
    return 0;
  }

  // ---- only the k rows of R are live from here on: re-pack X to the leading dimension
  //      k | 1 (batches of 16 columns through registers; a batch never lands on the source of
  //      a later one), which frees shared memory for the rotation accumulator J
  {
    const int ld2 = k | 1;
    if (ld2 < ld) {
      for (int c0 = 0; c0 < q; c0 += 16) {
        const int nc = min(16, q - c0), tot = nc * k;
        cplx v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int e = tid + t * BT;
          if (e < tot) { const int c = e / k, i = e - c * k; v[t] = X[i + (size_t)(c0 + c) * ld]; }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int e = tid + t * BT;
          if (e < tot) { const int c = e / k, i = e - c * k; X[i + (size_t)(c0 + c) * ld2] = v[t]; }
        }
        __syncthreads();
      }
      Jsm = X + (size_t)ld2 * q;
      jsm_elems += (ld - ld2) * q;
      ld = ld2;
    }
  }
  // ---- cyclic one-sided Jacobi on the rows 0..k-1 of R (k x q), rotations accumulated in J
  cplx* J = ((size_t)k * k <= (size_t)jsm_elems) ? Jsm : jglob;
  for (int e = tid; e < k * k; e += BT) J[e] = make_double2((e / k == e % k) ? 1.0 : 0.0, 0.0);
  const double fro = sqrt(S.fro2);
  double tol = 2.0 * sqrt((double)p) * 2.220446049250313e-16;
  if (tol < 1e-11) tol = 1e-11;
  const double tol2 = tol * tol;
  const double floor_ = 8.0 * 2.220446049250313e-16 * fro;
  const double floor2 = floor_ * floor_;
  const double negr = (eps > 0.0) ? 1e-2 * eps * fro : 0.0;
  const double neg2 = negr * negr;
  const int ke = (k + 1) & ~1;       // even number of players (a dummy row when k is odd)
  __syncthreads();
  int sweeps = 0;
  for (; sweeps < 60; ++sweeps) {
    if (tid == 0) S.flag = 0;
    __syncthreads();
    for (int round = 0; round < ke - 1; ++round) {
      for (int idx = warp; idx < ke / 2; idx += BW) {
        int a, b;
        rr_pair(idx, round, ke, a, b);
        if (b >= k) continue;
        cplx* ra = X + a;
        cplx* rb = X + b;
        double ga = 0.0, gb = 0.0, xr = 0.0, xi = 0.0;
        for (int c = lane; c < q; c += 32) {
          const cplx va = ra[(size_t)c * ld], vb = rb[(size_t)c * ld];
          ga = fma(va.x, va.x, ga); ga = fma(va.y, va.y, ga);
          gb = fma(vb.x, vb.x, gb); gb = fma(vb.y, vb.y, gb);
          // G_ab = l_a^H l_b with l = conj(row):  sum_c R[a,c] conj(R[b,c])
          xr = fma(va.x, vb.x, xr); xr = fma(va.y, vb.y, xr);
          xi = fma(va.y, vb.x, xi); xi = fma(-va.x, vb.y, xi);
        }
        ga = wsum(ga); gb = wsum(gb); xr = wsum(xr); xi = wsum(xi);
        const double g2 = fma(xr, xr, xi * xi);
        const double big = fmax(ga, gb), small = fmax(fmin(ga, gb), 0.0);
        bool conv;
        if (big <= 0.0) conv = true;
        else if (big < neg2) conv = g2 <= 1e-4 * big * small + big * floor2;
        else conv = g2 <= big * (tol2 * small + floor2);
        // rotate above a much lower threshold than the convergence test (as svd.cu)
        const double fl = 0.0625 * floor2;
        const double thr = (big < neg2) ? fma(1e-4 * big, small, big * fl)
                                        : big * fma(1e-28, small, fl);
        const bool act = (g2 > thr) && (big > 0.0);
        if (!conv && lane == 0) S.flag = 1;
        if (act) {
          const double h = 0.5 * (gb - ga);
          const double inv_r = rsqrt(fma(h, h, g2) + 1e-300);
          const double w = fma(0.5 * fabs(h), inv_r, 0.5);
          const double ic = rsqrt(w);
          const double kk = ((h >= 0.0) ? 0.5 : -0.5) * inv_r * ic;
          const double c = w * ic;
          const double sr = xr * kk, si = xi * kk;        // s e^{i phi}
          const double d = ((h >= 0.0) ? 0.5 : -0.5) * g2 * inv_r * ic * ic;
          const bool swap = (ga - d) < (gb + d);          // larger norm first (de Rijk)
          // l_a' = c l_a - conj(se) l_b,  l_b' = se l_a + c l_b   (l = conj(row))
          // rows: R_a' = c R_a - se R_b,  R_b' = conj(se) R_a + c R_b
          for (int cc = lane; cc < q; cc += 32) {
            const cplx va = ra[(size_t)cc * ld], vb = rb[(size_t)cc * ld];
            cplx na, nb;
            na.x = c * va.x - (sr * vb.x - si * vb.y);
            na.y = c * va.y - (sr * vb.y + si * vb.x);
            nb.x = c * vb.x + (sr * va.x + si * va.y);
            nb.y = c * vb.y + (sr * va.y - si * va.x);
            ra[(size_t)cc * ld] = swap ? nb : na;
            rb[(size_t)cc * ld] = swap ? na : nb;
          }
          cplx* ja = J + (size_t)a * k;
          cplx* jb = J + (size_t)b * k;
          for (int r = lane; r < k; r += 32) {            // J[:,a]' = c J_a - conj(se) J_b ...
            const cplx va = ja[r], vb = jb[r];
            cplx na, nb;
            na.x = c * va.x - (sr * vb.x + si * vb.y);
            na.y = c * va.y - (sr * vb.y - si * vb.x);
            nb.x = c * vb.x + (sr * va.x - si * va.y);
            nb.y = c * vb.y + (sr * va.y + si * va.x);
            ja[r] = swap ? nb : na;
            jb[r] = swap ? na : nb;
          }
        }
      }
      __syncthreads();
    }
    if (S.flag == 0) { ++sweeps; break; }
    __syncthreads();
  }

  // ---- singular values, order, the reference's tail-norm rule
  for (int a = warp; a < k; a += BW) {
    double s = 0.0;
    for (int c = lane; c < q; c += 32) { const cplx v = X[a + (size_t)c * ld]; s = fma(v.x, v.x, s); s = fma(v.y, v.y, s); }
    s = wsum(s);
    if (lane == 0) S.sig2[a] = s;
  }
  __syncthreads();
  if (tid < k) {
    const double mine = S.sig2[tid];
    int rank = 0;
    for (int i = 0; i < k; ++i) {
      const double o = S.sig2[i];
      rank += (o > mine || (o == mine && i < tid)) ? 1 : 0;
    }
    S.order[rank] = tid;
  }
  __syncthreads();
  if (tid == 0) {
    int keep = k;
    if (eps >= 0.0) {
      const double s0 = sqrt(fmax(S.sig2[S.order[0]], 0.0));
      const double thr = eps * s0;
      double tail = S.tail2;
      keep = 0;
      for (int j = k - 1; j >= 0; --j) {
        const double s = sqrt(fmax(S.sig2[S.order[j]], 0.0));
        tail += s * s;
        if (sqrt(tail) > thr) ++keep;
      }
    }
    S.keep = keep;
    S.sweeps = sweeps;
    hdr[3] += 1;
    hdr[4] += sweeps;
    if (sweeps >= 60 && S.flag) hdr[2] = 3;      // no convergence
  }
  __syncthreads();
  const int keep = S.keep;
  if (u_cj < 0) { u_ri = keep; u_cj = 1; }

  // ---- the R' side: S Vh (X = theta) or U (X = theta^H), scattered through the permutation
  for (int e = tid; e < keep * q; e += BT) {
    const int jj = e / q, c = e - jj * q;
    const int row = S.order[jj];
    const cplx v = X[row + (size_t)c * ld];
    const int pc = S.perm[c];
    if (!tr) {
      svh[(size_t)jj * n + pc] = v;
    } else {
      const double s = sqrt(fmax(S.sig2[row], 0.0));
      const double inv = (s > 0.0) ? 1.0 / s : 0.0;
      u[(long long)pc * u_ri + (long long)jj * u_cj] = make_double2(v.x * inv, -v.y * inv);
    }
  }
  // ---- the Q side: z = Q [J[:, row]; 0], one warp per kept triplet, z in registers
  constexpr int ZR = (MAXD + 31) / 32;
  for (int jj = warp; jj < keep; jj += BW) {
    const int row = S.order[jj];
    const cplx* jc = J + (size_t)row * k;
    cplx z[ZR];
#pragma unroll
    for (int t = 0; t < ZR; ++t) {
      const int i = lane + 32 * t;
      z[t] = (i < k) ? jc[i] : make_double2(0.0, 0.0);
    }
    for (int r = k - 1; r >= 0; --r) {
      const cplx* v = vg + (size_t)r * p;
      const cplx tau = vg[(size_t)MAXD * MAXD + r];
      cplx w = make_double2(0.0, 0.0);
      cplx vv[ZR];
#pragma unroll
      for (int t = 0; t < ZR; ++t) {
        const int i = lane + 32 * t;
        vv[t] = (i >= r && i < p) ? v[i] : make_double2(0.0, 0.0);
        w = b200::cfma(b200::cconj(vv[t]), z[t], w);
      }
      w.x = wsum(w.x);
      w.y = wsum(w.y);
      const cplx f = b200::cmul(tau, w);
#pragma unroll
      for (int t = 0; t < ZR; ++t) {
        z[t].x -= f.x * vv[t].x - f.y * vv[t].y;
        z[t].y -= f.x * vv[t].y + f.y * vv[t].x;
      }
    }
    const double s = sqrt(fmax(S.sig2[row], 0.0));
#pragma unroll
    for (int t = 0; t < ZR; ++t) {
      const int i = lane + 32 * t;
      if (i >= p) continue;
      if (!tr) u[(long long)i * u_ri + (long long)jj * u_cj] = z[t];
      else svh[(size_t)jj * n + i] = make_double2(z[t].x * s, -z[t].y * s);
    }
  }
  __syncthreads();
  return keep;
}

// ---------------------------------------------------------------- one member, one time step
__device__ void member_step(const BatchDev& P, const int e, SvdShared& S, unsigned char* bsm) {
  const int tid = threadIdx.x;
  if (e >= P.E) return;
  const int d2 = P.d2, NS = P.ns_slots;
  int* hdr = P.hdr + (size_t)e * 8;
  if (hdr[2] != 0) return;                     // a failed member stays where it failed
  int* dims = P.dims + (size_t)e * NS * 3;
  cplx* slots = P.slots + (size_t)e * NS * P.slot_elems;
  cplx* X = reinterpret_cast<cplx*>(bsm);
  const int smem_elems = BATCH_SMEM / (int)sizeof(cplx);
  cplx* carry0 = P.carry + (size_t)e * 2 * P.cap_chi * P.cap_chi * d2;
  cplx* carry1 = carry0 + (size_t)P.cap_chi * P.cap_chi * d2;
  cplx* vg = P.vg + (size_t)e * MAXD * (MAXD + 1);
  cplx* jg = P.jg + (size_t)e * MAXD * MAXD;
  cplx* tmp = P.tmp + (size_t)e * P.slot_elems;
  const cplx* mid = P.mid + (size_t)e * P.n_infl * d2 * d2;
  const cplx* start = P.start + (size_t)e * P.n_infl * d2 * d2;
  int n_sites = hdr[0], head = hdr[1];
  const int n_mpo = P.n_mpo;
  auto slot_of = [&](int i) { return (head + i) % NS; };
  auto fail = [&](int code) { if (tid == 0) hdr[2] = code; };

  // ---- first half propagator on the newest site: last[l, j] <- sum_i last[l, i] P1[j, i]
  {
    const int sl = slot_of(n_sites - 1);
    cplx* a = slots + (size_t)sl * P.slot_elems;
    const int nl = dims[sl * 3 + 0];
    const cplx* p1 = P.p1 + (size_t)e * d2 * d2;
    for (int idx = tid; idx < nl * d2; idx += BT) {
      const int l = idx / d2, j = idx - l * d2;
      cplx acc = make_double2(0.0, 0.0);
      for (int i = 0; i < d2; ++i) acc = b200::cfma(a[l * d2 + i], p1[j * d2 + i], acc);
      tmp[idx] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < nl * d2; idx += BT) a[idx] = tmp[idx];
    __syncthreads();
  }
  // ---- sum out the oldest leg beyond the memory cut-off
  if (n_sites == n_mpo + 1) {
    const int s0 = slot_of(0), s1 = slot_of(1);
    const cplx* f = slots + (size_t)s0 * P.slot_elems;
    cplx* g = slots + (size_t)s1 * P.slot_elems;
    const int fa = dims[s0 * 3 + 1], fr = dims[s0 * 3 + 2];
    const int ga = dims[s1 * 3 + 1], gr = dims[s1 * 3 + 2];
    cplx* vec = tmp;                            // [fr]
    for (int r = tid; r < fr; r += BT) {
      cplx acc = make_double2(0.0, 0.0);
      for (int a = 0; a < fa; ++a) acc = b200::cfma(P.sn[a], f[a * fr + r], acc);
      vec[r] = acc;
    }
    __syncthreads();
    cplx* merged = tmp + MAXD;
    for (int idx = tid; idx < ga * gr; idx += BT) {
      cplx acc = make_double2(0.0, 0.0);
      for (int r = 0; r < fr; ++r) acc = b200::cfma(vec[r], g[(size_t)r * ga * gr + idx], acc);
      merged[idx] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < ga * gr; idx += BT) g[idx] = merged[idx];
    if (tid == 0) { dims[s1 * 3 + 0] = 1; }
    head = (head + 1) % NS;
    n_sites -= 1;
    __syncthreads();
  } else if (n_sites != n_mpo) {
    fail(4);
    return;
  }

  // ---- zip-up, direction "right": Theta[(k,s),(r,e)] = M[s,e] sum_l C[k,l,e] A[l,s,r]
  cplx* carry_in = carry0;
  cplx* carry_out = carry1;
  int ck = 0, cl = 0;                           // carry (k, l, e)
  for (int ib = 0; ib < n_mpo; ++ib) {
    const int sl = slot_of(ib);
    cplx* a = slots + (size_t)sl * P.slot_elems;
    const int nl = dims[sl * 3 + 0], nn = dims[sl * 3 + 1], nr = dims[sl * 3 + 2];
    const int dk = n_mpo - 1 - ib;
    if (ib == n_mpo - 1) {                      // dense dk = 0 site, no SVD
      if (nr != 1 || nn != d2) { fail(5); return; }
      const int nw = (ib == 0) ? 1 : d2;
      const cplx* dmat = (ib == 0) ? P.dense0w + (size_t)e * d2 * d2 * d2
                                   : P.dense0 + (size_t)e * d2 * d2 * d2 * d2;
      const int nk = (ib == 0) ? 1 : ck;
      const int nse = d2 * d2;
      if (ib != 0 && cl != nl) { fail(6); return; }
      // T[k,w,n] = sum_l C[k,l,w] A[l,n]   (first site: T[0,0,n] = A[0,n])
      cplx* T = tmp;
      for (int idx = tid; idx < nk * nw * nn; idx += BT) {
        const int kk = idx / (nw * nn), rem = idx - kk * nw * nn;
        const int w = rem / nn, nq = rem - w * nn;
        cplx acc = make_double2(0.0, 0.0);
        if (ib == 0) acc = a[nq];
        else
          for (int l = 0; l < nl; ++l)
            acc = b200::cfma(carry_in[((size_t)kk * nl + l) * d2 + w], a[l * nn + nq], acc);
        T[idx] = acc;
      }
      __syncthreads();
      if (nk * nse > P.slot_elems) { fail(7); return; }
      for (int idx = tid; idx < nk * nse; idx += BT) {
        const int kk = idx / nse, c = idx - kk * nse;
        cplx acc = make_double2(0.0, 0.0);
        for (int t = 0; t < nw * nn; ++t) acc = b200::cfma(T[kk * nw * nn + t], dmat[(size_t)t * nse + c], acc);
        X[idx] = acc;                           // staged: T aliases nothing of the site
      }
      __syncthreads();
      for (int idx = tid; idx < nk * nse; idx += BT) a[idx] = X[idx];
      if (tid == 0) { dims[sl * 3 + 0] = nk; dims[sl * 3 + 1] = d2; dims[sl * 3 + 2] = d2; }
      __syncthreads();
      break;
    }
    const cplx* mat = (ib == 0) ? start + (size_t)dk * d2 * d2 : mid + (size_t)dk * d2 * d2;
    if (nn != d2) { fail(8); return; }
    const int nk = (ib == 0) ? 1 : ck;
    if (ib == 0 && nl != 1) { fail(9); return; }
    if (ib != 0 && cl != nl) { fail(6); return; }
    const int m = nk * d2, n = nr * d2;
    if (m > MAXD || n > MAXD) { fail(2); return; }
    const bool tr = m < n;
    const int p = tr ? n : m, q = tr ? m : n;
    const int ld = p | 1;
    const int jsm = smem_elems - ld * q;
    if (jsm < 0) { fail(2); return; }
    // build Theta straight into shared memory, in the orientation the SVD wants
    for (int idx = tid; idx < m * n; idx += BT) {
      const int i = idx / n, jx = idx - i * n;           // i = (k, s), jx = (r, e)
      const int kk = i / d2, s = i - kk * d2;
      const int r = jx / d2, ee = jx - r * d2;
      cplx acc = make_double2(0.0, 0.0);
      if (ib == 0) acc = a[s * nr + r];
      else
        for (int l = 0; l < nl; ++l)
          acc = b200::cfma(carry_in[((size_t)kk * nl + l) * d2 + ee], a[((size_t)l * d2 + s) * nr + r], acc);
      acc = b200::cmul(acc, mat[s * d2 + ee]);
      if (!tr) X[i + (size_t)jx * ld] = acc;
      else X[jx + (size_t)i * ld] = b200::cconj(acc);
    }
    __syncthreads();
    // new site = U as (k, s, j) row-major m x keep; carry' = S Vh as (j, r, e)
    const int keep = cta_svd(S, X, ld, X + (size_t)ld * q, jsm, jg, vg, m, n, P.eps, a, 0, -1,
                             carry_out, hdr);
    if (keep > P.cap_chi || keep < 1) { fail(keep < 1 ? 10 : 2); return; }
    if (tid == 0) {
      dims[sl * 3 + 0] = nk; dims[sl * 3 + 1] = d2; dims[sl * 3 + 2] = keep;
      if (keep > hdr[5]) hdr[5] = keep;
    }
    ck = keep; cl = nr;
    cplx* t = carry_in; carry_in = carry_out; carry_out = t;
    __syncthreads();
  }

  // ---- svd_sweep right -> left: site i as (a r) x l
  for (int i = n_sites - 1; i > 0; --i) {
    const int sl = slot_of(i), sb = slot_of(i - 1);
    cplx* a = slots + (size_t)sl * P.slot_elems;
    cplx* b = slots + (size_t)sb * P.slot_elems;
    const int nl = dims[sl * 3 + 0], na = dims[sl * 3 + 1], nr = dims[sl * 3 + 2];
    const int m = na * nr, n = nl;
    if (m > MAXD || n > MAXD) { fail(2); return; }
    const bool tr = m < n;
    const int p = tr ? n : m, q = tr ? m : n;
    const int ld = p | 1;
    const int jsm = smem_elems - ld * q;
    if (jsm < 0) { fail(2); return; }
    for (int idx = tid; idx < m * n; idx += BT) {
      const int l = idx / m, ii = idx - l * m;            // a[l][ii]: theta[ii][l]
      const cplx v = a[idx];
      if (!tr) X[ii + (size_t)l * ld] = v;
      else X[l + (size_t)ii * ld] = b200::cconj(v);
    }
    __syncthreads();
    cplx* svh = carry_out;                                // keep x nl
    const int keep = cta_svd(S, X, ld, X + (size_t)ld * q, jsm, jg, vg, m, n, P.eps, a, 1, m,
                             svh, hdr);
    if (keep > P.cap_chi || keep < 1) { fail(keep < 1 ? 10 : 2); return; }
    // neighbour: nb[(bl,ba), j] = sum_l b[(bl,ba), l] svh[j, l]   (staged: in-place reshape)
    const int bl = dims[sb * 3 + 0], ba = dims[sb * 3 + 1];
    const int rows = bl * ba;
    for (int idx = tid; idx < rows * keep; idx += BT) {
      const int rw = idx / keep, jj = idx - rw * keep;
      cplx acc = make_double2(0.0, 0.0);
      for (int l = 0; l < nl; ++l) acc = b200::cfma(b[(size_t)rw * nl + l], svh[(size_t)jj * nl + l], acc);
      tmp[idx] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < rows * keep; idx += BT) b[idx] = tmp[idx];
    if (tid == 0) {
      dims[sl * 3 + 0] = keep;
      dims[sb * 3 + 2] = keep;
      if (keep > hdr[5]) hdr[5] = keep;
    }
    __syncthreads();
  }

  // ---- append the second half propagator as a site (d2, d2, 1) = prop_2^T
  {
    if (n_sites + 1 > NS) { fail(11); return; }
    const int sl = slot_of(n_sites);
    cplx* a = slots + (size_t)sl * P.slot_elems;
    const cplx* p2t = P.p2t + (size_t)e * d2 * d2;
    for (int idx = tid; idx < d2 * d2; idx += BT) a[idx] = p2t[idx];
    if (tid == 0) { dims[sl * 3 + 0] = d2; dims[sl * 3 + 1] = d2; dims[sl * 3 + 2] = 1; }
    n_sites += 1;
    __syncthreads();
  }
  // ---- read-out: vec <- sum_a sn[a] (vec . A[:, a, :]) over all sites but the last
  {
    cplx* vec = tmp;                  // [<= cap]
    cplx* nxt = tmp + MAXD;
    if (tid == 0) vec[0] = make_double2(1.0, 0.0);
    __syncthreads();
    for (int i = 0; i < n_sites - 1; ++i) {
      const int sl = slot_of(i);
      const cplx* a = slots + (size_t)sl * P.slot_elems;
      const int nl = dims[sl * 3 + 0], na = dims[sl * 3 + 1], nr = dims[sl * 3 + 2];
      for (int r = tid; r < nr; r += BT) {
        cplx acc = make_double2(0.0, 0.0);
        for (int l = 0; l < nl; ++l) {
          const cplx vl = vec[l];
          for (int aa = 0; aa < na; ++aa)
            acc = b200::cfma(b200::cmul(vl, P.sn[aa]), a[((size_t)l * na + aa) * nr + r], acc);
        }
        nxt[r] = acc;
      }
      __syncthreads();
      for (int r = tid; r < nr; r += BT) vec[r] = nxt[r];
      __syncthreads();
    }
    const int sl = slot_of(n_sites - 1);
    const cplx* a = slots + (size_t)sl * P.slot_elems;     // (d2, d2, 1)
    if (tid < d2) {
      cplx acc = make_double2(0.0, 0.0);
      for (int l = 0; l < d2; ++l) acc = b200::cfma(vec[l], a[l * d2 + tid], acc);
      P.states[(size_t)e * d2 + tid] = acc;
    }
  }
  if (tid == 0) { hdr[0] = n_sites; hdr[1] = head; }
}

// PERSISTENT launch: gridDim.x CTAs (at most one per SM, fewer when SMs are kept free for
// members that are re-run on the general path next to the batch) draw members from a work
// queue in the order of P.order (longest-running first).  All exits of member_step are
// CTA-uniform (they depend on the member's shapes only).
__global__ void __launch_bounds__(BT, 1) tempo_batch_step_kernel(const BatchDev P) {
  extern __shared__ __align__(16) unsigned char bsm[];
  __shared__ SvdShared S;
  __shared__ int s_next;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_next = atomicAdd(P.counter, 1);
    __syncthreads();
    const int i = s_next;
    if (i >= P.E) break;
    member_step(P, P.order ? P.order[i] : P.first_member + i, S, bsm);
  }
}

struct Batch {
  BatchDev dev;
  cudaStream_t stream;
  int E, d2, dkmax, cap_chi;
  int step;
  size_t bytes;
  void* arena;
  int* order_buf;
  int reserve_sms;    // SMs the persistent launch leaves to other streams
  int sms;
};

}  // namespace

// ============================================================================ C-ABI
extern "C" {

/* Lock-step TEMPO ensemble (oqupy/backends/tempo_backend.py:439-575 for E members at once).
 * All members share d2, dkmax (>= 1, finite) and epsrel; chi_cap bounds the bond dimension
 * (chi_cap * d2 <= 128; every truncated SVD operand must fit shared memory). */
void* b200_tempo_batch_create(void* stream, int n_members, int d2, int dkmax, int chi_cap,
                              double epsrel) {
  if (n_members < 1 || d2 < 1 || dkmax < 1 || chi_cap < 1 || chi_cap * d2 > MAXD ||
      chi_cap < d2) {
    b200::set_error("b200_tempo_batch_create: invalid argument (chi_cap*d2 must be <= %d)", MAXD);
    return nullptr;
  }
  Batch* b = new Batch();
  b->stream = (cudaStream_t)stream;
  b->E = n_members; b->d2 = d2; b->dkmax = dkmax; b->cap_chi = chi_cap; b->step = 0;
  BatchDev& D = b->dev;
  D.E = n_members; D.d2 = d2; D.cap_chi = chi_cap; D.eps = epsrel;
  D.ns_slots = dkmax + 3;
  D.slot_elems = chi_cap * d2 * chi_cap;
  if (D.slot_elems < d2 * d2 * d2) D.slot_elems = d2 * d2 * d2;
  if (D.slot_elems < 2 * MAXD + chi_cap * d2 * d2) D.slot_elems = 2 * MAXD + chi_cap * d2 * d2;
  D.n_infl = dkmax + 1;
  D.first_member = 0;
  D.order = nullptr;
  const size_t E = (size_t)n_members, d4 = (size_t)d2 * d2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_slots = take(E * D.ns_slots * D.slot_elems * sizeof(cplx));
  const size_t o_dims = take(E * D.ns_slots * 3 * sizeof(int));
  const size_t o_hdr = take(E * 8 * sizeof(int));
  const size_t o_mid = take(E * D.n_infl * d4 * sizeof(cplx));
  const size_t o_start = take(E * D.n_infl * d4 * sizeof(cplx));
  const size_t o_d0 = take(E * d4 * d4 * sizeof(cplx));
  const size_t o_d0w = take(E * d2 * d4 * sizeof(cplx));
  const size_t o_sn = take(d2 * sizeof(cplx));
  const size_t o_p1 = take(E * d4 * sizeof(cplx));
  const size_t o_p2 = take(E * d4 * sizeof(cplx));
  const size_t o_carry = take(E * 2 * (size_t)chi_cap * chi_cap * d2 * sizeof(cplx));
  const size_t o_vg = take(E * (size_t)MAXD * (MAXD + 1) * sizeof(cplx));
  const size_t o_jg = take(E * (size_t)MAXD * MAXD * sizeof(cplx));
  const size_t o_tmp = take(E * D.slot_elems * sizeof(cplx));
  const size_t o_states = take(E * d2 * sizeof(cplx));
  const size_t o_order = take(E * sizeof(int));
  const size_t o_counter = take(sizeof(int));
  b->bytes = off;
  if (cudaMalloc(&b->arena, off) != cudaSuccess) {
    b200::set_error("b200_tempo_batch_create: cudaMalloc(%zu) failed", off);
    (void)cudaGetLastError();
    delete b;
    return nullptr;
  }
  cudaMemsetAsync(b->arena, 0, off, b->stream);
  unsigned char* base = (unsigned char*)b->arena;
  D.slots = (cplx*)(base + o_slots); D.dims = (int*)(base + o_dims); D.hdr = (int*)(base + o_hdr);
  D.mid = (cplx*)(base + o_mid); D.start = (cplx*)(base + o_start);
  D.dense0 = (cplx*)(base + o_d0); D.dense0w = (cplx*)(base + o_d0w); D.sn = (cplx*)(base + o_sn);
  D.p1 = (cplx*)(base + o_p1); D.p2t = (cplx*)(base + o_p2); D.carry = (cplx*)(base + o_carry);
  D.vg = (cplx*)(base + o_vg); D.jg = (cplx*)(base + o_jg); D.tmp = (cplx*)(base + o_tmp);
  D.states = (cplx*)(base + o_states);
  b->order_buf = (int*)(base + o_order);
  D.counter = (int*)(base + o_counter);
  b->reserve_sms = 0;
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
      (void)cudaGetLastError();
      sms = 148;
    }
    b->sms = sms;
  }
  static std::once_flag once;       // ensembles drive the library from several host threads
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(tempo_batch_step_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, BATCH_SMEM);
  });
  if (attr_rc != cudaSuccess) {
    b200::set_error("b200_tempo_batch_create: shared memory attribute failed");
    (void)cudaGetLastError();
    cudaFree(b->arena);
    delete b;
    return nullptr;
  }
  return b;
}

int b200_tempo_batch_destroy(void* h) {
  Batch* b = (Batch*)h;
  if (!b) return B200_OK;
  cudaStreamSynchronize(b->stream);
  cudaFree(b->arena);
  delete b;
  return B200_OK;
}

/* Device-to-device upload of the per-member tables (all complex128, member-major):
 *   mid, start   (E, dkmax+1, d2, d2)  influence matrices by dk: infl[dk][s,e] and
 *                infl[dk][s,e] * sum_west[e] (tempo_backend.py:419,426; :519)
 *   dense0       (E, d2*d2, d2*d2)     the dk = 0 site incl. the unitary transform,
 *                [(w,n),(s,e)] (:419-424);   dense0w (E, d2, d2*d2) with the west leg summed
 *   sum_north    (d2)
 *   state0       (E, d2)               initial states: site (1, d2, 1) */
int b200_tempo_batch_set(void* h, const void* mid, const void* start, const void* dense0,
                         const void* dense0w, const void* sum_north, const void* state0) {
  Batch* b = (Batch*)h;
  if (!b || !mid || !start || !dense0 || !dense0w || !sum_north || !state0) {
    b200::set_error("b200_tempo_batch_set: invalid argument");
    return B200_EINVAL;
  }
  BatchDev& D = b->dev;
  const size_t E = (size_t)b->E, d4 = (size_t)b->d2 * b->d2;
  cudaStream_t s = b->stream;
  B200_CUDA_CHECK(cudaMemcpyAsync((void*)D.mid, mid, E * D.n_infl * d4 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  B200_CUDA_CHECK(cudaMemcpyAsync((void*)D.start, start, E * D.n_infl * d4 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  B200_CUDA_CHECK(cudaMemcpyAsync((void*)D.dense0, dense0, E * d4 * d4 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  B200_CUDA_CHECK(cudaMemcpyAsync((void*)D.dense0w, dense0w, E * b->d2 * d4 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  B200_CUDA_CHECK(cudaMemcpyAsync((void*)D.sn, sum_north, b->d2 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  // chains: one site (1, d2, 1) per member
  B200_CUDA_CHECK(cudaMemsetAsync(D.hdr, 0, E * 8 * sizeof(int), s));
  B200_CUDA_CHECK(cudaMemcpy2DAsync(D.slots, (size_t)D.ns_slots * D.slot_elems * sizeof(cplx), state0,
                                    b->d2 * sizeof(cplx), b->d2 * sizeof(cplx), E,
                                    cudaMemcpyDeviceToDevice, s));
  std::vector<int> hdims(E * D.ns_slots * 3, 0), hhdr(E * 8, 0);
  for (size_t e = 0; e < E; ++e) {
    hdims[e * D.ns_slots * 3 + 0] = 1; hdims[e * D.ns_slots * 3 + 1] = b->d2; hdims[e * D.ns_slots * 3 + 2] = 1;
    hhdr[e * 8 + 0] = 1;
  }
  B200_CUDA_CHECK(cudaMemcpyAsync(D.dims, hdims.data(), hdims.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  B200_CUDA_CHECK(cudaMemcpyAsync(D.hdr, hhdr.data(), hhdr.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  B200_CUDA_CHECK(cudaStreamSynchronize(s));
  b->step = 0;
  return B200_OK;
}

/* One time step of every member: p1 (E, d2, d2) = prop_1 and p2t (E, d2, d2) = prop_2^T of
 * this step (device); states_out (E, d2) device, may be NULL (kept internally).  ONE kernel
 * launch; nothing is read back. */
int b200_tempo_batch_step(void* h, const void* p1, const void* p2t, void* states_out) {
  Batch* b = (Batch*)h;
  if (!b || !p1 || !p2t) { b200::set_error("b200_tempo_batch_step: invalid argument"); return B200_EINVAL; }
  BatchDev& D = b->dev;
  const size_t E = (size_t)b->E, d4 = (size_t)b->d2 * b->d2;
  cudaStream_t s = b->stream;
  B200_CUDA_CHECK(cudaMemcpyAsync((void*)D.p1, p1, E * d4 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  B200_CUDA_CHECK(cudaMemcpyAsync((void*)D.p2t, p2t, E * d4 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  b->step += 1;
  D.n_mpo = (b->step <= b->dkmax) ? b->step : b->dkmax + 1;
  b200::profile_begin(s, 3);
  B200_CUDA_CHECK(cudaMemsetAsync(D.counter, 0, sizeof(int), s));
  int grid = b->sms - b->reserve_sms;
  if (grid < 1) grid = 1;
  if (grid > b->E) grid = b->E;
  tempo_batch_step_kernel<<<grid, BT, BATCH_SMEM, s>>>(D);
  B200_LAUNCH_CHECK();
  b200::profile_end(s, 3, 0.0, nullptr);
  if (states_out)
    B200_CUDA_CHECK(cudaMemcpyAsync(states_out, D.states, E * b->d2 * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  return B200_OK;
}

/* Per member (host arrays, synchronises the stream): status (0 ok, 2 bond dimension or
 * operand exceeds the capacity, 3 Jacobi did not converge, >= 4 internal), truncated SVDs and
 * Jacobi sweeps so far, largest bond dimension; bond dimensions of the current chain into
 * bonds[e * (dkmax+2) + i] (-1 padded). */
int b200_tempo_batch_info(void* h, int32_t* status, int32_t* svds, int32_t* sweeps,
                          int32_t* max_chi, int32_t* bonds) {
  Batch* b = (Batch*)h;
  if (!b) { b200::set_error("b200_tempo_batch_info: invalid argument"); return B200_EINVAL; }
  BatchDev& D = b->dev;
  const size_t E = (size_t)b->E;
  std::vector<int> hhdr(E * 8), hdims(E * D.ns_slots * 3);
  B200_CUDA_CHECK(cudaMemcpyAsync(hhdr.data(), D.hdr, hhdr.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  B200_CUDA_CHECK(cudaMemcpyAsync(hdims.data(), D.dims, hdims.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  B200_CUDA_CHECK(cudaStreamSynchronize(b->stream));
  const int nb = b->dkmax + 2;
  for (size_t e = 0; e < E; ++e) {
    if (status) status[e] = hhdr[e * 8 + 2];
    if (svds) svds[e] = hhdr[e * 8 + 3];
    if (sweeps) sweeps[e] = hhdr[e * 8 + 4];
    if (max_chi) max_chi[e] = hhdr[e * 8 + 5];
    if (bonds) {
      const int n_sites = hhdr[e * 8 + 0], head = hhdr[e * 8 + 1];
      for (int i = 0; i < nb; ++i) {
        int v = -1;
        if (i < n_sites - 1) v = hdims[(e * D.ns_slots + (head + i) % D.ns_slots) * 3 + 2];
        bonds[e * nb + i] = v;
      }
    }
  }
  return B200_OK;
}

/* CTA i of the following steps takes member order[i] (a permutation of 0..E-1, host array;
 * NULL restores the identity).  One launch advances E members that differ widely in cost
 * (bond dimension^3): handing the longest-running members out FIRST keeps the last wave of
 * CTAs short (longest-processing-time-first; measured on the 4096-member grid of BASELINE
 * configs[4] at 512 members per GPU). */
int b200_tempo_batch_set_order(void* h, const int32_t* order) {
  Batch* b = (Batch*)h;
  if (!b) { b200::set_error("b200_tempo_batch_set_order: invalid argument"); return B200_EINVAL; }
  if (!order) { b->dev.order = nullptr; return B200_OK; }
  std::vector<char> seen((size_t)b->E, 0);
  for (int i = 0; i < b->E; ++i) {
    if (order[i] < 0 || order[i] >= b->E || seen[(size_t)order[i]]) {
      b200::set_error("b200_tempo_batch_set_order: not a permutation of 0..%d", b->E - 1);
      return B200_EINVAL;
    }
    seen[(size_t)order[i]] = 1;
  }
  // (pageable source: the copy is staged before the call returns)
  B200_CUDA_CHECK(cudaMemcpyAsync(b->order_buf, order, (size_t)b->E * sizeof(int),
                                  cudaMemcpyHostToDevice, b->stream));
  b->dev.order = b->order_buf;
  return B200_OK;
}

/* The following steps leave `n` SMs to other streams (0: none): work that runs NEXT TO the
 * batch -- members re-run on the general backend -- needs free SMs for its cooperative
 * launches, which a launch that fills the GPU never yields. */
int b200_tempo_batch_reserve_sms(void* h, int n) {
  Batch* b = (Batch*)h;
  if (!b || n < 0) { b200::set_error("b200_tempo_batch_reserve_sms: invalid argument"); return B200_EINVAL; }
  b->reserve_sms = (n >= b->sms) ? b->sms - 1 : n;
  return B200_OK;
}

size_t b200_tempo_batch_bytes(void* h) { return h ? ((Batch*)h)->bytes : 0; }

}  // extern "C"
