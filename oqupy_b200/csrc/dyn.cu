// compute_dynamics hot loop on the device (C-ABI: b200_dyn_step, b200_caps_step).
//
// Replaces _apply_system_superoperator / _apply_pt_mpos / _apply_caps
// (oqupy/system_dynamics.py:631-700) and SimpleProcessTensor.compute_caps
// (oqupy/process_tensor.py:380-406).  The rank-3 PT-MPO site T[l, r, x] is used as
// stored: the reference's dense (l, r, x, x') delta expansion
// (oqupy/process_tensor.py:346-347) is never materialised.
//
// HBM-bound: one step streams T (chi_l*chi_r*d2*16 B) exactly once; `nvec`
// ensemble members sharing the process tensor reuse each T element from registers.
#include <mutex>

#include "common.cuh"

namespace {

constexpr int DT = 128;   // threads per CTA of the streaming kernel
constexpr int EV = 4;     // ensemble members per thread

// u[e,l,x] = sum_i P1[e][x,i] v[e,l,i];  optional read-out rho[e,i] = sum_l cap[l] v[e,l,i]
__global__ void dyn_pre_kernel(int nvec, int chi_l, int d2, const cplx* __restrict__ p1,
                               const cplx* __restrict__ v, cplx* __restrict__ u,
                               const cplx* __restrict__ cap, cplx* __restrict__ rho) {
  const int e = blockIdx.y;
  const cplx* P = p1 + (size_t)e * d2 * d2;
  const cplx* V = v + (size_t)e * chi_l * d2;
  cplx* U = u + (size_t)e * chi_l * d2;
  const int total = chi_l * d2;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < total;
       c += gridDim.x * blockDim.x) {
    const int l = c / d2, x = c % d2;
    cplx acc = make_double2(0.0, 0.0);
    for (int i = 0; i < d2; ++i) acc = b200::cfma(P[x * d2 + i], V[l * d2 + i], acc);
    U[c] = acc;
  }
  if (cap && rho && blockIdx.x == 0) {
    // one warp per system index i (d2 <= 32 warps worth handled in a loop)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    for (int i = warp; i < d2; i += nwarp) {
      cplx acc = make_double2(0.0, 0.0);
      for (int l = lane; l < chi_l; l += 32) acc = b200::cfma(cap[l], V[l * d2 + i], acc);
      for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      }
      if (lane == 0) rho[(size_t)e * d2 + i] = acc;
    }
  }
}

// ---- TMA (bulk async copy) helpers: the rows of T stream into shared memory through a
// 4-stage mbarrier pipeline, one cp.async.bulk per row chunk.  (Plain LDG tops out near
// 2 TB/s here: the per-SM miss queues bound the bytes in flight; a bulk copy is ONE
// request for kilobytes.)
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes,
                                          unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

constexpr int DW = 256;    // columns of T per CTA (4 KB per row chunk)
constexpr int DST = 4;     // pipeline stages
constexpr int RPS = 4;     // rows of T per stage (4 bulk copies on one mbarrier): 64 KB in flight
constexpr int DCPT = DW / DT;   // columns per thread

// One launch per step:  u = P1 v for the CTA's own l-range (shared memory) -> stream T
// (TMA pipeline) -> partial sums -> the LAST CTA of a column block (atomic ticket) reduces
// the partials and applies P2.  CTA (0, 0, z) also does the read-out
// rho = sum_l cap[l] v[l, :].
template <int EVT>
__global__ void __launch_bounds__(DT)
dyn_fused_kernel(int nvec, int chi_l, int chi_r, int d2, int lchunk, int cols_per_cta,
                 const cplx* __restrict__ t, const cplx* __restrict__ p1,
                 const cplx* __restrict__ p2, const cplx* __restrict__ v,
                 cplx* __restrict__ v_out, const cplx* __restrict__ cap,
                 cplx* __restrict__ rho, cplx* __restrict__ partial,
                 unsigned int* __restrict__ tickets) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  cplx* stage = reinterpret_cast<cplx*>(dyn_smem);                 // [DST][RPS][DW]
  cplx* us = stage + DST * RPS * DW;                               // [EVT][lchunk][d2]
  __shared__ __align__(8) unsigned long long full[DST];
  __shared__ int s_last;
  const int ncol = chi_r * d2;
  const int s = blockIdx.y, ns = gridDim.y - 1;
  const int e0 = blockIdx.z * EVT;
  if (s == ns) {
    // ---- the extra CTA row: read-out of the INPUT state, rho[e, i] = sum_l cap[l] v[e, l, i]
    // (system_dynamics.py:143-147); it never delays a streaming CTA
    if (!cap || !rho || blockIdx.x != 0) return;
    cplx* red = reinterpret_cast<cplx*>(dyn_smem);                 // [DT]
    for (int k = 0; k < EVT; ++k) {
      if (e0 + k >= nvec) break;
      const cplx* V = v + (size_t)(e0 + k) * chi_l * d2;
      for (int i = 0; i < d2; ++i) {
        cplx a = make_double2(0.0, 0.0);
#pragma unroll 4
        for (int l = threadIdx.x; l < chi_l; l += DT) a = b200::cfma(cap[l], V[l * d2 + i], a);
        red[threadIdx.x] = a;
        __syncthreads();
        for (int o = DT / 2; o > 0; o >>= 1) {
          if (threadIdx.x < o) {
            red[threadIdx.x].x += red[threadIdx.x + o].x;
            red[threadIdx.x].y += red[threadIdx.x + o].y;
          }
          __syncthreads();
        }
        if (threadIdx.x == 0) rho[(size_t)(e0 + k) * d2 + i] = red[0];
        __syncthreads();
      }
    }
    return;
  }
  const int c0 = blockIdx.x * cols_per_cta;
  const int wcols = min(cols_per_cta, ncol - c0);
  const unsigned row_bytes = (unsigned)wcols * sizeof(cplx);
  const int l0 = s * lchunk;
  const int l1 = min(chi_l, l0 + lchunk);
  const int nl = max(0, l1 - l0);
  const cplx* tbase = t + (size_t)l0 * ncol + c0;
  if (threadIdx.x == 0) {
    for (int q = 0; q < DST; ++q) mbar_init(&full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ngroups = (nl + RPS - 1) / RPS;       // row groups = pipeline steps
  auto issue = [&](int grp) {                   // one thread: RPS bulk copies, one barrier
    const int q = grp % DST;
    const int r0 = grp * RPS;
    const int nr = min(RPS, nl - r0);
    mbar_expect_tx(&full[q], row_bytes * (unsigned)nr);
    for (int rr = 0; rr < nr; ++rr)
      bulk_load(stage + ((size_t)q * RPS + rr) * DW, tbase + (size_t)(r0 + rr) * ncol,
                row_bytes, &full[q]);
  };
  if (threadIdx.x == 0)                         // prologue: fill the pipeline
    for (int grp = 0; grp < DST && grp < ngroups; ++grp) issue(grp);
  // ---- u[e, l, x] = sum_i P1[e][x, i] v[e, l, i] for l in [l0, l1)  (overlaps the loads)
  for (int k = 0; k < EVT; ++k) {
    if (e0 + k >= nvec) break;
    const cplx* P = p1 + (size_t)(e0 + k) * d2 * d2;
    const cplx* V = v + ((size_t)(e0 + k) * chi_l + l0) * d2;
    for (int q = threadIdx.x; q < nl * d2; q += DT) {
      const int l = q / d2, x = q % d2;
      cplx acc = make_double2(0.0, 0.0);
      for (int i = 0; i < d2; ++i) acc = b200::cfma(P[x * d2 + i], V[l * d2 + i], acc);
      us[((size_t)k * lchunk + l) * d2 + x] = acc;
    }
  }
  __syncthreads();
  // ---- stream T: partial[s][e][c] = sum_{l in slice} T[l, c] u[e, l, x(c)]
  cplx acc[DCPT][EVT];
  int xs[DCPT];
#pragma unroll
  for (int j = 0; j < DCPT; ++j) {
    xs[j] = (c0 + threadIdx.x + j * DT) % d2;
#pragma unroll
    for (int k = 0; k < EVT; ++k) acc[j][k] = make_double2(0.0, 0.0);
  }
  for (int grp = 0; grp < ngroups; ++grp) {
    const int q = grp % DST;
    mbar_wait(&full[q], (unsigned)((grp / DST) & 1));
    const int r0 = grp * RPS;
    const int nr = min(RPS, nl - r0);
#pragma unroll
    for (int rr = 0; rr < RPS; ++rr) {
      if (rr < nr) {
        const cplx* row = stage + ((size_t)q * RPS + rr) * DW;
        const int l = r0 + rr;
#pragma unroll
        for (int j = 0; j < DCPT; ++j) {
          const int cc = threadIdx.x + j * DT;
          if (cc < wcols) {
            const cplx tv = row[cc];
#pragma unroll
            for (int k = 0; k < EVT; ++k)
              acc[j][k] = b200::cfma(tv, us[((size_t)k * lchunk + l) * d2 + xs[j]],
                                     acc[j][k]);
          }
        }
      }
    }
    __syncthreads();                            // everybody has consumed stage q
    if (threadIdx.x == 0 && grp + DST < ngroups) issue(grp + DST);
  }
#pragma unroll
  for (int j = 0; j < DCPT; ++j) {
    const int cc = threadIdx.x + j * DT;
    if (cc < wcols) {
#pragma unroll
      for (int k = 0; k < EVT; ++k)
        if (e0 + k < nvec)
          partial[((size_t)s * nvec + (e0 + k)) * ncol + c0 + cc] = acc[j][k];
    }
  }
  // ---- the last CTA of this (column block, member group) finishes the step
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* tk = tickets + (size_t)blockIdx.z * gridDim.x + blockIdx.x;
    const unsigned int prev = atomicAdd(tk, 1u);
    s_last = (prev == (unsigned int)(ns - 1));
    if (s_last) *tk = 0u;                       // self-resetting for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // w[e][c] = sum_s partial[s][e][c] (fixed order: deterministic), all loads of an entry
  // in flight at once; staged in shared memory (the pipeline buffers are free now), then
  // v_out[e][r, j] = sum_x P2[e][j, x] w[e][r, x]
  cplx* ws = stage;                                   // [wcols] per member, reused
  for (int k = 0; k < EVT; ++k) {
    if (e0 + k >= nvec) break;
    const cplx* pbase = partial + (size_t)(e0 + k) * ncol + c0;
    const size_t sstride = (size_t)nvec * ncol;
    for (int cc = threadIdx.x; cc < wcols; cc += DT) {
      cplx pv[16];
#pragma unroll
      for (int q = 0; q < 16; ++q)
        pv[q] = (q < ns) ? __ldcg(reinterpret_cast<const double2*>(pbase + q * sstride + cc))
                         : make_double2(0.0, 0.0);
      cplx w = make_double2(0.0, 0.0);
#pragma unroll
      for (int q = 0; q < 16; ++q) { w.x += pv[q].x; w.y += pv[q].y; }
      for (int q = 16; q < ns; ++q) {             // (ns <= 16 by construction)
        const cplx x = __ldcg(reinterpret_cast<const double2*>(pbase + q * sstride + cc));
        w.x += x.x; w.y += x.y;
      }
      ws[cc] = w;
    }
    __syncthreads();
    const cplx* P = p2 + (size_t)(e0 + k) * d2 * d2;
    for (int cc = threadIdx.x; cc < wcols; cc += DT) {
      const int rr = cc / d2, j = cc % d2;          // c0 is a multiple of d2
      cplx a = make_double2(0.0, 0.0);
      for (int x = 0; x < d2; ++x) a = b200::cfma(P[j * d2 + x], ws[rr * d2 + x], a);
      v_out[(size_t)(e0 + k) * ncol + c0 + cc] = a;
    }
    __syncthreads();
  }
}

// cap_out[l] = sum_{r,x} T[l,r,x] cap_next[r] tr2[x].  HBM-bound (T is read once): a CTA of
// 8 warps takes CAP_ROWS rows x CAP_SEG column segments, every lane keeps 8 independent
// 16-byte streaming loads in flight (the weights cap_next[r] tr2[x] are L1/L2 hits), the
// segments of a row meet in shared memory in a fixed order.
constexpr int CAP_ROWS = 2, CAP_SEG = 4, CAP_T = 32 * CAP_ROWS * CAP_SEG;
__global__ void __launch_bounds__(CAP_T)
caps_kernel(int chi_l, int chi_r, int d2, const cplx* __restrict__ t,
            const cplx* __restrict__ cap_next, const cplx* __restrict__ tr2,
            cplx* __restrict__ cap_out) {
  __shared__ cplx s_part[CAP_ROWS][CAP_SEG];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rloc = warp / CAP_SEG, seg = warp % CAP_SEG;
  const int l = blockIdx.x * CAP_ROWS + rloc;
  const int ncol = chi_r * d2;
  cplx acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
  if (l < chi_l) {
    const cplx* row = t + (size_t)l * ncol;
    // columns of this segment, in chunks of 8 x 32
    const int per = (((ncol + CAP_SEG - 1) / CAP_SEG) + 255) & ~255;
    const int cbeg = seg * per, cend = min(ncol, cbeg + per);
    for (int c0 = cbeg; c0 < cend; c0 += 256) {
      cplx tv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * 32 + lane;
        tv[u] = (c < cend) ? __ldcs(reinterpret_cast<const double2*>(row + c))
                           : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * 32 + lane;
        if (c < cend) {
          const cplx w = b200::cmul(cap_next[c / d2], tr2[c % d2]);
          if (u & 1) acc1 = b200::cfma(tv[u], w, acc1); else acc0 = b200::cfma(tv[u], w, acc0);
        }
      }
    }
  }
  cplx acc = make_double2(acc0.x + acc1.x, acc0.y + acc1.y);
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
  }
  if (lane == 0) s_part[rloc][seg] = acc;
  __syncthreads();
  if (threadIdx.x < CAP_ROWS) {
    const int lo = blockIdx.x * CAP_ROWS + threadIdx.x;
    if (lo < chi_l) {
      cplx tot = s_part[threadIdx.x][0];
      for (int q = 1; q < CAP_SEG; ++q) { tot.x += s_part[threadIdx.x][q].x; tot.y += s_part[threadIdx.x][q].y; }
      cap_out[lo] = tot;
    }
  }
}

constexpr size_t kTicketBytes = 64 * 1024;   // >= nx * ne tickets (checked)

struct DynGeom { int cols_per_cta, nx, ne, ev, ns, lchunk; size_t smem, part_bytes, ticket_bytes; };

inline DynGeom dyn_geom(int nvec, int chi_l, int chi_r, int d2) {
  DynGeom g;
  g.cols_per_cta = (DW / d2) * d2;                 // whole system-leg groups per CTA
  if (g.cols_per_cta < d2) g.cols_per_cta = d2;    // (d2 > DT is rejected by the caller)
  g.nx = (chi_r * d2 + g.cols_per_cta - 1) / g.cols_per_cta;
  g.ev = (nvec == 1) ? 1 : EV;
  g.ne = (nvec + g.ev - 1) / g.ev;
  // enough CTAs for ~8 per SM (E = 1): a streaming kernel needs the bytes in flight;
  // with several members per thread the split costs partial-sum traffic
  const int per_sm = 2;
  int ns = (148 * per_sm + g.nx * g.ne - 1) / (g.nx * g.ne);
  int max_ns = (chi_l + 31) / 32;            // at least 32 rows per CTA: a real pipeline
  if (max_ns > 16) max_ns = 16;              // the last CTA sums ns partials per entry
  if (ns > max_ns) ns = max_ns;
  if (ns < 1) ns = 1;
  g.lchunk = (chi_l + ns - 1) / ns;
  while ((size_t)g.lchunk * d2 * g.ev * sizeof(cplx) > 40 * 1024) {   // u must fit in smem
    ++ns;
    g.lchunk = (chi_l + ns - 1) / ns;
  }
  g.ns = (chi_l + g.lchunk - 1) / g.lchunk;
  g.smem = (size_t)DST * RPS * DW * sizeof(cplx) + (size_t)g.lchunk * d2 * g.ev * sizeof(cplx);
  g.part_bytes = (((size_t)g.ns * nvec * chi_r * d2 * sizeof(cplx)) + 255) & ~(size_t)255;
  g.ticket_bytes = kTicketBytes;                   // fixed: tickets live at offset 0
  return g;
}

}  // namespace

extern "C" size_t b200_dyn_workspace_bytes(int nvec, int chi_l, int chi_r, int d2) {
  if (nvec <= 0 || chi_l <= 0 || chi_r <= 0 || d2 <= 0 || d2 > DT) return 0;
  const DynGeom g = dyn_geom(nvec, chi_l, chi_r, d2);
  return g.part_bytes + g.ticket_bytes;
}

static int dyn_step_impl(void* stream_, int nvec, int chi_l, int chi_r, int d2,
                         const void* t, const void* p1, const void* p2, const void* v,
                         void* v_out, const void* cap, void* rho_out, void* work,
                         bool clear_tickets) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (nvec <= 0 || chi_l <= 0 || chi_r <= 0 || d2 <= 0 || !t || !p1 || !p2 || !v ||
      !v_out || !work) {
    b200::set_error("b200_dyn_step: invalid argument");
    return B200_EINVAL;
  }
  if (nvec > 65535 * EV || d2 > DT) {
    b200::set_error("b200_dyn_step: nvec or d2 too large");
    return B200_ESIZE;
  }
  const DynGeom g = dyn_geom(nvec, chi_l, chi_r, d2);
  if ((size_t)g.nx * g.ne * sizeof(unsigned int) > kTicketBytes) {
    b200::set_error("b200_dyn_step: too many column blocks");
    return B200_ESIZE;
  }
  unsigned int* tickets = (unsigned int*)work;
  cplx* partial = (cplx*)((unsigned char*)work + kTicketBytes);
  {   // once per process, safe against concurrent first calls (ensemble worker threads)
    static std::once_flag once;
    static cudaError_t attr_rc = cudaSuccess;
    std::call_once(once, [] {
      attr_rc = cudaFuncSetAttribute(dyn_fused_kernel<1>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (attr_rc == cudaSuccess)
        attr_rc = cudaFuncSetAttribute(dyn_fused_kernel<EV>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    B200_CUDA_CHECK(attr_rc);
  }
  // tickets are self-resetting, but `work` may be fresh memory
  if (clear_tickets) B200_CUDA_CHECK(cudaMemsetAsync(tickets, 0, kTicketBytes, stream));
  const dim3 grid(g.nx, g.ns + 1, g.ne);     // +1: the read-out CTA row
  if (g.ev == 1)
    dyn_fused_kernel<1><<<grid, DT, g.smem, stream>>>(
        nvec, chi_l, chi_r, d2, g.lchunk, g.cols_per_cta, (const cplx*)t, (const cplx*)p1,
        (const cplx*)p2, (const cplx*)v, (cplx*)v_out, (const cplx*)cap, (cplx*)rho_out,
        partial, tickets);
  else
    dyn_fused_kernel<EV><<<grid, DT, g.smem, stream>>>(
        nvec, chi_l, chi_r, d2, g.lchunk, g.cols_per_cta, (const cplx*)t, (const cplx*)p1,
        (const cplx*)p2, (const cplx*)v, (cplx*)v_out, (const cplx*)cap, (cplx*)rho_out,
        partial, tickets);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_dyn_step(void* stream_, int nvec, int chi_l, int chi_r, int d2,
                             const void* t, const void* p1, const void* p2,
                             const void* v, void* v_out, const void* cap,
                             void* rho_out, void* work) {
  return dyn_step_impl(stream_, nvec, chi_l, chi_r, d2, t, p1, p2, v, v_out, cap, rho_out,
                       work, true);
}

extern "C" int b200_caps_step(void* stream_, int chi_l, int chi_r, int d2,
                              const void* t, const void* cap_next, const void* tr2,
                              void* cap_out) {
  if (chi_l <= 0 || chi_r <= 0 || d2 <= 0 || !t || !cap_next || !tr2 || !cap_out) {
    b200::set_error("b200_caps_step: invalid argument");
    return B200_EINVAL;
  }
  const int blocks = (chi_l + CAP_ROWS - 1) / CAP_ROWS;
  caps_kernel<<<blocks, CAP_T, 0, (cudaStream_t)stream_>>>(
      chi_l, chi_r, d2, (const cplx*)t, (const cplx*)cap_next, (const cplx*)tr2,
      (cplx*)cap_out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ---------------------------------------------------------------------------
// The whole compute_dynamics loop (system_dynamics.py:131-170) in ONE call: the per-step
// launches are issued back to back from C, so the stream never waits for the interpreter
// and the PT-MPO sites stream from HBM at close to copy speed.
namespace {
struct DynRunLayout { size_t va, vb, step, total; };
DynRunLayout dyn_run_layout(int nsteps, int nvec, const int32_t* chi, int d2) {
  int cmax = 1;
  size_t step_max = 0;
  for (int k = 0; k <= nsteps; ++k) if (chi[k] > cmax) cmax = chi[k];
  for (int k = 0; k < nsteps; ++k) {
    const size_t b = b200_dyn_workspace_bytes(nvec, chi[k], chi[k + 1], d2);
    if (b > step_max) step_max = b;
  }
  // the final read-out (dyn_pre_kernel) needs room for u = P1 v of the last state
  const size_t fin = (size_t)nvec * chi[nsteps] * d2 * sizeof(cplx) + 256;
  if (fin > step_max) step_max = fin;
  DynRunLayout L;
  const size_t vbytes = (((size_t)nvec * cmax * d2 * sizeof(cplx)) + 255) & ~(size_t)255;
  L.va = 0; L.vb = vbytes; L.step = 2 * vbytes; L.total = 2 * vbytes + step_max;
  return L;
}
}  // namespace

extern "C" size_t b200_dyn_run_workspace_bytes(int nsteps, int nvec, const int32_t* chi,
                                               int d2) {
  if (nsteps <= 0 || nvec <= 0 || d2 <= 0 || !chi) return 0;
  return dyn_run_layout(nsteps, nvec, chi, d2).total;
}

extern "C" int b200_dyn_run(void* stream_, int nsteps, int nvec, int d2, const int32_t* chi,
                            const void* const* t, const void* p1, const void* p2,
                            int64_t prop_step_stride, const void* const* caps,
                            const void* v0, void* rho_out, void* work) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (nsteps <= 0 || nvec <= 0 || d2 <= 0 || !chi || !t || !p1 || !p2 || !caps || !v0 ||
      !rho_out || !work) {
    b200::set_error("b200_dyn_run: invalid argument");
    return B200_EINVAL;
  }
  const DynRunLayout L = dyn_run_layout(nsteps, nvec, chi, d2);
  unsigned char* base = (unsigned char*)work;
  cplx* bufs[2] = {(cplx*)(base + L.va), (cplx*)(base + L.vb)};
  void* step_work = base + L.step;
  const cplx* v = (const cplx*)v0;
  cplx* rho = (cplx*)rho_out;
  const cplx* P1 = (const cplx*)p1;
  const cplx* P2 = (const cplx*)p2;
  B200_CUDA_CHECK(cudaMemsetAsync(step_work, 0, kTicketBytes, stream));
  for (int k = 0; k < nsteps; ++k) {
    cplx* vo = bufs[k & 1];
    const int rc = dyn_step_impl(stream_, nvec, chi[k], chi[k + 1], d2, t[k],
                                 P1 + (size_t)k * prop_step_stride,
                                 P2 + (size_t)k * prop_step_stride, v, vo, caps[k],
                                 rho + (size_t)k * nvec * d2, step_work, false);
    if (rc != B200_OK) return rc;
    v = vo;
  }
  {   // final read-out rho[N] = sum_l cap_N[l] v[l]  (system_dynamics.py:167-170)
    const int chi_l = chi[nsteps];
    int bx = (chi_l * d2 + 255) / 256;
    if (bx > 64) bx = 64;
    dyn_pre_kernel<<<dim3(bx, nvec), 256, 0, stream>>>(
        nvec, chi_l, d2, P1, v, (cplx*)step_work, (const cplx*)caps[nsteps],
        rho + (size_t)nsteps * nvec * d2);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}
