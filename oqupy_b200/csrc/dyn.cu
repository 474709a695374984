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

// partial[s][e][r][x] = sum_{l in slice s} T[l,r,x] * u[e,l,x]
__global__ void __launch_bounds__(DT)
dyn_stream_kernel(int nvec, int chi_l, int chi_r, int d2, int lchunk,
                  const cplx* __restrict__ t, const cplx* __restrict__ u,
                  cplx* __restrict__ partial) {
  const int ncol = chi_r * d2;
  const int c = blockIdx.x * DT + threadIdx.x;
  const int s = blockIdx.y;
  const int e0 = blockIdx.z * EV;
  if (c >= ncol) return;
  const int x = c % d2;
  const int l0 = s * lchunk;
  const int l1 = min(chi_l, l0 + lchunk);
  cplx acc[EV];
#pragma unroll
  for (int k = 0; k < EV; ++k) acc[k] = make_double2(0.0, 0.0);
  const cplx* tp = t + (size_t)l0 * ncol + c;
#pragma unroll 4
  for (int l = l0; l < l1; ++l, tp += ncol) {
    const cplx tv = __ldg(tp);
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      if (e0 + k < nvec) {
        const cplx uv = u[((size_t)(e0 + k) * chi_l + l) * d2 + x];
        acc[k] = b200::cfma(tv, uv, acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < EV; ++k)
    if (e0 + k < nvec)
      partial[((size_t)s * nvec + (e0 + k)) * ncol + c] = acc[k];
}

// v_out[e,r,j] = sum_x P2[e][j,x] * sum_s partial[s][e][r][x]
__global__ void dyn_post_kernel(int nvec, int chi_r, int d2, int ns,
                                const cplx* __restrict__ p2,
                                const cplx* __restrict__ partial,
                                cplx* __restrict__ v_out) {
  const int e = blockIdx.y;
  const int ncol = chi_r * d2;
  const cplx* P = p2 + (size_t)e * d2 * d2;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol;
       c += gridDim.x * blockDim.x) {
    const int r = c / d2, j = c % d2;
    cplx acc = make_double2(0.0, 0.0);
    for (int x = 0; x < d2; ++x) {
      cplx w = make_double2(0.0, 0.0);
      for (int s = 0; s < ns; ++s) {
        const cplx pv = partial[((size_t)s * nvec + e) * ncol + r * d2 + x];
        w.x += pv.x; w.y += pv.y;
      }
      acc = b200::cfma(P[j * d2 + x], w, acc);
    }
    v_out[(size_t)e * ncol + c] = acc;
  }
}

// cap_out[l] = sum_{r,x} T[l,r,x] cap_next[r] tr2[x] ; one warp per l
__global__ void caps_kernel(int chi_l, int chi_r, int d2, const cplx* __restrict__ t,
                            const cplx* __restrict__ cap_next,
                            const cplx* __restrict__ tr2, cplx* __restrict__ cap_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= chi_l) return;
  const int ncol = chi_r * d2;
  const cplx* row = t + (size_t)warp * ncol;
  cplx acc = make_double2(0.0, 0.0);
  for (int c = lane; c < ncol; c += 32) {
    const cplx w = b200::cmul(cap_next[c / d2], tr2[c % d2]);
    acc = b200::cfma(row[c], w, acc);
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
  }
  if (lane == 0) cap_out[warp] = acc;
}

inline int dyn_splits(int nvec, int chi_l, int chi_r, int d2) {
  const int nx = (chi_r * d2 + DT - 1) / DT;
  const int ne = (nvec + EV - 1) / EV;
  // enough CTAs for ~8 per SM: a streaming kernel needs the bytes in flight
  const int per_sm = (ne == 1) ? 8 : 4;   // (E > 1: the split costs partial-sum traffic)
  int ns = (148 * per_sm + nx * ne - 1) / (nx * ne);
  const int max_ns = (chi_l + 15) / 16;
  if (ns > max_ns) ns = max_ns;
  if (ns < 1) ns = 1;
  return ns;
}

}  // namespace

extern "C" size_t b200_dyn_workspace_bytes(int nvec, int chi_l, int chi_r, int d2) {
  if (nvec <= 0 || chi_l <= 0 || chi_r <= 0 || d2 <= 0) return 0;
  const int ns = dyn_splits(nvec, chi_l, chi_r, d2);
  const size_t u = (size_t)nvec * chi_l * d2 * sizeof(cplx);
  const size_t part = (size_t)ns * nvec * chi_r * d2 * sizeof(cplx);
  return ((u + 255) & ~(size_t)255) + part;
}

extern "C" int b200_dyn_step(void* stream_, int nvec, int chi_l, int chi_r, int d2,
                             const void* t, const void* p1, const void* p2,
                             const void* v, void* v_out, const void* cap,
                             void* rho_out, void* work) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (nvec <= 0 || chi_l <= 0 || chi_r <= 0 || d2 <= 0 || !t || !p1 || !p2 || !v ||
      !v_out || !work) {
    b200::set_error("b200_dyn_step: invalid argument");
    return B200_EINVAL;
  }
  if (nvec > 65535) {
    b200::set_error("b200_dyn_step: nvec too large");
    return B200_ESIZE;
  }
  const int ns = dyn_splits(nvec, chi_l, chi_r, d2);
  const size_t u_bytes = (((size_t)nvec * chi_l * d2 * sizeof(cplx)) + 255) & ~(size_t)255;
  cplx* u = (cplx*)work;
  cplx* partial = (cplx*)((unsigned char*)work + u_bytes);
  {
    int bx = (chi_l * d2 + 255) / 256;
    if (bx > 64) bx = 64;
    dyn_pre_kernel<<<dim3(bx, nvec), 256, 0, stream>>>(
        nvec, chi_l, d2, (const cplx*)p1, (const cplx*)v, u, (const cplx*)cap,
        (cplx*)rho_out);
    B200_LAUNCH_CHECK();
  }
  {
    const int nx = (chi_r * d2 + DT - 1) / DT;
    const int ne = (nvec + EV - 1) / EV;
    const int lchunk = (chi_l + ns - 1) / ns;
    dyn_stream_kernel<<<dim3(nx, ns, ne), DT, 0, stream>>>(
        nvec, chi_l, chi_r, d2, lchunk, (const cplx*)t, u, partial);
    B200_LAUNCH_CHECK();
  }
  {
    int bx = (chi_r * d2 + 255) / 256;
    if (bx > 64) bx = 64;
    dyn_post_kernel<<<dim3(bx, nvec), 256, 0, stream>>>(
        nvec, chi_r, d2, ns, (const cplx*)p2, partial, (cplx*)v_out);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

extern "C" int b200_caps_step(void* stream_, int chi_l, int chi_r, int d2,
                              const void* t, const void* cap_next, const void* tr2,
                              void* cap_out) {
  if (chi_l <= 0 || chi_r <= 0 || d2 <= 0 || !t || !cap_next || !tr2 || !cap_out) {
    b200::set_error("b200_caps_step: invalid argument");
    return B200_EINVAL;
  }
  const int blocks = (chi_l * 32 + 255) / 256;
  caps_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(
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
  const size_t fin = b200_dyn_workspace_bytes(nvec, chi[nsteps], 1, d2);
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
  for (int k = 0; k < nsteps; ++k) {
    cplx* vo = bufs[k & 1];
    const int rc = b200_dyn_step(stream_, nvec, chi[k], chi[k + 1], d2, t[k],
                                 P1 + (size_t)k * prop_step_stride,
                                 P2 + (size_t)k * prop_step_stride, v, vo, caps[k],
                                 rho + (size_t)k * nvec * d2, step_work);
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
