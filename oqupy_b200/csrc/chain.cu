// Native matrix-product chain engine (host C++): the device-resident counterpart of
// NodeArray (oqupy/backends/node_array.py:226-299, 413-554) for the PT-TEMPO path.
//
// A chain owns its sites in device memory (stream-ordered pool: cudaMallocAsync) and runs
// a whole zip-up or svd-sweep in ONE C call: per site it launches the contraction GEMM
// (b200_zgemm_strided), the truncated SVD (b200_svd_factor), waits for `keep` (the one
// host round trip a dependent chain of exact-size sites needs), allocates the exact-size
// outputs and launches b200_svd_emit.  Doing the loop here instead of in Python removes
// ~70 us of interpreter time per SVD from a ~400-SVD dependency chain.
#include <stdlib.h>

#include <atomic>
#include <map>
#include <vector>

#include "common.cuh"

namespace {

struct Site {
  cplx* p;
  int dl, da, dr;       // (chi_l, array leg, chi_r), contiguous complex128
};

// Host-side first-fit allocator over ONE device arena.  Everything a chain allocates is
// used on its single stream, so a freed range may be handed out again at once (the kernels
// that read it are ordered before the kernels that will overwrite it): no CUDA call per
// allocation -- ten cudaMallocAsync / cudaFreeAsync per truncated SVD were the largest
// host cost of a step.  Requests that do not fit fall back to the stream-ordered pool.
struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  std::map<size_t, size_t> free_;     // offset -> size, coalesced
  void init(char* b, size_t c) { base = b; cap = c; free_.clear(); if (c) free_[0] = c; }
  void* alloc(size_t n) {
    n = (n + 255) & ~(size_t)255;
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= n) {
        const size_t off = it->first, rest = it->second - n;
        free_.erase(it);
        if (rest) free_[off + n] = rest;
        return base + off;
      }
    }
    return nullptr;
  }
  bool owns(const void* p) const { return base && (const char*)p >= base && (const char*)p < base + cap; }
  void release(void* p, size_t n) {
    n = (n + 255) & ~(size_t)255;
    size_t off = (size_t)((char*)p - base);
    auto nxt = free_.lower_bound(off);
    if (nxt != free_.begin()) {
      auto prv = std::prev(nxt);
      if (prv->first + prv->second == off) { off = prv->first; n += prv->second; free_.erase(prv); }
    }
    if (nxt != free_.end() && off + n == nxt->first) { n += nxt->second; free_.erase(nxt); }
    free_[off] = n;
  }
};

struct Chain {
  std::vector<Arena> arenas;            // grows by doubling (64 MB, 128 MB, ...)
  size_t next_arena_bytes;
  std::map<const void*, size_t> live;   // arena allocations -> size
  cudaStream_t stream;
  std::vector<Site> sites;
  void* work;
  size_t work_bytes;
  int32_t* info;        // pinned, 8 words: {keep, sweeps, status, rotations, qr pivots, ...}
  cplx* one;            // device scalar 1
  // statistics since the last reset
  uint64_t nsvd, sweeps, d2h_bytes;
  std::vector<int32_t> log;   // (m, n, keep, sweeps) per SVD when enabled
  bool log_on;
};

int dev_alloc(Chain* c, cplx** out, size_t n) {
  if (n == 0) n = 1;
  const size_t bytes = n * sizeof(cplx);
  for (int attempt = 0; attempt < 2; ++attempt) {
    for (auto& a : c->arenas) {
      if (void* p = a.alloc(bytes)) {
        c->live[p] = bytes;
        *out = (cplx*)p;
        return B200_OK;
      }
    }
    // no room: add a chunk (a rare, synchronous cudaMalloc), at most 10 of them
    if (attempt || c->next_arena_bytes == 0 || c->arenas.size() >= 16) break;
    size_t want = c->next_arena_bytes;
    while (want < 2 * bytes) want *= 2;
    void* p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) { (void)cudaGetLastError(); break; }
    Arena a;
    a.init((char*)p, want);
    c->arenas.push_back(a);
    // every new chunk is a blocking cudaMalloc: grow fast (x4) up to 16 GB chunks
    c->next_arena_bytes = (want >= ((size_t)16 << 30)) ? want : want * 4;
  }
  B200_CUDA_CHECK(cudaMallocAsync((void**)out, bytes, c->stream));
  return B200_OK;
}
Arena* arena_of(Chain* c, const void* p) {
  for (auto& a : c->arenas) if (a.owns(p)) return &a;
  return nullptr;
}
int dev_free(Chain* c, cplx* p) {
  if (!p) return B200_OK;
  if (Arena* a = arena_of(c, p)) {
    auto it = c->live.find(p);
    if (it != c->live.end()) { a->release(p, it->second); c->live.erase(it); }
    return B200_OK;
  }
  B200_CUDA_CHECK(cudaFreeAsync(p, c->stream));
  return B200_OK;
}

#define B200_TRY(expr)                \
  do {                                \
    const int _rc = (expr);           \
    if (_rc != B200_OK) return _rc;   \
  } while (0)

b200_operand op(const cplx* p, int64_t row, int64_t col, int64_t b1 = 0, int64_t b2 = 0) {
  b200_operand o;
  o.ptr = p; o.row = row; o.col = col; o.b1 = b1; o.b2 = b2; o.conj = 0;
  return o;
}

int gemm(Chain* c, int m, int n, int k, int nb1, int nb2, const b200_operand& a,
         const b200_operand& b, cplx* out, int64_t c_row, int64_t c_col, int64_t c_b1,
         int64_t c_b2, const cplx* scale = nullptr, int64_t s_b1 = 0, int64_t s_b2 = 0) {
  return b200_zgemm_strided(c->stream, m, n, k, nb1, nb2, &a, &b, out, c_row, c_col, c_b1,
                            c_b2, scale, s_b1, s_b2, 0);
}

// truncated SVD of theta (m x n, element strides rs, cs); returns keep
int split(Chain* c, const cplx* theta, int m, int n, int64_t rs, int64_t cs, double eps,
          int* keep) {
  const size_t need = b200_svd_workspace_bytes(m, n);
  if (need > c->work_bytes) {
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->work) B200_CUDA_CHECK(cudaFree(c->work));
    c->work = nullptr;
    c->work_bytes = 0;
    const size_t want = 2 * need;       // geometric growth: a reallocation synchronises
    B200_CUDA_CHECK(cudaMalloc(&c->work, want));
    c->work_bytes = want;
  }
  // `keep` comes back through pinned host memory: the kernel writes info[0..2], a system
  // fence, then info[3] (>= 0).  Spinning on that word costs ~1 us; a blocking stream
  // synchronisation costs a scheduler wake-up per SVD.
  volatile int32_t* vinfo = c->info;
  vinfo[3] = -1;
  B200_TRY(b200_svd_factor(c->stream, theta, m, n, rs, cs, eps, c->work, c->info));
  for (uint64_t it = 1;; ++it) {
    if (vinfo[3] != -1) break;
    if ((it & 0x3FFF) == 0) {            // a faulted kernel never writes the word
      const cudaError_t q = cudaStreamQuery(c->stream);
      if (q == cudaSuccess) break;
      if (q != cudaErrorNotReady) B200_CUDA_CHECK(q);
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (vinfo[3] == -1) B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->d2h_bytes += 16;
  ++c->nsvd;
  c->sweeps += (uint64_t)c->info[1];
  if (c->log_on) {
    c->log.push_back(m); c->log.push_back(n);
    c->log.push_back(c->info[0]); c->log.push_back(c->info[1]);
  }
  if (c->info[2] != 0) {
    b200::set_error("Jacobi SVD did not converge (%dx%d, %d sweeps)", m, n, c->info[1]);
    return B200_ENOCONV;
  }
  *keep = c->info[0];
  return B200_OK;
}

int emit(Chain* c, int m, int n, int keep, cplx* u, int u_na, int64_t u_so, int64_t u_sa,
         int64_t u_sj, cplx* svh) {
  return b200_svd_emit(c->stream, c->work, nullptr, m, n, 0, 0, keep, u, u_na, u_so, u_sa,
                       u_sj, svh);
}

}  // namespace

extern "C" {

void* b200_chain_create(void* stream) {
  Chain* c = new Chain();
  c->stream = (cudaStream_t)stream;
  c->work = nullptr;
  c->work_bytes = 0;
  c->info = nullptr;
  c->one = nullptr;
  c->nsvd = c->sweeps = c->d2h_bytes = 0;
  c->log_on = false;
  if (cudaMallocHost((void**)&c->info, 8 * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc((void**)&c->one, sizeof(cplx)) != cudaSuccess) {
    b200::set_error("b200_chain_create: allocation failed");
    delete c;
    return nullptr;
  }
  {   // keep freed site memory cached in the stream-ordered pool across the per-SVD syncs
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep_all = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all);
    }
  }
  {   // optional: reserve pool memory up front (B200_POOL_RESERVE_MB; measured: no gain
      // for the config-2 build, the pool's release threshold above is what matters)
    size_t mb = 0;
    if (const char* e = getenv("B200_POOL_RESERVE_MB")) mb = (size_t)atoll(e);
    void* p = nullptr;
    if (mb > 0 && cudaMallocAsync(&p, mb << 20, c->stream) == cudaSuccess)
      cudaFreeAsync(p, c->stream);
    else
      (void)cudaGetLastError();
  }
  {   // device arenas for the sites and temporaries: first chunk on first use
      // (B200_CHAIN_ARENA_MB, default 64; 0 = stream-ordered pool only)
    size_t mb = 64;
    if (const char* e = getenv("B200_CHAIN_ARENA_MB")) mb = (size_t)atoll(e);
    c->next_arena_bytes = mb << 20;
  }
  const cplx h1 = make_double2(1.0, 0.0);
  cudaMemcpyAsync(c->one, &h1, sizeof(cplx), cudaMemcpyHostToDevice, c->stream);
  cudaStreamSynchronize(c->stream);
  return c;
}

int b200_chain_destroy(void* h) {
  Chain* c = (Chain*)h;
  if (!c) return B200_OK;
  cudaStreamSynchronize(c->stream);
  for (auto& s : c->sites) if (s.p && !arena_of(c, s.p)) cudaFreeAsync(s.p, c->stream);
  cudaStreamSynchronize(c->stream);
  for (auto& a : c->arenas) if (a.base) cudaFree(a.base);
  if (c->work) cudaFree(c->work);
  if (c->info) cudaFreeHost(c->info);
  if (c->one) cudaFree(c->one);
  delete c;
  return B200_OK;
}

int b200_chain_len(void* h) { return h ? (int)((Chain*)h)->sites.size() : 0; }

int b200_chain_push(void* h, const void* dev_src, int dl, int da, int dr) {
  Chain* c = (Chain*)h;
  if (!c || !dev_src || dl < 1 || da < 1 || dr < 1) {
    b200::set_error("b200_chain_push: invalid argument");
    return B200_EINVAL;
  }
  Site s;
  s.dl = dl; s.da = da; s.dr = dr;
  const size_t n = (size_t)dl * da * dr;
  B200_TRY(dev_alloc(c, &s.p, n));
  B200_CUDA_CHECK(cudaMemcpyAsync(s.p, dev_src, n * sizeof(cplx), cudaMemcpyDeviceToDevice,
                                  c->stream));
  c->sites.push_back(s);
  return B200_OK;
}

int b200_chain_shape(void* h, int i, int32_t* out3) {
  Chain* c = (Chain*)h;
  if (!c || !out3 || i < 0 || i >= (int)c->sites.size()) {
    b200::set_error("b200_chain_shape: invalid argument");
    return B200_EINVAL;
  }
  out3[0] = c->sites[i].dl; out3[1] = c->sites[i].da; out3[2] = c->sites[i].dr;
  return B200_OK;
}

int b200_chain_read(void* h, int i, void* dev_dst) {
  Chain* c = (Chain*)h;
  if (!c || !dev_dst || i < 0 || i >= (int)c->sites.size()) {
    b200::set_error("b200_chain_read: invalid argument");
    return B200_EINVAL;
  }
  const Site& s = c->sites[i];
  B200_CUDA_CHECK(cudaMemcpyAsync(dev_dst, s.p, (size_t)s.dl * s.da * s.dr * sizeof(cplx),
                                  cudaMemcpyDeviceToDevice, c->stream));
  return B200_OK;
}

int b200_chain_stats(void* h, uint64_t* nsvd, uint64_t* sweeps, uint64_t* d2h_bytes,
                     int reset) {
  Chain* c = (Chain*)h;
  if (!c) return B200_EINVAL;
  if (nsvd) *nsvd = c->nsvd;
  if (sweeps) *sweeps = c->sweeps;
  if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
  if (reset) c->nsvd = c->sweeps = c->d2h_bytes = 0;
  return B200_OK;
}

int b200_chain_log(void* h, int enable, int32_t* out, int cap_entries) {
  Chain* c = (Chain*)h;
  if (!c) return B200_EINVAL;
  int n = (int)(c->log.size() / 4);
  if (out) {
    if (n > cap_entries) n = cap_entries;
    for (int i = 0; i < 4 * n; ++i) out[i] = c->log[i];
    c->log.clear();
  }
  c->log_on = (enable != 0);
  return n;
}

// svd_sweep(from, to)  (oqupy/backends/node_array.py:226-299)
int b200_chain_svd_sweep(void* h, int from, int to, double eps) {
  Chain* c = (Chain*)h;
  if (!c) return B200_EINVAL;
  const int n = (int)c->sites.size();
  if (from < 0) from += n;
  if (to < 0) to += n;
  if (from < 0 || from >= n || to < 0 || to >= n) {
    b200::set_error("b200_chain_svd_sweep: index out of range");
    return B200_EINVAL;
  }
  if (from < to) {          // left -> right: rows (l, a), cols r  (:252-273)
    for (int i = from; i < to; ++i) {
      Site a = c->sites[i];
      const int m = a.dl * a.da, nr = a.dr;
      int nj = 0;
      B200_TRY(split(c, a.p, m, nr, nr, 1, eps, &nj));
      cplx *u = nullptr, *svh = nullptr;
      B200_TRY(dev_alloc(c, &u, (size_t)m * nj));
      B200_TRY(dev_alloc(c, &svh, (size_t)nj * nr));
      B200_TRY(emit(c, m, nr, nj, u, 1, nj, 0, 1, svh));
      Site b = c->sites[i + 1];
      const int bn = b.da * b.dr;
      cplx* nb = nullptr;
      B200_TRY(dev_alloc(c, &nb, (size_t)nj * bn));
      B200_TRY(gemm(c, nj, bn, nr, 1, 1, op(svh, nr, 1), op(b.p, bn, 1), nb, bn, 1, 0, 0));
      B200_TRY(dev_free(c, a.p));
      B200_TRY(dev_free(c, b.p));
      B200_TRY(dev_free(c, svh));
      c->sites[i] = Site{u, a.dl, a.da, nj};
      c->sites[i + 1] = Site{nb, nj, b.da, b.dr};
    }
  } else {                  // right -> left: rows (a, r), cols l  (:274-296)
    for (int i = from; i > to; --i) {
      Site a = c->sites[i];
      const int m = a.da * a.dr, nl = a.dl;
      int nj = 0;
      B200_TRY(split(c, a.p, m, nl, 1, m, eps, &nj));
      cplx *u = nullptr, *svh = nullptr;
      B200_TRY(dev_alloc(c, &u, (size_t)m * nj));
      B200_TRY(dev_alloc(c, &svh, (size_t)nj * nl));
      B200_TRY(emit(c, m, nl, nj, u, 1, 1, 0, m, svh));
      Site b = c->sites[i - 1];
      const int bm = b.dl * b.da;
      cplx* nb = nullptr;
      B200_TRY(dev_alloc(c, &nb, (size_t)bm * nj));
      B200_TRY(gemm(c, bm, nj, nl, 1, 1, op(b.p, nl, 1), op(svh, 1, nl), nb, nj, 1, 0, 0));
      B200_TRY(dev_free(c, a.p));
      B200_TRY(dev_free(c, b.p));
      B200_TRY(dev_free(c, svh));
      c->sites[i] = Site{u, nj, a.da, a.dr};
      c->sites[i - 1] = Site{nb, b.dl, b.da, nj};
    }
  }
  return B200_OK;
}

// mps.zip_up(mpo, right_index=-1, direction="left") with the implicit PT-TEMPO MPO
// (oqupy/backends/node_array.py:482-552, pt_tempo_backend.py:142-152):
//   Theta[(k,y),(l,e)] = M[e,y] sum_r carry[k,r,e] A[l,y,r];
//   new site = U as (j, y, k), carry' = S Vh as (j, l, e).
int b200_chain_pt_zip_up_left(void* h, const b200_pt_site* mpo, int nb, double eps) {
  Chain* c = (Chain*)h;
  if (!c || !mpo || nb < 1 || nb > (int)c->sites.size()) {
    b200::set_error("b200_chain_pt_zip_up_left: invalid argument");
    return B200_EINVAL;
  }
  const int left = (int)c->sites.size() - nb;
  cplx* carry = nullptr;
  int ck = 0, cr = 0, ce = 0;            // carry (k, r, e)
  for (int ib = nb - 1; ib >= 0; --ib) {
    const int ia = left + ib;
    Site a = c->sites[ia];
    const int nl = a.dl, nx = a.da, nr = a.dr;
    const b200_pt_site& site = mpo[ib];
    const cplx* mat = (const cplx*)site.mat;
    if (site.kind == B200_PT_FIRST && site.north_map && site.west_map) {
      // unique=True (pt_tempo_backend.py:125-137): B[w,y,n] = [w=wmap[y]] [n=nmap[y]] vec[n];
      // one small product per array value y
      const int ny = site.cols;
      const int32_t *nm = site.north_map, *wm = site.west_map;
      for (int y = 0; y < ny; ++y)
        if (wm[y] < 0 || wm[y] >= nx || nm[y] < 0 || nm[y] >= site.rows ||
            (carry && nm[y] >= ce)) {
          b200::set_error("pt_zip_up_left: degeneracy map out of range");
          return B200_EINVAL;
        }
      cplx* out = nullptr;
      if (!carry) {       // out[l,y,0] = vec[n(y)] A[l,w(y),0]
        if (nr != 1) { b200::set_error("pt_zip_up_left: bad first site"); return B200_EINVAL; }
        B200_TRY(dev_alloc(c, &out, (size_t)nl * ny));
        for (int y = 0; y < ny; ++y)
          B200_TRY(gemm(c, nl, 1, 1, 1, 1, op(a.p + (size_t)wm[y] * nr, (int64_t)nx * nr, 0),
                        op(c->one, 0, 0), out + y, ny, 0, 0, 0, mat + nm[y], 0, 0));
        c->sites[ia] = Site{out, nl, ny, 1};
      } else {            // out[l,y,k] = vec[n(y)] sum_r A[l,w(y),r] C[k,r,n(y)]
        if (cr != nr) { b200::set_error("pt_zip_up_left: carry mismatch"); return B200_EINVAL; }
        B200_TRY(dev_alloc(c, &out, (size_t)nl * ny * ck));
        for (int y = 0; y < ny; ++y)
          B200_TRY(gemm(c, nl, ck, nr, 1, 1, op(a.p + (size_t)wm[y] * nr, (int64_t)nx * nr, 1),
                        op(carry + nm[y], ce, (int64_t)nr * ce), out + (size_t)y * ck,
                        (int64_t)ny * ck, 1, 0, 0, mat + nm[y], 0, 0));
        c->sites[ia] = Site{out, nl, ny, ck};
      }
      B200_TRY(dev_free(c, a.p));
      if (ib != 0) { b200::set_error("pt_zip_up_left: 'first' site must be leftmost"); return B200_EINVAL; }
      break;
    }
    if (site.kind == B200_PT_FIRST) {
      const int ny = nx;
      cplx* out = nullptr;
      if (!carry) {       // closed single-site MPO: out[l,y,0] = vec[y] A[l,y,0]
        if (nr != 1) { b200::set_error("pt_zip_up_left: bad first site"); return B200_EINVAL; }
        B200_TRY(dev_alloc(c, &out, (size_t)nl * ny));
        B200_TRY(gemm(c, nl, 1, 1, ny, 1, op(a.p, (int64_t)nx * nr, 0, nr), op(c->one, 0, 0),
                      out, ny, 0, 1, 0, mat, 1, 0));
        c->sites[ia] = Site{out, nl, ny, 1};
      } else {            // out[l,y,k] = vec[y] sum_r A[l,y,r] C[k,r,y]
        if (ce != ny || cr != nr) { b200::set_error("pt_zip_up_left: carry mismatch"); return B200_EINVAL; }
        B200_TRY(dev_alloc(c, &out, (size_t)nl * ny * ck));
        B200_TRY(gemm(c, nl, ck, nr, ny, 1, op(a.p, (int64_t)nx * nr, 1, nr),
                      op(carry, ce, (int64_t)nr * ce, 1), out, (int64_t)ny * ck, 1, ck, 0,
                      mat, 1, 0));
        c->sites[ia] = Site{out, nl, ny, ck};
      }
      B200_TRY(dev_free(c, a.p));
      if (ib != 0) { b200::set_error("pt_zip_up_left: 'first' site must be leftmost"); return B200_EINVAL; }
      break;
    }
    const int ne = site.rows, ny = site.cols;
    int nk = 0;
    cplx* theta = nullptr;
    if (!carry) {         // newest site: Theta[0,y,l,e] = M[e,y] A[l,x(y),0]
      if (nr != 1 || (site.kind != B200_PT_LAST && site.kind != B200_PT_CLOSED)) {
        b200::set_error("pt_zip_up_left: bad newest site");
        return B200_EINVAL;
      }
      nk = 1;
      B200_TRY(dev_alloc(c, &theta, (size_t)ny * nl * ne));
      const int64_t xs = (site.kind == B200_PT_LAST) ? 0 : nr;
      B200_TRY(gemm(c, 1, nl, 1, ny, ne, op(c->one, 0, 0), op(a.p, 0, (int64_t)nx * nr, xs),
                    theta, 0, ne, (int64_t)nl * ne, 1, mat, 1, ny));
    } else {
      if (site.kind != B200_PT_MID || ny != nx || cr != nr || ce != ne) {
        b200::set_error("pt_zip_up_left: bad middle site");
        return B200_EINVAL;
      }
      nk = ck;
      B200_TRY(dev_alloc(c, &theta, (size_t)nk * ny * nl * ne));
      B200_TRY(gemm(c, nk, nl, nr, ny, ne, op(carry, (int64_t)nr * ne, ne, 0, 1),
                    op(a.p, 1, (int64_t)nx * nr, nr), theta, (int64_t)ny * nl * ne, ne,
                    (int64_t)nl * ne, 1, mat, 1, ny));
    }
    const int m = nk * ny, n = nl * ne;
    int nj = 0;
    B200_TRY(split(c, theta, m, n, n, 1, eps, &nj));
    cplx *new_site = nullptr, *new_carry = nullptr;
    B200_TRY(dev_alloc(c, &new_site, (size_t)nj * ny * nk));
    B200_TRY(dev_alloc(c, &new_carry, (size_t)nj * nl * ne));
    B200_TRY(emit(c, m, n, nj, new_site, ny, 1, nk, (int64_t)ny * nk, new_carry));
    B200_TRY(dev_free(c, theta));
    B200_TRY(dev_free(c, a.p));
    if (carry) B200_TRY(dev_free(c, carry));
    c->sites[ia] = Site{new_site, nj, ny, nk};
    carry = new_carry; ck = nj; cr = nl; ce = ne;
  }
  if (carry) B200_TRY(dev_free(c, carry));
  return B200_OK;
}

}  // extern "C"

// mps.zip_up(mpo, left_index=0, right_index=-1, direction="right") with the implicit TEMPO
// influence MPO (oqupy/backends/tempo_backend.py:539-547 -> node_array.py:482-552):
//   Theta[(k,s),(r,e)] = M[s,e] sum_l carry[k,l,e] A[l,s,r];  new site = U as (k, s, j),
//   carry' = S Vh as (j, r, e); the last (dk = 0) site is dense and takes the carry.
static int tempo_zip_up_right(Chain* c, const b200_tempo_site* mpo, int nb, double eps) {
  if (nb != (int)c->sites.size()) {
    b200::set_error("tempo_zip_up_right: MPO and MPS lengths differ");
    return B200_EINVAL;
  }
  cplx* carry = nullptr;
  int ck = 0, cl = 0, ce = 0;            // carry (k, l, e)
  for (int ib = 0; ib < nb; ++ib) {
    Site a = c->sites[ib];
    const int nl = a.dl, nn = a.da, nr = a.dr;
    const b200_tempo_site& site = mpo[ib];
    const cplx* mat = (const cplx*)site.mat;
    if (ib == nb - 1) {                  // dense dk = 0 site, no SVD
      if (site.kind != B200_TEMPO_DENSE || nr != 1) {
        b200::set_error("tempo_zip_up_right: last site must be dense");
        return B200_EINVAL;
      }
      const int nw = site.nw, nse = site.cols, ns = site.ns;
      int nk = 1;
      b200_operand cv = op(c->one, 0, 0);
      if (carry) {
        if (ce != nw || cl != nl) { b200::set_error("tempo_zip_up_right: carry mismatch"); return B200_EINVAL; }
        nk = ck;
        cv = op(carry, (int64_t)nl * nw, nw, 1);
      } else if (nl != 1 || nw != 1) {
        b200::set_error("tempo_zip_up_right: bad single-site step");
        return B200_EINVAL;
      }
      cplx *tmp = nullptr, *out = nullptr;         // T[k,w,n] = sum_l C[k,l,w] A[l,n]
      B200_TRY(dev_alloc(c, &tmp, (size_t)nk * nw * nn));
      B200_TRY(gemm(c, nk, nn, nl, nw, 1, cv, op(a.p, (int64_t)nn * nr, nr), tmp,
                    (int64_t)nw * nn, 1, nn, 0));
      B200_TRY(dev_alloc(c, &out, (size_t)nk * nse));
      B200_TRY(gemm(c, nk, nse, nw * nn, 1, 1, op(tmp, (int64_t)nw * nn, 1), op(mat, nse, 1), out,
                    nse, 1, 0, 0));
      B200_TRY(dev_free(c, tmp));
      B200_TRY(dev_free(c, a.p));
      if (carry) B200_TRY(dev_free(c, carry));
      c->sites[ib] = Site{out, nk, ns, nse / ns};
      return B200_OK;
    }
    const int ns = site.rows, ne = site.cols;
    if (ns != nn) { b200::set_error("tempo_zip_up_right: array leg mismatch"); return B200_EINVAL; }
    int nk = 0;
    cplx* theta = nullptr;
    if (!carry) {                        // Theta[0,s,r,e] = A[0,s,r] M[s,e]
      if (site.kind != B200_TEMPO_START || nl != 1) {
        b200::set_error("tempo_zip_up_right: bad first site");
        return B200_EINVAL;
      }
      nk = 1;
      B200_TRY(dev_alloc(c, &theta, (size_t)ns * nr * ne));
      B200_TRY(gemm(c, 1, nr, 1, ns, ne, op(c->one, 0, 0), op(a.p, 0, 1, nr), theta, 0, ne,
                    (int64_t)nr * ne, 1, mat, ne, 1));
    } else {
      if (site.kind != B200_TEMPO_MID || cl != nl || ce != ne) {
        b200::set_error("tempo_zip_up_right: bad middle site");
        return B200_EINVAL;
      }
      nk = ck;
      B200_TRY(dev_alloc(c, &theta, (size_t)nk * ns * nr * ne));
      B200_TRY(gemm(c, nk, nr, nl, ns, ne, op(carry, (int64_t)nl * ne, ne, 0, 1),
                    op(a.p, (int64_t)nn * nr, 1, nr), theta, (int64_t)ns * nr * ne, ne,
                    (int64_t)nr * ne, 1, mat, ne, 1));
    }
    const int m = nk * ns, n = nr * ne;
    int nj = 0;
    B200_TRY(split(c, theta, m, n, n, 1, eps, &nj));
    cplx *new_site = nullptr, *new_carry = nullptr;
    B200_TRY(dev_alloc(c, &new_site, (size_t)m * nj));
    B200_TRY(dev_alloc(c, &new_carry, (size_t)nj * n));
    B200_TRY(emit(c, m, n, nj, new_site, 1, nj, 0, 1, new_carry));
    B200_TRY(dev_free(c, theta));
    B200_TRY(dev_free(c, a.p));
    if (carry) B200_TRY(dev_free(c, carry));
    c->sites[ib] = Site{new_site, nk, ns, nj};
    carry = new_carry; ck = nj; cl = nr; ce = ne;
  }
  if (carry) B200_TRY(dev_free(c, carry));
  return B200_OK;
}

extern "C" int b200_chain_tempo_step(void* h, const b200_tempo_site* mpo, int nb, const void* p1_,
                                     const void* p2site_, const void* sum_north_, int d2,
                                     double eps, void* state_out_) {
  Chain* c = (Chain*)h;
  const cplx *p1 = (const cplx*)p1_, *p2site = (const cplx*)p2site_,
             *sn = (const cplx*)sum_north_;
  cplx* state = (cplx*)state_out_;
  if (!c || !mpo || !p1 || !p2site || !sn || !state || nb < 1 || d2 < 1 || c->sites.empty()) {
    b200::set_error("b200_chain_tempo_step: invalid argument");
    return B200_EINVAL;
  }
  // first half propagator on the newest site (tempo_backend.py:521-529)
  {
    Site last = c->sites.back();
    if (last.da != d2 || last.dr != 1) { b200::set_error("tempo_step: bad newest site"); return B200_EINVAL; }
    cplx* nl_ = nullptr;
    B200_TRY(dev_alloc(c, &nl_, (size_t)last.dl * d2));
    B200_TRY(gemm(c, last.dl, d2, d2, 1, 1, op(last.p, d2, 1), op(p1, 1, d2), nl_, d2, 1, 0, 0));
    B200_TRY(dev_free(c, last.p));
    c->sites.back().p = nl_;
  }
  // sum out the oldest leg beyond the memory cut-off (:531-537)
  if ((int)c->sites.size() != nb) {
    if ((int)c->sites.size() != nb + 1 || c->sites[0].dl != 1) {
      b200::set_error("tempo_step: MPS / MPO length mismatch");
      return B200_EINVAL;
    }
    Site first = c->sites[0], second = c->sites[1];
    cplx *vec = nullptr, *merged = nullptr;
    B200_TRY(dev_alloc(c, &vec, (size_t)first.dr));
    B200_TRY(gemm(c, 1, first.dr, first.da, 1, 1, op(sn, 0, 1), op(first.p, first.dr, 1), vec, 0, 1, 0, 0));
    const int sn2 = second.da * second.dr;
    B200_TRY(dev_alloc(c, &merged, (size_t)sn2));
    B200_TRY(gemm(c, 1, sn2, first.dr, 1, 1, op(vec, 0, 1), op(second.p, sn2, 1), merged, 0, 1, 0, 0));
    B200_TRY(dev_free(c, vec));
    B200_TRY(dev_free(c, first.p));
    B200_TRY(dev_free(c, second.p));
    c->sites[1] = Site{merged, 1, second.da, second.dr};
    c->sites.erase(c->sites.begin());
  }
  B200_TRY(tempo_zip_up_right(c, mpo, nb, eps));                       // :539-547
  B200_TRY(b200_chain_svd_sweep(h, (int)c->sites.size() - 1, 0, eps));    // :549-553
  B200_TRY(b200_chain_push(h, p2site, d2, d2, 1));                     // :555-558
  // read-out (:560-573): vec <- sum_a sn[a] (vec . A[:, a, :])
  const cplx* vec = c->one;
  cplx* owned = nullptr;
  const int n = (int)c->sites.size();
  for (int i = 0; i < n - 1; ++i) {
    const Site& a = c->sites[i];
    cplx *tmp = nullptr, *nxt = nullptr;
    B200_TRY(dev_alloc(c, &tmp, (size_t)a.da * a.dr));
    B200_TRY(gemm(c, 1, a.da * a.dr, a.dl, 1, 1, op(vec, 0, 1), op(a.p, (int64_t)a.da * a.dr, 1), tmp,
                  0, 1, 0, 0));
    B200_TRY(dev_alloc(c, &nxt, (size_t)a.dr));
    B200_TRY(gemm(c, 1, a.dr, a.da, 1, 1, op(sn, 0, 1), op(tmp, a.dr, 1), nxt, 0, 1, 0, 0));
    B200_TRY(dev_free(c, tmp));
    if (owned) B200_TRY(dev_free(c, owned));
    owned = nxt; vec = nxt;
  }
  B200_TRY(gemm(c, 1, d2, d2, 1, 1, op(vec, 0, 1), op(c->sites.back().p, d2, 1), state, 0, 1, 0, 0));
  if (owned) B200_TRY(dev_free(c, owned));
  return B200_OK;
}
