/* oqupy_b200 -- C-ABI of the B200-native TEMPO / PT-TEMPO engine.
 *
 * Plain pointers and sizes only (no torch types).  All tensor pointers are DEVICE
 * pointers to complex128 (interleaved re,im doubles) unless stated otherwise;
 * `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success
 * or a negative B200_E* code; b200_last_error() gives the message.
 *
 * The reference (OQuPy 0.5.0) is pure Python and has no FFI: its hot path reaches
 * the arithmetic through the third-party `tensornetwork` calls listed below.  Each
 * entry point names the reference call site(s) it replaces.
 */
#ifndef OQUPY_B200_H
#define OQUPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_EINVAL (-1)
#define B200_ECUDA (-2)
#define B200_ENOCONV (-3)
#define B200_ESIZE (-4)

const char* b200_last_error(void);
int b200_abi_version(void);
/* number of kernel launches issued through this library since load */
uint64_t b200_launch_count(void);

/* Live per-kernel timing for the roofline report (bench.py): when enabled, every
 * launch of the dominant kernel (jacobi_kernel, the truncated SVD) is bracketed by
 * CUDA events on its own stream.  b200_profile_read synchronises the device, adds
 * up the finished launches and returns: total kernel milliseconds, total
 * ALGORITHMIC flops (4*(14 m n^2 + 8 n^3), m >= n, per truncated SVD),
 * number of launches and total Jacobi sweeps; then resets the counters. */
int b200_profile_enable(int on);
int b200_profile_read(double* kernel_ms, double* algorithmic_flops,
                      uint64_t* launches, uint64_t* sweeps);
/* The same, split by kernel family of the truncated SVD: ms4 / launches4 index
 * 0 = Jacobi stage (jacobi_kernel), 1 = rank-revealing QR (qrcp_kernel), 2 = back-
 * transformation + emit (apply_q_kernel, emit_l_kernel, emit_kernel), 3 = other. */
int b200_profile_read_kinds(double* ms4, uint64_t* launches4, double* algorithmic_flops,
                            uint64_t* sweeps);

/* ---------------------------------------------------------------------------
 * Strided, doubly-batched complex GEMM with per-batch scale:
 *   C[b1,b2][i,j] = scale[b1,b2] * sum_t opA(A[b1,b2])[i,t] * opB(B[b1,b2])[t,j]
 * Every operand is addressed with ELEMENT strides (row, col, batch1, batch2), so
 * the delta-structured influence MPO sites never have to be materialised.
 * scale may be NULL (== 1).  conj_a / conj_b conjugate the operand elements.
 * If accumulate != 0 the result is added to C.
 *
 * Replaces: tn.contractors.greedy([A_site, B_site, carry])  oqupy/backends/node_array.py:395,519
 *           svh @ nodes[i+-1]                               oqupy/backends/node_array.py:272-273,295-296
 *           node @ node in compute_dynamics                 oqupy/system_dynamics.py:638,649,697
 *           SimpleProcessTensor.compute_caps                oqupy/process_tensor.py:398,403
 */
typedef struct {
  const void* ptr;
  int64_t row, col, b1, b2; /* element strides */
  int conj;
} b200_operand;

int b200_zgemm_strided(void* stream, int m, int n, int k, int nb1, int nb2,
                       const b200_operand* a, const b200_operand* b,
                       void* c, int64_t c_row, int64_t c_col, int64_t c_b1,
                       int64_t c_b2, const void* scale, int64_t s_b1,
                       int64_t s_b2, int accumulate);

/* ---------------------------------------------------------------------------
 * eps-truncated SVD: rank-revealing QR stopped at the deflation level (roughly square
 * operands, eps > 0), then one-sided block Jacobi (row-sliced 2-D grid, cooperative launches)
 * on the surviving columns, back-transformation by the stored reflectors.
 *
 * `info_host` is PINNED host memory of 8 x int32: {keep, sweeps, status, rotations,
 * pivots of the QR stage (-1 = not used), QR grid, spare, spare}.  The library itself spins
 * on word 4 inside b200_svd_factor* when the QR stage runs (its pivot count fixes the shape
 * of the Jacobi stage); the caller waits for word 3 (or the stream) before reading keep.
 *
 * b200_svd_factor: theta is an m x n matrix addressed theta[i*rs + j*cs].
 *   Computes all singular triplets on the device, sorts them, applies the
 *   reference's tail-norm rule
 *       keep = #{ j : sqrt(sum_{i>=j} s_i^2) > eps * s_0 }      (eps < 0: keep all)
 *   and writes {keep, sweeps, status, rotations} to `info_host` (see above).  The caller synchronises the stream, reads
 *   keep, allocates exact-size outputs and calls b200_svd_emit.
 *   `work` must hold b200_svd_workspace_bytes(m, n) bytes of device memory.
 *
 * b200_svd_emit: writes U (m x keep) with the row index i split as
 *   (i / u_na, i % u_na) -> address (i/u_na)*u_so + (i%u_na)*u_sa + j*u_sj, and
 *   S*Vh (keep x n) row-major into svh.  Either output may be NULL.  `theta`, rs, cs
 *   are the same matrix as given to b200_svd_factor (reserved; the right factor is
 *   taken from the accumulated rotations, so theta may already have been reused).
 *
 * b200_svd_values: copies the min(m,n) sorted singular values (doubles) to s_out.
 *
 * b200_svd_factor2: as b200_svd_factor with two-level row and column indices,
 *   theta[i][j] at (i / rin)*rso + (i % rin)*rsi + (j / cin)*cso + (j % cin)*csi, so that
 *   any leg grouping of a rank-4 tensor is factorised in place -- the PT-TEBD splits
 *   (left_edges / right_edges of oqupy/backends/pt_tebd_backend.py:487-531).  cos_tol > 0
 *   sets the orthogonality target |cos| of the Jacobi iteration (default 1e-11, never below
 *   the rounding level 2 sqrt(max(m,n)) eps_mach): PT-TEBD multiplies the factors by
 *   inverse singular values (pt_tebd_backend.py:533-559), which amplifies a residual
 *   non-orthogonality by 1/lambda, and asks for 1e-12 (which also selects the relative-accuracy mode).
 * b200_svd_emit_parts: as b200_svd_emit; `vh` receives S*Vh (vh_unscaled = 0) or Vh
 *   (vh_unscaled = 1), `lam` / `inv_lam` (complex128[keep], may be NULL) the kept singular
 *   values and their inverses (the lambda matrices of pt_tebd_backend.py:526, 573-578).
 *
 * Replaces: tn.split_node_full_svd(node, left_edges, right_edges,
 *             max_truncation_err=eps, relative=True)   oqupy/backends/node_array.py:262,285,541
 *           and `s @ vh` at node_array.py:272,295,552; truncation rule mirrored at
 *           oqupy/mps_mpo.py:452-457.
 */
size_t b200_svd_workspace_bytes(int m, int n);
int b200_svd_factor(void* stream, const void* theta, int m, int n, int64_t rs,
                    int64_t cs, double eps, void* work, int32_t* info_host);
int b200_svd_factor2(void* stream, const void* theta, int m, int n, int rin, int64_t rso,
                     int64_t rsi, int cin, int64_t cso, int64_t csi, double eps, double cos_tol,
                     void* work, int32_t* info_host);
int b200_svd_emit_parts(void* stream, const void* work, int m, int n, int keep, void* u,
                        int u_na, int64_t u_so, int64_t u_sa, int64_t u_sj, void* vh,
                        int vh_unscaled, void* lam, void* inv_lam);
int b200_svd_emit(void* stream, const void* work, const void* theta, int m, int n,
                  int64_t rs, int64_t cs, int keep, void* u, int u_na, int64_t u_so,
                  int64_t u_sa, int64_t u_sj, void* svh);
int b200_svd_values(void* stream, const void* work, int m, int n, double* s_out);
/* diagnostics of the factorisation that last ran in `work`: out4 = {QR path used (0/1),
 * pivots above the stop level, CTAs of the QR grid, columns resident in shared memory} */
int b200_svd_plan(const void* work, int m, int n, int32_t* out4);
/* test access to the QR-path workspace (byte offsets of its arrays inside `work`):
 * out16 = {eligible, p, q, transposed, QR grid, resident columns, perm, tau, work array a,
 * Jacobi stage, header, tail partials, shared-memory bytes, and for k > 0 pivots the Jacobi
 * stage's y, singular values, column order} */
int b200_svd_qr_layout(int m, int n, int k, int64_t* out16);
/* runtime switches (tests, A/B measurements): "qr" (0/1: rank-revealing QR front end),
 * "qr_minq" (smallest min(m,n) that takes it), "qr_cols" (columns per CTA of the QR grid);
 * "max_slices" (CALLING THREAD only; < 0 clears it): cap of the row slices per pair slot, i.e.
 * of the cooperative grid, for runs that share the GPU with a persistent kernel */
int b200_svd_config(const char* key, double value);
/* diagnostics: SM-clock cycles CTA 0 (leader of pair slot 0) spent per phase of the
 * Jacobi kernel: {0 wait for input blocks, 1 load + partial Gram, 2 publish, 3 wait
 * for all partials, 4 reduce + convergence test, 5 inner 32x32 sweep, 6 sort + publish
 * J, 7 apply + hand-over, 8 sweep vote, 9 norms/rank, 10..14 spare, 15 #stages} */
int b200_svd_phase_cycles(void* stream, const void* work, long long* out16);
/* the same for the rank-revealing QR stage: {0 pick the CTA's best column, 1 wait for all offers,
 * 2 select the panel, 3 fetch it, 4 factorise it, 5 file the pivot columns, 6 apply the
 * reflectors to the CTA's own columns, 7 publish the offer} */
int b200_svd_qr_phase_cycles(void* stream, const void* work, long long* out8);

/* ---------------------------------------------------------------------------
 * Native matrix-product chain: the device-resident counterpart of NodeArray for the
 * PT-TEMPO path (oqupy/backends/node_array.py).  A chain owns its sites
 * (chi_l, a, chi_r) in device memory and runs a whole zip-up / svd-sweep in ONE call
 * (contraction GEMM -> truncated SVD -> exact-size outputs, site after site; the only
 * host round trip per site is the 16-byte `keep` read-back).
 *
 * b200_chain_pt_zip_up_left replaces mps.zip_up(mpo, axes=[(0,0)], right_index=-1,
 *   direction="left", max_truncation_err=eps, relative=True)
 *   (oqupy/backends/pt_tempo_backend.py:267-274 -> node_array.py:482-552) for the
 *   implicit (delta-structured) PT-TEMPO influence MPO (pt_tempo_backend.py:142-152):
 *     kind FIRST  B[x,y,r]   = d_xy d_xr vec[x]          mat = vec (rows entries)
 *     kind MID    B[l,x,y,r] = d_lr d_xy mat[l,x]         mat (rows x cols)
 *     kind LAST   B[l,0,y,0] = mat[l,y]                   newest site
 *     kind CLOSED B[l,x,y,0] = d_xy mat[l,x]              end phase (mat times closing vector)
 * b200_chain_svd_sweep replaces mps.svd_sweep(from_index, to_index, eps, relative=True)
 *   (pt_tempo_backend.py:171-181, 276-280 -> node_array.py:226-299); negative indices
 *   count from the end.
 */
#define B200_PT_FIRST 0
#define B200_PT_MID 1
#define B200_PT_LAST 2
#define B200_PT_CLOSED 3
typedef struct {
  int kind;
  int rows, cols;
  const void* mat; /* device, complex128, row-major */
  /* kind FIRST with degeneracy maps (unique=True, pt_tempo_backend.py:125-137), else NULL:
   * HOST arrays of `cols` entries; B[w,y,n] = [w = west_map[y]] [n = north_map[y]] vec[n],
   * mat = vec (rows = number of distinct north values) */
  const int32_t* north_map;
  const int32_t* west_map;
} b200_pt_site;

void* b200_chain_create(void* stream);
int b200_chain_destroy(void* chain);
int b200_chain_len(void* chain);
/* append a site (copied device-to-device) */
int b200_chain_push(void* chain, const void* dev_src, int chi_l, int a, int chi_r);
int b200_chain_shape(void* chain, int i, int32_t* out3);
/* copy site i (contiguous) into caller-owned device memory */
int b200_chain_read(void* chain, int i, void* dev_dst);
int b200_chain_svd_sweep(void* chain, int from_index, int to_index, double eps);
int b200_chain_pt_zip_up_left(void* chain, const b200_pt_site* mpo, int n_mpo, double eps);
/* One whole TEMPO time step on the chain (oqupy/backends/tempo_backend.py:439-575):
 * first half propagator on the newest site (:521-529), sum out the oldest leg beyond the
 * memory cut-off (:531-537), mps.zip_up(mpo, direction="right") (:539-547 -> node_array.py:
 * 482-552) with the implicit influence MPO (tempo_backend.py:419,426)
 *     kind START  first aligned site, west leg summed:  mat[s,e] = infl[s,e] * sum_west[e]
 *     kind MID    B[w,n,s,e] = d_we d_ns mat[s,e]
 *     kind DENSE  dk = 0 site incl. the unitary transform, mat[(w,n),(s,e)] (rows x cols),
 *                 nw = size of its west leg (1 when it is the first aligned site), ns = size of s
 * svd_sweep right-to-left (:549-553), append the second half propagator `p2site`
 * ((d2,d2,1) = prop_2^T, :555-558) and the read-out of the state (:560-573) into
 * `state_out` (device, d2).  p1 is prop_1 (d2 x d2, row-major); sum_north complex128. */
#define B200_TEMPO_START 0
#define B200_TEMPO_MID 1
#define B200_TEMPO_DENSE 2
typedef struct {
  int kind;
  int rows, cols;
  int nw, ns;
  const void* mat; /* device, complex128, row-major */
} b200_tempo_site;
int b200_chain_tempo_step(void* chain, const b200_tempo_site* mpo, int n_mpo, const void* p1,
                          const void* p2site, const void* sum_north, int d2, double eps,
                          void* state_out);
/* counters since the last reset: truncated SVDs, Jacobi sweeps, D2H bytes (keep read-backs) */
int b200_chain_stats(void* chain, uint64_t* nsvd, uint64_t* sweeps, uint64_t* d2h_bytes,
                     int reset);
/* per-SVD log (m, n, keep, sweeps as 4 x int32): returns the entries copied to `out`
 * (and clears them); `enable` switches logging on/off */
int b200_chain_log(void* chain, int enable, int32_t* out, int cap_entries);

/* ---------------------------------------------------------------------------
 * compute_dynamics step for ONE environment and `nvec` ensemble members that
 * share the process tensor (oqupy/system_dynamics.py:631-700):
 *   v'[e, r, j] = sum_x P2[e][j,x] * sum_l T[l, r, x] * (sum_i P1[e][x,i] v[e, l, i])
 * T is the rank-3 PT-MPO site (chi_l, chi_r, d2) (oqupy/process_tensor.py:301-305),
 * v (nvec, chi_l, d2), v_out (nvec, chi_r, d2), p1/p2 (nvec, d2, d2) row-major.
 * If cap (chi_l) and rho_out (nvec, d2) are non-NULL, the read-out
 *   rho[e, i] = sum_l cap[l] * v[e, l, i]          (system_dynamics.py:643-651)
 * of the INPUT state is fused into the same pass.  `work` holds
 * b200_dyn_workspace_bytes(...) bytes of device scratch.
 */
size_t b200_dyn_workspace_bytes(int nvec, int chi_l, int chi_r, int d2);
int b200_dyn_step(void* stream, int nvec, int chi_l, int chi_r, int d2,
                  const void* t, const void* p1, const void* p2, const void* v,
                  void* v_out, const void* cap, void* rho_out, void* work);

/* The whole loop of compute_dynamics (oqupy/system_dynamics.py:131-170) for ONE
 * environment in one call: for k = 0..nsteps-1 the fused step above (read-out of step k
 * included), then the final read-out.  t[k] / caps[k] are HOST arrays of device pointers
 * to the PT-MPO sites (chi[k], chi[k+1], d2) and cap vectors (chi[k]); p1 / p2 hold the
 * propagators of step k at element offset k*prop_step_stride (0: time independent), each
 * (nvec, d2, d2); v0 (nvec, chi[0], d2); rho_out (nsteps+1, nvec, d2). */
size_t b200_dyn_run_workspace_bytes(int nsteps, int nvec, const int32_t* chi, int d2);
int b200_dyn_run(void* stream, int nsteps, int nvec, int d2, const int32_t* chi,
                 const void* const* t, const void* p1, const void* p2,
                 int64_t prop_step_stride, const void* const* caps, const void* v0,
                 void* rho_out, void* work);

/* cap_k[l] = sum_{r,x} T[l,r,x] * cap_next[r] * tr2[x]   (oqupy/process_tensor.py:380-406) */
int b200_caps_step(void* stream, int chi_l, int chi_r, int d2, const void* t,
                   const void* cap_next, const void* tr2, void* cap_out);

/* ---------------------------------------------------------------------------
 * Lock-step TEMPO ensemble: E independent TEMPO runs (same d2, dkmax, epsrel; different
 * influence matrices / propagators / initial states) advance one time step per call with
 * ONE kernel launch -- one CTA per member runs BaseTempoBackend.compute_system_step
 * (oqupy/backends/tempo_backend.py:439-575) for its member entirely on the device (chain in
 * capacity-padded device slots, shapes on the device, every truncated SVD in shared memory:
 * stopped column-pivoted QR + one-sided Jacobi on R + the reference's tail-norm rule).  The
 * loop it replaces is the ensemble of oqupy.Tempo(...).compute() calls of BASELINE configs[4]
 * (oqupy/tempo.py:479-484 per member) and MeanFieldTempoBackend's per-system loop
 * (tempo_backend.py:747-773).  chi_cap bounds the bond dimension (chi_cap * d2 <= 128; an operand p x q must fit shared memory, (p|1)*q <= 14272).
 *
 * b200_tempo_batch_set (device pointers, complex128, member-major):
 *   mid, start (E, dkmax+1, d2, d2): infl[dk][s,e] and infl[dk][s,e]*sum_west[e]
 *   (tempo_backend.py:419,426,519); dense0 (E, d2*d2, d2*d2): the dk = 0 site incl. the
 *   unitary transform as [(w,n),(s,e)] (:419-424); dense0w (E, d2, d2*d2): its west leg summed;
 *   sum_north (d2); state0 (E, d2).
 * b200_tempo_batch_step: p1 (E, d2, d2) = prop_1, p2t (E, d2, d2) = prop_2^T of this step;
 *   states_out (E, d2) device or NULL.
 * b200_tempo_batch_info (host arrays, synchronises): per member status (0 ok, 2 capacity
 *   exceeded, 3 no convergence, >= 4 internal), truncated SVDs, Jacobi sweeps, largest bond
 *   dimension; bonds[e*(dkmax+2) + i] = bond dimensions of the chain (-1 padded). */
void* b200_tempo_batch_create(void* stream, int n_members, int d2, int dkmax, int chi_cap,
                              double epsrel);
int b200_tempo_batch_destroy(void* batch);
int b200_tempo_batch_set(void* batch, const void* mid, const void* start, const void* dense0,
                         const void* dense0w, const void* sum_north, const void* state0);
int b200_tempo_batch_step(void* batch, const void* p1, const void* p2t, void* states_out);
int b200_tempo_batch_info(void* batch, int32_t* status, int32_t* svds, int32_t* sweeps,
                          int32_t* max_chi, int32_t* bonds);
/* CTA i of the following steps takes member order[i] (host array, a permutation of
 * 0..n_members-1; NULL: identity): longest-running members first (a scheduling hint, the
 * results do not depend on it). */
int b200_tempo_batch_set_order(void* batch, const int32_t* order);
/* The following steps leave n SMs to other streams (0: none): members that outgrow shared
 * memory are re-run on the general backend NEXT TO the batch and need free SMs for their
 * cooperative launches. */
int b200_tempo_batch_reserve_sms(void* batch, int n);
size_t b200_tempo_batch_bytes(void* batch);

#ifdef __cplusplus
}
#endif
#endif /* OQUPY_B200_H */
