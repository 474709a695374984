"""Drop-in ``PtTebdBackend`` (oqupy/backends/pt_tebd_backend.py:46-565, SURVEY.md 8a row
A7): the augmented MPS

    lam[0] - Gam[0] - lam[1] - ... - Gam[n-1] - lam[n],     Gam (chi_l, d2, chi_pt, chi_r)

lives on the B200; every contraction is a strided batched GEMM and every split an
eps-truncated SVD of the C-ABI library (oqupy_b200/csrc).  The leg groupings of the three
splits of a nearest-neighbour gate (pt_tebd_backend.py:487-531) are factorised IN PLACE
through the two-level row / column strides of ``b200_svd_factor2`` -- no tensor is ever
permuted in memory.  lambda matrices are kept as their diagonals (complex128 vectors, next
to their inverses).

Gauge: singular vectors are fixed up to phases, so ``get_gamma`` agrees with the reference
up to a bond gauge; density matrices, norms, lambdas and bond dimensions are gauge
invariant and are what the parity tests compare.
"""
import os

import numpy as np

from ._lib import View, default_ops
from .process_tensor import as_device_process_tensor

CDTYPE = np.complex128
# Orthogonality target of the truncated SVDs.  The gate update multiplies the factors by
# inverse singular values (pt_tebd_backend.py:533-559): a residual |cos| between columns of U
# is amplified by 1/lambda <= 1/eps.  What matters is the relative-accuracy mode any
# cos_tol > 0 selects (tiny absolute floor, Rutishauser updates, no loose columns); the
# tolerance itself was measured: 1e-12 and 1e-15 give the same 5e-10 agreement with the
# LAPACK oracle on the config-4 shape, 1e-12 is 19 % faster.
COS_TOL = float(os.environ.get("B200_TEBD_COS_TOL", "1e-12"))


def _isqrt(x):
    n = int(round(np.sqrt(x)))
    if n ** 2 != x:
        raise ValueError(f"{x} is not a perfect square. Can't take integer square root!")
    return n


class PtTebdBackend:
    """PT-TEBD backend on the device: same constructor and methods as the reference
    (pt_tebd_backend.py:46-445).  The gates of a layer are independent
    (pt_tebd_backend.py:134-155 maps them over a process / thread pool): with
    ``config['parallel']`` set ('multithread' / 'multiprocess' as in the reference, or a number
    of workers) -- or ``B200_TEBD_STREAMS`` > 1 -- they run on that many host threads, each with
    its own ops object and CUDA stream, so that their dependent chains of small launches (three
    truncated SVDs and six contractions per gate) overlap on the SMs; the results are
    bit-identical to the serial order."""

    def __init__(self, gammas, lambdas, epsrel, config=None, ops=None):
        assert len(gammas) == len(lambdas) + 1                         # :71
        self._n = len(gammas)
        self._epsrel = epsrel
        self._config = {} if config is None else config
        self._ops = ops = default_ops() if ops is None else ops
        self._gammas = []
        for gam in gammas:
            gam = np.asarray(gam, dtype=CDTYPE)
            assert gam.ndim == 4, "gamma tensors are (chi_l, d2, chi_pt, chi_r)"
            self._gammas.append(ops.from_host(gam))
        host = ([np.ones(gammas[0].shape[0], dtype=CDTYPE)]            # :97-102
                + [np.asarray(lam, dtype=CDTYPE).reshape(-1) for lam in lambdas]
                + [np.ones(gammas[-1].shape[3], dtype=CDTYPE)])
        for i in range(self._n):
            assert host[i].size == gammas[i].shape[0]
            assert host[i + 1].size == gammas[i].shape[3]
        self._lams = [ops.from_host(lam) for lam in host]
        self._inv_lams = [ops.from_host(1.0 / lam) for lam in host]
        self._gate_cache = {}
        par = self._config.get("parallel") if isinstance(self._config, dict) else None
        env = int(os.environ.get("B200_TEBD_STREAMS", "0"))
        if isinstance(par, (int, np.integer)) and not isinstance(par, bool):
            self._nworkers = int(par)
        elif par in ("multithread", "multiprocess"):
            self._nworkers = max(env, 4)
        elif par is None or par is False:
            self._nworkers = env
        else:
            raise NotImplementedError(f"Parallelisation method '{par}' is not implemented!")
        self._pool = None
        self.clear_traces()

    # -------------------------------------------------------------- small helpers
    def _scale_rows(self, x, lam, nrows, ncols, ops=None):
        """out[r, c] = lam[r] x[r, c]  (x contiguous, nrows x ncols)."""
        ops = self._ops if ops is None else ops
        out = ops.empty(*x.shape)
        ops.gemm(1, ncols, 1, View(ops.one), View(x, col=1, b1=ncols),
                 View(out, col=1, b1=ncols), nb1=nrows, scale=View(lam, b1=1))
        return out

    def _scale_cols(self, x, lam, nrows, ncols, ops=None):
        """out[r, c] = x[r, c] lam[c]."""
        ops = self._ops if ops is None else ops
        out = ops.empty(*x.shape)
        ops.gemm(nrows, 1, 1, View(x, row=ncols, b1=1), View(ops.one),
                 View(out, row=ncols, b1=1), nb1=ncols, scale=View(lam, b1=1))
        return out

    # -------------------------------------------------------------- properties
    @property
    def n(self):
        """Number of chain sites. """
        return self._n

    def get_gamma(self, site):
        """Gamma tensor of a site (host copy; fixed up to the bond gauge)."""
        return self._ops.to_host(self._gammas[site])

    def get_lambda(self, site):
        """Lambda matrix to the right of a site (host, dense diagonal like the reference)."""
        return np.diag(self._ops.to_host(self._lams[site + 1]))

    def get_bond_dimensions(self):
        return np.array([int(g.shape[3]) for g in self._gammas[:-1]])

    # -------------------------------------------------------------- gates
    def apply_nn_gate_layer(self, gate_layer):                         # :134-155
        gates = list(gate_layer.gates)
        nw = min(self._nworkers, len(gates))
        if nw <= 1 or getattr(self._ops, "name", "") != "cuda":
            for gate in gates:
                self.apply_nn_gate(gate)
            return
        self._apply_nn_gates_concurrently(gates, nw)

    def _apply_nn_gates_concurrently(self, gates, nw):
        """The gates of a layer touch disjoint site pairs (and only READ the outer lambdas):
        worker w takes gates w, w + nw, ... on its own stream; all streams start after the
        work already queued on the caller's stream and the caller's stream waits for all of
        them; tensors that cross streams are registered with the caching allocator."""
        import threading  # pylint: disable=import-outside-toplevel
        import torch  # pylint: disable=import-outside-toplevel
        from ._lib import CudaOps  # pylint: disable=import-outside-toplevel
        dev = self._ops.device
        if self._pool is None:
            self._pool = []
        while len(self._pool) < nw:
            self._pool.append((CudaOps(dev.index), torch.cuda.Stream(device=dev)))
        main = torch.cuda.current_stream(dev)
        streams = [main] + [st for _, st in self._pool[:nw]]
        for gate in gates:                    # gate matrices are uploaded on the main stream
            self._gate_matrix(gate)
        start = main.record_event()
        results, errors, done = [None] * len(gates), [], [None] * nw

        def work(w):
            try:
                wops, stream = self._pool[w]
                with torch.cuda.stream(stream):
                    stream.wait_event(start)
                    for gi in range(w, len(gates), nw):
                        results[gi] = self._nn_gate_compute(gates[gi], wops, streams)
                    done[w] = stream.record_event()
            except Exception as exc:  # pylint: disable=broad-except
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(w,)) for w in range(nw)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        for ev in done:
            main.wait_event(ev)
        for res in results:
            self._nn_gate_commit(*res)

    def apply_site_gate_layer(self, gate_layer):                       # :233-236
        for gate in gate_layer.gates:
            self.apply_site_gate(gate)

    def apply_site_gate(self, gate):
        """gam'[l,a,b,r] = sum_p M[a,p] gam[l,p,b,r]  (:238-251)."""
        ops = self._ops
        site = gate.sites[0]
        mat = np.asarray(gate.tensors[0], dtype=CDTYPE)
        gam = self._gammas[site]
        nl, d2, npt, nr = gam.shape
        na = mat.shape[0]
        assert mat.shape[1] == d2
        mdev = ops.from_host(mat)
        out = ops.empty(nl, na, npt, nr)
        ops.gemm(na, npt * nr, d2, View(mdev, row=d2, col=1),
                 View(gam, row=npt * nr, col=1, b1=d2 * npt * nr),
                 View(out, row=npt * nr, col=1, b1=na * npt * nr), nb1=nl)
        self._gammas[site] = out

    def _gate_matrix(self, gate):
        """G[(a,b),(p,q)] = sum_g gate_l[a,p,g] gate_r[g,b,q] on the device (edge wiring
        pt_tebd_backend.py:474-481); cached per gate object."""
        key = id(gate)
        hit = self._gate_cache.get(key)
        if hit is not None and hit[0] is gate:
            return hit[1:]
        g_l = np.asarray(gate.tensors[0], dtype=CDTYPE)
        g_r = np.asarray(gate.tensors[1], dtype=CDTYPE)
        mat = np.einsum("apg,gbq->abpq", g_l, g_r)
        na, nb, np_, nq = mat.shape
        dev = self._ops.from_host(mat.reshape(na * nb, np_ * nq))
        self._gate_cache[key] = (gate, dev, na, nb, np_, nq)
        return dev, na, nb, np_, nq

    def apply_nn_gate(self, gate):
        """_apply_nn_gate (pt_tebd_backend.py:447-565), Figs. S2(c-h) of [Fux2023]."""
        self._nn_gate_commit(*self._nn_gate_compute(gate, self._ops, None))

    def _nn_gate_commit(self, sl, new_l, new_r, lam, inv_lam):         # :561-565
        self._gammas[sl], self._gammas[sl + 1] = new_l, new_r
        self._lams[sl + 1], self._inv_lams[sl + 1] = lam, inv_lam

    def _nn_gate_compute(self, gate, ops, streams):
        """The gate on (sl, sl + 1) with the kernels of `ops` on the current stream; returns
        what :meth:`_nn_gate_commit` stores.  ``streams``: every stream that may touch the
        chain (concurrent layers) -- inputs and outputs are registered with all of them."""
        eps = self._epsrel
        sl, sr = gate.sites[0], gate.sites[1]
        assert sr == sl + 1
        gmat, na, nb, d2l, d2r = self._gate_matrix(gate)
        gam_l, gam_r = self._gammas[sl], self._gammas[sr]
        lam_l, lam_m, lam_r = self._lams[sl], self._lams[sl + 1], self._lams[sr + 1]
        inv_l, inv_r = self._inv_lams[sl], self._inv_lams[sr + 1]

        def share(*tensors):
            if streams is not None:
                for t in tensors:
                    for st in streams:
                        t.record_stream(st)
        share(gmat, gam_l, gam_r, lam_l, lam_m, lam_r, inv_l, inv_r)
        nl, _, pl, nm = gam_l.shape
        _, _, pr, nr = gam_r.shape
        assert gam_l.shape[1] == d2l and gam_r.shape[1] == d2r and gam_r.shape[0] == nm
        # -- split the process-tensor leg off the left site (:487-496):
        #    (lam_l Gam_l)[(L,b),(p,M)] = U1 [(L,b),k1] . S1 Vh1 [k1,(p,M)]
        left = self._scale_rows(gam_l, lam_l, nl, d2l * pl * nm, ops)
        h1 = ops.svd_factor(left, nl * pl, d2l * nm, d2l * pl * nm, pl * nm, eps,
                            rin=pl, rsi=nm, cin=nm, csi=1, cos_tol=COS_TOL)
        k1 = h1.keep
        u1 = ops.empty(nl, pl, k1)
        svh1 = ops.empty(k1, d2l, nm)
        ops.svd_emit(h1, u=u1, u_na=1, u_so=k1, u_sa=0, u_sj=1, svh=svh1)
        left_mid = self._scale_cols(svh1, lam_m, k1 * d2l, nm, ops)            # times lam_m
        # -- and off the right site (:498-507): (Gam_r lam_r)[(M,p),(b,R)] = U2 S2 . Vh2.
        #    The transposed matrix is factorised, so that U' = Vh2^T and S Vh' = (U2 S2)^T
        right = self._scale_cols(gam_r, lam_r, nm * d2r * pr, nr, ops)
        h2 = ops.svd_factor(right, pr * nr, nm * d2r, 1, pr * nr, eps, cos_tol=COS_TOL)
        k2 = h2.keep
        right_temp = ops.empty(k2, pr, nr)
        us2t = ops.empty(k2, nm, d2r)
        ops.svd_emit(h2, u=right_temp, u_na=1, u_so=1, u_sa=0, u_sj=pr * nr, svh=us2t)
        # -- theta (:508-519): X[k1,p,q,k2] = sum_M left_mid[k1,p,M] us2t[k2,M,q], then the gate
        x = ops.empty(k1, d2l, d2r, k2)
        ops.gemm(k1, k2, nm, View(left_mid, row=d2l * nm, col=1, b1=nm),
                 View(us2t, row=d2r, col=nm * d2r, b2=1),
                 View(x, row=d2l * d2r * k2, col=1, b1=d2r * k2, b2=k2), nb1=d2l, nb2=d2r)
        theta = ops.empty(k1, na, nb, k2)
        ops.gemm(na * nb, k2, d2l * d2r, View(gmat, row=d2l * d2r, col=1),
                 View(x, row=k2, col=1, b1=d2l * d2r * k2),
                 View(theta, row=k2, col=1, b1=na * nb * k2), nb1=k1)
        # -- split theta (:521-531): rows (k1, a), columns (b, k2)
        h3 = ops.svd_factor(theta, k1 * na, nb * k2, nb * k2, 1, eps, cos_tol=COS_TOL)
        nj = h3.keep
        u3 = ops.empty(k1, na, nj)
        vh3 = ops.empty(nj, nb, k2)
        lam = ops.empty(nj)
        inv_lam = ops.empty(nj)
        ops.svd_emit(h3, u=u3, u_na=1, u_so=nj, u_sa=0, u_sj=1, vh=vh3, lam=lam,
                     inv_lam=inv_lam)
        # -- new gammas with the inverted outer lambdas (:533-559)
        new_l = ops.empty(nl, na, pl, nj)
        ops.gemm(pl, nj, k1, View(u1, row=k1, col=1, b1=pl * k1),
                 View(u3, row=na * nj, col=1, b2=nj),
                 View(new_l, row=nj, col=1, b1=na * pl * nj, b2=pl * nj), nb1=nl, nb2=na,
                 scale=View(inv_l, b1=1))
        rts = self._scale_cols(right_temp, inv_r, k2 * pr, nr, ops)
        new_r = ops.empty(nj, nb, pr, nr)
        ops.gemm(nj, pr * nr, k2, View(vh3, row=nb * k2, col=1, b1=k2),
                 View(rts, row=pr * nr, col=1),
                 View(new_r, row=nb * pr * nr, col=1, b1=pr * nr), nb1=nb)
        share(new_l, new_r, lam, inv_lam)
        return sl, new_l, new_r, lam, inv_lam

    # -------------------------------------------------------------- process tensors
    def apply_process_tensors(self, step, process_tensors):
        """Contract the step-1 PT-MPO site of every chain site into its gamma (:158-175).
        A device-resident process tensor (oqupy_b200.DeviceProcessTensor, or a host process
        tensor with rank-3 sites and no transforms, uploaded once: rank-3 site = delta
        between the system legs, process_tensor.py:346-347) is applied without ever
        forming the 4-leg tensor; any other process tensor through its public 4-leg
        ``get_mpo_tensor``."""
        ops = self._ops
        for site in range(self._n):
            pt = process_tensors[site]
            gam = self._gammas[site]
            nl, d2, npt, nr = gam.shape
            t3 = None
            dev = as_device_process_tensor(pt, ops)
            if dev is not None and step - 1 < len(dev):
                t3 = dev.get_mpo_tensor_device(step - 1)
            if t3 is not None:
                assert t3.shape[0] == npt and t3.shape[2] == d2
                nq = int(t3.shape[1])
                out = ops.empty(nl, d2, nq, nr)      # out[l,p,c,r] = sum_b T[b,c,p] gam[l,p,b,r]
                ops.gemm(nq, nr, npt, View(t3, row=d2, col=nq * d2, b2=1),
                         View(gam, row=nr, col=1, b1=d2 * npt * nr, b2=npt * nr),
                         View(out, row=nr, col=1, b1=d2 * nq * nr, b2=nq * nr),
                         nb1=nl, nb2=d2)
                self._gammas[site] = out
                continue
            t4 = pt.get_mpo_tensor(step - 1)
            if t4 is None:
                continue
            t4 = np.asarray(t4, dtype=CDTYPE)          # (b, c, p, q)
            if t4.ndim == 3:                           # delta between p and q
                t4 = np.einsum("bcp,pq->bcpq", t4, np.identity(t4.shape[2]))
            assert t4.shape[0] == npt and t4.shape[2] == d2
            nq, d2o = t4.shape[1], t4.shape[3]
            mdev = ops.from_host(t4.transpose(3, 1, 2, 0).reshape(d2o * nq, d2 * npt))
            out = ops.empty(nl, d2o, nq, nr)
            ops.gemm(d2o * nq, nr, d2 * npt, View(mdev, row=d2 * npt, col=1),
                     View(gam, row=nr, col=1, b1=d2 * npt * nr),
                     View(out, row=nr, col=1, b1=d2o * nq * nr), nb1=nl)
            self._gammas[site] = out

    # -------------------------------------------------------------- traces
    def clear_traces(self):                                            # :253-259
        self._bath_tr = None
        self._full_tr = None
        self._left_tr = None
        self._right_tr = None
        self._total = None

    def compute_traces(self, step, process_tensors):                   # :261-358
        ops = self._ops
        self.clear_traces()
        self._bath_tr, self._full_tr = [], []
        for site in range(self._n):
            gam = self._gammas[site]
            nl, d2, npt, nr = gam.shape
            cap = ops.from_host(np.asarray(process_tensors[site].get_cap_tensor(step),
                                           dtype=CDTYPE).reshape(-1))
            assert cap.shape[0] == npt
            bath = ops.empty(nl, d2, nr)               # sum_b gam[l,p,b,r] cap[b]
            ops.gemm(1, nr, npt, View(cap, col=1), View(gam, row=nr, col=1, b1=npt * nr),
                     View(bath, col=1, b1=nr), nb1=nl * d2)
            d = _isqrt(d2)
            tcap = ops.from_host(np.identity(d, dtype=CDTYPE).reshape(-1))
            full = ops.empty(nl, nr)                   # sum_p bath[l,p,r] vec(1)[p]
            ops.gemm(1, nr, d2, View(tcap, col=1), View(bath, row=nr, col=1, b1=d2 * nr),
                     View(full, col=1, b1=nr), nb1=nl)
            self._bath_tr.append(bath)
            self._full_tr.append(full)
        left = self._lams[0]
        self._left_tr = [left]
        for site in range(self._n):
            full = self._full_tr[site]
            nl, nr = full.shape
            tmp = ops.empty(nr)
            ops.gemm(1, nr, nl, View(left, col=1), View(full, row=nr, col=1),
                     View(tmp, col=1))
            left = self._scale_cols(tmp, self._lams[site + 1], 1, nr)
            if site == self._n - 1:
                tot = ops.empty(1)
                ops.gemm(1, 1, nr, View(left, col=1), View(ops.one), View(tot))
                self._total = complex(ops.to_host(tot)[0])
            else:
                self._left_tr.append(left)
        right = self._lams[-1]
        self._right_tr = [right]
        for site in range(self._n - 1, 0, -1):
            full = self._full_tr[site]
            nl, nr = full.shape
            tmp = ops.empty(nl)
            ops.gemm(nl, 1, nr, View(full, row=nr, col=1), View(right, row=1),
                     View(tmp, row=1))
            right = self._scale_cols(tmp, self._lams[site], 1, nl)
            self._right_tr.insert(0, right)

    def get_norm(self):
        """Total trace of the current chain state. """
        return complex(self._total)

    def get_site_density_matrix(self, site):                           # :364-376
        return self.get_density_matrix([site])

    def get_density_matrix(self, sites):
        """Reduced density matrix of a sorted list of sites (:378-445)."""
        assert isinstance(sites, list)
        assert len(sites) >= 1
        assert sites == sorted(sites)
        assert self._bath_tr is not None, "compute_traces() first"
        ops = self._ops
        cur = self._left_tr[sites[0]]                  # (rows = 1, bond)
        rows = 1
        dims = []
        for a, b in zip(sites, sites[1:] + [None]):
            bath = self._bath_tr[a]
            nl, d2, nr = bath.shape
            nxt = ops.empty(rows * d2, nr)
            ops.gemm(rows, d2 * nr, nl, View(cur, row=nl, col=1),
                     View(bath, row=d2 * nr, col=1), View(nxt, row=d2 * nr, col=1))
            cur, rows = nxt, rows * d2
            dims.append(_isqrt(d2))
            if b is None:
                break
            cur = self._scale_cols(cur, self._lams[a + 1], rows, nr)
            for i in range(a + 1, b):
                full = self._full_tr[i]
                fl, fr = full.shape
                nxt = ops.empty(rows, fr)
                ops.gemm(rows, fr, fl, View(cur, row=fl, col=1), View(full, row=fr, col=1),
                         View(nxt, row=fr, col=1))
                cur = self._scale_cols(nxt, self._lams[i + 1], rows, fr)
        right = self._right_tr[sites[-1]]
        nr = right.shape[0]
        vec = ops.empty(rows)
        ops.gemm(rows, 1, nr, View(cur, row=nr, col=1), View(right, row=1), View(vec, row=1))
        dm = ops.to_host(vec)
        k = len(dims)
        dm = dm.reshape([x for d in dims for x in (d, d)])
        perm = [2 * i for i in range(k)] + [2 * i + 1 for i in range(k)]
        tot = int(np.prod(dims))
        return dm.transpose(perm).reshape(tot, tot)
