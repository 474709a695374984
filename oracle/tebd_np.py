"""TEST INFRASTRUCTURE ONLY -- numpy restatement ("port") of the PT-TEBD backend
``PtTebdBackend`` (/root/reference/oqupy/backends/pt_tebd_backend.py:46-565, SURVEY.md
8a row A7).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; nothing under oqupy_b200/ does.

Parity pinned: tests/golden/pt_tebd_F*.npz are produced by the UNMODIFIED reference
(tests/golden/make_golden_tebd.py) on its own test F (tests/physics/pt_tebd_test.py) and
carry the reference's golden density matrices tests/data/correct_results/example_F{1,2}_rhos.npy.

The augmented MPS is  lam[0] - Gam[0] - lam[1] - ... - Gam[n-1] - lam[n]  with
Gam (chi_l, d2, chi_pt, chi_r) and lam the DIAGONALS (lam[0] = lam[n] = ones), see
pt_tebd_backend.py:36-45.
"""
import numpy as np

from .tempo_np import truncated_svd

CDTYPE = np.complex128


def apply_nn_gate(lam_l, gam_l, lam_m, gam_r, lam_r, gate_l, gate_r, eps):
    """_apply_nn_gate (pt_tebd_backend.py:447-565).  gate_l (new_l, old_l, g), gate_r
    (g, new_r, old_r) (edge wiring :474-481).  Returns (new_gam_l, new_lam_m, new_gam_r)."""
    nl, d2l, pl, nm = gam_l.shape
    _, d2r, pr, nr = gam_r.shape
    # -- split off the process tensor legs (:487-507)
    left = lam_l[:, None, None, None] * gam_l
    u1, s1, vh1, _ = truncated_svd(left.transpose(0, 2, 1, 3).reshape(nl * pl, d2l * nm), eps)
    k1 = s1.size
    left_temp = u1.reshape(nl, pl, k1)
    left_mid = (s1[:, None] * vh1).reshape(k1, d2l, nm)
    right = gam_r * lam_r[None, None, None, :]
    u2, s2, vh2, _ = truncated_svd(right.reshape(nm * d2r, pr * nr), eps)
    k2 = s2.size
    right_mid = (u2 * s2[None, :]).reshape(nm, d2r, k2)
    right_temp = vh2.reshape(k2, pr, nr)
    # -- apply the gate (:508-519), split (:521-531)
    theta = np.einsum("kpm,apg,m,gbq,mqj->kabj", left_mid, gate_l, lam_m, gate_r, right_mid,
                      optimize=True)
    na, nb = theta.shape[1], theta.shape[2]
    u3, s3, vh3, _ = truncated_svd(theta.reshape(k1 * na, nb * k2), eps)
    nj = s3.size
    # -- contract with the inverted outer lambdas (:533-559)
    new_gam_l = np.einsum("l,lbk,kaj->labj", 1.0 / lam_l, left_temp, u3.reshape(k1, na, nj),
                          optimize=True)
    new_gam_r = np.einsum("jaq,qbr,r->jabr", vh3.reshape(nj, nb, k2), right_temp,
                          1.0 / lam_r, optimize=True)
    return new_gam_l, s3.astype(CDTYPE), new_gam_r


class PtTebdOracle:
    """PtTebdBackend (pt_tebd_backend.py:46-445) with plain arrays."""

    def __init__(self, gammas, lambdas, epsrel):
        assert len(gammas) == len(lambdas) + 1                       # :71
        self.n = len(gammas)
        self.eps = epsrel
        self.gammas = [np.array(g, dtype=CDTYPE) for g in gammas]
        self.lams = ([np.ones(gammas[0].shape[0], dtype=CDTYPE)]      # :97-102
                     + [np.array(lam, dtype=CDTYPE).reshape(-1) for lam in lambdas]
                     + [np.ones(gammas[-1].shape[3], dtype=CDTYPE)])
        self.clear_traces()

    # ---- evolution
    def apply_nn_gate(self, sites, tensors):                          # :226-231
        sl, sr = sites
        assert sr == sl + 1
        g_l, lam, g_r = apply_nn_gate(self.lams[sl], self.gammas[sl], self.lams[sl + 1],
                                      self.gammas[sr], self.lams[sr + 1],
                                      np.asarray(tensors[0], dtype=CDTYPE),
                                      np.asarray(tensors[1], dtype=CDTYPE), self.eps)
        self.gammas[sl], self.lams[sl + 1], self.gammas[sr] = g_l, lam, g_r

    def apply_nn_gate_layer(self, gates):                             # :134-155
        for sites, tensors in gates:
            self.apply_nn_gate(sites, tensors)

    def apply_site_gate(self, site, matrix):                          # :238-251
        self.gammas[site] = np.einsum("ap,lpbr->labr", np.asarray(matrix, dtype=CDTYPE),
                                      self.gammas[site])

    def apply_process_tensors(self, step, mpo_tensors):               # :158-175
        """mpo_tensors[site]: the 4-leg (b, b', p, p') tensor of step-1, rank-3 (b, b', p)
        meaning a delta between p and p' (process_tensor.py:346-347), or None."""
        for site, t in enumerate(mpo_tensors):
            if t is None:
                continue
            t = np.asarray(t, dtype=CDTYPE)
            if t.ndim == 3:
                self.gammas[site] = np.einsum("lpbr,bcp->lpcr", self.gammas[site], t)
            else:
                self.gammas[site] = np.einsum("lpbr,bcpq->lqcr", self.gammas[site], t)

    # ---- read-out
    def clear_traces(self):                                           # :253-259
        self.bath_tr = self.full_tr = self.left_tr = self.right_tr = self.total = None

    def compute_traces(self, caps):                                   # :261-358
        """caps[site]: cap vector of the site's process tensor at the current step."""
        self.bath_tr = [np.einsum("lpbr,b->lpr", g, np.asarray(c, dtype=CDTYPE))
                        for g, c in zip(self.gammas, caps)]
        self.full_tr = []
        for g in self.bath_tr:
            d = int(round(np.sqrt(g.shape[1])))
            self.full_tr.append(np.einsum("lpr,p->lr", g, np.identity(d).reshape(-1)))
        left = self.lams[0].copy()
        self.left_tr = [left]
        for s in range(self.n):
            left = (left @ self.full_tr[s]) * self.lams[s + 1]
            if s == self.n - 1:
                self.total = left.sum()
            else:
                self.left_tr.append(left)
        right = self.lams[-1].copy()
        self.right_tr = [right]
        for s in range(self.n - 1, 0, -1):
            right = self.lams[s] * (self.full_tr[s] @ right)
            self.right_tr.insert(0, right)

    def get_norm(self):
        return complex(self.total)

    def get_bond_dimensions(self):
        return np.array([g.shape[3] for g in self.gammas[:-1]])

    def get_density_matrix(self, sites):                              # :378-445
        s = list(sites)
        assert s == sorted(s) and len(s) >= 1
        cur = self.left_tr[s[0]][None, :]                 # (phys..., bond)
        for a, b in zip(s, s[1:] + [None]):
            cur = np.tensordot(cur, self.bath_tr[a], (-1, 0))         # (..., p, r)
            if b is None:
                break
            m = np.diag(self.lams[a + 1])
            for i in range(a + 1, b):
                m = (m @ self.full_tr[i]) * self.lams[i + 1][None, :]
            cur = cur @ m
        cur = (cur @ self.right_tr[s[-1]])[0]             # (p_1, ..., p_k)
        dims = [int(round(np.sqrt(x))) for x in cur.shape]
        k = len(dims)
        cur = cur.reshape([x for d in dims for x in (d, d)])
        perm = [2 * i for i in range(k)] + [2 * i + 1 for i in range(k)]
        tot = int(np.prod(dims))
        return cur.transpose(perm).reshape(tot, tot)
