"""TEST-ONLY stand-in for ``oqupy_b200._lib.CudaOps`` so that the HOST LOGIC of the
engine (index bookkeeping of oqupy_b200/chain.py, backends.py, process_tensor.py)
can be checked on the CPU-only build container.

This is NOT a product fallback: nothing under ``oqupy_b200/`` imports it, there is
no switch that selects it, and the product classes raise without CUDA.  It honours
the exact strided-view contract of the C-ABI (include/oqupy_b200.h) with torch CPU
tensors, so a wrong stride in the host code fails here exactly as it would on the
device.
"""
import numpy as np
import torch


class _Handle:
    pass


class HostModelOps:
    name = "host-model"

    def __init__(self):
        self.one = torch.ones(1, dtype=torch.complex128)
        self.svd_log = []
        self.launches = 0

    def empty(self, *shape):
        # NaN-filled: reading an element the kernels never wrote shows up at once
        return torch.full(shape, complex(float("nan"), float("nan")),
                          dtype=torch.complex128)

    def from_host(self, array):
        a = np.ascontiguousarray(np.asarray(array, dtype=np.complex128))
        return torch.from_numpy(a.copy())

    def to_host(self, tensor):
        return tensor.contiguous().numpy().copy()

    @staticmethod
    def _view(v, shape, strides):
        base = v.t
        # zero-size safe strided view on the tensor's storage
        return torch.as_strided(base, shape, strides,
                                base.storage_offset() + v.off)

    def gemm(self, m, n, k, a, b, c, nb1=1, nb2=1, scale=None, accumulate=False):
        self.launches += 1
        av = self._view(a, (nb1, nb2, m, k), (a.b1, a.b2, a.row, a.col))
        bv = self._view(b, (nb1, nb2, k, n), (b.b1, b.b2, b.row, b.col))
        if a.conj:
            av = av.conj()
        if b.conj:
            bv = bv.conj()
        res = torch.matmul(av, bv)
        if scale is not None:
            sv = self._view(scale, (nb1, nb2), (scale.b1, scale.b2))
            res = res * sv[:, :, None, None]
        cv = self._view(c, (nb1, nb2, m, n), (c.b1, c.b2, c.row, c.col))
        if accumulate:
            cv += res
        else:
            cv.copy_(res)

    def svd_factor(self, theta, m, n, rs, cs, eps, off=0, rin=1, rsi=0, cin=1, csi=0,
                   cos_tol=0.0):
        self.launches += 4
        assert m % rin == 0 and n % cin == 0
        mat = torch.as_strided(theta, (m // rin, rin, n // cin, cin), (rs, rsi, cs, csi),
                               theta.storage_offset() + off).reshape(m, n).numpy()
        u, s, vh = np.linalg.svd(mat, full_matrices=False)
        if eps is None:
            keep = s.size
        else:
            tail = np.sqrt(np.cumsum(np.square(s[::-1])))
            keep = int(np.count_nonzero(tail > eps * s[0]))
        h = _Handle()
        h.m, h.n, h.keep, h.sweeps, h.status = m, n, keep, 0, 0
        h.u, h.s, h.vh = u, s, vh
        self.svd_log.append((m, n, keep, 0))
        return h

    def svd_emit(self, h, u=None, u_na=1, u_so=0, u_sa=0, u_sj=0, svh=None, vh=None,
                 lam=None, inv_lam=None):
        self.launches += 1
        k = h.keep
        if vh is not None:
            vh.view(k, h.n).copy_(torch.from_numpy(np.ascontiguousarray(h.vh[:k])))
        if lam is not None:
            lam.copy_(torch.from_numpy(h.s[:k].astype(np.complex128)))
        if inv_lam is not None:
            inv_lam.copy_(torch.from_numpy((1.0 / h.s[:k]).astype(np.complex128)))
        if u is not None:
            no = h.m // u_na
            uv = torch.as_strided(u, (no, u_na, k), (u_so, u_sa, u_sj),
                                  u.storage_offset())
            uv.copy_(torch.from_numpy(
                np.ascontiguousarray(h.u[:, :k]).reshape(no, u_na, k)))
        if svh is not None:
            svh.view(k, h.n).copy_(torch.from_numpy(h.s[:k, None] * h.vh[:k]))

    def svd_values(self, h):
        return h.s.copy()

    def dyn_step(self, nvec, chi_l, chi_r, d2, t, p1, p2, v, v_out, cap=None,
                 rho_out=None):
        self.launches += 3
        if cap is not None and rho_out is not None:
            rho_out.copy_(torch.einsum("l,eli->ei", cap, v))
        u = torch.einsum("exi,eli->elx", p1, v)
        w = torch.einsum("lrx,elx->erx", t, u)
        v_out.copy_(torch.einsum("ejx,erx->erj", p2, w))

    def caps_step(self, chi_l, chi_r, d2, t, cap_next, tr2, cap_out):
        self.launches += 1
        cap_out.copy_(torch.einsum("lrx,r,x->l", t, cap_next, tr2))

    def launch_count(self):
        return self.launches

    def synchronize(self):
        pass
