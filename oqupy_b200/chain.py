"""Device-resident matrix-product chain: the two primitives of the hot path.

Host-side mirror of ``NodeArray.zip_up`` / ``NodeArray.svd_sweep``
(/root/reference/oqupy/backends/node_array.py:413-554, 226-299), re-designed for
the device:

* every MPS site is ONE contiguous complex128 device tensor ``(chi_l, a, chi_r)``;
* influence MPO sites are never materialised: ``B[l,x,y,r] = d_lr d_xy M[l,x]``
  (pt_tempo_backend.py:151) / ``B[w,n,s,e] = d_we d_ns M[n,w]``
  (tempo_backend.py:419,426) turn the three-tensor contraction into d2*d2
  independent (k x chi)(chi x chi') products scaled by one entry of ``M`` -- one
  strided batched GEMM launch that writes Theta directly in SVD layout;
* the eps-truncated SVD returns ``keep`` to the host (one sync per bond) so that
  every site keeps its exact shape.

All arithmetic happens inside ``ops`` (C-ABI kernels, see _lib.CudaOps).
"""
from ._lib import View


class PtSite:
    """One PT-TEMPO influence MPO site in implicit (delta-structured) form.

    kind: 'first'  B[x,y,r]   = d_xy d_xr vec[x]                (dk = 0)
                   with ``maps = (north_map, west_map)`` (unique=True,
                   pt_tempo_backend.py:125-137):
                   B[w,y,n] = [w = west_map[y]] [n = north_map[y]] vec[n]
          'mid'    B[l,x,y,r] = d_lr d_xy mat[l,x]
          'last'   B[l,0,y,0] = mat[l,y]          (newest site, array dim 1)
          'closed' B[l,x,y,0] = d_xy mat[l,x]     (mat already times the closing
                                                   vector, end phase)
    """
    __slots__ = ("kind", "mat", "maps")

    def __init__(self, kind, mat, maps=None):
        self.kind, self.mat, self.maps = kind, mat, maps


def _split(ops, theta, m, n, rs, cs, eps):
    return ops.svd_factor(theta, m, n, rs, cs, eps)


# ------------------------------------------------------------------ PT-TEMPO
def pt_zip_up_left(ops, mps, mpo, eps):
    """mps.zip_up(mpo, right_index=-1, direction='left')  (node_array.py:482-552).

    Theta[(k,y),(l,e)] = M[e,y] sum_r carry[k,r,e] A[l,y,r];  new site = U as
    (j, y, k), carry' = S Vh as (j, l, e).
    """
    nb = len(mpo)
    left = len(mps) - nb
    carry = None
    for ib in range(nb - 1, -1, -1):
        ia = left + ib
        a = mps[ia]
        nl, nx, nr = a.shape
        site = mpo[ib]
        mat = site.mat
        if site.kind == "first" and site.maps is not None:
            nmap, wmap = site.maps          # one small product per array value y
            ny = len(nmap)
            if carry is None:       # out[l,y,0] = vec[n(y)] A[l,w(y),0]
                assert nr == 1
                out = ops.empty(nl, ny, 1)
                for y in range(ny):
                    ops.gemm(nl, 1, 1, View(a, row=nx * nr, off=int(wmap[y]) * nr),
                             View(ops.one), View(out, row=ny, off=y),
                             scale=View(mat, off=int(nmap[y])))
            else:                   # out[l,y,k] = vec[n(y)] sum_r A[l,w(y),r] C[k,r,n(y)]
                nk, _, ne = carry.shape
                out = ops.empty(nl, ny, nk)
                for y in range(ny):
                    ops.gemm(nl, nk, nr,
                             View(a, row=nx * nr, col=1, off=int(wmap[y]) * nr),
                             View(carry, row=ne, col=nr * ne, off=int(nmap[y])),
                             View(out, row=ny * nk, col=1, off=y * nk),
                             scale=View(mat, off=int(nmap[y])))
            mps[ia] = out
            assert ib == 0
            break
        if site.kind == "first":
            ny = nx
            if carry is None:       # closed single-site MPO: out[l,y,0] = vec[y] A[l,y,0]
                assert nr == 1
                out = ops.empty(nl, ny, 1)
                ops.gemm(nl, 1, 1, View(a, row=nx * nr, b1=nr), View(ops.one),
                         View(out, row=ny, b1=1), nb1=ny, scale=View(mat, b1=1))
            else:                   # out[l,y,k] = vec[y] sum_r A[l,y,r] C[k,r,y]
                nk, _, ne = carry.shape
                assert ne == ny
                out = ops.empty(nl, ny, nk)
                ops.gemm(nl, nk, nr, View(a, row=nx * nr, col=1, b1=nr),
                         View(carry, row=ne, col=nr * ne, b1=1),
                         View(out, row=ny * nk, col=1, b1=nk), nb1=ny,
                         scale=View(mat, b1=1))
            mps[ia] = out
            assert ib == 0
            break
        ne, ny = mat.shape
        if carry is None:           # newest site: Theta[0,y,l,e] = M[e,y] A[l,x(y),0]
            assert nr == 1 and site.kind in ("last", "closed")
            nk = 1
            theta = ops.empty(1, ny, nl, ne)
            xs = 0 if site.kind == "last" else nr
            ops.gemm(1, nl, 1, View(ops.one), View(a, col=nx * nr, b1=xs),
                     View(theta, col=ne, b1=nl * ne, b2=1), nb1=ny, nb2=ne,
                     scale=View(mat, b1=1, b2=ny))
        else:
            assert site.kind == "mid" and ny == nx
            nk = carry.shape[0]
            theta = ops.empty(nk, ny, nl, ne)
            ops.gemm(nk, nl, nr, View(carry, row=nr * ne, col=ne, b2=1),
                     View(a, row=1, col=nx * nr, b1=nr),
                     View(theta, row=ny * nl * ne, col=ne, b1=nl * ne, b2=1),
                     nb1=ny, nb2=ne, scale=View(mat, b1=1, b2=ny))
        m, n = nk * ny, nl * ne
        h = _split(ops, theta, m, n, n, 1, eps)
        nj = h.keep
        new_site = ops.empty(nj, ny, nk)
        carry = ops.empty(nj, nl, ne)
        ops.svd_emit(h, u=new_site, u_na=ny, u_so=1, u_sa=nk, u_sj=ny * nk,
                     svh=carry)
        mps[ia] = new_site


def svd_sweep_right(ops, mps, fi, ti, eps):
    """svd_sweep(from_index=fi, to_index=ti), fi < ti  (node_array.py:252-273)."""
    for i in range(fi, ti):
        a = mps[i]
        nl, nx, nr = a.shape
        h = _split(ops, a, nl * nx, nr, nr, 1, eps)
        nj = h.keep
        u = ops.empty(nl, nx, nj)
        svh = ops.empty(nj, nr)
        ops.svd_emit(h, u=u, u_na=1, u_so=nj, u_sa=0, u_sj=1, svh=svh)
        mps[i] = u
        b = mps[i + 1]
        _, bx, br = b.shape
        nb = ops.empty(nj, bx, br)
        ops.gemm(nj, bx * br, nr, View(svh, row=nr, col=1),
                 View(b, row=bx * br, col=1), View(nb, row=bx * br, col=1))
        mps[i + 1] = nb


def svd_sweep_left(ops, mps, fi, ti, eps):
    """svd_sweep(from_index=fi, to_index=ti), fi > ti  (node_array.py:274-296)."""
    for i in range(fi, ti, -1):
        a = mps[i]
        nl, nx, nr = a.shape
        m = nx * nr
        h = _split(ops, a, m, nl, 1, m, eps)       # rows (x, r), cols l
        nj = h.keep
        u = ops.empty(nj, nx, nr)
        svh = ops.empty(nj, nl)
        ops.svd_emit(h, u=u, u_na=1, u_so=1, u_sa=0, u_sj=m, svh=svh)
        mps[i] = u
        b = mps[i - 1]
        bl, bx, _ = b.shape
        nb = ops.empty(bl, bx, nj)
        ops.gemm(bl * bx, nj, nl, View(b, row=nl, col=1),
                 View(svh, row=1, col=nl), View(nb, row=nj, col=1))
        mps[i - 1] = nb


# ------------------------------------------------------------------ TEMPO
class TempoSite:
    """One TEMPO influence MPO site, B[w,n,s,e] = d_we d_ns mat[n,w] in implicit form.

    kind: 'start' first aligned site, west leg summed: mat[s,e] = infl[s,e]*sum_west[e]
          'mid'   mat[s,e] = infl[s,e]
          'dense' last aligned site (dk=0), explicit tensor mat[(w,n),(s,e)]
    """
    __slots__ = ("kind", "mat", "nw", "ns")

    def __init__(self, kind, mat, nw=None, ns=None):
        self.kind, self.mat, self.nw, self.ns = kind, mat, nw, ns


def tempo_zip_up_right(ops, mps, mpo, eps):
    """mps.zip_up(mpo, left_index=0, right_index=-1, direction='right')
    (tempo_backend.py:539-547).  Theta[(k,s),(r,e)] = M[s,e] sum_l carry[k,l,e] A[l,s,r]."""
    nb = len(mpo)
    assert nb == len(mps)
    carry = None
    for ib in range(nb):
        a = mps[ib]
        nl, nn, nr = a.shape
        site = mpo[ib]
        mat = site.mat
        if ib == nb - 1:            # dense dk=0 site, no SVD
            assert site.kind == "dense" and nr == 1
            nw = site.nw
            nse = mat.shape[1]
            if carry is None:
                nk = 1
                assert nl == 1 and nw == 1
                cview = View(ops.one)
            else:
                nk = carry.shape[0]
                assert carry.shape[2] == nw
                cview = View(carry, row=nl * nw, col=nw, b1=1)
            tmp = ops.empty(nk, nw, nn)      # T[k,w,n] = sum_l C[k,l,w] A[l,n]
            ops.gemm(nk, nn, nl, cview, View(a, row=nn * nr, col=nr),
                     View(tmp, row=nw * nn, col=1, b1=nn), nb1=nw)
            ns = nn if site.ns is None else site.ns   # unique=True: s is the reduced leg
            out = ops.empty(nk, ns, nse // ns)        # (k, s, e): e dangles right
            ops.gemm(nk, nse, nw * nn, View(tmp, row=nw * nn, col=1),
                     View(mat, row=nse, col=1), View(out, row=nse, col=1))
            mps[ib] = out
            return out
        ns, ne = mat.shape
        assert ns == nn
        if carry is None:           # Theta[0,s,r,e] = A[0,s,r] M[s,e]
            assert site.kind == "start" and nl == 1
            nk = 1
            theta = ops.empty(1, ns, nr, ne)
            ops.gemm(1, nr, 1, View(ops.one), View(a, col=1, b1=nr),
                     View(theta, col=ne, b1=nr * ne, b2=1), nb1=ns, nb2=ne,
                     scale=View(mat, b1=ne, b2=1))
        else:
            assert site.kind == "mid"
            nk = carry.shape[0]
            theta = ops.empty(nk, ns, nr, ne)
            ops.gemm(nk, nr, nl, View(carry, row=nl * ne, col=ne, b2=1),
                     View(a, row=nn * nr, col=1, b1=nr),
                     View(theta, row=ns * nr * ne, col=ne, b1=nr * ne, b2=1),
                     nb1=ns, nb2=ne, scale=View(mat, b1=ne, b2=1))
        m, n = nk * ns, nr * ne
        h = _split(ops, theta, m, n, n, 1, eps)
        nj = h.keep
        new_site = ops.empty(nk, ns, nj)
        carry = ops.empty(nj, nr, ne)
        ops.svd_emit(h, u=new_site, u_na=1, u_so=nj, u_sa=0, u_sj=1, svh=carry)
        mps[ib] = new_site
    return None
