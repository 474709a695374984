"""GPU: every C-ABI kernel against a plain torch complex128 reference of the same op."""
import numpy as np
import pytest
import torch

from oqupy_b200._lib import View, default_ops

pytestmark = pytest.mark.gpu


def rnd(rng, *shape):
    return rng.normal(size=shape) + 1j * rng.normal(size=shape)


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (5, 7, 3), (64, 64, 8), (130, 67, 45),
                                   (257, 300, 129)])
def test_zgemm_plain(m, n, k):
    ops = default_ops()
    rng = np.random.default_rng(m * 1000 + n)
    a, b = rnd(rng, m, k), rnd(rng, k, n)
    da, db = ops.from_host(a), ops.from_host(b)
    c = ops.empty(m, n)
    ops.gemm(m, n, k, View(da, row=k, col=1), View(db, row=n, col=1),
             View(c, row=n, col=1))
    np.testing.assert_allclose(ops.to_host(c), a @ b, atol=1e-11 * k)
    # transposed / conjugated operands and accumulate
    c2 = ops.from_host(np.ones((n, m)))
    ops.gemm(n, m, k, View(db, row=1, col=n, conj=True),
             View(da, row=1, col=k), View(c2, row=m, col=1), accumulate=True)
    np.testing.assert_allclose(ops.to_host(c2), 1 + b.conj().T @ a.T,
                               atol=1e-11 * k)


def test_zgemm_zip_layout():
    """The PT zip-up contraction: Theta[k,y,l,e] = M[e,y] sum_r C[k,r,e] A[l,y,r]."""
    ops = default_ops()
    rng = np.random.default_rng(5)
    nk, nr, nl, d2 = 37, 29, 41, 4
    carry, a, mat = rnd(rng, nk, nr, d2), rnd(rng, nl, d2, nr), rnd(rng, d2, d2)
    dc, da, dm = ops.from_host(carry), ops.from_host(a), ops.from_host(mat)
    theta = ops.empty(nk, d2, nl, d2)
    ops.gemm(nk, nl, nr, View(dc, row=nr * d2, col=d2, b2=1),
             View(da, row=1, col=d2 * nr, b1=nr),
             View(theta, row=d2 * nl * d2, col=d2, b1=nl * d2, b2=1),
             nb1=d2, nb2=d2, scale=View(dm, b1=1, b2=d2))
    ref = np.einsum("ey,kre,lyr->kyle", mat, carry, a)
    np.testing.assert_allclose(ops.to_host(theta), ref, atol=1e-10)


def graded(rng, m, n, lo=-14.0):
    k = min(m, n)
    s = np.sort(10.0 ** rng.uniform(lo, 0.0, size=k))[::-1]
    s[0] = 1.0
    q1, _ = np.linalg.qr(rnd(rng, m, k))
    q2, _ = np.linalg.qr(rnd(rng, n, k))
    return (q1 * s) @ q2.conj().T, s


def ref_keep(s, eps):
    tail = np.sqrt(np.cumsum(np.square(s[::-1])))
    return int(np.count_nonzero(tail > eps * s[0]))


@pytest.mark.parametrize("m,n", [(4, 4), (4, 16), (16, 4), (3, 5), (9, 7), (33, 33),
                                 (64, 48), (48, 130), (200, 96), (96, 200),
                                 (260, 260), (520, 130), (640, 600), (500, 900)])
@pytest.mark.parametrize("eps", [1e-7, None])
def test_trunc_svd(m, n, eps):
    ops = default_ops()
    rng = np.random.default_rng(m * 7919 + n)
    mat, _ = graded(rng, m, n)
    dm = ops.from_host(mat)
    h = ops.svd_factor(dm, m, n, n, 1, eps)
    s_ref = np.linalg.svd(mat, compute_uv=False)
    s = ops.svd_values(h)
    k = h.keep
    assert k == (min(m, n) if eps is None else ref_keep(s_ref, eps))
    # absolute accuracy relative to s0 (what the truncation rule needs).  With a
    # truncation eps, columns below 1e-2*eps*||A||_F are deliberately left unresolved
    # among themselves: only their joint Frobenius norm (the tail) is meaningful.
    if eps is None:
        np.testing.assert_allclose(s, s_ref, atol=2e-13 * s_ref[0], rtol=1e-9)
    else:
        rel = s_ref > 2e-2 * eps * np.linalg.norm(s_ref)
        np.testing.assert_allclose(s[rel], s_ref[rel], atol=2e-13 * s_ref[0], rtol=1e-7)
        np.testing.assert_allclose(np.linalg.norm(s[k:]), np.linalg.norm(s_ref[k:]),
                                   rtol=1e-7, atol=1e-15 * s_ref[0])
    u, svh = ops.empty(m, k), ops.empty(k, n)
    ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sj=1, svh=svh)
    u, svh = ops.to_host(u), ops.to_host(svh)
    # kept columns with sigma ~ eps*s0 are orthogonal to the rest only down to the
    # absolute rounding floor (~1e-14 s0), i.e. to ~1e-14/sigma relative
    otol = 1e-12 if eps is None else 1e-5
    if eps is not None:
        np.testing.assert_allclose(u.conj().T @ u, np.eye(k), atol=otol)
    ur, sr, vhr = np.linalg.svd(mat, full_matrices=False)
    best = (ur[:, :k] * sr[:k]) @ vhr[:k]
    np.testing.assert_allclose(u @ svh, best, atol=2e-13)
    # rows of S Vh are orthogonal with norms s
    g = svh @ svh.conj().T
    np.testing.assert_allclose(g, np.diag(s[:k] ** 2), atol=1e-12)


def test_trunc_svd_strided_and_emit_layout():
    """Transposed input view and the (j, y, k) site layout used by the PT zip-up."""
    ops = default_ops()
    rng = np.random.default_rng(11)
    nk, ny, n = 13, 4, 40
    mat, _ = graded(rng, nk * ny, n, lo=-9.0)
    dt = ops.from_host(np.ascontiguousarray(mat.T))      # stored transposed
    h = ops.svd_factor(dt, nk * ny, n, 1, nk * ny, 1e-6)
    k = h.keep
    site, svh = ops.empty(k, ny, nk), ops.empty(k, n)
    ops.svd_emit(h, u=site, u_na=ny, u_so=1, u_sa=nk, u_sj=ny * nk, svh=svh)
    site, svh = ops.to_host(site), ops.to_host(svh)
    u = site.transpose(2, 1, 0).reshape(nk * ny, k)
    ur, sr, vhr = np.linalg.svd(mat, full_matrices=False)
    np.testing.assert_allclose(u @ svh, (ur[:, :k] * sr[:k]) @ vhr[:k], atol=1e-12)


def test_trunc_svd_rank_deficient_and_zero_columns():
    ops = default_ops()
    rng = np.random.default_rng(2)
    a = rnd(rng, 60, 5) @ rnd(rng, 5, 44)          # exact rank 5
    a[:, 7] = 0.0
    h = ops.svd_factor(ops.from_host(a), 60, 44, 44, 1, 1e-9)
    assert h.keep == 5
    s = ops.svd_values(h)
    np.testing.assert_allclose(s[:5], np.linalg.svd(a, compute_uv=False)[:5],
                               rtol=1e-12)


def test_dyn_and_caps():
    ops = default_ops()
    rng = np.random.default_rng(8)
    for nvec, cl, cr, d2 in [(1, 1, 4, 4), (1, 57, 63, 4), (6, 130, 90, 4),
                             (3, 20, 33, 9)]:
        t, v = rnd(rng, cl, cr, d2), rnd(rng, nvec, cl, d2)
        p1, p2 = rnd(rng, nvec, d2, d2), rnd(rng, nvec, d2, d2)
        cap, capn, tr2 = rnd(rng, cl), rnd(rng, cr), rnd(rng, d2)
        dv_out, drho, dcap = ops.empty(nvec, cr, d2), ops.empty(nvec, d2), ops.empty(cl)
        dt = ops.from_host(t)
        ops.dyn_step(nvec, cl, cr, d2, dt, ops.from_host(p1), ops.from_host(p2),
                     ops.from_host(v), dv_out, cap=ops.from_host(cap), rho_out=drho)
        u = np.einsum("exi,eli->elx", p1, v)
        w = np.einsum("lrx,elx->erx", t, u)
        np.testing.assert_allclose(ops.to_host(dv_out),
                                   np.einsum("ejx,erx->erj", p2, w), atol=1e-10)
        np.testing.assert_allclose(ops.to_host(drho),
                                   np.einsum("l,eli->ei", cap, v), atol=1e-11)
        ops.caps_step(cl, cr, d2, dt, ops.from_host(capn), ops.from_host(tr2), dcap)
        np.testing.assert_allclose(ops.to_host(dcap),
                                   np.einsum("lrx,r,x->l", t, capn, tr2), atol=1e-10)
    assert ops.launch_count() > 0


@pytest.mark.parametrize("m,n", [(1300, 1200), (2300, 2200)])
def test_trunc_svd_large_multichunk(m, n):
    """Slices taller than one shared-memory chunk (256 rows): 1300x1200 runs 2 X slices
    (3 chunks each, DMMA Gram/apply) + 1 W slice; 2300x2200 has more pair slots than
    SMs / 2, i.e. ONE slice that mixes X and W rows (FMA Gram, chunked)."""
    ops = default_ops()
    rng = np.random.default_rng(m + n)
    k = min(m, n)
    # graded spectrum without forming dense random orthogonal factors on the host
    s = np.sort(10.0 ** rng.uniform(-12.0, 0.0, size=k))[::-1]
    s[0] = 1.0
    a = rnd(rng, m, 48) @ (rnd(rng, 48, n) * 1e-3)
    q1, _ = np.linalg.qr(rnd(rng, m, 64))
    q2, _ = np.linalg.qr(rnd(rng, n, 64))
    mat = (q1 * s[:64]) @ q2.conj().T + 1e-7 * a
    eps = 1e-6
    h = ops.svd_factor(ops.from_host(mat), m, n, n, 1, eps)
    s_ref = np.linalg.svd(mat, compute_uv=False)
    assert h.keep == ref_keep(s_ref, eps)
    sv = ops.svd_values(h)
    kk = h.keep
    np.testing.assert_allclose(sv[:kk], s_ref[:kk], atol=2e-13 * s_ref[0], rtol=1e-7)
    np.testing.assert_allclose(np.linalg.norm(sv[kk:]), np.linalg.norm(s_ref[kk:]),
                               rtol=1e-6, atol=1e-15 * s_ref[0])
    u, svh = ops.empty(m, kk), ops.empty(kk, n)
    ops.svd_emit(h, u=u, u_na=1, u_so=kk, u_sj=1, svh=svh)
    u, svh = ops.to_host(u), ops.to_host(svh)
    ur, sr, vhr = np.linalg.svd(mat, full_matrices=False)
    np.testing.assert_allclose(u @ svh, (ur[:, :kk] * sr[:kk]) @ vhr[:kk], atol=1e-12)


@pytest.mark.parametrize("shape,rows", [((6, 4, 5, 7), (0, 2)), ((5, 4, 9, 3), (0, 1)),
                                        ((13, 4, 11, 17), (0, 2)), ((3, 4, 40, 30), (2, 3))])
def test_svd_split_strides_and_parts(shape, rows):
    """b200_svd_factor2 / b200_svd_emit_parts: a rank-4 tensor factorised in place with
    rows = two of its legs, columns = the other two (the PT-TEBD splits,
    pt_tebd_backend.py:487-531); U, Vh, lambda, 1/lambda separately."""
    ops = default_ops()
    rng = np.random.default_rng(sum(shape))
    t = rnd(rng, *shape)
    cols = tuple(a for a in range(4) if a not in rows)
    strides = [int(np.prod(shape[a + 1:])) for a in range(4)]
    m = shape[rows[0]] * shape[rows[1]]
    n = shape[cols[0]] * shape[cols[1]]
    mat = t.transpose(rows + cols).reshape(m, n)
    dt = ops.from_host(t)
    h = ops.svd_factor(dt, m, n, strides[rows[0]], strides[cols[0]], 1e-9,
                       rin=shape[rows[1]], rsi=strides[rows[1]],
                       cin=shape[cols[1]], csi=strides[cols[1]])
    s_ref = np.linalg.svd(mat, compute_uv=False)
    k = h.keep
    assert k == ref_keep(s_ref, 1e-9)
    u, vh = ops.empty(m, k), ops.empty(k, n)
    lam, inv = ops.empty(k), ops.empty(k)
    ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sa=0, u_sj=1, vh=vh, lam=lam, inv_lam=inv)
    hu, hvh, hl, hi = (ops.to_host(x) for x in (u, vh, lam, inv))
    np.testing.assert_allclose(hl.real, s_ref[:k], atol=1e-12 * s_ref[0])
    np.testing.assert_allclose(hl * hi, np.ones(k), atol=1e-13)
    np.testing.assert_allclose((hu * hl) @ hvh, mat, atol=1e-11 * s_ref[0])
    np.testing.assert_allclose(hu.conj().T @ hu, np.eye(k), atol=1e-10)
    np.testing.assert_allclose(hvh @ hvh.conj().T, np.eye(k), atol=1e-10)


@pytest.mark.parametrize("m,n", [(300, 280), (1500, 96), (96, 1500), (700, 130)])
def test_trunc_svd_relative_mode_qr_path(m, n):
    """cos_tol > 0 (PT-TEBD): the rank-revealing QR front end in the RELATIVE-accuracy mode,
    square and tall splits.  Graded operand over 12 decades, eps = 1e-5: same `keep` as
    LAPACK, kept singular values to 1e-9 RELATIVE, orthonormal factors, and the product
    lambda^-1-safe: U diag(lambda) Vh reproduces theta to 1e-12 s_0 beyond the truncation."""
    ops = default_ops()
    rng = np.random.default_rng(m + n)
    q = min(m, n)
    uu = np.linalg.qr(rnd(rng, m, q))[0]
    vv = np.linalg.qr(rnd(rng, n, q))[0]
    sv = np.logspace(0, -12, q)
    a = (uu * sv) @ vv.conj().T
    eps = 1e-5
    h = ops.svd_factor(ops.from_host(a), m, n, n, 1, eps, cos_tol=1e-12)
    assert ops.svd_plan(h)[0] == 1                       # the QR path ran
    s_ref = np.linalg.svd(a, compute_uv=False)
    k = h.keep
    assert k == ref_keep(s_ref, eps)
    u, vh = ops.empty(m, k), ops.empty(k, n)
    lam, inv = ops.empty(k), ops.empty(k)
    ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sa=0, u_sj=1, vh=vh, lam=lam, inv_lam=inv)
    hu, hvh, hl, hi = (ops.to_host(x) for x in (u, vh, lam, inv))
    np.testing.assert_allclose(hl.real, s_ref[:k], rtol=1e-9)
    np.testing.assert_allclose(hl * hi, np.ones(k), atol=1e-13)
    np.testing.assert_allclose(hu.conj().T @ hu, np.eye(k), atol=1e-10)
    np.testing.assert_allclose(hvh @ hvh.conj().T, np.eye(k), atol=1e-10)
    trunc = np.linalg.norm(s_ref[k:])
    err = np.linalg.norm((hu * hl) @ hvh - a)
    assert err <= trunc * (1 + 1e-6) + 1e-12 * s_ref[0], (err, trunc)


@pytest.mark.parametrize("m,n", [(1, 1), (1, 9), (9, 1), (2, 33), (33, 2)])
def test_trunc_svd_degenerate_shapes(m, n):
    """Vectors and 1x1 operands (the first and last sites of a chain)."""
    ops = default_ops()
    rng = np.random.default_rng(10 * m + n)
    a = rnd(rng, m, n)
    h = ops.svd_factor(ops.from_host(a), m, n, n, 1, 1e-9)
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert h.keep == ref_keep(s_ref, 1e-9)
    k = h.keep
    u, svh = ops.empty(m, k), ops.empty(k, n)
    ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sj=1, svh=svh)
    np.testing.assert_allclose(ops.to_host(u) @ ops.to_host(svh), a, atol=1e-13)
    np.testing.assert_allclose(ops.svd_values(h)[:k], s_ref[:k], rtol=1e-12)


def test_trunc_svd_zero_matrix_and_clusters():
    """s_0 = 0 keeps nothing (the only way the rule returns 0); exactly degenerate singular
    values (an isometry times a scalar) keep everything."""
    ops = default_ops()
    h = ops.svd_factor(ops.from_host(np.zeros((12, 7))), 12, 7, 7, 1, 1e-9)
    assert h.keep == 0
    ops.svd_emit(h, u=ops.empty(12, 1), u_na=1, u_so=1, u_sj=1, svh=ops.empty(1, 7))
    rng = np.random.default_rng(4)
    q, _ = np.linalg.qr(rnd(rng, 80, 40))
    a = 3.0 * q
    h = ops.svd_factor(ops.from_host(a), 80, 40, 40, 1, 1e-9)
    assert h.keep == 40
    np.testing.assert_allclose(ops.svd_values(h), 3.0 * np.ones(40), rtol=1e-13)
    u, svh = ops.empty(80, 40), ops.empty(40, 40)
    ops.svd_emit(h, u=u, u_na=1, u_so=40, u_sj=1, svh=svh)
    np.testing.assert_allclose(ops.to_host(u) @ ops.to_host(svh), a, atol=1e-13)
