"""CPU: the numpy model of the rank-revealing QR front end (tests/qr_model.py, which mirrors
oqupy_b200/csrc/qrcp.cuh) against the oracle's truncated SVD on graded operands."""
import numpy as np
import pytest

from oracle import tempo_np
import qr_model


def graded(rng, m, n, lo=-20.0):
    k = min(m, n)
    s = np.sort(10.0 ** rng.uniform(lo, 0.0, size=k))[::-1]
    s[0] = 1.0
    q1, _ = np.linalg.qr(rng.normal(size=(m, k)) + 1j * rng.normal(size=(m, k)))
    q2, _ = np.linalg.qr(rng.normal(size=(n, k)) + 1j * rng.normal(size=(n, k)))
    return (q1 * s) @ q2.conj().T


@pytest.mark.parametrize("m,n", [(60, 52), (52, 60), (96, 96), (130, 70), (70, 130)])
@pytest.mark.parametrize("eps", [1e-7, 1e-9])
def test_model_matches_oracle(m, n, eps):
    rng = np.random.default_rng(m * 131 + n)
    theta = graded(rng, m, n)
    u, svh, keep, k, s = qr_model.truncated_svd_model(theta, eps)
    u_ref, s_ref, vh_ref = tempo_np.truncated_svd(theta, eps)[:3]
    assert keep == u_ref.shape[1]
    assert k <= min(m, n)
    ref = (u_ref * s_ref) @ vh_ref
    np.testing.assert_allclose(u @ svh, ref, atol=5e-14)
    np.testing.assert_allclose(s[:keep], s_ref, rtol=1e-6, atol=1e-15)
    np.testing.assert_allclose(u.conj().T @ u, np.eye(keep), atol=1e-4)


@pytest.mark.parametrize("m,n,grid,panel", [(96, 90, 23, 8), (70, 130, 18, 4), (130, 120, 30, 8)])
def test_panel_pivoting_matches_oracle(m, n, grid, panel):
    """The kernel's relaxed (panel) pivoting gives the same truncated factorisation."""
    rng = np.random.default_rng(m + 7 * n)
    theta = graded(rng, m, n)
    eps = 1e-9
    u, svh, keep, k, s = qr_model.truncated_svd_model(theta, eps, grid=grid, panel=panel)
    u_ref, s_ref, vh_ref = tempo_np.truncated_svd(theta, eps)[:3]
    assert keep == u_ref.shape[1]
    np.testing.assert_allclose(u @ svh, (u_ref * s_ref) @ vh_ref, atol=5e-14)
    x = theta.conj().T if m < n else theta
    a, perm, tau, kk, tail2, shakes = qr_model.qrcp_panel(x, 1e-14, grid, panel)
    assert kk == k and shakes < k      # several pivots per hand-shake
    # a valid QR of the permuted operand whatever the pivot order
    r = np.zeros((k, x.shape[1]), dtype=complex)
    for pos in range(x.shape[1]):
        top = min(pos + 1, k)
        r[:top, pos] = a[:top, perm[pos]]
    np.testing.assert_allclose(qr_model.apply_q(a, perm, tau, k, r), x[:, perm], atol=1e-12)


def test_model_qr_identities():
    rng = np.random.default_rng(3)
    x = graded(rng, 80, 64, lo=-6.0)
    a, perm, tau, k, tail2 = qr_model.qrcp_stopped(x, 0.0)
    assert k == 64 and tail2 == 0.0
    r = np.triu(a[:64, perm])
    qfull = qr_model.apply_q(a, perm, tau, k, np.eye(64, dtype=complex))
    np.testing.assert_allclose(qfull.conj().T @ qfull, np.eye(64), atol=1e-13)
    np.testing.assert_allclose(qfull @ r, x[:, perm], atol=1e-13)
    # diagonal of R is real and non-increasing in magnitude (column pivoting)
    d = np.abs(np.diag(r))
    assert np.all(np.abs(np.diag(r).imag) < 1e-300) and np.all(d[:-1] >= d[1:] * (1 - 1e-12))
