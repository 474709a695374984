"""The numpy statement of the round-2 SVD front end (tools/study_precond.py: stopped column-pivoted
QR -> QR -> block Jacobi) against the oracle's truncated SVD
(tn.split_node_full_svd, oqupy/backends/node_array.py:262,285,541) on graded operands.
CPU only: this pins the specification the round-2 kernel will be tested against."""
import os
import sys

import numpy as np
import pytest
import scipy.linalg as sla

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
study = pytest.importorskip("study_precond")


def graded(rng, m, n, decades):
    k = min(m, n)
    s = np.sort(10.0 ** rng.uniform(-decades, 0.0, size=k))[::-1]
    s[0] = 1.0
    q1 = np.linalg.qr(rng.normal(size=(m, k)) + 1j * rng.normal(size=(m, k)))[0]
    q2 = np.linalg.qr(rng.normal(size=(n, k)) + 1j * rng.normal(size=(n, k)))[0]
    return (q1 * s) @ q2.conj().T


@pytest.mark.parametrize("m,n", [(96, 80), (40, 72), (64, 16)])
def test_pipeline_matches_oracle(m, n):
    theta = graded(np.random.default_rng(m * 1000 + n), m, n, 22.0)
    out = study.pipeline(theta)
    assert out["keep"] == out["keep_lapack"]
    assert out["product_vs_lapack_over_s0"] < 1e-12
    assert out["columns"] <= min(m, n)


def test_qrcp_stopped_matches_scipy():
    theta = graded(np.random.default_rng(7), 72, 56, 24.0)
    out = study.check_qrcp_stopped(theta)
    assert out["k"] == out["k_scipy"]
    assert out["resid_AP_minus_QR_over_norm"] < 1e-12
    # layout: R on and above the diagonal of the pivoted columns, reflectors below
    stop = 1e-5 * study.EPSREL * np.linalg.norm(theta)
    a, tau, perm, k, tail2 = study.qrcp_stopped(theta, stop)
    assert sorted(perm.tolist()) == list(range(theta.shape[1]))
    d = np.abs([a[i, perm[i]] for i in range(k)])
    assert np.all(d[:-1] >= d[1:] * (1 - 1e-8))           # pivots decrease
    r_ref = sla.qr(theta, mode="r", pivoting=True)[0]
    assert tail2 == pytest.approx(np.linalg.norm(r_ref[k:, k:]) ** 2, rel=0.1, abs=1e-40)   # the block sits at the rounding level
