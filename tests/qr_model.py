"""numpy model of the rank-revealing front end of the truncated SVD (test infrastructure).

Follows oqupy_b200/csrc/qrcp.cuh statement by statement at the level of the data the kernels
exchange: the stopped column-pivoted Householder QR (zlarfg reflectors, H = I - tau v v^H,
exact column norms after every pivot, lowest physical column on ties), the Jacobi operand
L = [R11 R12]^H read out of the work array through the pivot permutation, the
back-transformation by the stored reflectors and the scatter of the other factor through the
permutation.  The Jacobi stage itself is stood in for by LAPACK (any orthogonal J with
L J = Y, orthogonal columns, is a valid outcome of it).

Used by tests/test_qr_model.py (CPU: the conventions against the oracle's truncated SVD) and
by tests/test_kernels_gpu.py / tools/qr_check.py (GPU: the kernels' intermediate arrays
against this model).  Nothing under oqupy_b200/ imports it.
"""
import numpy as np


def qrcp_stopped(x, stop_rel):
    """x: p x q (p >= q).  Returns a (p x q work array: R above/on the diagonal of the pivot
    columns, reflector below; R12 in rows < k of the other columns), perm (position ->
    physical column), tau (k), k, tail2 = ||R22||_F^2."""
    a = np.array(x, dtype=complex, order="F")
    p, q = a.shape
    done = np.zeros(q, dtype=bool)
    vn2 = np.sum(np.abs(a) ** 2, axis=0)
    stop2 = stop_rel ** 2 * float(np.sum(vn2))
    perm, taus = [], []
    k = 0
    for j in range(q):
        cand = np.where(done, -1.0, vn2)
        widx = int(np.argmax(cand))            # first maximum: lowest physical column on ties
        if not cand[widx] > stop2:
            break
        x_col = a[:, widx].copy()
        alpha = x_col[j]
        xnorm2 = float(np.sum(np.abs(x_col[j + 1:]) ** 2))
        beta, tau, scale = alpha.real, 0.0, 0.0
        if xnorm2 > 0.0 or alpha.imag != 0.0:
            an = np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm2)
            beta = -an if alpha.real >= 0.0 else an
            tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
            scale = 1.0 / (alpha - beta)
        v = np.zeros(p, dtype=complex)
        v[j] = 1.0
        v[j + 1:] = x_col[j + 1:] * scale
        done[widx] = True
        a[j, widx] = beta
        a[j + 1:, widx] = v[j + 1:]
        perm.append(widx)
        taus.append(tau)
        rest = np.where(~done)[0]
        if len(rest):
            w = v[j:].conj() @ a[j:, rest]
            a[j:, rest] -= np.conj(tau) * np.outer(v[j:], w)
            vn2[rest] = np.sum(np.abs(a[j + 1:, rest]) ** 2, axis=0)
        k = j + 1
    rest = np.where(~done)[0]
    tail2 = float(np.sum(vn2[rest])) if len(rest) else 0.0
    perm = np.array(perm + list(rest), dtype=int)
    return a, perm, np.array(taus, dtype=complex), k, tail2


def _trunc35(x):
    """norm^2 as the kernel publishes it: top 35 bits of the double."""
    b = np.float64(max(x, 0.0)).view(np.uint64)
    return float(((b >> np.uint64(29)) << np.uint64(29)).view(np.float64))


def qrcp_panel(x, stop_rel, grid, panel, theta=0.5):
    """The kernel's pivot selection (qrcp.cuh): column c belongs to CTA c % grid; per grid
    hand-shake every CTA offers its best remaining column, the `panel` largest offers above
    the stop level are factorised greedily by their exact remaining norms; a panel column is
    only taken while its remaining norm^2 exceeds max(stop^2, theta^2 * best offer outside
    the panel).  Returns (a, perm, tau, k, tail2, hand-shakes)."""
    a = np.array(x, dtype=complex, order="F")
    p, q = a.shape
    done = np.zeros(q, dtype=bool)
    vn2 = np.sum(np.abs(a) ** 2, axis=0)
    stop2 = stop_rel ** 2 * float(np.sum(vn2))
    owner = np.arange(q) % grid
    perm, taus = [], []
    j = shakes = 0
    while j < q:
        shakes += 1
        offers = []
        for g in range(min(grid, q)):
            mine = np.where((owner == g) & ~done)[0]
            if len(mine):
                c = mine[np.argmax(vn2[mine])]
                offers.append((_trunc35(vn2[c]), -int(c)))
        offers.sort(reverse=True)
        cand = [(-c, v) for v, c in offers[:panel] if v > stop2]
        if not cand:
            break
        outside = offers[panel][0] if len(offers) > panel and offers[panel][0] > stop2 else 0.0
        thr = max(stop2, theta * theta * outside)
        pan = [c for c, _ in cand]
        taken = 0
        while pan and j < q:
            rem = [float(np.sum(np.abs(a[j:, c]) ** 2)) for c in pan]
            t = int(np.argmax(rem))
            if taken and not rem[t] > thr:
                break
            widx = pan.pop(t)
            col = a[:, widx].copy()
            alpha = col[j]
            xnorm2 = float(np.sum(np.abs(col[j + 1:]) ** 2))
            beta, tau, scale = alpha.real, 0.0, 0.0
            if xnorm2 > 0.0 or alpha.imag != 0.0:
                an = np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm2)
                beta = -an if alpha.real >= 0.0 else an
                tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
                scale = 1.0 / (alpha - beta)
            v = np.zeros(p, dtype=complex)
            v[j] = 1.0
            v[j + 1:] = col[j + 1:] * scale
            done[widx] = True
            a[j, widx] = beta
            a[j + 1:, widx] = v[j + 1:]
            perm.append(widx)
            taus.append(tau)
            rest = np.where(~done)[0]
            if len(rest):
                w = v[j:].conj() @ a[j:, rest]
                a[j:, rest] -= np.conj(tau) * np.outer(v[j:], w)
            j += 1
            taken += 1
        rest = np.where(~done)[0]
        if len(rest):
            vn2[rest] = np.sum(np.abs(a[j:, rest]) ** 2, axis=0)
    rest = np.where(~done)[0]
    tail2 = float(np.sum(vn2[rest])) if len(rest) else 0.0
    return (a, np.array(perm + list(rest), dtype=int), np.array(taus, dtype=complex), j,
            tail2, shakes)


def l_operand(a, perm, k):
    """L[i, j] = conj(R[j, position i]) for j <= i (the Jacobi kernel's QR-mode loader)."""
    q = a.shape[1]
    lmat = np.zeros((q, k), dtype=complex)
    for i in range(q):
        jmax = min(i, k - 1)
        lmat[i, :jmax + 1] = a[:jmax + 1, perm[i]].conj()
    return lmat


def apply_q(a, perm, tau, k, jsel):
    """Q[:, :k] @ jsel with Q = H_0 ... H_{k-1}, H_r = I - tau_r v_r v_r^H (apply_q_kernel)."""
    p = a.shape[0]
    z = np.zeros((p, jsel.shape[1]), dtype=complex)
    z[:k] = jsel
    for r in range(k - 1, -1, -1):
        v = np.zeros(p, dtype=complex)
        v[r] = 1.0
        v[r + 1:] = a[r + 1:, perm[r]]
        z -= tau[r] * np.outer(v, v.conj() @ z)
    return z


def keep_rule(s, eps, tail2=0.0):
    tail = tail2
    keep = 0
    for x in s[::-1]:
        tail += x * x
        if np.sqrt(tail) > eps * s[0]:
            keep += 1
    return keep


def truncated_svd_model(theta, eps, stop_rel=None, grid=None, panel=1):
    """The whole QR path on theta (m x n): returns (u, svh, keep, k, sigma[:k]).  With `grid`
    the pivots are chosen the way the kernel does it (panels of up to `panel` per hand-shake)."""
    m, n = theta.shape
    transposed = m < n
    x = theta.conj().T if transposed else theta
    if stop_rel is None:
        stop_rel = 1e-5 * eps
    if grid is None:
        a, perm, tau, k, tail2 = qrcp_stopped(x, stop_rel)
    else:
        a, perm, tau, k, tail2, _ = qrcp_panel(x, stop_rel, grid, panel)
    lmat = l_operand(a, perm, k)
    ul, s, vlh = np.linalg.svd(lmat, full_matrices=False)      # L = ul s vlh; J = vlh^H
    j = vlh.conj().T
    y = ul * s                                                 # L J
    keep = keep_rule(s, eps, tail2)
    zq = apply_q(a, perm, tau, k, j[:, :keep])                 # p x keep
    q = x.shape[1]
    if not transposed:        # theta = X: U = Q J, S Vh[j, perm[i]] = conj(Y[i, j])
        u = zq
        svh = np.zeros((keep, n), dtype=complex)
        svh[:, perm[:q]] = y[:, :keep].conj().T
    else:                     # theta = X^H: U[perm[i], j] = Y[i, j] / s_j, S Vh = s (Q J)^H
        u = np.zeros((m, keep), dtype=complex)
        u[perm[:q], :] = y[:, :keep] / s[:keep]
        svh = (zq * s[:keep]).conj().T
    return u, svh, keep, k, s
