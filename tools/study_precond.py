"""CPU study for DESIGN.md section 7 item 1: does a QR preconditioner cut the outer sweep count
of the block one-sided Jacobi SVD on the operands of the config-2 PT-TEMPO build?

The emulation follows the kernel's structure (svd.cu): blocks of 16 columns, block pairs in a
round-robin tournament, ONE cyclic sweep of two-sided Jacobi on the 32x32 Gram matrix of a
pair (tools/micro/jacobi_inner_emu.c, compiled on first use; columns sorted by norm
afterwards) -- a LAPACK eigh as inner solve is NOT usable for this study: it resolves the
Gram matrix only to eps*||G||, keeps re-mixing the columns below that level and the outer
iteration never terminates on these graded operands (16 decades) --
the kernel's convergence rule  |g_ij|^2 > big*(tol^2*small + floor^2),  tol = 1e-11,
floor = 8*eps_mach*||X||_F, and one final rotation-free sweep counted like the kernel does.

Operands: the SVD inputs of one step of the config-2 build, captured from the CPU oracle
(test infrastructure; this tool is a study, not a product path).

Variants
  plain       Jacobi on the columns of Theta (m >= n after the kernel's own transposition)
  sort        columns pre-sorted by decreasing norm
  qr          Theta = Q R (no pivoting), Jacobi on the columns of L = R^H
  sort+qr     norm-sorted columns, then as qr   (the "cheap pivoting" of Drmac-Veselic)
  qrcp        column-pivoted QR, Jacobi on L = R^H
  qrcp+qr     second (unpivoted) QR of L, Jacobi on the columns of R2^H
  qrcp-trunc  QRCP stopped at the first pivot below 1e-5*eps_rel*||X||_F; Jacobi on the k
              columns of [R11 R12]^H; ||R22||_F^2 joins the tail norm of the rank rule;  +qr as above
  grampiv+qr  column order from a pivoted Cholesky of the fp64 Gram matrix (no pivoting inside
              the QR: a blocked Householder QR, GEMM-rich), Jacobi on L = R^H;  +qr as above

  python tools/study_precond.py [step=25] [max_ops=6] [only_operand=k]   (STUDY_NPASS: inner sweeps;
  STUDY_VARIANTS=a,b | pipeline | qrcp-check | kdist (max_ops=0: every SVD of the step))
"""
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, ".")

B = 16
TOL = 1e-11
EPS = np.finfo(float).eps
EPSREL = 1e-9


def keep_rule(s, extra_tail2=0.0):
    """The reference's rank rule (oracle/tempo_np.py truncated_svd), with the Frobenius mass
    of columns that were dropped before the iteration added to the tail."""
    s = np.sort(np.asarray(s))[::-1]
    tail = np.sqrt(np.cumsum(np.concatenate(([extra_tail2], s[::-1] ** 2)))[1:])
    return int(np.count_nonzero(tail > EPSREL * s[0]))


def _inner_lib():
    so = "/tmp/oqupy_b200_jacobi_inner_emu.so"
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro", "jacobi_inner_emu.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    lib = ctypes.CDLL(so)
    lib.inner_sweeps.restype = ctypes.c_int
    lib.inner_sweeps.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_double, ctypes.c_double, ctypes.c_int]
    return lib


LIB = _inner_lib()
NPASS = int(os.environ.get("STUDY_NPASS", "1"))


def round_robin(nb):
    """Rounds of disjoint block pairs (circle method); nb padded to even."""
    n = nb + (nb & 1)
    ring = list(range(n))
    for _ in range(n - 1):
        yield [(min(ring[i], ring[n - 1 - i]), max(ring[i], ring[n - 1 - i]))
               for i in range(n // 2) if ring[i] < nb and ring[n - 1 - i] < nb]
        ring = [ring[0]] + [ring[-1]] + ring[1:-1]


def block_jacobi(x, max_sweeps=60, accumulate=False):
    """One-sided block Jacobi on the columns of x (copy).  Returns (sweeps, singular values,
    rotated stage-slots, tested stage-slots); with accumulate=True the identity is stacked
    under x like the kernel's [X ; W] and (sweeps, rotated columns, accumulated J) come back."""
    x = np.array(x, dtype=complex, order="F")
    n = x.shape[1]
    nb = (n + B - 1) // B
    floor2 = (8 * EPS * np.linalg.norm(x)) ** 2
    mg = x.shape[0]
    if accumulate:
        x = np.asfortranarray(np.vstack([x, np.eye(n, dtype=complex)]))
    rotated = tested = 0
    for sweep in range(1, max_sweeps + 1):
        dirty = False
        for pairs in round_robin(nb):
            for a, b in pairs:
                idx = np.r_[a * B:min((a + 1) * B, n), b * B:min((b + 1) * B, n)]
                t = x[:, idx]
                g = np.ascontiguousarray(t[:mg].conj().T @ t[:mg])
                j = np.eye(len(idx), dtype=complex)
                tested += 1
                if LIB.inner_sweeps(len(idx), g.ctypes.data, j.ctypes.data,
                                    TOL ** 2, floor2, NPASS) == 0:
                    continue
                dirty = True
                rotated += 1
                order = np.argsort(-np.real(np.diag(g)), kind="stable")
                x[:, idx] = t @ j[:, order]
        if not dirty:
            break
    if accumulate:
        return sweep, x[:mg], x[mg:]
    s = np.sort(np.linalg.norm(x, axis=0))[::-1]
    return sweep, s, rotated, tested


def pipeline(theta):
    """The whole proposed factorisation in numpy, next to the
    oracle's truncated SVD: stopped QRCP -> QR of [R11 R12]^H -> block Jacobi on R2^H ->
    U = Q1[:, :k] W_hat,  S.Vh = (Q2 J Sigma)^H P^T.  Returns the figures a parity test of the
    round-2 kernel will assert."""
    from oracle import tempo_np
    flip = theta.shape[0] < theta.shape[1]
    a = theta.conj().T if flip else theta
    q1, r, piv = sla.qr(a, mode="economic", pivoting=True)
    tau = 1e-5 * EPSREL * np.linalg.norm(a)
    k = int(np.count_nonzero(np.abs(np.diag(r)) > tau))
    tail2 = float(np.linalg.norm(r[k:, k:]) ** 2)
    q2, r2 = sla.qr(r[:k, :].conj().T, mode="economic")
    sweeps, w, j = block_jacobi(r2.conj().T, accumulate=True)
    sig = np.linalg.norm(w, axis=0)
    order = np.argsort(-sig, kind="stable")
    sig, w, j = sig[order], w[:, order], j[:, order]
    keep = keep_rule(sig, tail2)
    u = q1[:, :k] @ (w[:, :keep] / sig[:keep])
    svh_p = ((q2 @ j[:, :keep]) * sig[:keep]).conj().T
    svh = np.empty_like(svh_p)
    svh[:, piv] = svh_p
    if flip:                                       # Theta = (U S Vh)^H = V (S U^H)
        u, svh = svh.conj().T / sig[:keep], (u * sig[:keep]).conj().T
    u_ref, s_lap, vh_lap = tempo_np.truncated_svd(theta, EPSREL)[:3]
    svh_ref = s_lap[:, None] * vh_lap
    s0 = sig[0]
    return {"columns": k, "sweeps": sweeps, "keep": keep, "keep_lapack": int(u_ref.shape[1]),
            "recon_vs_theta_over_s0": float(np.linalg.norm(u @ svh - theta, 2) / s0),
            "lapack_recon_vs_theta_over_s0":
                float(np.linalg.norm(u_ref @ svh_ref - theta, 2) / s0),
            "product_vs_lapack_over_s0": float(np.linalg.norm(u @ svh - u_ref @ svh_ref, 2) / s0)
            if keep == u_ref.shape[1] else None,
            "u_orthogonality": float(np.linalg.norm(u.conj().T @ u - np.eye(keep), 2))}


def variants(theta):
    m, n = theta.shape
    if m < n:                                     # the kernel factors Theta^H then
        theta = theta.conj().T
    yield "plain", theta
    order = np.argsort(-np.linalg.norm(theta, axis=0))
    yield "sort", theta[:, order]
    yield "qr", sla.qr(theta, mode="economic")[1].conj().T
    yield "sort+qr", sla.qr(theta[:, order], mode="economic")[1].conj().T
    r = sla.qr(theta, mode="economic", pivoting=True)[1]
    yield "qrcp", r.conj().T
    yield "qrcp+qr", sla.qr(r.conj().T, mode="economic")[1].conj().T
    # early-terminated QRCP: stop at the first pivot below 1e-5*eps_rel*||X||_F (the kernel's
    # deflation level); the trailing block only enters the tail norm
    tau = 1e-5 * EPSREL * np.linalg.norm(theta)
    k = int(np.count_nonzero(np.abs(np.diag(r)) > tau))
    tail2 = float(np.linalg.norm(r[k:, k:]) ** 2)
    yield "qrcp-trunc", (r[:k, :].conj().T, tail2)
    yield "qrcp-trunc+qr", (sla.qr(r[:k, :].conj().T, mode="economic")[1].conj().T, tail2)
    lg = sla.qr(theta[:, gram_pivot_order(theta)], mode="economic")[1].conj().T
    yield "grampiv+qr", lg
    yield "grampiv+qr+qr", sla.qr(lg, mode="economic")[1].conj().T


def gram_pivot_order(theta):
    """Pivot order of a diagonally pivoted Cholesky factorisation of the fp64 Gram matrix: the
    QRCP order in exact arithmetic, reliable over the leading ~8 decades only."""
    g = theta.conj().T @ theta
    n = g.shape[0]
    perm = np.arange(n)
    d = np.real(np.diag(g)).copy()
    l = np.zeros((n, n), dtype=complex)
    for k in range(n):
        p = k + int(np.argmax(d[perm[k:]]))
        perm[[k, p]] = perm[[p, k]]
        pk = perm[k]
        if d[pk] <= 0.0:
            break
        rest = perm[k + 1:]
        l[pk, k] = np.sqrt(d[pk])
        l[rest, k] = (g[rest, pk] - l[rest, :k] @ l[pk, :k].conj()) / l[pk, k]
        d[rest] -= np.abs(l[rest, k]) ** 2
    return perm


def qrcp_stopped(a, stop):
    """numpy statement, step by step, of oqupy_b200/csrc/draft/qrcp.cu (physical columns stay
    in place, perm[] holds the pivot order, LAPACK zlarfg reflectors with real beta, xGEQP3
    norm down-dating, stop at the first pivot norm <= stop).  Returns (a_out, tau, perm, k,
    tail2) in the kernel's output layout."""
    a = np.array(a, dtype=complex, order="F")
    m, n = a.shape
    vn = np.linalg.norm(a, axis=0)
    vn_ref = vn.copy()
    done = np.zeros(n, dtype=bool)
    tau = np.zeros(n, dtype=complex)
    perm = []
    k = 0
    while k < min(m, n):
        cand = np.where(done, -1.0, vn)
        p = int(np.argmax(cand))                   # first maximum = lowest physical index
        if not cand[p] > stop:
            break
        done[p] = True
        perm.append(p)
        alpha = a[k, p]
        xnorm2 = float(np.sum(np.abs(a[k + 1:, p]) ** 2))
        v = np.zeros(m - k, dtype=complex)
        v[0] = 1.0
        if xnorm2 > 0.0 or alpha.imag != 0.0:
            an = np.sqrt(abs(alpha) ** 2 + xnorm2)
            beta = -an if alpha.real >= 0.0 else an
            tau[k] = complex((beta - alpha.real) / beta, -alpha.imag / beta)
            v[1:] = a[k + 1:, p] / (alpha - beta)
            a[k, p] = beta
        a[k + 1:, p] = v[1:]
        for c in np.nonzero(~done)[0]:
            w = np.vdot(v, a[k:, c])
            a[k:, c] -= np.conj(tau[k]) * w * v
            if vn[c] > 0.0:
                r = abs(a[k, c]) / vn[c]
                t = max(0.0, (1.0 + r) * (1.0 - r))
                if t * (vn[c] / vn_ref[c]) ** 2 <= np.sqrt(EPS):
                    vn[c] = vn_ref[c] = np.linalg.norm(a[k + 1:, c])
                else:
                    vn[c] *= np.sqrt(t)
        k += 1
    rest = np.nonzero(~done)[0]
    tail2 = float(np.sum(np.abs(a[k:, rest]) ** 2))
    return a, tau, np.array(perm + list(rest)), k, tail2


def check_qrcp_stopped(theta):
    """qrcp_stopped against scipy's QRCP on one operand: |diag R|, the pivot count at the stop
    level, ||A P - Q R|| with Q rebuilt from the stored reflectors the way apply_q_kernel does."""
    a0 = theta.conj().T if theta.shape[0] < theta.shape[1] else theta
    m, n = a0.shape
    stop = 1e-5 * EPSREL * np.linalg.norm(a0)
    a, tau, perm, k, tail2 = qrcp_stopped(a0, stop)
    r_ref = sla.qr(a0, mode="r", pivoting=True)[0]
    k_ref = int(np.count_nonzero(np.abs(np.diag(r_ref)) > stop))
    r = np.zeros((k, n), dtype=complex)
    for pos in range(n):
        top = min(pos + 1, k)
        r[:top, pos] = a[:top, perm[pos]]
    y = np.vstack([r, np.zeros((m - k, n))])       # Q [R; 0], reflectors k-1 ... 0
    for i in range(k - 1, -1, -1):
        v = np.concatenate(([1.0], a[i + 1:, perm[i]]))
        y[i:] -= tau[i] * np.outer(v, v.conj() @ y[i:])
    resid = np.linalg.norm(y - a0[:, perm], 2) / np.linalg.norm(a0, 2)
    return {"k": k, "k_scipy": k_ref, "tail2": tail2,
            "tail2_scipy": float(np.linalg.norm(r_ref[k_ref:, k_ref:]) ** 2),
            "diagR_rel_dev": float(np.max(np.abs(np.abs(np.diag(r)[:k])
                                                 - np.abs(np.diag(r_ref)[:k]))
                                          / np.abs(np.diag(r_ref)[:k]))),
            "resid_AP_minus_QR_over_norm": float(resid)}


def capture(step, max_ops):
    from oracle import tempo_np
    with np.load("tests/golden/c2_operands.npz") as f:
        infl = f["influences"]
    orc = tempo_np.PtTempoOracle(2, lambda dk: None if dk < 0 else infl[dk], 1000, 200, 1e-9)
    orc.initialize()
    while orc.step < step - 1:
        orc.compute_step()
    mats = []
    inner = tempo_np.truncated_svd

    def spy(mat, eps):
        mats.append(np.array(mat))
        return inner(mat, eps)
    tempo_np.truncated_svd = spy
    try:
        orc.compute_step()
    finally:
        tempo_np.truncated_svd = inner
    # the widest operands, one from the middle of the zip-up chain, one sweep operand
    if max_ops <= 0:
        return len(mats), list(enumerate(mats))
    by_size = sorted(range(len(mats)), key=lambda i: -min(mats[i].shape))
    pick = by_size[:max(1, max_ops - 2)]
    nz = len(mats) // 2                            # zip-up SVDs come first, then the sweep
    pick += [nz // 2, nz + nz // 2]
    seen, out = set(), []
    for i in pick[:max_ops]:
        if i not in seen:
            seen.add(i)
            out.append((i, mats[i]))
    return len(mats), out


def main():
    step = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    max_ops = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    nsvd, ops = capture(step, max_ops)
    if len(sys.argv) > 3:
        ops = [ops[int(sys.argv[3])]]
    if os.environ.get("STUDY_VARIANTS") == "kdist":
        # every SVD of the step: columns n, pivots k above the stop level, kept rank
        rows = []
        for i, theta in ops:
            a = theta.conj().T if theta.shape[0] < theta.shape[1] else theta
            r = sla.qr(a, mode="r", pivoting=True)[0]
            d = np.abs(np.diag(r))
            k = int(np.count_nonzero(d > 1e-5 * EPSREL * np.linalg.norm(a)))
            keep = keep_rule(np.linalg.svd(theta, compute_uv=False))
            rows.append((i, a.shape[0], a.shape[1], k, keep))
        arr = np.array(rows, dtype=float)
        nb, kb = np.ceil(arr[:, 2] / B), np.ceil(arr[:, 3] / B)
        print(json.dumps({
            "step": step, "svds": len(rows),
            "sum_columns": int(arr[:, 2].sum()), "sum_pivots": int(arr[:, 3].sum()),
            "sum_keep": int(arr[:, 4].sum()),
            "block_pairs_per_sweep_today": int((nb * (nb - 1) / 2).sum()),
            "block_pairs_per_sweep_stopped": int((kb * (kb - 1) / 2).sum()),
            "tournament_rounds_per_sweep_today": int(np.maximum(nb - 1, 0).sum()),
            "tournament_rounds_per_sweep_stopped": int(np.maximum(kb - 1, 0).sum()),
            "per_svd_m_n_k_keep": [list(map(int, r[1:])) for r in rows]}), flush=True)
        return
    print(f"step {step}: {nsvd} SVDs, studying {len(ops)}", file=sys.stderr, flush=True)
    for i, theta in ops:
        s_ref = np.linalg.svd(theta, compute_uv=False)
        row = {"step": step, "svd_index": i, "shape": list(theta.shape),
               "decades": float(np.log10(s_ref[0] / max(s_ref[-1], 1e-300))), "variants": {}}
        only = os.environ.get("STUDY_VARIANTS")
        if only == "pipeline":
            row["pipeline"] = pipeline(theta)
            print(f"  {theta.shape} pipeline: {row['pipeline']}", file=sys.stderr, flush=True)
            print(json.dumps(row), flush=True)
            continue
        if only == "qrcp-check":
            row["qrcp_stopped_vs_scipy"] = check_qrcp_stopped(theta)
            print(json.dumps(row), flush=True)
            continue
        for name, x in variants(theta):
            if only and name not in only.split(","):
                continue
            t0 = time.perf_counter()
            tail2 = 0.0
            if isinstance(x, tuple):
                x, tail2 = x
            sweeps, s, rot, tst = block_jacobi(x)
            keep_ref = keep_rule(s_ref)
            keep = keep_rule(s, tail2)
            kk = min(keep, keep_ref)
            row["variants"][name] = {
                "columns": int(x.shape[1]),
                "sweeps": sweeps, "rotating_slots": rot, "tested_slots": tst,
                "max_ds_over_s0": float(np.max(np.abs(s[:kk] - s_ref[:kk])) / s_ref[0]),
                "max_rel_ds_kept": float(np.max(np.abs(s[:kk] - s_ref[:kk]) / s_ref[:kk])),
                "keep": keep, "keep_lapack": keep_ref,
                "emu_s": round(time.perf_counter() - t0, 2)}
            print(f"  {theta.shape} {name}: {row['variants'][name]}", file=sys.stderr, flush=True)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
