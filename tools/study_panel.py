"""CPU study: how much does PANEL pivoting (one grid hand-shake selects up to b pivots: the
best remaining column of b different CTAs, ordered greedily inside the panel) cost in Jacobi
sweeps compared with strict column pivoting?  Operands: the config-2 build, captured from the
oracle.  Emulation of the Jacobi stage: tools/study_precond.py::block_jacobi.

  python tools/study_panel.py [step=25] [G=148] [b list=1,4,8,16]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
sys.path.insert(0, "tests")
import study_precond as sp  # noqa: E402
import qr_model  # noqa: E402

EPSREL = 1e-9


def panel_qrcp(x, stop_rel, grid, b, theta=0.0, group=1):
    """Stopped Householder QR with panel pivoting.  Column c belongs to CTA c % grid; per
    hand-shake every CTA offers its best remaining column, the b largest offers form the
    panel, inside the panel pivots are taken greedily by their exact remaining norms while
    these stay above max(stop, theta * best offer outside the panel)."""
    a = np.array(x, dtype=complex, order="F")
    p, q = a.shape
    done = np.zeros(q, dtype=bool)
    vn2 = np.sum(np.abs(a) ** 2, axis=0)
    stop2 = stop_rel ** 2 * float(np.sum(vn2))
    owner = (np.arange(q) // group) % grid
    perm, taus = [], []
    j = 0
    shakes = 0
    while j < q:
        shakes += 1
        offers = []
        for g in range(min(grid, q)):
            mine = np.where((owner == g) & ~done)[0]
            if len(mine):
                c = mine[np.argmax(vn2[mine])]
                offers.append((vn2[c], -c))
        offers.sort(reverse=True)
        if not offers or not offers[0][0] > stop2:
            break
        panel = [-c for v, c in offers[:b] if v > stop2]
        outside = offers[b][0] if len(offers) > b else 0.0
        thr = max(stop2, theta * theta * outside)
        taken = 0
        while panel and j < q:
            rem = [float(np.sum(np.abs(a[j:, c]) ** 2)) for c in panel]
            t = int(np.argmax(rem))
            if taken and not rem[t] > thr:
                break
            widx = panel.pop(t)
            col = a[:, widx].copy()
            alpha = col[j]
            xnorm2 = float(np.sum(np.abs(col[j + 1:]) ** 2))
            beta, tau, scale = alpha.real, 0.0, 0.0
            if xnorm2 > 0.0 or alpha.imag != 0.0:
                an = np.sqrt(abs(alpha) ** 2 + xnorm2)
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
    return a, np.array(perm + list(rest), dtype=int), np.array(taus), j, tail2, shakes


def main():
    step = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    grid = int(sys.argv[2]) if len(sys.argv) > 2 else 148
    bs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,4,8,16").split(",")]
    thetas = [float(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "0").split(",")]
    group = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    nsvd, ops = sp.capture(step, 4)
    for i, theta_mat in ops:
        x = theta_mat.conj().T if theta_mat.shape[0] < theta_mat.shape[1] else theta_mat
        if x.shape[0] > 2 * x.shape[1]:
            continue
        s_ref = np.linalg.svd(theta_mat, compute_uv=False)
        keep_ref = sp.keep_rule(s_ref)
        g = min(grid, (x.shape[1] + 3) // 4)
        for b in bs:
            for th in thetas:
                t0 = time.perf_counter()
                a, perm, tau, k, tail2, shakes = panel_qrcp(x, 1e-5 * EPSREL, g, b, th, group)
                lmat = qr_model.l_operand(a, perm, k)
                sweeps, s, rot, tst = sp.block_jacobi(lmat)
                keep = sp.keep_rule(s, tail2)
                d = np.abs(np.array([a[t, perm[t]] for t in range(k)]))
                print(json.dumps({
                    "step": step, "svd": i, "shape": list(theta_mat.shape), "grid": g, "b": b,
                    "theta": th, "group": group, "k": k, "handshakes": shakes, "sweeps": sweeps,
                    "rotating_slots": rot, "tested_slots": tst, "keep": keep,
                    "keep_lapack": keep_ref,
                    "diag_inversions": int(np.count_nonzero(d[1:] > d[:-1] * 1.0000001)),
                    "s": round(time.perf_counter() - t0, 1)}), flush=True)


if __name__ == "__main__":
    main()
