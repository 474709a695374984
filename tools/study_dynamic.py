"""CPU study (not a GPU measurement): Gram-guided dynamic ordering of the block one-sided Jacobi
iteration behind the stopped QRCP (svd.cu / qrcp.cuh), on operands captured from the config-2
build (tools/study_precond.py::capture).

The kernel's tournament visits every block pair once per sweep; after the first two sweeps
most stages only TEST (Gram matrix + hand-shakes, no rotation).  Here: after S0 full sweeps the
full Gram matrix G = L^H L decides which block pairs still hold a violating column pair; only
those are scheduled (greedy matching by violation weight -> rounds of disjoint pairs), G is
refreshed after every round, the iteration ends when G shows no violation.

  python tools/study_dynamic.py [step=25] [S0=2] [operand indices ...]
"""
import json
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import study_precond as sp  # noqa: E402

B = sp.B
TOL2 = sp.TOL ** 2


def viol_matrix(g, floor2):
    d = np.real(np.diag(g))
    big = np.maximum.outer(d, d)
    small = np.minimum.outer(d, d)
    v = (np.abs(g) ** 2) > big * (TOL2 * small + floor2)
    np.fill_diagonal(v, False)
    return v


def block_flags(v, nb, n):
    """w[a, b] = number of violating column pairs between blocks a and b (a == b: inside a)."""
    w = np.zeros((nb, nb), dtype=int)
    for a in range(nb):
        for b in range(a, nb):
            w[a, b] = w[b, a] = int(v[a * B:min((a + 1) * B, n), b * B:min((b + 1) * B, n)].sum())
    return w


def stage(x, mg, a, b, n, floor2):
    idx = np.r_[a * B:min((a + 1) * B, n), b * B:min((b + 1) * B, n)]
    t = x[:, idx]
    g = np.ascontiguousarray(t[:mg].conj().T @ t[:mg])
    j = np.eye(len(idx), dtype=complex)
    if sp.LIB.inner_sweeps(len(idx), g.ctypes.data, j.ctypes.data, TOL2, floor2, 1) == 0:
        return False
    order = np.argsort(-np.real(np.diag(g)), kind="stable")
    x[:, idx] = t @ j[:, order]
    return True


def run(x, s0, max_rounds=4000):
    x = np.array(x, dtype=complex, order="F")
    mg, n = x.shape
    nb = (n + B - 1) // B
    floor2 = (8 * sp.EPS * np.linalg.norm(x)) ** 2
    log = {"nb": nb, "rounds_per_sweep": nb - 1 + (nb & 1) - 0, "sweeps": []}
    # --- baseline statistics: per sweep rotated / tested (on a copy), full tournament
    xb = x.copy()
    total_rounds = 0
    for sweep in range(1, 61):
        rot = tst = 0
        rounds_rot = 0
        for pairs in sp.round_robin(nb):
            r = 0
            for a, b in pairs:
                tst += 1
                r += stage(xb, mg, a, b, n, floor2)
            rot += r
            rounds_rot += (r > 0)
            total_rounds += 1
        log["sweeps"].append({"sweep": sweep, "rotated": rot, "tested": tst,
                              "rounds_with_rotation": rounds_rot})
        if rot == 0:
            break
    log["baseline_rounds"] = total_rounds
    log["baseline_sigma"] = np.sort(np.linalg.norm(xb, axis=0))[::-1]
    # --- dynamic: S0 full sweeps, then Gram-guided rounds
    rounds = 0
    for sweep in range(s0):
        for pairs in sp.round_robin(nb):
            for a, b in pairs:
                stage(x, mg, a, b, n, floor2)
            rounds += 1
    dyn_rounds = []
    while rounds < max_rounds:
        g = x.conj().T @ x
        w = block_flags(viol_matrix(g, floor2), nb, n)
        if not w.any():
            break
        cross = [(w[a, b], a, b) for a in range(nb) for b in range(a + 1, nb) if w[a, b]]
        cross.sort(reverse=True)
        used, pairs = set(), []
        for _, a, b in cross:
            if a not in used and b not in used:
                used.update((a, b))
                pairs.append((a, b))
        # blocks with an internal violation that found no partner: pair them with a free block
        for a in range(nb):
            if w[a, a] and a not in used:
                free = [b for b in range(nb) if b != a and b not in used]
                if free:
                    b = free[0]
                    used.update((a, b))
                    pairs.append((min(a, b), max(a, b)))
        nrot = sum(stage(x, mg, a, b, n, floor2) for a, b in pairs)
        dyn_rounds.append((len(pairs), int(nrot), int((w > 0).sum() // 2)))
        rounds += 1
    log["dynamic_rounds_total"] = rounds
    log["dynamic_guided_rounds"] = len(dyn_rounds)
    log["dynamic_rounds_detail_first20"] = dyn_rounds[:20]
    log["dynamic_sigma"] = np.sort(np.linalg.norm(x, axis=0))[::-1]
    return log


def main():
    step = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    want = [int(a) for a in sys.argv[3:]] or [100, 185, 199]
    nsvd, ops = sp.capture(step, 0)
    for i, theta in ops:
        if i not in want:
            continue
        a = theta.conj().T if theta.shape[0] < theta.shape[1] else theta
        r = sla.qr(a, mode="r", pivoting=True)[0]
        tau = 1e-5 * sp.EPSREL * np.linalg.norm(a)
        k = int(np.count_nonzero(np.abs(np.diag(r)) > tau))
        tail2 = float(np.linalg.norm(r[k:, k:]) ** 2)
        lmat = r[:k, :].conj().T
        log = run(lmat, s0)
        s_ref = np.linalg.svd(theta, compute_uv=False)
        keep_ref = sp.keep_rule(s_ref)
        sb, sd = log.pop("baseline_sigma"), log.pop("dynamic_sigma")
        log.update({"svd_index": i, "shape": list(theta.shape), "k": k, "S0": s0,
                    "keep_lapack": keep_ref, "keep_baseline": sp.keep_rule(sb, tail2),
                    "keep_dynamic": sp.keep_rule(sd, tail2),
                    "max_rel_ds_kept_dynamic":
                        float(np.max(np.abs(sd[:keep_ref] - s_ref[:keep_ref]) / s_ref[:keep_ref])),
                    "max_rel_ds_kept_baseline":
                        float(np.max(np.abs(sb[:keep_ref] - s_ref[:keep_ref]) / s_ref[:keep_ref]))})
        print(json.dumps(log), flush=True)


if __name__ == "__main__":
    main()
