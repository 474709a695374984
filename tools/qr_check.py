"""GPU: stage-by-stage check and timing of the rank-revealing QR path of the truncated SVD
against its numpy model (tests/qr_model.py) and LAPACK, next to the plain Jacobi path.

    python tools/qr_check.py [synthetic] [oracle25] [time]      (default: all)

Prints one JSON line per operand: shapes, k, keep (QR path / plain path / LAPACK), deviations
of the intermediate arrays from the model, reconstruction error, milliseconds per stage.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qr_model  # noqa: E402
from oqupy_b200._lib import default_ops  # noqa: E402

EPS = 1e-9
PANEL = int(os.environ.get("B200_SVD_QR_PANEL", "0"))


def graded(rng, m, n, lo=-25.0):
    k = min(m, n)
    s = np.sort(10.0 ** rng.uniform(lo, 0.0, size=k))[::-1]
    s[0] = 1.0
    q1, _ = np.linalg.qr(rng.normal(size=(m, k)) + 1j * rng.normal(size=(m, k)))
    q2, _ = np.linalg.qr(rng.normal(size=(n, k)) + 1j * rng.normal(size=(n, k)))
    return (q1 * s) @ q2.conj().T


def ref_keep(s, eps):
    tail = np.sqrt(np.cumsum(np.square(s[::-1])))
    return int(np.count_nonzero(tail > eps * s[0]))


def run_one(ops, theta, eps, qr, reps=1):
    m, n = theta.shape
    ops.svd_config("qr", 1 if qr else 0)
    d = ops.from_host(theta)
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        h = ops.svd_factor(d, m, n, n, 1, eps)
        e1.record()
        k = h.keep
        u, svh = ops.empty(m, max(k, 1)), ops.empty(max(k, 1), n)
        ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sj=1, svh=svh)
        e2.record()
        torch.cuda.synchronize()
        t = (e0.elapsed_time(e1), e1.elapsed_time(e2))
        if best is None or sum(t) < sum(best):
            best = t
    return h, ops.to_host(u)[:, :k], ops.to_host(svh)[:k], best


def check(ops, name, theta, eps=EPS, model=True):
    m, n = theta.shape
    row = {"name": name, "shape": [m, n]}
    ur, sr, vhr = np.linalg.svd(theta, full_matrices=False)
    kref = ref_keep(sr, eps)
    row["keep_lapack"] = kref
    # threshold margin of the reference rule at its cut: |tail/(eps s0) - 1|
    tail = np.sqrt(np.cumsum(np.square(sr[::-1])))[::-1]
    edge = [abs(tail[j] / (eps * sr[0]) - 1.0) for j in (kref - 1, kref) if 0 <= j < len(tail)]
    row["tie_margin"] = float(min(edge)) if edge else None
    try:
        ops.profile_enable(True)
        ops.profile_read_kinds()
        h, u, svh, tq = run_one(ops, theta, eps, True, reps=0)
        kinds = ops.profile_read_kinds()[0]
        ops.profile_enable(False)
        row["ms_kinds_qr"] = {k: round(v[0], 3) for k, v in kinds.items() if v[1]}
    except Exception as exc:  # pylint: disable=broad-except
        row["qr_error"] = str(exc)[:300]
        print(json.dumps(row), flush=True)
        return row
    plan = ops.svd_plan(h)
    row.update({"qr_used": plan[0], "k": plan[1], "grid": plan[2], "resident": plan[3],
                "keep_qr": h.keep, "sweeps_qr": h.sweeps,
                "ms_qr_factor": round(tq[0], 3), "ms_qr_emit": round(tq[1], 3)})
    best = (ur[:, :h.keep] * sr[:h.keep]) @ vhr[:h.keep]
    if h.keep > 0:
        row["recon_qr"] = float(np.abs(u @ svh - best).max() / sr[0])
        row["orth_qr"] = float(np.abs(u.conj().T @ u - np.eye(h.keep)).max())
    if plan[0] and os.environ.get("B200_SVD_PHASES"):
        row["qr_phase_kcyc"] = [round(c / 1e3, 1) for c in ops.svd_qr_phase_cycles(h)]
        row["jac_phase_kcyc"] = [round(c / 1e3, 1) for c in ops.svd_phase_cycles(h)]
    if plan[0] and model:
        dbg = ops.svd_qr_debug(h)
        x = theta.conj().T if m < n else theta
        a_m, perm_m, tau_m, k_m, tail2_m, shakes = qr_model.qrcp_panel(
            x, 1e-5 * eps, dbg["grid"], PANEL if PANEL else (8 if x.shape[0] <= 1024 else 4))
        row["handshakes_model"] = shakes
        row["k_model"] = k_m
        row["tail2"] = [dbg["tail2"], tail2_m]
        kk = min(k_m, dbg["k"])
        same = int(np.count_nonzero(dbg["perm"][:kk] == perm_m[:kk]))
        row["perm_equal_prefix"] = same
        first_diff = next((i for i in range(kk) if dbg["perm"][i] != perm_m[i]), kk)
        row["perm_first_diff"] = first_diff
        f = first_diff
        if f > 0:
            cols = perm_m[:f]
            row["a_dev"] = float(np.abs(dbg["a"][:, cols] - a_m[:, cols]).max()
                                 / np.abs(a_m).max())
            row["tau_dev"] = float(np.abs(dbg["tau"][:f] - tau_m[:f]).max())
        # |diag R| must decay like the model's even if ties reorder pivots
        dg = np.abs(np.array([dbg["a"][i, dbg["perm"][i]] for i in range(dbg["k"])]))
        dm = np.abs(np.array([a_m[i, perm_m[i]] for i in range(k_m)]))
        row["diagR_dev"] = float(np.abs(dg[:kk] - dm[:kk]).max() / dm[0])
        # A P = Q R from the GPU's own arrays
        r = np.zeros((dbg["k"], x.shape[1]), dtype=complex)
        for pos in range(x.shape[1]):
            top = min(pos + 1, dbg["k"])
            r[:top, pos] = dbg["a"][:top, dbg["perm"][pos]]
        qr_ = qr_model.apply_q(dbg["a"], dbg["perm"], dbg["tau"], dbg["k"], r)
        row["resid_AP_QR"] = float(np.abs(qr_ - x[:, dbg["perm"]]).max() / np.abs(x).max())
    try:
        h2, u2, svh2, tp = run_one(ops, theta, eps, False)
        row.update({"keep_plain": h2.keep, "sweeps_plain": h2.sweeps,
                    "ms_plain_factor": round(tp[0], 3), "ms_plain_emit": round(tp[1], 3)})
    except Exception as exc:  # pylint: disable=broad-except
        row["plain_error"] = str(exc)[:300]
    ops.svd_config("qr", 1)
    print(json.dumps(row), flush=True)
    return row


def oracle_operands(step, picks):
    from oracle import tempo_np
    with np.load(os.path.join(ROOT, "tests", "golden", "c2_operands.npz")) as f:
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
    return [(i, mats[i]) for i in picks if i < len(mats)]


def main():
    what = sys.argv[1:] or ["synthetic", "oracle25"]
    ops = default_ops()
    rng = np.random.default_rng(7)
    if "tiny" in what:
        check(ops, "graded", graded(rng, 64, 60))
        check(ops, "graded", graded(rng, 60, 100))
    if "synthetic" in what:
        for (m, n) in [(64, 60), (60, 64), (100, 96), (208, 200), (200, 208), (260, 180),
                       (520, 500), (968, 864), (864, 968), (1500, 1400)]:
            check(ops, "graded", graded(rng, m, n))
    if "big" in what:
        for (m, n) in [(1756, 1608), (2352, 2160)]:
            check(ops, "graded", graded(rng, m, n), model=False)
    if "oracle25" in what:
        t0 = time.perf_counter()
        sel = oracle_operands(25, [40, 100, 150, 185, 190, 195, 198, 199, 201, 300])
        print(json.dumps({"oracle_capture_s": round(time.perf_counter() - t0, 1)}), flush=True)
        for i, th in sel:
            check(ops, f"step25#{i}", th)


if __name__ == "__main__":
    main()
