"""GPU diagnostics for the truncated-SVD kernel: accuracy + timing per size."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oqupy_b200._lib import default_ops  # noqa: E402


def graded(rng, m, n, lo):
    k = min(m, n)
    s = np.sort(10.0 ** rng.uniform(lo, 0.0, size=k))[::-1]
    s[0] = 1.0
    q1, _ = np.linalg.qr(rng.normal(size=(m, k)) + 1j * rng.normal(size=(m, k)))
    q2, _ = np.linalg.qr(rng.normal(size=(n, k)) + 1j * rng.normal(size=(n, k)))
    return (q1 * s) @ q2.conj().T


def main():
    ops = default_ops()
    rng = np.random.default_rng(0)
    sizes = [(8, 8), (64, 64), (128, 128), (256, 256), (512, 512), (1024, 256),
             (768, 768), (1152, 1088), (1600, 1464)]
    if len(sys.argv) > 1:
        sizes = sizes[:int(sys.argv[1])]
    eps = 1e-9
    print("m n | keep ref | sweeps | ms | max|ds|/s0 | UhU-I | recon")
    for m, n in sizes:
        a = graded(rng, m, n, -16.0)
        da = ops.from_host(a)
        h = ops.svd_factor(da, m, n, n, 1, eps)       # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h = ops.svd_factor(da, m, n, n, 1, eps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        s = ops.svd_values(h)
        pc = ops.svd_phase_cycles(h)
        st = max(pc[7], 1)
        print("   cycles/stage: wait %d loadX %d gram %d eig %d apply %d store %d | vote/sweep %d | stages %d"
              % (pc[0] // st, pc[1] // st, pc[2] // st, pc[3] // st, pc[4] // st, pc[5] // st,
                 pc[6] // max(h.sweeps, 1), pc[7]))
        t0 = time.perf_counter()
        ur, sr, vhr = np.linalg.svd(a, full_matrices=False)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        tail = np.sqrt(np.cumsum(sr[::-1] ** 2))
        kref = int(np.count_nonzero(tail > eps * sr[0]))
        k = h.keep
        u, svh = ops.empty(m, k), ops.empty(k, n)
        ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sj=1, svh=svh)
        u, svh = ops.to_host(u), ops.to_host(svh)
        orth = np.abs(u.conj().T @ u - np.eye(k)).max()
        recon = np.abs(u @ svh - (ur[:, :k] * sr[:k]) @ vhr[:k]).max()
        print(f"{m} {n} | {k} {kref} | {h.sweeps} | {ms:.2f} (cpu {cpu_ms:.1f}) | "
              f"{np.abs(s - sr).max():.1e} | {orth:.1e} | {recon:.1e}", flush=True)


if __name__ == "__main__":
    main()
