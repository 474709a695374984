"""GPU: where does the per-SVD time outside jacobi_kernel go?  (tiny operands)"""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402
from oqupy_b200._lib import View  # noqa: E402

ops = ob.default_ops()
rng = np.random.default_rng(0)
for (m, n) in [(8, 8), (44, 36), (168, 42)]:
    a = ops.from_host(rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n)))
    reps = 300

    def timeit(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    h = ops.svd_factor(a, m, n, n, 1, 1e-9)
    pc = ops.svd_phase_cycles(h)
    kern_us = sum(pc[:10]) / 1965.0
    k = h.keep
    u, svh = ops.empty(m, k), ops.empty(k, n)
    c = ops.empty(k, n)

    def f_factor():
        return ops.svd_factor(a, m, n, n, 1, 1e-9)

    def f_emit():
        hh = ops.svd_factor(a, m, n, n, 1, 1e-9)
        ops.svd_emit(hh, u=u, u_na=1, u_so=k, u_sj=1, svh=svh)

    def f_all():
        hh = ops.svd_factor(a, m, n, n, 1, 1e-9)
        ops.svd_emit(hh, u=u, u_na=1, u_so=k, u_sj=1, svh=svh)
        ops.gemm(k, n, k, View(svh, row=n, col=1), View(svh, row=n, col=1), View(c, row=n, col=1))

    def f_sync():
        torch.cuda.current_stream().synchronize()

    def f_empty():
        ops.empty(m, k)

    print(f"{m}x{n}: kernel (cycle counters) {kern_us:.1f} us | factor+sync {timeit(f_factor):.1f} us | "
          f"+emit {timeit(f_emit):.1f} us | +gemm {timeit(f_all):.1f} us | bare sync {timeit(f_sync):.1f} us | "
          f"torch.empty {timeit(f_empty):.1f} us")
