"""GPU: host overhead per SVD of the native chain (tiny sites, repeated sweeps)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402
from oqupy_b200._lib import NativeChain  # noqa: E402

ops = ob.default_ops()
rng = np.random.default_rng(0)
for chi in (4, 16, 40):
    ch = NativeChain(ops)
    dims = [1] + [chi] * 40 + [1]
    for k in range(41):
        ch.push(ops.from_host(rng.normal(size=(dims[k], 4, dims[k + 1]))
                              + 1j * rng.normal(size=(dims[k], 4, dims[k + 1]))))
    ch.svd_sweep(0, -1, 1e-14)
    ch.svd_sweep(-1, 0, 1e-14)
    torch.cuda.synchronize()
    ops.profile_enable(True)
    ops.profile_read()
    t0 = time.perf_counter()
    reps = 6
    for _ in range(reps):
        ch.svd_sweep(0, -1, 1e-14)
        ch.svd_sweep(-1, 0, 1e-14)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    kms, _, n, sw = ops.profile_read()
    ops.profile_enable(False)
    print(f"chi={chi}: {n} SVDs, wall {dt / n * 1e6:.1f} us per SVD, jacobi kernel "
          f"{kms / n * 1e3:.1f} us per SVD, outside the kernel {(dt - kms / 1e3) / n * 1e6:.1f} us, "
          f"sweeps/SVD {sw / n:.1f}")
