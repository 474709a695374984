"""GPU: factor captured operands (tools/data/thetas.npz) and print sweeps / phase cycles."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402

ops = ob.default_ops()
Z = np.load("tools/data/thetas.npz")
for key in Z.files:
    a = Z[key]
    m, n = a.shape
    d = ops.from_host(a)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = ops.svd_factor(d, m, n, n, 1, 1e-9)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
    pc = ops.svd_phase_cycles(h)
    s = ops.svd_values(h)
    sref = np.linalg.svd(a, compute_uv=False)
    print(key, a.shape, "keep", h.keep, "sweeps", h.sweeps, "rot stages", h.rotations,
          f"{ms:.3f} ms", "err/s0 %.1e" % (np.abs(s - sref).max() / sref[0]),
          "phase kcyc", [round(c / 1e3) for c in pc[:10]], "stages", pc[15])
