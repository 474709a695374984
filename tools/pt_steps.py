"""GPU: run the first steps of the config-2 PT-TEMPO build and print per-step timing."""
import os
import sys
import time

os.environ.setdefault("OQUPY_B200_PYCHAIN", "1")   # per-SVD instrumentation needs the Python chain

import numpy as np
import torch

sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402


def main():
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    with np.load("tests/golden/c2_operands.npz") as f:
        g = {k: f[k] for k in f.files}
    infl = g["influences"]
    ops = ob.default_ops()
    ops.svd_log = []
    be = ob.PtTempoBackend(2, lambda dk: None if dk < 0 else infl[dk], None,
                           np.ones(4), np.ones(4), 1000, 200, 1e-9)
    t0 = time.perf_counter()
    be.initialize()
    torch.cuda.synchronize()
    print("init", round(time.perf_counter() - t0, 3), "s; svds", len(ops.svd_log))
    for _ in range(nsteps):
        ops.svd_log = []
        l0 = ops.launch_count()
        t0 = time.perf_counter()
        be.compute_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        log = ops.svd_log
        big = max(log, key=lambda x: x[0] * x[1])
        print(f"step {be.step}: {dt:.3f} s, svds {len(log)}, biggest {big}, "
              f"max keep {max(x[2] for x in log)}, mean sweeps "
              f"{np.mean([x[3] for x in log]):.1f}, max sweeps "
              f"{max(x[3] for x in log)}, launches {ops.launch_count() - l0}",
              flush=True)


if __name__ == "__main__":
    main()
