"""GPU: per-step wall time and per-kernel-family device time of the config-2 PT-TEMPO build.

  python tools/step_profile.py [last_step=50] [budget_s=200]
One JSON line per step: step, wall ms, kernel ms by family (jacobi / qrcp / emit), launches,
largest SVD, max bond; every 10th step the shape histogram of the step's SVDs.
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402


def main():
    last = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 200.0
    with np.load("tests/golden/c2_operands.npz") as f:
        infl = f["influences"]
    ops = ob.default_ops()
    be = ob.PtTempoBackend(2, lambda dk: None if dk < 0 else infl[dk], None,
                           np.ones(4), np.ones(4), 1000, 200, 1e-9)
    t0 = time.perf_counter()
    be.initialize()
    torch.cuda.synchronize()
    print(json.dumps({"initialize_s": round(time.perf_counter() - t0, 3)}), flush=True)
    be.pop_svd_log()
    ops.profile_enable(True)
    ops.profile_read_kinds()
    t_start = time.perf_counter()
    while be.step < last and time.perf_counter() - t_start < budget:
        l0 = ops.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        be.compute_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        kinds, flops, sweeps = ops.profile_read_kinds()
        log = be.pop_svd_log()
        big = max(log, key=lambda x: x[0] * x[1])
        row = {"step": be.step, "ms": round(dt * 1e3, 1),
               "kernels_ms": {k: round(v[0], 1) for k, v in kinds.items() if v[1]},
               "launches": ops.launch_count() - l0, "sweeps": int(sweeps),
               "largest": list(big), "max_bond": max(be.get_bond_dimensions())}
        if be.step % 10 == 0:
            row["svd_log"] = [list(x) for x in log]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
