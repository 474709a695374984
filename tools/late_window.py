"""GPU: supplementary measurement for DESIGN.md -- the config-2 PT-TEMPO build further into
the grow phase than bench.py's window (bond dimensions several hundred), next to the time
numpy/LAPACK (zgesdd, all host threads) needs for the truncated SVDs of the SAME operand
shapes.  The SVD is >= 90 % of the reference's step, so the LAPACK figure is a
LOWER bound of the reference's step time.

  python tools/late_window.py [last_step=60] [gpu_budget_s=200]
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402


def main():
    last = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 200.0
    with np.load("tests/golden/c2_operands.npz") as f:
        infl = f["influences"]
    be = ob.PtTempoBackend(2, lambda dk: None if dk < 0 else infl[dk], None,
                           np.ones(4), np.ones(4), 1000, 200, 1e-9)
    be.initialize()
    be.pop_svd_log()
    t_start = time.perf_counter()
    rows = []
    while be.step < last and time.perf_counter() - t_start < budget:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        be.compute_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        log = be.pop_svd_log()
        rows.append((be.step, dt, log))
        if be.step % 5 == 0:
            big = max(log, key=lambda x: x[0] * x[1])
            print(f"step {be.step}: {dt:.3f} s, biggest svd {big}, max bond "
                  f"{max(be.get_bond_dimensions())}", file=sys.stderr, flush=True)
    step, gpu_s, log = rows[-1]
    # LAPACK on the same shapes (bucketed to multiples of 8 to bound the number of timings)
    rng = np.random.default_rng(0)
    buckets = {}
    for m, n, keep, sweeps in log:
        key = ((m + 7) // 8 * 8, (n + 7) // 8 * 8)
        buckets[key] = buckets.get(key, 0) + 1
    cpu_s = 0.0
    for (m, n), cnt in sorted(buckets.items()):
        a = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        t0 = time.perf_counter()
        np.linalg.svd(a, full_matrices=False)
        cpu_s += (time.perf_counter() - t0) * cnt
    import os
    print(json.dumps({
        "row": "PT-TEMPO config 2, late grow-phase step (supplementary to bench.py)",
        "step": step, "gpu_step_s": gpu_s, "gpu_steps_per_s": 1.0 / gpu_s,
        "lapack_svd_only_step_s": cpu_s, "lapack_svd_only_steps_per_s": 1.0 / cpu_s,
        "ratio_lower_bound": cpu_s / gpu_s, "host_threads": os.cpu_count(),
        "svds": len(log), "largest_svd": max(log, key=lambda x: x[0] * x[1]),
        "max_bond": int(max(be.get_bond_dimensions())),
        "note": "LAPACK zgesdd on random matrices of the step's operand shapes (multiples of 8), "
                "all host threads; a lower bound of the reference's step time (SVD only)"}))


if __name__ == "__main__":
    main()
