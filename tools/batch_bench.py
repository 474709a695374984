"""GPU: throughput of the lock-step TEMPO ensemble (config-5 shape: K=20, eps=1e-7, d2=4).

  python tools/batch_bench.py [members=592] [steps=60] [warm=25]
Members: coupling scan alpha = 0.08 * f, f in linspace(0.25, 2.0) (element-wise powers of the
config-1 influence matrices).  Prints one JSON line; the CPU figure is the oracle on ONE member
(one host thread) for the same steps.
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
import oqupy_b200 as ob  # noqa: E402
from conftest import load_golden  # noqa: E402
from test_batch_gpu import oracle_run, scaled  # noqa: E402


def main():
    members = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    warm = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    g = load_golden("tempo_c1_k20_eps7_n60")
    base = g["influences"][:21]
    fac = np.linspace(0.25, 2.0, members)
    infl = np.array([scaled(base, f) for f in fac])
    be = ob.BatchedTempoBackend(np.array([g["initial_state"].reshape(-1)] * members), infl,
                                g["unitary"], lambda s: (g["prop_1"], g["prop_2"]),
                                np.ones(4), np.ones(4), 20, 1e-7)
    be.initialize()
    be.compute_steps(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    st = be.compute_steps(steps)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    info = be.info()
    # CPU: one member, one thread
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:  # pylint: disable=broad-except
        pass
    t0 = time.perf_counter()
    ref, _ = oracle_run(infl[members // 2], g, 20, 1e-7, warm + steps)
    cpu = time.perf_counter() - t0
    dev = float(np.abs(st[:, members // 2] - ref[warm + 1:]).max())
    print(json.dumps({
        "row": "config-5 shape: lock-step TEMPO ensemble, K=20 eps=1e-7 d2=4",
        "members": members, "steps": steps, "first_step": warm + 1,
        "member_steps_per_s": members * steps / (e0.elapsed_time(e1) * 1e-3),
        "wall_member_steps_per_s": members * steps / wall,
        "ms_per_ensemble_step": e0.elapsed_time(e1) / steps,
        "max_chi": int(info["max_chi"].max()), "svds": int(info["svds"].sum()),
        "sweeps_per_svd": float(info["sweeps"].sum() / max(1, info["svds"].sum())),
        "cpu_oracle_steps_per_s_one_thread": (warm + steps) / cpu,
        "host_cores": os.cpu_count(),
        "max_dev_vs_oracle_member": dev,
        "device_bytes": be.device_bytes()}), flush=True)


if __name__ == "__main__":
    main()
