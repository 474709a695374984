"""GPU: which members of the config-5 grid leave the lock-step path, at which step, and how long
their re-run on the general backend takes alone (diagnosis of the N = 8 tail).
  python tools/c5_overflow_probe.py [grid=32] [steps=200]
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
from oqupy_b200.ensemble import tempo_member  # noqa: E402


def scaled(m, f):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(m == 0, 0, np.exp(np.log(np.where(m == 0, 1, m)) * f))


def main():
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    g = load_golden("c5_grid_operands")
    ti = np.linspace(0, 63, grid).round().astype(int)
    alphas = np.linspace(0.02, 0.30, grid)
    # the `rows` strongest couplings of the grid
    infl = np.array([scaled(g["influences"][t], a / float(g["alpha_ref"]))
                     for a in alphas[-rows:] for t in ti])
    n = infl.shape[0]
    props = lambda s: (g["prop_1"], g["prop_2"])  # noqa: E731
    be = ob.BatchedTempoBackend(np.array([g["initial_state"].reshape(-1)] * n), infl,
                                g["unitary"], props, np.ones(4), np.ones(4), int(g["dkmax"]),
                                float(g["epsrel"]))
    be.initialize()
    first_fail = {}
    for s in range(1, steps + 1):
        be.compute_steps(1, strict=False)
        st = be.info()["status"]
        for k in np.nonzero(st)[0]:
            first_fail.setdefault(int(k), (s, int(st[k])))
    info = be.info()
    out = {"members": n, "left_lock_step": {str(k): v for k, v in first_fail.items()},
           "alpha_T_index": {str(k): [float(alphas[-rows:][k // grid]), int(ti[k % grid])]
                             for k in first_fail},
           "max_chi_top5": sorted(info["max_chi"].tolist())[-5:]}
    for k in list(first_fail)[:2]:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = tempo_member(infl[k], props, g["initial_state"], int(g["dkmax"]),
                           float(g["epsrel"]), steps, unitary=g["unitary"])
        torch.cuda.synchronize()
        out[f"general_path_alone_s_member_{k}"] = round(time.perf_counter() - t0, 3)
        out[f"trace_member_{k}"] = float(abs(np.trace(res[-1]) - 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
