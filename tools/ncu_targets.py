"""GPU: a short program whose launches ncu captures (profiles/r02_*): the truncated SVD
pipeline on two graded operands (408 x 400, 968 x 864; QR path) and three steps of the
lock-step TEMPO ensemble (148 members).  Usage (under gpurun):
  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 \
      -o gpurun_out/<name> python tools/ncu_targets.py [svd|batch]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oqupy_b200 as ob  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from qr_check import graded  # noqa: E402


def main():
    what = sys.argv[1:] or ["svd", "batch"]
    ops = ob.default_ops()
    rng = np.random.default_rng(3)
    if "svd" in what:
        for (m, n) in [(408, 400), (968, 864)]:
            theta = graded(rng, m, n)
            d = ops.from_host(theta)
            h = ops.svd_factor(d, m, n, n, 1, 1e-9)
            k = h.keep
            u, svh = ops.empty(m, k), ops.empty(k, n)
            ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sj=1, svh=svh)
            ops.synchronize()
            print(m, n, "keep", k, "plan", ops.svd_plan(h))
    if "batch" in what:
        from conftest import load_golden
        from test_batch_gpu import scaled
        g = load_golden("tempo_c1_k20_eps7_n60")
        members = 148
        infl = np.array([scaled(g["influences"][:21], f) for f in np.linspace(0.5, 1.5, members)])
        be = ob.BatchedTempoBackend(np.array([g["initial_state"].reshape(-1)] * members), infl,
                                    g["unitary"], lambda s: (g["prop_1"], g["prop_2"]),
                                    np.ones(4), np.ones(4), 20, 1e-7)
        be.initialize()
        be.compute_steps(24)
        print("batch max chi", be.info()["max_chi"].max())


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()
