"""GPU: run_ensemble(..., concurrent=k) (threads + ops + CUDA streams per member) against the
serial run of the same 8 TEMPO members."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import load_golden  # noqa: E402
from oqupy_b200._lib import CudaOps  # noqa: E402
from oqupy_b200.ensemble import run_ensemble, tempo_member  # noqa: E402

g = load_golden("tempo_c1_k20_eps7_n60")
p1, p2 = g["prop_1"], g["prop_2"]


def infl(i):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(g["influences"] == 0, 0,
                        np.exp(np.log(g["influences"]) * (1.0 + 0.02 * i)))


def member(i, ops=None):
    return tempo_member(infl(i), lambda s: (p1, p2), g["initial_state"], 20, 1e-7, 30, ops=ops)


t0 = time.perf_counter()
serial = run_ensemble(8, member)
t1 = time.perf_counter()
conc = run_ensemble(8, member, concurrent=8, make_ops=lambda: CudaOps(0))
t2 = time.perf_counter()
print("serial %.2fs concurrent %.2fs  max diff %.1e" % (t1 - t0, t2 - t1,
                                                          np.abs(serial - conc).max()))
