"""GPU: several independent PT-TEMPO builds (config 2 operands, different couplings) running
concurrently on ONE B200, one host thread + CUDA stream + native chain per member."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oqupy_b200 as ob  # noqa: E402
from oqupy_b200._lib import CudaOps  # noqa: E402
import bench  # noqa: E402


def member(rank, warm, steps, ops, out, barrier):
    _, infl = bench.load_operands(rank)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        be = ob.PtTempoBackend(2, bench.influence_fn(infl), None, np.ones(4), np.ones(4),
                               1000, 200, 1e-9, ops=ops)
        be.initialize()
        for _ in range(warm):
            be.compute_step()
        stream.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            be.compute_step()
        stream.synchronize()
        out[rank] = (time.perf_counter() - t0, be.get_bond_dimensions())


def main():
    warm, steps = 3, 20
    single = None
    for w in (1, 2, 3, 4):
        ops_pool = [CudaOps(0) for _ in range(w)]
        out = {}
        barrier = threading.Barrier(w)
        threads = [threading.Thread(target=member, args=(k, warm, steps, ops_pool[k], out,
                                                         barrier)) for k in range(w)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        dt = max(v[0] for v in out.values())
        agg = w * steps / dt
        if w == 1:
            single = out[0][1]
        print(json.dumps({"row": "concurrent PT-TEMPO builds on one B200 (config 2, steps 5-24)",
                          "concurrent_members": w, "aggregate_steps_per_s": agg,
                          "member0_bonds_equal_single": out[0][1] == single}), flush=True)


if __name__ == "__main__":
    main()
