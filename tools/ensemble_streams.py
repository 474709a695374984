"""GPU: ensemble of independent TEMPO runs (BASELINE config 5 shape) on ONE B200, one host
thread + CUDA stream + CudaOps per concurrent member (ctypes releases the GIL during the
C-ABI calls and stream waits).  Prints aggregate steps/s for several widths, next to the
CPU oracle run with the same number of worker PROCESSES-worth of threads."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oqupy_b200 as ob  # noqa: E402
from oqupy_b200._lib import CudaOps  # noqa: E402
from conftest import golden_callables, load_golden  # noqa: E402
from oracle import tempo_np as onp  # noqa: E402


def member(g, scale, nsteps, ops, out, idx):
    influence, propagators = golden_callables(g)
    infl = g["influences"]
    # another coupling strength = element-wise power of the influence functions
    with np.errstate(divide="ignore", invalid="ignore"):
        infl_s = np.where(infl == 0, 0, np.exp(np.log(infl) * scale))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        be = ob.TempoBackend(g["initial_state"], lambda dk: None if dk < 0 else infl_s[dk],
                             g["unitary"], propagators, np.ones(4), np.ones(4), 20, 1e-7,
                             ops=ops)
        be.initialize()
        states = []
        for _ in range(nsteps):
            states.append(be.compute_step()[1])
        stream.synchronize()
    out[idx] = np.array(states)


def main():
    g = load_golden("tempo_c1_k20_eps7_n60")
    nsteps = 40
    widths = [1, 2, 4, 8, 16, 32]
    members = 32
    ops_pool = [CudaOps(0) for _ in range(max(widths))]
    # warm-up (module load, first launches)
    out = {}
    member(g, 1.0, 5, ops_pool[0], out, 0)
    for w in widths:
        out = {}
        t0 = time.perf_counter()
        idx = 0
        while idx < members:
            threads = []
            for k in range(w):
                if idx >= members:
                    break
                th = threading.Thread(target=member, args=(
                    g, 1.0 + 0.01 * idx, nsteps, ops_pool[k], out, idx))
                th.start()
                threads.append(th)
                idx += 1
            for th in threads:
                th.join()
        dt = time.perf_counter() - t0
        print(json.dumps({"row": "config-5 shape: ensemble of TEMPO runs (K=20, eps=1e-7) on one B200",
                          "concurrent_members": w, "members": members, "steps_each": nsteps,
                          "aggregate_steps_per_s": members * nsteps / dt}), flush=True)
    # CPU oracle, one member, all host threads
    influence, propagators = golden_callables(g)
    orc = onp.TempoOracle(g["initial_state"], influence, g["unitary"], propagators,
                          np.ones(4), np.ones(4), 20, 1e-7)
    orc.initialize()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        orc.compute_step()
    cpu = nsteps / (time.perf_counter() - t0)
    print(json.dumps({"row": "CPU oracle, one member", "steps_per_s": cpu,
                      "host_threads": os.cpu_count()}), flush=True)
    # sanity: member 0 equals the single-run result of the same inputs
    ref = {}
    member(g, 1.0, nsteps, ops_pool[0], ref, 0)
    print("max |concurrent - single| member 0:", float(np.abs(out[0] - ref[0]).max()))


if __name__ == "__main__":
    main()
