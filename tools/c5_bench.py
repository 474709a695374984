"""GPU (1..8 ranks): BASELINE configs[4] -- the ensemble sweep of independent TEMPO runs over
the coupling x temperature grid (alpha in linspace(0.02, 0.30), T in
linspace(0.2, 3.2), K=20, eps=1e-7, dt=0.05), sharded over the ranks by
oqupy_b200.ensemble.tempo_grid and advanced in lock-step on every GPU.

    python tools/c5_bench.py [grid=64] [steps=200]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \\
        tools/c5_bench.py [grid] [steps]                        (N GPUs, NCCL gather)

Rank 0 prints one JSON line: runs, member-steps/s (device time, max over ranks), the CPU
oracle on a sample of members (one host thread each), and the deviation from it.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from oqupy_b200._lib import CudaOps  # noqa: E402
from oqupy_b200.ensemble import tempo_grid  # noqa: E402


def scaled(m, f):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(m == 0, 0, np.exp(np.log(np.where(m == 0, 1, m)) * f))


def main():
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ops = CudaOps(local)
    g = load_golden("c5_grid_operands")
    ti = np.linspace(0, 63, grid).round().astype(int)
    alphas = np.linspace(0.02, 0.30, grid)
    infl = np.array([scaled(g["influences"][t], a / float(g["alpha_ref"]))
                     for a in alphas for t in ti])              # member = (alpha, T) pair
    n = infl.shape[0]
    if world > 1:
        # warm-up of the collectives the gather uses (NCCL sets its rings / trees up on the first
        # call of every collective type: a per-process cost, not a per-job one)
        from oqupy_b200.ensemble import run_ensemble
        run_ensemble(world * 2, lambda i: np.zeros((3, 2, 2), dtype=complex) + i,
                     device=ops.device)
        dist.barrier()
    torch.cuda.synchronize()
    timings = {}
    t0 = time.perf_counter()
    res, rerun = tempo_grid(infl, g["initial_state"], g["unitary"],
                            lambda s: (g["prop_1"], g["prop_2"]), int(g["dkmax"]),
                            float(g["epsrel"]), steps, device=ops.device, ops=ops,
                            timings=timings,
                            reserve=int(os.environ.get("B200_BATCH_RESERVE", "24")))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tt = torch.tensor([wall], dtype=torch.float64, device=ops.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        from oracle import tempo_np as onp
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=1)
        except Exception:  # pylint: disable=broad-except
            pass
        sample = [0, n // 3, n // 2, n - 1]
        dev, cpu_s = 0.0, 0.0
        for i in sample:
            tb = onp.TempoOracle(g["initial_state"], lambda dk, m=infl[i]: None if dk < 0 else m[dk],
                                 g["unitary"], lambda s: (g["prop_1"], g["prop_2"]),
                                 np.ones(4), np.ones(4), int(g["dkmax"]), float(g["epsrel"]))
            c0 = time.perf_counter()
            _, s0 = tb.initialize()
            ref = np.array([s0] + [tb.compute_step()[1] for _ in range(steps)]).reshape(-1, 2, 2)
            cpu_s += time.perf_counter() - c0
            dev = max(dev, float(np.abs(res[i] - ref).max()))
        print(json.dumps({
            "row": "BASELINE configs[4]: TEMPO ensemble over the alpha x T grid, lock-step",
            "n_gpus": world, "runs": n, "steps_per_run": steps,
            "member_steps_per_s": n * steps / float(tt[0]),
            "runs_per_s": n / float(tt[0]), "seconds": float(tt[0]),
            "members_on_general_path": len(rerun),
            "rank0_phase_seconds": {k: round(v, 3) for k, v in timings.items()},
            "cpu_oracle_steps_per_s_one_thread": len(sample) * steps / cpu_s,
            "host_cores": os.cpu_count(),
            "max_dev_vs_oracle_sample": dev,
            "trace_error": float(np.abs(np.trace(res, axis1=2, axis2=3) - 1).max())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
