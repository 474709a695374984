"""bench.py -- PT-TEMPO steps/s at (dkmax=200, epsrel=1e-9) on B200 (BASELINE.json metric).

A "step" is one ``PtTempoBackend.compute_step`` (one column of the PT-TEMPO network:
201 influence-MPO x MPS site contractions + 201 truncated SVDs in the zip-up, then the
200 truncated SVDs of the sweep) of BASELINE.json configs[1]: spin-boson process tensor,
ohmic alpha=0.08, wc=4, T=1.6, dt=0.05, dkmax=200, epsrel=1e-9, N=1000.  The run is
time-sequential and the cost of a step grows with the bond dimension, so the timed window is
part of the workload definition: after ``initialize()`` the build runs PREROLL = 25 untimed
steps (bond dimension 50 -> 260), then W warm-up steps, then EXACTLY K timed steps (steps
PREROLL+W+2 .. PREROLL+W+K+1 of the 1000; with the driver's W=5, K=20: steps 32..51, bond
dimension 320 -> 520, largest truncated SVD 1200x1100 -> 1950x1800).  Round 1 timed steps
7..26; ``--preroll 0`` reproduces that window.  Inputs are the reference's own influence
matrices (tests/golden/c2_operands.npz).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--preroll P] [--impl reference]

N > 1 (torchrun): one process per GPU, every rank builds the SAME process tensor (identical
cost per member: "weak" scaling, no data-path collective) through
``oqupy_b200.ensemble.run_ensemble``; the one collective is its final gather (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PT-TEMPO steps/s at dkmax=200, epsrel=1e-9"
WORKLOAD = ("spin-boson PT-TEMPO process tensor build, ohmic alpha=0.08 wc=4 T=1.6, "
            "dt=0.05, dkmax=200, epsrel=1e-9, N=1000 (BASELINE configs[1])")
NUM_STEPS, DKMAX, EPSREL = 1000, 200, 1e-9


def load_operands():
    with np.load(os.path.join(ROOT, "tests", "golden", "c2_operands.npz")) as f:
        return f["influences"]


def influence_fn(infl):
    def influence(dk):
        return None if dk < 0 else infl[dk]
    return influence


def window(args):
    """(first timed step, last timed step) of the 1000-step build."""
    first = args.preroll + args.warmup + 2
    return [first, first + args.steps - 1]


def config(args):
    # identical in both arms (the driver compares the two config objects)
    return {"workload": WORKLOAD, "timed_steps": window(args), "preroll": args.preroll}


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.thread, self.index = [], None, None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms",
                 os.environ.get("B200_BENCH_CLOCK_MS", "1000")],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi missing"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for s in self.samples:
            if len(s) < 7:
                continue
            try:
                sm.append(float(s[0]))
                smax.append(float(s[1]))
            except ValueError:
                continue
            for name, val in zip(names, s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle)
def all_host_threads():
    try:    # torchrun exports OMP_NUM_THREADS=1: give LAPACK/BLAS all host cores back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:  # pylint: disable=broad-except
        pass


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:  # pylint: disable=broad-except
        return os.cpu_count() or 1


def oracle_at(infl, sites=None, step=None):
    """The oracle port (numpy/LAPACK restatement of the reference path), either freshly
    initialised or positioned at `step` with the given MPS sites (chi_l, d2, chi_r)."""
    from oracle import tempo_np as onp
    orc = onp.PtTempoOracle(2, influence_fn(infl), NUM_STEPS, DKMAX, EPSREL)
    orc.initialize()
    if sites is not None:
        orc.mps = [np.array(s) for s in sites]
        orc.step = step
    return orc


def time_oracle(orc, steps, budget_s):
    """Timed oracle steps; stops early when the time budget is exhausted."""
    done, t0 = 0, time.perf_counter()
    for _ in range(steps):
        orc.compute_step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0


def reference_arm(args):
    """The reference path on the host cores, no GPU involved: the oracle port builds the
    same process tensor from step 1 (pre-roll and warm-up untimed) and its steps in the same
    window are timed until the CPU budget runs out."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    all_host_threads()
    infl = load_operands()
    orc = oracle_at(infl)
    t0 = time.perf_counter()
    for _ in range(args.preroll + args.warmup):
        orc.compute_step()
    pre_s = time.perf_counter() - t0
    done, dt = time_oracle(orc, args.steps, args.cpu_budget)
    val = done / dt
    cores = host_threads()
    first = window(args)[0]
    sample = (f"oracle port (numpy/LAPACK gesdd restatement of the reference path), steps "
              f"{first}..{first + done - 1} of the same build ({done} of {args.steps} requested, "
              f"{dt:.1f} s; untimed pre-roll on the host {pre_s:.1f} s)")
    line = {
        "impl": "reference", "metric": METRIC,
        "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": config(args),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores,
                         "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def fp64_gemm_peak(torch, device):
    """cuBLAS DGEMM 4096^3 burst, the FP64 (DMMA) denominator: MEASURED_PEAKS.json has
    no fp64 entry (SURVEY 8d asks to measure it on the box)."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def build_member(ops, infl, args, torch, sync, snapshot=False):
    """One PT-TEMPO build up to the end of the timed window.  Returns the backend and the
    measurements of this member."""
    import oqupy_b200 as ob
    be = ob.PtTempoBackend(2, influence_fn(infl), None, np.ones(4), np.ones(4),
                           NUM_STEPS, DKMAX, EPSREL, ops=ops)
    t0 = time.perf_counter()
    be.initialize()
    torch.cuda.synchronize(ops.device)
    init_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(args.preroll):
        be.compute_step()
    torch.cuda.synchronize(ops.device)
    pre_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        be.compute_step()
    sites = None
    snap_s = snap_bytes = None
    if snapshot:     # start state of the CPU sample (device -> host, before the timed region)
        torch.cuda.synchronize(ops.device)
        t0 = time.perf_counter()
        sites = [ops.to_host(be._site(k)) for k in range(be._num_sites())]  # pylint: disable=protected-access
        snap_s = time.perf_counter() - t0
        snap_bytes = int(sum(x.nbytes for x in sites))
    bond_hist = []
    ops.profile_enable(True)
    ops.profile_read_kinds()
    be.pop_svd_log()                 # native chain: switch the per-SVD log on
    be.chain_stats(reset=True)
    l0, h0, d0 = ops.launch_count(), ops.h2d_bytes, ops.d2h_bytes
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    sync()
    cuprof = os.environ.get("B200_BENCH_CUPROF") == "1"     # ncu --profile-from-start off
    if cuprof:
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        be.compute_step()
        bond_hist.append(be.get_bond_dimensions())      # host-side shapes, no device sync
    e1.record()
    sync()
    wall = time.perf_counter() - t0
    if cuprof:
        torch.cuda.cudart().cudaProfilerStop()
    out = {"dev_ms": e0.elapsed_time(e1), "wall_ms": wall * 1e3, "sites": sites,
           "bond_hist": bond_hist,
           "launches": ops.launch_count() - l0, "h2d": ops.h2d_bytes - h0,
           "d2h": ops.d2h_bytes - d0 + be.chain_stats()[2],
           "initialize_s": init_s, "preroll_s": pre_s,
           "snapshot_s": snap_s, "snapshot_bytes": snap_bytes}
    kinds, flops, sweeps = ops.profile_read_kinds()
    ops.profile_enable(False)
    out.update({"kinds": kinds, "flops": flops, "sweeps": sweeps,
                "svd_log": be.pop_svd_log()})
    return be, out


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oqupy_b200._lib import CudaOps
    from oqupy_b200.ensemble import run_ensemble
    ops = CudaOps(local)
    dev = ops.device
    infl = load_operands()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    state = {}

    def run_member(_index):
        be, out = build_member(ops, infl, args, torch, sync,
                               snapshot=(world == 1 and not args.no_cpu))
        state["be"], state["out"] = be, out
        bonds = np.zeros(1024)
        b = be.get_bond_dimensions()
        bonds[:len(b)] = b
        # per-member record gathered by run_ensemble: timings + bond dimensions
        ksum = sum(v[0] for v in out["kinds"].values())
        return np.concatenate(([out["dev_ms"], out["wall_ms"], ksum, out["launches"]], bonds))

    # product API: one member per rank, sharded and gathered by run_ensemble (NCCL)
    gathered = run_ensemble(world, run_member, device=dev)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    be, out = state["be"], state["out"]
    dev_ms = float(np.max(gathered[:, 0]))          # max over ranks
    wall_ms = float(np.max(gathered[:, 1]))
    launches = int(np.sum(gathered[:, 3]))
    per_rank = [{"rank": r, "dev_ms": round(float(gathered[r, 0]), 1),
                 "wall_ms": round(float(gathered[r, 1]), 1),
                 "svd_kernel_ms": round(float(gathered[r, 2]), 1)} for r in range(world)]

    peak = fp64_gemm_peak(torch, dev)
    value = world * args.steps / (dev_ms * 1e-3)
    e2e = world * args.steps / (wall_ms * 1e-3)
    kinds = out["kinds"]
    k_ms = sum(v[0] for v in kinds.values())
    k_launches = sum(v[1] for v in kinds.values())
    achieved = out["flops"] / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    n_svd = kinds["jacobi"][1]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_svd_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_svd")
    svd_log = out["svd_log"]
    big = max(svd_log, key=lambda x: x[0] * x[1]) if svd_log else None

    # bounded CPU baseline on this box's host cores (oracle port): the FIRST steps of the
    # timed window (the cheapest ones: bond dimensions grow), started from the device chain
    # as it stood when the timed region began -- rank 0, N=1 only
    cpu = None
    parity = None
    if not (args.no_cpu or world > 1):
        all_host_threads()
        first = window(args)[0]
        orc = oracle_at(infl, out["sites"], first - 1)
        cdone, cdt = time_oracle(orc, min(args.cpu_steps, args.steps), args.cpu_budget)
        # the CUDA path took the same steps from the same state: per-bond dimensions
        mine = out["bond_hist"][cdone - 1][1:-1]
        theirs = orc.bond_dimensions()
        diff = [abs(int(a) - int(b)) for a, b in zip(mine, theirs)]
        parity = {"steps_compared": cdone, "bonds_compared": len(diff),
                  "bonds_equal": sum(x == 0 for x in diff),
                  "max_abs_diff": max(diff) if diff else 0,
                  "max_bond": int(max(theirs)) if len(theirs) else 0,
                  "note": ("the oracle starts from the device chain at the start of the timed "
                           "window and takes the steps of the CPU sample; per-bond dimensions "
                           "of both after those steps (differences are truncation-threshold "
                           "ties, tests/test_parity_gpu.py logs their margins)")}
        cpu = {"value": cdone / cdt, "unit": "steps/s", "cores": host_threads(), "kind": "port",
               "sample": (f"oracle port (numpy/LAPACK), steps {first}..{first + cdone - 1} of "
                          f"the same build (the first, cheapest steps of the timed window) "
                          f"started from the device chain, {cdt:.1f} s")}

    line = {
        "metric": METRIC,
        "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c128", "data": "synthetic",
        "config": config(args),
        "detail": {
            "parallelism": f"{world} identical ensemble members, one per GPU, via "
                           "oqupy_b200.ensemble.run_ensemble (one NCCL gather at the end)",
            "l2": ("time-sequential build: every SVD operand is new data produced by "
                   "the previous kernel; no repeated iteration, no L2 flush needed"),
            "largest_svd": None if big is None else [big[0], big[1], big[2]],
            "svds_per_step": len(svd_log) / args.steps if svd_log else None,
            "max_bond": max(be.get_bond_dimensions()),
            "initialize_s": round(out["initialize_s"], 3),
            "preroll_s": round(out["preroll_s"], 2),
            "update_process_tensor": {
                "note": ("not in the timed loop (the reference reads the sites out once, after "
                         "the last step: pt_tempo.py:275-278).  What it costs with a HOST "
                         "process tensor is one device->host copy of every site; measured here "
                         "as the read-back of the whole chain at the start of the timed window "
                         "(untimed; null when the CPU sample is off).  With a "
                         "DeviceProcessTensor nothing leaves the device."),
                "chain_readback_s": (None if out["snapshot_s"] is None
                                     else round(out["snapshot_s"], 4)),
                "chain_bytes": out["snapshot_bytes"]},
            "per_rank": per_rank,
        },
        "e2e": {"value": e2e, "unit": "steps/s",
                "h2d_bytes_per_step": out["h2d"] / args.steps,
                "d2h_bytes_per_step": out["d2h"] / args.steps,
                "note": ("wall clock around PtTempoBackend.compute_step() (public "
                         "API): host influence callback + H2D of its operands and "
                         "the per-bond rank read-backs (D2H) happen inside every step")},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {
            "kernel": ("truncated SVD = qrcp_kernel + jacobi_kernel + apply_q/emit kernels "
                       "(fp64; one dependent chain of ~400 factorisations per step)"),
            "bound": "tensor", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "traffic": traffic,
            "peak_source": ("fp64 cuBLAS DGEMM 4096^3 measured in this run "
                            "(MEASURED_PEAKS.json carries no fp64 figure)"),
            "note": ("algorithmic flops 4(14 m n^2 + 8 n^3) per truncated SVD (SURVEY 8d) over "
                     "the summed device time of the three SVD kernel families; the chain is "
                     "latency-bound (hand-shakes per pivot / per tournament round), see "
                     "DESIGN.md section 5"),
            "kernel_ms": k_ms, "kernel_share_of_step": k_ms / out["dev_ms"],
            "kernel_ms_by_family": {k: round(v[0], 1) for k, v in kinds.items()},
            "launches": int(k_launches), "svds": int(n_svd),
            "jacobi_sweeps": int(out["sweeps"]),
            "algorithmic_flops": out["flops"],
        },
        "parity": parity,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--preroll", type=int, default=25,
                    help="untimed build steps before the warm-up (part of the workload)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-budget", type=float, default=25.0,
                    help="seconds of host time for the CPU baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.cpu_budget == 25.0:
            args.cpu_budget = 150.0
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
