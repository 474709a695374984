"""GPU: measurement of the other hot-path rows (SURVEY 8a A5/A8/A9, 8f gradient) next to
the CPU oracle, one JSON line per row (-> profiles/).  Not the headline (bench.py is).

  python tools/bench_rows.py > gpurun_out/rows.jsonl
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
from conftest import golden_callables, load_golden, tebd_fixture  # noqa: E402
from oracle import tempo_np as onp  # noqa: E402
from oracle import tebd_np  # noqa: E402

HBM_GBS = 6533.8
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        HBM_GBS = json.load(f)["hbm_gbs"]
except (OSError, KeyError):
    pass


def emit(**kw):
    print(json.dumps(kw), flush=True)


def tempo_c1():
    """BASELINE configs[0]: spin-boson TEMPO, K=20, eps=1e-7 (latency-bound: chi <= 22)."""
    g = load_golden("tempo_c1_k20_eps7_n60")
    influence, propagators = golden_callables(g)
    n = int(g["num_steps"])
    be = ob.TempoBackend(g["initial_state"], influence, g["unitary"], propagators,
                         np.ones(4), np.ones(4), 20, 1e-7)
    be.initialize()
    for _ in range(5):
        be.compute_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n - 5):
        be.compute_step()
    torch.cuda.synchronize()
    gpu = (n - 5) / (time.perf_counter() - t0)
    orc = onp.TempoOracle(g["initial_state"], influence, g["unitary"], propagators,
                          np.ones(4), np.ones(4), 20, 1e-7)
    orc.initialize()
    for _ in range(5):
        orc.compute_step()
    t0 = time.perf_counter()
    for _ in range(n - 5):
        orc.compute_step()
    cpu = (n - 5) / (time.perf_counter() - t0)
    emit(row="A5 TEMPO step (config 1: K=20, eps=1e-7, chi<=22)", metric="steps/s",
         gpu=gpu, cpu_oracle=cpu, bound="latency (39 dependent SVDs <= 88x88 per step)",
         note="single run; the batched small-matrix path for ensembles is a next item")


def synthetic_pt(ops, n, chi, d=2, seed=0):
    """A process tensor of N sites with bond dimension chi (random, contractive)."""
    rng = np.random.default_rng(seed)
    pt = ob.DeviceProcessTensor(d, dt=0.1, ops=ops)
    dims = [1] + [chi] * (n - 1) + [1]
    for k in range(n):
        t = rng.normal(size=(dims[k], dims[k + 1], d * d)) \
            + 1j * rng.normal(size=(dims[k], dims[k + 1], d * d))
        t *= 0.5 / np.sqrt(dims[k] * d * d)
        pt.set_mpo_tensor(k, t)
    pt.compute_caps()
    return pt, dims


def dynamics_rows():
    ops = ob.default_ops()
    n, chi, d2 = 64, 1024, 4
    pt, dims = synthetic_pt(ops, n, chi)
    rng = np.random.default_rng(1)
    p1 = np.eye(4) + 0.01 * (rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    p2 = np.eye(4) + 0.01 * (rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    rho0 = np.array([[1.0, 0.0], [0.0, 0.0]], dtype=complex)
    t_bytes = sum(16 * d2 * dims[k] * dims[k + 1] for k in range(n))
    mpos = [pt.get_mpo_tensor(k) for k in range(n)]
    caps = [pt.get_cap_tensor(k) for k in range(n + 1)]
    t0 = time.perf_counter()
    ref = onp.compute_dynamics([mpos], [caps], lambda s: (p1, p2), rho0)
    cpu_s = time.perf_counter() - t0
    for nvec in (1, 64):
        rho = np.array([rho0] * nvec) if nvec > 1 else rho0
        p1s = np.array([p1] * nvec) if nvec > 1 else p1
        p2s = np.array([p2] * nvec) if nvec > 1 else p2
        props = lambda s: (p1s, p2s)   # noqa: E731
        out = ob.dynamics_device(pt, props, rho)
        err = float(np.abs((out[0] if nvec > 1 else out) - ref).max())
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            ob.dynamics_device(pt, props, rho)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        emit(row=f"A8 compute_dynamics, 1 environment, E={nvec} members sharing the PT",
             metric="PT steps/s", gpu=n / best, cpu_oracle=n / cpu_s,
             workload=f"N={n} sites, chi={chi}, d2=4 (synthetic PT, {t_bytes / 1e6:.0f} MB)",
             roofline={"bound": "hbm", "achieved": t_bytes / best / 1e9, "peak": HBM_GBS,
                       "unit": "GB/s", "frac": t_bytes / best / 1e9 / HBM_GBS,
                       "algorithmic_bytes": t_bytes,
                       "note": "end-to-end call incl. per-step launches; T_k read once"},
             max_abs_err_vs_oracle=err)
    # caps (best of 3: the first call after the ensemble run above pays for torch's
    # allocator returning the large workspace)
    dt = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pt.compute_caps()
        torch.cuda.synchronize()
        dt = min(dt, time.perf_counter() - t0)
    emit(row="A9 compute_caps", metric="sites/s", gpu=n / dt,
         roofline={"bound": "hbm", "achieved": t_bytes / dt / 1e9, "peak": HBM_GBS,
                   "unit": "GB/s", "frac": t_bytes / dt / 1e9 / HBM_GBS,
                   "algorithmic_bytes": t_bytes})
    # gradient
    target = np.array([[0.0, 0.0], [0.0, 1.0]], dtype=complex)
    t0 = time.perf_counter()
    dref, sref = onp.compute_gradient_and_dynamics(mpos, caps, lambda s: (p1, p2), rho0,
                                                   target)
    cpu_g = time.perf_counter() - t0
    derivs, states = ob.gradient_device(pt, lambda s: (p1, p2), rho0, target)
    err = float(np.abs(np.array(derivs) - np.array(dref)).max())
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ob.gradient_device(pt, lambda s: (p1, p2), rho0, target)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    emit(row="8f-1 gradient (forward + back-propagation + adjoint tensors)",
         metric="PT steps/s", gpu=n / best, cpu_oracle=n / cpu_g,
         workload=f"N={n} sites, chi={chi}",
         roofline={"bound": "hbm", "achieved": 3 * t_bytes / best / 1e9, "peak": HBM_GBS,
                   "unit": "GB/s", "frac": 3 * t_bytes / best / 1e9 / HBM_GBS,
                   "algorithmic_bytes": 3 * t_bytes,
                   "note": "T_k is read by the forward, the backward and the adjoint pass"},
         max_abs_err_vs_oracle=err)
    # two environments
    pt_b, _ = synthetic_pt(ops, 16, 96, seed=5)
    pt_a, _ = synthetic_pt(ops, 16, 96, seed=6)
    ma = [pt_a.get_mpo_tensor(k) for k in range(16)]
    mb = [pt_b.get_mpo_tensor(k) for k in range(16)]
    ca = [pt_a.get_cap_tensor(k) for k in range(17)]
    cb = [pt_b.get_cap_tensor(k) for k in range(17)]
    t0 = time.perf_counter()
    ref2 = onp.compute_dynamics([ma, mb], [ca, cb], lambda s: (p1, p2), rho0)
    cpu2 = time.perf_counter() - t0
    out2 = ob.dynamics_device([pt_a, pt_b], lambda s: (p1, p2), rho0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ob.dynamics_device([pt_a, pt_b], lambda s: (p1, p2), rho0)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    emit(row="A8 compute_dynamics, 2 environments (config 3 shape)", metric="PT steps/s",
         gpu=16 / dt2, cpu_oracle=16 / cpu2, workload="N=16, chi=96 per environment",
         max_abs_err_vs_oracle=float(np.abs(out2 - ref2).max()))


def tebd_row(n_sites=16):
    """SURVEY 8a A7 / BASELINE configs[3] shape: PT-TEBD chain of 16 spins, a process tensor
    on EVERY site.  Inputs: the gates and the process tensor of the reference's test F
    (tests/golden/pt_tebd_F2.npz; chi_pt up to 51), the bond gates tiled to 16 sites."""
    from types import SimpleNamespace
    g = load_golden("pt_tebd_F2")
    gammas, lambdas, layers, _, mpos, caps = tebd_fixture(g)
    # eps_tebd = 1e-5 (BASELINE configs[3]); 4 steps: the bond dimension reaches 74 and
    # chi_pt 51, the splits are tall (3800 x 300) and the CPU oracle needs ~30 s
    eps, steps = 1.0e-5, 4
    by_bond = [{} for _ in layers]
    for li, layer in enumerate(layers):
        for sites, tensors in layer:
            by_bond[li][sites[0]] = tensors
    big_layers = []
    for li, layer in enumerate(layers):
        parity = layer[0][0][0] % 2
        src = sorted(by_bond[li])
        big_layers.append([((b, b + 1), by_bond[li][src[(b // 2) % len(src)]])
                           for b in range(parity, n_sites - 1, 2)])
    gam = [gammas[0]] * n_sites
    lam = [lambdas[0]] * (n_sites - 1)
    # --- device
    ops = ob.default_ops()
    pt = ob.DeviceProcessTensor(2, dt=float(g["dt"]), ops=ops)
    for k, t in enumerate(mpos):
        pt.set_mpo_tensor(k, t)
    pt.compute_caps()
    gate_layers = [SimpleNamespace(gates=[SimpleNamespace(sites=list(s), tensors=list(t))
                                          for s, t in layer]) for layer in big_layers]

    def run_device(steps=steps):
        par = int(os.environ.get("TEBD_PARALLEL", "0"))
        be = ob.PtTebdBackend(gam, lam, eps, {"parallel": par} if par else {}, ops=ops)
        nsvd0 = ops.launch_count()
        for step in range(1, steps + 1):
            for layer in gate_layers:
                be.apply_nn_gate_layer(layer)
            be.apply_process_tensors(step, [pt] * n_sites)
            for layer in gate_layers:
                be.apply_nn_gate_layer(layer)
            be.compute_traces(step, [pt] * n_sites)
        rho = [be.get_density_matrix([s]) for s in range(n_sites)]
        return be, np.array(rho), ops.launch_count() - nsvd0

    run_device(steps=2)          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    be, rho, launches = run_device()
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    # --- CPU oracle
    t0 = time.perf_counter()
    orc = tebd_np.PtTebdOracle(gam, lam, eps)
    for step in range(1, steps + 1):
        for layer in big_layers:
            orc.apply_nn_gate_layer(layer)
        orc.apply_process_tensors(step, [mpos[step - 1]] * n_sites)
        for layer in big_layers:
            orc.apply_nn_gate_layer(layer)
        orc.compute_traces([caps[step]] * n_sites)
    cpu_s = time.perf_counter() - t0
    rho_ref = np.array([orc.get_density_matrix([s]) for s in range(n_sites)])
    gates_per_step = 2 * sum(len(layer) for layer in big_layers)
    emit(row="A7 PT-TEBD step (config 4 shape: 16 spins, a PT on every site)",
         metric="PT-TEBD steps/s", gpu=steps / gpu_s, cpu_oracle=steps / cpu_s,
         workload=f"{n_sites} sites, {gates_per_step} nn gates = {3 * gates_per_step} truncated "
                  f"SVDs per step, {steps} steps, eps={eps}, chi_pt<=51",
         bond_dims_gpu=[int(x) for x in be.get_bond_dimensions()],
         bond_dims_oracle=[int(x) for x in orc.get_bond_dimensions()],
         max_abs_err_vs_oracle=float(np.abs(rho - rho_ref).max()), gpu_launches=launches,
         parallel=int(os.environ.get("TEBD_PARALLEL", "0")),
         bound="latency (a chain of dependent SVDs per gate; TEBD_PARALLEL=k runs the gates "
               "of a layer on k streams)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tebd":
        tebd_row()
        sys.exit(0)
    tempo_c1()
    dynamics_rows()
