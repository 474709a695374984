"""CPU, world_size 2 (gloo): sharding + the one collective of the design (the gather of
per-member dynamics).  Members are real TEMPO runs executed through the engine's host
logic with the test-only ops model (tests/host_model_ops.py)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_members, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from conftest import load_golden
    from host_model_ops import HostModelOps
    from oqupy_b200.ensemble import run_ensemble, shard_indices, tempo_member
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("tempo_c1_k20_eps7_n60")
    ops = HostModelOps()
    p1, p2 = g["prop_1"], g["prop_2"]

    def member(i):
        # coupling scan: influence^(1 + 0.1 i)  (eta is linear in alpha)
        with np.errstate(divide="ignore", invalid="ignore"):
            infl = np.where(g["influences"] == 0, 0,
                            np.exp(np.log(g["influences"]) * (1.0 + 0.1 * i)))
        return tempo_member(infl[:7], lambda s: (p1, p2), g["initial_state"], 6,
                            1e-6, 8, ops=ops)

    res = run_ensemble(n_members, member)
    # the same ensemble with two members of this rank in flight (thread + ops object each)
    res_c = run_ensemble(n_members, lambda i, o: tempo_member(
        np.where(g["influences"] == 0, 0,
                 np.exp(np.log(np.where(g["influences"] == 0, 1, g["influences"]))
                        * (1.0 + 0.1 * i)))[:7],
        lambda s: (p1, p2), g["initial_state"], 6, 1e-6, 8, ops=o),
        concurrent=2, make_ops=HostModelOps)
    np.testing.assert_array_equal(res_c, res)
    assert shard_indices(n_members, rank, world) == list(range(rank, n_members, world))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), res)
    dist.destroy_process_group()


def test_ensemble_two_ranks(tmp_path):
    n_members, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_members, str(tmp_path)), nprocs=world,
             join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert r0.shape == (n_members, 9, 2, 2)
    np.testing.assert_array_equal(r0, r1)          # every rank holds the full result
    # member 0 is the un-scaled coupling: compare with a serial single-process run
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden
    from host_model_ops import HostModelOps
    from oqupy_b200.ensemble import tempo_member
    g = load_golden("tempo_c1_k20_eps7_n60")
    serial = tempo_member(g["influences"][:7], lambda s: (g["prop_1"], g["prop_2"]),
                          g["initial_state"], 6, 1e-6, 8, ops=HostModelOps())
    np.testing.assert_allclose(r0[0], serial, atol=1e-12)
    # different couplings give different dynamics, traces stay 1
    assert np.abs(r0[4] - r0[0]).max() > 1e-4
    np.testing.assert_allclose(np.trace(r0, axis1=2, axis2=3), 1.0, atol=1e-4)


def _bcast_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from conftest import golden_callables, load_golden
    from host_model_ops import HostModelOps
    import oqupy_b200 as ob
    from oqupy_b200.ensemble import broadcast_process_tensor, run_ensemble
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("pt_k8_eps9_n24")
    influence, propagators = golden_callables(g)
    ops = HostModelOps()
    pt = None
    if rank == 0:       # only rank 0 builds the process tensor
        pt = ob.DeviceProcessTensor(2, dt=float(g["dt"]), ops=ops)
        be = ob.PtTempoBackend(2, influence, pt, np.ones(4), np.ones(4), 10, 5, 1e-8, ops=ops)
        be.initialize()
        while be.compute_step():
            pass
        be.update_process_tensor()
    pt = broadcast_process_tensor(pt, src=0, ops=ops)
    assert len(pt) == 10
    p1, p2 = propagators(0)

    def member(i):      # a panel of system Hamiltonians sharing the one process tensor
        ph = np.exp(0.05j * i)
        return ob.dynamics_device(pt, lambda s: (p1 * ph, p2), g["initial_state"], ops=ops)

    res = run_ensemble(5, member)
    np.save(os.path.join(out_dir, f"bc{rank}.npy"), res)
    dist.destroy_process_group()


def test_broadcast_process_tensor_two_ranks(tmp_path):
    """one process tensor built on rank 0, broadcast (device tensors; gloo here,
    NCCL on the GPUs), reused by every rank for its share of a Hamiltonian panel."""
    port = _free_port()
    mp.spawn(_bcast_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "bc0.npy"), np.load(tmp_path / "bc1.npy")
    assert r0.shape == (5, 11, 2, 2)
    np.testing.assert_array_equal(r0, r1)
    assert np.abs(r0[1] - r0[0]).max() > 1e-6


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from conftest import gradient_multi_setup, load_golden
    from host_model_ops import HostModelOps
    import oqupy_b200 as ob
    from oqupy_b200.ensemble import run_ensemble
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("gradient_multi")
    ops = HostModelOps()
    pts, props, _ = gradient_multi_setup(g, ops)
    n = int(g["num_steps"])
    rng = np.random.default_rng(0)              # the same perturbation directions on every rank
    direction = rng.normal(size=(6, n))

    def member(i):      # BASELINE configs[4]: control-gradient perturbations, base +- 1e-3 * unit
        def perturbed(step):
            p1, p2 = props(step)
            return p1 * (1.0 + 1e-3 * direction[i, step]), p2
        derivs, states = ob.gradient_device(pts, perturbed, g["initial_state"],
                                            g["target_derivative"], num_steps=n, ops=ops)
        return np.concatenate((np.array(derivs).reshape(-1), np.array(states).reshape(-1)))

    res = run_ensemble(6, member)
    if rank == 0:
        serial = np.array([member(i) for i in range(6)])
        np.testing.assert_array_equal(res, serial)
    np.save(os.path.join(out_dir, f"gr{rank}.npy"), res)
    dist.destroy_process_group()


def test_gradient_perturbation_panel_two_ranks(tmp_path):
    """BASELINE configs[4], second half: a panel of control-gradient evaluations (two
    environments) sharded over the ranks, adjoint tensors and dynamics gathered everywhere."""
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "gr0.npy"), np.load(tmp_path / "gr1.npy")
    assert r0.shape[0] == 6
    np.testing.assert_array_equal(r0, r1)
    assert np.abs(r0[1] - r0[0]).max() > 1e-6
