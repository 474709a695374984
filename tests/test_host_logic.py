"""CPU: host-side index logic of the engine (oqupy_b200/chain.py, backends.py,
process_tensor.py) against the oracle and the reference-generated fixtures, using
the test-only strided-view model of the C-ABI (tests/host_model_ops.py)."""
import numpy as np
import pytest

from conftest import (check_mean_field_run, golden_callables, load_golden,
                      mean_field_callables, unique_callables)
from host_model_ops import HostModelOps
from oracle import tempo_np as onp
import oqupy_b200 as ob

PT_CASES = ["pt_k12_eps7_n30", "pt_k8_eps9_n24", "pt_refA", "pt_refC"]
TEMPO_CASES = ["tempo_c1_k20_eps7_n60", "tempo_refA", "tempo_refC",
               "tempo_nondiag"]


def build_pt(g, ops):
    influence, propagators = golden_callables(g)
    d = int(g["dim"])
    pt = ob.DeviceProcessTensor(d, dt=float(g["dt"]), ops=ops)
    be = ob.PtTempoBackend(d, influence, pt, np.ones(d * d), np.ones(d * d),
                           int(g["num_steps"]), int(g["dkmax"]),
                           float(g["epsrel"]), ops=ops)
    be.initialize()
    while be.compute_step():
        pass
    be.update_process_tensor()
    return be, pt, propagators


@pytest.mark.parametrize("name", PT_CASES)
def test_pt_backend_host_logic(name):
    g = load_golden(name)
    ops = HostModelOps()
    be, pt, propagators = build_pt(g, ops)
    assert list(pt.get_bond_dimensions()) == list(g["bond_dims"])
    states = ob.dynamics_device(pt, propagators, g["initial_state"], ops=ops)
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(pt.get_cap_tensor(0), g["cap_first"], atol=1e-10)


def test_dynamics_ensemble_host_logic():
    g = load_golden("pt_k12_eps7_n30")
    ops = HostModelOps()
    _, pt, propagators = build_pt(g, ops)
    rng = np.random.default_rng(0)
    p1, p2 = propagators(0)
    e = 5
    p1s = np.array([p1 * (1 + 0.01 * k) for k in range(e)])
    p2s = np.array([p2] * e)
    rho0 = np.array([g["initial_state"]] * e)
    out = ob.dynamics_device(pt, lambda s: (p1s, p2s), rho0, ops=ops)
    mpos = [pt.get_mpo_tensor(k) for k in range(len(pt))]
    caps = [pt.get_cap_tensor(k) for k in range(len(pt) + 1)]
    for k in range(e):
        ref = onp.compute_dynamics([mpos], [caps], lambda s: (p1s[k], p2s[k]),
                                   g["initial_state"])
        np.testing.assert_allclose(out[k], ref, atol=1e-12, rtol=0)
    del rng


@pytest.mark.parametrize("name", TEMPO_CASES)
def test_tempo_backend_host_logic(name):
    g = load_golden(name)
    influence, propagators = golden_callables(g)
    d = int(g["dim"])
    d2 = d * d
    dkmax = None if int(g["dkmax"]) < 0 else int(g["dkmax"])
    ops = HostModelOps()
    be = ob.TempoBackend(g["initial_state"], influence, g["unitary"],
                         propagators, np.ones(d2), np.ones(d2), dkmax,
                         float(g["epsrel"]), ops=ops)
    _, s0 = be.initialize()
    states = [s0]
    for _ in range(int(g["num_steps"])):
        states.append(be.compute_step()[1])
    states = np.array(states).reshape(-1, d, d)
    assert be.get_bond_dimensions() == list(g["bond_dims"])
    np.testing.assert_allclose(states, g["states"],
                               atol=50 * float(g["epsrel"]), rtol=0)
    np.testing.assert_allclose(states[:5], g["states"][:5], atol=1e-9, rtol=0)


def run_unique_pt(g, ops):
    """PT-TEMPO with degeneracy maps (unique=True, pt_tempo_backend.py:114-140)."""
    influence, propagators = unique_callables(g)
    maps = [g["north_map"], g["west_map"]]
    nn, nw = int(maps[0].max()) + 1, int(maps[1].max()) + 1
    d = int(g["dim"])
    pt = ob.DeviceProcessTensor(d, dt=float(g["dt"]), ops=ops)
    be = ob.PtTempoBackend(d, influence, pt, np.ones(nn), np.ones(nw), int(g["num_steps"]),
                           int(g["dkmax"]), float(g["epsrel"]), degeneracy_maps=maps,
                           ops=ops)
    be.initialize()
    while be.compute_step():
        pass
    be.update_process_tensor()
    return pt, ob.dynamics_device(pt, propagators, g["initial_state"], ops=ops)


def run_unique_tempo(g, ops):
    """TEMPO with degeneracy maps (unique=True, tempo_backend.py:400-417)."""
    influence, propagators = unique_callables(g)
    maps = [g["north_map"], g["west_map"]]
    nn, nw = int(maps[0].max()) + 1, int(maps[1].max()) + 1
    d = int(g["dim"])
    be = ob.TempoBackend(g["initial_state"], influence, g["unitary"], propagators,
                         np.ones(nn), np.ones(nw), int(g["dkmax"]), float(g["epsrel"]),
                         degeneracy_maps=maps, dim=d, ops=ops)
    _, s0 = be.initialize()
    states = [s0]
    for _ in range(int(g["num_steps"])):
        states.append(be.compute_step()[1])
    return be, np.array(states).reshape(-1, d, d)


@pytest.mark.parametrize("tag", ["spin12", "spin1"])
def test_unique_pt_host_logic(tag):
    g = load_golden(f"pt_unique_{tag}")
    pt, states = run_unique_pt(g, HostModelOps())
    assert list(pt.get_bond_dimensions()) == list(g["bond_dims"])
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)


@pytest.mark.parametrize("tag", ["spin12", "spin1"])
def test_unique_tempo_host_logic(tag):
    g = load_golden(f"tempo_unique_{tag}")
    be, states = run_unique_tempo(g, HostModelOps())
    assert be.get_bond_dimensions() == list(g["bond_dims"])
    np.testing.assert_allclose(states, g["states"], atol=50 * float(g["epsrel"]), rtol=0)
    np.testing.assert_allclose(states[:3], g["states"][:3], atol=1e-9, rtol=0)


def test_bad_degeneracy_maps_are_rejected():
    with pytest.raises(AssertionError):
        ob.PtTempoBackend(2, lambda k: None, None, np.ones(4), np.ones(4), 10, 5,
                          1e-6, degeneracy_maps=[np.arange(3), np.arange(4)],
                          ops=HostModelOps())


def test_mean_field_backend_host_logic():
    g = load_golden("mean_field_G")
    influence, props, cfield, cdfield, seen = mean_field_callables(g)
    d2 = int(g["dim"]) ** 2
    be = ob.MeanFieldTempoBackend([g["initial_state"]], complex(g["initial_field"]),
                                  [influence], [g["unitary"]], [props], cfield, cdfield,
                                  [np.ones(d2)], [np.ones(d2)], None, float(g["epsrel"]),
                                  ops=HostModelOps())
    check_mean_field_run(be, g, seen)
    assert be.get_bond_dimensions()[0] == list(g["bond_dims"])


def test_gradient_host_logic():
    """gradient_device (forward + back-propagation + adjoint tensors) against the
    reference's own compute_gradient_and_dynamics on its test J."""
    g = load_golden("gradient_J")
    ops = HostModelOps()
    _, pt, _ = build_pt(load_golden(str(g["pt_fixture"])), ops)

    def props(k):
        return g["props_1"][k], g["props_2"][k]
    derivs, states = ob.gradient_device(pt, props, g["initial_state"],
                                        g["target_derivative"], ops=ops)
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(np.array(derivs), g["propagator_derivatives"], atol=1e-10,
                               rtol=0)
    grad = onp.chain_rule(derivs, lambda k: (g["dprops_1"][k], g["dprops_2"][k]), props,
                          int(g["num_steps"]), 1)
    np.testing.assert_allclose(grad, g["grad_params"], atol=1e-10)
    np.testing.assert_almost_equal(grad.real[:, 0], g["grad_params_golden"], decimal=4)


def test_gradient_two_environments_and_controls_host_logic():
    """gradient_device with two environments, and with pre-/post-measurement controls, against
    the reference's own compute_gradient_and_dynamics (tests/golden/gradient_multi.npz)."""
    from conftest import gradient_multi_setup
    g = load_golden("gradient_multi")
    ops = HostModelOps()
    pts, props, controls = gradient_multi_setup(g, ops)
    n = int(g["num_steps"])
    for e in range(2):          # the caps the reference stored are what compute_caps gives
        for k in range(n + 1):
            np.testing.assert_allclose(pts[e].get_cap_tensor(k), g[f"cap_{e}_{k}"], atol=1e-12)
    derivs, states = ob.gradient_device(pts, props, g["initial_state"], g["target_derivative"],
                                        num_steps=n, ops=ops)
    np.testing.assert_allclose(states, g["states_two_env"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(np.array(derivs), g["derivs_two_env"], atol=1e-10, rtol=0)
    derivs, states = ob.gradient_device(pts[0], props, g["initial_state"],
                                        g["target_derivative"], num_steps=n, ops=ops,
                                        controls=controls)
    np.testing.assert_allclose(states, g["states_controls"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(np.array(derivs), g["derivs_controls"], atol=1e-10, rtol=0)
    # controls through the several-environment code path (one environment listed twice is
    # not the same physics; use [pt] * 1 via the general routine)
    from oqupy_b200.process_tensor import _gradient_multi_env
    derivs, states = _gradient_multi_env([pts[0]], props, g["initial_state"],
                                         g["target_derivative"], n, ops, controls)
    np.testing.assert_allclose(states, g["states_controls"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(np.array(derivs), g["derivs_controls"], atol=1e-10, rtol=0)


def _build_two_pts(g, ops):
    pts = []
    for key in ("influences_a", "influences_b"):
        infl = g[key]
        pt = ob.DeviceProcessTensor(2, dt=float(g["dt"]), ops=ops)
        be = ob.PtTempoBackend(2, lambda dk, infl=infl: None if dk < 0 else infl[dk], pt,
                               np.ones(4), np.ones(4), int(g["num_steps"]),
                               int(g["dkmax"]), float(g["epsrel"]), ops=ops)
        be.initialize()
        while be.compute_step():
            pass
        be.update_process_tensor()
        pts.append(pt)
    return pts


def test_multi_environment_dynamics_host_logic():
    """compute_dynamics with two process tensors (system_dynamics.py:689-700)."""
    g = load_golden("multi_env")
    ops = HostModelOps()
    pts = _build_two_pts(g, ops)
    assert list(pts[0].get_bond_dimensions()) == list(g["bond_dims_a"])
    assert list(pts[1].get_bond_dimensions()) == list(g["bond_dims_b"])
    props = lambda step: (g["prop_1"], g["prop_2"])   # noqa: E731
    states = ob.dynamics_device(pts, props, g["initial_state"], ops=ops)
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)
    swapped = ob.dynamics_device(pts[::-1], props, g["initial_state"], ops=ops)
    np.testing.assert_allclose(swapped, g["states_swapped"], atol=1e-10, rtol=0)


def test_process_tensor_file_round_trip(tmp_path):
    """DeviceProcessTensor.export / import_process_tensor (the side-format for the
    reference's HDF5 FileProcessTensor, process_tensor.py:501-559, 801-823) and export_to a
    host process tensor through the reference's set_mpo_tensor / set_cap_tensor protocol."""
    g = load_golden("pt_k8_eps9_n24")
    ops = HostModelOps()
    _, pt, propagators = build_pt(g, ops)
    path = tmp_path / "pt.b200pt"
    pt.export(str(path))
    back = ob.import_process_tensor(str(path), ops=ops)
    assert len(back) == len(pt) and back.dt == pt.dt
    assert list(back.get_bond_dimensions()) == list(pt.get_bond_dimensions())
    states = ob.dynamics_device(back, propagators, g["initial_state"], ops=ops)
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)

    class HostPt:                     # the two setters of SimpleProcessTensor
        def __init__(self):
            self.mpo, self.cap = {}, {}

        def set_mpo_tensor(self, step, tensor):
            self.mpo[step] = np.array(tensor)

        def set_cap_tensor(self, step, tensor):
            self.cap[step] = np.array(tensor)
    host = pt.export_to(HostPt())
    assert len(host.mpo) == len(pt) and len(host.cap) == len(pt) + 1
    np.testing.assert_array_equal(host.mpo[3], pt.get_mpo_tensor(3))


def test_run_concurrent_background_mode():
    """ensemble._run_concurrent(background=True): returns at once, join() hands back the
    results in member order, alive() reports the threads, thread_init runs once per thread,
    errors of a member surface at join()."""
    import threading
    import time
    from oqupy_b200.ensemble import _run_concurrent
    seen, gate = [], threading.Event()

    def member(i, ops):
        gate.wait(5.0)
        return np.full(3, float(i))

    join = _run_concurrent([4, 7, 9], member, 2, None, background=True,
                           thread_init=lambda ops: seen.append(threading.get_ident()))
    assert join.alive()
    gate.set()
    res = join()
    assert not join.alive()
    assert [float(r[0]) for r in res] == [4.0, 7.0, 9.0]
    assert len(seen) == 2 and len(set(seen)) == 2

    def bad(i, ops):
        raise ValueError(f"member {i}")
    join = _run_concurrent([1], bad, 1, None, background=True)
    time.sleep(0.05)
    with pytest.raises(ValueError):
        join()
    # blocking mode returns the list directly
    assert [float(r[0]) for r in _run_concurrent([2, 3], member, 2, None)] == [2.0, 3.0]


def test_tempo_grid_orchestration_with_a_model_backend(monkeypatch):
    """tempo_grid's host logic (chunked stepping, early / late detection of members that
    leave the lock-step path, their re-run, assembly of the result) against a stand-in for the
    CUDA-only BatchedTempoBackend: every state encodes (member, step), members 2 and 5 leave
    the lock-step path at steps 7 and 33 of 40."""
    import oqupy_b200.batch as batch_mod
    import oqupy_b200.ensemble as ens
    fail_at = {2: 7, 5: 33}
    calls = {"rebalance": 0, "chunks": []}

    class ModelBatch:
        def __init__(self, st0, infl, unitary, props, sn, sw, dkmax, eps, chi_cap=None,
                     ops=None):
            self.E, self.d2, self.step = len(st0), st0.shape[1], 0
            self.members = [int(round(m[0, 0, 0].real)) for m in infl]

        def initialize(self):
            return 0, np.array([[m, 0, 0, 0] for m in self.members], dtype=complex)

        def compute_steps(self, n, strict=True):
            calls["chunks"].append(n)
            out = np.zeros((n, self.E, self.d2), dtype=complex)
            for k in range(n):
                self.step += 1
                for e, m in enumerate(self.members):
                    ok = m not in fail_at or self.step < fail_at[m]
                    out[k, e] = [m, self.step, 0, 0] if ok else np.nan
            return out

        def info(self):
            st = np.array([2 if (m in fail_at and self.step >= fail_at[m]) else 0
                           for m in self.members])
            return {"status": st, "max_chi": np.arange(self.E)}

        def rebalance(self):
            calls["rebalance"] += 1

        def reserve_sms(self, n):
            raise AssertionError("no reservation without a CUDA device")

    def model_member(infl, props, st, dkmax, eps, num_steps, unitary=None, ops=None):
        m = int(round(infl[0, 0, 0].real))
        return np.array([[[m, s], [0, -1]] for s in range(num_steps + 1)], dtype=complex)

    monkeypatch.setattr(batch_mod, "BatchedTempoBackend", ModelBatch)
    monkeypatch.setattr(ens, "tempo_member", model_member)
    n, steps, dkmax = 8, 40, 4
    infl = np.zeros((n, dkmax + 1, 4, 4), dtype=complex)
    infl[:, :, 0, 0] = np.arange(n)[:, None]
    ops = HostModelOps()
    res, rerun = ens.tempo_grid(infl, np.eye(2), np.eye(2), lambda s: (None, None), dkmax,
                                1e-7, steps, ops=ops, check_every=10)
    assert sorted(rerun) == [2, 5]
    assert sum(calls["chunks"]) == steps and calls["rebalance"] == 2
    assert res.shape == (n, steps + 1, 2, 2)
    for m in range(n):
        for s in range(steps + 1):
            if m in fail_at:
                np.testing.assert_array_equal(res[m, s], [[m, s], [0, -1]])   # the re-run
            else:
                np.testing.assert_array_equal(res[m, s].reshape(-1), [m, s, 0, 0])
