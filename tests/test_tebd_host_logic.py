"""CPU: host-side index logic of the device PT-TEBD backend (oqupy_b200/tebd.py) on the
test-only strided-view model of the C-ABI, against the reference fixtures of its test F."""
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import load_golden, tebd_fixture
from host_model_ops import HostModelOps
from test_tebd_oracle import check_tebd
import oqupy_b200 as ob


class HostPt:
    """A plain host process tensor returning the public 4-leg site
    (process_tensor.py:326-355)."""

    def __init__(self, mpos, caps):
        self.mpos, self.caps = mpos, caps

    def get_mpo_tensor(self, step):
        t = self.mpos[step]
        return np.einsum("bcp,pq->bcpq", t, np.identity(t.shape[2]))

    def get_cap_tensor(self, step):
        return self.caps[step]


class TrivialPt:
    """TrivialProcessTensor (process_tensor.py:199-250)."""

    def get_mpo_tensor(self, step):
        return None

    def get_cap_tensor(self, step):
        return np.array([1.0], dtype=complex)


def run_tebd_backend(g, ops, device_pt):
    """PtTebd.compute_step loop (oqupy/pt_tebd.py:408-419) on the device backend."""
    gammas, lambdas, layers, pt_sites, mpos, caps = tebd_fixture(g)
    n = int(g["n"])
    be = ob.PtTebdBackend(gammas, lambdas, float(g["epsrel"]), {}, ops=ops)
    gate_layers = [SimpleNamespace(gates=[SimpleNamespace(sites=list(s), tensors=list(t))
                                          for s, t in layer]) for layer in layers]
    pt = None
    if pt_sites:
        if device_pt:
            pt = ob.DeviceProcessTensor(2, dt=float(g["dt"]), ops=ops)
            for k, t in enumerate(mpos):
                pt.set_mpo_tensor(k, t)
            pt.compute_caps()
            np.testing.assert_allclose(pt.get_cap_tensor(0), caps[0], atol=1e-12)
        else:
            pt = HostPt(mpos, caps)
    pts = [pt if s in pt_sites else TrivialPt() for s in range(n)]
    states, states_13, norms, bonds = [], [], [], []

    def record(step):
        be.compute_traces(step, pts)
        states.append([be.get_density_matrix([s]) for s in range(n)])
        states_13.append(be.get_density_matrix([1, 3]))
        norms.append(be.get_norm())
        bonds.append(list(be.get_bond_dimensions()))
        be.clear_traces()

    record(0)
    for step in range(1, int(g["num_steps"]) + 1):
        for layer in gate_layers:
            be.apply_nn_gate_layer(layer)
        be.apply_process_tensors(step, pts)
        for layer in gate_layers:
            be.apply_nn_gate_layer(layer)
        record(step)
    return be, (np.array(states).transpose(1, 0, 2, 3), np.array(states_13),
                np.array(norms), np.array(bonds))


@pytest.mark.parametrize("tag,device_pt", [("F1", False), ("F2", False), ("F2", True)])
def test_tebd_backend_host_logic(tag, device_pt):
    g = load_golden(f"pt_tebd_{tag}")
    be, res = run_tebd_backend(g, HostModelOps(), device_pt)
    check_tebd(res, g, 1e-9)
    lam = be.get_lambda(1)
    assert lam.shape == (res[3][-1][1],) * 2 and np.allclose(lam, np.diag(np.diag(lam)))
    assert be.get_gamma(2).shape[0] == res[3][-1][1]


def test_tebd_site_gate_host_logic():
    g = load_golden("pt_tebd_F1")
    gammas, lambdas, _, _, _, _ = tebd_fixture(g)
    ops = HostModelOps()
    be = ob.PtTebdBackend(gammas, lambdas, 1e-7, {}, ops=ops)
    rng = np.random.default_rng(1)
    mat = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    be.apply_site_gate_layer(SimpleNamespace(gates=[SimpleNamespace(sites=[2], tensors=[mat])]))
    np.testing.assert_allclose(be.get_gamma(2),
                               np.einsum("ap,lpbr->labr", mat, gammas[2]), atol=1e-14)
