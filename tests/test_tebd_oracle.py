"""CPU: the numpy PT-TEBD oracle (oracle/tebd_np.py) against fixtures produced by the
UNMODIFIED reference on its own test F (tests/physics/pt_tebd_test.py) and against the
reference's golden matrices example_F{1,2}_rhos.npy."""
import numpy as np
import pytest

from conftest import load_golden, tebd_fixture
from oracle import tebd_np


def run_tebd(backend_factory, g, apply_layer, traces):
    """PtTebd.compute_step loop (oqupy/pt_tebd.py:408-419) on a fixture."""
    gammas, lambdas, layers, pt_sites, mpos, caps = tebd_fixture(g)
    n = int(g["n"])
    be = backend_factory(gammas, lambdas, float(g["epsrel"]))
    one = np.array([1.0], dtype=complex)
    states, states_13, norms, bonds = [], [], [], []

    def record(step):
        traces(be, [caps[step] if s in pt_sites else one for s in range(n)])
        states.append([be.get_density_matrix([s]) for s in range(n)])
        states_13.append(be.get_density_matrix([1, 3]))
        norms.append(be.get_norm())
        bonds.append(list(be.get_bond_dimensions()))

    record(0)
    for step in range(1, int(g["num_steps"]) + 1):
        for layer in layers:
            apply_layer(be, layer)
        be.apply_process_tensors(step, [mpos[step - 1] if s in pt_sites else None
                                        for s in range(n)])
        for layer in layers:
            apply_layer(be, layer)
        record(step)
    return (np.array(states).transpose(1, 0, 2, 3), np.array(states_13), np.array(norms),
            np.array(bonds))


def check_tebd(result, g, atol):
    states, states_13, norms, bonds = result
    assert bonds.tolist() == g["bond_dims"].tolist()
    np.testing.assert_allclose(states, g["states"], atol=atol, rtol=0)
    np.testing.assert_allclose(states_13, g["states_13"], atol=atol, rtol=0)
    np.testing.assert_allclose(norms, g["norm"], atol=atol, rtol=0)
    # the reference's own pin (pt_tebd_test.py:112-116, 134-138)
    np.testing.assert_almost_equal(states[:, -1], g["rho_golden"], decimal=4)


@pytest.mark.parametrize("tag", ["F1", "F2"])
def test_tebd_oracle_matches_reference(tag):
    g = load_golden(f"pt_tebd_{tag}")
    res = run_tebd(tebd_np.PtTebdOracle, g, lambda be, layer: be.apply_nn_gate_layer(layer),
                   lambda be, caps: be.compute_traces(caps))
    check_tebd(res, g, 1e-9)
