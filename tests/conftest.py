"""pytest configuration: the ``gpu`` marker and shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def TEMPO_STATE_ATOL(eps):  # pylint: disable=invalid-name
    """Tolerance of TEMPO states against the reference: the reproducibility floor of the
    reference ALGORITHM, measured by tests/test_oracle_vs_reference.py (inputs perturbed by
    1e-15 move the states by up to ~0.5 eps; oracle and reference, both LAPACK, differ by up
    to ~20 eps once the memory cut-off sets in)."""
    return 50.0 * eps


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def golden_callables(g):
    """(influence(dk), propagators(step)) callables from a golden fixture."""
    infl = g["influences"]

    def influence(dk):
        if dk < 0:
            return None
        return infl[dk]

    p1, p2 = g["prop_1"], g["prop_2"]

    def propagators(step):
        return p1, p2

    return influence, propagators


def mean_field_callables(g):
    """Replay of the host callbacks of the reference's mean-field run: the propagator
    pair and the field it used at every step (tests/golden/make_golden_mean_field.py).
    Returns (influence, propagators, compute_field, compute_field_derivative, seen)
    where ``seen`` collects the (field, field derivative) values handed to propagators."""
    infl = g["influences"]
    seen = []

    def influence(dk):
        return None if dk < 0 else infl[dk]

    def propagators(step, field, dfield):
        seen.append((step, field, dfield))
        return g["props_1"][step], g["props_2"][step]

    def compute_field(step, states, field, next_states=None):
        return complex(g["fields"][step + 1])

    def compute_field_derivative(step, states, field):
        return complex(g["dfields_in"][step])

    return influence, propagators, compute_field, compute_field_derivative, seen


def check_mean_field_run(backend, g, seen):
    """Drive a MeanFieldTempoBackend-like object through the fixture and compare."""
    d = int(g["dim"])
    step, states, field = backend.initialize()
    assert step == 0 and complex(field) == complex(g["initial_field"])
    out = [np.asarray(states[0]).reshape(d, d)]
    for k in range(int(g["num_steps"])):
        step, states, field = backend.compute_step()
        assert step == k + 1
        assert complex(field) == complex(g["fields"][k + 1])
        out.append(np.asarray(states[0]).reshape(d, d))
    out = np.array(out)
    np.testing.assert_allclose(out, g["states"], atol=TEMPO_STATE_ATOL(float(g["epsrel"])),
                               rtol=0)
    np.testing.assert_allclose(out[:5], g["states"][:5], atol=1e-9, rtol=0)
    np.testing.assert_almost_equal(out[-1], g["rho_golden"], decimal=4)
    # plumbing: every step's propagators saw the field of that step and its derivative
    assert [s[0] for s in seen] == list(range(int(g["num_steps"])))
    np.testing.assert_array_equal(np.array([s[1] for s in seen]), g["fields_in"])
    np.testing.assert_array_equal(np.array([s[2] for s in seen]), g["dfields_in"])
    return out


def unique_callables(g):
    """(influence(dk), propagators(step)) from a ``*_unique_*`` fixture: dk = 0 gives the
    n_north reduced values, dk > 0 the reduced (n_north, n_west) matrices
    (oqupy/tempo.py:969-1020 with the degeneracy maps of oqupy/bath.py:87-89)."""
    infl0, infl = g["influence_0"], g["influences"]

    def influence(dk):
        if dk < 0:
            return None
        return infl0 if dk == 0 else infl[dk - 1]

    p1, p2 = g["prop_1"], g["prop_2"]
    return influence, (lambda step: (p1, p2))


def tebd_fixture(g):
    """(gammas, lambdas, gate layers [[(sites, (gate_l, gate_r))]], pt sites / caps per
    chain site or None) from a ``pt_tebd_F*`` fixture (tests/golden/make_golden_tebd.py)."""
    n = int(g["n"])
    gammas = [g[f"gamma_{i}"] for i in range(n)]
    lambdas = [g[f"lambda_{i}"] for i in range(n - 1)]
    layers = []
    for li, size in enumerate(g["layer_sizes"]):
        layers.append([(tuple(int(x) for x in g[f"gate_{li}_{gi}_sites"]),
                        (g[f"gate_{li}_{gi}_l"], g[f"gate_{li}_{gi}_r"]))
                       for gi in range(int(size))])
    pt_sites = set(int(x) for x in g["pt_sites"])
    mpos = caps = None
    if pt_sites:
        mpos = [g[f"pt_mpo_{k}"] for k in range(int(g["pt_len"]))]
        caps = [g[f"pt_cap_{k}"] for k in range(int(g["pt_len"]) + 1)]
    return gammas, lambdas, layers, pt_sites, mpos, caps


def gradient_multi_setup(g, ops):
    """Process tensors (device objects on `ops`), propagators and controls of
    tests/golden/gradient_multi.npz (make_golden_gradient_multi.py)."""
    import oqupy_b200 as ob  # pylint: disable=import-outside-toplevel
    n = int(g["num_steps"])
    pts = []
    for e in range(2):
        pt = ob.DeviceProcessTensor(int(g["dim"]), dt=float(g["dt"]), ops=ops)
        for k in range(n):
            pt.set_mpo_tensor(k, g[f"mpo_{e}_{k}"])
        pt.compute_caps()
        pts.append(pt)

    def props(k):
        return g["props_1"][k], g["props_2"][k]
    controls = [(g["controls"][k, 0] if g["has_control"][k, 0] else None,
                 g["controls"][k, 1] if g["has_control"][k, 1] else None) for k in range(n + 1)]
    return pts, props, controls
