"""CPU: the numpy oracle (oracle/tempo_np.py) against the UNMODIFIED reference run LIVE in
this container (oracle/ref_loader.py: /root/reference on top of the restated tensornetwork
slice oracle/tn_shim).  Skipped where the reference tree is absent (the GPU box): there the
committed fixtures the same reference produced (tests/test_oracle_golden.py) stand in.

Also pins, as a TEST instead of prose, the reproducibility floor of the TEMPO path that the
GPU tolerance for TEMPO states is derived from (tests/conftest.py::TEMPO_STATE_ATOL):
perturbing the influence matrices by 1e-15 relative moves the ORACLE's own states by up to a
few 1e-8 once the memory cut-off sets in, and oracle and reference (both LAPACK, different
contraction order) differ by up to a few 1e-6 at epsrel = 1e-7.
"""
import os
import sys

import numpy as np
import pytest

from conftest import TEMPO_STATE_ATOL, golden_callables, load_golden
from oracle import tempo_np as onp

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
from ref_loader import load_reference, reference_available  # noqa: E402

needs_reference = pytest.mark.skipif(not reference_available(),
                                     reason="reference tree not present")


def _tempo_states(g, infl_scale=None):
    influence, propagators = golden_callables(g)
    if infl_scale is not None:
        base = influence

        def influence(dk):          # pylint: disable=function-redefined
            m = base(dk)
            return None if m is None else m * infl_scale[dk]
    d = int(g["dim"])
    d2 = d * d
    dkmax = None if int(g["dkmax"]) < 0 else int(g["dkmax"])
    tb = onp.TempoOracle(g["initial_state"], influence, g["unitary"], propagators,
                         np.ones(d2), np.ones(d2), dkmax, float(g["epsrel"]))
    _, s0 = tb.initialize()
    states = [s0]
    for _ in range(int(g["num_steps"])):
        states.append(tb.compute_step()[1])
    return np.array(states).reshape(-1, d, d), tb.bond_dimensions()


@pytest.mark.parametrize("name", ["tempo_c1_k20_eps7_n60", "tempo_refA"])
def test_tempo_reproducibility_floor(name):
    """The oracle against itself with inputs perturbed at the last bit: the deviation is the
    floor below which no implementation can be held to the reference, and the tolerance the
    GPU TEMPO tests use must sit above it but stay of the order of epsrel."""
    g = load_golden(name)
    eps = float(g["epsrel"])
    rng = np.random.default_rng(0)
    n_infl = len(g["influences"])
    base, bonds0 = _tempo_states(g)
    worst = 0.0
    ties = 0
    for _ in range(3):
        scale = 1.0 + 1e-15 * rng.standard_normal((n_infl, 1, 1))
        pert, bonds1 = _tempo_states(g, scale)
        worst = max(worst, float(np.abs(pert - base).max()))
        ties += sum(a != b for a, b in zip(bonds0, bonds1))
    ref_dev = float(np.abs(base - g["states"]).max())
    print(f"{name}: eps={eps:g}  oracle vs 1e-15-perturbed oracle {worst:.2e} (bond ties {ties}); "
          f"oracle vs reference {ref_dev:.2e}; GPU tolerance {TEMPO_STATE_ATOL(eps):.1e}")
    # the floor is real (far above rounding) ...
    assert worst < TEMPO_STATE_ATOL(eps)
    # ... the oracle-vs-reference deviation sits inside the GPU tolerance too
    assert ref_dev < TEMPO_STATE_ATOL(eps)
    # ... and the tolerance is no looser than a small multiple of the truncation threshold
    assert TEMPO_STATE_ATOL(eps) <= 50 * eps


@needs_reference
def test_pt_tempo_oracle_vs_live_reference():
    """Small PT-TEMPO + compute_dynamics, reference run here, oracle on the same inputs:
    states to 1e-10, equal bond dimensions (tempo_spin_boson_test.py:64-94 geometry)."""
    oqupy = load_reference()
    sx, sz = oqupy.operators.sigma("x"), oqupy.operators.sigma("z")
    system = oqupy.System(0.5 * sx)
    corr = oqupy.PowerLawSD(alpha=0.1, zeta=1, cutoff=4.0, cutoff_type="exponential",
                            temperature=1.6)
    bath = oqupy.Bath(0.5 * sz, corr)
    params = oqupy.TempoParameters(dt=0.1, dkmax=7, epsrel=1e-8)
    rho0 = oqupy.operators.spin_dm("z+")
    pt_obj = oqupy.PtTempo(bath, 0.0, 1.6, params)
    infl = [np.asarray(pt_obj._influence(k), dtype=complex) for k in range(8)]  # pylint: disable=protected-access
    pt = pt_obj.get_process_tensor(progress_type="silent")
    dyn = oqupy.compute_dynamics(system=system, process_tensor=pt, initial_state=rho0,
                                 progress_type="silent")
    p1, p2 = system.get_propagators(params.dt, 0.0, 256, 2 ** -26)(0)
    orc = onp.PtTempoOracle(2, lambda dk: None if dk < 0 else infl[dk], 16, 7, 1e-8)
    orc.compute()
    mpos = orc.mpo_tensors()
    states = onp.compute_dynamics([mpos], [onp.compute_caps(mpos, 2)], lambda s: (p1, p2),
                                  np.asarray(rho0, dtype=complex))
    assert [1] + orc.bond_dimensions() + [1] == list(pt.get_bond_dimensions())
    np.testing.assert_allclose(states, np.array(dyn.states), atol=1e-10, rtol=0)


@needs_reference
def test_tempo_oracle_vs_live_reference():
    oqupy = load_reference()
    sx, sz = oqupy.operators.sigma("x"), oqupy.operators.sigma("z")
    system = oqupy.System(0.5 * sx)
    corr = oqupy.PowerLawSD(alpha=0.1, zeta=1, cutoff=4.0, cutoff_type="exponential",
                            temperature=1.6)
    bath = oqupy.Bath(0.5 * sz, corr)
    params = oqupy.TempoParameters(dt=0.1, dkmax=6, epsrel=1e-7)
    rho0 = oqupy.operators.spin_dm("z+")
    tempo = oqupy.Tempo(system, bath, params, rho0, 0.0)
    infl = [np.asarray(tempo._influence(k), dtype=complex) for k in range(7)]  # pylint: disable=protected-access
    dyn = tempo.compute(1.2, progress_type="silent")
    p1, p2 = system.get_propagators(params.dt, 0.0, 256, 2 ** -26)(0)
    tb = onp.TempoOracle(np.asarray(rho0, dtype=complex), lambda dk: None if dk < 0 else infl[dk],
                         np.eye(2), lambda s: (p1, p2), np.ones(4), np.ones(4), 6, 1e-7)
    _, s0 = tb.initialize()
    ref = np.array(dyn.states)
    states = [s0] + [tb.compute_step()[1] for _ in range(len(ref) - 1)]
    states = np.array(states).reshape(-1, 2, 2)
    assert len(ref) >= 12
    np.testing.assert_allclose(states, ref, atol=TEMPO_STATE_ATOL(1e-7), rtol=0)
    np.testing.assert_allclose(states[:5], ref[:5], atol=1e-9, rtol=0)
