"""CPU: the numpy oracle (oracle/tempo_np.py) against fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py) and against the golden density
matrices quoted from the reference's own tests."""
import numpy as np
import pytest

from conftest import (check_mean_field_run, golden_callables, load_golden,
                      mean_field_callables, unique_callables)
from oracle import tempo_np as onp

PT_CASES = ["pt_k12_eps7_n30", "pt_k8_eps9_n24", "pt_refA", "pt_refC"]
TEMPO_CASES = ["tempo_c1_k20_eps7_n60", "tempo_refA", "tempo_refC",
               "tempo_nondiag"]


def run_pt_oracle(g):
    influence, propagators = golden_callables(g)
    pt = onp.PtTempoOracle(int(g["dim"]), influence, int(g["num_steps"]),
                           int(g["dkmax"]), float(g["epsrel"]))
    pt.compute()
    mpos = pt.mpo_tensors()
    caps = onp.compute_caps(mpos, int(g["dim"]))
    states = onp.compute_dynamics([mpos], [caps], propagators,
                                  g["initial_state"])
    return pt, mpos, caps, states


@pytest.mark.parametrize("name", PT_CASES)
def test_pt_oracle_matches_reference(name):
    g = load_golden(name)
    pt, mpos, caps, states = run_pt_oracle(g)
    bonds = [1] + pt.bond_dimensions() + [1]
    assert bonds == list(g["bond_dims"])
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(caps[0], g["cap_first"], atol=1e-10)
    if "rho_golden" in g:   # the reference's own decimal=4 pin
        np.testing.assert_almost_equal(states[-1], g["rho_golden"], decimal=4)


@pytest.mark.parametrize("name", TEMPO_CASES)
def test_tempo_oracle_matches_reference(name):
    g = load_golden(name)
    influence, propagators = golden_callables(g)
    d2 = int(g["dim"]) ** 2
    dkmax = None if int(g["dkmax"]) < 0 else int(g["dkmax"])
    tb = onp.TempoOracle(g["initial_state"], influence, g["unitary"],
                         propagators, np.ones(d2), np.ones(d2), dkmax,
                         float(g["epsrel"]))
    _, s0 = tb.initialize()
    states = [s0]
    for _ in range(int(g["num_steps"])):
        states.append(tb.compute_step()[1])
    d = int(g["dim"])
    states = np.array(states).reshape(-1, d, d)
    assert tb.bond_dimensions() == list(g["bond_dims"])
    # TEMPO amplifies 1e-16 rounding differences (contraction order) to ~10*eps
    # once the memory cut-off sets in: perturbing the influence matrices of the
    # REFERENCE by 1e-15 relative moves its own states by 4e-8 (DESIGN.md,
    # "reproducibility floor").  Hence eps-scaled tolerance + exact bond dims.
    np.testing.assert_allclose(states, g["states"],
                               atol=50 * float(g["epsrel"]), rtol=0)
    k = min(5, len(states))
    np.testing.assert_allclose(states[:k], g["states"][:k], atol=1e-9, rtol=0)
    if "rho_golden" in g:
        np.testing.assert_almost_equal(states[-1], g["rho_golden"], decimal=4)


def test_truncation_rule_matches_in_repo_pin():
    """oqupy/mps_mpo.py:452-457: chi = count_nonzero(cumsum(flip(s)^2) > (s0 eps)^2)."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        s = np.sort(10.0 ** rng.uniform(-12, 0, size=40))[::-1]
        q1, _ = np.linalg.qr(rng.normal(size=(60, 40)) + 1j * rng.normal(size=(60, 40)))
        q2, _ = np.linalg.qr(rng.normal(size=(40, 40)) + 1j * rng.normal(size=(40, 40)))
        mat = (q1 * s) @ q2
        eps = 10.0 ** rng.uniform(-9, -3)
        chi = np.count_nonzero(np.cumsum(np.flip(s) ** 2) > (s[0] * eps) ** 2)
        u, sk, vh, rest = onp.truncated_svd(mat, eps)
        assert sk.size == chi and rest.size == 40 - chi


def test_mean_field_oracle_matches_reference():
    """MeanFieldTempoBackend (tempo_backend.py:629-773) replayed with the host callbacks
    of the reference's own run of its test G (mean_field_tempo_test.py:21-69)."""
    g = load_golden("mean_field_G")
    influence, props, cfield, cdfield, seen = mean_field_callables(g)
    d2 = int(g["dim"]) ** 2
    mf = onp.MeanFieldTempoOracle([g["initial_state"]], complex(g["initial_field"]),
                                  [influence], [g["unitary"]], [props], cfield, cdfield,
                                  [np.ones(d2)], [np.ones(d2)], None, float(g["epsrel"]))
    check_mean_field_run(mf, g, seen)
    assert mf.networks[0].bond_dimensions() == list(g["bond_dims"])
    np.testing.assert_almost_equal(complex(g["fields"][-1]), complex(g["field_golden"]),
                                   decimal=4)


def test_gradient_oracle_matches_reference():
    """compute_gradient_and_dynamics + _chain_rule (gradient.py:114-437) restated in the
    oracle against the reference's own run of its test J (gradient_target_state_test.py)."""
    g = load_golden("gradient_J")
    ga = load_golden(str(g["pt_fixture"]))
    pt, mpos, caps, _ = run_pt_oracle(ga)

    def props(k):
        return g["props_1"][k], g["props_2"][k]
    derivs, states = onp.compute_gradient_and_dynamics(
        mpos, caps, props, g["initial_state"], g["target_derivative"])
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(np.array(derivs), g["propagator_derivatives"], atol=1e-10,
                               rtol=0)
    grad = onp.chain_rule(derivs, lambda k: (g["dprops_1"][k], g["dprops_2"][k]), props,
                          int(g["num_steps"]), 1)
    np.testing.assert_allclose(grad, g["grad_params"], atol=1e-12)
    # the reference test's own pin (grad_params_J, decimal=4)
    np.testing.assert_almost_equal(grad.real[:, 0], g["grad_params_golden"], decimal=4)


def test_multi_environment_oracle_matches_reference():
    """compute_dynamics with two process tensors (system_dynamics.py:689-700; the
    setting of tests/physics/multi_environments_test.py with two different baths)."""
    g = load_golden("multi_env")
    mpos, caps = [], []
    for key, bkey in (("influences_a", "bond_dims_a"), ("influences_b", "bond_dims_b")):
        infl = g[key]
        pt = onp.PtTempoOracle(2, lambda dk, infl=infl: None if dk < 0 else infl[dk],
                               int(g["num_steps"]), int(g["dkmax"]), float(g["epsrel"]))
        pt.compute()
        assert [1] + pt.bond_dimensions() + [1] == list(g[bkey])
        mpos.append(pt.mpo_tensors())
        caps.append(onp.compute_caps(mpos[-1], 2))
    props = lambda step: (g["prop_1"], g["prop_2"])   # noqa: E731
    states = onp.compute_dynamics(mpos, caps, props, g["initial_state"])
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)


UNIQUE_TAGS = ["spin12", "spin1"]


@pytest.mark.parametrize("tag", UNIQUE_TAGS)
def test_pt_oracle_unique(tag):
    """unique=True (pt_tempo_backend.py:114-140): reduced north / west legs."""
    g = load_golden(f"pt_unique_{tag}")
    influence, propagators = unique_callables(g)
    maps = [g["north_map"], g["west_map"]]
    pt = onp.PtTempoOracle(int(g["dim"]), influence, int(g["num_steps"]), int(g["dkmax"]),
                           float(g["epsrel"]), sum_north=np.ones(int(maps[0].max()) + 1),
                           degeneracy_maps=maps)
    pt.compute()
    mpos = pt.mpo_tensors()
    caps = onp.compute_caps(mpos, int(g["dim"]))
    states = onp.compute_dynamics([mpos], [caps], propagators, g["initial_state"])
    assert [1] + pt.bond_dimensions() + [1] == list(g["bond_dims"])
    np.testing.assert_allclose(states, g["states"], atol=1e-10, rtol=0)


@pytest.mark.parametrize("tag", UNIQUE_TAGS)
def test_tempo_oracle_unique(tag):
    """unique=True (tempo_backend.py:400-417)."""
    g = load_golden(f"tempo_unique_{tag}")
    influence, propagators = unique_callables(g)
    maps = [g["north_map"], g["west_map"]]
    nn, nw = int(maps[0].max()) + 1, int(maps[1].max()) + 1
    tb = onp.TempoOracle(g["initial_state"], influence, g["unitary"], propagators,
                         np.ones(nn), np.ones(nw), int(g["dkmax"]), float(g["epsrel"]),
                         degeneracy_maps=maps)
    _, s0 = tb.initialize()
    states = [s0]
    for _ in range(int(g["num_steps"])):
        states.append(tb.compute_step()[1])
    d = int(g["dim"])
    states = np.array(states).reshape(-1, d, d)
    assert tb.bond_dimensions() == list(g["bond_dims"])
    # spin-1 truncates from step 3 on (d2 = 9): 2.5e-8 there, the TEMPO reproducibility
    # floor at eps = 1e-7 (see test_tempo_oracle_matches_reference); exact before
    np.testing.assert_allclose(states, g["states"], atol=50 * float(g["epsrel"]), rtol=0)
    np.testing.assert_allclose(states[:3], g["states"][:3], atol=1e-9, rtol=0)


def test_gradient_oracle_two_environments_and_controls():
    """The oracle's general gradient (several environments, controls) against the reference's
    own compute_gradient_and_dynamics (tests/golden/gradient_multi.npz)."""
    g = load_golden("gradient_multi")
    n = int(g["num_steps"])
    mpos = [[g[f"mpo_{e}_{k}"] for k in range(n)] for e in range(2)]
    caps = [[g[f"cap_{e}_{k}"] for k in range(n + 1)] for e in range(2)]
    props = lambda k: (g["props_1"][k], g["props_2"][k])   # noqa: E731
    derivs, states = onp.compute_gradient_and_dynamics_general(
        mpos, caps, props, g["initial_state"], g["target_derivative"], n)
    np.testing.assert_allclose(states, g["states_two_env"], atol=1e-12, rtol=0)
    np.testing.assert_allclose(np.array(derivs), g["derivs_two_env"], atol=1e-12, rtol=0)
    controls = [(g["controls"][k, 0] if g["has_control"][k, 0] else None,
                 g["controls"][k, 1] if g["has_control"][k, 1] else None) for k in range(n + 1)]
    derivs, states = onp.compute_gradient_and_dynamics_general(
        mpos[:1], caps[:1], props, g["initial_state"], g["target_derivative"], n,
        controls=controls)
    np.testing.assert_allclose(states, g["states_controls"], atol=1e-12, rtol=0)
    np.testing.assert_allclose(np.array(derivs), g["derivs_controls"], atol=1e-12, rtol=0)
