"""GPU: the lock-step TEMPO ensemble (csrc/batch.cu, one CTA per member) against the CPU
oracle member by member, and against the single-run device backend."""
import numpy as np
import pytest

import oqupy_b200 as ob
from conftest import TEMPO_STATE_ATOL, load_golden
from oracle import tempo_np as onp

pytestmark = pytest.mark.gpu


def scaled(infl, factor):
    """Another coupling strength: eta is linear in alpha (oqupy/tempo.py:1008-1015), the
    influence matrices are element-wise powers of each other."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(infl == 0, 0, np.exp(np.log(np.where(infl == 0, 1, infl)) * factor))


def oracle_run(infl, g, dkmax, eps, steps, p1=None):
    d2 = infl.shape[-1]
    tb = onp.TempoOracle(g["initial_state"], lambda dk: None if dk < 0 else infl[dk],
                         g["unitary"], lambda s: (g["prop_1"] if p1 is None else p1, g["prop_2"]),
                         np.ones(d2), np.ones(d2), dkmax, eps)
    _, s0 = tb.initialize()
    out = [s0] + [tb.compute_step()[1] for _ in range(steps)]
    return np.array(out), tb.bond_dimensions()


@pytest.mark.parametrize("dkmax,eps,steps", [(6, 1e-6, 14), (20, 1e-7, 30)])
def test_batched_tempo_matches_oracle(dkmax, eps, steps):
    g = load_golden("tempo_c1_k20_eps7_n60")
    base = g["influences"][:dkmax + 1]
    factors = [1.0, 0.25, 1.7, 3.0, 0.6]
    infl = np.array([scaled(base, f) for f in factors])
    e_ = len(factors)
    d2 = 4
    # members also differ in their propagators (a detuning phase on prop_1)
    p1s = np.array([g["prop_1"] * np.exp(0.02j * k) for k in range(e_)])
    be = ob.BatchedTempoBackend(np.array([g["initial_state"].reshape(-1)] * e_), infl,
                                g["unitary"], lambda s: (p1s, g["prop_2"]), np.ones(d2),
                                np.ones(d2), dkmax, eps)
    _, s0 = be.initialize()
    states = np.concatenate((s0[None], be.compute_steps(steps)))       # (steps+1, E, d2)
    info = be.info()
    assert not info["status"].any()
    for k in range(e_):
        ref, bonds = oracle_run(infl[k], g, dkmax, eps, steps, p1=p1s[k])
        np.testing.assert_allclose(states[:, k], ref, atol=TEMPO_STATE_ATOL(eps), rtol=0)
        np.testing.assert_allclose(states[:4, k], ref[:4], atol=1e-9, rtol=0)
        mine = info["bonds"][k]
        assert len(mine) == len(bonds)
        diff = [abs(a - b) for a, b in zip(mine, bonds)]
        assert max(diff) <= 1 and sum(diff) <= 3, (k, mine, bonds)


def test_batched_tempo_step_by_step_equals_block():
    """compute_step (one D2H per step) and compute_steps give bit-identical states, and
    identical members give bit-identical results (one CTA each, deterministic)."""
    g = load_golden("tempo_c1_k20_eps7_n60")
    infl = np.array([g["influences"][:9]] * 3)
    mk = lambda: ob.BatchedTempoBackend(  # noqa: E731
        np.array([g["initial_state"].reshape(-1)] * 3), infl, g["unitary"],
        lambda s: (g["prop_1"], g["prop_2"]), np.ones(4), np.ones(4), 8, 1e-7)
    a, b = mk(), mk()
    a.initialize()
    b.initialize()
    blk = a.compute_steps(12)
    one = np.array([b.compute_step()[1] for _ in range(12)])
    np.testing.assert_array_equal(blk, one)
    np.testing.assert_array_equal(blk[:, 0], blk[:, 1])
    np.testing.assert_array_equal(blk[:, 0], blk[:, 2])


def test_batched_tempo_launch_order_is_only_a_scheduling_hint():
    """set_order / rebalance (longest members first): bit-identical states and bond dimensions
    whatever CTA takes which member; a non-permutation is rejected."""
    g = load_golden("tempo_c1_k20_eps7_n60")
    base = g["influences"][:9]
    infl = np.array([scaled(base, f) for f in (0.1, 2.0, 0.5, 1.3, 0.02, 0.9)])
    mk = lambda: ob.BatchedTempoBackend(  # noqa: E731
        np.array([g["initial_state"].reshape(-1)] * 6), infl, g["unitary"],
        lambda s: (g["prop_1"], g["prop_2"]), np.ones(4), np.ones(4), 8, 1e-7)
    a, b = mk(), mk()
    a.initialize()
    b.initialize()
    ref = a.compute_steps(14)
    b.set_order([5, 3, 1, 0, 2, 4])
    part1 = b.compute_steps(6)
    order = b.rebalance()
    assert sorted(order.tolist()) == list(range(6))
    chi = b.info()["max_chi"]
    assert all(chi[order[i]] >= chi[order[i + 1]] for i in range(5))
    part2 = b.compute_steps(8)
    np.testing.assert_array_equal(np.concatenate((part1, part2)), ref)
    assert a.get_bond_dimensions() == b.get_bond_dimensions()
    b.set_order(None)
    with pytest.raises(ob.B200Error):
        b.set_order([0, 1, 2, 3, 4, 4])


def test_batched_tempo_capacity_is_reported():
    g = load_golden("tempo_c1_k20_eps7_n60")
    infl = np.array([g["influences"][:21]])
    be = ob.BatchedTempoBackend(g["initial_state"].reshape(1, -1), infl, g["unitary"],
                                lambda s: (g["prop_1"], g["prop_2"]), np.ones(4), np.ones(4),
                                20, 1e-7, chi_cap=8)
    be.initialize()
    with pytest.raises(ob.B200Error):
        be.compute_steps(30)


def test_mean_field_lock_step_matches_oracle():
    """MeanFieldTempoBackend (tempo_backend.py:629-773) with a finite memory cut-off: all
    systems of the model are members of ONE lock-step backend (one launch per time step);
    live field feedback through the host callbacks, against MeanFieldTempoOracle."""
    g = load_golden("tempo_c1_k20_eps7_n60")
    dkmax, eps, steps = 8, 1e-7, 16
    base = g["influences"][:dkmax + 1]
    infls = [scaled(base, f) for f in (1.0, 1.6, 0.4)]
    sz = np.array([1.0, 0.0, 0.0, -1.0])            # <sigma_z> from the vectorised state

    def make_props(k):
        def props(step, field, dfield):
            ph = np.exp(-0.05j * (k + 1) * (field.real + 0.5 * dfield.real))
            return g["prop_1"] * ph, g["prop_2"] * np.conj(ph)
        return props

    def dfield(step, states, field):
        return -0.3 * field + 0.1 * sum(complex(sz @ s) for s in states)

    def cfield(step, states, field, next_states):
        return field + 0.05 * (dfield(step, states, field) + dfield(step, next_states, field)) / 2

    args = ([g["initial_state"]] * 3, 0.2 + 0.1j,
            [lambda dk, m=m: None if dk < 0 else m[dk] for m in infls],
            [g["unitary"]] * 3, [make_props(k) for k in range(3)], cfield, dfield,
            [np.ones(4)] * 3, [np.ones(4)] * 3, dkmax, eps)
    be = ob.MeanFieldTempoBackend(*args)
    assert be._batched is not None                      # the lock-step path is taken
    orc = onp.MeanFieldTempoOracle(*args)
    be.initialize()
    orc.initialize()
    for k in range(steps):
        s1, st1, f1 = be.compute_step()
        s2, st2, f2 = orc.compute_step()
        assert s1 == s2 == k + 1
        tol = 1e-9 if k < 4 else TEMPO_STATE_ATOL(eps)
        np.testing.assert_allclose(np.array(st1), np.array(st2), atol=tol, rtol=0)
        assert abs(f1 - f2) < tol
    for mine, ref in zip(be.get_bond_dimensions(), [n.bond_dimensions() for n in orc.networks]):
        diff = [abs(a - b) for a, b in zip(mine, ref)]
        assert len(mine) == len(ref) and max(diff) <= 1 and sum(diff) <= 3


def test_tempo_grid_general_path_fallback():
    """tempo_grid: members whose bond dimension outgrows the lock-step capacity are re-run on
    the general device backend WHILE the others carry on; the result is what a single run of
    that member gives, the others are untouched."""
    from oqupy_b200.ensemble import tempo_grid, tempo_member
    g = load_golden("tempo_c1_k20_eps7_n60")
    dkmax, eps, steps = 6, 1e-7, 24
    base = g["influences"][:dkmax + 1]
    factors = [0.05, 1.0, 0.02, 2.5, 0.03]
    infl = np.array([scaled(base, f) for f in factors])
    props = lambda s: (g["prop_1"], g["prop_2"])  # noqa: E731
    probe = ob.BatchedTempoBackend(np.array([g["initial_state"].reshape(-1)] * len(factors)),
                                   infl, g["unitary"], props, np.ones(4), np.ones(4), dkmax, eps)
    probe.initialize()
    probe.compute_steps(steps)
    max_chi = probe.info()["max_chi"]
    assert max_chi.max() > max_chi.min()
    cap = int(max_chi.min())           # the weakest coupling just fits, the strongest does not
    expect = {k for k in range(len(factors)) if max_chi[k] > cap}
    res, rerun = tempo_grid(infl, g["initial_state"], g["unitary"], props, dkmax, eps, steps,
                            chi_cap=cap)
    assert res.shape == (len(factors), steps + 1, 2, 2)
    assert set(rerun) == expect and 0 < len(expect) < len(factors), (rerun, max_chi)
    full, none = tempo_grid(infl, g["initial_state"], g["unitary"], props, dkmax, eps, steps)
    assert not none
    for k in range(len(factors)):
        if k in rerun:
            one = tempo_member(infl[k], props, g["initial_state"], dkmax, eps, steps,
                               unitary=g["unitary"])
            np.testing.assert_array_equal(res[k], one)
            np.testing.assert_allclose(res[k], full[k], atol=TEMPO_STATE_ATOL(eps), rtol=0)
        else:
            np.testing.assert_array_equal(res[k], full[k])
