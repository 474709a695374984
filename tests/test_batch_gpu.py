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


def test_batched_tempo_capacity_is_reported():
    g = load_golden("tempo_c1_k20_eps7_n60")
    infl = np.array([g["influences"][:21]])
    be = ob.BatchedTempoBackend(g["initial_state"].reshape(1, -1), infl, g["unitary"],
                                lambda s: (g["prop_1"], g["prop_2"]), np.ones(4), np.ones(4),
                                20, 1e-7, chi_cap=8)
    be.initialize()
    with pytest.raises(ob.B200Error):
        be.compute_steps(30)
