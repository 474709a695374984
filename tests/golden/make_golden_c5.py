"""Generate tests/golden/c5_grid_operands.npz: the influence matrices of the config-5 sweep
(BASELINE configs[4]: TEMPO runs over a coupling x temperature grid; SURVEY 8d: ohmic,
wc=4, dt=0.05, dkmax=20, epsrel=1e-7, alpha in linspace(0.02, 0.30, 64), T in
linspace(0.2, 3.2, 64)).  The UNMODIFIED reference (oqupy.Tempo._influence ->
oqupy/tempo.py:969-1020) is run for the 64 temperatures at alpha_ref = 0.08; the coupling
axis needs no further reference runs: eta is linear in alpha, so the influence matrices of
another alpha are element-wise powers alpha/alpha_ref of these (checked below for one pair).

    python tests/golden/make_golden_c5.py        (build container only, ~2 min)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from ref_loader import load_reference  # noqa: E402

oqupy = load_reference()
ALPHA_REF, K, DT, EPS = 0.08, 20, 0.05, 1e-7


def influences(alpha, temperature):
    corr = oqupy.PowerLawSD(alpha=alpha, zeta=1, cutoff=4.0, cutoff_type="exponential",
                            temperature=temperature)
    bath = oqupy.Bath(0.5 * oqupy.operators.sigma("z"), corr)
    system = oqupy.System(0.5 * oqupy.operators.sigma("x"))
    params = oqupy.TempoParameters(dt=DT, dkmax=K, epsrel=EPS)
    tempo = oqupy.Tempo(system, bath, params, oqupy.operators.spin_dm("z+"), 0.0)
    return np.array([np.asarray(tempo._influence(k), dtype=complex) for k in range(K + 1)])  # pylint: disable=protected-access


def main():
    temps = np.linspace(0.2, 3.2, 64)
    infl = np.array([influences(ALPHA_REF, t) for t in temps])
    # the coupling axis: element-wise power
    other = influences(0.2, temps[10])
    with np.errstate(divide="ignore", invalid="ignore"):
        power = np.where(infl[10] == 0, 0, np.exp(np.log(np.where(infl[10] == 0, 1, infl[10]))
                                                   * (0.2 / ALPHA_REF)))
    err = float(np.abs(power - other).max())
    print("alpha scaling check, max abs deviation:", err)
    assert err < 1e-12
    system = oqupy.System(0.5 * oqupy.operators.sigma("x"))
    p1, p2 = system.get_propagators(DT, 0.0, 256, 2 ** -26)(0)
    np.savez_compressed(os.path.join(HERE, "c5_grid_operands.npz"), influences=infl,
                        temperatures=temps, alpha_ref=ALPHA_REF, alphas=np.linspace(0.02, 0.30, 64),
                        dkmax=K, dt=DT, epsrel=EPS, prop_1=p1, prop_2=p2,
                        initial_state=np.asarray(oqupy.operators.spin_dm("z+"), dtype=complex),
                        unitary=np.identity(2, dtype=complex))


if __name__ == "__main__":
    main()
