"""Generate tests/golden/multi_env.npz by running the UNMODIFIED reference's
compute_dynamics with TWO process tensors (system_dynamics.py:41-182, 689-700; the setting
of tests/physics/multi_environments_test.py:20-68, with two DIFFERENT baths so that the
order of the bond legs matters).  Build-container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from ref_loader import load_reference  # noqa: E402

oqupy = load_reference()


def main():
    sx, sz = oqupy.operators.sigma("x"), oqupy.operators.sigma("z")
    system = oqupy.System(sx)
    rho0 = oqupy.operators.spin_dm("z+")
    params = oqupy.TempoParameters(dt=0.1, dkmax=5, epsrel=1e-6)
    baths = [oqupy.Bath(0.5 * sz, oqupy.PowerLawSD(alpha=a, zeta=1.0, cutoff=5.0,
                                                    cutoff_type="exponential",
                                                    temperature=t))
             for a, t in ((0.05, 0.2), (0.1, 0.6))]
    pts, infl = [], []
    for bath in baths:
        obj = oqupy.PtTempo(bath, 0.0, 1.0, params)
        infl.append(np.array([np.asarray(obj._influence(k), dtype=complex)
                              for k in range(6)]))
        obj.compute(progress_type="silent")
        pts.append(obj.get_process_tensor(progress_type="silent"))
    dyn = oqupy.compute_dynamics(system, process_tensor=pts, initial_state=rho0,
                                 progress_type="silent")
    dyn_swapped = oqupy.compute_dynamics(system, process_tensor=pts[::-1],
                                         initial_state=rho0, progress_type="silent")
    p1, p2 = system.get_propagators(0.1, 0.0, 256, 2 ** -26)(0)
    np.savez_compressed(
        os.path.join(HERE, "multi_env.npz"), kind="multi_env", dim=2, dt=0.1, dkmax=5,
        epsrel=1e-6, num_steps=10, influences_a=infl[0], influences_b=infl[1],
        prop_1=p1, prop_2=p2, initial_state=np.asarray(rho0, dtype=complex),
        states=np.array(dyn.states), states_swapped=np.array(dyn_swapped.states),
        bond_dims_a=pts[0].get_bond_dimensions(), bond_dims_b=pts[1].get_bond_dimensions())
    print("multi_env: bonds", list(pts[0].get_bond_dimensions()),
          list(pts[1].get_bond_dimensions()), "order dependence",
          np.abs(np.array(dyn.states) - np.array(dyn_swapped.states)).max())


if __name__ == "__main__":
    main()
