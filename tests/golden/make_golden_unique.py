"""Generate tests/golden/{pt,tempo}_unique_*.npz by running the UNMODIFIED reference with
``unique=True`` (degeneracy reduction of the north / west legs: oqupy/bath.py:87-89,
tempo_backend.py:400-417, pt_tempo_backend.py:114-140).  Build-container only.

The fixtures store what crosses the backend boundary in this mode: the REDUCED influence
matrices (n_north x n_west; dk = 0: n_north values), the degeneracy maps, propagators and
initial state, next to the reference's outputs.  The reference's own check is that
unique=True reproduces unique=False (tests/physics/degeneracy_test.py); the generator
asserts that too."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from ref_loader import load_reference  # noqa: E402

oqupy = load_reference()


def main():
    sx, sz = oqupy.operators.sigma("x"), oqupy.operators.sigma("z")
    up = oqupy.operators.spin_dm("z+")
    corr = oqupy.PowerLawSD(alpha=0.1, zeta=1, cutoff=4.0, cutoff_type="exponential",
                            temperature=0.5)
    # spin-1 with S_z coupling: d2 = 9, 5 distinct commutator values (west), 9 north
    szz = np.diag([1.0, 0.0, -1.0])
    sxx = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]]) / np.sqrt(2.0)
    for tag, op, hs, rho0 in (("spin12", 0.5 * sz, 0.5 * sx, up),
                              ("spin1", szz, sxx + 0.3 * szz, np.diag([1.0, 0.0, 0.0]))):
        bath = oqupy.Bath(op, corr)
        system = oqupy.System(hs)
        nmap, wmap = bath.north_degeneracy_map, bath.west_degeneracy_map
        n_north, n_west = int(nmap.max()) + 1, int(wmap.max()) + 1
        d = bath.dimension
        # ---- PT-TEMPO
        params = oqupy.TempoParameters(dt=0.1, dkmax=6, epsrel=1e-7)
        n_steps = 14
        obj = oqupy.PtTempo(bath, 0.0, n_steps * 0.1, params, unique=True)
        infl = [np.asarray(obj._influence(k), dtype=complex) for k in range(7)]
        obj.compute(progress_type="silent")
        pt = obj.get_process_tensor(progress_type="silent")
        dyn = oqupy.compute_dynamics(system=system, process_tensor=pt, initial_state=rho0,
                                     progress_type="silent")
        ref = oqupy.PtTempo(bath, 0.0, n_steps * 0.1, params, unique=False)
        ref.compute(progress_type="silent")
        dyn_ref = oqupy.compute_dynamics(
            system=system, process_tensor=ref.get_process_tensor(progress_type="silent"),
            initial_state=rho0, progress_type="silent")
        assert np.abs(np.array(dyn.states) - np.array(dyn_ref.states)).max() < 1e-5
        p1, p2 = system.get_propagators(0.1, 0.0, 256, 2 ** -26)(0)
        np.savez_compressed(
            os.path.join(HERE, f"pt_unique_{tag}.npz"), kind="pt_unique", dim=d, dt=0.1,
            dkmax=6, epsrel=1e-7, num_steps=n_steps, influence_0=infl[0],
            influences=np.array(infl[1:]), north_map=nmap, west_map=wmap,
            prop_1=p1, prop_2=p2, initial_state=np.asarray(rho0, dtype=complex),
            states=np.array(dyn.states), bond_dims=pt.get_bond_dimensions())
        print(f"pt_unique_{tag}: n_north {n_north} n_west {n_west} d2 {d * d} bonds",
              list(pt.get_bond_dimensions()))
        # ---- TEMPO
        tparams = oqupy.TempoParameters(dt=0.1, dkmax=6, epsrel=1e-7)
        tempo = oqupy.Tempo(system, bath, tparams, rho0, start_time=0.0, unique=True)
        tinfl = [np.asarray(tempo._influence(k), dtype=complex) for k in range(7)]
        tempo.compute(end_time=2.0, progress_type="silent")
        tdyn = tempo.get_dynamics()
        tref = oqupy.Tempo(system, bath, tparams, rho0, start_time=0.0, unique=False)
        tref.compute(end_time=2.0, progress_type="silent")
        assert np.abs(np.array(tdyn.states) - np.array(tref.get_dynamics().states)).max() < 1e-5
        np.savez_compressed(
            os.path.join(HERE, f"tempo_unique_{tag}.npz"), kind="tempo_unique", dim=d,
            dt=0.1, dkmax=6, epsrel=1e-7, num_steps=20, influence_0=tinfl[0],
            influences=np.array(tinfl[1:]), north_map=nmap, west_map=wmap,
            unitary=bath.unitary_transform, prop_1=p1, prop_2=p2,
            initial_state=np.asarray(rho0, dtype=complex), states=np.array(tdyn.states),
            bond_dims=np.array(tempo._backend_instance._mps.bond_dimensions))
        print(f"tempo_unique_{tag}: bonds",
              list(tempo._backend_instance._mps.bond_dimensions))


if __name__ == "__main__":
    main()
