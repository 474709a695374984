"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Build-container only (needs /root/reference; third-party ``tensornetwork`` is
provided by the restated slice in oracle/tn_shim).  Run:  python tests/golden/make_golden.py

Each fixture stores the INPUTS crossing the drop-in boundary (influence matrices
per dk, half-step propagators, initial state) and the reference's OUTPUTS
(states per step, bond dimensions, PT-MPO caps).  The golden density matrices
quoted from the reference's own tests are stored next to them:
  rho_A  /root/reference/tests/physics/tempo_spin_boson_test.py:45-46
  rho_C  /root/reference/tests/physics/tempo_qutrit_test.py:59-63
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from ref_loader import load_reference  # noqa: E402

oqupy = load_reference()


def _influences(pt_or_tempo, ks):
    out = []
    for k in ks:
        m = pt_or_tempo._influence(k)
        out.append(np.asarray(m, dtype=complex))
    return np.array(out)


def pt_case(name, bath, system, params, t_end, rho0, extra=None):
    n_steps = int((t_end - 0.0) / params.dt)
    pt_obj = oqupy.PtTempo(bath, 0.0, t_end, params)
    dkmax = params.dkmax if params.dkmax is not None else n_steps
    num_infl = min(n_steps, dkmax + 1)
    infl = _influences(pt_obj, range(num_infl))
    pt_obj.compute(progress_type="silent")
    pt = pt_obj.get_process_tensor(progress_type="silent")
    dyn = oqupy.compute_dynamics(system=system, process_tensor=pt,
                                 initial_state=rho0, progress_type="silent")
    p1, p2 = system.get_propagators(params.dt, 0.0, 256, 2 ** -26)(0)
    caps = [pt.get_cap_tensor(k) for k in range(len(pt) + 1)]
    data = dict(
        kind="pt", dim=bath.dimension, dt=params.dt, dkmax=dkmax,
        epsrel=params.epsrel, num_steps=n_steps, influences=infl,
        prop_1=p1, prop_2=p2, initial_state=np.asarray(rho0, dtype=complex),
        states=np.array(dyn.states), bond_dims=pt.get_bond_dimensions(),
        cap_first=caps[0], cap_mid=caps[len(caps) // 2],
        mid_index=len(caps) // 2)
    if extra:
        data.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print(name, "bond dims", list(pt.get_bond_dimensions()))
    return pt


def tempo_case(name, bath, system, params, n_steps, rho0, extra=None):
    tempo = oqupy.Tempo(system, bath, params, rho0, start_time=0.0)
    dkmax = params.dkmax if params.dkmax is not None else n_steps
    infl = _influences(tempo, range(min(dkmax, n_steps) + 1))
    tempo.compute(end_time=n_steps * params.dt, progress_type="silent")
    dyn = tempo.get_dynamics()
    p1, p2 = system.get_propagators(params.dt, 0.0, 256, 2 ** -26)(0)
    data = dict(
        kind="tempo", dim=bath.dimension, dt=params.dt,
        dkmax=-1 if params.dkmax is None else params.dkmax,
        epsrel=params.epsrel, num_steps=n_steps, influences=infl,
        prop_1=p1, prop_2=p2, initial_state=np.asarray(rho0, dtype=complex),
        unitary=bath.unitary_transform,
        states=np.array(dyn.states),
        bond_dims=np.array(tempo._backend_instance._mps.bond_dimensions))
    if extra:
        data.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print(name, "bond dims", data["bond_dims"].tolist())


def main():
    sx, sy, sz = (oqupy.operators.sigma(c) for c in "xyz")
    up = oqupy.operators.spin_dm("z+")

    # -- configs 1/2 bath: ohmic alpha=0.08, wc=4, T=1.6, dt=0.05 ------
    corr = oqupy.PowerLawSD(alpha=0.08, zeta=1, cutoff=4.0,
                            cutoff_type="exponential", temperature=1.6)
    bath = oqupy.Bath(0.5 * sz, corr)
    system = oqupy.System(0.5 * sx)

    pt_case("pt_k12_eps7_n30", bath, system,
            oqupy.TempoParameters(dt=0.05, dkmax=12, epsrel=1e-7), 1.5, up)
    pt_case("pt_k8_eps9_n24", bath, system,
            oqupy.TempoParameters(dt=0.05, dkmax=8, epsrel=1e-9), 1.2, up)
    tempo_case("tempo_c1_k20_eps7_n60", bath, system,
               oqupy.TempoParameters(dt=0.05, dkmax=20, epsrel=1e-7), 60, up)

    # -- config 2 operands only (influences at eps=1e-9, dk=0..200) -----------------
    p2 = oqupy.TempoParameters(dt=0.05, dkmax=200, epsrel=1e-9)
    obj = oqupy.PtTempo(bath, 0.0, 50.0, p2)
    p1m, p2m = system.get_propagators(0.05, 0.0, 256, 2 ** -26)(0)
    np.savez_compressed(os.path.join(HERE, "c2_operands.npz"),
                        kind="operands", dim=2, dt=0.05, dkmax=200,
                        epsrel=1e-9, num_steps=1000,
                        influences=_influences(obj, range(201)),
                        prop_1=p1m, prop_2=p2m,
                        initial_state=np.asarray(up, dtype=complex))

    # -- reference test A (dkmax=None, Lindblad system) ----------------------------
    rho_a = np.array([[0.7809559 + 0.j, -0.09456333 + 0.16671419j],
                      [-0.09456333 - 0.16671419j, 0.2190441 + 0.j]])
    corr_a = oqupy.PowerLawSD(alpha=0.3, zeta=1.0, cutoff=5.0,
                              cutoff_type="exponential", temperature=0.2)
    bath_a = oqupy.Bath(np.array([[0.5, 0.0], [0.0, -0.5]]), corr_a)
    sys_a = oqupy.System(np.array([[0.0, 0.5], [0.5, 0.0]]), gammas=[0.1, 0.2],
                         lindblad_operators=[oqupy.operators.sigma("-"), sz])
    rho0_a = np.array([[1.0, 0.0], [0.0, 0.0]])
    par_a = oqupy.TempoParameters(dt=0.05, tcut=None, epsrel=1e-7)
    pt_case("pt_refA", bath_a, sys_a, par_a, 1.0, rho0_a,
            extra=dict(rho_golden=rho_a))
    tempo_case("tempo_refA", bath_a, sys_a, par_a, 20, rho0_a,
               extra=dict(rho_golden=rho_a))

    # -- reference test C (qutrit, dkmax=10: grow + end phase, d2=9) ---------------
    rho_c = np.array(
        [[0.12576653 + 0.j, -0.11739956 - 0.14312036j, 0.12211454 - 0.05963583j],
         [-0.11739956 + 0.14312036j, 0.61315893 + 0.j, -0.06636825 + 0.26917271j],
         [0.12211454 + 0.05963583j, -0.06636825 - 0.26917271j, 0.26107455 + 0.j]])
    h_c = np.array([[1.0, 0.5, 0.0], [0.5, 0.5, 0.5], [0.0, 0.5, 0.0]])
    l1 = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    l2 = np.array([[0.5, 0.0, 0.0], [0.0, -0.5, 0.0], [0.0, 0.0, 0.0]])
    corr_c = oqupy.PowerLawSD(alpha=0.3, zeta=1.0, cutoff=5.0,
                              cutoff_type="exponential", temperature=0.0)
    bath_c = oqupy.Bath(l2.copy(), corr_c)
    sys_c = oqupy.System(h_c, gammas=[0.1, 0.2], lindblad_operators=[l1, l2])
    rho0_c = np.diag([0.0, 1.0, 0.0])
    par_c = oqupy.TempoParameters(dt=0.05, dkmax=10, epsrel=1e-7)
    pt_case("pt_refC", bath_c, sys_c, par_c, 1.0, rho0_c,
            extra=dict(rho_golden=rho_c))
    tempo_case("tempo_refC", bath_c, sys_c, par_c, 20, rho0_c,
               extra=dict(rho_golden=rho_c))

    # -- non-diagonal coupling (sigma_x), TEMPO: exercises unitary_transform -------
    corr_b = oqupy.PowerLawSD(alpha=0.1, zeta=1.0, cutoff=5.0,
                              cutoff_type="exponential", temperature=0.2)
    bath_b = oqupy.Bath(0.5 * sx, corr_b)
    sys_b = oqupy.System(0.5 * sz + 0.3 * sy)
    par_b = oqupy.TempoParameters(dt=0.1, dkmax=8, epsrel=1e-6)
    tempo_case("tempo_nondiag", bath_b, sys_b, par_b, 24,
               oqupy.operators.spin_dm("y+"))


if __name__ == "__main__":
    main()
