"""Generate tests/golden/mean_field_G.npz by running the UNMODIFIED reference's
MeanFieldTempo (reference test G: tests/physics/mean_field_tempo_test.py:21-69).

Build-container only (needs /root/reference + oracle/tn_shim).  The fixture stores what
crosses the MeanFieldTempoBackend boundary (tempo_backend.py:629-773): the influence
matrices, and -- because the half-step propagators and the field update are host
callables of the reference's System objects -- the propagator pair, field and field
derivative the reference actually used at every step, next to its outputs (states,
fields) and the golden values quoted in the reference test (rho_G, field_G).

numpy-2 note: the reference validates Hamiltonians with np.vectorize(h)(t), which numpy 2
rejects for array-valued h (the reference pins numpy<2); the generator neutralises that
CHECK only (np.vectorize -> identity); no arithmetic of the reference is touched.
"""
import os
import sys

import numpy as np

np.vectorize = lambda f, *a, **k: f   # see module docstring

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from ref_loader import load_reference  # noqa: E402

oqupy = load_reference()


def main():
    from oqupy import operators
    rho0 = np.array([[0.5, 0.5], [0.5, 0.5]])
    field0 = 1.0
    rho_g = np.array([[0.6245009 + 3.19373236e-15j, 0.14243496 - 2.19523032e-01j],
                      [0.14243496 + 2.19523032e-01j, 0.3754991 - 2.26692588e-15j]])
    field_g = 0.10602369935009 - 0.46986388684474406j

    def h_sys(t, field):
        return 0.5 * operators.sigma("z") + np.real(field) * operators.sigma("x")

    def field_eom(t, states, field):
        return -(1j + 1) * field \
            - 0.5j * np.matmul(operators.sigma("x"), states[0]).trace().real

    corr = oqupy.PowerLawSD(alpha=0.1, zeta=1.0, cutoff=5.0, cutoff_type="gaussian",
                            temperature=0.0)
    bath = oqupy.Bath(0.5 * operators.sigma("z"), corr)
    system = oqupy.TimeDependentSystemWithField(h_sys)
    mfs = oqupy.MeanFieldSystem([system], field_eom)
    params = oqupy.TempoParameters(dt=0.05, tcut=None, epsrel=1e-7)
    tempo = oqupy.MeanFieldTempo(mfs, [bath], params, [rho0], field0, start_time=0.0)
    be = tempo._backend_instance
    rec = dict(p1=[], p2=[], field_in=[], dfield_in=[], field_out=[])
    orig_prop = be._propagators_list[0]

    def prop(step, field, dfield):
        p1, p2 = orig_prop(step, field, dfield)
        rec["p1"].append(np.array(p1)); rec["p2"].append(np.array(p2))
        rec["field_in"].append(field); rec["dfield_in"].append(dfield)
        return p1, p2
    be._propagators_list[0] = prop
    orig_cf = be._compute_field

    def cf(*args):
        out = orig_cf(*args)
        rec["field_out"].append(out)
        return out
    be._compute_field = cf
    n_steps = 20
    infl = np.array([np.asarray(be._backend_list[0]._influence(k), dtype=complex)
                     for k in range(n_steps + 1)])
    tempo.compute(end_time=1.0, progress_type="silent")
    dyn = tempo.get_dynamics()
    states = np.array(dyn.system_dynamics[0].states)
    np.savez_compressed(
        os.path.join(HERE, "mean_field_G.npz"), kind="mean_field", dim=2, dt=0.05,
        dkmax=-1, epsrel=1e-7, num_steps=n_steps, influences=infl,
        unitary=bath.unitary_transform, initial_state=rho0.astype(complex),
        initial_field=complex(field0), props_1=np.array(rec["p1"]),
        props_2=np.array(rec["p2"]), fields_in=np.array(rec["field_in"]),
        dfields_in=np.array(rec["dfield_in"]), fields=np.array(dyn.fields),
        states=states, rho_golden=rho_g, field_golden=complex(field_g),
        bond_dims=np.array(be._backend_list[0]._mps.bond_dimensions))
    print("mean_field_G: final field", dyn.fields[-1], "golden", field_g)
    print("bond dims", be._backend_list[0]._mps.bond_dimensions)


if __name__ == "__main__":
    main()
