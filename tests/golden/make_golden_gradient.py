"""Generate tests/golden/gradient_J.npz by running the UNMODIFIED reference's
compute_gradient_and_dynamics + _chain_rule on its own test J
(tests/physics/gradient_target_state_test.py:30-113).

Build-container only (needs /root/reference + oracle/tn_shim).  The bath of test J is the
bath of test A, so the process tensor is rebuilt from tests/golden/pt_refA.npz (its
influence matrices); the derivative tensors are gauge invariant (bond legs contracted).
Stored: the half-step propagators and their parameter derivatives at every step (host
callables of the reference's ParameterizedSystem), initial state, target derivative, and
the reference's outputs: propagator_derivatives (gradient.py:216-224), states, the
chain-rule gradient and the golden values quoted in the reference test (grad_params_J).

numpy-2 note: see make_golden_mean_field.py (np.vectorize input CHECK neutralised).
"""
import os
import sys

import numpy as np

np.vectorize = lambda f, *a, **k: f

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from ref_loader import load_reference  # noqa: E402

oqupy = load_reference()

GRAD_PARAMS_J = [0.00507649, 0.00534207, 0.0053693, 0.00557299, 0.00559541, 0.00575529,
                 0.00577277, 0.0059047, 0.00591734, 0.0060279, 0.0060359, 0.00612581,
                 0.00612941, 0.00619729, 0.00619673, 0.00624089, 0.00623642, 0.00625545,
                 0.00624727, 0.00624009, 0.00622842, 0.00619391, 0.00617894, 0.00611593,
                 0.00609786, 0.00600489, 0.00598388, 0.00585915, 0.00583539, 0.00567655,
                 0.00565021, 0.00545466, 0.0054259, 0.00519144, 0.00516043, 0.00488701,
                 0.00485394, 0.00454699, 0.00451208, 0.00418564]


def main():
    dt, num_steps = 0.05, 20
    x0 = np.ones((2 * num_steps, 1))
    rho0 = np.array([[1.0, 0.0], [0.0, 0.0]])
    target = np.array([[0.0, 0.0], [0.0, 1.0]])
    corr = oqupy.PowerLawSD(alpha=0.3, zeta=1.0, cutoff=5.0, cutoff_type="exponential",
                            temperature=0.2)
    bath = oqupy.Bath(np.array([[0.5, 0.0], [0.0, -0.5]]), corr)
    system = oqupy.ParameterizedSystem(
        hamiltonian=lambda hx: 0.5 * hx * oqupy.operators.sigma("x"),
        gammas=[lambda t: 0.1, lambda t: 0.2],
        lindblad_operators=[lambda t: oqupy.operators.sigma("-"),
                            lambda t: oqupy.operators.sigma("z")])
    params = oqupy.TempoParameters(dt=dt, tcut=None, epsrel=1e-7)
    pt = oqupy.pt_tempo_compute(bath, start_time=0.0, end_time=1.0, parameters=params,
                                progress_type="silent")
    grad_prop, dyn = oqupy.compute_gradient_and_dynamics(
        system=system, parameters=x0, process_tensors=[pt], initial_state=rho0,
        target_derivative=target.T, progress_type="silent")
    derivs = np.array([np.asarray(getattr(g, "tensor", g)) for g in grad_prop])
    props = system.get_propagators(dt, parameters=x0)
    # the reference differentiates its propagators with numdifftools (absent here):
    # central differences of the reference's own get_propagators instead (test J uses
    # the same parameter value at every half step)
    h = 1e-6
    pp, pm = (system.get_propagators(dt, parameters=x0 * (1.0 + sgn * h))
              for sgn in (1.0, -1.0))
    dhalf = [(np.asarray(pp(0)[i]) - np.asarray(pm(0)[i])) / (2.0 * h) for i in (0, 1)]

    def dprops(step):
        return [dhalf[0]], [dhalf[1]]
    grad = oqupy.gradient._chain_rule(adjoint_tensor=grad_prop, dprop_dparam=dprops,
                                      propagators=props, num_steps=num_steps,
                                      num_parameters=1, progress_type="silent")
    p1 = np.array([props(k)[0] for k in range(num_steps)])
    p2 = np.array([props(k)[1] for k in range(num_steps)])
    dp1 = np.array([np.asarray(dprops(k)[0]) for k in range(num_steps)])
    dp2 = np.array([np.asarray(dprops(k)[1]) for k in range(num_steps)])
    np.savez_compressed(
        os.path.join(HERE, "gradient_J.npz"), kind="gradient", dim=2, dt=dt,
        num_steps=num_steps, pt_fixture="pt_refA", initial_state=rho0.astype(complex),
        target_derivative=target.T.astype(complex), props_1=p1, props_2=p2,
        dprops_1=dp1, dprops_2=dp2, propagator_derivatives=derivs,
        states=np.array(dyn.states), grad_params=grad,
        grad_params_golden=np.array(GRAD_PARAMS_J))
    print("gradient_J: derivs", derivs.shape, "max |grad - golden|",
          np.abs(grad.real[:, 0] - np.array(GRAD_PARAMS_J)).max())


if __name__ == "__main__":
    main()
