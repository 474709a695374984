"""Generate tests/golden/gradient_multi.npz: the UNMODIFIED reference's
compute_gradient_and_dynamics (oqupy/gradient.py:169-437) with (a) TWO environments (two
different ohmic baths, cf. tests/physics/multi_environments_test.py:70-93) and (b) ONE
environment plus pre- and post-measurement controls (oqupy/control.py) at several steps,
step N included.

Build-container only (needs /root/reference + oracle/tn_shim).  Stored: the PT-MPO tensors
(rank 3: past, future, array) and cap tensors of both process tensors as the reference built
them, the half-step propagators, the control superoperators per step, and the reference's
outputs (propagator_derivatives, states) of both runs.

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


def pt_arrays(pt, n):
    mpos = [np.asarray(pt._mpo_tensors[k]) for k in range(n)]      # rank 3 as stored
    caps = [np.asarray(pt.get_cap_tensor(k)) for k in range(n + 1)]
    return mpos, caps


def main():
    dt, num_steps = 0.05, 12
    x0 = np.ones((2 * num_steps, 1)) * 1.3
    rho0 = np.array([[0.7, 0.2 - 0.1j], [0.2 + 0.1j, 0.3]])
    target = np.array([[0.1, 0.3j], [-0.3j, 0.9]])
    sz = np.array([[0.5, 0.0], [0.0, -0.5]])
    baths = [oqupy.Bath(sz, oqupy.PowerLawSD(alpha=a, zeta=1.0, cutoff=5.0,
                                             cutoff_type="exponential", temperature=t))
             for a, t in ((0.3, 0.2), (0.12, 1.1))]
    system = oqupy.ParameterizedSystem(
        hamiltonian=lambda hx: 0.5 * hx * oqupy.operators.sigma("x")
        + 0.2 * oqupy.operators.sigma("z"),
        gammas=[lambda t: 0.1], lindblad_operators=[lambda t: oqupy.operators.sigma("-")])
    params = oqupy.TempoParameters(dt=dt, tcut=6 * dt, epsrel=1e-7)
    pts = [oqupy.pt_tempo_compute(b, start_time=0.0, end_time=dt * num_steps,
                                  parameters=params, progress_type="silent") for b in baths]
    props = system.get_propagators(dt, parameters=x0)
    p1 = np.array([props(k)[0] for k in range(num_steps)])
    p2 = np.array([props(k)[1] for k in range(num_steps)])
    out = {"kind": "gradient_multi", "dim": 2, "dt": dt, "num_steps": num_steps,
           "initial_state": rho0.astype(complex), "target_derivative": target.T.astype(complex),
           "props_1": p1, "props_2": p2}
    for e, pt in enumerate(pts):
        mpos, caps = pt_arrays(pt, num_steps)
        for k, t in enumerate(mpos):
            assert t.ndim == 3
            out[f"mpo_{e}_{k}"] = t
        for k, c in enumerate(caps):
            out[f"cap_{e}_{k}"] = c
    # (a) two environments
    g2, dyn2 = oqupy.compute_gradient_and_dynamics(
        system=system, parameters=x0, process_tensors=pts, initial_state=rho0,
        target_derivative=target.T, dt=dt, progress_type="silent")
    out["derivs_two_env"] = np.array([np.asarray(getattr(g, "tensor", g)) for g in g2])
    out["states_two_env"] = np.array(dyn2.states)
    # (b) one environment + controls
    rng = np.random.default_rng(5)
    ctrl = oqupy.Control(2)
    steps_pre, steps_post = [0, 4, num_steps], [2, 4, 9]
    ctl = np.zeros((num_steps + 1, 2, 4, 4), dtype=complex)
    has = np.zeros((num_steps + 1, 2), dtype=bool)
    for post, steps in ((False, steps_pre), (True, steps_post)):
        for st in steps:
            u = np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]
            sup = 0.9 * np.kron(u, u.conj()) + 0.1 * np.identity(4)   # not trace preserving
            ctrl.add_single(st, sup, post=post)
            ctl[st, int(post)] = sup
            has[st, int(post)] = True
    g1, dyn1 = oqupy.compute_gradient_and_dynamics(
        system=system, parameters=x0, process_tensors=[pts[0]], initial_state=rho0,
        target_derivative=target.T, dt=dt, control=ctrl, progress_type="silent")
    out["derivs_controls"] = np.array([np.asarray(getattr(g, "tensor", g)) for g in g1])
    out["states_controls"] = np.array(dyn1.states)
    out["controls"] = ctl
    out["has_control"] = has
    np.savez_compressed(os.path.join(HERE, "gradient_multi.npz"), **out)
    print("gradient_multi: two-env derivs", out["derivs_two_env"].shape, "bonds",
          [out[f"mpo_0_{k}"].shape[1] for k in range(num_steps)],
          [out[f"mpo_1_{k}"].shape[1] for k in range(num_steps)])


if __name__ == "__main__":
    main()
