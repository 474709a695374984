"""Generate tests/golden/pt_tebd_F{1,2}.npz by running the UNMODIFIED reference on its own
PT-TEBD test F (tests/physics/pt_tebd_test.py: XYZ chain of 5 spins, without baths (F1)
and with a process tensor on sites 0 and 3 (F2)).  Build-container only.

The fixtures hold what crosses the PtTebdBackend boundary (pt_tebd_backend.py:46-114):
initial gammas / lambdas, the gate tensors of every layer of the second-order propagator
(mps_mpo.py:329-338), the process-tensor sites and caps, next to the reference's outputs
per step (single-site density matrices, one two-site density matrix, norm, bond
dimensions) and the reference's golden matrices example_F{1,2}_rhos.npy."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from ref_loader import load_reference, REFERENCE_ROOT  # noqa: E402

oqupy = load_reference()
from oqupy.mps_mpo import compute_tebd_propagator  # noqa: E402


def main():
    dt, num_steps, n = 0.1, 10, 5
    corr = oqupy.PowerLawSD(alpha=0.3, zeta=3, cutoff=3.0, cutoff_type="exponential",
                            temperature=0.8)
    bath = oqupy.Bath(0.5 * oqupy.operators.sigma("z"), corr)
    tp = oqupy.TempoParameters(dt=dt, dkmax=10, epsrel=1.0e-6, add_correlation_time=5.0)
    pt = oqupy.pt_tempo_compute(bath=bath, start_time=0.0, end_time=num_steps * dt,
                                parameters=tp, progress_type="silent")
    h = np.array([[1.0, 0, 0], [2.0, 0, 0], [3.0, 0, 0], [4.0, 0, 0], [5.0, 0, 0]]) * np.pi / 10
    jj = np.array([[1.2, 1.3, 0.7]] * (n - 1))
    chain = oqupy.SystemChain(hilbert_space_dimensions=[2] * n)
    for s in range(n):
        for i, xyz in enumerate("xyz"):
            chain.add_site_hamiltonian(site=s, hamiltonian=0.5 * h[s, i] * oqupy.operators.sigma(xyz))
    for s in range(n - 1):
        for i, xyz in enumerate("xyz"):
            chain.add_nn_hamiltonian(site=s, hamiltonian_l=0.5 * jj[s, i] * oqupy.operators.sigma(xyz),
                                     hamiltonian_r=0.5 * oqupy.operators.sigma(xyz))
    params = oqupy.PtTebdParameters(dt=dt, order=2, epsrel=1.0e-7)
    amps = oqupy.AugmentedMPS([oqupy.operators.spin_dm("z-")] * n)
    prop = compute_tebd_propagator(system_chain=chain, time_step=dt / 2.0, epsrel=1.0e-7,
                                   order=2)
    gates = {}
    for li, layer in enumerate(prop.gate_layers):
        for gi, gate in enumerate(layer.gates):
            gates[f"gate_{li}_{gi}_sites"] = np.array(gate.sites)
            gates[f"gate_{li}_{gi}_l"] = np.array(gate.tensors[0], dtype=complex)
            gates[f"gate_{li}_{gi}_r"] = np.array(gate.tensors[1], dtype=complex)
    gates["layer_sizes"] = np.array([len(layer.gates) for layer in prop.gate_layers])
    for tag, pts in (("F1", [None] * n), ("F2", [pt, None, None, pt, None])):
        run = oqupy.PtTebd(initial_augmented_mps=amps, system_chain=chain, process_tensors=pts,
                           parameters=params, dynamics_sites=list(range(n)) + [(1, 3)],
                           chain_control=None)
        res = run.compute(num_steps, progress_type="silent")
        golden = np.load(os.path.join(REFERENCE_ROOT, "tests", "data", "correct_results",
                                      f"example_{tag}_rhos.npy"))
        out = dict(gates)
        out.update(
            kind="pt_tebd", n=n, dt=dt, num_steps=num_steps, epsrel=1.0e-7,
            pt_sites=np.array([i for i, p in enumerate(pts) if p is not None], dtype=int),
            states=np.array([res["dynamics"][s].states for s in range(n)]),
            states_13=np.array(res["dynamics"][(1, 3)].states),
            norm=np.array(res["norm"]), bond_dims=np.array(res["bond_dimensions"]),
            rho_golden=golden)
        for i, g in enumerate(amps.gammas):
            out[f"gamma_{i}"] = np.array(g, dtype=complex)
        for i, lam in enumerate(amps.lambdas):
            out[f"lambda_{i}"] = np.array(lam, dtype=complex)
        if tag == "F2":      # rank-3 sites (delta between the system legs), no transforms
            assert pt._transform_in is None and pt._transform_out is None
            for k in range(len(pt)):
                assert pt._mpo_tensors[k].ndim == 3
                out[f"pt_mpo_{k}"] = np.array(pt._mpo_tensors[k], dtype=complex)
            for k in range(len(pt) + 1):
                out[f"pt_cap_{k}"] = np.array(pt.get_cap_tensor(k), dtype=complex)
        out["pt_len"] = len(pt)
        np.savez_compressed(os.path.join(HERE, f"pt_tebd_{tag}.npz"), **out)
        for s in range(n):
            # the reference's own pin (pt_tebd_test.py:116, 138: decimal=4)
            np.testing.assert_almost_equal(res["dynamics"][s].states[-1], golden[s], decimal=4)
        print(tag, "bonds", res["bond_dimensions"][-1], "norm", res["norm"][-1],
              "pt bonds", list(pt.get_bond_dimensions()))


if __name__ == "__main__":
    main()
