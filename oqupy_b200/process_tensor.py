"""Device-resident process tensor + the compute_dynamics hot loop.

Mirrors the protocol of ``oqupy.process_tensor.SimpleProcessTensor``
(/root/reference/oqupy/process_tensor.py:249-430) that ``compute_dynamics``
(/root/reference/oqupy/system_dynamics.py:41-182) consumes, with the PT-MPO sites
kept on the B200 as rank-3 tensors (past bond, future bond, array leg).
"""
import numpy as np

from ._lib import default_ops

CDTYPE = np.complex128


class DeviceProcessTensor:
    """PT-MPO on the device.  Rank-3 sites only (diagonalised coupling)."""

    def __init__(self, hilbert_space_dimension, dt=None, transform_in=None,
                 transform_out=None, name=None, description=None, ops=None):
        if transform_in is not None or transform_out is not None:
            raise NotImplementedError(
                "oqupy_b200: non-diagonal coupling transforms on the device "
                "process tensor are not supported yet")
        self._hs_dim = hilbert_space_dimension
        self._dt = dt
        self.name = name
        self.description = description
        self._ops = default_ops() if ops is None else ops
        self._sites = []
        self._caps = []
        d = self._hs_dim
        self._trace_square = (np.identity(d, dtype=CDTYPE)
                              / np.sqrt(float(d))).flatten() ** 2   # :55-57

    # -- protocol used by compute_dynamics ------------------------------------
    @property
    def hilbert_space_dimension(self):
        return self._hs_dim

    @property
    def dt(self):
        return self._dt

    @property
    def max_step(self):
        return len(self)

    def __len__(self):
        return len(self._sites)

    def get_initial_tensor(self):
        return None

    def set_mpo_tensor_device(self, step, tensor):
        if step >= len(self._sites):
            self._sites.extend([None] * (step - len(self._sites) + 1))
        self._sites[step] = tensor

    def set_mpo_tensor(self, step, tensor):
        self.set_mpo_tensor_device(step, self._ops.from_host(tensor))

    def get_mpo_tensor(self, step, transformed=True):
        """Host copy, rank-3 (past, future, array) -- process_tensor.py:326-355."""
        if step >= len(self._sites) or step < 0:
            raise IndexError("Process tensor index out of bound. ")
        return self._ops.to_host(self._sites[step])

    def get_mpo_tensor_device(self, step):
        return self._sites[step]

    def get_cap_tensor(self, step):
        if step >= len(self._caps) or step < 0:
            return None
        return self._ops.to_host(self._caps[step])

    def get_cap_tensor_device(self, step):
        return self._caps[step]

    def get_bond_dimensions(self):
        dims = [int(t.shape[0]) for t in self._sites]
        dims.append(int(self._sites[-1].shape[1]))
        return np.array(dims)

    def compute_caps(self):
        """cap_N = [1]; cap_k = sum T_k cap_{k+1} tr^2   (process_tensor.py:380-406)."""
        ops = self._ops
        tr2 = ops.from_host(self._trace_square)
        caps = [ops.from_host(np.array([1.0]))]
        for t in reversed(self._sites):
            chi_l, chi_r, d2 = t.shape
            cap = ops.empty(chi_l)
            ops.caps_step(chi_l, chi_r, d2, t, caps[0], tr2, cap)
            caps.insert(0, cap)
        self._caps = caps


def dynamics_device(pt, propagators, initial_states, num_steps=None, ops=None):
    """compute_dynamics hot loop (system_dynamics.py:131-170), one environment,
    ``E`` ensemble members sharing the process tensor.

    propagators(step) -> (P1, P2), each (d2, d2) or (E, d2, d2).
    initial_states: (d, d) or (E, d, d).  Returns ndarray (E, num_steps+1, d, d)
    (E squeezed if the input was a single state).
    """
    ops = default_ops() if ops is None else ops
    rho0 = np.asarray(initial_states, dtype=CDTYPE)
    single = rho0.ndim == 2
    if single:
        rho0 = rho0[None]
    nvec, d = rho0.shape[0], rho0.shape[1]
    d2 = d * d
    if num_steps is None:
        num_steps = len(pt)
    v = ops.from_host(rho0.reshape(nvec, 1, d2))
    rho = ops.empty(num_steps + 1, nvec, d2)
    cache = {}

    def dev_props(step):
        p1, p2 = propagators(step)
        key = (id(p1), id(p2))
        if key not in cache:
            cache.clear()
            a = np.broadcast_to(np.asarray(p1, dtype=CDTYPE), (nvec, d2, d2))
            b = np.broadcast_to(np.asarray(p2, dtype=CDTYPE), (nvec, d2, d2))
            cache[key] = (ops.from_host(a), ops.from_host(b), p1, p2)
        return cache[key][0], cache[key][1]

    for step in range(num_steps):
        t = pt.get_mpo_tensor_device(step)
        chi_l, chi_r, _ = t.shape
        p1, p2 = dev_props(step)
        v_out = ops.empty(nvec, chi_r, d2)
        ops.dyn_step(nvec, chi_l, chi_r, d2, t, p1, p2, v, v_out,
                     cap=pt.get_cap_tensor_device(step), rho_out=rho[step])
        v = v_out
    # final read-out: rho[N] = sum_l cap_N[l] v[l]   (system_dynamics.py:167-170)
    from ._lib import View  # pylint: disable=import-outside-toplevel
    cap = pt.get_cap_tensor_device(num_steps)
    chi = v.shape[1]
    ops.gemm(1, d2, chi, View(cap, col=1), View(v, row=d2, col=1, b1=chi * d2),
             View(rho[num_steps], col=1, b1=d2), nb1=nvec)
    out = ops.to_host(rho).transpose(1, 0, 2).reshape(nvec, num_steps + 1, d, d)
    return out[0] if single else out
