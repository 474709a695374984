"""Drop-in backend classes: same constructor signatures and method contracts as

* ``oqupy.backends.pt_tempo_backend.PtTempoBackend``   (pt_tempo_backend.py:28-316)
* ``oqupy.backends.tempo_backend.BaseTempoBackend``    (tempo_backend.py:314-575)
* ``oqupy.backends.tempo_backend.TempoBackend``        (tempo_backend.py:578-626)

but with the matrix-product state resident on the B200 and every contraction /
truncated SVD executed by the C-ABI kernels (oqupy_b200/csrc).  Host work is
limited to what the reference also does on the host per step: calling the
``influence`` / ``propagators`` callbacks (d2 x d2 matrices) and O(d2^2) operand
preparation.
"""
from copy import copy, deepcopy

import numpy as np

import os

from . import chain
from ._lib import NativeChain, View, default_ops
from .chain import PtSite, TempoSite

CDTYPE = np.complex128


def _check_maps(degeneracy_maps, d2, n_north, n_west):
    """[north_map, west_map] of unique=True (oqupy/bath.py:87-89, 158-164) as int arrays,
    or None.  The reduced legs carry n_north / n_west distinct values."""
    if degeneracy_maps is None:
        return None
    nmap, wmap = (np.asarray(m).astype(np.int64).reshape(-1) for m in degeneracy_maps)
    assert nmap.size == d2 and wmap.size == d2, "degeneracy maps must have dim**2 entries"
    assert nmap.min() >= 0 and wmap.min() >= 0
    assert int(nmap.max()) + 1 == n_north, "north map does not match sum_north"
    assert int(wmap.max()) + 1 == n_west, "west map does not match sum_west"
    return nmap, wmap


class PtTempoBackend:
    """PT-TEMPO process-tensor build on the device (pt_tempo_backend.py:28-316)."""

    def __init__(self, dimension, influence, process_tensor, sum_north,
                 sum_west, num_steps, dkmax, epsrel, config=None,
                 degeneracy_maps=None, ops=None):
        self._maps = _check_maps(degeneracy_maps, dimension ** 2, len(sum_north),
                                 len(sum_west))
        self._dimension = dimension
        self._influence = influence
        self._process_tensor = process_tensor
        self._sum_north = np.asarray(sum_north, dtype=float)
        self._sum_west = np.asarray(sum_west, dtype=float)
        self._num_steps = num_steps
        self._dkmax = dkmax
        self._epsrel = epsrel
        self._config = {} if config is None else config
        self._step = None
        self._num_infl = min(num_steps, dkmax + 1)
        assert self._num_infl >= 2, "need at least two influence functions"
        self._ops = default_ops() if ops is None else ops
        self._mps = None
        self._mpo = None
        self._closing = None
        # On the device the chain lives in the native C++ engine (csrc/chain.cu: one
        # C-ABI call per zip-up / sweep).  The Python chain (chain.py, same kernels) is
        # the readable specification; it serves the TEMPO path, the per-SVD profiling
        # tools (OQUPY_B200_PYCHAIN=1) and the CPU-only host-logic tests.
        self._native = (getattr(self._ops, "name", "") == "cuda"
                        and os.environ.get("OQUPY_B200_PYCHAIN", "0") != "1")
        self._chain = None

    @property
    def step(self):
        """The current step in the PT-TEMPO computation."""
        return self._step

    @property
    def num_steps(self):
        """The number of steps of the process tensor."""
        return self._num_steps

    def initialize(self):
        """Build the num_infl MPO sites and the initial MPS (:105-190)."""
        ops, d = self._ops, self._dimension
        d2 = d * d
        self._closing = self._sum_north * d                 # :112
        mpo, mps = [], []
        for i in range(self._num_infl):
            infl = np.asarray(self._influence(i), dtype=CDTYPE)
            if i == 0 and self._maps is not None:           # :125-137 (unique=True)
                nmap, wmap = self._maps
                vec = infl.reshape(-1) / d                  # n_north distinct values
                mpo.append(PtSite("first", ops.from_host(vec), maps=(nmap, wmap)))
                a = np.zeros((1, d2, vec.size), dtype=CDTYPE)
                a[0, np.arange(d2), nmap] = vec[nmap] / d
            elif i == 0:                                    # :122-143
                vec = np.diag(infl) / d
                mpo.append(PtSite("first", ops.from_host(vec)))
                a = np.zeros((1, d2, d2), dtype=CDTYPE)
                a[0, np.arange(d2), np.arange(d2)] = vec / d
            elif i == self._num_infl - 1:                   # :144-148
                mpo.append(PtSite("last", ops.from_host(infl)))
                a = infl.reshape(infl.shape[0], infl.shape[1], 1)
            else:                                           # :149-152
                mpo.append(PtSite("mid", ops.from_host(infl)))
                nn = infl.shape[0]          # (north, west): d2 x d2 unless unique=True
                a = np.zeros((nn, infl.shape[1], nn), dtype=CDTYPE)
                idx = np.arange(nn)
                a[idx, :, idx] = infl / d
            mps.append(ops.from_host(a))
        self._mpo, self._mps = mpo, mps
        n = len(mps)
        if self._native:
            self._chain = NativeChain(ops)
            for t in mps:
                self._chain.push(t)
            self._mps = None
            self._chain.svd_sweep(n - 1, 0, self._epsrel)           # :171-175
            self._chain.svd_sweep(0, n - 1, self._epsrel)           # :177-181
        else:
            chain.svd_sweep_left(ops, mps, n - 1, 0, self._epsrel)      # :171-175
            chain.svd_sweep_right(ops, mps, 0, n - 1, self._epsrel)     # :177-181
        self._step = 1

    def compute_step(self):
        """One column of the PT-TEMPO network (:192-282)."""
        ops = self._ops
        self._step += 1
        end_phase = self._step > self._num_steps - self._num_infl + 1
        if end_phase:                                               # :229-236
            self._mpo = self._mpo[:-1]
            last = self._mpo[-1]
            if last.kind == "first":
                vec = ops.to_host(last.mat) * self._closing
                self._mpo[-1] = PtSite("first", ops.from_host(vec), maps=last.maps)
            else:
                mat = ops.to_host(last.mat) * self._closing[:, None]
                self._mpo[-1] = PtSite("closed", ops.from_host(mat))
        else:
            infl = self._influence(int(0 - self._step))             # :238-257
            if infl is not None:
                self._mpo[-1] = PtSite("last", ops.from_host(
                    np.asarray(infl, dtype=CDTYPE)))
            new_site = ops.from_host(np.ones((1, 1, 1)))            # :259-263
            if self._native:
                self._chain.push(new_site)
            else:
                self._mps.append(new_site)
        if self._native:
            self._chain.pt_zip_up_left(self._mpo, self._epsrel)     # :267-274
            self._chain.svd_sweep(self._step - 2, len(self._chain) - 1,
                                  self._epsrel)                     # :276-280
        else:
            chain.pt_zip_up_left(ops, self._mps, self._mpo, self._epsrel)   # :267-274
            chain.svd_sweep_right(ops, self._mps, self._step - 2,
                                  len(self._mps) - 1, self._epsrel)     # :276-280
        return self._step < self._num_steps

    def _site(self, k):
        return self._chain.site(k) if self._native else self._mps[k]

    def _num_sites(self):
        return len(self._chain) if self._native else len(self._mps)

    def pop_svd_log(self):
        """[(m, n, keep, sweeps)] of the truncated SVDs since the last call (native chain)."""
        return self._chain.log(True) if self._native and self._chain else []

    def chain_stats(self, reset=False):
        """(truncated SVDs, Jacobi sweeps, D2H bytes) of the native chain."""
        return self._chain.stats(reset) if self._native and self._chain else (0, 0, 0)

    # -- results ----------------------------------------------------------------
    def get_mpo_tensor_device(self, step):
        """(past bond, future bond, array leg) * d on the device (:284-307)."""
        assert self._num_sites() == self._num_steps
        return (self._site(step).permute(0, 2, 1) * self._dimension).contiguous()

    def get_mpo_tensor(self, step):
        return self._ops.to_host(self.get_mpo_tensor_device(step))

    def get_bond_dimensions(self):
        if self._native:
            return [1] + [self._chain.shape(k)[2] for k in range(len(self._chain))]
        return [1] + [int(t.shape[2]) for t in self._mps]

    def update_process_tensor(self):
        """Write the sites into the process tensor and build its caps (:309-316)."""
        assert self._step >= self._num_steps
        pt = self._process_tensor
        if hasattr(pt, "set_mpo_tensor_device"):
            for step in reversed(range(self._num_steps)):
                pt.set_mpo_tensor_device(step, self.get_mpo_tensor_device(step))
        else:
            # a host process tensor (the reference's SimpleProcessTensor) gets host copies
            # through its own protocol; the device copies stay attached to it, so that the
            # rebound oqupy.compute_dynamics (install.py) does not upload them again
            from .process_tensor import DeviceProcessTensor  # pylint: disable=import-outside-toplevel
            dev = DeviceProcessTensor(self._dimension, dt=getattr(pt, "dt", None),
                                      ops=self._ops)
            for step in reversed(range(self._num_steps)):
                site = self.get_mpo_tensor_device(step)
                dev.set_mpo_tensor_device(step, site)
                pt.set_mpo_tensor(step, self._ops.to_host(site))
            dev.compute_caps()
            try:
                pt._b200_device = (self._num_steps, dev)   # pylint: disable=protected-access
            except AttributeError:
                pass
        pt.compute_caps()


class BaseTempoBackend:
    """TEMPO time step on the device (tempo_backend.py:314-575)."""

    def __init__(self, initial_state, influence, unitary_transform, sum_north,
                 sum_west, dkmax, epsrel, config=None, degeneracy_maps=None,
                 dim=None, ops=None):
        self._maps = _check_maps(degeneracy_maps, np.asarray(initial_state).size,
                                 len(sum_north), len(sum_west))
        self._initial_state = initial_state
        self._influence = influence
        self._unitary_transform = np.asarray(unitary_transform, dtype=CDTYPE)
        self._sum_north = np.asarray(sum_north, dtype=float)
        self._sum_west = np.asarray(sum_west, dtype=float)
        self._dkmax = dkmax
        self._epsrel = epsrel
        self._step = None
        self._state = None
        self._config = {} if config is None else config
        self._dim = dim
        self._ops = default_ops() if ops is None else ops
        self._mps = None
        # device: the whole time step is ONE C-ABI call on the native chain engine
        # (b200_chain_tempo_step); the Python chain below is the readable specification
        # (OQUPY_B200_PYCHAIN=1, and the CPU-only host-logic tests)
        self._native = (getattr(self._ops, "name", "") == "cuda"
                        and os.environ.get("OQUPY_B200_PYCHAIN", "0") != "1")
        self._chain = None
        self._infl = None        # influence matrices by dk (host, d2 x d2)
        self._dense0 = None      # (w, n, s, e) dk=0 site incl. unitary transform

    @property
    def step(self):
        """The current step in the TEMPO computation."""
        return self._step

    def initialize_mps_mpo(self):
        """:379-437"""
        ops = self._ops
        self._initial_state = copy(self._initial_state).reshape(-1)
        d2 = self._initial_state.shape[0]
        u = self._unitary_transform
        ud = u.conjugate().T
        super_u = np.kron(u, ud.T)           # operators.left_right_super(u, u^dagger)
        super_u_dagg = np.kron(ud, u.T)
        npre = 1 if self._dkmax is None else self._dkmax + 1
        self._infl = [np.asarray(self._influence(i), dtype=CDTYPE)
                      for i in range(npre)]
        # dk = 0 site: B[w,n,s,e] = d_we d_ns infl0[n,w], then the unitary
        # transform on legs n and e (:419-424)
        infl0 = self._infl[0]
        if self._maps is not None:           # :408-417 (unique=True): reduced w and s legs
            nmap, wmap = self._maps
            nn, nw = self._sum_north.size, self._sum_west.size
            b0 = np.zeros((nw, d2, nn, d2), dtype=CDTYPE)
            idx = np.arange(d2)
            b0[wmap, idx, nmap, idx] = infl0.reshape(-1)[nmap]
        else:
            nn = nw = d2
            b0 = np.zeros((d2, d2, d2, d2), dtype=CDTYPE)
            for n in range(d2):
                for w in range(d2):
                    b0[w, n, n, w] = infl0[n, w]
        b0 = np.einsum("wnse,nm->wmse", b0, super_u_dagg)
        b0 = np.einsum("wmse,fe->wmsf", b0, super_u)
        self._dense0 = b0
        self._nn, self._nw = nn, nw
        self._dense0_dev = ops.from_host(b0.reshape(nw * d2, nn * d2))
        self._dense0_west = ops.from_host(
            np.tensordot(self._sum_west, b0, (0, 0)).reshape(d2, nn * d2))
        self._infl_dev = {}
        self._sn_dev = ops.from_host(self._sum_north)
        self._d2 = d2
        self._mps = [ops.from_host(self._initial_state.reshape(1, d2, 1))]
        if self._native:
            self._chain = NativeChain(ops)
            self._chain.push(self._mps[0])
            self._mps = None

    def _mid_site(self, dk, start):
        key = (dk, start)
        if key not in self._infl_dev:
            mat = self._infl[dk]
            if start:
                mat = mat * self._sum_west[None, :]
            self._infl_dev[key] = self._ops.from_host(mat)
        return TempoSite("start" if start else "mid", self._infl_dev[key])

    def compute_system_step(self, current_step, prop_1, prop_2):
        """One TEMPO step (:439-575); returns the state as a host (d2,) vector."""
        ops, d2 = self._ops, self._d2
        # -- which influence functions take part (:487-516)
        if self._dkmax is None:
            dks = list(range(len(self._infl) - 1, -1, -1))
            self._infl.append(np.asarray(self._influence(len(self._infl)),
                                         dtype=CDTYPE))
            override = None
        elif current_step <= self._dkmax:
            dks = list(range(current_step - 1, -1, -1))
            override = None
        else:
            dks = list(range(self._dkmax, -1, -1))
            override = self._influence(self._dkmax - current_step)
        mpo = []
        for pos, dk in enumerate(dks):
            if dk == 0:
                if pos == 0:
                    mpo.append(TempoSite("dense", self._dense0_west, nw=1, ns=self._nn))
                else:
                    mpo.append(TempoSite("dense", self._dense0_dev, nw=self._nw,
                                         ns=self._nn))
            elif pos == 0 and override is not None:
                mat = np.asarray(override, dtype=CDTYPE) * self._sum_west[None, :]
                mpo.append(TempoSite("start", ops.from_host(mat)))
            else:
                mpo.append(self._mid_site(dk, pos == 0))
        if self._native:
            state = ops.empty(1, d2)
            self._chain.tempo_step(
                mpo, ops.from_host(prop_1),
                ops.from_host(np.asarray(prop_2, dtype=CDTYPE).T.reshape(d2, d2, 1)),
                self._sn_dev, d2, self._epsrel, state)
            return ops.to_host(state).reshape(-1)
        mps = self._mps
        # -- first half propagator on the newest site (:521-529)
        last = mps[-1]
        nl = last.shape[0]
        p1 = ops.from_host(prop_1)
        new_last = ops.empty(nl, d2, 1)
        ops.gemm(nl, d2, d2, View(last, row=d2, col=1), View(p1, row=1, col=d2),
                 View(new_last, row=d2, col=1))
        mps[-1] = new_last
        # -- sum out the oldest leg beyond the memory cut-off (:531-537)
        if len(mps) != len(mpo):
            first = mps[0]
            _, na, nr = first.shape
            vec = ops.empty(1, nr)
            ops.gemm(1, nr, na, View(self._sn_dev, col=1),
                     View(first, row=nr, col=1), View(vec, col=1))
            second = mps[1]
            _, sa, sr = second.shape
            merged = ops.empty(1, sa, sr)
            ops.gemm(1, sa * sr, nr, View(vec, col=1),
                     View(second, row=sa * sr, col=1), View(merged, col=1))
            mps[1] = merged
            del mps[0]
        chain.tempo_zip_up_right(ops, mps, mpo, self._epsrel)        # :539-547
        chain.svd_sweep_left(ops, mps, len(mps) - 1, 0, self._epsrel)  # :549-553
        p2 = ops.from_host(np.asarray(prop_2, dtype=CDTYPE).T.reshape(d2, d2, 1))
        mps.append(p2)                                                # :555-558
        # -- read-out (:560-573): vec <- sum_a sn[a] (vec . A[:, a, :])
        vec = ops.one
        for site in mps[:-1]:
            nl, na, nr = site.shape
            tmp = ops.empty(na, nr)
            ops.gemm(1, na * nr, nl, View(vec, col=1),
                     View(site, row=na * nr, col=1), View(tmp, col=1))
            nxt = ops.empty(1, nr)
            ops.gemm(1, nr, na, View(self._sn_dev, col=1),
                     View(tmp, row=nr, col=1), View(nxt, col=1))
            vec = nxt
        state = ops.empty(1, d2)
        ops.gemm(1, d2, d2, View(vec, col=1), View(mps[-1], row=d2, col=1),
                 View(state, col=1))
        return ops.to_host(state).reshape(-1)

    def get_bond_dimensions(self):
        if self._native:
            return [self._chain.shape(k)[2] for k in range(len(self._chain) - 1)]
        return [int(t.shape[2]) for t in self._mps[:-1]]


class TempoBackend(BaseTempoBackend):
    """tempo_backend.py:578-626"""

    def __init__(self, initial_state, influence, unitary_transform, propagators,
                 sum_north, sum_west, dkmax, epsrel, config=None,
                 degeneracy_maps=None, dim=None, ops=None):
        super().__init__(initial_state, influence, unitary_transform, sum_north,
                         sum_west, dkmax, epsrel, config, degeneracy_maps, dim,
                         ops=ops)
        self._propagators = propagators

    def initialize(self):
        self._step = 0
        self.initialize_mps_mpo()
        self._state = self._initial_state
        return self._step, copy(self._state)

    def compute_step(self):
        self._step += 1
        prop_1, prop_2 = self._propagators(self._step - 1)
        self._state = self.compute_system_step(self._step, prop_1, prop_2)
        return self._step, copy(self._state)


class MeanFieldTempoBackend:
    """One or more TEMPO networks with a coherent mean field (tempo_backend.py:629-773).

    The systems of a mean-field model advance in LOCK-STEP: when they share the Hilbert-space
    dimension and a finite ``dkmax`` (and use full legs, no correlation-time tail) all of
    them are members of ONE :class:`oqupy_b200.batch.BatchedTempoBackend` and a time step of
    the whole model is one kernel launch (one CTA per system) instead of the reference's
    host loop over per-system backends (:761-764).  Otherwise (``unique=True`` maps,
    ``dkmax=None``, mixed dimensions) every system keeps its own device-resident
    :class:`BaseTempoBackend`.  The field equation of motion stays a host callback, as the
    propagators do (they depend on the field of the step, :755-759)."""

    def __init__(self, initial_state_list, initial_field, influence_list,
                 unitary_transform_list, propagators_list, compute_field,
                 compute_field_derivative, sum_north_list, sum_west_list, dkmax,
                 epsrel, config=None, degeneracy_maps_list=None, dim_list=None,
                 ops=None):
        n = len(initial_state_list)
        if degeneracy_maps_list is None:
            degeneracy_maps_list = [None] * n
        if dim_list is None:
            dim_list = [None] * n
        self._initial_state_list = initial_state_list
        self._initial_field = initial_field
        self._compute_field = compute_field
        self._compute_field_derivative = compute_field_derivative
        self._field = initial_field
        self._state_list = initial_state_list
        self._step = None
        self._propagators_list = propagators_list
        self._batched = None
        self._backend_list = None
        ops_ = default_ops() if ops is None else ops
        sizes = {np.asarray(s).size for s in initial_state_list}
        lockstep = (getattr(ops_, "name", "") == "cuda" and dkmax is not None and dkmax >= 1
                    and len(sizes) == 1
                    and all(m is None for m in degeneracy_maps_list)
                    and os.environ.get("OQUPY_B200_PYCHAIN", "0") != "1")
        if lockstep:
            d2 = sizes.pop()
            lockstep = (all(len(sn) == d2 and len(sw) == d2 and np.all(np.asarray(sn) == 1)
                            and np.all(np.asarray(sw) == 1)
                            for sn, sw in zip(sum_north_list, sum_west_list))
                        and all(infl(-1) is None for infl in influence_list))
        if lockstep:
            from .batch import MAX_OPERAND, BatchedTempoBackend  # pylint: disable=import-outside-toplevel
            infl = np.array([[np.asarray(f(dk), dtype=CDTYPE) for dk in range(dkmax + 1)]
                             for f in influence_list])
            self._batched = BatchedTempoBackend(
                np.array([np.asarray(s, dtype=CDTYPE).reshape(-1) for s in initial_state_list]),
                infl, np.array([np.asarray(u, dtype=CDTYPE) for u in unitary_transform_list]),
                None, np.ones(d2), np.ones(d2), dkmax, epsrel,
                chi_cap=MAX_OPERAND // d2, ops=ops_)
            return
        self._backend_list = [
            BaseTempoBackend(state, influence, unitary, sum_north, sum_west, dkmax,
                             epsrel, config, maps, dim, ops=ops)
            for state, influence, unitary, sum_north, sum_west, maps, dim in zip(
                initial_state_list, influence_list, unitary_transform_list,
                sum_north_list, sum_west_list, degeneracy_maps_list, dim_list)]

    @property
    def step(self):
        """The current step in the TEMPO computation."""
        return self._step

    def initialize(self):
        """:740-745"""
        self._step = 0
        if self._batched is not None:
            self._batched.initialize()
            self._state_list = [copy(s).reshape(-1) for s in self._initial_state_list]
            return self._step, deepcopy(self._state_list), self._field
        for backend in self._backend_list:
            backend.initialize_mps_mpo()
        return self._step, deepcopy(self._state_list), self._field

    def compute_step(self):
        """:747-773"""
        current_step = self._step
        next_step = current_step + 1
        current_state_list = deepcopy(self._state_list)
        current_field = self._field
        current_field_derivative = self._compute_field_derivative(
            current_step, current_state_list, current_field)
        # the field enters each system's dynamics through its propagators
        prop_tuple_list = [propagators(current_step, current_field,
                                       current_field_derivative)
                           for propagators in self._propagators_list]
        if self._batched is not None:      # every system in ONE launch
            _, states = self._batched.compute_step_with(
                np.array([np.asarray(p[0], dtype=CDTYPE) for p in prop_tuple_list]),
                np.array([np.asarray(p[1], dtype=CDTYPE) for p in prop_tuple_list]))
            next_state_list = [states[i].copy() for i in range(states.shape[0])]
        else:
            next_state_list = [backend.compute_system_step(next_step, *prop_tuple)
                               for backend, prop_tuple in zip(self._backend_list,
                                                              prop_tuple_list)]
        next_field = self._compute_field(current_step, current_state_list,
                                         current_field, next_state_list)
        self._state_list = next_state_list
        self._field = next_field
        self._step = next_step
        return self._step, deepcopy(self._state_list), self._field

    def get_bond_dimensions(self):
        if self._batched is not None:
            return self._batched.get_bond_dimensions()
        return [backend.get_bond_dimensions() for backend in self._backend_list]
