"""Lock-step ensembles of TEMPO runs on one GPU (BASELINE configs[4]).

:class:`BatchedTempoBackend` is ``oqupy.backends.tempo_backend.TempoBackend``
(/root/reference/oqupy/backends/tempo_backend.py:578-626) with a leading ensemble dimension:
E members that share the system dimension, ``dkmax`` and ``epsrel`` but have their own
influence functions (coupling strength, temperature, spectral density), propagators and
initial states.  One ``compute_step`` advances ALL members with ONE kernel launch
(csrc/batch.cu: one CTA per member, its chain and every truncated SVD stay on the device);
nothing is read back between the sites, the states come back once per call.

There is no CPU fallback: constructing the backend needs the CUDA library.
"""
import ctypes
from ctypes import c_int32, c_void_p

import numpy as np

from ._lib import B200Error, default_ops

CDTYPE = np.complex128
MAX_OPERAND = 128           # chi * d2 must fit the shared-memory resident SVD (csrc/batch.cu)


def _dense0(infl0, unitary, d2):
    """The dk = 0 MPO site incl. the unitary transform, (w, n, s, e)
    (tempo_backend.py:383-388, 419-424)."""
    u = np.asarray(unitary, dtype=CDTYPE)
    ud = u.conjugate().T
    super_u = np.kron(u, ud.T)
    super_u_dagg = np.kron(ud, u.T)
    b0 = np.zeros((d2, d2, d2, d2), dtype=CDTYPE)
    idx = np.arange(d2)
    for w in range(d2):
        b0[w, idx, idx, w] = infl0[idx, w]
    b0 = np.einsum("wnse,nm->wmse", b0, super_u_dagg)
    return np.einsum("wmse,fe->wmsf", b0, super_u)


class BatchedTempoBackend:
    """E TEMPO runs in lock-step.

    initial_states   (E, d2) or (E, d, d)
    influences       (E, dkmax+1, d2, d2): influence matrices by dk of every member, exactly
                     what each member's ``influence(dk)`` callback returns
                     (oqupy/tempo.py:969-1020); dk < 0 (``add_correlation_time``) is not
                     supported in lock-step
    unitary_transforms  (d, d) for all members or (E, d, d)
    propagators      ``step -> (P1, P2)``, each (d2, d2) (shared) or (E, d2, d2)
    sum_north, sum_west  (d2,) as for TempoBackend (unique=False)
    chi_cap          largest bond dimension a member may reach (default: the largest that
                     keeps every SVD operand in shared memory, 104 // d2); a member that
                     would exceed it stops with status 2 and ``compute_step`` raises
    """

    def __init__(self, initial_states, influences, unitary_transforms, propagators, sum_north,
                 sum_west, dkmax, epsrel, chi_cap=None, ops=None):
        self._ops = default_ops() if ops is None else ops
        infl = np.asarray(influences, dtype=CDTYPE)
        assert infl.ndim == 4 and dkmax is not None and dkmax >= 1
        self.E, _, self.d2, _ = infl.shape
        assert infl.shape[1] == dkmax + 1 and infl.shape[2] == infl.shape[3]
        self._infl = infl
        st = np.asarray(initial_states, dtype=CDTYPE).reshape(self.E, -1)
        assert st.shape[1] == self.d2
        self._state0 = st
        ut = np.asarray(unitary_transforms, dtype=CDTYPE)
        self._unitary = np.broadcast_to(ut, (self.E,) + ut.shape[-2:])
        self._propagators = propagators
        self._sum_north = np.asarray(sum_north, dtype=float)
        self._sum_west = np.asarray(sum_west, dtype=float)
        assert self._sum_north.size == self.d2 and self._sum_west.size == self.d2, \
            "lock-step ensembles take full (unique=False) legs"
        self._dkmax = int(dkmax)
        self._epsrel = float(epsrel)
        self._chi_cap = MAX_OPERAND // self.d2 if chi_cap is None else int(chi_cap)
        if self._chi_cap * self.d2 > MAX_OPERAND or self._chi_cap < self.d2:
            raise ValueError(f"chi_cap * d2 must be <= {MAX_OPERAND} and chi_cap >= d2")
        self._h = None
        self._step = None
        self._prop_cache = None
        self._states_dev = None

    @property
    def step(self):
        return self._step

    def __del__(self):
        try:
            if self._h:
                self._ops.lib.b200_tempo_batch_destroy(c_void_p(self._h))
                self._h = None
        except Exception:  # pylint: disable=broad-except
            pass

    def initialize(self):
        """Upload the per-member tables; returns (0, states (E, d2))."""
        ops, lib = self._ops, self._ops.lib
        e_, d2, k = self.E, self.d2, self._dkmax
        self._h = lib.b200_tempo_batch_create(ops._stream(), e_, d2, k, self._chi_cap,  # pylint: disable=protected-access
                                              self._epsrel)
        if not self._h:
            raise B200Error("b200_tempo_batch_create failed: "
                            f"{lib.b200_last_error().decode()}")
        mid = self._infl                                         # [e, dk, s, e']
        start = self._infl * self._sum_west[None, None, None, :]
        dense0 = np.empty((e_, d2 * d2, d2 * d2), dtype=CDTYPE)
        dense0w = np.empty((e_, d2, d2 * d2), dtype=CDTYPE)
        cache = {}
        for i in range(e_):
            key = (self._infl[i, 0].tobytes(), self._unitary[i].tobytes())
            if key not in cache:
                b0 = _dense0(self._infl[i, 0], self._unitary[i], d2)
                cache[key] = (b0.reshape(d2 * d2, d2 * d2),
                              np.tensordot(self._sum_west, b0, (0, 0)).reshape(d2, d2 * d2))
            dense0[i], dense0w[i] = cache[key]
        dev = [ops.from_host(a) for a in (mid, start, dense0, dense0w,
                                          self._sum_north.astype(CDTYPE), self._state0)]
        ops._check(lib.b200_tempo_batch_set(c_void_p(self._h), *[c_void_p(t.data_ptr()) for t in dev]),  # pylint: disable=protected-access
                   "b200_tempo_batch_set")
        self._states_dev = ops.empty(e_, d2)
        self._step = 0
        return 0, self._state0.copy()

    def _props(self, step):
        p1, p2 = self._propagators(step)
        key = (id(p1), id(p2))
        if self._prop_cache is None or self._prop_cache[0] != key:
            a = np.broadcast_to(np.asarray(p1, dtype=CDTYPE), (self.E, self.d2, self.d2))
            b = np.broadcast_to(np.asarray(p2, dtype=CDTYPE), (self.E, self.d2, self.d2))
            bt = np.ascontiguousarray(np.swapaxes(b, 1, 2))
            self._prop_cache = (key, self._ops.from_host(a), self._ops.from_host(bt), p1, p2)
        return self._prop_cache[1], self._prop_cache[2]

    def _launch(self, out):
        p1, p2t = self._props(self._step)
        self._step += 1
        self._ops._check(self._ops.lib.b200_tempo_batch_step(  # pylint: disable=protected-access
            c_void_p(self._h), c_void_p(p1.data_ptr()), c_void_p(p2t.data_ptr()),
            c_void_p(out.data_ptr())), "b200_tempo_batch_step")

    def compute_step_with(self, prop_1, prop_2):
        """One time step with explicitly given propagators, each (E, d2, d2) or (d2, d2)
        (mean-field systems: the propagators of a step depend on the field of that step,
        tempo_backend.py:755-764).  Returns (step, states (E, d2)) on the host."""
        saved = self._propagators
        self._propagators = lambda step: (prop_1, prop_2)
        self._prop_cache = None
        try:
            return self.compute_step()
        finally:
            self._propagators = saved
            self._prop_cache = None

    def compute_step(self):
        """One time step of every member; returns (step, states (E, d2)) on the host."""
        self._launch(self._states_dev)
        states = self._ops.to_host(self._states_dev)
        self.check()
        return self._step, states

    def compute_steps(self, num_steps, strict=True):
        """``num_steps`` time steps back to back (one launch each, no host round trip in
        between); returns the states (num_steps, E, d2) with ONE device-to-host copy.
        ``strict=False``: members that left the device path (see :meth:`info`) do not raise;
        their states are undefined from the step at which they stopped."""
        buf = self._ops.empty(num_steps, self.E, self.d2)
        for k in range(num_steps):
            self._launch(buf[k])
        out = self._read_back(buf)
        if strict:
            self.check()
        return out

    def set_order(self, order=None):
        """CTA i of the following launches takes member ``order[i]`` (None: identity).  A
        scheduling hint only -- every member is still advanced exactly once per step."""
        if order is None:
            arr = None
        else:
            arr = (c_int32 * self.E)(*[int(x) for x in order])
        self._ops._check(self._ops.lib.b200_tempo_batch_set_order(c_void_p(self._h), arr),  # pylint: disable=protected-access
                         "b200_tempo_batch_set_order")

    def reserve_sms(self, n):
        """Leave ``n`` SMs to other streams in the following steps (0: none)."""
        self._ops._check(self._ops.lib.b200_tempo_batch_reserve_sms(c_void_p(self._h), int(n)),  # pylint: disable=protected-access
                         "b200_tempo_batch_reserve_sms")

    def rebalance(self):
        """Longest-processing-time-first: hand the members with the largest bond dimensions
        (cost ~ chi^3 per SVD) to the first CTAs of every launch, so that the last wave of a
        step does not wait for one heavy member.  Returns the order."""
        chi = self.info()["max_chi"].astype(np.int64)
        order = np.argsort(-chi, kind="stable")
        self.set_order(order)
        return order

    def _read_back(self, buf):
        """Device -> host WITHOUT holding the interpreter lock while the launches drain (an
        asynchronous copy into pinned memory, then a sleeping poll on an event): members that
        are re-run on the general backend in other host threads of this process (tempo_grid)
        keep stepping while the batch runs.  A blocking ``tensor.cpu()`` starved them: measured
        on the 4096-member grid, the 5.3 s re-run of one member made no progress during 17-37 s
        of batch time."""
        import time  # pylint: disable=import-outside-toplevel
        import torch  # pylint: disable=import-outside-toplevel
        if getattr(self._ops, "name", "") != "cuda":
            return self._ops.to_host(buf)
        host = torch.empty(buf.shape, dtype=buf.dtype, pin_memory=True)
        host.copy_(buf, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(buf.device))
        while not done.query():
            time.sleep(0.0005)
        self._ops.d2h_bytes += buf.numel() * buf.element_size()
        return host.numpy().copy()

    def info(self):
        """dict of per-member arrays: status, svds, sweeps, max_chi, bonds (list of lists)."""
        e_ = self.E
        nb = self._dkmax + 2
        st, sv, sw, mc = ((c_int32 * e_)() for _ in range(4))
        bonds = (c_int32 * (e_ * nb))()
        self._ops._check(self._ops.lib.b200_tempo_batch_info(c_void_p(self._h), st, sv, sw, mc, bonds),  # pylint: disable=protected-access
                         "b200_tempo_batch_info")
        b = np.ctypeslib.as_array(bonds).reshape(e_, nb)
        return {"status": np.array(st), "svds": np.array(sv), "sweeps": np.array(sw),
                "max_chi": np.array(mc), "bonds": [[int(x) for x in row if x >= 0] for row in b]}

    def check(self):
        """Raise if any member left the device path (capacity, convergence)."""
        status = self.info()["status"]
        bad = np.nonzero(status)[0]
        if bad.size:
            what = {2: "bond dimension / operand exceeds chi_cap", 3: "Jacobi did not converge"}
            raise B200Error(
                f"lock-step TEMPO: member {int(bad[0])} stopped with status {int(status[bad[0]])}"
                f" ({what.get(int(status[bad[0]]), 'internal')}); {bad.size} member(s) affected")

    def get_bond_dimensions(self):
        return self.info()["bonds"]

    def device_bytes(self):
        return int(self._ops.lib.b200_tempo_batch_bytes(c_void_p(self._h)))
