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
    """PT-MPO on the device: rank-3 sites (past bond, future bond, array leg).  A
    non-diagonal system-bath coupling keeps the rank-3 sites of the diagonalised problem
    plus the pair ``transform_in`` / ``transform_out`` (d2 x d2, oqupy/pt_tempo.py:159-167);
    the device loop folds them into the system propagators (see :func:`fold_transforms`)
    instead of expanding every site to four dense legs (process_tensor.py:346-355)."""

    def __init__(self, hilbert_space_dimension, dt=None, transform_in=None,
                 transform_out=None, name=None, description=None, ops=None):
        d2 = hilbert_space_dimension ** 2
        self._transform_in = self._transform_out = None
        if transform_in is not None:
            self._transform_in = np.array(transform_in, dtype=CDTYPE)
            assert self._transform_in.shape == (d2, d2)
        if transform_out is not None:
            self._transform_out = np.array(transform_out, dtype=CDTYPE)
            assert self._transform_out.shape == (d2, d2)
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

    @property
    def transform_in(self):
        return self._transform_in

    @property
    def transform_out(self):
        return self._transform_out

    def get_mpo_tensor(self, step, transformed=True):
        """Host copy (process_tensor.py:326-355): rank-3 (past, future, array); with
        transforms and ``transformed`` the dense 4-leg tensor
        ``T4[l,r,a,b] = sum_x tin[a,x] T3[l,r,x] tout[x,b]`` exactly as the reference forms it
        (diagnostics / protocol fidelity: the device loop never calls this)."""
        if step >= len(self._sites) or step < 0:
            raise IndexError("Process tensor index out of bound. ")
        t3 = self._ops.to_host(self._sites[step])
        if not transformed or (self._transform_in is None and self._transform_out is None):
            return t3
        d2 = t3.shape[2]
        tin = np.identity(d2) if self._transform_in is None else self._transform_in
        tout = np.identity(d2) if self._transform_out is None else self._transform_out
        return np.einsum("ax,lrx,xb->lrab", tin, t3, tout)

    def get_mpo_tensor_device(self, step):
        return self._sites[step]

    def get_mpo_tensor_device_swapped(self, step):
        """The site with its bond legs swapped, ``(chi_k+1, chi_k, d2)`` (the MPO the
        back-propagation of the gradient runs through, system_dynamics.py:588-628); a
        device copy made on first use and kept."""
        cache = self.__dict__.setdefault("_sites_swapped", {})
        t = self._sites[step]
        hit = cache.get(step)
        if hit is None or hit[0] is not t:
            hit = (t, t.permute(1, 0, 2).contiguous())
            cache[step] = hit
        return hit[1]

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

    # -- persistence -------------------------------------------
    def export(self, filename):
        """Write the PT-MPO to ``filename`` (numpy ``.npz`` container, uncompressed): the
        side-format of this package for the reference's HDF5 ``FileProcessTensor``
        (process_tensor.py:501-559, 801-823; h5py is not part of this image).  Same
        content: dimension, dt, name, description, the rank-3 sites ``(chi_k, chi_k+1, d2)``
        and the cap vectors.  Sites are copied device -> host one at a time."""
        data = {"format": "oqupy_b200-pt-1", "hilbert_space_dimension": self._hs_dim,
                "dt": np.nan if self._dt is None else float(self._dt),
                "name": "" if self.name is None else str(self.name),
                "description": "" if self.description is None else str(self.description),
                "num_sites": len(self._sites), "num_caps": len(self._caps)}
        for k, t in enumerate(self._sites):
            data[f"mpo_{k}"] = self._ops.to_host(t)
        for k, c in enumerate(self._caps):
            data[f"cap_{k}"] = self._ops.to_host(c)
        with open(filename, "wb") as f:
            np.savez(f, **data)

    def export_to(self, process_tensor):
        """Fill a reference process tensor (``SimpleProcessTensor`` or the HDF5-backed
        ``FileProcessTensor``) through its own ``set_mpo_tensor`` / ``set_cap_tensor``
        (process_tensor.py:291-324): the route to the reference's on-disk format where h5py
        is installed."""
        for k, t in enumerate(self._sites):
            process_tensor.set_mpo_tensor(k, self._ops.to_host(t))
        for k, c in enumerate(self._caps):
            process_tensor.set_cap_tensor(k, self._ops.to_host(c))
        return process_tensor

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


def as_device_process_tensor(pt, ops=None):
    """``pt`` on the device, or None when the device path does not cover it: a
    DeviceProcessTensor as is; a host process tensor with rank-3 sites, no transforms and no
    initial tensor (what PT-TEMPO produces for a diagonalised coupling,
    process_tensor.py:249-430) is uploaded once and the copy kept on the object -- its
    public ``get_mpo_tensor`` would expand every site to four legs in an interpreted loop
    (process_tensor.py:346-347, util.py:30-57) on every call."""
    if isinstance(pt, DeviceProcessTensor):
        return pt
    sites = getattr(pt, "_mpo_tensors", None)
    if sites is None or len(sites) == 0 or pt.get_initial_tensor() is not None:
        return None
    if any(t is None or getattr(t, "ndim", 0) != 3 for t in sites):
        return None
    caps = getattr(pt, "_cap_tensors", None)
    if not caps or len(caps) != len(sites) + 1 or any(c is None for c in caps):
        return None          # no caps: the reference raises (system_dynamics.py:562-575)
    # the cached device copy is valid for exactly these site / cap arrays (a later
    # set_mpo_tensor / set_cap_tensor replaces the array object)
    key = tuple(id(t) for t in sites) + tuple(id(c) for c in caps)
    cached = getattr(pt, "_b200_device", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    dev = DeviceProcessTensor(pt.hilbert_space_dimension, dt=pt.dt,
                              transform_in=getattr(pt, "_transform_in", None),
                              transform_out=getattr(pt, "_transform_out", None), ops=ops)
    for k, t in enumerate(sites):
        dev.set_mpo_tensor(k, t)
    dev._caps = [dev._ops.from_host(np.asarray(c).reshape(-1)) for c in caps]  # pylint: disable=protected-access
    try:
        pt._b200_device = (key, dev)   # pylint: disable=protected-access
    except AttributeError:
        pass
    return dev


def fold_transforms(pts, propagators):
    """Propagators that carry the coupling transforms of ONE-environment process tensors.

    With T4[l,r,a,b] = sum_x tin[a,x] T3[l,r,x] tout[x,b] (process_tensor.py:349-354) and the
    conventions v'[j] = sum_i P[j,i] v[i], v'[r,b] = sum_{l,a} v[l,a] T4[l,r,a,b]
    (system_dynamics.py:631-640, 689-700), a step  P2 . T4 . P1  equals
    (P2 tout^T) . T3 . (tin^T P1): the rank-3 device kernel with two d2 x d2 products on the
    host per distinct propagator pair."""
    pt = pts[0]
    tin, tout = pt.transform_in, pt.transform_out
    if tin is None and tout is None:
        return propagators
    cache = {}

    def folded(step):
        p1, p2 = propagators(step)
        key = (id(p1), id(p2))
        if key not in cache:
            cache.clear()
            q1 = np.asarray(p1, dtype=CDTYPE)
            q2 = np.asarray(p2, dtype=CDTYPE)
            if tin is not None:
                q1 = tin.T @ q1
            if tout is not None:
                q2 = q2 @ tout.T
            cache[key] = (q1, q2, p1, p2)
        return cache[key][0], cache[key][1]
    return folded


def import_process_tensor(filename, ops=None):
    """Read a file written by :meth:`DeviceProcessTensor.export` straight onto the device
    (counterpart of ``oqupy.import_process_tensor``, process_tensor.py:801-823)."""
    with np.load(filename, allow_pickle=False) as f:
        if str(f["format"]) != "oqupy_b200-pt-1":
            raise ValueError(f"{filename}: not an oqupy_b200 process tensor file")
        dt = float(f["dt"])
        pt = DeviceProcessTensor(int(f["hilbert_space_dimension"]),
                                 dt=None if np.isnan(dt) else dt,
                                 name=str(f["name"]) or None,
                                 description=str(f["description"]) or None, ops=ops)
        for k in range(int(f["num_sites"])):
            pt.set_mpo_tensor(k, f[f"mpo_{k}"])
        caps = [pt._ops.from_host(f[f"cap_{k}"]) for k in range(int(f["num_caps"]))]
    pt._caps = caps
    return pt


def dynamics_device(pt, propagators, initial_states, num_steps=None, ops=None):
    """compute_dynamics hot loop (system_dynamics.py:131-170).

    ``pt``: one process tensor, or a list of them (one per environment,
    system_dynamics.py:689-700: the state carries one bond leg per environment).
    With ONE environment, ``E`` ensemble members may share the process tensor:
    propagators(step) -> (P1, P2), each (d2, d2) or (E, d2, d2); initial_states (d, d) or
    (E, d, d); returns ndarray (E, num_steps+1, d, d) (E squeezed for a single state).
    With several environments: one state, (d2, d2) propagators, returns
    (num_steps+1, d, d).
    """
    ops = default_ops() if ops is None else ops
    if isinstance(pt, (list, tuple)):
        if len(pt) > 1:
            return _dynamics_multi_env(list(pt), propagators, initial_states, num_steps,
                                       ops)
        pt = pt[0]
    propagators = fold_transforms([pt], propagators)
    rho0 = np.asarray(initial_states, dtype=CDTYPE)
    single = rho0.ndim == 2
    if single:
        rho0 = rho0[None]
    nvec, d = rho0.shape[0], rho0.shape[1]
    d2 = d * d
    if num_steps is None:
        num_steps = len(pt)
    for k in range(num_steps):       # an out-of-bounds device read otherwise
        if int(pt.get_mpo_tensor_device(k).shape[2]) != d2:
            raise ValueError(f"process tensor site {k} has array leg "
                             f"{int(pt.get_mpo_tensor_device(k).shape[2])}, the state needs {d2}")
    v = ops.from_host(rho0.reshape(nvec, 1, d2))
    rho = ops.empty(num_steps + 1, nvec, d2)
    if hasattr(ops, "dyn_run"):
        # native loop: all propagators are gathered on the host first (time-independent
        # systems hand back the same pair every step -> stride 0), ONE C-ABI call then
        # issues every launch back to back
        pairs = [propagators(step) for step in range(num_steps)]
        same = all(p[0] is pairs[0][0] and p[1] is pairs[0][1] for p in pairs)
        sel = pairs[:1] if same else pairs
        p1 = np.stack([np.broadcast_to(np.asarray(p[0], dtype=CDTYPE), (nvec, d2, d2))
                       for p in sel])
        p2 = np.stack([np.broadcast_to(np.asarray(p[1], dtype=CDTYPE), (nvec, d2, d2))
                       for p in sel])
        sites = [pt.get_mpo_tensor_device(k) for k in range(num_steps)]
        caps = [pt.get_cap_tensor_device(k) for k in range(num_steps + 1)]
        keep = ops.dyn_run(nvec, d2, sites, caps, ops.from_host(p1), ops.from_host(p2),
                           0 if same else nvec * d2 * d2, v, rho)
        out = ops.to_host(rho).transpose(1, 0, 2).reshape(nvec, num_steps + 1, d, d)
        del keep
        return out[0] if single else out
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


def _dynamics_multi_env(pts, propagators, initial_state, num_steps, ops):
    """m >= 2 environments: the state is (chi_1, ..., chi_m, d2); each environment's
    rank-3 PT-MPO site contracts its own bond leg, diagonal in the system leg x
    (system_dynamics.py:689-700 with T4[l,r,x,x'] = T3[l,r,x] d_xx')."""
    from ._lib import View  # pylint: disable=import-outside-toplevel
    rho0 = np.asarray(initial_state, dtype=CDTYPE)
    if rho0.ndim != 2:
        raise NotImplementedError(
            "oqupy_b200: ensembles over several environments are not supported yet")
    m = len(pts)
    d = rho0.shape[0]
    d2 = d * d
    if num_steps is None:
        num_steps = min(len(p) for p in pts)
    chis = [1] * m
    v = ops.from_host(rho0.reshape(1, d2))
    rho = ops.empty(num_steps + 1, d2)

    def readout(state, dims, step, dst):
        """rho = sum_{l_1..l_m} cap_1[l_1] .. cap_m[l_m] v[l_1..l_m, :]  (:643-651)"""
        cur = state
        rest = int(np.prod(dims)) * d2
        for i in range(m):
            rest //= dims[i]
            cap = pts[i].get_cap_tensor_device(step)
            out = dst if i == m - 1 else ops.empty(rest)
            ops.gemm(1, rest, dims[i], View(cap, col=1), View(cur, row=rest, col=1),
                     View(out, col=1))
            cur = out

    tcache = {}

    def system_leg_product(state, mat):
        """state[..., j] <- sum_i mat[j, i] state[..., i]"""
        key = id(mat)
        if key not in tcache:
            tcache[key] = (ops.from_host(np.ascontiguousarray(mat)), mat)
        rows_ = state.numel() // d2
        out = ops.empty(rows_, d2)
        ops.gemm(rows_, d2, d2, View(state, row=d2, col=1), View(tcache[key][0], row=1, col=d2),
                 View(out, row=d2, col=1))
        return out

    cache = {}
    for step in range(num_steps):
        readout(v, chis, step, rho[step])
        p1, p2 = propagators(step)
        key = (id(p1), id(p2))
        if key not in cache:
            cache.clear()
            cache[key] = (ops.from_host(np.asarray(p1, dtype=CDTYPE)),
                          ops.from_host(np.asarray(p2, dtype=CDTYPE)), p1, p2)
        dp1, dp2 = cache[key][0], cache[key][1]
        rows = int(np.prod(chis))
        nxt = ops.empty(rows, d2)
        ops.gemm(rows, d2, d2, View(v, row=d2, col=1), View(dp1, row=1, col=d2),
                 View(nxt, row=d2, col=1))                          # v <- v P1^T
        v = nxt
        for i in range(m):
            t = pts[i].get_mpo_tensor_device(step)
            chi_l, chi_r, _ = t.shape
            assert chi_l == chis[i] and int(t.shape[2]) == d2
            if pts[i].transform_in is not None:                     # v <- v tin
                v = system_leg_product(v, pts[i].transform_in.T)
            na = int(np.prod(chis[:i]))                # legs before the contracted one
            nb = int(np.prod(chis[i + 1:]))            # legs after it
            nxt = ops.empty(na * chi_r * nb, d2)
            # out[a, r, b, x] = sum_l T[l, r, x] v[a, l, b, x]; batches: x, a
            ops.gemm(chi_r, nb, chi_l,
                     View(t, row=d2, col=chi_r * d2, b1=1),
                     View(v, row=nb * d2, col=d2, b1=1, b2=chi_l * nb * d2),
                     View(nxt, row=nb * d2, col=d2, b1=1, b2=chi_r * nb * d2),
                     nb1=d2, nb2=na)
            v = nxt
            chis[i] = chi_r
            if pts[i].transform_out is not None:                    # v <- v tout
                v = system_leg_product(v, pts[i].transform_out.T)
        rows = int(np.prod(chis))
        nxt = ops.empty(rows, d2)
        ops.gemm(rows, d2, d2, View(v, row=d2, col=1), View(dp2, row=1, col=d2),
                 View(nxt, row=d2, col=1))                          # v <- v P2^T
        v = nxt
    readout(v, chis, num_steps, rho[num_steps])
    return ops.to_host(rho).reshape(num_steps + 1, d, d)


class _SteppedState:
    """One system's state (chi_1, ..., chi_m, d2) on the device, advanced step by step through
    its m >= 1 process tensors -- the loop body of system_dynamics.py:131-170 / 366-447 for
    callers that need the state between the steps (mean-field systems: the propagators of a
    step depend on the field, the field on the states)."""

    def __init__(self, pts, initial_state, ops):
        from ._lib import View  # pylint: disable=import-outside-toplevel
        self._view = View
        self.pts, self.ops = list(pts), ops
        rho0 = np.asarray(initial_state, dtype=CDTYPE)
        self.d = rho0.shape[0]
        self.d2 = self.d * self.d
        self.dims = [1] * len(self.pts)
        self.v = ops.from_host(rho0.reshape(1, self.d2))
        self._mats = {}

    def sys_leg(self, mat):
        """v[..., j] <- sum_i mat[j, i] v[..., i]  (_apply_system_superoperator, :631-640)"""
        ops, d2, view = self.ops, self.d2, self._view
        dm = ops.from_host(np.ascontiguousarray(np.asarray(mat, dtype=CDTYPE)))
        rows = int(self.v.numel()) // d2
        out = ops.empty(rows, d2)
        ops.gemm(rows, d2, d2, view(self.v, row=d2, col=1), view(dm, row=1, col=d2),
                 view(out, row=d2, col=1))
        self.v = out

    def readout(self, step):
        """rho = sum cap_1[l_1] .. cap_m[l_m] v[l_1..l_m, :]  (_apply_caps, :643-651)"""
        ops, d2, view = self.ops, self.d2, self._view
        cur = self.v
        rest = int(np.prod(self.dims)) * d2
        for e, pt in enumerate(self.pts):
            rest //= self.dims[e]
            cap = pt.get_cap_tensor_device(step)
            out = ops.empty(rest)
            ops.gemm(1, rest, self.dims[e], view(cap, col=1), view(cur, row=rest, col=1),
                     view(out, col=1))
            cur = out
        return ops.to_host(cur).reshape(self.d, self.d)

    def through(self, step):
        """_apply_pt_mpos (:654-700) with rank-3 sites and their transforms"""
        ops, d2, view = self.ops, self.d2, self._view
        for e, pt in enumerate(self.pts):
            t = pt.get_mpo_tensor_device(step)
            chi_l, chi_r, _ = (int(x) for x in t.shape)
            assert chi_l == self.dims[e] and int(t.shape[2]) == d2
            if pt.transform_in is not None:
                self.sys_leg(np.asarray(pt.transform_in).T)
            na = int(np.prod(self.dims[:e]))
            nb = int(np.prod(self.dims[e + 1:]))
            nxt = ops.empty(na * chi_r * nb, d2)
            ops.gemm(chi_r, nb, chi_l, view(t, row=d2, col=chi_r * d2, b1=1),
                     view(self.v, row=nb * d2, col=d2, b1=1, b2=chi_l * nb * d2),
                     view(nxt, row=nb * d2, col=d2, b1=1, b2=chi_r * nb * d2),
                     nb1=d2, nb2=na)
            self.v = nxt
            self.dims[e] = chi_r
            if pt.transform_out is not None:
                self.sys_leg(np.asarray(pt.transform_out).T)


def dynamics_with_field_device(pts_list, propagators_list, initial_state_list, initial_field,
                               field_eom, dt, start_time, num_steps, ops=None,
                               controls_list=None):
    """compute_dynamics_with_field hot loop (system_dynamics.py:366-470): every system of a
    mean-field model is propagated through its own process tensor(s) on the device; per step
    the reduced states come back (d x d each), the field takes its Heun step on the host
    (:324-330) and the field-dependent propagators of the step are uploaded.

    ``pts_list[s]``: list of device process tensors of system s; ``propagators_list[s]``:
    ``(step, field, field_derivative) -> (P1, P2)`` (system.py: TimeDependentSystemWithField);
    ``controls_list[s]``: ``step -> (pre, post)`` or None.  Returns (states, fields):
    states[k][s] the (d, d) state of system s at step k (num_steps + 1 entries), fields[k]."""
    ops = default_ops() if ops is None else ops
    nsys = len(pts_list)
    if controls_list is None:
        controls_list = [None] * nsys
    sts = [_SteppedState(pts, rho0, ops) for pts, rho0 in zip(pts_list, initial_state_list)]

    def heun(t, states, field, next_states):
        rk1 = field_eom(t, states, field)
        rk2 = field_eom(t + dt, next_states, field + rk1 * dt)
        return field + dt * (rk1 + rk2) / 2

    all_states, fields = [], []
    field, prev = initial_field, None
    t = start_time
    for step in range(num_steps + 1):
        t = start_time + step * dt
        ctl = [(None, None) if c is None else c(step) for c in controls_list]
        for st, (pre, _) in zip(sts, ctl):
            if pre is not None:
                st.sys_leg(pre)
        if step == num_steps:
            break
        states = [st.readout(step) for st in sts]
        field = initial_field if step == 0 else heun(t, prev, field, states)
        prev = states
        all_states.append(states)
        fields.append(field)
        for st, (_, post) in zip(sts, ctl):
            if post is not None:
                st.sys_leg(post)
        dfield = field_eom(t, states, field)
        for st, props in zip(sts, propagators_list):
            p1, p2 = props(step, field, dfield)
            st.sys_leg(p1)
            st.through(step)
            st.sys_leg(p2)
    final = [st.readout(num_steps) for st in sts]
    all_states.append(final)
    fields.append(heun(t, prev, field, final))
    return all_states, fields


def _gradient_multi_env(pts, propagators, initial_state, target_derivative, num_steps, ops,
                        controls):
    """gradient_device for m >= 2 environments (gradient.py:266-425 with one bond leg per
    environment; adjoint tensor of system_dynamics.py:702-786 for MPOs that are diagonal in
    the system leg):

        D3[x][i, j] = sum F[l_1..l_m, i] prod_e T_e[l_e, r_e, x] B[r_1..r_m, j].

    Every contraction is a strided, doubly batched GEMM (batches: the system index x and the
    bond legs in front of the contracted one)."""
    from ._lib import View  # pylint: disable=import-outside-toplevel
    m = len(pts)
    if num_steps is None:
        num_steps = min(len(p) for p in pts)
    for p_ in pts:
        if p_.transform_in is not None or p_.transform_out is not None:
            raise NotImplementedError(
                "oqupy_b200: gradients through transformed process tensors are not supported")
    propagators, rho0, controls = _fold_controls(propagators, controls, initial_state)
    d = rho0.shape[0]
    d2 = d * d
    mats = {}

    def dev(mat):
        key = id(mat)
        if key not in mats:
            mats[key] = (ops.from_host(np.ascontiguousarray(np.asarray(mat, dtype=CDTYPE))), mat)
        return mats[key][0]

    def sys_leg(x, mat):                   # x[..., j] <- sum_i mat[j, i] x[..., i]
        rows_ = int(x.numel()) // d2
        out = ops.empty(rows_, d2)
        ops.gemm(rows_, d2, d2, View(x, row=d2, col=1), View(dev(mat), row=1, col=d2),
                 View(out, row=d2, col=1))
        return out

    def through(v, dims, step, swapped):
        """v[a, l, b, x] -> sum_l T_e[l, r, x] v[a, l, b, x] for every environment e (or the
        bond-swapped site, system_dynamics.py:588-628)."""
        dims = list(dims)
        for e in range(m):
            t = pts[e].get_mpo_tensor_device(step)
            chi_l, chi_r, _ = (int(x) for x in t.shape)
            if swapped:
                k_, n_ = chi_r, chi_l
                ta = View(t, row=chi_r * d2, col=d2, b1=1)        # A[l][r] = T[l, r, x]
            else:
                k_, n_ = chi_l, chi_r
                ta = View(t, row=d2, col=chi_r * d2, b1=1)        # A[r][l] = T[l, r, x]
            assert dims[e] == k_
            na = int(np.prod(dims[:e]))
            nb = int(np.prod(dims[e + 1:]))
            nxt = ops.empty(na * n_ * nb, d2)
            ops.gemm(n_, nb, k_, ta,
                     View(v, row=nb * d2, col=d2, b1=1, b2=k_ * nb * d2),
                     View(nxt, row=nb * d2, col=d2, b1=1, b2=n_ * nb * d2), nb1=d2, nb2=na)
            v = nxt
            dims[e] = n_
        return v, dims

    def readout(v, dims, step, dst):
        cur = v
        rest = int(np.prod(dims)) * d2
        for e in range(m):
            rest //= dims[e]
            cap = pts[e].get_cap_tensor_device(step)
            out = dst if e == m - 1 else ops.empty(rest)
            ops.gemm(1, rest, dims[e], View(cap, col=1), View(cur, row=rest, col=1),
                     View(out, col=1))
            cur = out

    # ---- forward, every state kept
    dims = [1] * m
    v = ops.from_host(rho0.reshape(1, d2))
    rho = ops.empty(num_steps + 1, d2)
    forward, fdims = [], []
    for step in range(num_steps):
        readout(v, dims, step, rho[step])
        forward.append(v)
        fdims.append(list(dims))
        p1, p2 = propagators(step)
        v = sys_leg(v, p1)
        v, dims = through(v, dims, step, False)
        v = sys_leg(v, p2)
    readout(v, dims, num_steps, rho[num_steps])
    states = ops.to_host(rho).reshape(num_steps + 1, d, d)
    target = target_derivative(states[-1]) if callable(target_derivative) \
        else target_derivative
    back = ops.from_host(np.asarray(target, dtype=CDTYPE).reshape(1, d2))
    out = ops.empty(num_steps, d2, d2, d2)                    # D3[step][x][i][j]

    def adjoint(step, b, bdims):
        f = forward[step]
        if controls is not None:
            if controls[step][1] is not None:
                f = sys_leg(f, controls[step][1])
            if controls[step + 1][0] is not None:
                b = sys_leg(b, np.asarray(controls[step + 1][0]).T)
        cur, cdims, has_x = f, list(fdims[step]), False
        for e in range(m):
            t = pts[e].get_mpo_tensor_device(step)
            chi_l, chi_r, _ = (int(x) for x in t.shape)
            na = int(np.prod(cdims[:e]))
            nc = int(np.prod(cdims[e + 1:])) * d2             # later bond legs and i
            nxt = ops.empty(na * chi_r * nc, d2)              # G[a, r, c, x]
            xs = d2 if has_x else 1                           # element stride of c in `cur`
            ops.gemm(chi_r, nc, chi_l, View(t, row=d2, col=chi_r * d2, b1=1),
                     View(cur, row=nc * xs, col=xs, b1=1 if has_x else 0,
                          b2=chi_l * nc * xs),
                     View(nxt, row=nc * d2, col=d2, b1=1, b2=chi_r * nc * d2),
                     nb1=d2, nb2=na)
            cur, has_x = nxt, True
            cdims[e] = chi_r
        assert cdims == list(bdims)
        r_ = int(np.prod(cdims))
        ops.gemm(d2, d2, r_, View(cur, row=d2, col=d2 * d2, b1=1),
                 View(b, row=d2, col=1),
                 View(out[step], row=d2, col=1, b1=d2 * d2), nb1=d2)

    bdims = list(dims)
    adjoint(num_steps - 1, back, bdims)
    for step in range(num_steps - 1, 0, -1):
        p1, p2 = propagators(step)
        back = sys_leg(back, np.asarray(p2, dtype=CDTYPE).T)
        back, bdims = through(back, bdims, step, True)
        back = sys_leg(back, np.asarray(p1, dtype=CDTYPE).T)
        adjoint(step - 1, back, bdims)
    d3 = ops.to_host(out)
    derivs = []
    for step in range(num_steps):
        full = np.zeros((d2, d2, d2, d2), dtype=CDTYPE)
        for x in range(d2):
            full[:, x, x, :] = d3[step, x]
        derivs.append(full)
    return derivs, states


def _fold_controls(propagators, controls, initial_state):
    """Controls (gradient.py:252-256, 287-301) folded into the half-step propagators:
    P1'_k = P1_k C^post_k,  P2'_k = C^pre_{k+1} P2_k,  rho_0' = C^pre_0 rho_0.  The state
    recorded at step k (after the pre-, before the post-measurement control) is then what
    the folded loop reads out."""
    rho0 = np.asarray(initial_state, dtype=CDTYPE)
    if controls is None or all(c[0] is None and c[1] is None for c in controls):
        return propagators, rho0, None
    d = rho0.shape[0]
    if controls[0][0] is not None:
        rho0 = (np.asarray(controls[0][0], dtype=CDTYPE) @ rho0.reshape(d * d)).reshape(d, d)

    def folded(step):
        p1, p2 = propagators(step)
        if controls[step][1] is not None:
            p1 = np.asarray(p1, dtype=CDTYPE) @ np.asarray(controls[step][1], dtype=CDTYPE)
        if controls[step + 1][0] is not None:
            p2 = np.asarray(controls[step + 1][0], dtype=CDTYPE) @ np.asarray(p2, dtype=CDTYPE)
        return p1, p2
    return folded, rho0, controls


def gradient_device(pt, propagators, initial_state, target_derivative, num_steps=None,
                    ops=None, controls=None):
    """compute_gradient_and_dynamics hot loops (oqupy/gradient.py:275-425): forward
    propagation through the process tensor keeping every intermediate state on the
    device, back-propagation of the target derivative through the bond-/leg-swapped
    PT-MPO (system_dynamics.py:588-628), and per step the adjoint tensor
    ``D[i, x, x', j] = sum_{l,r} F[l, i] T[l, r, x] d_xx' B[r, j]``
    (system_dynamics.py:702-786; leg order of gradient.py:216-224).

    ``pt``: one process tensor or a list of them (one bond leg per environment,
    gradient.py:266-270); rank-3 PT-MPO sites.  ``propagators(step)`` -> (P1, P2) as
    (d2, d2) superoperators; ``target_derivative`` is a (d, d) array or a callable of the
    final state; ``controls``: optional list of (pre, post) measurement controls per step
    0..num_steps ((d2, d2) superoperators or None, what ``Control.get_controls`` returns).
    Returns (propagator_derivatives, states): a list of ``num_steps`` ndarrays
    (d2, d2, d2, d2) -- what ``oqupy.gradient._chain_rule`` takes as ``adjoint_tensor`` --
    and the (num_steps+1, d, d) dynamics.
    """
    from ._lib import View  # pylint: disable=import-outside-toplevel
    ops = default_ops() if ops is None else ops
    if isinstance(pt, (list, tuple)):
        if len(pt) > 1:
            return _gradient_multi_env(list(pt), propagators, initial_state, target_derivative,
                                       num_steps, ops, controls)
        pt = pt[0]
    if num_steps is None:
        num_steps = len(pt)
    propagators, rho0, controls = _fold_controls(propagators, controls, initial_state)
    d = rho0.shape[0]
    d2 = d * d

    def sys_leg(x, mat):
        """x[..., j] <- sum_i mat[j, i] x[..., i] on a (1, chi, d2) device tensor."""
        dm = ops.from_host(np.ascontiguousarray(np.asarray(mat, dtype=CDTYPE)))
        rows_ = int(x.numel()) // d2
        out = ops.empty(*x.shape)
        ops.gemm(rows_, d2, d2, View(x, row=d2, col=1), View(dm, row=1, col=d2),
                 View(out, row=d2, col=1))
        return out
    # ---- forward (gradient.py:275-316): the fused dynamics step, states kept
    v = ops.from_host(rho0.reshape(1, 1, d2))
    rho = ops.empty(num_steps + 1, 1, d2)
    forward, props, props_t = [], [], []
    for step in range(num_steps):
        t = pt.get_mpo_tensor_device(step)
        chi_l, chi_r, _ = t.shape
        p1, p2 = (np.ascontiguousarray(np.asarray(p, dtype=CDTYPE))
                  for p in propagators(step))
        dp1, dp2 = ops.from_host(p1.reshape(1, d2, d2)), ops.from_host(p2.reshape(1, d2, d2))
        props.append((dp1, dp2))
        props_t.append((ops.from_host(p1.T.reshape(1, d2, d2)),
                        ops.from_host(p2.T.reshape(1, d2, d2))))
        forward.append(v)
        v_out = ops.empty(1, chi_r, d2)
        ops.dyn_step(1, chi_l, chi_r, d2, t, dp1, dp2, v, v_out,
                     cap=pt.get_cap_tensor_device(step), rho_out=rho[step])
        v = v_out
    cap = pt.get_cap_tensor_device(num_steps)
    chi = v.shape[1]
    ops.gemm(1, d2, chi, View(cap, col=1), View(v, row=d2, col=1),
             View(rho[num_steps], col=1))
    states = ops.to_host(rho).reshape(num_steps + 1, d, d)
    # ---- backward (gradient.py:336-425).  Both passes over T are the fused streaming
    # kernel of compute_dynamics (b200_dyn_step, T read once and coalesced):
    #  * adjoint tensor: W[i][r,x] = sum_l T[l,r,x] F[l,i] is a dynamics step of d2
    #    "members" that all carry F and select column i with P1[i][x,i'] = d_{i',i};
    #  * back-propagation: B <- ((B P2) o T^swap) P1 is a dynamics step through the
    #    bond-swapped site with the propagators P2^T, P1^T.
    target = target_derivative(states[-1]) if callable(target_derivative) \
        else target_derivative
    back = ops.from_host(np.asarray(target, dtype=CDTYPE).reshape(1, 1, d2))  # (1, chi_N=1, d2)
    out = ops.empty(num_steps, d2, d2, d2)        # D3[step][x][i][j]
    sel = np.zeros((d2, d2, d2), dtype=CDTYPE)
    for e in range(d2):
        sel[e, :, e] = 1.0
    p1_sel = ops.from_host(sel)
    p2_id = ops.from_host(np.array([np.identity(d2)] * d2, dtype=CDTYPE))

    def adjoint(step, b):
        """D3[x][i, j] = sum_{l,r} F[l, i] T[l, r, x] B[r, j] for the MPO of `step`."""
        t = pt.get_mpo_tensor_device(step)
        chi_l, chi_r, _ = t.shape
        f = forward[step]                                     # (1, chi_l, d2)
        if controls is not None:
            # the folded loop keeps F before the post-measurement control of its step and B
            # before the pre-measurement control of the next one (gradient.py:299-303, 406-412)
            if controls[step][1] is not None:
                f = sys_leg(f, controls[step][1])
            if controls[step + 1][0] is not None:
                b = sys_leg(b, np.asarray(controls[step + 1][0]).T)
        frep = ops.empty(d2, chi_l, d2)                       # d2 copies of F
        ops.gemm(1, chi_l * d2, 1, View(ops.one), View(f, col=1),
                 View(frep, col=1, b1=chi_l * d2), nb1=d2)
        w = ops.empty(d2, chi_r, d2)                          # W[i][r][x]
        ops.dyn_step(d2, chi_l, chi_r, d2, t, p1_sel, p2_id, frep, w)
        ops.gemm(d2, d2, chi_r, View(w, row=chi_r * d2, col=d2, b1=1),
                 View(b, row=d2, col=1),
                 View(out[step], row=d2, col=1, b1=d2 * d2), nb1=d2)

    adjoint(num_steps - 1, back)
    for step in range(num_steps - 1, 0, -1):
        ts = pt.get_mpo_tensor_device_swapped(step)           # (chi_r, chi_l, d2)
        chi_r, chi_l, _ = ts.shape
        dp1t, dp2t = props_t[step]
        nxt = ops.empty(1, chi_l, d2)
        ops.dyn_step(1, chi_r, chi_l, d2, ts, dp2t, dp1t, back, nxt)
        back = nxt
        adjoint(step - 1, back)
    d3 = ops.to_host(out)                                     # (N, x, i, j)
    derivs = []
    for step in range(num_steps):
        full = np.zeros((d2, d2, d2, d2), dtype=CDTYPE)
        for x in range(d2):
            full[:, x, x, :] = d3[step, x]
        derivs.append(full)
    return derivs, states
