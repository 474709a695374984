"""Drop-in hook for an existing OQuPy installation.

OQuPy's front-ends resolve the backend classes by module-global name at call time
(oqupy/tempo.py:424, 818, oqupy/pt_tempo.py:218, oqupy/backends/tempo_backend.py:715), so
rebinding those names is a complete drop-in: ``oqupy.Tempo(...).compute()``,
``oqupy.PtTempo(...)`` / ``oqupy.pt_tempo_compute`` then run on the B200.  Selection
follows the reference's config mechanism: the B200 classes are used when
``backend_config["backend"] == "b200"`` (oqupy/config.py:38,53 dictionaries, or the
``backend_config`` argument); any other value falls through to the original class.
The key is consumed here because the reference forwards ``config["backend"]`` to
tensornetwork (oqupy/backends/pt_tempo_backend.py:81-84).
"""
from . import backends as _b200
from . import tebd as _tebd

_ORIGINALS = {}
_MISSING = object()
# PtTebd and compute_dynamics have no module-level config dictionary to mutate
_FORCE = {"tebd": False, "dynamics": False}


def _dispatch(b200_cls, original_cls, force_key=None):
    def factory(*args, **kwargs):
        config = kwargs.get("config")
        if config is None:
            for a in args:
                if isinstance(a, dict):
                    config = a
        if isinstance(config, dict) and config.get("backend") == "b200":
            return b200_cls(*args, **kwargs)
        if force_key is not None and _FORCE[force_key] and not (
                isinstance(config, dict) and "backend" in config):
            return b200_cls(*args, **kwargs)
        return original_cls(*args, **kwargs)
    factory.__name__ = original_cls.__name__
    factory.__doc__ = original_cls.__doc__
    return factory


def _device_process_tensor(pt):
    from .process_tensor import as_device_process_tensor  # pylint: disable=import-outside-toplevel
    return as_device_process_tensor(pt)


def _compute_dynamics_factory(original, dynamics_cls):
    """``oqupy.compute_dynamics`` (system_dynamics.py:41-182) with the hot loop on the device
    whenever every process tensor is (or can be put) on the device.  The reference's own
    argument checks run first (``_compute_dynamics_input_parse``, system_dynamics.py:478-559:
    system type, state shape, Hilbert-space dimensions, equal time steps, num_steps), so bad
    input raises exactly what the reference raises.  Controls (system_dynamics.py:131-155) are
    folded into the per-step propagators:  P1'_k = P1_k C^post_k,  P2'_k = C^pre_{k+1} P2_k,
    rho_0' = C^pre_0 rho_0;  ``record_all=False`` returns the last state under the reference's
    time stamp (system_dynamics.py:176-180)."""
    import numpy as np  # pylint: disable=import-outside-toplevel
    from .process_tensor import dynamics_device  # pylint: disable=import-outside-toplevel

    def compute_dynamics(system, initial_state=None, dt=None, num_steps=None, start_time=0.0,
                         process_tensor=None, control=None, record_all=True, **kwargs):
        def reference():
            return original(system, initial_state=initial_state, dt=dt, num_steps=num_steps,
                            start_time=start_time, process_tensor=process_tensor,
                            control=control, record_all=record_all, **kwargs)
        if not _FORCE["dynamics"] or process_tensor is None:
            return reference()
        from oqupy.system_dynamics import _compute_dynamics_input_parse  # pylint: disable=import-outside-toplevel
        parsed = _compute_dynamics_input_parse(   # raises like the reference on bad input
            False, system, initial_state, dt, num_steps, start_time, process_tensor, control,
            record_all)
        sys_, rho0, step, n, t0, pts, ctrl, rec_all, hs_dim = parsed
        devs = [_device_process_tensor(p) for p in pts]
        if not devs or any(d is None for d in devs) or n < 1:
            return reference()
        from oqupy.config import INTEGRATE_EPSREL, SUBDIV_LIMIT  # pylint: disable=import-outside-toplevel
        props = sys_.get_propagators(step, t0, kwargs.get("subdiv_limit", SUBDIV_LIMIT),
                                     kwargs.get("liouvillian_epsrel", INTEGRATE_EPSREL))
        controls = [ctrl.get_controls(k, dt=step, start_time=t0) for k in range(n + 1)]
        rho0 = np.asarray(rho0, dtype=complex)
        if any(c[0] is not None or c[1] is not None for c in controls):
            d2 = hs_dim * hs_dim
            if controls[0][0] is not None:
                rho0 = (np.asarray(controls[0][0]) @ rho0.reshape(d2)).reshape(hs_dim, hs_dim)
            base = props

            def props(k):          # pylint: disable=function-redefined
                p1, p2 = base(k)
                if controls[k][1] is not None:
                    p1 = np.asarray(p1) @ np.asarray(controls[k][1])
                if controls[k + 1][0] is not None:
                    p2 = np.asarray(controls[k + 1][0]) @ np.asarray(p2)
                return p1, p2
        states = dynamics_device(devs if len(devs) > 1 else devs[0], props, rho0, num_steps=n)
        if not rec_all:
            return dynamics_cls(times=[t0 + 1 * step], states=[states[-1]])
        times = [t0 + step * k for k in range(n + 1)]
        return dynamics_cls(times=times, states=list(states))
    compute_dynamics.__doc__ = original.__doc__
    return compute_dynamics


def _compute_gradient_factory(original, dynamics_cls):
    """``oqupy.gradient.compute_gradient_and_dynamics`` (gradient.py:169-437) with the forward
    pass, the back-propagation and the adjoint tensors on the device
    (``oqupy_b200.gradient_device``: one or several environments, controls).  The reference's
    own argument checks run first; ``state_gradient`` (gradient.py:32-111) looks the function
    up in its module, so it takes the device path too and keeps its host ``_chain_rule``."""
    import numpy as np  # pylint: disable=import-outside-toplevel
    from .process_tensor import gradient_device  # pylint: disable=import-outside-toplevel

    def compute_gradient_and_dynamics(system, initial_state, target_derivative,
                                      process_tensors, parameters, start_time=0.0, dt=None,
                                      num_steps=None, control=None, record_all=True,
                                      progress_type=None):
        def reference():
            return original(system, initial_state, target_derivative, process_tensors,
                            parameters, start_time=start_time, dt=dt, num_steps=num_steps,
                            control=control, record_all=record_all,
                            progress_type=progress_type)
        if not _FORCE["dynamics"]:
            return reference()
        from oqupy.system import ParameterizedSystem  # pylint: disable=import-outside-toplevel
        from oqupy.system_dynamics import _compute_dynamics_input_parse  # pylint: disable=import-outside-toplevel
        from oqupy.util import check_isinstance  # pylint: disable=import-outside-toplevel
        parsed = _compute_dynamics_input_parse(   # raises like the reference on bad input
            False, system, initial_state, dt, num_steps, start_time, process_tensors, control,
            record_all)
        sys_, rho0, step, n, t0, pts, ctrl, rec_all, _ = parsed
        check_isinstance(sys_, ParameterizedSystem, "system")
        assert target_derivative is not None, "target state must be given explicitly"
        devs = [_device_process_tensor(p) for p in pts]
        if (not devs or any(d is None for d in devs) or n < 1
                or any(d.transform_in is not None or d.transform_out is not None for d in devs)):
            return reference()
        props = sys_.get_propagators(step, parameters)
        controls = [ctrl.get_controls(k, dt=step, start_time=t0) for k in range(n + 1)]
        derivs, states = gradient_device(devs if len(devs) > 1 else devs[0], props,
                                         np.asarray(rho0, dtype=complex), target_derivative,
                                         num_steps=n, controls=controls)
        if rec_all:
            times = [t0 + step * k for k in range(n + 1)]
            return derivs, dynamics_cls(times=times, states=list(states))
        return derivs, dynamics_cls(times=[t0 + 1 * step], states=[states[-1]])
    compute_gradient_and_dynamics.__doc__ = original.__doc__
    return compute_gradient_and_dynamics


def _compute_dynamics_with_field_factory(original, dynamics_cls):
    """``oqupy.compute_dynamics_with_field`` (system_dynamics.py:184-475): every system of the
    mean-field model runs through its process tensor(s) on the device
    (``oqupy_b200.dynamics_with_field_device``); the field equation of motion and the
    field-dependent propagators stay host callbacks, as in the reference.  The reference's
    own argument checks run first; anything the device path does not cover runs the
    reference code."""
    from .process_tensor import dynamics_with_field_device  # pylint: disable=import-outside-toplevel

    def compute_dynamics_with_field(mean_field_system, initial_field, process_tensor_list=None,
                                    dt=None, num_steps=None, initial_state_list=None,
                                    start_time=0.0, control_list=None, record_all=True,
                                    **kwargs):
        def reference():
            return original(mean_field_system, initial_field,
                            process_tensor_list=process_tensor_list, dt=dt,
                            num_steps=num_steps, initial_state_list=initial_state_list,
                            start_time=start_time, control_list=control_list,
                            record_all=record_all, **kwargs)
        from oqupy.system import MeanFieldSystem  # pylint: disable=import-outside-toplevel
        if (not _FORCE["dynamics"] or process_tensor_list is None
                or not isinstance(mean_field_system, MeanFieldSystem)
                or not isinstance(process_tensor_list, list)):
            return reference()
        systems = mean_field_system.system_list
        nsys = len(systems)
        states0 = [None] * nsys if initial_state_list is None else initial_state_list
        ctrls = [None] * nsys if control_list is None else control_list
        if nsys == 0 or not nsys == len(states0) == len(ctrls) == len(process_tensor_list):
            return reference()                 # raises the reference's own message
        from oqupy.config import INTEGRATE_EPSREL, SUBDIV_LIMIT  # pylint: disable=import-outside-toplevel
        from oqupy.system_dynamics import _compute_dynamics_input_parse  # pylint: disable=import-outside-toplevel
        parsed = [_compute_dynamics_input_parse(True, sys_, st, dt, num_steps, start_time, pt,
                                                ct, record_all)
                  for sys_, st, pt, ct in zip(systems, states0, process_tensor_list, ctrls)]
        step, n, rec_all = parsed[0][2], parsed[0][3], parsed[0][7]
        devs = [[_device_process_tensor(p) for p in par[5]] for par in parsed]
        if n < 1 or any(not d or any(x is None for x in d) for d in devs):
            return reference()
        props = [par[0].get_propagators(step, start_time,
                                        kwargs.get("subdiv_limit", SUBDIV_LIMIT),
                                        kwargs.get("liouvillian_epsrel", INTEGRATE_EPSREL))
                 for par in parsed]
        controls = [(lambda k, c=par[6]: c.get_controls(k, dt=step, start_time=start_time))
                    for par in parsed]
        states, fields = dynamics_with_field_device(
            devs, props, [par[1] for par in parsed], complex(initial_field),
            mean_field_system.field_eom, step, start_time, n, controls_list=controls)
        if rec_all:
            times = [start_time + step * k for k in range(n + 1)]
            return dynamics_cls(times=times, system_states_list=states, fields=fields)
        return dynamics_cls(times=[start_time + 1 * step], system_states_list=[states[-1]],
                            fields=[fields[-1]])
    compute_dynamics_with_field.__doc__ = original.__doc__
    return compute_dynamics_with_field


def install(default=False, dynamics=True):
    """Rebind OQuPy's backend names.  With ``default=True`` the config dictionaries are
    also mutated in place so that every Tempo / PtTempo uses the B200 backend.
    ``dynamics=True`` also routes ``oqupy.compute_dynamics`` through the device loop
    whenever its process tensors are (or can be uploaded to) device process tensors."""
    import oqupy  # pylint: disable=import-outside-toplevel
    import oqupy.backends.tempo_backend as tb  # pylint: disable=import-outside-toplevel
    import oqupy.pt_tebd as tebdm  # pylint: disable=import-outside-toplevel
    import oqupy.pt_tempo as ptm  # pylint: disable=import-outside-toplevel
    import oqupy.tempo as tm  # pylint: disable=import-outside-toplevel
    if not _ORIGINALS:
        _ORIGINALS.update(PtTebdBackend=tebdm.PtTebdBackend,
                          TempoBackend=tm.TempoBackend,
                          BaseTempoBackend=tb.BaseTempoBackend,
                          MeanFieldTempoBackend=tm.MeanFieldTempoBackend,
                          PtTempoBackend=ptm.PtTempoBackend)
    tm.TempoBackend = _dispatch(_b200.TempoBackend, _ORIGINALS["TempoBackend"])
    tb.BaseTempoBackend = _dispatch(_b200.BaseTempoBackend,
                                    _ORIGINALS["BaseTempoBackend"])
    tm.MeanFieldTempoBackend = _dispatch(_b200.MeanFieldTempoBackend,
                                         _ORIGINALS["MeanFieldTempoBackend"])
    ptm.PtTempoBackend = _dispatch(_b200.PtTempoBackend,
                                   _ORIGINALS["PtTempoBackend"])
    # oqupy/pt_tebd.py:26, 244: PtTebd(..., backend_config={"backend": "b200"})
    tebdm.PtTebdBackend = _dispatch(_tebd.PtTebdBackend, _ORIGINALS["PtTebdBackend"],
                                    force_key="tebd")
    _FORCE["tebd"] = bool(default)
    # oqupy.compute_dynamics (system_dynamics.py:41-182): module attribute of
    # oqupy.system_dynamics, re-exported by oqupy/__init__.py
    import oqupy.system_dynamics as sdm  # pylint: disable=import-outside-toplevel
    if "compute_dynamics" not in _ORIGINALS:
        _ORIGINALS["compute_dynamics"] = sdm.compute_dynamics
    shim = _compute_dynamics_factory(_ORIGINALS["compute_dynamics"], oqupy.Dynamics)
    sdm.compute_dynamics = shim
    oqupy.compute_dynamics = shim
    if "compute_dynamics_with_field" not in _ORIGINALS:
        _ORIGINALS["compute_dynamics_with_field"] = sdm.compute_dynamics_with_field
    fshim = _compute_dynamics_with_field_factory(_ORIGINALS["compute_dynamics_with_field"],
                                                 oqupy.MeanFieldDynamics)
    sdm.compute_dynamics_with_field = fshim
    oqupy.compute_dynamics_with_field = fshim
    # oqupy.gradient.compute_gradient_and_dynamics (gradient.py:169-437), also reached by
    # oqupy.state_gradient through its module
    import oqupy.gradient as gm  # pylint: disable=import-outside-toplevel
    if "compute_gradient_and_dynamics" not in _ORIGINALS:
        _ORIGINALS["compute_gradient_and_dynamics"] = gm.compute_gradient_and_dynamics
    gshim = _compute_gradient_factory(_ORIGINALS["compute_gradient_and_dynamics"],
                                      oqupy.Dynamics)
    gm.compute_gradient_and_dynamics = gshim
    oqupy.compute_gradient_and_dynamics = gshim
    _FORCE["dynamics"] = bool(dynamics)
    if default:
        for name, cfg in (("TEMPO", oqupy.config.TEMPO_BACKEND_CONFIG),
                          ("PT_TEMPO", oqupy.config.PT_TEMPO_BACKEND_CONFIG)):
            _ORIGINALS.setdefault("cfg_" + name, cfg.get("backend", _MISSING))
            cfg["backend"] = "b200"


def uninstall():
    """Restore the reference classes."""
    if not _ORIGINALS:
        return
    import oqupy  # pylint: disable=import-outside-toplevel
    import oqupy.backends.tempo_backend as tb  # pylint: disable=import-outside-toplevel
    import oqupy.pt_tebd as tebdm  # pylint: disable=import-outside-toplevel
    import oqupy.pt_tempo as ptm  # pylint: disable=import-outside-toplevel
    import oqupy.tempo as tm  # pylint: disable=import-outside-toplevel
    tebdm.PtTebdBackend = _ORIGINALS["PtTebdBackend"]
    _FORCE["tebd"] = False
    _FORCE["dynamics"] = False
    import oqupy.system_dynamics as sdm  # pylint: disable=import-outside-toplevel
    sdm.compute_dynamics = _ORIGINALS["compute_dynamics"]
    oqupy.compute_dynamics = _ORIGINALS["compute_dynamics"]
    sdm.compute_dynamics_with_field = _ORIGINALS["compute_dynamics_with_field"]
    oqupy.compute_dynamics_with_field = _ORIGINALS["compute_dynamics_with_field"]
    import oqupy.gradient as gm  # pylint: disable=import-outside-toplevel
    gm.compute_gradient_and_dynamics = _ORIGINALS["compute_gradient_and_dynamics"]
    oqupy.compute_gradient_and_dynamics = _ORIGINALS["compute_gradient_and_dynamics"]
    tm.TempoBackend = _ORIGINALS["TempoBackend"]
    tb.BaseTempoBackend = _ORIGINALS["BaseTempoBackend"]
    tm.MeanFieldTempoBackend = _ORIGINALS["MeanFieldTempoBackend"]
    ptm.PtTempoBackend = _ORIGINALS["PtTempoBackend"]
    for name, cfg in (("TEMPO", oqupy.config.TEMPO_BACKEND_CONFIG),
                      ("PT_TEMPO", oqupy.config.PT_TEMPO_BACKEND_CONFIG)):
        saved = _ORIGINALS.pop("cfg_" + name, None)
        if saved is None:
            continue                       # install(default=True) never touched it
        if saved is _MISSING:
            cfg.pop("backend", None)
        else:
            cfg["backend"] = saved         # the value the user had set before install()
