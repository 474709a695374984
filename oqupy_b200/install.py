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
    """``oqupy.compute_dynamics`` (system_dynamics.py:41-182) with the hot loop on the
    device when every process tensor is (or can be put) on the device, there are no
    controls and every step is recorded; anything else runs the reference code."""
    import numpy as np  # pylint: disable=import-outside-toplevel
    from .process_tensor import dynamics_device  # pylint: disable=import-outside-toplevel

    def compute_dynamics(system, initial_state=None, dt=None, num_steps=None, start_time=0.0,
                         process_tensor=None, control=None, record_all=True, **kwargs):
        if _FORCE["dynamics"] and process_tensor is not None and control is None \
                and record_all and initial_state is not None:
            pts = process_tensor if isinstance(process_tensor, (list, tuple)) \
                else [process_tensor]
            devs = [_device_process_tensor(p) for p in pts]
            if devs and all(d is not None for d in devs):
                step = dt if dt is not None else devs[0].dt
                n = num_steps if num_steps is not None else min(len(d) for d in devs)
                if step is not None and all(len(d) >= n for d in devs):
                    from oqupy.config import INTEGRATE_EPSREL, SUBDIV_LIMIT  # pylint: disable=import-outside-toplevel
                    props = system.get_propagators(
                        step, start_time, kwargs.get("subdiv_limit", SUBDIV_LIMIT),
                        kwargs.get("liouvillian_epsrel", INTEGRATE_EPSREL))
                    states = dynamics_device(devs if len(devs) > 1 else devs[0], props,
                                             np.asarray(initial_state), num_steps=n)
                    times = [start_time + step * k for k in range(n + 1)]
                    return dynamics_cls(times=times, states=list(states))
        return original(system, initial_state=initial_state, dt=dt, num_steps=num_steps,
                        start_time=start_time, process_tensor=process_tensor,
                        control=control, record_all=record_all, **kwargs)
    compute_dynamics.__doc__ = original.__doc__
    return compute_dynamics


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
    _FORCE["dynamics"] = bool(dynamics)
    if default:
        oqupy.config.TEMPO_BACKEND_CONFIG["backend"] = "b200"
        oqupy.config.PT_TEMPO_BACKEND_CONFIG["backend"] = "b200"


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
    tm.TempoBackend = _ORIGINALS["TempoBackend"]
    tb.BaseTempoBackend = _ORIGINALS["BaseTempoBackend"]
    tm.MeanFieldTempoBackend = _ORIGINALS["MeanFieldTempoBackend"]
    ptm.PtTempoBackend = _ORIGINALS["PtTempoBackend"]
    oqupy.config.TEMPO_BACKEND_CONFIG.pop("backend", None)
    oqupy.config.PT_TEMPO_BACKEND_CONFIG.pop("backend", None)
