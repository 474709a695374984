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
_FORCE = {"tebd": False}     # PtTebd has no module-level config dictionary to mutate


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


def install(default=False):
    """Rebind OQuPy's backend names.  With ``default=True`` the config dictionaries are
    also mutated in place so that every Tempo / PtTempo uses the B200 backend."""
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
    tm.TempoBackend = _ORIGINALS["TempoBackend"]
    tb.BaseTempoBackend = _ORIGINALS["BaseTempoBackend"]
    tm.MeanFieldTempoBackend = _ORIGINALS["MeanFieldTempoBackend"]
    ptm.PtTempoBackend = _ORIGINALS["PtTempoBackend"]
    oqupy.config.TEMPO_BACKEND_CONFIG.pop("backend", None)
    oqupy.config.PT_TEMPO_BACKEND_CONFIG.pop("backend", None)
