"""CPU (build container only): the install() hook rebinds the names OQuPy's front-ends
look up and dispatches on backend_config['backend']; with any other value
the reference classes run unchanged.  Skipped where /root/reference is absent."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ref_loader import load_reference, reference_available  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_available(),
                                reason="reference tree not present")


def test_install_dispatch(monkeypatch):
    oqupy = load_reference()
    from oqupy_b200 import backends, install
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    # product default_ops raises without CUDA; inject the test model for this check
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    install.install()
    try:
        corr = oqupy.PowerLawSD(alpha=0.1, zeta=1, cutoff=4.0,
                                cutoff_type="exponential", temperature=0.5)
        bath = oqupy.Bath(0.5 * oqupy.operators.sigma("z"), corr)
        system = oqupy.System(0.5 * oqupy.operators.sigma("x"))
        params = oqupy.TempoParameters(dt=0.1, dkmax=5, epsrel=1e-6)
        rho0 = oqupy.operators.spin_dm("z+")
        ref = oqupy.Tempo(system, bath, params, rho0, 0.0)
        ref.compute(1.0, progress_type="silent")
        new = oqupy.Tempo(system, bath, params, rho0, 0.0,
                          backend_config={"backend": "b200"})
        assert isinstance(new._backend_instance, backends.TempoBackend)
        assert not isinstance(ref._backend_instance, backends.TempoBackend)
        new.compute(1.0, progress_type="silent")
        np.testing.assert_allclose(new.get_dynamics().states,
                                   ref.get_dynamics().states, atol=1e-9)
        # PT-TEMPO through the reference front-end, host SimpleProcessTensor
        ptr = oqupy.PtTempo(bath, 0.0, 1.0, params)
        ptn = oqupy.PtTempo(bath, 0.0, 1.0, params, backend_config={"backend": "b200"})
        assert isinstance(ptn._backend_instance, backends.PtTempoBackend)
        d_ref = oqupy.compute_dynamics(system, process_tensor=ptr.get_process_tensor(
            progress_type="silent"), initial_state=rho0, progress_type="silent")
        d_new = oqupy.compute_dynamics(system, process_tensor=ptn.get_process_tensor(
            progress_type="silent"), initial_state=rho0, progress_type="silent")
        np.testing.assert_allclose(d_new.states, d_ref.states, atol=1e-9)
    finally:
        install.uninstall()
    assert oqupy.tempo.TempoBackend is install._ORIGINALS["TempoBackend"]


def test_install_dispatch_mean_field(monkeypatch):
    """oqupy.MeanFieldTempo (reference front-end, test G of
    tests/physics/mean_field_tempo_test.py) on the rebound MeanFieldTempoBackend."""
    oqupy = load_reference()
    from oqupy_b200 import backends, install
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    # numpy 2 rejects np.vectorize(h)(t) for array-valued h; the reference (numpy<2)
    # only uses it as an input check
    monkeypatch.setattr(np, "vectorize", lambda f, *a, **k: f)
    sx, sz = oqupy.operators.sigma("x"), oqupy.operators.sigma("z")
    system = oqupy.TimeDependentSystemWithField(
        lambda t, field: 0.5 * sz + np.real(field) * sx)
    mfs = oqupy.MeanFieldSystem(
        [system], lambda t, states, field: -(1j + 1) * field
        - 0.5j * np.matmul(sx, states[0]).trace().real)
    corr = oqupy.PowerLawSD(alpha=0.1, zeta=1.0, cutoff=5.0, cutoff_type="gaussian",
                            temperature=0.0)
    bath = oqupy.Bath(0.5 * sz, corr)
    params = oqupy.TempoParameters(dt=0.05, tcut=None, epsrel=1e-7)
    rho0 = np.array([[0.5, 0.5], [0.5, 0.5]])
    install.install()
    try:
        ref = oqupy.MeanFieldTempo(mfs, [bath], params, [rho0], 1.0, start_time=0.0)
        new = oqupy.MeanFieldTempo(mfs, [bath], params, [rho0], 1.0, start_time=0.0,
                                   backend_config={"backend": "b200"})
        assert isinstance(new._backend_instance, backends.MeanFieldTempoBackend)
        assert not isinstance(ref._backend_instance, backends.MeanFieldTempoBackend)
        ref.compute(0.5, progress_type="silent")
        new.compute(0.5, progress_type="silent")
        dr, dn = ref.get_dynamics(), new.get_dynamics()
        np.testing.assert_allclose(dn.fields, dr.fields, atol=1e-8)
        np.testing.assert_allclose(dn.system_dynamics[0].states,
                                   dr.system_dynamics[0].states, atol=1e-8)
    finally:
        install.uninstall()
