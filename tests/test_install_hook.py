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
    from oqupy_b200 import backends, install, process_tensor
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    # product default_ops raises without CUDA; inject the test model for this check
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    monkeypatch.setattr(process_tensor, "default_ops", lambda: ops)
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
        pt_new = ptn.get_process_tensor(progress_type="silent")
        # the device copies of the sites stay attached to the host process tensor
        assert pt_new._b200_device[0] == len(pt_new)
        d_new = oqupy.compute_dynamics(system, process_tensor=pt_new, initial_state=rho0,
                                       progress_type="silent")
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
        # the same with degeneracy-reduced legs (tests/physics/degeneracy_mean_field_test.py)
        ref_u = oqupy.MeanFieldTempo(mfs, [bath], params, [rho0], 1.0, start_time=0.0,
                                     unique=True)
        new_u = oqupy.MeanFieldTempo(mfs, [bath], params, [rho0], 1.0, start_time=0.0,
                                     unique=True, backend_config={"backend": "b200"})
        assert isinstance(new_u._backend_instance, backends.MeanFieldTempoBackend)
        ref_u.compute(0.4, progress_type="silent")
        new_u.compute(0.4, progress_type="silent")
        du, dnu = ref_u.get_dynamics(), new_u.get_dynamics()
        np.testing.assert_allclose(dnu.fields, du.fields, atol=1e-8)
        np.testing.assert_allclose(dnu.system_dynamics[0].states,
                                   du.system_dynamics[0].states, atol=1e-8)
    finally:
        install.uninstall()


def test_install_dispatch_pt_tebd_and_unique(monkeypatch):
    """oqupy.PtTebd (reference front-end: gate layers, TrivialProcessTensor, host
    SimpleProcessTensor) on the rebound PtTebdBackend; oqupy.PtTempo / Tempo with
    unique=True on the rebound TEMPO backends."""
    oqupy = load_reference()
    from oqupy_b200 import backends, install, process_tensor, tebd
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    monkeypatch.setattr(process_tensor, "default_ops", lambda: ops)
    monkeypatch.setattr(tebd, "default_ops", lambda: ops)
    sig = oqupy.operators.sigma
    corr = oqupy.PowerLawSD(alpha=0.3, zeta=3, cutoff=3.0, cutoff_type="exponential",
                            temperature=0.8)
    bath = oqupy.Bath(0.5 * sig("z"), corr)
    tparams = oqupy.TempoParameters(dt=0.1, dkmax=5, epsrel=1e-6)
    n = 4
    chain = oqupy.SystemChain(hilbert_space_dimensions=[2] * n)
    for s in range(n):
        chain.add_site_hamiltonian(site=s, hamiltonian=0.3 * (s + 1) * sig("x"))
    for s in range(n - 1):
        for i, xyz in enumerate("xyz"):
            chain.add_nn_hamiltonian(site=s, hamiltonian_l=0.5 * (1.0 + 0.2 * i) * sig(xyz),
                                     hamiltonian_r=0.5 * sig(xyz))
    params = oqupy.PtTebdParameters(dt=0.1, order=2, epsrel=1e-7)
    amps = oqupy.AugmentedMPS([oqupy.operators.spin_dm("z-")] * n)
    install.install()
    try:
        pt = oqupy.pt_tempo_compute(bath=bath, start_time=0.0, end_time=1.0,
                                    parameters=tparams, progress_type="silent")
        pts = [pt, None, pt, None]
        kw = dict(initial_augmented_mps=amps, system_chain=chain, process_tensors=pts,
                  parameters=params, dynamics_sites=[0, 1, 2, 3, (0, 2)])
        ref = oqupy.PtTebd(**kw)
        new = oqupy.PtTebd(backend_config={"backend": "b200"}, **kw)
        r_ref = ref.compute(6, progress_type="silent")
        r_new = new.compute(6, progress_type="silent")
        assert isinstance(new._t_mps, tebd.PtTebdBackend)
        assert not isinstance(ref._t_mps, tebd.PtTebdBackend)
        np.testing.assert_array_equal(r_new["bond_dimensions"], r_ref["bond_dimensions"])
        np.testing.assert_allclose(r_new["norm"], r_ref["norm"], atol=1e-9)
        for key in [0, 1, 2, 3, (0, 2)]:
            np.testing.assert_allclose(r_new["dynamics"][key].states,
                                       r_ref["dynamics"][key].states, atol=1e-9)
        a_ref, a_new = ref.get_augmented_mps(), new.get_augmented_mps()
        for lr, ln in zip(a_ref.lambdas, a_new.lambdas):
            np.testing.assert_allclose(np.abs(np.diag(ln) if ln.ndim == 2 else ln),
                                       np.abs(np.diag(lr) if lr.ndim == 2 else lr), atol=1e-9)
        # unique=True through the front-ends (degeneracy maps handed to the backends)
        system = oqupy.System(0.5 * sig("x"))
        rho0 = oqupy.operators.spin_dm("z+")
        t_ref = oqupy.Tempo(system, bath, tparams, rho0, 0.0, unique=True)
        t_new = oqupy.Tempo(system, bath, tparams, rho0, 0.0, unique=True,
                            backend_config={"backend": "b200"})
        assert isinstance(t_new._backend_instance, backends.TempoBackend)
        t_ref.compute(1.0, progress_type="silent")
        t_new.compute(1.0, progress_type="silent")
        np.testing.assert_allclose(t_new.get_dynamics().states, t_ref.get_dynamics().states,
                                   atol=1e-8)
        p_ref = oqupy.PtTempo(bath, 0.0, 1.0, tparams, unique=True)
        p_new = oqupy.PtTempo(bath, 0.0, 1.0, tparams, unique=True,
                              backend_config={"backend": "b200"})
        assert isinstance(p_new._backend_instance, backends.PtTempoBackend)
        d_ref = oqupy.compute_dynamics(system, process_tensor=p_ref.get_process_tensor(
            progress_type="silent"), initial_state=rho0, progress_type="silent")
        d_new = oqupy.compute_dynamics(system, process_tensor=p_new.get_process_tensor(
            progress_type="silent"), initial_state=rho0, progress_type="silent")
        np.testing.assert_allclose(d_new.states, d_ref.states, atol=1e-9)
    finally:
        install.uninstall()
    assert oqupy.pt_tebd.PtTebdBackend is install._ORIGINALS["PtTebdBackend"]


def test_install_compute_dynamics(monkeypatch):
    """oqupy.compute_dynamics rebound (system_dynamics.py:41-182): host process tensors from
    the reference's PT-TEMPO are uploaded once, one and two environments, controls,
    record_all=False, non-diagonal coupling transforms, the reference's own input checks."""
    oqupy = load_reference()
    from oqupy_b200 import backends, install, process_tensor
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    monkeypatch.setattr(process_tensor, "default_ops", lambda: ops)
    sig = oqupy.operators.sigma
    corr = oqupy.PowerLawSD(alpha=0.1, zeta=1, cutoff=4.0, cutoff_type="exponential",
                            temperature=0.5)
    bath = oqupy.Bath(0.5 * sig("z"), corr)
    system = oqupy.System(0.5 * sig("x") + 0.2 * sig("z"))
    params = oqupy.TempoParameters(dt=0.1, dkmax=5, epsrel=1e-6)
    rho0 = oqupy.operators.spin_dm("z+")
    pt = oqupy.pt_tempo_compute(bath=bath, start_time=0.0, end_time=1.0, parameters=params,
                                progress_type="silent")
    ref1 = oqupy.compute_dynamics(system, process_tensor=pt, initial_state=rho0,
                                  progress_type="silent")
    ref2 = oqupy.compute_dynamics(system, process_tensor=[pt, pt], initial_state=rho0,
                                  progress_type="silent")
    install.install()
    try:
        launches = ops.launches
        new1 = oqupy.compute_dynamics(system, process_tensor=pt, initial_state=rho0,
                                      progress_type="silent")
        assert ops.launches > launches            # ran on the (model) device
        assert hasattr(pt, "_b200_device")
        np.testing.assert_allclose(new1.times, ref1.times, atol=1e-12)
        np.testing.assert_allclose(new1.states, ref1.states, atol=1e-10)
        new2 = oqupy.compute_dynamics(system, process_tensor=[pt, pt], initial_state=rho0,
                                      num_steps=6, progress_type="silent")
        np.testing.assert_allclose(new2.states, ref2.states[:7], atol=1e-10)
        # controls (system_dynamics.py:131-155) are folded into the per-step propagators
        control = oqupy.Control(2)
        control.add_single(3, oqupy.operators.left_super(sig("x")))
        control.add_single(5, oqupy.operators.right_super(sig("y")), post=True)
        control.add_single(0, oqupy.operators.left_right_super(sig("x"), sig("x")))
        refc = install._ORIGINALS["compute_dynamics"](
            system, process_tensor=pt, initial_state=rho0, control=control,
            progress_type="silent")
        launches = ops.launches
        withc = oqupy.compute_dynamics(system, process_tensor=pt, initial_state=rho0,
                                       control=control, progress_type="silent")
        assert ops.launches > launches
        np.testing.assert_allclose(withc.states, refc.states, atol=1e-10)
        # record_all=False: the last state under the reference's time stamp (:176-180)
        ref_last = install._ORIGINALS["compute_dynamics"](
            system, process_tensor=pt, initial_state=rho0, record_all=False,
            progress_type="silent")
        new_last = oqupy.compute_dynamics(system, process_tensor=pt, initial_state=rho0,
                                          record_all=False, progress_type="silent")
        np.testing.assert_allclose(new_last.times, ref_last.times, atol=1e-12)
        np.testing.assert_allclose(new_last.states, ref_last.states, atol=1e-10)
        # bad input raises what the reference raises (system_dynamics.py:478-559)
        with pytest.raises(ValueError):
            oqupy.compute_dynamics(system, process_tensor=pt, initial_state=rho0, dt=0.07,
                                   progress_type="silent")
        with pytest.raises(ValueError):
            oqupy.compute_dynamics(system, process_tensor=pt, initial_state=np.eye(3),
                                   progress_type="silent")
        with pytest.raises(ValueError):
            oqupy.compute_dynamics(system, process_tensor=pt, initial_state=rho0,
                                   num_steps=99, progress_type="silent")
        # a later set_mpo_tensor invalidates the cached device copy
        dev_before = pt._b200_device[1]
        pt.set_mpo_tensor(2, pt._mpo_tensors[2] * 1.0)
        oqupy.compute_dynamics(system, process_tensor=pt, initial_state=rho0,
                               progress_type="silent")
        assert pt._b200_device[1] is not dev_before
        # non-diagonal coupling (pt_tempo.py:159-167, process_tensor.py:349-354): the
        # transforms are folded into the propagators, one and two environments
        bath_x = oqupy.Bath(0.5 * sig("x"), corr)
        pt_x = oqupy.pt_tempo_compute(bath=bath_x, start_time=0.0, end_time=1.0,
                                      parameters=params, progress_type="silent")
        assert pt_x._transform_in is not None
        for pts in (pt_x, [pt_x, pt]):
            refx = install._ORIGINALS["compute_dynamics"](
                system, process_tensor=pts, initial_state=rho0, progress_type="silent")
            launches = ops.launches
            newx = oqupy.compute_dynamics(system, process_tensor=pts, initial_state=rho0,
                                          progress_type="silent")
            assert ops.launches > launches
            np.testing.assert_allclose(newx.states, refx.states, atol=1e-10)
    finally:
        install.uninstall()
    assert oqupy.compute_dynamics is install._ORIGINALS["compute_dynamics"]


def test_install_compute_gradient_and_dynamics(monkeypatch):
    """oqupy.compute_gradient_and_dynamics rebound (gradient.py:169-437): one and two
    environments, controls, record_all=False, against the reference's own function on the
    same reference-built process tensors."""
    oqupy = load_reference()
    np.vectorize = lambda f, *a, **k: f     # numpy-2: see tests/golden/make_golden_mean_field.py
    from oqupy_b200 import backends, install, process_tensor
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    monkeypatch.setattr(process_tensor, "default_ops", lambda: ops)
    sig = oqupy.operators.sigma
    dt, n = 0.1, 8
    pts = []
    for alpha, temp in ((0.2, 0.5), (0.1, 1.3)):
        corr = oqupy.PowerLawSD(alpha=alpha, zeta=1, cutoff=4.0, cutoff_type="exponential",
                                temperature=temp)
        pts.append(oqupy.pt_tempo_compute(
            bath=oqupy.Bath(0.5 * sig("z"), corr), start_time=0.0, end_time=dt * n,
            parameters=oqupy.TempoParameters(dt=dt, dkmax=4, epsrel=1e-6),
            progress_type="silent"))
    system = oqupy.ParameterizedSystem(hamiltonian=lambda hx: 0.5 * hx * sig("x") + 0.1 * sig("z"))
    x0 = np.linspace(0.5, 1.5, 2 * n).reshape(-1, 1)
    rho0 = oqupy.operators.spin_dm("z+")
    target = oqupy.operators.spin_dm("x-").T
    control = oqupy.Control(2)
    control.add_single(2, oqupy.operators.left_super(sig("x")))
    control.add_single(5, oqupy.operators.right_super(sig("y")), post=True)
    control.add_single(n, oqupy.operators.left_right_super(sig("x"), sig("x")))
    cases = [dict(process_tensors=[pts[0]]), dict(process_tensors=pts),
             dict(process_tensors=[pts[0]], control=control),
             dict(process_tensors=pts, control=control),
             dict(process_tensors=[pts[1]], record_all=False)]
    common = dict(system=system, initial_state=rho0, target_derivative=target, parameters=x0,
                  dt=dt, progress_type="silent")
    refs = [oqupy.compute_gradient_and_dynamics(**common, **c) for c in cases]
    install.install()
    try:
        assert oqupy.gradient.compute_gradient_and_dynamics is not \
            install._ORIGINALS["compute_gradient_and_dynamics"]
        for case, (gref, dref) in zip(cases, refs):
            launches = ops.launches
            gnew, dnew = oqupy.compute_gradient_and_dynamics(**common, **case)
            assert ops.launches > launches            # ran on the (model) device
            np.testing.assert_allclose(dnew.times, dref.times, atol=1e-12)
            np.testing.assert_allclose(dnew.states, dref.states, atol=1e-10)
            ref_arr = np.array([np.asarray(getattr(t, "tensor", t)) for t in gref])
            np.testing.assert_allclose(np.array(gnew), ref_arr, atol=1e-10)
        with pytest.raises(ValueError):
            oqupy.compute_gradient_and_dynamics(**{**common, "dt": 0.07}, process_tensors=pts)
    finally:
        install.uninstall()
    assert oqupy.gradient.compute_gradient_and_dynamics is \
        install._ORIGINALS["compute_gradient_and_dynamics"]


def test_install_compute_dynamics_with_field(monkeypatch):
    """oqupy.compute_dynamics_with_field rebound (system_dynamics.py:184-475): a mean-field
    model of two systems, each with its own reference-built process tensor (the second one
    with two environments), a control on one system, against the reference's own function."""
    oqupy = load_reference()
    np.vectorize = lambda f, *a, **k: f     # numpy-2: see tests/golden/make_golden_mean_field.py
    from oqupy_b200 import backends, install, process_tensor
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    monkeypatch.setattr(process_tensor, "default_ops", lambda: ops)
    sig = oqupy.operators.sigma
    dt, n = 0.1, 8
    pts = []
    for alpha, temp in ((0.2, 0.5), (0.1, 1.3)):
        corr = oqupy.PowerLawSD(alpha=alpha, zeta=1, cutoff=4.0, cutoff_type="exponential",
                                temperature=temp)
        pts.append(oqupy.pt_tempo_compute(
            bath=oqupy.Bath(0.5 * sig("z"), corr), start_time=0.0, end_time=dt * n,
            parameters=oqupy.TempoParameters(dt=dt, dkmax=4, epsrel=1e-6),
            progress_type="silent"))

    def make_system(w):
        return oqupy.TimeDependentSystemWithField(
            lambda t, field: 0.5 * w * sig("z") + 0.3 * (field * sig("+") + np.conj(field) * sig("-")))

    def field_eom(t, states, field):
        return -(0.2 + 1.0j) * field - 0.3j * sum(np.trace(sig("-") @ s) for s in states)

    mfs = oqupy.MeanFieldSystem([make_system(1.0), make_system(0.6)], field_eom=field_eom)
    rho0 = [oqupy.operators.spin_dm("z-"), oqupy.operators.spin_dm("x+")]
    control = oqupy.Control(2)
    control.add_single(3, oqupy.operators.left_super(sig("x")))
    control.add_single(5, oqupy.operators.right_super(sig("y")), post=True)
    cases = [dict(process_tensor_list=[pts[0], pts[1]]),
             dict(process_tensor_list=[pts[0], [pts[1], pts[0]]], control_list=[control, None]),
             dict(process_tensor_list=[pts[0], pts[1]], record_all=False)]
    common = dict(mean_field_system=mfs, initial_field=0.4 + 0.1j, initial_state_list=rho0,
                  progress_type="silent")
    refs = [oqupy.compute_dynamics_with_field(**common, **c) for c in cases]
    install.install()
    try:
        for case, ref in zip(cases, refs):
            launches = ops.launches
            new = oqupy.compute_dynamics_with_field(**common, **case)
            assert ops.launches > launches            # ran on the (model) device
            np.testing.assert_allclose(new.times, ref.times, atol=1e-12)
            np.testing.assert_allclose(new.fields, ref.fields, atol=1e-10)
            for a, b in zip(new.system_dynamics, ref.system_dynamics):
                np.testing.assert_allclose(a.states, b.states, atol=1e-10)
        with pytest.raises(AssertionError):
            oqupy.compute_dynamics_with_field(**common, process_tensor_list=[pts[0]])
    finally:
        install.uninstall()
    assert oqupy.compute_dynamics_with_field is install._ORIGINALS["compute_dynamics_with_field"]


def test_install_compute_correlations(monkeypatch):
    """oqupy.compute_correlations / compute_correlations_nt (system_dynamics.py:791-1095) call
    compute_dynamics with controls once per operator ordering: with the hook installed they
    run through the device loop unchanged and return what the reference returns."""
    oqupy = load_reference()
    from oqupy_b200 import backends, install, process_tensor
    from host_model_ops import HostModelOps
    ops = HostModelOps()
    monkeypatch.setattr(backends, "default_ops", lambda: ops)
    monkeypatch.setattr(process_tensor, "default_ops", lambda: ops)
    sig = oqupy.operators.sigma
    corr = oqupy.PowerLawSD(alpha=0.1, zeta=1, cutoff=4.0, cutoff_type="exponential",
                            temperature=0.5)
    pt = oqupy.pt_tempo_compute(bath=oqupy.Bath(0.5 * sig("z"), corr), start_time=0.0,
                                end_time=0.8, progress_type="silent",
                                parameters=oqupy.TempoParameters(dt=0.1, dkmax=5, epsrel=1e-6))
    system = oqupy.System(0.5 * sig("x") + 0.2 * sig("z"))
    args = dict(system=system, process_tensor=pt, operator_a=sig("x"), operator_b=sig("z"),
                times_a=0.2, times_b=(0.2, 0.7), time_order="ordered",
                initial_state=oqupy.operators.spin_dm("z+"), progress_type="silent")
    t_ref, c_ref = oqupy.compute_correlations(**args)
    install.install()
    try:
        launches = ops.launches
        t_new, c_new = oqupy.compute_correlations(**args)
        assert ops.launches > launches            # ran on the (model) device
        for a, b in zip(t_new, t_ref):
            np.testing.assert_allclose(a, b, atol=1e-12)
        np.testing.assert_allclose(c_new, c_ref, atol=1e-10, equal_nan=True)
    finally:
        install.uninstall()
