"""ctypes binding of the C-ABI library ``liboqupy_b200.so`` (include/oqupy_b200.h).

torch is used for device memory and streams only.  There is NO CPU fallback: if the
shared library or a CUDA device is missing, constructing :class:`CudaOps` raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_int, c_int32,
                    c_int64, c_size_t, c_uint64, c_void_p)

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboqupy_b200.so")

EXPORTS = [
    "b200_last_error", "b200_abi_version", "b200_launch_count",
    "b200_profile_enable", "b200_profile_read",
    "b200_zgemm_strided", "b200_svd_workspace_bytes", "b200_svd_factor",
    "b200_svd_emit", "b200_svd_factor2", "b200_svd_emit_parts", "b200_svd_values",
    "b200_svd_phase_cycles", "b200_svd_qr_phase_cycles", "b200_svd_plan", "b200_svd_qr_layout", "b200_svd_config", "b200_profile_read_kinds", "b200_dyn_workspace_bytes",
    "b200_dyn_step", "b200_caps_step", "b200_dyn_run", "b200_dyn_run_workspace_bytes",
    "b200_chain_create", "b200_chain_destroy", "b200_chain_len", "b200_chain_push",
    "b200_chain_shape", "b200_chain_read", "b200_chain_svd_sweep",
    "b200_chain_pt_zip_up_left", "b200_chain_tempo_step", "b200_chain_stats",
    "b200_chain_log",
    "b200_tempo_batch_create", "b200_tempo_batch_destroy", "b200_tempo_batch_set",
    "b200_tempo_batch_step", "b200_tempo_batch_info", "b200_tempo_batch_bytes",
    "b200_tempo_batch_set_order", "b200_tempo_batch_reserve_sms",
]


class B200Error(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status."""


class _Operand(Structure):
    _fields_ = [("ptr", c_void_p), ("row", c_int64), ("col", c_int64),
                ("b1", c_int64), ("b2", c_int64), ("conj", c_int)]


class _PtSite(Structure):
    _fields_ = [("kind", c_int), ("rows", c_int), ("cols", c_int), ("mat", c_void_p),
                ("north_map", POINTER(c_int32)), ("west_map", POINTER(c_int32))]


class _TempoSite(Structure):
    _fields_ = [("kind", c_int), ("rows", c_int), ("cols", c_int), ("nw", c_int),
                ("ns", c_int), ("mat", c_void_p)]


PT_KINDS = {"first": 0, "mid": 1, "last": 2, "closed": 3}
TEMPO_KINDS = {"start": 0, "mid": 1, "dense": 2}

_lib = None


def load_library():
    """Load (once) and return the ctypes handle; raises if the .so is missing."""
    global _lib  # pylint: disable=global-statement
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} not found: build it with oqupy_b200/csrc/build.sh "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.b200_last_error.restype = c_char_p
    lib.b200_abi_version.restype = c_int
    lib.b200_launch_count.restype = c_uint64
    lib.b200_profile_enable.restype = c_int
    lib.b200_profile_enable.argtypes = [c_int]
    lib.b200_profile_read.restype = c_int
    lib.b200_profile_read.argtypes = [POINTER(c_double), POINTER(c_double),
                                      POINTER(c_uint64), POINTER(c_uint64)]
    lib.b200_zgemm_strided.restype = c_int
    lib.b200_zgemm_strided.argtypes = [
        c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(_Operand),
        POINTER(_Operand), c_void_p, c_int64, c_int64, c_int64, c_int64,
        c_void_p, c_int64, c_int64, c_int]
    lib.b200_svd_workspace_bytes.restype = c_size_t
    lib.b200_svd_workspace_bytes.argtypes = [c_int, c_int]
    lib.b200_svd_factor.restype = c_int
    lib.b200_svd_factor.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int64,
                                    c_int64, c_double, c_void_p, c_void_p]
    lib.b200_svd_emit.restype = c_int
    lib.b200_svd_emit.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int64, c_int64, c_int, c_void_p, c_int,
                                  c_int64, c_int64, c_int64, c_void_p]
    lib.b200_svd_factor2.restype = c_int
    lib.b200_svd_factor2.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int64,
                                     c_int64, c_int, c_int64, c_int64, c_double, c_double,
                                     c_void_p, c_void_p]
    lib.b200_svd_emit_parts.restype = c_int
    lib.b200_svd_emit_parts.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                        c_int, c_int64, c_int64, c_int64, c_void_p, c_int,
                                        c_void_p, c_void_p]
    lib.b200_svd_phase_cycles.restype = c_int
    lib.b200_svd_phase_cycles.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.b200_svd_qr_phase_cycles.restype = c_int
    lib.b200_svd_qr_phase_cycles.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.b200_svd_values.restype = c_int
    lib.b200_svd_values.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.b200_svd_plan.restype = c_int
    lib.b200_svd_plan.argtypes = [c_void_p, c_int, c_int, POINTER(c_int32)]
    lib.b200_svd_qr_layout.restype = c_int
    lib.b200_svd_qr_layout.argtypes = [c_int, c_int, c_int, POINTER(c_int64)]
    lib.b200_svd_config.restype = c_int
    lib.b200_svd_config.argtypes = [c_char_p, c_double]
    lib.b200_profile_read_kinds.restype = c_int
    lib.b200_profile_read_kinds.argtypes = [POINTER(c_double), POINTER(c_uint64),
                                            POINTER(c_double), POINTER(c_uint64)]
    lib.b200_dyn_workspace_bytes.restype = c_size_t
    lib.b200_dyn_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.b200_dyn_step.restype = c_int
    lib.b200_dyn_step.argtypes = [c_void_p, c_int, c_int, c_int, c_int] + \
        [c_void_p] * 8
    lib.b200_caps_step.restype = c_int
    lib.b200_caps_step.argtypes = [c_void_p, c_int, c_int, c_int] + \
        [c_void_p] * 4
    lib.b200_dyn_run_workspace_bytes.restype = c_size_t
    lib.b200_dyn_run_workspace_bytes.argtypes = [c_int, c_int, POINTER(c_int32), c_int]
    lib.b200_dyn_run.restype = c_int
    lib.b200_dyn_run.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(c_int32),
                                 POINTER(c_void_p), c_void_p, c_void_p, c_int64,
                                 POINTER(c_void_p), c_void_p, c_void_p, c_void_p]
    lib.b200_chain_create.restype = c_void_p
    lib.b200_chain_create.argtypes = [c_void_p]
    lib.b200_chain_destroy.restype = c_int
    lib.b200_chain_destroy.argtypes = [c_void_p]
    lib.b200_chain_len.restype = c_int
    lib.b200_chain_len.argtypes = [c_void_p]
    lib.b200_chain_push.restype = c_int
    lib.b200_chain_push.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int]
    lib.b200_chain_shape.restype = c_int
    lib.b200_chain_shape.argtypes = [c_void_p, c_int, POINTER(c_int32)]
    lib.b200_chain_read.restype = c_int
    lib.b200_chain_read.argtypes = [c_void_p, c_int, c_void_p]
    lib.b200_chain_svd_sweep.restype = c_int
    lib.b200_chain_svd_sweep.argtypes = [c_void_p, c_int, c_int, c_double]
    lib.b200_chain_pt_zip_up_left.restype = c_int
    lib.b200_chain_pt_zip_up_left.argtypes = [c_void_p, POINTER(_PtSite), c_int, c_double]
    lib.b200_chain_tempo_step.restype = c_int
    lib.b200_chain_tempo_step.argtypes = [c_void_p, POINTER(_TempoSite), c_int, c_void_p,
                                          c_void_p, c_void_p, c_int, c_double, c_void_p]
    lib.b200_chain_stats.restype = c_int
    lib.b200_chain_stats.argtypes = [c_void_p, POINTER(c_uint64), POINTER(c_uint64),
                                     POINTER(c_uint64), c_int]
    lib.b200_chain_log.restype = c_int
    lib.b200_chain_log.argtypes = [c_void_p, c_int, POINTER(c_int32), c_int]
    lib.b200_tempo_batch_create.restype = c_void_p
    lib.b200_tempo_batch_create.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_double]
    lib.b200_tempo_batch_destroy.restype = c_int
    lib.b200_tempo_batch_destroy.argtypes = [c_void_p]
    lib.b200_tempo_batch_set.restype = c_int
    lib.b200_tempo_batch_set.argtypes = [c_void_p] * 7
    lib.b200_tempo_batch_step.restype = c_int
    lib.b200_tempo_batch_step.argtypes = [c_void_p] * 4
    lib.b200_tempo_batch_info.restype = c_int
    lib.b200_tempo_batch_info.argtypes = [c_void_p] + [POINTER(c_int32)] * 5
    lib.b200_tempo_batch_set_order.restype = c_int
    lib.b200_tempo_batch_set_order.argtypes = [c_void_p, POINTER(c_int32)]
    lib.b200_tempo_batch_reserve_sms.restype = c_int
    lib.b200_tempo_batch_reserve_sms.argtypes = [c_void_p, c_int]
    lib.b200_tempo_batch_bytes.restype = c_size_t
    lib.b200_tempo_batch_bytes.argtypes = [c_void_p]
    _lib = lib
    return lib


class View:
    """Strided 2-D (+ two batch strides) view of a complex128 tensor, in elements."""
    __slots__ = ("t", "row", "col", "b1", "b2", "off", "conj")

    def __init__(self, t, row=0, col=0, b1=0, b2=0, off=0, conj=False):
        self.t, self.row, self.col = t, int(row), int(col)
        self.b1, self.b2, self.off, self.conj = int(b1), int(b2), int(off), conj


class SvdHandle:
    """Result of svd_factor: keeps the device workspace alive until emit."""
    __slots__ = ("work", "m", "n", "keep", "sweeps", "status", "rotations",
                 "theta", "theta_ptr", "rs", "cs")


class CudaOps:
    """The product ops object: every method is a C-ABI call on ``cuda:device``."""

    name = "cuda"

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise B200Error("oqupy_b200 needs a CUDA device (no CPU fallback).")
        self.lib = load_library()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self._info = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._info_np = self._info.numpy()
        self._work = None
        self.one = torch.ones(1, dtype=torch.complex128, device=self.device)
        self.svd_log = None       # optional list: (m, n, keep, sweeps)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # -- memory ---------------------------------------------------------------
    def empty(self, *shape):
        return torch.empty(shape, dtype=torch.complex128, device=self.device)

    def from_host(self, array):
        a = np.ascontiguousarray(np.asarray(array, dtype=np.complex128))
        self.h2d_bytes += a.nbytes
        return torch.from_numpy(a).to(self.device)

    def to_host(self, tensor):
        self.d2h_bytes += tensor.numel() * tensor.element_size()
        return tensor.cpu().numpy()

    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, code, what):
        if code != 0:
            raise B200Error(f"{what} failed ({code}): "
                            f"{self.lib.b200_last_error().decode()}")

    @staticmethod
    def _ptr(t, off=0):
        return t.data_ptr() + 16 * off

    # -- kernels --------------------------------------------------------------
    def gemm(self, m, n, k, a, b, c, nb1=1, nb2=1, scale=None, accumulate=False):
        oa = _Operand(self._ptr(a.t, a.off), a.row, a.col, a.b1, a.b2,
                      1 if a.conj else 0)
        ob = _Operand(self._ptr(b.t, b.off), b.row, b.col, b.b1, b.b2,
                      1 if b.conj else 0)
        if scale is None:
            sp, s1, s2 = None, 0, 0
        else:
            sp, s1, s2 = self._ptr(scale.t, scale.off), scale.b1, scale.b2
        code = self.lib.b200_zgemm_strided(
            self._stream(), m, n, k, nb1, nb2, ctypes.byref(oa),
            ctypes.byref(ob), self._ptr(c.t, c.off), c.row, c.col, c.b1, c.b2,
            sp, s1, s2, 1 if accumulate else 0)
        self._check(code, "b200_zgemm_strided")

    def svd_factor(self, theta, m, n, rs, cs, eps, off=0, rin=1, rsi=0, cin=1, csi=0,
                   cos_tol=0.0):
        """theta[i][j] at (i // rin)*rs + (i % rin)*rsi + (j // cin)*cs + (j % cin)*csi;
        ``cos_tol`` > 0: orthogonality target of the iteration (default 1e-11)."""
        nbytes = self.lib.b200_svd_workspace_bytes(m, n)
        work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        code = self.lib.b200_svd_factor2(
            self._stream(), self._ptr(theta, off), m, n, rin, rs, rsi, cin, cs, csi,
            -1.0 if eps is None else float(eps), float(cos_tol), work.data_ptr(),
            self._info.data_ptr())
        self._check(code, "b200_svd_factor")
        torch.cuda.current_stream(self.device).synchronize()
        h = SvdHandle()
        h.work, h.m, h.n = work, m, n
        h.theta, h.theta_ptr, h.rs, h.cs = theta, self._ptr(theta, off), rs, cs
        h.keep, h.sweeps, h.status, h.rotations = (int(x) for x in self._info_np[:4])
        self.d2h_bytes += 16
        if h.status != 0:
            raise B200Error(f"Jacobi SVD did not converge ({m}x{n}, "
                            f"{h.sweeps} sweeps)")
        if self.svd_log is not None:
            self.svd_log.append((m, n, h.keep, h.sweeps))
        return h

    def svd_emit(self, h, u=None, u_na=1, u_so=0, u_sa=0, u_sj=0, svh=None, vh=None,
                 lam=None, inv_lam=None):
        """U, and either S*Vh (``svh``) or Vh (``vh``); ``lam`` / ``inv_lam``: the kept
        singular values and their inverses as complex128 vectors."""
        assert svh is None or vh is None
        right = svh if vh is None else vh
        code = self.lib.b200_svd_emit_parts(
            self._stream(), h.work.data_ptr(), h.m, h.n, h.keep,
            None if u is None else u.data_ptr(), u_na, u_so, u_sa, u_sj,
            None if right is None else right.data_ptr(), 0 if vh is None else 1,
            None if lam is None else lam.data_ptr(),
            None if inv_lam is None else inv_lam.data_ptr())
        self._check(code, "b200_svd_emit_parts")

    def svd_phase_cycles(self, h):
        out = (ctypes.c_longlong * 16)()
        self._check(self.lib.b200_svd_phase_cycles(self._stream(), h.work.data_ptr(),
                                                   out), "b200_svd_phase_cycles")
        return list(out)

    def svd_plan(self, h):
        """(qr path used, pivots above the stop level, QRCP CTAs, resident columns)."""
        out = (c_int32 * 4)()
        self._check(self.lib.b200_svd_plan(h.work.data_ptr(), h.m, h.n, out), "b200_svd_plan")
        return tuple(int(x) for x in out)

    def svd_config(self, key, value):
        """Runtime switch of the truncated SVD ("qr", "qr_minq", "qr_cols")."""
        self._check(self.lib.b200_svd_config(key.encode(), float(value)), "b200_svd_config")

    def svd_qr_debug(self, h):
        """Intermediate arrays of the QR path (test access): dict with a (p x q work array),
        perm, tau, k, tail2, or None when the plain path ran."""
        qr, k, _, _ = self.svd_plan(h)
        if not qr:
            return None
        out = (c_int64 * 16)()
        self._check(self.lib.b200_svd_qr_layout(h.m, h.n, k, out), "b200_svd_qr_layout")
        o = [int(x) for x in out]
        p, q, grid = o[1], o[2], o[4]
        raw = h.work.cpu().numpy()

        def arr(off, count, dtype):
            return np.frombuffer(raw, dtype=dtype, count=count, offset=off).copy()
        return {"p": p, "q": q, "k": k, "transposed": bool(o[3]), "grid": grid,
                "resident": o[5],
                "perm": arr(o[6], q, np.int32), "tau": arr(o[7], k, np.complex128),
                "a": arr(o[8], p * q, np.complex128).reshape(q, p).T,
                "tail2": float(arr(o[11], grid, np.float64).sum()),
                "sval": arr(o[14], k, np.float64)}

    def svd_qr_phase_cycles(self, h):
        out = (ctypes.c_longlong * 8)()
        self._check(self.lib.b200_svd_qr_phase_cycles(self._stream(), h.work.data_ptr(), out),
                    "b200_svd_qr_phase_cycles")
        return list(out)

    def svd_values(self, h):
        k = min(h.m, h.n)
        out = torch.empty(k, dtype=torch.float64, device=self.device)
        code = self.lib.b200_svd_values(self._stream(), h.work.data_ptr(), h.m,
                                        h.n, out.data_ptr())
        self._check(code, "b200_svd_values")
        return out.cpu().numpy()

    def dyn_step(self, nvec, chi_l, chi_r, d2, t, p1, p2, v, v_out, cap=None,
                 rho_out=None):
        nbytes = self.lib.b200_dyn_workspace_bytes(nvec, chi_l, chi_r, d2)
        if self._work is None or self._work.numel() < nbytes:
            self._work = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8,
                                     device=self.device)
        code = self.lib.b200_dyn_step(
            self._stream(), nvec, chi_l, chi_r, d2, t.data_ptr(), p1.data_ptr(),
            p2.data_ptr(), v.data_ptr(), v_out.data_ptr(),
            None if cap is None else cap.data_ptr(),
            None if rho_out is None else rho_out.data_ptr(),
            self._work.data_ptr())
        self._check(code, "b200_dyn_step")

    def dyn_run(self, nvec, d2, sites, caps, p1, p2, prop_step_stride, v0, rho_out):
        """The whole compute_dynamics loop in one C call (b200_dyn_run)."""
        n = len(sites)
        chi = (c_int32 * (n + 1))(*([int(t.shape[0]) for t in sites]
                                    + [int(sites[-1].shape[1])]))
        tp = (c_void_p * n)(*[t.data_ptr() for t in sites])
        cp = (c_void_p * (n + 1))(*[c.data_ptr() for c in caps])
        nbytes = self.lib.b200_dyn_run_workspace_bytes(n, nvec, chi, d2)
        work = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        code = self.lib.b200_dyn_run(self._stream(), n, nvec, d2, chi, tp, p1.data_ptr(),
                                     p2.data_ptr(), int(prop_step_stride), cp,
                                     v0.data_ptr(), rho_out.data_ptr(), work.data_ptr())
        self._check(code, "b200_dyn_run")
        return work        # keep alive until the stream has consumed it

    def caps_step(self, chi_l, chi_r, d2, t, cap_next, tr2, cap_out):
        code = self.lib.b200_caps_step(
            self._stream(), chi_l, chi_r, d2, t.data_ptr(), cap_next.data_ptr(),
            tr2.data_ptr(), cap_out.data_ptr())
        self._check(code, "b200_caps_step")

    def launch_count(self):
        return int(self.lib.b200_launch_count())

    def profile_enable(self, on=True):
        self.lib.b200_profile_enable(1 if on else 0)

    def profile_read(self):
        """(kernel_ms, algorithmic_flops, launches, sweeps) of the SVD sweep kernel."""
        ms, fl = c_double(0.0), c_double(0.0)
        n, sw = c_uint64(0), c_uint64(0)
        self._check(self.lib.b200_profile_read(ctypes.byref(ms), ctypes.byref(fl),
                                               ctypes.byref(n), ctypes.byref(sw)),
                    "b200_profile_read")
        return ms.value, fl.value, n.value, sw.value

    def profile_read_kinds(self):
        """Per kernel family of the truncated SVD: ({'jacobi','qrcp','emit','other'} ->
        (ms, launches)), algorithmic flops, Jacobi sweeps."""
        ms, n = (c_double * 4)(), (c_uint64 * 4)()
        fl, sw = c_double(0.0), c_uint64(0)
        self._check(self.lib.b200_profile_read_kinds(ms, n, ctypes.byref(fl),
                                                     ctypes.byref(sw)),
                    "b200_profile_read_kinds")
        names = ("jacobi", "qrcp", "emit", "other")
        return ({k: (ms[i], int(n[i])) for i, k in enumerate(names)}, fl.value, sw.value)

    def synchronize(self):
        torch.cuda.synchronize(self.device)


class NativeChain:
    """Device-resident matrix-product chain owned by the C++ engine (csrc/chain.cu):
    the counterpart of NodeArray for the PT-TEMPO path.  A whole zip-up / svd-sweep is
    ONE C-ABI call."""

    def __init__(self, ops):
        self.ops = ops
        self.lib = ops.lib
        self.h = self.lib.b200_chain_create(ops._stream())
        if not self.h:
            raise B200Error("b200_chain_create failed: "
                            f"{self.lib.b200_last_error().decode()}")
        self._shape = (c_int32 * 3)()
        self._stream_id = ops._stream().value

    def _same_stream(self):
        """The engine launches on the stream it was created with; torch allocations and the
        other ops calls use the CURRENT stream.  Both must be the same stream (create and
        drive a backend inside one ``torch.cuda.stream(...)`` context, or outside of any)."""
        if self.ops._stream().value != self._stream_id:
            raise B200Error("NativeChain was created on another CUDA stream than the current "
                            "one: create and step a backend under the same torch stream")

    def __del__(self):
        try:
            if self.h:
                self.lib.b200_chain_destroy(c_void_p(self.h))
                self.h = None
        except Exception:  # pylint: disable=broad-except
            pass

    def __len__(self):
        return int(self.lib.b200_chain_len(c_void_p(self.h)))

    def push(self, tensor):
        """Append a (chi_l, a, chi_r) complex128 device tensor (copied)."""
        self._same_stream()
        t = tensor.contiguous()
        dl, da, dr = (int(x) for x in t.shape)
        self.ops._check(self.lib.b200_chain_push(c_void_p(self.h), t.data_ptr(), dl, da, dr),
                        "b200_chain_push")

    def shape(self, i):
        self.ops._check(self.lib.b200_chain_shape(c_void_p(self.h), i, self._shape),
                        "b200_chain_shape")
        return tuple(int(x) for x in self._shape)

    def site(self, i):
        """Copy of site i as a torch tensor (chi_l, a, chi_r)."""
        self._same_stream()
        out = self.ops.empty(*self.shape(i))
        self.ops._check(self.lib.b200_chain_read(c_void_p(self.h), i, out.data_ptr()),
                        "b200_chain_read")
        return out

    def svd_sweep(self, from_index, to_index, eps):
        self._same_stream()
        self.ops._check(self.lib.b200_chain_svd_sweep(c_void_p(self.h), from_index, to_index,
                                                      float(eps)), "b200_chain_svd_sweep")

    def pt_zip_up_left(self, mpo, eps):
        """mpo: list of chain.PtSite (kind, device matrix)."""
        self._same_stream()
        arr = (_PtSite * len(mpo))()
        keep = []           # host map arrays stay alive for the call
        for k, site in enumerate(mpo):
            arr[k].kind = PT_KINDS[site.kind]
            m = site.mat
            if m.dim() == 1:
                arr[k].rows, arr[k].cols = int(m.shape[0]), 1
            else:
                arr[k].rows, arr[k].cols = int(m.shape[0]), int(m.shape[1])
            arr[k].mat = m.data_ptr()
            if getattr(site, "maps", None) is not None:
                nmap, wmap = site.maps
                keep += [(c_int32 * len(nmap))(*[int(x) for x in nmap]),
                         (c_int32 * len(wmap))(*[int(x) for x in wmap])]
                arr[k].north_map, arr[k].west_map = keep[-2], keep[-1]
                arr[k].cols = len(nmap)
        self.ops._check(self.lib.b200_chain_pt_zip_up_left(c_void_p(self.h), arr, len(mpo),
                                                           float(eps)),
                        "b200_chain_pt_zip_up_left")

    def tempo_step(self, mpo, p1, p2site, sum_north, d2, eps, state_out):
        """One whole TEMPO time step (b200_chain_tempo_step); mpo: list of chain.TempoSite."""
        self._same_stream()
        arr = (_TempoSite * len(mpo))()
        for k, site in enumerate(mpo):
            arr[k].kind = TEMPO_KINDS[site.kind]
            arr[k].rows, arr[k].cols = int(site.mat.shape[0]), int(site.mat.shape[1])
            arr[k].nw = 0 if site.nw is None else int(site.nw)
            arr[k].ns = 0 if site.ns is None else int(site.ns)
            arr[k].mat = site.mat.data_ptr()
        self.ops._check(self.lib.b200_chain_tempo_step(
            c_void_p(self.h), arr, len(mpo), p1.data_ptr(), p2site.data_ptr(),
            sum_north.data_ptr(), int(d2), float(eps), state_out.data_ptr()),
            "b200_chain_tempo_step")

    def stats(self, reset=False):
        """(truncated SVDs, Jacobi sweeps, D2H bytes) since the last reset."""
        a, b, c = c_uint64(0), c_uint64(0), c_uint64(0)
        self.lib.b200_chain_stats(c_void_p(self.h), ctypes.byref(a), ctypes.byref(b),
                                  ctypes.byref(c), 1 if reset else 0)
        return a.value, b.value, c.value

    def log(self, enable=True, cap=1 << 16):
        """Pop the per-SVD log [(m, n, keep, sweeps)] and switch logging on/off."""
        buf = (c_int32 * (4 * cap))()
        n = self.lib.b200_chain_log(c_void_p(self.h), 1 if enable else 0, buf, cap)
        return [tuple(buf[4 * i: 4 * i + 4]) for i in range(n)]


_default_ops = None


def default_ops():
    """The process-wide CudaOps (raises without CUDA / the built library)."""
    global _default_ops  # pylint: disable=global-statement
    if _default_ops is None:
        dev = int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.is_available() \
            and torch.cuda.device_count() > 1 else 0
        _default_ops = CudaOps(dev)
    return _default_ops
