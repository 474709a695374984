"""GPU: factor captured operands (npz of matrices) at a given eps; report keep vs numpy,
singular-value error, reconstruction and orthogonality of the emitted factors."""
import sys
import numpy as np
sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402

path, eps = sys.argv[1], float(sys.argv[2])
ops = ob.default_ops()
Z = np.load(path)
for key in Z.files:
    a = Z[key]
    m, n = a.shape
    d = ops.from_host(a)
    h = ops.svd_factor(d, m, n, n, 1, eps)
    k = h.keep
    s = ops.svd_values(h)
    sref = np.linalg.svd(a, compute_uv=False)
    tail = np.sqrt(np.cumsum(sref[::-1] ** 2))
    kref = int(np.count_nonzero(tail > eps * sref[0]))
    u, vh = ops.empty(m, k), ops.empty(k, n)
    lam = ops.empty(k)
    ops.svd_emit(h, u=u, u_na=1, u_so=k, u_sa=0, u_sj=1, vh=vh, lam=lam)
    hu, hvh, hl = ops.to_host(u), ops.to_host(vh), ops.to_host(lam).real
    ur, sr, vhr = np.linalg.svd(a, full_matrices=False)
    best = (ur[:, :k] * sr[:k]) @ vhr[:k]
    rec = (hu * hl) @ hvh
    print(key, a.shape, "keep", k, "ref", kref, "sweeps", h.sweeps,
          "s err/s0 %.1e" % (np.abs(s - sref).max() / sref[0]),
          "kept s relerr %.1e" % (np.abs(s[:k] - sref[:k]) / sref[:k]).max(),
          "rec-vs-best/s0 %.1e" % (np.abs(rec - best).max() / sref[0]),
          "orthU %.1e orthV %.1e" % (np.abs(hu.conj().T @ hu - np.eye(k)).max(),
                                     np.abs(hvh @ hvh.conj().T - np.eye(k)).max()))
