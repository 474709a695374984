"""GPU diagnostic: per-split (m, n, keep) of the device PT-TEBD run against the oracle's on the
bench_rows.tebd_row workload; prints the first differing split and the final deviation."""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oqupy_b200 as ob  # noqa: E402
from conftest import load_golden, tebd_fixture  # noqa: E402
from oracle import tebd_np  # noqa: E402


def main(n_sites=8, eps=1e-5, steps=4):
    g = load_golden("pt_tebd_F2")
    gammas, lambdas, layers, _, mpos, caps = tebd_fixture(g)
    by_bond = [{} for _ in layers]
    for li, layer in enumerate(layers):
        for sites, tensors in layer:
            by_bond[li][sites[0]] = tensors
    big = []
    for li, layer in enumerate(layers):
        parity = layer[0][0][0] % 2
        src = sorted(by_bond[li])
        big.append([((b, b + 1), by_bond[li][src[(b // 2) % len(src)]])
                    for b in range(parity, n_sites - 1, 2)])
    gam, lam = [gammas[0]] * n_sites, [lambdas[0]] * (n_sites - 1)
    olog = []
    orig = tebd_np.truncated_svd

    def logged(mat, e):
        out = orig(mat, e)
        s_all = np.linalg.svd(mat, compute_uv=False)
        tail = np.sqrt(np.cumsum(s_all[::-1] ** 2))[::-1]
        k = out[1].size
        margin = min(abs(tail[k - 1] / (e * s_all[0]) - 1.0),
                     abs(tail[k] / (e * s_all[0]) - 1.0) if k < s_all.size else 1.0)
        olog.append((mat.shape[0], mat.shape[1], k, margin))
        return out
    tebd_np.truncated_svd = logged
    orc = tebd_np.PtTebdOracle(gam, lam, eps)
    ops = ob.default_ops()
    ops.svd_log = []
    pt = ob.DeviceProcessTensor(2, dt=0.1, ops=ops)
    for k, t in enumerate(mpos):
        pt.set_mpo_tensor(k, t)
    pt.compute_caps()
    be = ob.PtTebdBackend(gam, lam, eps, {}, ops=ops)
    gl = [SimpleNamespace(gates=[SimpleNamespace(sites=list(s), tensors=list(t))
                                 for s, t in layer]) for layer in big]
    for step in range(1, steps + 1):
        for layer, dl in zip(big, gl):
            orc.apply_nn_gate_layer(layer)
            be.apply_nn_gate_layer(dl)
        orc.apply_process_tensors(step, [mpos[step - 1]] * n_sites)
        be.apply_process_tensors(step, [pt] * n_sites)
        for layer, dl in zip(big, gl):
            orc.apply_nn_gate_layer(layer)
            be.apply_nn_gate_layer(dl)
        orc.compute_traces([caps[step]] * n_sites)
        be.compute_traces(step, [pt] * n_sites)
        dev = max(np.abs(orc.get_density_matrix([s]) - be.get_density_matrix([s])).max()
                  for s in range(n_sites))
        print(f"step {step}: max |rho_gpu - rho_oracle| = {dev:.2e}, splits so far "
              f"{len(olog)} / {len(ops.svd_log)}")
    nbad = 0
    for i, (o, d) in enumerate(zip(olog, ops.svd_log)):
        dm, dn = d[0], d[1]
        same_shape = (o[0], o[1]) in ((dm, dn), (dn, dm))
        if not same_shape or o[2] != d[2]:
            nbad += 1
            if nbad <= 10:
                print(f"split {i}: oracle (m,n,keep,margin)={o} device (m,n,keep,sweeps)={d}")
    print("differing splits:", nbad, "of", len(olog))


if __name__ == "__main__":
    main(*(int(x) if i != 1 else float(x) for i, x in enumerate(sys.argv[1:])))
