"""GPU: time every SVD of one late PT-TEMPO step (config 2) and print the heavy hitters."""
import os
import sys
import time

os.environ.setdefault("OQUPY_B200_PYCHAIN", "1")   # per-SVD instrumentation needs the Python chain

import numpy as np
import torch

sys.path.insert(0, ".")
import oqupy_b200 as ob  # noqa: E402


def main():
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    with np.load("tests/golden/c2_operands.npz") as f:
        infl = f["influences"]
    ops = ob.default_ops()
    be = ob.PtTempoBackend(2, lambda dk: None if dk < 0 else infl[dk], None,
                           np.ones(4), np.ones(4), 1000, 200, 1e-9)
    be.initialize()
    for _ in range(nsteps):
        be.compute_step()
    rec = []
    orig = ops.svd_factor

    def timed(theta, m, n, rs, cs, eps, off=0, **kw):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = orig(theta, m, n, rs, cs, eps, off, **kw)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        pc = ops.svd_phase_cycles(h)
        rec.append((m, n, h.keep, h.sweeps, ms, pc, h.rotations))
        return h

    ops.svd_factor = timed
    t0 = time.perf_counter()
    be.compute_step()
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
    svd_ms = sum(r[4] for r in rec)
    print(f"step {be.step}: total {total:.1f} ms, svd_factor {svd_ms:.1f} ms over {len(rec)} calls")
    rec_sorted = sorted(rec, key=lambda r: -r[4])
    print("top 15 (m, n, keep, sweeps, ms):")
    for r in rec_sorted[:15]:
        q = min(r[0], r[1]); nb = (q + 15) // 16; nb += nb & 1
        print("  ", r[0], r[1], r[2], r[3], round(r[4], 2), "phase kcyc", [round(c / 1e3) for c in r[5][:10]], "stages", r[5][15],
              "rotated stage-slots", r[6], "of", r[3] * (nb - 1) * (nb // 2))
    print("typical mid-size (every 25th call):")
    for r in rec[5::25]:
        print("  ", r[0], r[1], r[2], r[3], round(r[4], 3), "phase kcyc", [round(c / 1e3) for c in r[5][:10]], "stages", r[5][15])
    # histogram by min dimension
    bins = [0, 32, 64, 128, 256, 512, 1024, 4096]
    for lo, hi in zip(bins[:-1], bins[1:]):
        sel = [r for r in rec if lo < min(r[0], r[1]) <= hi]
        if sel:
            print(f"min(m,n) in ({lo},{hi}]: {len(sel)} svds, {sum(r[4] for r in sel):.1f} ms, "
                  f"mean sweeps {np.mean([r[3] for r in sel]):.1f}")


if __name__ == "__main__":
    main()
