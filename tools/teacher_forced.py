"""GPU: teacher-forced parity of the truncated SVD at bench scale.

The CPU oracle runs the config-2 build; every truncated-SVD operand of the chosen steps is
also factorised by the CUDA kernels.  Output (one JSON document): counts, and every call
whose ranks differ or whose reference cut is closer than 1e-4 to a tie, with its margin
|tail/(eps s0) - 1|.

    python tools/teacher_forced.py [first=7] [last=26] > profiles/r02_teacher_forced_7_26.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_parity_gpu as tp  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    log = []
    t0 = time.perf_counter()
    stats = tp.teacher_forced_stream(first, last, log)
    stats["mismatch"] = [list(map(float, x)) for x in stats["mismatch"]]
    print(json.dumps({"steps": [first, last], "seconds": round(time.perf_counter() - t0, 1),
                      "stats": stats, "calls_near_a_tie_or_different": log}, indent=1))


if __name__ == "__main__":
    main()
