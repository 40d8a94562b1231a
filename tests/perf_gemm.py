"""All GEMM perf cases of tests/bringup_gemm.py in one process (device-timed, algorithmic TFLOP/s)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bringup_gemm as bg  # noqa: E402

if __name__ == "__main__":
    names = sys.argv[1:] or [n for n in bg.CASES if n.startswith(("perf_", "wperf_"))]
    for n in names:
        fn, kw = bg.CASES[n]
        fn(**kw)
