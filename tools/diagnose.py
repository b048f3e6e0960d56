"""Prints the full per-tensor parity table (CUDA learner vs float64 oracle) -- run on the GPU box:
    python tools/diagnose.py [B] [T] [steps]  > gpurun_out/diagnose.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity  # noqa: E402

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    errs = parity.compare_step(B, T=T, steps=steps)
    bad = parity.failures(errs)
    for k, v in errs.items():
        print("%-28s %.3e%s" % (k, v, "   <-- FAIL" if k in bad else ""))
    print("B=%d T=%d steps=%d: %d entries break the bar" % (B, T, steps, len(bad)))
