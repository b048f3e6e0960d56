"""Diagnostics for math mode 5 (16-bit split operands): per-case errors of the debug GEMM and per-tensor parity errors of
a learner step, printed rather than asserted."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from distributed_reinforcement_learning_b200 import _native as native
    from test_gpu_umma import _gemm
    print("== debug GEMM (core 2 = 3xTF32, 5 = bf16 split, 6 = fp16 split, 7/8 = pre-tiled B)")
    for core in (2, 5, 6, 7, 8):
        for a_km, b_km in ((1, 0), (1, 1), (0, 0), (0, 1)):
            for bn in (32, 64, 128, 256):
                if core in (5, 6) and bn == 32 and not b_km:
                    continue
                try:
                    err, cs = _gemm(native, core, bn, a_km, b_km, 256, 128, 192, 1)
                    print("core %d a_km %d b_km %d bn %3d  err %.3e  colsum %s" % (core, a_km, b_km, bn, err, cs))
                except Exception as ex:
                    print("core %d a_km %d b_km %d bn %3d  FAILED %s" % (core, a_km, b_km, bn, ex))
    for (M, N, K, sp, bn) in ((640, 1024, 3648, 7, 256), (3648, 1024, 576, 1, 256), (576, 3392, 1024, 1, 128), (51840, 64, 512, 1, 64)):
        for core in (2, 5, 6, 7):
            try:
                err, cs = _gemm(native, core, bn, 1, 0 if core != 7 else 0, M, N, K, sp, seed=1)
                print("big  core %d %dx%dx%d splits %d bn %d  err %.3e" % (core, M, N, K, sp, bn, err))
            except Exception as ex:
                print("big  core %d %dx%dx%d FAILED %s" % (core, M, N, K, ex))
    import parity
    for mode in (2, 5):
        for (B, T, kw) in ((4, 20, {}), (32, 20, {"layers": False})):
            try:
                errs = parity.compare_step(B, T=T, math_mode=mode, **kw)
            except Exception as ex:
                print("mode %d B %d: FAILED %s" % (mode, B, ex))
                continue
            bad = parity.failures(errs)
            top = sorted(((v, k) for k, v in errs.items() if not k.startswith(("kink", "lr"))), reverse=True)[:14]
            print("== step mode %d B=%d T=%d: %d entries, %d failures; worst:" % (mode, B, T, len(errs), len(bad)))
            for v, k in top:
                print("   %-28s %.3e" % (k, v))
            print("   kink flip fraction %.2e max |x| at flip %.2e" % (errs["kink/flip_fraction"], errs["kink/max_abs_at_flip"]))


if __name__ == "__main__":
    main()
