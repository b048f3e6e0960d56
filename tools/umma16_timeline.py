"""In-kernel timeline of the 16-bit split GEMM kernels (gemm_umma16.cuh::Trace16): the middle CTA of every launch of
one eager learner step reports when its producers, MMA warp, bulk-copy loader and epilogue reach their milestones.
    python tools/umma16_timeline.py [--batch 32]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_reinforcement_learning_b200 import _native as N   # noqa: E402
from distributed_reinforcement_learning_b200.learner import NativeLearner   # noqa: E402
from bench import synth_batch, FIELDS   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    eng = NativeLearner(batch=a.batch, trajectory=20, num_action=18, use_cuda_graph=a.graph, math_mode=5)
    bt = synth_batch(a.batch, 1)
    for s in range(2):
        eng.stage(s, *[bt[f] for f in FIELDS])
    for i in range(4):
        eng.step(i % 2)
    buf = torch.zeros(8002 + 1 + 2 * 12000 + 8, dtype=torch.int64, device="cuda")
    buf[8001] = 0xD17A16
    N.check(N.lib.drl_debug_trace(C.c_void_p(buf.data_ptr())))
    torch.cuda.synchronize()
    eng.step(0)
    torch.cuda.synchronize()
    h = buf.cpu().numpy()
    N.check(N.lib.drl_debug_trace(C.c_void_p(0)))
    n = int(h[8002])
    rec = [(int(h[8003 + 2 * j]), int(h[8004 + 2 * j])) for j in range(min(n, 12000))]
    # split into launches at the start tags
    launches, cur = [], None
    for tag, t in sorted(rec, key=lambda r: r[1]):
        if tag >> 40:
            cur = {"grid": (tag >> 20) & 0xFFFFF, "yz": (tag & 0xFFFFF) >> 10, "bn": tag & 1023, "t0": t, "ev": []}
            launches.append(cur)
        elif cur is not None:
            cur["ev"].append((tag, t - cur["t0"]))
    names = {1: "start", 2: "setup done", 6000: "acc complete", 6001: "epilogue done"}
    for L in launches:
        print("== launch grid.x %d  y*z %d  BN %d: %d events, CTA lifetime %.2f us" % (
            L["grid"], L["yz"], L["bn"], len(L["ev"]), (max(t for _, t in L["ev"]) if L["ev"] else 0) / 1e3))
        for tag, t in L["ev"]:
            if tag in names:
                nm = names[tag]
            elif 600 <= tag < 1000:
                nm = "gather sub %d issued" % (tag - 600)
            elif 1000 <= tag < 2000:
                nm = "  producer: stage free for sub %d" % (tag - 1000)
            elif 2000 <= tag < 3000:
                nm = "  producer: sub %d stored" % (tag - 2000)
            elif 3000 <= tag < 4000:
                nm = "    mma: stage %d full" % (tag - 3000)
            elif 4000 <= tag < 5000:
                nm = "    mma: stage %d issued" % (tag - 4000)
            elif 5000 <= tag < 6000:
                nm = "      loader: bulk copy %d issued" % (tag - 5000)
            else:
                nm = str(tag)
            print("   %8.2f us  %s" % (t / 1e3, nm))
    eng.close()


if __name__ == "__main__":
    main()
