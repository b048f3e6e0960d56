"""Feasibility probe for micro-batch pipelining: do TWO independent B/2 learner handles stepping concurrently on one
GPU (each with its own streams and CUDA graph) deliver more frames/s than one handle at B?  If the step is a
latency-bound chain that leaves SMs idle, two chains should interleave; if it is throughput-bound they should not.

    python tools/two_handles_probe.py [--batch 32] [--steps 40]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_reinforcement_learning_b200.learner import NativeLearner   # noqa: E402
from distributed_reinforcement_learning_b200.model import impala_actor_critic as model   # noqa: E402
from oracle import synthetic   # noqa: E402


def run(handles, steps, warm=5):
    for e, b in handles:
        for s in (0, 1):
            e.stage(s, *[b[k] for k in synthetic.TRAIN_FIELDS])
    for i in range(warm):
        for e, _ in handles:
            e.step_async(i % 2)
        for e, _ in handles:
            e.wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        for e, _ in handles:
            e.step_async(i % 2)
        if i % 4 == 3 or i == steps - 1:    # keep the host a few steps ahead, not unboundedly
            for e, _ in handles:
                e.wait()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    T = 20
    params = model.init_params(seed=0)
    for parts in (1, 2, 4):
        Bp = a.batch // parts
        hs = []
        for p in range(parts):
            e = NativeLearner(batch=Bp, trajectory=T, num_action=18, device=0, use_cuda_graph=True)
            e.set_params(params)
            hs.append((e, synthetic.make_batch(Bp, T=T, A=18, seed=3 + p)))
        ms = run(hs, a.steps)
        print("%d handle(s) x B=%d: %.4f ms per round of %d trajectories = %.0f frames/s" %
              (parts, Bp, ms, a.batch, a.batch * T / ms * 1e3), flush=True)
        for e, _ in hs:
            e.close()


if __name__ == "__main__":
    main()
