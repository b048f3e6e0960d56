"""Where does the e2e loop lose time against the device-resident loop?  Times, per step: (a) the bench's e2e loop,
(b) the same without the per-step H2D staging, (c) the H2D staging alone (copy stream only)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_reinforcement_learning_b200.learner import NativeLearner   # noqa: E402
from bench import synth_batch   # noqa: E402

B, T, K = 32, 20, 100
F = ("state", "reward", "action", "done", "behavior_policy", "previous_action", "initial_h", "initial_c")
eng = NativeLearner(batch=B, trajectory=T, num_action=18, use_cuda_graph=True)
hb = []
for i in range(3):
    bt = synth_batch(B, i)
    hb.append([torch.from_numpy(np.ascontiguousarray(bt[f])).pin_memory().numpy() for f in F])
for s in range(2):
    eng.stage(s, *hb[s])
for i in range(6):
    eng.step(i % 2)


def loop(stage):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        eng.step_async(i % 2)
        if stage and i + 1 < K:
            eng.stage((i + 1) % 2, *hb[(i + 1) % 3])
        if i >= 1:
            eng.wait((i - 1) % 2)
    eng.wait((K - 1) % 2)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3


for rep in range(2):
    print("e2e loop with H2D staging   : %.4f ms/step" % loop(True))
    print("e2e loop without staging    : %.4f ms/step" % loop(False))
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    eng.stage(i % 2, *hb[i % 3])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 50
nbytes = sum(a.nbytes for a in hb[0])
print("H2D staging alone           : %.4f ms/batch of %.2f MB = %.1f GB/s" % (dt * 1e3, nbytes / 1e6, nbytes / dt / 1e9))
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dev = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    dev.copy_(big, non_blocking=True)
torch.cuda.synchronize()
print("plain 256 MB pinned H2D     : %.1f GB/s" % (10 * big.numel() / (time.perf_counter() - t0) / 1e9))
