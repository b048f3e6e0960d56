"""clock64() timeline of CTA 0 of conv1_fwd_tma_kernel (converter thread 0, MMA lane, epilogue warp 8 lane 0, TMA lane)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_reinforcement_learning_b200 import _native as N   # noqa: E402
from distributed_reinforcement_learning_b200.learner import NativeLearner   # noqa: E402
from bench import synth_batch, FIELDS   # noqa: E402

eng = NativeLearner(batch=32, trajectory=20, num_action=18, use_cuda_graph=False, math_mode=5)
bt = synth_batch(32, 1)
eng.stage(0, *[bt[f] for f in FIELDS])
for i in range(3):
    eng.forward(0)
buf = torch.zeros(8002 + 4 * 96 + 8, dtype=torch.int64, device="cuda")
buf[8001] = 0xC0171
N.check(N.lib.drl_debug_trace(C.c_void_p(buf.data_ptr())))
torch.cuda.synchronize()
eng.forward(0)
torch.cuda.synchronize()
h = buf.cpu().numpy()
N.check(N.lib.drl_debug_trace(C.c_void_p(0)))
d = h[8002:8002 + 4 * 96].reshape(4, 96)
t0 = min(int(x) for x in d.ravel() if x > 0)
names = ["converter t0 (pairs: stage free, stage published)", "mma lane (stage full)", "epilogue warp 8 (pairs: acc full, stored)", "tma lane (slot free)"]
for r in range(4):
    v = [int(x) - t0 for x in d[r] if x > 0]
    print(names[r], len(v))
    print("   ", " ".join("%d" % x for x in v))
eng.close()
