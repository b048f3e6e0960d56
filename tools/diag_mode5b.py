import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from distributed_reinforcement_learning_b200 import _native as native
from test_gpu_umma import _gemm
for core in ():
    for a_km, b_km in ((1, 0), (1, 1)):
        try:
            print("mixed core", core, a_km, b_km, _gemm(native, core, 64, a_km, b_km, 256, 128, 192, 1), flush=True)
        except Exception as ex:
            print("mixed core", core, "FAILED", ex, flush=True)
from distributed_reinforcement_learning_b200.learner import NativeLearner
from oracle import impala_torch as it, synthetic
B, T = 2, 5
batch = synthetic.make_batch(B, T=T, seed=5)
eng = NativeLearner(batch=B, trajectory=T, num_action=18, math_mode=5, use_cuda_graph=False)
eng.set_params(it.flatten_params(it.init_params(0)))
eng.stage(0, *[batch[f] for f in synthetic.TRAIN_FIELDS])
try:
    eng.forward(0)
    print("forward ok", flush=True)
    print(eng.step(0), flush=True)
except Exception as ex:
    print("FAILED:", ex, flush=True)
