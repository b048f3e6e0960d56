"""Times the CPU restatement of the reference at several torch thread counts (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import impala_torch as it, synthetic
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip(), flush=True)
except Exception as e:
    print("no cgroup info", e)
b = synthetic.make_batch(32)
args = [b[k] for k in synthetic.TRAIN_FIELDS]
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    L = it.Learner(None, torch.float32, "reference")
    t0 = time.time(); L.train(*args); t1 = time.time(); L.train(*args); t2 = time.time()
    print("threads %3d: first %.2fs second %.2fs -> %.0f frames/s" % (nt, t1 - t0, t2 - t1, 640 / (t2 - t1)), flush=True)
    if t2 - t1 > 30:
        break
