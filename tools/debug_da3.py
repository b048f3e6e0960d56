"""Debug: where do the backward activations differ from the oracle? (run on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import parity
from oracle import impala_torch as it, synthetic
B, T, A = int(sys.argv[1]), int(sys.argv[2]), 18
batch, params, cfg = parity.make_case(B, T, A)
L = it.Learner(params, torch.float64, "dedup", **cfg)
args = [batch[k] for k in synthetic.TRAIN_FIELDS]
o, g = L.gradients(*args)
eng = parity.native_learner(batch, params, cfg)
eng.stage(0, *args); eng.forward_backward(0)
Mb, M = B * (T - 2), B * T
tp, ag = o["taps"], o["act_grads"]
for nm, shp in (("a3", (7, 7, 64)), ("a2", (9, 9, 64)), ("a1", (20, 20, 32))):
    post = tp[nm].detach().numpy().reshape((B, T) + shp)
    gr = ag[nm].detach().numpy().reshape((B, T) + shp) * (post > 0)
    ref = parity.to_time_major(gr)[:Mb].reshape(Mb, -1)
    got = eng.read_buffer("d" + nm, Mb * int(np.prod(shp))).reshape(Mb, -1)
    err = np.abs(got - ref)
    thr = 1e-4 * np.abs(ref).max()
    bad = err > thr
    rows = np.where(bad.any(1))[0]
    cols = np.where(bad.any(0))[0]
    print("d" + nm, "max err", err.max(), "ref max", np.abs(ref).max(), "bad elems", bad.sum(), "of", bad.size)
    print("  bad rows (count %d):" % len(rows), rows[:40])
    print("  bad cols (count %d):" % len(cols), cols[:40])
    if bad.any():
        i, j = np.unravel_index(np.argmax(err), err.shape)
        print("  worst at row %d col %d: got %g ref %g ; got==0? %s" % (i, j, got[i, j], ref[i, j], got[i, j] == 0))
        # is `got` equal to ref at a shifted location?
        r = rows[0]
        print("  row %d: n bad cols %d ; first bad cols %s" % (r, bad[r].sum(), np.where(bad[r])[0][:20]))
        print("  got:", got[r, np.where(bad[r])[0][:6]], " ref:", ref[r, np.where(bad[r])[0][:6]])
# dz check through h1 grads is not available; check du via emb grads and the LSTM input-gradient directly
h1g = ag.get("h1")
eng.close()
