"""Generates the committed golden vectors for the IMPALA learner step from the float64 oracle.

    python tests/golden/make_golden.py        # writes tests/golden/*.npz

PARITY UNPINNED: the reference (TF 1.14) cannot be imported in this image, so these vectors pin the
ORACLE (oracle/impala_torch.py, float64) -- not TensorFlow -- against regressions, and give the GPU tests
a fixture that does not depend on re-running the oracle.  Inputs are regenerated from the stored seed by
oracle/synthetic.py::make_batch and oracle/impala_torch.py::init_params(0).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import impala_torch as it  # noqa: E402
from oracle import synthetic, vtrace_np  # noqa: E402


STRIDE = 61


def learner_case(B, T, A, seed, path):
    batch = synthetic.make_batch(B, T=T, A=A, seed=seed)
    params = it.init_params(0, torch.float32, num_action=A)
    L = it.Learner(params, torch.float64, "reference", trajectory=T, num_action=A)
    (pi, bl, en, lr), out, g, gn = L.train(*[batch[k] for k in synthetic.TRAIN_FIELDS], return_all=True)
    rec = dict(B=B, T=T, A=A, seed=seed, pi_loss=pi, baseline_loss=bl, entropy=en, learning_rate=lr, grad_norm=gn,
               total_loss=float(out["total_loss"]))
    for k in ("vs", "clipped_rho", "vs_plus_1", "pg_advantage", "first_policy", "first_value"):
        rec[k] = out[k].detach().numpy()
    for n, v in g.items():
        v = v.detach().numpy()
        # big tensors (LSTM kernel: 3.7M floats): strided sample + l2 norm; full tensors for the rest
        if v.size > 70000:
            rec["gradsample_" + n] = v.ravel()[::STRIDE].astype(np.float32)
            rec["gradl2_" + n] = np.float64(np.sqrt(np.sum(v.astype(np.float64) ** 2)))
        else:
            rec["grad_" + n] = v.astype(np.float32)
    np.savez_compressed(path, **rec)
    return rec


def vtrace_case(path):
    rng = np.random.default_rng(2024)
    T, B = 18, 8
    kw = dict(log_rhos=rng.standard_normal((T, B)) * 0.7, discounts=(rng.random((T, B)) > 0.1) * 0.99,
              rewards=rng.standard_normal((T, B)), values=rng.standard_normal((T, B)),
              bootstrap_value=rng.standard_normal(B))
    vs, rho = vtrace_np.from_importance_weights(**kw)
    np.savez_compressed(path, vs=vs, clipped_rhos=rho, **kw)


if __name__ == "__main__":
    r = learner_case(2, 6, 18, 4321, os.path.join(HERE, "impala_step_B2_T6.npz"))
    print("pi %.6f bl %.6f ent %.6f gn %.6f" % (r["pi_loss"], r["baseline_loss"], r["entropy"], r["grad_norm"]))
    vtrace_case(os.path.join(HERE, "vtrace_T18_B8.npz"))
