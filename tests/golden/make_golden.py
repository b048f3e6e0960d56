"""Generates the committed golden vectors for the IMPALA learner step (and the Ape-X / R2D2 / A3C steps) from the
float64 oracles.

    python tests/golden/make_golden.py [impala] [vtrace] [apex] [r2d2] [a3c]     # default: all; writes tests/golden/*.npz

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


def _pack_grads(rec, g):
    for n, v in g.items():
        v = v.detach().numpy()
        if v.size > 70000:
            rec["gradsample_" + n] = v.ravel()[::STRIDE].astype(np.float32)
            rec["gradl2_" + n] = np.float64(np.sqrt(np.sum(v.astype(np.float64) ** 2)))
        else:
            rec["grad_" + n] = v.astype(np.float32)


def apex_case(path, B=3, A=4, seed=1357):
    """oracle/apex_torch.py: one distributed_train step (inputs regenerated from make_transitions(B, A, seed),
    parameters from init_params(0) / init_params(1))."""
    from oracle import apex_torch as ax
    b = ax.make_transitions(B, A=A, seed=seed)
    L = ax.Learner(dtype=torch.float64, num_action=A)
    (loss, td), out, g, gn, lr = L.distributed_train(*[b[k] for k in ax.TRAIN_FIELDS], return_all=True)
    rec = dict(B=B, A=A, seed=seed, loss=loss, td_error=td, grad_norm=gn, learning_rate=lr)
    for k in ("main_q", "next_main_q", "target_q", "target_value", "state_action_value"):
        rec[k] = out[k].detach().numpy()
    _pack_grads(rec, g)
    np.savez_compressed(path, **rec)
    return rec


def r2d2_case(path, B=2, S=6, bi=2, seed=2468):
    from oracle import r2d2_torch as rt
    b = rt.make_sequences(B, S=S, seed=seed)
    L = rt.Learner(dtype=torch.float64, seq_len=S, burn_in=bi)
    (loss, td), out, g, gn = L.train(*[b[k] for k in rt.TRAIN_FIELDS], return_all=True)
    rec = dict(B=B, S=S, burn_in=bi, seed=seed, loss=loss, td_error=td, grad_norm=gn)
    for k in ("main_q", "target_q", "target_value", "state_action_value"):
        rec[k] = out[k].detach().numpy()
    _pack_grads(rec, g)
    np.savez_compressed(path, **rec)
    return rec


def a3c_case(path, B=3, A=4, seed=97531):
    from oracle import a3c_torch as at
    b = at.make_transitions(B, A=A, seed=seed)
    L = at.Learner(dtype=torch.float64, num_action=A)
    (pi, bl, en, lr), out, g, gn = L.train(*[b[k] for k in at.TRAIN_FIELDS], return_all=True)
    rec = dict(B=B, A=A, seed=seed, pi_loss=pi, baseline_loss=bl, entropy=en, learning_rate=lr, grad_norm=gn)
    for k in ("policy", "value", "next_value", "advantage"):
        rec[k] = out[k].detach().numpy()
    _pack_grads(rec, g)
    np.savez_compressed(path, **rec)
    return rec


if __name__ == "__main__":
    which = set(sys.argv[1:]) or {"impala", "vtrace", "apex", "r2d2", "a3c"}
    if "impala" in which:
        r = learner_case(2, 6, 18, 4321, os.path.join(HERE, "impala_step_B2_T6.npz"))
        print("pi %.6f bl %.6f ent %.6f gn %.6f" % (r["pi_loss"], r["baseline_loss"], r["entropy"], r["grad_norm"]))
    if "vtrace" in which:
        vtrace_case(os.path.join(HERE, "vtrace_T18_B8.npz"))
    if "apex" in which:
        r = apex_case(os.path.join(HERE, "apex_step_B3.npz"))
        print("apex loss %.6f gn %.6f" % (r["loss"], r["grad_norm"]))
    if "r2d2" in which:
        r = r2d2_case(os.path.join(HERE, "r2d2_step_B2_S6.npz"))
        print("r2d2 loss %.6f gn %.6f" % (r["loss"], r["grad_norm"]))
    if "a3c" in which:
        r = a3c_case(os.path.join(HERE, "a3c_step_B3.npz"))
        print("a3c pi %.6f bl %.6f ent %.6f gn %.6f" % (r["pi_loss"], r["baseline_loss"], r["entropy"], r["grad_norm"]))
