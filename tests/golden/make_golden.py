"""Generates the committed golden vectors of the IMPALA learner step, of optimizer/vtrace.py and of the Ape-X / R2D2
steps by EXECUTING THE UNMODIFIED REFERENCE FILES (``/root/reference/agent/impala.py`` etc.) over the TF 1.14 API
stand-in ``oracle/tf1_shim`` in float64 (``oracle/ref_exec.py``), the A3C family included.

    python tests/golden/make_golden.py [impala] [vtrace] [apex] [r2d2] [a3c]     # default: all; writes tests/golden/*.npz

Needs the reference checkout (build container only).  Every file records ``source``.  What these vectors pin: the
restatements under ``oracle/`` (CPU suite, float64, ~1e-12) and the CUDA path (``-m gpu`` suite, 1e-4) against the
reference's own graph-construction code; TensorFlow's op kernels themselves are restated by the shim (see its
docstring) -- ``tests/golden/make_golden_tf1.py`` regenerates the same files under a real tensorflow==1.14.0 where
one exists.  Inputs are regenerated from the stored seed by oracle/synthetic.py::make_batch and
oracle/impala_torch.py::init_params(0).
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


SOURCE = "reference files executed over oracle/tf1_shim (float64)"


def learner_case(B, T, A, seed, path):
    """agent/impala.py:132-148 ``Agent.train`` executed once; taps/gradients fetched from the agent's own graph."""
    from oracle import ref_exec
    batch = synthetic.make_batch(B, T=T, A=A, seed=seed)
    params = it.init_params(0, torch.float32, num_action=A)
    args = [batch[k] for k in synthetic.TRAIN_FIELDS]
    R = ref_exec.ReferenceImpala(params, trajectory=T, num_action=A)
    f = R.fetch(args, ["vs", "clipped_rho", "vs_plus_1", "pg_advantage", "unrolled_first_policy",
                       "unrolled_first_value", "total_loss"])
    g = {n: torch.from_numpy(np.asarray(v)) for n, v in R.gradients(args).items()}
    gn = float(np.sqrt(sum(float(torch.sum(v.double() ** 2)) for v in g.values())))
    pi, bl, en, lr = R.train(*args)
    rec = dict(B=B, T=T, A=A, seed=seed, pi_loss=pi, baseline_loss=bl, entropy=en, learning_rate=lr, grad_norm=gn,
               total_loss=float(f["total_loss"]), source=SOURCE)
    for k, src in (("vs", "vs"), ("clipped_rho", "clipped_rho"), ("vs_plus_1", "vs_plus_1"),
                   ("pg_advantage", "pg_advantage"), ("first_policy", "unrolled_first_policy"),
                   ("first_value", "unrolled_first_value")):
        rec[k] = f[src]
    _pack_grads(rec, g)
    # the applied update (clip_by_global_norm 40 + TF1 RMSProp, one step): strided samples of the new parameters / slots
    for n, v in R.params().items():
        rec["paramsample_" + n] = v.ravel()[::STRIDE]
    for n, v in R.rms().items():
        rec["rmssample_" + n] = v.ravel()[::STRIDE]
    np.savez_compressed(path, **rec)
    return rec


def vtrace_case(path):
    rng = np.random.default_rng(2024)
    T, B = 18, 8
    kw = dict(log_rhos=rng.standard_normal((T, B)) * 0.7, discounts=(rng.random((T, B)) > 0.1) * 0.99,
              rewards=rng.standard_normal((T, B)), values=rng.standard_normal((T, B)),
              bootstrap_value=rng.standard_normal(B))
    from oracle import ref_exec
    ref = ref_exec.load(("optimizer.vtrace",), float_dtype=torch.float64)
    tf, vt = ref.tf, ref["optimizer.vtrace"]
    vs, rho = tf.Session().run(list(vt.from_importance_weights(**{k: tf.constant(np.asarray(v, np.float64))
                                                                   for k, v in kw.items()})))
    np.savez_compressed(path, vs=vs, clipped_rhos=rho, source=SOURCE, **kw)


def _pack_grads(rec, g):
    for n, v in g.items():
        v = v.detach().numpy() if hasattr(v, "detach") else np.asarray(v)
        if v.size > 70000:
            rec["gradsample_" + n] = v.ravel()[::STRIDE].astype(np.float32)
            rec["gradl2_" + n] = np.float64(np.sqrt(np.sum(v.astype(np.float64) ** 2)))
        else:
            rec["grad_" + n] = v.astype(np.float32)


def apex_case(path, B=3, A=4, seed=1357):
    """agent/apex.py:135-154 ``distributed_train`` executed once (inputs regenerated from make_transitions(B, A, seed),
    parameters from init_params(0) / init_params(1))."""
    from oracle import apex_torch as ax
    from oracle import ref_exec
    b = ax.make_transitions(B, A=A, seed=seed)
    p, tp = ax.init_params(0, torch.float32, num_action=A), ax.init_params(1, torch.float32, num_action=A)
    R = ref_exec.ReferenceApex(p, tp, num_action=A)
    args = [b[k] for k in ax.TRAIN_FIELDS]
    f = R.fetch(args, ["main_q_value", "next_main_q_value", "target_q_value", "target_value", "state_action_value",
                       "learning_rate"])
    g = {n: torch.from_numpy(np.asarray(v)) for n, v in R.gradients(R.feed(*args)).items()}
    gn = float(np.sqrt(sum(float(torch.sum(v.double() ** 2)) for v in g.values())))
    loss, td = R.agent.distributed_train(*args)
    rec = dict(B=B, A=A, seed=seed, loss=loss, td_error=td, grad_norm=gn, learning_rate=float(f["learning_rate"]),
               source=SOURCE)
    for k, src in (("main_q", "main_q_value"), ("next_main_q", "next_main_q_value"), ("target_q", "target_q_value"),
                   ("target_value", "target_value"), ("state_action_value", "state_action_value")):
        rec[k] = f[src]
    _pack_grads(rec, g)
    np.savez_compressed(path, **rec)
    return rec


def r2d2_case(path, B=2, S=6, bi=2, seed=2468):
    """agent/r2d2.py:132-159 ``Agent.train`` executed once."""
    from oracle import r2d2_torch as rt
    from oracle import ref_exec
    b = rt.make_sequences(B, S=S, seed=seed)
    p, tp = rt.init_params(0, torch.float32), rt.init_params(1, torch.float32)
    R = ref_exec.ReferenceR2D2(p, tp, seq_len=S, burn_in=bi)
    args = [b[k] for k in rt.TRAIN_FIELDS]
    f = R.fetch(args, ["main_q", "target_q", "target_value", "state_action_value"])
    g = {n: torch.from_numpy(np.asarray(v)) for n, v in R.gradients(R.feed(*args)).items()}
    gn = float(np.sqrt(sum(float(torch.sum(v.double() ** 2)) for v in g.values())))
    loss, td = R.agent.train(*args)
    rec = dict(B=B, S=S, burn_in=bi, seed=seed, loss=loss, td_error=td, grad_norm=gn, source=SOURCE)
    for k in ("main_q", "target_q", "target_value", "state_action_value"):
        rec[k] = f[k]
    _pack_grads(rec, g)
    np.savez_compressed(path, **rec)
    return rec


def a3c_case(path, B=3, A=4, seed=97531):
    """agent/a3c.py:85-103 ``Agent.train`` executed once (inputs regenerated from make_transitions(B, A, seed), parameters
    from init_params(0)); taps and gradients fetched from the agent's own graph."""
    from oracle import a3c_torch as at
    from oracle import ref_exec
    b = at.make_transitions(B, A=A, seed=seed)
    p = at.init_params(0, torch.float32, num_action=A)
    R = ref_exec.ReferenceA3C(p, num_action=A)
    args = [b[k] for k in at.TRAIN_FIELDS]
    f = R.fetch(args, ["policy", "value", "next_value", "clipped_r_ph", "discounts"])
    g = R.gradients(args)
    gn = float(np.sqrt(sum(np.sum(np.asarray(v, np.float64) ** 2) for v in g.values())))
    pi, bl, en, lr = R.agent.train(*args)
    rec = dict(B=B, A=A, seed=seed, pi_loss=float(pi), baseline_loss=float(bl), entropy=float(en), learning_rate=float(lr),
               grad_norm=gn, source=SOURCE)
    for k in ("policy", "value", "next_value"):
        rec[k] = np.asarray(f[k])
    rec["advantage"] = np.asarray(f["clipped_r_ph"]) + np.asarray(f["discounts"]) * rec["next_value"] - rec["value"]
    _pack_grads(rec, g)
    np.savez_compressed(path, **rec)
    return rec


if __name__ == "__main__":
    which = set(sys.argv[1:]) or {"impala", "vtrace", "apex", "r2d2", "a3c"}
    if "impala" in which:
        for B, T, seed in ((2, 6, 4321), (4, 20, 1234)):        # small case; BASELINE configs[0] (T=20, B=4)
            r = learner_case(B, T, 18, seed, os.path.join(HERE, "impala_step_B%d_T%d.npz" % (B, T)))
            print("B%d T%d: pi %.6f bl %.6f ent %.6f gn %.6f" % (B, T, r["pi_loss"], r["baseline_loss"], r["entropy"],
                                                                 r["grad_norm"]))
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
