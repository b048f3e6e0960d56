"""Shared parity harness: runs the CUDA learner (through the C-ABI) and the CPU oracle on the same
seeded synthetic batch and reports per-tensor relative errors.  Used by the -m gpu tests, by
``__graft_entry__.smoke()`` and by ``tools/diagnose.py``.

Error metric (the "1e-4 relative fp32" bar of BASELINE.json's north_star): for each tensor,
    rel = max|gpu - oracle| / max(max|oracle|, 1e-30)
with the oracle evaluated in float64 from the same float32 parameters and inputs.
"""
import numpy as np
import torch

from oracle import impala_torch as it
from oracle import synthetic

TOL = 1e-4          # north_star: within 1e-4 relative fp32
TOL_TIGHT = 2e-5    # what the FP32-FFMA path is expected to meet


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.shape != b.shape:
        return float("inf")
    if not np.all(np.isfinite(a)):
        return float("inf")
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) if b.size else 0.0


def to_time_major(x_bt):
    """[B, T, ...] -> [T*B, ...] rows m = t*B + b (device activation layout)."""
    x = np.asarray(x_bt)
    return np.swapaxes(x, 0, 1).reshape((x.shape[0] * x.shape[1],) + x.shape[2:])


def make_case(B, T=20, A=18, seed=1234, param_seed=0, reward_clipping="abs_one"):
    batch = synthetic.make_batch(B, T=T, A=A, seed=seed)
    params = it.init_params(param_seed, torch.float32, num_action=A)
    cfg = dict(trajectory=T, num_action=A, reward_clipping=reward_clipping)
    return batch, params, cfg


def oracle_step(batch, params, cfg, dtype=torch.float64, steps=1, shaped="dedup"):
    L = it.Learner(params, dtype, shaped, **cfg)
    args = [batch[k] for k in synthetic.TRAIN_FIELDS]
    recs = []
    for _ in range(steps):
        res, out, g, gn = L.train(*args, return_all=True)
        recs.append(dict(res=res, out=out, grads=g, grad_norm=gn))
    return L, recs


def native_learner(batch, params, cfg, **kw):
    from distributed_reinforcement_learning_b200.learner import NativeLearner
    B = batch["state"].shape[0]
    eng = NativeLearner(batch=B, trajectory=cfg["trajectory"], num_action=cfg["num_action"],
                        reward_clipping=cfg.get("reward_clipping", "abs_one"), **kw)
    eng.set_params(it.flatten_params(params))
    return eng


def compare_step(B, T=20, A=18, seed=1234, steps=1, reward_clipping="abs_one", layers=True, **kw):
    """Returns {name: rel_err} for taps, losses, every parameter gradient, and the parameters / RMSProp
    slots after `steps` updates."""
    batch, params, cfg = make_case(B, T, A, seed, reward_clipping=reward_clipping)
    L, recs = oracle_step(batch, params, cfg, torch.float64, steps)
    eng = native_learner(batch, params, cfg, **kw)
    args = [batch[k] for k in synthetic.TRAIN_FIELDS]
    errs = {}
    try:
        for s in range(steps):
            eng.stage(s % eng.num_slots, *args)
            out = eng.step(s % eng.num_slots)
            rec = recs[s]
            tag = "" if steps == 1 else "@%d" % s
            o = rec["out"]
            for k, ok in (("pi_loss", "pi_loss"), ("baseline_loss", "baseline_loss"), ("entropy", "entropy"),
                          ("total_loss", "total_loss")):
                errs["loss/" + k + tag] = rel_err(out[k], float(o[ok].detach()))
            errs["lr" + tag] = abs(out["learning_rate"] - rec["res"][3])
            errs["grad_norm" + tag] = rel_err(out["grad_norm"], rec["grad_norm"])
            if out["step"] != s + 1:
                errs["step" + tag] = float("inf")
            if s == 0:
                taps = eng.taps()
                for k in ("vs", "clipped_rho", "vs_plus_1", "pg_advantage"):
                    errs["tap/" + k] = rel_err(taps[k], o[k].detach().numpy())
                gflat = eng.get_grads()
                gd = it.unflatten_params(gflat, torch.float64, num_action=A)
                for n in rec["grads"]:
                    errs["grad/" + n] = rel_err(gd[n].numpy(), rec["grads"][n].detach().numpy())
                if layers:
                    M = B * T
                    tp = o["taps"]
                    errs["act/policy"] = rel_err(eng.read_buffer("policy", M * A).reshape(M, A),
                                                 to_time_major(tp["policy"].detach().numpy()))
                    errs["act/value"] = rel_err(eng.read_buffer("value", M), to_time_major(tp["value"].detach().numpy()))
                    errs["act/h1"] = rel_err(eng.read_buffer("h1", M * 256).reshape(M, 256),
                                             to_time_major(tp["h1"].detach().numpy()))
                    errs["act/c1"] = rel_err(eng.read_buffer("c1", M * 256).reshape(M, 256),
                                             to_time_major(tp["c1"].detach().numpy()))
                    for nm, shp in (("a1", (20, 20, 32)), ("a2", (9, 9, 64)), ("a3", (7, 7, 64))):
                        ref = tp[nm].detach().numpy().reshape((B, T) + shp)
                        got = eng.read_buffer(nm, M * int(np.prod(shp))).reshape((M,) + shp)
                        errs["act/" + nm] = rel_err(got, to_time_major(ref))
        pf = eng.get_params()
        ms, st = eng.get_opt_state()
        pd = it.unflatten_params(pf, torch.float64, num_action=A)
        md = it.unflatten_params(ms, torch.float64, num_action=A)
        for n in L.params:
            errs["param/" + n] = rel_err(pd[n].numpy(), L.params[n].detach().numpy())
            # the applied update w_after - w_before, judged against its own size plus the float32 storage
            # granularity of the parameter it is added to (the GPU stores w in float32, the oracle in float64)
            p0 = params[n].double().numpy()
            du_got, du_exp = pd[n].numpy() - p0, L.params[n].detach().numpy() - p0
            floor = 4.0 * np.finfo(np.float32).eps * max(np.max(np.abs(p0)), 1e-30) / TOL
            errs["update/" + n] = float(np.max(np.abs(du_got - du_exp)) / (np.max(np.abs(du_exp)) + floor))
            errs["ms/" + n] = rel_err(md[n].numpy(), L.ms[n].detach().numpy())
        if st != steps:
            errs["opt_step"] = float("inf")
    finally:
        eng.close()
    return errs


def worst(errs, prefix=""):
    items = [(v, k) for k, v in errs.items() if k.startswith(prefix)]
    return max(items) if items else (0.0, "")
