"""ctypes loader of the plain-C V-trace restatement (oracle/c/vtrace_c.c) -- TEST INFRASTRUCTURE ONLY.

``build()`` runs ``make -C oracle/c`` (gcc); the library lands in ``oracle/_build/`` (git-ignored, travels with gpurun
snapshots).  Same function names and layouts as ``oracle/vtrace_np.py``."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libvtrace_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "c", "vtrace_c.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(_HERE, "c"), "-B"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for f in (_lib.vtrace_from_importance_weights, _lib.vtrace_from_softmax, _lib.vtrace_losses):
            f.restype = None
    return _lib


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def _p(a):
    return C.c_void_p(a.ctypes.data)


def from_importance_weights(log_rhos, discounts, rewards, values, bootstrap_value, clip_rho_threshold=1.0,
                            clip_pg_rho_threshold=1.0):
    lr, g, r, v, bo = map(_d, (log_rhos, discounts, rewards, values, bootstrap_value))
    T, B = lr.shape
    vs, rho = np.empty_like(lr), np.empty_like(lr)
    clip = -1.0 if clip_rho_threshold is None else float(clip_rho_threshold)
    lib().vtrace_from_importance_weights(_p(lr), _p(g), _p(r), _p(v), _p(bo), C.c_int(T), C.c_int(B), C.c_double(clip),
                                         _p(vs), _p(rho))
    return vs, rho


def from_softmax(behavior_policy_softmax, target_policy_softmax, actions, discounts, rewards, values, next_values,
                 action_size, clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0):
    mu, pi, g, r, v, nv = map(_d, (behavior_policy_softmax, target_policy_softmax, discounts, rewards, values, next_values))
    a = np.ascontiguousarray(actions, np.int32)
    B, T = a.shape
    vs, rho = np.empty((B, T)), np.empty((B, T))
    clip = -1.0 if clip_rho_threshold is None else float(clip_rho_threshold)
    lib().vtrace_from_softmax(_p(mu), _p(pi), _p(a), _p(g), _p(r), _p(v), _p(nv), C.c_int(B), C.c_int(T),
                              C.c_int(int(action_size)), C.c_double(clip), _p(vs), _p(rho))
    return vs, rho


def losses(softmax, actions, advantages, vs, value):
    """-> (policy-gradient loss, baseline loss, entropy 'loss') of optimizer/vtrace.py:105-126."""
    p, adv, vs_, val = map(_d, (softmax, advantages, vs, value))
    a = np.ascontiguousarray(actions, np.int32)
    B, T = a.shape
    out = np.empty(3)
    lib().vtrace_losses(_p(p), _p(a), _p(adv), _p(vs_), _p(val), C.c_int(B), C.c_int(T), C.c_int(p.shape[-1]), _p(out))
    return float(out[0]), float(out[1]), float(out[2])
