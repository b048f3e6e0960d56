"""NumPy restatement of the reference's ``optimizer/vtrace.py`` -- TEST INFRASTRUCTURE ONLY.

Every function keeps the reference's name, argument order and layout and cites the lines it
follows.  The dtype of the computation is the dtype of the inputs (float64 for the truth
oracle, float32 to mimic the TF1 CPU kernels).  PINNED: ``tests/test_oracle_refexec.py`` executes the unmodified
``optimizer/vtrace.py`` over ``oracle/tf1_shim`` and every function here equals it to ~1e-12 (see ``oracle/__init__.py``).
"""
import numpy as np


def split_data(x):
    """optimizer/vtrace.py:3-14 -- first/middle/last windows shifted by 0/1/2 along axis 1."""
    return x[:, :-2], x[:, 1:-1], x[:, 2:]


def _one_hot(actions, depth, dtype):
    # tf.one_hot: out-of-range indices give an all-zero row.
    a = np.asarray(actions)
    return (a[..., None] == np.arange(depth)).astype(dtype)


def log_probs_from_softmax_and_actions(policy_softmax, actions, action_size):
    """optimizer/vtrace.py:16-27 -- log(sum_a p * onehot(a)); no epsilon."""
    p = np.asarray(policy_softmax)
    onehot = _one_hot(actions, action_size, p.dtype)
    selected = np.sum(p * onehot, axis=2)
    with np.errstate(divide="ignore"):
        return np.log(selected)


def from_importance_weights(log_rhos, discounts, rewards, values, bootstrap_value,
                            clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0):
    """optimizer/vtrace.py:71-103 -- time-major [T', B] inputs, bootstrap [B].

    ``clip_pg_rho_threshold`` is accepted and never used, exactly like the reference (:72).
    ``cs = min(1, rho)`` is hard-coded (:80).  Returns (vs, clipped_rhos), both [T', B].
    """
    log_rhos = np.asarray(log_rhos)
    dt = log_rhos.dtype
    discounts = np.asarray(discounts, dtype=dt)
    rewards = np.asarray(rewards, dtype=dt)
    values = np.asarray(values, dtype=dt)
    bootstrap_value = np.asarray(bootstrap_value, dtype=dt)

    rhos = np.exp(log_rhos)                                              # :74
    if clip_rho_threshold is not None:                                   # :75-78
        clipped_rhos = np.minimum(dt.type(clip_rho_threshold), rhos)
    else:
        clipped_rhos = rhos
    cs = np.minimum(dt.type(1.0), rhos)                                  # :80
    values_t_plus_1 = np.concatenate([values[1:], bootstrap_value[None]], axis=0)   # :81-82
    deltas = clipped_rhos * (rewards + discounts * values_t_plus_1 - values)        # :84

    acc = np.zeros_like(bootstrap_value)                                 # :92
    out = np.zeros_like(values)
    for t in range(values.shape[0] - 1, -1, -1):                         # :93-100 reverse scan
        acc = deltas[t] + discounts[t] * cs[t] * acc                     # :88-90
        out[t] = acc
    vs = out + values                                                    # :101
    return vs, clipped_rhos


def from_importance_weights_direct(log_rhos, discounts, rewards, values, bootstrap_value,
                                   clip_rho_threshold=1.0):
    """Independent second oracle: the published V-trace definition evaluated directly,
    vs_t - V_t = sum_{k>=t} (prod_{j=t}^{k-1} gamma_j c_j) * delta_k   (O(T^2) loops, float64).
    """
    lr = np.asarray(log_rhos, dtype=np.float64)
    g = np.asarray(discounts, dtype=np.float64)
    r = np.asarray(rewards, dtype=np.float64)
    v = np.asarray(values, dtype=np.float64)
    boot = np.asarray(bootstrap_value, dtype=np.float64)
    T = v.shape[0]
    rho = np.exp(lr)
    rho_bar = np.minimum(clip_rho_threshold, rho) if clip_rho_threshold is not None else rho
    c = np.minimum(1.0, rho)
    vs = np.zeros_like(v)
    for t in range(T):
        total = np.zeros_like(boot)
        for k in range(t, T):
            v_next = v[k + 1] if k + 1 < T else boot
            delta = rho_bar[k] * (r[k] + g[k] * v_next - v[k])
            coef = np.ones_like(boot)
            for j in range(t, k):
                coef = coef * g[j] * c[j]
            total = total + coef * delta
        vs[t] = v[t] + total
    return vs, rho_bar


def from_softmax(behavior_policy_softmax, target_policy_softmax, actions, discounts,
                 rewards, values, next_values, action_size, clip_rho_threshold=1.0,
                 clip_pg_rho_threshold=1.0):
    """optimizer/vtrace.py:29-69 -- batch-major [B, T', A] / [B, T'] in and out."""
    target_lp = log_probs_from_softmax_and_actions(target_policy_softmax, actions, action_size)
    behavior_lp = log_probs_from_softmax_and_actions(behavior_policy_softmax, actions, action_size)
    log_rhos = target_lp - behavior_lp                                   # :51
    dt = log_rhos.dtype
    t_vs, t_rho = from_importance_weights(
        log_rhos=log_rhos.T,
        discounts=np.asarray(discounts, dtype=dt).T,
        rewards=np.asarray(rewards, dtype=dt).T,
        values=np.asarray(values, dtype=dt).T,
        bootstrap_value=np.asarray(next_values, dtype=dt).T[-1],          # :62
        clip_rho_threshold=clip_rho_threshold,
        clip_pg_rho_threshold=clip_pg_rho_threshold)
    return t_vs.T, t_rho.T                                               # :66-67


def compute_policy_gradient_loss(softmax, actions, advantages, output_size):
    """optimizer/vtrace.py:105-112 -- -sum log(pi(a) + 1e-8) * adv."""
    p = np.asarray(softmax)
    onehot = _one_hot(actions, output_size, p.dtype)
    selected = np.sum(p * onehot, axis=2)
    return -np.sum(np.log(selected + p.dtype.type(1e-8)) * np.asarray(advantages, dtype=p.dtype))


def compute_baseline_loss(vs, value):
    """optimizer/vtrace.py:114-118 -- 0.5 * sum (vs - V)^2."""
    err = np.asarray(vs) - np.asarray(value)
    return np.sum(np.square(err)) * err.dtype.type(0.5)


def compute_entropy_loss(softmax):
    """optimizer/vtrace.py:120-126 -- sum pi*log(pi) (= -entropy); no epsilon."""
    p = np.asarray(softmax)
    with np.errstate(divide="ignore", invalid="ignore"):
        per = -p * np.log(p)
    return -np.sum(np.sum(per, axis=1))
