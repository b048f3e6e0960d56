"""Stand-in for the reference's ``optimizer/a2c.py`` (host NumPy versions of the stand-alone functions; inside the
learner step they are one fused kernel, csrc/apex.cu::a3c_loss_kernel)."""
import numpy as np


def compute_entropy_loss(policy):
    """optimizer/a2c.py:3-7: -mean_b(sum_a -pi log pi)  (no epsilon, like the reference)."""
    p = np.asarray(policy, np.float64)
    return float(-np.mean(np.sum(-p * np.log(p), axis=1)))


def _td(value, next_value, discounts, reward):
    return np.asarray(reward, np.float64) + np.asarray(discounts, np.float64) * np.asarray(next_value, np.float64) - \
        np.asarray(value, np.float64)


def compute_baseline_loss(value, next_value, discounts, reward):
    """optimizer/a2c.py:9-15."""
    d = _td(value, next_value, discounts, reward)
    return float(np.mean(d * d))


def compute_policy_loss(policy, action, value, next_value, discounts, reward, num_action):
    """optimizer/a2c.py:17-26: -mean(advantage * pi(a)) -- the probability, not its logarithm."""
    p = np.asarray(policy, np.float64)
    sel = p[np.arange(p.shape[0]), np.asarray(action).astype(np.int64)]
    return float(-np.mean(_td(value, next_value, discounts, reward) * sel))
