"""Stand-in for the reference's ``optimizer/burn_in.py`` (host NumPy versions of the stand-alone functions; inside
the learner step they are part of the fused TD kernel, csrc/r2d2.cu::r2d2_td_kernel, evaluated in float64)."""
import numpy as np


def slice_in_burnin(size, tensor):
    """optimizer/burn_in.py:3-4."""
    return np.asarray(tensor)[:, size:]


def reformat_tensor(main_q_value, target_q_value, reward, action, done, mask):
    """optimizer/burn_in.py:6-15."""
    return (main_q_value[:, :-1], main_q_value[:, 1:], target_q_value[:, 1:], reward[:, :-1], action[:, :-1],
            done[:, :-1], mask[:, :-1])


def select_state_value_action(q_value, action, num_action):
    """optimizer/burn_in.py:17-21."""
    q = np.asarray(q_value)
    onehot = (np.asarray(action)[..., None] == np.arange(num_action)).astype(q.dtype)
    return np.sum(q * onehot, axis=2)


def value_function_rescaling(x, eps):
    """optimizer/burn_in.py:23-25 (R2D2 paper, table 2)."""
    x = np.asarray(x, np.float64)
    return np.sign(x) * (np.sqrt(np.abs(x) + 1.) - 1.) + eps * x


def inverse_value_function_rescaling(x, eps):
    """optimizer/burn_in.py:27-32."""
    x = np.asarray(x, np.float64)
    return np.sign(x) * (np.square(((np.sqrt(1. + 4. * eps * (np.abs(x) + 1. + eps))) - 1.) / (2. * eps)) - 1.)
