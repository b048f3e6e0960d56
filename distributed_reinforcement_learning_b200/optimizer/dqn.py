"""Stand-in for the reference's ``optimizer/dqn.py``."""
import numpy as np


def take_state_action_value(state_value, action, num_action):
    """optimizer/dqn.py:3-7: sum(state_value * one_hot(action), axis=1).  Inside the learner step this selection is
    part of the fused TD kernel (csrc/apex.cu::apex_td_kernel); this host version serves the stand-alone call."""
    sv = np.asarray(state_value)
    a = np.asarray(action).astype(np.int64)
    if sv.ndim != 2 or sv.shape[1] != num_action:
        raise ValueError("state_value must be [N, %d]" % num_action)
    onehot = (a[:, None] == np.arange(num_action)[None, :]).astype(sv.dtype)     # tf.one_hot: out-of-range -> zeros
    return np.sum(sv * onehot, axis=1)
