"""Learner-side helpers of the reference's ``utils.py`` that ``train_impala.py`` touches.

  check_properties   utils.py:34-45  (config validation before the learner is built)
  copy_src_to_dst    utils.py:6-22   (learner -> actor variable copy; here a callable shim)

The actor-side trajectory accumulators (utils.py:47-119) are out of scope.
"""


def check_properties(data):
    """Same five checks as the reference, raising AssertionError on a bad config."""
    n_actors = data['num_actors']
    avail, envs = data['available_action'], data['env']
    assert all(data['model_output'] >= a for a in avail), "available_action exceeds model_output"
    assert n_actors == len(avail), "num_actors != len(available_action)"
    assert n_actors == len(envs), "num_actors != len(env)"
    assert len(avail) == len(envs), "len(available_action) != len(env)"
    assert data['reward_clipping'] in ['abs_one', 'soft_asymmetric'], "unknown reward_clipping"


def copy_src_to_dst(from_scope, to_scope):
    """Returns a zero-argument callable that copies the variables of agent `from_scope` into agent
    `to_scope` (what running the reference's assign ops does)."""
    from .agent import impala

    def run():
        dst = impala._AGENTS.get(to_scope)
        if dst is not None:
            dst.parameter_sync()
    return run
