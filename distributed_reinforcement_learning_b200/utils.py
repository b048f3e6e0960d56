"""Learner-side helpers of the reference's ``utils.py`` that ``train_impala.py`` touches.

  check_properties   utils.py:34-45  (config validation before the learner is built)
  copy_src_to_dst    utils.py:6-22   (learner -> actor variable copy; here a callable shim)

  UnrolledTrajectory utils.py:80-119  (the actor-side accumulator whose ``extract()`` feeds ``FIFOQueue.append_to_queue``,
                                       train_impala.py:117,165-189: closes the actor -> ring -> learner loop)
"""
import collections


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


class _Accumulator:
    """Per-actor accumulator of one unroll: ``initialize()`` empties it, ``append(...)`` adds one step."""
    _FIELDS = ()

    def __init__(self):
        self.trajectory_data = collections.namedtuple('trajectory_data', list(self._FIELDS))

    def initialize(self):
        self.unroll_data = self.trajectory_data(*[[] for _ in self._FIELDS])

    def _add(self, values):
        for field, value in zip(self.unroll_data, values):
            field.append(value)


class UnrolledA3CTrajectory(_Accumulator):
    """utils.py:47-78: the A3C actor's unroll; ``extract()`` stacks every field (train_a3c.py feeds it to A3CFIFOQueue)."""
    _FIELDS = ('state', 'next_state', 'reward', 'done', 'action', 'previous_action')

    def append(self, state, next_state, previous_action, action, reward, done):
        self._add((state, next_state, reward, done, action, previous_action))

    def extract(self):
        import numpy as np
        return {k: np.stack(v) for k, v in self.unroll_data._asdict().items()}


class UnrolledTrajectory:
    """utils.py:80-119: per-actor accumulator of one unroll; ``extract()`` returns the nine per-step lists that
    ``FIFOQueue.append_to_queue`` takes (train_impala.py:178-189)."""

    def __init__(self):
        self.trajectory_data = collections.namedtuple(
            'trajectory_data',
            ['state', 'next_state', 'reward', 'done', 'action', 'behavior_policy', 'previous_action',
             'initial_h', 'initial_c'])

    def initialize(self):
        self.unroll_data = self.trajectory_data(*[[] for _ in range(9)])

    def append(self, state, next_state, reward, done, action, behavior_policy, previous_action, initial_h, initial_c):
        for field, value in zip(self.unroll_data, (state, next_state, reward, done, action, behavior_policy,
                                                   previous_action, initial_h, initial_c)):
            field.append(value)

    def extract(self):
        return dict(self.unroll_data._asdict())
