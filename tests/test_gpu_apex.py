"""GPU parity tests of the Ape-X learner: every call goes through the C-ABI
(``drl_apex_*``) and is compared with the float64 oracle to 1e-4 relative (tests/apex_parity.py)."""
import numpy as np
import pytest
import torch

import apex_parity as ap
from oracle import apex_torch as ax

pytestmark = pytest.mark.gpu


def _check(errs):
    bad = ap.apex_failures(errs)
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:10]


@pytest.mark.parametrize("mode", [1, 2])
def test_step_small(native, mode):
    _check(ap.compare_step(3, A=4, math_mode=mode))


def test_step_reference_config(native):
    """config.json:146-184: batch 32, 4 actions."""
    _check(ap.compare_step(32, A=4))


@pytest.mark.parametrize("B,A", [(1, 2), (5, 18), (16, 7)])
def test_step_ragged(native, B, A):
    _check(ap.compare_step(B, A=A))


def test_three_steps_with_target_sync(native):
    _check(ap.compare_step(4, A=4, steps=3, sync_target_at=1))


def test_no_reward_clipping(native):
    _check(ap.compare_step(4, A=4, reward_clipping="none"))


def test_cuda_graph_path_matches(native):
    _check(ap.compare_step(4, A=4, steps=3, use_cuda_graph=True, sync_target_at=2))


def test_td_error_act_and_unit_weights(native):
    A, B = 4, 6
    params = ax.init_params(0, torch.float32, num_action=A)
    target = ax.init_params(1, torch.float32, num_action=A)
    L = ax.Learner(params, target, torch.float64, num_action=A)
    eng = ap.native_apex(B, A, params, target)
    try:
        b = ax.make_transitions(4, A=A, seed=99)          # n < batch
        args = [b[k] for k in ax.TRAIN_FIELDS[:-1]]
        td = eng.td_error(*args)
        assert ap.rel_err(td, L.get_td_error(*args)) < ap.TOL
        t = eng.taps(4)
        o = L.losses(*args)
        for k in ("main_q", "next_main_q", "target_q", "target_value"):
            assert ap.rel_err(t[k], o[k].detach().numpy()) < ap.TOL, k
        q = eng.act(np.concatenate([b["state"], b["next_state"]]), np.concatenate([b["previous_action"], b["action"]]))
        ref = np.concatenate([o["main_q"].detach().numpy(), o["next_main_q"].numpy()])
        assert ap.rel_err(q, ref) < ap.TOL
        # Agent.train = unit importance weights (agent/apex.py:167)
        b6 = ax.make_transitions(B, A=A, seed=5)
        a6 = [b6[k] for k in ax.TRAIN_FIELDS[:-1]]
        eng.stage(0, *a6, None)
        out, td6 = eng.step(0)
        (loss, tdo), *_ = L.distributed_train(*a6, np.ones(B, np.float32), return_all=True)
        assert ap.rel_err(out["loss"], loss) < ap.TOL and ap.rel_err(td6, tdo) < ap.TOL
    finally:
        eng.close()


def test_agent_surface_runs_learner_loop(native):
    """The learner branch of train_apex.py:82-155 on the stand-in modules: get_td_error -> Memory.add,
    Memory.sample -> distributed_train -> Memory.update, target_to_main every few steps."""
    from distributed_reinforcement_learning_b200.agent import apex
    from distributed_reinforcement_learning_b200.distributed_queue import buffer_queue
    data = dict(model_input=[84, 84, 4], model_output=4, discount_factor=0.99, gradient_clip_norm=40.0,
                reward_clipping="abs_one", start_learning_rate=1e-4, end_learning_rate=0.0,
                learning_frame=100000000000000, batch_size=8, trajectory=8)
    learner = apex.Agent(input_shape=data['model_input'], num_action=data['model_output'],
                         discount_factor=data['discount_factor'], gradient_clip_norm=data['gradient_clip_norm'],
                         reward_clipping=data['reward_clipping'], start_learning_rate=data['start_learning_rate'],
                         end_learning_rate=data['end_learning_rate'], learning_frame=data['learning_frame'],
                         model_name='learner', learner_name='learner')
    learner.set_session(None)
    learner.target_to_main()
    replay_buffer = buffer_queue.Memory(capacity=64)
    losses = []
    for it_ in range(4):
        fa = ax.make_transitions(data['trajectory'], seed=100 + it_)
        td_error = learner.get_td_error(state=fa['state'], next_state=fa['next_state'],
                                        previous_action=fa['previous_action'], action=fa['action'],
                                        reward=fa['reward'], done=fa['done'])
        assert td_error.shape == (8,) and np.all(np.isfinite(td_error)) and np.all(td_error >= 0)
        for i in range(len(td_error)):
            replay_buffer.add(td_error[i], [fa['state'][i], fa['next_state'][i], fa['previous_action'][i],
                                            fa['action'][i], fa['reward'][i], fa['done'][i]])
        minibatch, idxs, is_weight = replay_buffer.sample(data['batch_size'])
        minibatch = np.array(minibatch, dtype=object)
        loss, td = learner.distributed_train(
            state=np.stack(minibatch[:, 0]), next_state=np.stack(minibatch[:, 1]),
            previous_action=np.stack(minibatch[:, 2]), action=np.stack(minibatch[:, 3]),
            reward=np.stack(minibatch[:, 4]), done=np.stack(minibatch[:, 5]), is_weight=is_weight)
        assert np.isfinite(loss) and td.shape == (8,)
        losses.append(loss)
        if it_ % 2 == 0:
            learner.target_to_main()
        for i in range(len(idxs)):
            replay_buffer.update(idxs[i], td[i])
    assert learner.num_env_frames == 4
    action, q, qa = learner.get_policy_and_action(fa['state'][0], 0, epsilon=0.0)
    assert q.shape == (4,) and qa == q[action] and action == int(np.argmax(q))


def test_errors_are_loud(native):
    from distributed_reinforcement_learning_b200.apex_learner import NativeApexLearner
    with pytest.raises(native.DrlError):
        NativeApexLearner(batch=2, num_action=64)
    with pytest.raises(native.DrlError):
        NativeApexLearner(batch=2, input_shape=(64, 64, 4))
    eng = NativeApexLearner(batch=2, num_action=4)
    try:
        with pytest.raises(native.DrlError):
            eng.step(0)                                   # not staged
        with pytest.raises(ValueError):
            eng.stage(0, np.zeros((3, 84, 84, 4), np.uint8), np.zeros((3, 84, 84, 4), np.uint8), [0] * 3, [0] * 3,
                      [0.0] * 3, [0] * 3)
        with pytest.raises(native.DrlError):
            eng.taps()
    finally:
        eng.close()
