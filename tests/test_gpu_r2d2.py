"""GPU parity tests of the R2D2 learner: every call goes through the C-ABI
(``drl_r2d2_*``) and is compared with the float64 oracle to 1e-4 relative (tests/r2d2_parity.py)."""
import numpy as np
import pytest
import torch

import r2d2_parity as rp
from oracle import r2d2_torch as rt

pytestmark = pytest.mark.gpu


def _check(errs):
    bad = rp.r2d2_failures(errs)
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:10]


@pytest.mark.parametrize("mode", [1, 2])
def test_step_small(native, mode):
    _check(rp.compare_step(2, S=6, bi=2, math_mode=mode))


def test_step_reference_config(native):
    """config.json:84-101: batch 16, seq_len 15, burn_in 7, 84x84x1 frames, lstm 64, 4 actions."""
    _check(rp.compare_step(16, S=15, bi=7, A=4, C=1))


@pytest.mark.parametrize("B,S,bi,A,C", [(1, 2, 0, 2, 1), (3, 9, 0, 18, 4), (2, 40, 20, 6, 1), (5, 7, 5, 4, 4)])
def test_step_ragged(native, B, S, bi, A, C):
    _check(rp.compare_step(B, S=S, bi=bi, A=A, C=C))


def test_three_steps_with_target_sync(native):
    _check(rp.compare_step(3, S=8, bi=3, steps=3, sync_target_at=1))


def test_cuda_graph_path_matches(native):
    _check(rp.compare_step(3, S=8, bi=3, steps=3, sync_target_at=2, use_cuda_graph=True))


def test_td_error_act_and_q_tests(native):
    A, S, bi, B = 4, 8, 3, 4
    kwp = dict(num_action=A, lstm_size=64, input_shape=(84, 84, 1))
    params = rt.init_params(0, torch.float32, **kwp)
    target = rt.init_params(1, torch.float32, **kwp)
    L = rt.Learner(params, target, torch.float64, seq_len=S, burn_in=bi, **kwp)
    eng = rp.native_r2d2(B, S, bi, A, 1, params, target)
    try:
        b = rt.make_sequences(3, S=S, A=A, seed=77)              # n < batch sequences at once
        td = eng.td_error(b["state"], b["previous_action"], b["action"], b["h"][:, 0], b["c"][:, 0], b["reward"], b["done"])
        ref = [L.get_td_error(*[b[k][i] for k in rt.TRAIN_FIELDS[:-1]]) for i in range(3)]
        assert rp.rel_err(td, np.asarray(ref)) < rp.TOL
        t = eng.taps(3)
        o = L.losses(*[b[k] for k in rt.TRAIN_FIELDS[:-1]])
        for k in ("main_q", "target_q", "target_value", "state_action_value"):
            assert rp.rel_err(t[k], o[k].detach().numpy()) < rp.TOL, k
        # get_action: one network step from a stored state
        q, h2, c2 = eng.act(b["state"][:, 0], b["previous_action"][:, 0], b["h"][:, 0], b["c"][:, 0])
        rq, rh, rc = L.step_q(b["state"][:, 0], b["h"][:, 0], b["c"][:, 0], b["previous_action"][:, 0])
        assert rp.rel_err(q, rq) < rp.TOL and rp.rel_err(h2, rh) < rp.TOL and rp.rel_err(c2, rc) < rp.TOL
        assert rp.rel_err(q, o["main_q"].detach().numpy()[:, 0]) < rp.TOL       # first step of the unroll
    finally:
        eng.close()


def test_agent_surface_runs_learner_loop(native):
    """The learner branch of train_r2d2.py:87-165 on the stand-in modules."""
    from distributed_reinforcement_learning_b200.agent import r2d2
    from distributed_reinforcement_learning_b200.distributed_queue import buffer_queue
    data = dict(seq_len=6, burn_in=2, model_input=[84, 84, 1], model_output=4, lstm_size=64, discount_factor=0.997,
                start_learning_rate=1e-4, end_learning_rate=0.0, learning_frame=1000000000, gradient_clip_norm=40.0,
                batch_size=4)
    learner = r2d2.Agent(seq_len=data['seq_len'], burn_in=data['burn_in'], input_shape=data['model_input'],
                         num_action=data['model_output'], lstm_size=data['lstm_size'],
                         discount_factor=data['discount_factor'], start_learning_rate=data['start_learning_rate'],
                         end_learning_rate=data['end_learning_rate'], learning_frame=data['learning_frame'],
                         gradient_clip_norm=data['gradient_clip_norm'], model_name='learner', learner_name='learner')
    learner.set_session(None)
    learner.main_to_target()
    per = buffer_queue.Memory(capacity=64)
    q = rt.make_sequences(8, S=6, seed=5)
    for i in range(8):
        td = learner.get_td_error(state=q['state'][i], previous_action=q['previous_action'][i], action=q['action'][i],
                                  reward=q['reward'][i], done=q['done'][i], h=q['h'][i], c=q['c'][i])
        assert np.isfinite(td) and td >= 0
        per.add(td, [q['state'][i], q['previous_action'][i], q['action'][i], q['reward'][i], q['done'][i], q['h'][i],
                     q['c'][i]])
    for step in range(3):
        minibatch, idxs, weight = per.sample(data['batch_size'])
        cols = list(zip(*minibatch))
        loss, td_error = learner.train(state=list(cols[0]), previous_action=list(cols[1]), action=list(cols[2]),
                                       h=list(cols[5]), c=list(cols[6]), reward=list(cols[3]), done=list(cols[4]),
                                       weight=weight)
        assert np.isfinite(loss) and td_error.shape == (4,)
        for i in range(len(idxs)):
            per.update(idxs[i], td_error[i])
    action, qa, h2, c2 = learner.get_action(q['state'][0][0], q['h'][0][0], q['c'][0][0], 0, epsilon=0.0)
    assert h2.shape == (64,) and c2.shape == (64,) and np.isfinite(qa)
    mq = learner.main_q_value_test(q['state'][0], q['h'][0], q['c'][0], q['done'][0], q['previous_action'][0])
    assert mq.shape == (6, 4)


def test_errors_are_loud(native):
    from distributed_reinforcement_learning_b200.r2d2_learner import NativeR2D2Learner
    with pytest.raises(native.DrlError):
        NativeR2D2Learner(batch=2, lstm_size=128)
    with pytest.raises(native.DrlError):
        NativeR2D2Learner(batch=2, seq_len=5, burn_in=4)
    with pytest.raises(native.DrlError):
        NativeR2D2Learner(batch=2, input_shape=(84, 84, 3))
    eng = NativeR2D2Learner(batch=2, seq_len=4, burn_in=1)
    try:
        with pytest.raises(native.DrlError):
            eng.step(0)
        with pytest.raises(native.DrlError):
            eng.taps()
    finally:
        eng.close()
