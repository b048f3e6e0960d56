"""B200-native stand-in for the reference's ``agent/r2d2.py``: same constructor kwargs, methods and return values, so
the learner branch of ``train_r2d2.py:87-165`` runs on it unchanged, with the TF1 graph replaced by ``drl_r2d2_*``.

  Agent.__init__       agent/r2d2.py:13-95
  Agent.get_td_error   agent/r2d2.py:97-130  (ONE sequence -> scalar)
  Agent.train          agent/r2d2.py:132-159 -> (loss, td_error [B])
  Agent.parameter_sync agent/r2d2.py:161-162 ; Agent.main_to_target :164-165 ; Agent.set_session :167-169
  Agent.get_action     agent/r2d2.py:171-191 -> (action, q_value[action], h', c')
  Agent.main_q_value_test / target_q_value_test   agent/r2d2.py:193-217
"""
import os
import threading

import numpy as np

from ..model import r2d2_lstm
from ..r2d2_learner import MAIN, TARGET, NativeR2D2Learner

_AGENTS = {}


class Agent:

    def __init__(self, seq_len, burn_in, input_shape, num_action, lstm_size,
                 discount_factor, start_learning_rate, end_learning_rate,
                 learning_frame, gradient_clip_norm, model_name, learner_name):
        self.seq_len = seq_len
        self.burn_in = burn_in
        self.input_shape = list(input_shape)
        self.num_action = num_action
        self.lstm_size = lstm_size
        self.discount_factor = discount_factor
        self.start_learning_rate = start_learning_rate     # stored like the reference; its optimizer is a constant
        self.end_learning_rate = end_learning_rate         # AdamOptimizer(1e-4) without clipping (agent/r2d2.py:91-92)
        self.learning_frame = learning_frame
        self.model_name, self.learner_name = model_name, learner_name
        self.device = int(os.environ.get("LOCAL_RANK", "0"))
        self.use_cuda_graph = os.environ.get("DRL_B200_CUDA_GRAPH", "0") == "1"
        self.sess = None
        self._kw = dict(num_action=num_action, lstm_size=lstm_size, input_shape=tuple(input_shape))
        self._main = self._target = self._opt = None
        self._engine = None
        self._slot = 0
        self._last = {}
        self._lock = threading.RLock()   # actor threads sync from / act on this agent while the learner thread trains
        _AGENTS[model_name] = self

    def _ensure_init(self):
        if self._main is None:
            self._main = r2d2_lstm.init_params(**self._kw)
            self._target = r2d2_lstm.init_params(**self._kw)
            z = np.zeros_like(self._main)
            self._opt = dict(m=z, v=z.copy(), step=0, beta1_power=0.9, beta2_power=0.999)

    def _pull_state(self):
        if self._engine is not None:
            self._main, self._target = self._engine.get_params(MAIN), self._engine.get_params(TARGET)
            self._opt = self._engine.get_opt_state()

    def _push_state(self):
        e, o = self._engine, self._opt
        e.set_params(self._main, MAIN)
        e.set_params(self._target, TARGET)
        e.set_opt_state(o["m"], o["v"], o["step"], o["beta1_power"], o["beta2_power"])

    def _get_engine(self, batch):
        self._ensure_init()
        if self._engine is None or self._engine.B < batch:
            if self._engine is not None:
                self._pull_state()
                self._engine.close()
            self._engine = NativeR2D2Learner(
                batch=batch, seq_len=self.seq_len, burn_in=self.burn_in, num_action=self.num_action,
                lstm_size=self.lstm_size, input_shape=tuple(self.input_shape), discount_factor=self.discount_factor,
                device=self.device, num_slots=2, use_cuda_graph=self.use_cuda_graph)
            self._push_state()
        return self._engine

    @staticmethod
    def _u8(state):
        st = np.asarray(state)
        if st.dtype != np.uint8:
            raise TypeError("states must be uint8 frames (the /255 normalisation runs on the GPU)")
        return st

    def get_td_error(self, state, previous_action, action, h, c, reward, done):
        eng = self._get_engine(1)
        td = eng.td_error(self._u8(state)[None], np.asarray(previous_action)[None], np.asarray(action)[None],
                          np.asarray(h, np.float32)[None, 0], np.asarray(c, np.float32)[None, 0],
                          np.asarray(reward, np.float32)[None], np.asarray(done)[None])
        return float(td[0])

    def train(self, state, previous_action, action, h, c, reward, done, weight):
        st = self._u8(np.stack(state))
        if self._engine is not None and self._engine.B != st.shape[0]:
            self._pull_state()
            self._engine.close()
            self._engine = None
        eng = self._get_engine(st.shape[0])
        slot = self._slot
        self._slot = (self._slot + 1) % eng.num_slots
        eng.stage(slot, st, np.stack(previous_action), np.stack(action), np.stack(h)[:, 0], np.stack(c)[:, 0],
                  np.stack(reward), np.stack(done), weight)
        out, td = eng.step(slot)
        self._last = out
        return out["loss"], td

    def _snapshot_params(self):
        with self._lock:
            self._ensure_init()
            if self._engine is not None:
                return self._engine.get_params(MAIN), self._engine.get_params(TARGET)
            return self._main.copy(), self._target.copy()

    def parameter_sync(self):
        src = _AGENTS.get(self.learner_name)
        if src is None or src is self:
            return
        main, target = src._snapshot_params()            # trainable variables only, into caller-owned buffers
        with self._lock:
            self._ensure_init()
            self._main, self._target = main, target
            if self._engine is not None:
                self._engine.set_params(self._main, MAIN)
                self._engine.set_params(self._target, TARGET)

    def main_to_target(self):
        self._ensure_init()
        if self._engine is not None:
            self._engine.main_to_target()
        else:
            self._target = self._main.copy()

    def set_session(self, sess):
        self.sess = sess
        self._main = None
        self._ensure_init()
        if self._engine is not None:
            self._push_state()

    def get_action(self, state, h, c, previous_action, epsilon):
        eng = self._engine if self._engine is not None else self._get_engine(1)
        q, h2, c2 = eng.act(self._u8(state)[None], np.asarray([previous_action], np.int32),
                            np.asarray(h, np.float32)[None], np.asarray(c, np.float32)[None])
        q = q[0]
        if np.random.rand() > epsilon:
            action = np.argmax(q)
        else:
            action = np.random.choice(self.num_action)
        return action, q[action], h2[0], c2[0]

    def _unroll_q(self, key, state, h, c, done, previous_action):
        eng = self._get_engine(1)
        S = self.seq_len
        eng.td_error(self._u8(np.stack(state))[None], np.asarray(previous_action)[None], np.zeros((1, S), np.int32),
                     np.asarray(h, np.float32)[None, 0], np.asarray(c, np.float32)[None, 0], np.zeros((1, S), np.float32),
                     np.asarray(done)[None])
        return eng.taps(1)[key][0]

    def main_q_value_test(self, state, h, c, done, previous_action):
        return self._unroll_q("main_q", state, h, c, done, previous_action)

    def target_q_value_test(self, state, h, c, done, previous_action):
        return self._unroll_q("target_q", state, h, c, done, previous_action)

    value_loss = property(lambda self: self._last.get("loss"))
    grad_norm = property(lambda self: self._last.get("grad_norm"))


def _locked(fn):
    def wrapper(self, *a, **k):
        with self._lock:
            return fn(self, *a, **k)
    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


# actor threads call these on a shared agent while the learner thread trains (in-process launch mode)
for _n in ('save_weights', 'load_weights', 'set_session', 'main_to_target', 'get_action', 'get_td_error', 'train', 'main_q_value_test', 'target_q_value_test'):
    if hasattr(Agent, _n):
        setattr(Agent, _n, _locked(getattr(Agent, _n)))
