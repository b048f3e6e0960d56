"""B200-native stand-ins for ``distributed_queue/buffer_queue.py``: ``FIFOQueue`` (reference :418-512, the IMPALA
trajectory queue) and ``SumTree`` / ``Memory`` (reference :326-416, the Ape-X prioritized replay memory).

The reference keeps a ``tf.FIFOQueue`` of 9-field trajectories on the learner's CPU and dequeues a
batch with ``batch_size`` serial ``sess.run`` RPCs, after which the launcher ``np.stack``s eight
fields (train_impala.py:98-108).  Here the queue is a pinned-host ring (``drl_ring_*``) organised
as batch slots: a popped batch is already eight contiguous, page-locked ``[B, ...]`` arrays that
``Agent.train`` hands to ``cudaMemcpyAsync`` without another host copy.

Same constructor arguments, method names and ``batch_tuple`` field order as the reference.
``next_state`` is accepted by ``append_to_queue`` and dropped: the learner never reads it
(train_impala.py:100-108); ``batch.next_state`` is returned as ``None``.
"""
import collections
import ctypes as C

import numpy as np

from .. import _native as N

batch_tuple = collections.namedtuple(
    'batch_tuple', ['state', 'next_state', 'reward', 'done', 'behavior_policy', 'action',
                    'previous_action', 'previous_h', 'previous_c'])


def _view(addr, shape, dtype):
    n = int(np.prod(shape))
    buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


class FIFOQueue:

    def __init__(self, trajectory, input_shape, output_size,
                 queue_size, batch_size, num_actors, lstm_size, pinned=True):
        self.trajectory = trajectory
        self.input_shape = list(input_shape)
        self.output_size = output_size
        self.batch_size = batch_size
        self.lstm_size = lstm_size
        self.queue_size = queue_size
        self.num_actors = num_actors
        self.sess = None
        h, w, c = self.input_shape
        self._r = C.c_void_p()
        N.check(N.lib.drl_ring_create(trajectory, h, w, c, output_size, lstm_size, max(queue_size, batch_size),
                                      batch_size, 1 if pinned else 0, C.byref(self._r)))
        self.pinned = bool(N.lib.drl_ring_is_pinned(self._r))
        self._held = None

    def close(self):
        if getattr(self, "_r", None) is not None and self._r.value:
            N.lib.drl_ring_destroy(self._r)
            self._r = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def append_to_queue(self, task, unrolled_state, unrolled_next_state,
                        unrolled_reward, unrolled_done, unrolled_behavior_policy,
                        unrolled_action, unrolled_previous_action,
                        unrolled_previous_h, unrolled_previous_c, timeout_ms=-1):
        """buffer_queue.py:468-484.  Blocks while the queue is full (like a full tf.FIFOQueue)."""
        T, A, L = self.trajectory, self.output_size, self.lstm_size
        st = N.as_c(unrolled_state, np.uint8, (T, *self.input_shape), "unrolled_state")
        rw = N.as_c(unrolled_reward, np.float32, (T,), "unrolled_reward")
        dn = N.as_c(np.asarray(unrolled_done).astype(np.uint8), np.uint8, (T,), "unrolled_done")
        mu = N.as_c(unrolled_behavior_policy, np.float32, (T, A), "unrolled_behavior_policy")
        ac = N.as_c(unrolled_action, np.int32, (T,), "unrolled_action")
        pa = N.as_c(unrolled_previous_action, np.int32, (T,), "unrolled_previous_action")
        ph = N.as_c(unrolled_previous_h, np.float32, (T, L), "unrolled_previous_h")
        pc = N.as_c(unrolled_previous_c, np.float32, (T, L), "unrolled_previous_c")
        N.check(N.lib.drl_ring_push(self._r, N.ptr(st), N.ptr(rw), N.ptr(dn), N.ptr(mu), N.ptr(ac), N.ptr(pa),
                                    N.ptr(ph), N.ptr(pc), int(timeout_ms)))

    def sample_batch(self, timeout_ms=-1):
        """buffer_queue.py:486-505: the oldest ``batch_size`` trajectories, FIFO order.  The returned
        arrays are views into the ring's pinned memory and stay valid until the next ``sample_batch``."""
        if self._held is not None:
            N.check(N.lib.drl_ring_release(self._r, self._held))
            self._held = None
        rb = N.RingBatch()
        N.check(N.lib.drl_ring_pop_batch(self._r, C.byref(rb), int(timeout_ms)))
        self._held = int(rb.slot)
        B, T, A, L = self.batch_size, self.trajectory, self.output_size, self.lstm_size
        return batch_tuple(
            _view(rb.state, (B, T, *self.input_shape), np.uint8),
            None,
            _view(rb.reward, (B, T), np.float32),
            _view(rb.done, (B, T), np.uint8).view(np.bool_),
            _view(rb.behavior_policy, (B, T, A), np.float32),
            _view(rb.action, (B, T), np.int32),
            _view(rb.previous_action, (B, T), np.int32),
            _view(rb.previous_h, (B, T, L), np.float32),
            _view(rb.previous_c, (B, T, L), np.float32))

    def get_size(self):
        """buffer_queue.py:507-509."""
        return int(N.lib.drl_ring_size(self._r))

    def set_session(self, sess):
        """buffer_queue.py:511-512 (kept for call compatibility; there is no session)."""
        self.sess = sess


class SumTree:
    """buffer_queue.py:326-369 over the native float64 sum tree (``drl_per_*``): same attributes and methods; the
    stored objects stay in a host list, the priorities live in the native tree."""

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.data = np.zeros(self.capacity, dtype=object)
        self._p = C.c_void_p()
        N.check(N.lib.drl_per_create(self.capacity, C.byref(self._p)))

    def close(self):
        if getattr(self, "_p", None) is not None and self._p.value:
            N.lib.drl_per_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_entries(self):
        n = C.c_int64()
        N.check(N.lib.drl_per_size(self._p, C.byref(n)))
        return int(n.value)

    def total(self):
        t = C.c_double()
        N.check(N.lib.drl_per_total(self._p, C.byref(t)))
        return float(t.value)


class Memory(object):
    """buffer_queue.py:371-415: prioritized replay with p = (|td| + 0.001) ** 0.6, stratified sampling and importance
    weights (n_entries * p / total) ** -beta / max, beta 0.4 -> 1 by +0.001 per ``sample`` call."""
    e = 0.001
    a = 0.6
    beta_increment_per_sampling = 0.001

    def __init__(self, capacity):
        self.capacity = capacity
        self.tree = SumTree(capacity)

    def reset(self):
        self.tree.close()
        self.tree = SumTree(self.capacity)

    @property
    def beta(self):
        b = C.c_double()
        N.check(N.lib.drl_per_beta(self.tree._p, C.byref(b)))
        return float(b.value)

    def _getPriority(self, error):
        return (error + self.e) ** self.a

    def add(self, error, sample):
        idx = C.c_int64()
        N.check(N.lib.drl_per_add(self.tree._p, float(error), C.byref(idx)))
        self.tree.data[idx.value] = sample

    def sample(self, n, u01=None):
        """-> (batch, idxs, is_weight) like the reference; ``u01`` ([n] uniforms in [0, 1)) makes the draw
        reproducible, by default they come from ``random.random`` as in the reference."""
        if u01 is None:
            import random
            u01 = [random.random() for _ in range(n)]
        u = N.as_c(u01, np.float64, (n,), "u01")
        ti = np.empty(n, np.int64)
        di = np.empty(n, np.int64)
        pr = np.empty(n, np.float64)
        w = np.empty(n, np.float64)
        N.check(N.lib.drl_per_sample(self.tree._p, n, N.ptr(u), N.ptr(ti), N.ptr(di), N.ptr(pr), N.ptr(w)))
        batch = [self.tree.data[i] for i in di]
        self.last_priorities = pr
        return batch, [int(i) for i in ti], w

    def update(self, idx, error):
        N.check(N.lib.drl_per_update(self.tree._p, int(idx), float(error)))
