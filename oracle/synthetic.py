"""Seeded synthetic trajectories -- TEST INFRASTRUCTURE ONLY.

Field order, dtypes and shapes follow the learner call at train_impala.py:100-108 and the
queue placeholders at distributed_queue/buffer_queue.py:427-435 (batch-major [B, T, ...]).
"""
import numpy as np


def make_batch(B, T=20, A=18, L=256, input_shape=(84, 84, 4), seed=1234, rank=0):
    rng = np.random.default_rng(seed + rank)
    state = rng.integers(0, 256, (B, T, *input_shape), dtype=np.uint8)
    previous_action = rng.integers(0, A, (B, T)).astype(np.int32)
    action = rng.integers(0, A, (B, T)).astype(np.int32)
    logits = rng.standard_normal((B, T, A)).astype(np.float32)
    e = np.exp(logits - logits.max(axis=-1, keepdims=True))
    behavior_policy = (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)
    reward = rng.standard_normal((B, T)).astype(np.float32)
    reward[rng.random((B, T)) < 0.10] = 0.0                       # 10 % exact zeros
    big = rng.random((B, T)) < 0.05                               # a few |r| > 1 (exercise the clip)
    reward[big] = (reward[big] * 4.0).astype(np.float32)
    done = rng.random((B, T)) < 0.05
    initial_h = np.clip(rng.standard_normal((B, T, L)) * 0.5, -0.999, 0.999).astype(np.float32)
    initial_c = rng.standard_normal((B, T, L)).astype(np.float32)
    return dict(state=state, reward=reward, action=action, done=done,
                behavior_policy=behavior_policy, previous_action=previous_action,
                initial_h=initial_h, initial_c=initial_c)


TRAIN_FIELDS = ("state", "reward", "action", "done", "behavior_policy", "previous_action",
                "initial_h", "initial_c")


def slice_batch(batch, lo, hi):
    return {k: v[lo:hi] for k, v in batch.items()}
