"""CPU tests of the Ape-X oracle (oracle/apex_torch.py, oracle/per_np.py), the native prioritized-replay index
(drl_per_*, host code) against it, and the host-side Ape-X mirror modules.  PARITY UNPINNED: the reference ships no
tests; the oracle is pinned by hand-worked known answers and invariants."""
import numpy as np
import pytest
import torch

from oracle import apex_torch as ax
from oracle import per_np


def test_param_inventory():
    # conv 8192+32 + 32768+64 + 36864+64, embedding 4*256+256 + 65536+256, two streams of 3392*256+256 + 65536+256,
    # outputs 256*4+4 and 256+1
    n = 8224 + 32832 + 36928 + 1280 + 65792 + 2 * (868608 + 65792) + 1028 + 257
    assert ax.param_count(num_action=4) == n == 2015141
    from distributed_reinforcement_learning_b200.model import apex_value
    assert apex_value.param_count(num_action=4) == n
    assert [s for s in apex_value.param_specs(num_action=6)] == [s for s in ax.param_specs(num_action=6)]


def _tiny_learner(dtype=torch.float64, **cfg):
    return ax.Learner(dtype=dtype, **cfg)


def test_td_target_known_answer():
    """Hand-worked double-DQN target (agent/apex.py:56-61): next_action comes from MAIN(s'), its value from TARGET(s')."""
    L = _tiny_learner()
    b = ax.make_transitions(3, seed=7)
    o = L.losses(b["state"], b["next_state"], b["previous_action"], b["action"], b["reward"], b["done"], b["is_weight"])
    nq, tq, mq = o["next_main_q"].numpy(), o["target_q"].numpy(), o["main_q"].detach().numpy()
    for i in range(3):
        na = int(np.argmax(nq[i]))
        r = float(np.clip(b["reward"][i], -1, 1))
        disc = 0.0 if b["done"][i] else 0.99
        assert o["target_value"][i].item() == pytest.approx(tq[i, na] * disc + r, rel=1e-12)
        assert o["state_action_value"][i].item() == pytest.approx(mq[i, b["action"][i]], rel=1e-12)
    td2 = (o["target_value"].numpy() - o["state_action_value"].detach().numpy()) ** 2
    assert o["value_loss"].item() == pytest.approx(float(np.mean(td2 * b["is_weight"].astype(np.float64))), rel=1e-12)


def test_dueling_is_value_minus_separate_mean_stream():
    L = _tiny_learner()
    b = ax.make_transitions(2, seed=3)
    x = L._img(b["state"])
    q, taps = ax.dueling_network(L.params, x, torch.from_numpy(b["previous_action"].astype(np.int64)), 4, return_taps=True)
    assert taps["mean"].shape == (2, 1) and taps["value"].shape == (2, 4)
    assert torch.allclose(q, taps["value"] - taps["mean"])
    # NOT the mean of the value stream:
    assert not torch.allclose(taps["mean"].squeeze(1), taps["value"].mean(dim=1))


def test_reward_clipping_switch_and_done_mask():
    b = ax.make_transitions(4, seed=11)
    b["reward"][:] = [3.0, -2.5, 0.25, 0.0]
    b["done"][:] = [True, False, True, False]
    args = (b["state"], b["next_state"], b["previous_action"], b["action"], b["reward"], b["done"])
    o1 = _tiny_learner().losses(*args)
    o2 = _tiny_learner(reward_clipping="none").losses(*args)
    d = (o2["target_value"] - o1["target_value"]).numpy()
    assert d == pytest.approx([2.0, -1.5, 0.0, 0.0], abs=1e-12)
    # done rows: the target is the (clipped) reward alone
    assert o1["target_value"][0].item() == pytest.approx(1.0) and o1["target_value"][2].item() == pytest.approx(0.25)


def test_gradient_only_through_main_of_current_state():
    L = _tiny_learner()
    b = ax.make_transitions(2, seed=5)
    args = [b[k] for k in ax.TRAIN_FIELDS]
    out, g = L.gradients(*args[:-1], is_weight=args[-1])
    # head gradient, hand-derived: dL/dq[b, a_b] = -2 w (target - q_sa) / B ; the mean stream gets its negative
    diff = (out["target_value"] - out["state_action_value"]).detach().numpy()
    gq = -2.0 * b["is_weight"].astype(np.float64) * diff / 2
    assert g["value3.b"].numpy()[b["action"][0]] != 0.0
    exp_b = np.zeros(4)
    for i in range(2):
        exp_b[b["action"][i]] += gq[i]
    assert g["value3.b"].numpy() == pytest.approx(exp_b, rel=1e-9)
    assert g["mean3.b"].numpy()[0] == pytest.approx(-gq.sum(), rel=1e-9)


def test_adam_first_step_hand_computed():
    """TF1 ApplyAdam, step 1: m = 0.1 g, v = 0.001 g^2, alpha = lr sqrt(1-0.999)/(1-0.9); the update is
    -alpha * 0.1 g / (sqrt(0.001) |g| + 1e-8)  ~  -lr * sign(g) for |g| >> 1e-8."""
    L = _tiny_learner()
    b = ax.make_transitions(2, seed=9)
    args = [b[k] for k in ax.TRAIN_FIELDS]
    p0 = {k: v.detach().clone() for k, v in L.params.items()}
    (loss, td), out, g, gn, lr = L.distributed_train(*args, return_all=True)
    assert lr == pytest.approx(1e-4, rel=1e-6) and L.step == 1
    scale = 40.0 * min(1.0 / gn, 1.0 / 40.0)
    alpha = lr * np.sqrt(1 - 0.999) / (1 - 0.9)
    for k in ("value3.w", "conv1.w"):
        gc = g[k].numpy() * scale
        exp = p0[k].numpy() - alpha * 0.1 * gc / (np.sqrt(0.001 * gc * gc) + 1e-8)
        assert L.params[k].detach().numpy() == pytest.approx(exp, rel=0, abs=2e-9)     # beta powers are float32
    assert float(L.beta1_power) == pytest.approx(0.81, rel=1e-6) and float(L.beta2_power) == pytest.approx(0.998001, rel=1e-6)
    assert td == pytest.approx(np.abs(out["target_value"].numpy() - out["state_action_value"].detach().numpy()))
    # the target network is untouched by the update and only moves with target_to_main (which copies main INTO target)
    t0 = L.target["conv1.w"].clone()
    L.distributed_train(*args)
    assert torch.equal(L.target["conv1.w"], t0)
    L.target_to_main()
    assert torch.equal(L.target["conv1.w"], L.params["conv1.w"].detach())


def test_float32_oracle_tracks_float64():
    b = ax.make_transitions(3, seed=21)
    args = [b[k] for k in ax.TRAIN_FIELDS]
    (l64, td64), _, g64, _, _ = ax.Learner(dtype=torch.float64).distributed_train(*args, return_all=True)
    (l32, td32), _, g32, _, _ = ax.Learner(dtype=torch.float32).distributed_train(*args, return_all=True)
    assert l32 == pytest.approx(l64, rel=1e-5)
    assert td32 == pytest.approx(td64, rel=1e-4, abs=1e-6)
    for k in g64:
        a, c = g64[k].numpy(), g32[k].numpy().astype(np.float64)
        assert np.max(np.abs(a - c)) <= 2e-4 * max(np.max(np.abs(a)), 1e-30), k


# ---- prioritized replay ---------------------------------------------------------------------
def test_sum_tree_known_answer():
    m = per_np.MemoryNP(4)
    for e in (0.0, 0.999, 0.0, 0.0):
        m.add(e)
    p0, p1 = 0.001 ** 0.6, 1.0
    assert m.tree.total() == pytest.approx(3 * p0 + p1, rel=1e-15)
    idx, didx, pr, w = m.sample(2, [0.5, 0.5])
    # segment 0 = [0, total/2): mass total/4 > p0 -> leaf 1 (the big one); segment 1 centre 3 total/4 -> also leaf 1
    assert list(didx) == [1, 1] and list(idx) == [4, 4]
    assert w == pytest.approx([1.0, 1.0])
    assert m.beta == pytest.approx(0.401)


def test_native_per_is_bit_identical_to_numpy(native):
    from distributed_reinforcement_learning_b200.distributed_queue import buffer_queue as bq
    rng = np.random.default_rng(0)
    cap = 37                                   # not a power of two: unbalanced tree
    ref, mem = per_np.MemoryNP(cap), bq.Memory(cap)
    for i in range(90):                        # wraps around 2.4 times
        e = float(abs(rng.standard_normal()))
        ref.add(e)
        mem.add(e, ("transition", i))
        if i % 7 == 3 and i > 10:
            u = rng.random(8)
            ri, rd, rp, rw = ref.sample(8, u)
            batch, idxs, w = mem.sample(8, u)
            assert idxs == list(ri)
            assert np.array_equal(mem.last_priorities, rp)            # bit-identical float64
            assert w == pytest.approx(rw, rel=1e-13)
            for j, d in zip(rd, batch):
                assert d[0] == "transition" and d[1] % cap == j % cap
            for k in idxs[:3]:
                ne = float(abs(rng.standard_normal()))
                ref.update(k, ne)
                mem.update(k, ne)
        assert mem.tree.total() == ref.tree.total()                   # exact
    assert mem.tree.n_entries == cap == ref.tree.n_entries
    assert mem.beta == pytest.approx(float(ref.beta))
    # invariant: every internal node is the sum of its children up to rounding
    t = ref.tree.nodes
    for node in range(cap - 1):
        assert t[node] == pytest.approx(t[2 * node + 1] + t[2 * node + 2], rel=1e-9)


def test_per_errors(native):
    import ctypes as C
    N = native
    p = C.c_void_p()
    assert N.lib.drl_per_create(1, C.byref(p)) == N.DRL_ERR_INVALID
    assert N.lib.drl_per_create(4, C.byref(p)) == 0
    u = np.zeros(2)
    out = [np.zeros(2, np.int64), np.zeros(2, np.int64), np.zeros(2), np.zeros(2)]
    assert N.lib.drl_per_sample(p, 2, N.ptr(u), *[N.ptr(a) for a in out]) == N.DRL_ERR_STATE     # empty
    assert N.lib.drl_per_update(p, 0, 1.0) == N.DRL_ERR_INVALID                                     # not a leaf
    N.lib.drl_per_destroy(p)


def test_apex_agent_surface_and_no_cpu_fallback(native, tmp_path):
    """agent/apex.py constructor kwargs (train_apex.py:48-58), checkpoint round trip on the host copy, and a loud
    failure of every compute call without a CUDA device."""
    from distributed_reinforcement_learning_b200.agent import apex
    from distributed_reinforcement_learning_b200.optimizer import dqn
    kw = dict(input_shape=[84, 84, 4], num_action=4, discount_factor=0.99, gradient_clip_norm=40.0,
              reward_clipping="abs_one", start_learning_rate=1e-4, end_learning_rate=0.0,
              learning_frame=100000000000000, model_name="learner", learner_name="learner")
    ag = apex.Agent(**kw)
    ag.set_session(None)
    m0, t0 = ag._main.copy(), ag._target.copy()
    assert not np.array_equal(m0, t0)
    ag.target_to_main()
    assert np.array_equal(ag._target, m0)                 # target <- main
    ag.save_weights(str(tmp_path / "ck"))
    ag2 = apex.Agent(**dict(kw, model_name="actor_0"))
    ag2.load_weights(str(tmp_path / "ck"))
    assert np.array_equal(ag2._main, m0) and ag2._opt["step"] == 0
    ag2.set_session(None)
    ag2.parameter_sync()
    assert np.array_equal(ag2._main, m0)
    assert dqn.take_state_action_value(np.arange(8.0).reshape(2, 4), [1, 3], 4).tolist() == [1.0, 7.0]
    if native.device_count() == 0:
        b = ax.make_transitions(2)
        with pytest.raises(native.DrlError) as ei:
            ag.distributed_train(*[b[k] for k in ax.TRAIN_FIELDS])
        assert "no CPU fallback" in str(ei.value)


def test_native_per_random_op_sequences(native):
    """Property test (hypothesis): arbitrary interleavings of add / sample / update keep the native sum tree
    bit-identical to the NumPy restatement (totals, sampled leaves, priorities) for arbitrary capacities."""
    from hypothesis import given, settings, strategies as st
    from distributed_reinforcement_learning_b200.distributed_queue import buffer_queue as bq

    ops = st.lists(st.tuples(st.sampled_from(["add", "add", "sample", "update"]),
                             st.floats(min_value=0.0, max_value=50.0, allow_nan=False),
                             st.floats(min_value=0.0, max_value=0.999999)), min_size=1, max_size=60)

    @settings(max_examples=60, deadline=None, derandomize=True, database=None)
    @given(cap=st.integers(min_value=2, max_value=33), seq=ops)
    def run(cap, seq):
        ref, mem = per_np.MemoryNP(cap), bq.Memory(cap)
        last = None
        for kind, err, u in seq:
            if kind == "add":
                ref.add(err)
                mem.add(err, None)
            elif kind == "sample" and ref.tree.n_entries > 0 and ref.tree.total() > 0:
                n = 1 + int(u * 4)
                us = [(u * (i + 1)) % 1.0 for i in range(n)]
                with np.errstate(all="ignore"):
                    ri, rd, rp, rw = ref.sample(n, us)
                _, idxs, w = mem.sample(n, us)
                assert idxs == list(ri) and np.array_equal(mem.last_priorities, rp)
                # a draw of mass exactly 0 can land on a still-empty leaf (priority 0): the reference then yields
                # inf / inf = nan weights, and so do both implementations here
                with np.errstate(all="ignore"):
                    assert np.allclose(w, rw, rtol=1e-12, atol=0.0, equal_nan=True)
                last = idxs
            elif kind == "update" and last:
                k = last[int(u * len(last)) % len(last)]
                ref.update(k, err)
                mem.update(k, err)
            assert mem.tree.total() == ref.tree.total()
        mem.tree.close()
    run()


def test_conv2d_restatement_against_naive_loops():
    """tf.layers.conv2d(padding='VALID', NHWC input, HWIO kernel) = cross-correlation + bias, restated through
    F.conv2d in oracle/impala_torch.py::_conv2d_tf; pinned here by explicit loops."""
    from oracle import impala_torch as it
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 8, 3))
    w = rng.standard_normal((4, 3, 3, 5))
    b = rng.standard_normal(5)
    for stride in (1, 2):
        got = it._conv2d_tf(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride).numpy()
        OH, OW = (9 - 4) // stride + 1, (8 - 3) // stride + 1
        ref = np.zeros((2, OH, OW, 5))
        for n in range(2):
            for oy in range(OH):
                for ox in range(OW):
                    patch = x[n, oy * stride:oy * stride + 4, ox * stride:ox * stride + 3, :]
                    ref[n, oy, ox] = np.tensordot(patch, w, axes=([0, 1, 2], [0, 1, 2])) + b
        assert got.shape == ref.shape and np.allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_tf1_adam_restatement_tracks_torch_adam_when_epsilon_is_negligible():
    """TF1 ApplyAdam (lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); w -= lr_t m / (sqrt(v) + eps)) and torch.optim.Adam
    (bias-corrected m, v; eps outside the corrected sqrt) are the same update up to where eps enters: with gradients far
    above eps the oracle's update must follow torch's over several steps."""
    rng = np.random.default_rng(1)
    w0 = rng.standard_normal(50)
    grads = [np.sign(rng.standard_normal(50)) * (1.0 + 3.0 * np.abs(rng.standard_normal(50))) for _ in range(5)]   # |g| >= 1
    wt = torch.tensor(w0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([wt], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    w, m, v, b1p, b2p = w0.copy(), np.zeros(50), np.zeros(50), 0.9, 0.999
    for g in grads:
        wt.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
        m += (g - m) * (1 - 0.9)
        v += (g * g - v) * (1 - 0.999)
        w -= m * (1e-3 * np.sqrt(1 - b2p) / (1 - b1p)) / (np.sqrt(v) + 1e-8)
        b1p *= 0.9
        b2p *= 0.999
    assert np.allclose(w, wt.detach().numpy(), rtol=0, atol=1e-3 * 1e-6)      # eps placement: <= lr * 3e-7 / |g|
