"""CPU tests of the R2D2 oracle (oracle/r2d2_torch.py) and the host-side R2D2 mirror modules.  PARITY UNPINNED: the
reference ships no tests; pinned by hand-worked known answers and invariants."""
import numpy as np
import pytest
import torch

from oracle import r2d2_torch as rt


def test_param_inventory():
    n = (8 * 8 * 1 * 32 + 32) + 32832 + 36928 + 1280 + 65792 + (3456 * 256 + 256) + (64 * 128 + 128) + (128 * 4 + 4) + 129
    assert rt.param_count() == n == 1032869
    from distributed_reinforcement_learning_b200.model import r2d2_lstm
    assert r2d2_lstm.param_specs(num_action=6, input_shape=(84, 84, 4)) == rt.param_specs(num_action=6, input_shape=(84, 84, 4))


def test_value_rescaling_is_inverse_pair_and_matches_formula():
    x = torch.tensor([-30.0, -1.5, -1e-3, 0.0, 2e-3, 0.7, 12.0, 400.0], dtype=torch.float64)
    assert torch.allclose(rt.vf_rescale(rt.vf_rescale_inv(x)), x, rtol=1e-10, atol=1e-12)
    assert torch.allclose(rt.vf_rescale_inv(rt.vf_rescale(x)), x, rtol=1e-10, atol=1e-12)
    assert rt.vf_rescale(torch.tensor(3.0, dtype=torch.float64)).item() == pytest.approx(1.0 + 3e-3)      # sqrt(4) - 1 + eps x
    from distributed_reinforcement_learning_b200.optimizer import burn_in
    assert burn_in.value_function_rescaling(x.numpy(), 1e-3) == pytest.approx(rt.vf_rescale(x).numpy(), rel=1e-14)
    assert burn_in.inverse_value_function_rescaling(x.numpy(), 1e-3) == pytest.approx(rt.vf_rescale_inv(x).numpy(), rel=1e-14)


def test_done_resets_carried_state_after_the_step():
    """model/r2d2_lstm.py:79-81: the q of step i uses the un-masked output; steps after a done start from zero state."""
    L = rt.Learner(dtype=torch.float64, seq_len=5, burn_in=1)
    b = rt.make_sequences(1, S=5, seed=3)
    b["done"][:] = False
    b["done"][0, 1] = True
    x = L._img(b["state"])
    pa = torch.from_numpy(b["previous_action"].astype(np.int64))
    d = torch.from_numpy(b["done"])
    h0 = torch.from_numpy(b["h"][:, 0]).double()
    c0 = torch.from_numpy(b["c"][:, 0]).double()
    q, _ = rt.unroll(L.params, x, pa, d, h0, c0, 4)
    z = torch.zeros(1, 64, dtype=torch.float64)
    q_from_zero, _ = rt.unroll(L.params, x[:, 2:], pa[:, 2:], d[:, 2:], z, z, 4)
    assert torch.allclose(q[:, 2:], q_from_zero, rtol=1e-12, atol=1e-14)          # steps 2.. restart from zeros
    q_no_done, _ = rt.unroll(L.params, x, pa, torch.zeros_like(d), h0, c0, 4)
    assert torch.allclose(q[:, :2], q_no_done[:, :2])                             # step 1 itself is unaffected
    assert not torch.allclose(q[:, 2], q_no_done[:, 2])


def test_loss_window_and_td_error_known_answer():
    S, bi = 7, 2
    L = rt.Learner(dtype=torch.float64, seq_len=S, burn_in=bi)
    b = rt.make_sequences(2, S=S, seed=9)
    args = [b[k] for k in rt.TRAIN_FIELDS]
    o = L.losses(*args[:-1], weight=args[-1])
    mq, tq = o["main_q"].detach().numpy(), o["target_q"].numpy()
    Nt = S - bi - 1
    assert o["target_value"].shape == (2, Nt)
    from distributed_reinforcement_learning_b200.optimizer import burn_in
    tot = 0.0
    for i in range(2):
        sq = 0.0
        for k in range(Nt):
            t = bi + k
            na = int(np.argmax(mq[i, t + 1]))
            disc = 0.0 if b["done"][i, t] else 0.997
            tgt = burn_in.value_function_rescaling(
                burn_in.inverse_value_function_rescaling(tq[i, t + 1, na], 1e-3) * disc + float(b["reward"][i, t]), 1e-3)
            assert o["target_value"][i, k].item() == pytest.approx(float(tgt), rel=1e-12)
            sq += (float(tgt) - mq[i, t, b["action"][i, t]]) ** 2
        tot += float(b["weight"][i]) * sq / Nt
    assert o["value_loss"].item() == pytest.approx(tot / 2, rel=1e-12)
    td = L.get_td_error(*[b[k][0] for k in rt.TRAIN_FIELDS[:-1]])
    diff = (o["target_value"] - o["state_action_value"]).detach().numpy()
    assert td == pytest.approx(abs(diff[0].mean()), rel=1e-10)


def test_gradient_flows_through_burn_in_steps():
    """burn_in only slices the loss (agent/r2d2.py:64-68): frames before the window still receive gradient through (h, c)."""
    S, bi = 6, 3
    L = rt.Learner(dtype=torch.float64, seq_len=S, burn_in=bi)
    b = rt.make_sequences(1, S=S, seed=4)
    b["done"][:] = False
    x = L._img(b["state"]).requires_grad_(True)
    pa = torch.from_numpy(b["previous_action"].astype(np.int64))
    d = torch.from_numpy(b["done"])
    q, _ = rt.unroll(L.params, x, pa, d, torch.from_numpy(b["h"][:, 0]).double(), torch.from_numpy(b["c"][:, 0]).double(), 4)
    q[:, bi:].sum().backward()
    g = x.grad.abs().reshape(S, -1).sum(1)
    assert (g[:bi] > 0).all()
    # ...and a done inside the burn-in cuts it
    b["done"][0, 1] = True
    x2 = L._img(b["state"]).requires_grad_(True)
    q2, _ = rt.unroll(L.params, x2, pa, torch.from_numpy(b["done"]), torch.from_numpy(b["h"][:, 0]).double(),
                      torch.from_numpy(b["c"][:, 0]).double(), 4)
    q2[:, bi:].sum().backward()
    g2 = x2.grad.abs().reshape(S, -1).sum(1)
    assert g2[0] == 0 and g2[1] == 0 and g2[2] > 0


def test_float32_oracle_tracks_float64():
    b = rt.make_sequences(2, S=6, seed=21)
    args = [b[k] for k in rt.TRAIN_FIELDS]
    (l64, td64), _, g64, _ = rt.Learner(dtype=torch.float64, seq_len=6, burn_in=2).train(*args, return_all=True)
    (l32, td32), _, g32, _ = rt.Learner(dtype=torch.float32, seq_len=6, burn_in=2).train(*args, return_all=True)
    assert l32 == pytest.approx(l64, rel=2e-4)            # float32 h^-1 loses ~3 digits (sqrt(1.004..) - 1)
    for k in g64:
        a, c = g64[k].numpy(), g32[k].numpy().astype(np.float64)
        assert np.max(np.abs(a - c)) <= 1e-3 * max(np.max(np.abs(a)), 1e-30), k


def test_r2d2_agent_surface_and_no_cpu_fallback(native):
    from distributed_reinforcement_learning_b200.agent import r2d2
    kw = dict(seq_len=15, burn_in=7, input_shape=[84, 84, 1], num_action=4, lstm_size=64, discount_factor=0.997,
              start_learning_rate=1e-4, end_learning_rate=0.0, learning_frame=1000000000, gradient_clip_norm=40.0,
              model_name="learner", learner_name="learner")          # train_r2d2.py:50-63
    ag = r2d2.Agent(**kw)
    ag.set_session(None)
    m0 = ag._main.copy()
    assert not np.array_equal(m0, ag._target)
    ag.main_to_target()
    assert np.array_equal(ag._target, m0)
    a2 = r2d2.Agent(**dict(kw, model_name="actor_0"))
    a2.set_session(None)
    a2.parameter_sync()
    assert np.array_equal(a2._main, m0)
    if native.device_count() == 0:
        b = rt.make_sequences(2)
        with pytest.raises(native.DrlError) as ei:
            ag.train(list(b["state"]), list(b["previous_action"]), list(b["action"]), list(b["h"]), list(b["c"]),
                     list(b["reward"]), list(b["done"]), b["weight"])
        assert "no CPU fallback" in str(ei.value)


def test_lstm_cell_restatement_agrees_with_torch_lstmcell():
    """The TF 1.14 LSTMCell restatement (kernel [x|h] -> gates i, j, f, o; forget_bias 1.0 outside the bias variable)
    against an independent implementation: torch.nn.LSTMCell (gates i, f, g, o) with the weights permuted and the
    forget bias folded in.  Pins the cell used by oracle/impala_torch.py::lstm and oracle/r2d2_torch.py::network."""
    from oracle import impala_torch as it
    torch.manual_seed(0)
    n, X, L = 5, 37, 16
    W = torch.randn(X + L, 4 * L, dtype=torch.float64) * 0.3
    bias = torch.randn(4 * L, dtype=torch.float64) * 0.1
    x, h0, c0 = (torch.randn(n, X, dtype=torch.float64), torch.randn(n, L, dtype=torch.float64) * 0.5,
                 torch.randn(n, L, dtype=torch.float64))
    h1, c1, _ = it.lstm({"lstm.w": W, "lstm.b": bias}, x, h0, c0)
    cell = torch.nn.LSTMCell(X, L).double()
    i, j, f, o = torch.chunk(W, 4, dim=1)                      # TF column blocks
    bi, bj, bf, bo = torch.chunk(bias, 4)
    with torch.no_grad():
        Wt = torch.cat([i, f, j, o], dim=1)                    # torch row blocks: i, f, g, o
        cell.weight_ih.copy_(Wt[:X].t())
        cell.weight_hh.copy_(Wt[X:].t())
        cell.bias_ih.copy_(torch.cat([bi, bf + 1.0, bj, bo]))  # forget_bias = 1.0
        cell.bias_hh.zero_()
        h_ref, c_ref = cell(x, (h0, c0))
    assert torch.allclose(h1, h_ref, rtol=1e-12, atol=1e-13) and torch.allclose(c1, c_ref, rtol=1e-12, atol=1e-13)
    # and the R2D2 network step uses the same cell
    p = rt.init_params(0, torch.float64)
    b = rt.make_sequences(2, S=2, seed=1)
    q, h2, c2, _ = rt.network(p, torch.from_numpy(b["state"][:, 0]).double() / 255, torch.from_numpy(b["previous_action"][:, 0]).long(),
                              torch.from_numpy(b["h"][:, 0]).double(), torch.from_numpy(b["c"][:, 0]).double(), 4)
    assert h2.shape == (2, 64) and torch.all(h2.abs() < 1) and q.shape == (2, 4)
