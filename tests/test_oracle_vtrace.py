"""CPU: pins the V-trace oracle (oracle/vtrace_np.py, restating optimizer/vtrace.py) with the analytic
known-answer cases of SURVEY.md Appendix C, an independent O(T^2) closed form, the torch restatement and
hypothesis property tests.  PARITY UNPINNED against TensorFlow (not installable) -- see oracle/__init__.py."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import impala_torch as it
from oracle import vtrace_np as vt


def _rand(rng, T, B):
    return dict(log_rhos=rng.standard_normal((T, B)) * 0.7, discounts=(rng.random((T, B)) > 0.1) * 0.99,
                rewards=rng.standard_normal((T, B)), values=rng.standard_normal((T, B)),
                bootstrap_value=rng.standard_normal(B))


def test_hand_worked_example():
    """App. C.1 (time-major, B=1, T'=2)."""
    pi = np.array([[[.5, .5]], [[.8, .2]]])
    mu = np.array([[[.25, .75]], [[.4, .6]]])
    a = np.array([[1], [0]])
    lr = np.log(np.take_along_axis(pi, a[..., None], 2)[..., 0]) - np.log(np.take_along_axis(mu, a[..., None], 2)[..., 0])
    vs, rho = vt.from_importance_weights(lr, np.full((2, 1), .9), np.array([[1.], [2.]]), np.array([[.5], [1.]]),
                                         np.array([2.]))
    np.testing.assert_allclose(rho[:, 0], [2 / 3, 1.0], rtol=1e-12)
    np.testing.assert_allclose(vs[:, 0], [0.5 + (2 / 3) * 1.4 + 0.9 * (2 / 3) * 2.8, 3.8], rtol=1e-12)
    np.testing.assert_allclose(vs[:, 0], [3.113333333333, 3.8], rtol=1e-9)


def test_on_policy_n_step_return():
    """App. C.2."""
    rng = np.random.default_rng(0)
    T, B = 18, 4
    r, v, boot = rng.standard_normal((T, B)), rng.standard_normal((T, B)), rng.standard_normal(B)
    vs, rho = vt.from_importance_weights(np.zeros((T, B)), np.full((T, B), .99), r, v, boot)
    acc, exp = boot.copy(), np.zeros((T, B))
    for t in range(T - 1, -1, -1):
        acc = r[t] + .99 * acc
        exp[t] = acc
    np.testing.assert_allclose(vs, exp, rtol=1e-12)
    assert np.all(rho == 1.0)


def test_all_done():
    """App. C.3."""
    kw = _rand(np.random.default_rng(1), 7, 3)
    kw["discounts"] = np.zeros((7, 3))
    vs, rho = vt.from_importance_weights(**kw)
    np.testing.assert_allclose(vs, kw["values"] + rho * (kw["rewards"] - kw["values"]), rtol=1e-12)


def test_clip_pg_rho_threshold_is_dead_and_cs_hardcoded():
    """optimizer/vtrace.py:72,80 quirks."""
    kw = _rand(np.random.default_rng(2), 9, 3)
    a = vt.from_importance_weights(clip_pg_rho_threshold=1.0, **kw)
    b = vt.from_importance_weights(clip_pg_rho_threshold=123.0, **kw)
    np.testing.assert_array_equal(a[0], b[0])
    vs_none, rho_none = vt.from_importance_weights(clip_rho_threshold=None, **kw)
    np.testing.assert_allclose(rho_none, np.exp(kw["log_rhos"]))
    assert not np.allclose(vs_none, a[0])


@settings(max_examples=40, deadline=None, derandomize=True, database=None)
@given(T=st.integers(1, 32), B=st.integers(1, 9), seed=st.integers(0, 2 ** 31 - 1))
def test_recursion_equals_direct_definition(T, B, seed):
    """Serial reverse scan == sum_k gamma^k (prod c) delta evaluated directly (second oracle)."""
    kw = _rand(np.random.default_rng(seed), T, B)
    vs, rho = vt.from_importance_weights(**kw)
    dvs, drho = vt.from_importance_weights_direct(**kw)
    np.testing.assert_allclose(vs, dvs, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(rho, drho, rtol=1e-12)


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(T=st.integers(1, 32), B=st.integers(1, 6), seed=st.integers(0, 2 ** 31 - 1))
def test_scan_is_associative(T, B, seed):
    """App. C.5: (a1,b1) o (a2,b2) = (a1 a2, b1 + a1 b2) reproduces the serial scan (what the warp scan uses)."""
    kw = _rand(np.random.default_rng(seed), T, B)
    rho = np.exp(kw["log_rhos"])
    c = np.minimum(1, rho)
    v1 = np.concatenate([kw["values"][1:], kw["bootstrap_value"][None]])
    delta = np.minimum(1, rho) * (kw["rewards"] + kw["discounts"] * v1 - kw["values"])
    a, b = kw["discounts"] * c, delta.copy()
    d = 1
    while d < T:                                  # Hillis-Steele suffix scan
        a2 = np.concatenate([a[d:], np.ones((d, B))])[:T]
        b2 = np.concatenate([b[d:], np.zeros((d, B))])[:T]
        b, a = b + a * b2, a * a2
        d *= 2
    vs, _ = vt.from_importance_weights(**kw)
    np.testing.assert_allclose(b + kw["values"], vs, rtol=1e-9, atol=1e-9)


def test_numpy_and_torch_restatements_agree():
    rng = np.random.default_rng(3)
    B, T, A = 5, 18, 18
    x = rng.standard_normal((2, B, T, A))
    sm = np.exp(x) / np.exp(x).sum(-1, keepdims=True)
    a = rng.integers(0, A, (B, T))
    g, r, v, nv = (rng.random((B, T)) > .1) * .99, rng.standard_normal((B, T)), rng.standard_normal((B, T)), \
        rng.standard_normal((B, T))
    vs, rho = vt.from_softmax(sm[0], sm[1], a, g, r, v, nv, A)
    tv, tr = it.from_softmax(*(torch.from_numpy(z) for z in (sm[0], sm[1], a, g, r, v, nv)), A)
    np.testing.assert_allclose(vs, tv.numpy(), rtol=1e-12)
    np.testing.assert_allclose(rho, tr.numpy(), rtol=1e-12)
    adv = rng.standard_normal((B, T))
    np.testing.assert_allclose(vt.compute_policy_gradient_loss(sm[1], a, adv, A),
                               float(it.compute_policy_gradient_loss(torch.from_numpy(sm[1]), torch.from_numpy(a),
                                                                     torch.from_numpy(adv), A)), rtol=1e-12)
    np.testing.assert_allclose(vt.compute_baseline_loss(vs, v), float(it.compute_baseline_loss(tv, torch.from_numpy(v))),
                               rtol=1e-12)
    np.testing.assert_allclose(vt.compute_entropy_loss(sm[1]), float(it.compute_entropy_loss(torch.from_numpy(sm[1]))),
                               rtol=1e-12)


def test_split_data_and_log_probs():
    x = np.arange(2 * 6 * 3).reshape(2, 6, 3)
    f, m, l = vt.split_data(x)
    assert np.array_equal(f, x[:, :4]) and np.array_equal(m, x[:, 1:5]) and np.array_equal(l, x[:, 2:])
    p = np.array([[[.1, .9], [.7, .3]]])
    np.testing.assert_allclose(vt.log_probs_from_softmax_and_actions(p, np.array([[1, 0]]), 2), np.log([[.9, .7]]))
    # tf.one_hot gives zeros for out-of-range actions -> log(0) = -inf, no epsilon
    assert np.isneginf(vt.log_probs_from_softmax_and_actions(p, np.array([[5, 0]]), 2)[0, 0])


def test_entropy_without_epsilon_is_nan_on_zero_prob():
    """optimizer/vtrace.py:122: log(0)*0 -> NaN, preserved."""
    assert np.isnan(vt.compute_entropy_loss(np.array([[[1.0, 0.0]]])))


def test_golden_vtrace_fixture():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "vtrace_T18_B8.npz"))
    vs, rho = vt.from_importance_weights(z["log_rhos"], z["discounts"], z["rewards"], z["values"], z["bootstrap_value"])
    np.testing.assert_allclose(vs, z["vs"], rtol=1e-12)
    np.testing.assert_allclose(rho, z["clipped_rhos"], rtol=1e-12)


def test_c_restatement_agrees_with_numpy_oracle():
    """oracle/c/vtrace_c.c (plain C, gcc) against oracle/vtrace_np.py and the O(T^2) closed form: two independent
    restatements of optimizer/vtrace.py pinning each other."""
    from oracle import vtrace_c
    rng = np.random.default_rng(5)
    for T, B, A in ((1, 1, 2), (18, 5, 18), (32, 3, 7)):
        lr = rng.standard_normal((T, B)) * 0.8
        g = (rng.random((T, B)) > 0.15) * 0.99
        r, v, boot = rng.standard_normal((T, B)), rng.standard_normal((T, B)), rng.standard_normal(B)
        for clip in (1.0, 0.7, None):
            vs_c, rho_c = vtrace_c.from_importance_weights(lr, g, r, v, boot, clip)
            vs_n, rho_n = vt.from_importance_weights(lr, g, r, v, boot, clip)
            assert np.allclose(vs_c, vs_n, rtol=1e-13, atol=1e-13) and np.allclose(rho_c, rho_n, rtol=1e-15)
        vs_d, _ = vt.from_importance_weights_direct(lr, g, r, v, boot, 1.0)
        assert np.allclose(vtrace_c.from_importance_weights(lr, g, r, v, boot, 1.0)[0], vs_d, rtol=1e-10, atol=1e-12)
        logits = rng.standard_normal((2, B, T, A))
        sm = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
        act = rng.integers(0, A, (B, T)).astype(np.int32)
        args = (sm[0], sm[1], act, g.T.copy(), r.T.copy(), v.T.copy(), rng.standard_normal((B, T)), A)
        vs_c, rho_c = vtrace_c.from_softmax(*args)
        vs_n, rho_n = vt.from_softmax(*args)
        assert np.allclose(vs_c, vs_n, rtol=1e-12, atol=1e-13) and np.allclose(rho_c, rho_n, rtol=1e-13)
        adv = rng.standard_normal((B, T))
        pg, bl, en = vtrace_c.losses(sm[1], act, adv, vs_n, v.T.copy())
        assert pg == pytest.approx(float(vt.compute_policy_gradient_loss(sm[1], act, adv, A)), rel=1e-12)
        assert bl == pytest.approx(float(vt.compute_baseline_loss(vs_n, v.T)), rel=1e-12)
        assert en == pytest.approx(float(vt.compute_entropy_loss(sm[1])), rel=1e-12)
