"""-m gpu: the stand-alone V-trace kernels (through the C-ABI / the optimizer.vtrace mirror) against the
NumPy oracle of optimizer/vtrace.py, on the analytic known-answer cases and on random
shapes including ragged / extreme ones.  Tolerance: 1e-4 relative (north_star); observed ~1e-6."""
import numpy as np
import pytest

from oracle import vtrace_np

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _vt(native):
    from distributed_reinforcement_learning_b200.optimizer import vtrace
    return vtrace


def _rel(a, b):
    return np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / max(np.max(np.abs(b)), 1e-30)


def _rand_fiw(rng, T, B):
    return dict(log_rhos=(rng.standard_normal((T, B)) * 0.7).astype(np.float32),
                discounts=((rng.random((T, B)) > 0.1) * 0.99).astype(np.float32),
                rewards=rng.standard_normal((T, B)).astype(np.float32),
                values=rng.standard_normal((T, B)).astype(np.float32),
                bootstrap_value=rng.standard_normal(B).astype(np.float32))


def test_hand_worked_example(native):
    """App. C.1: rho=[2/3,2] -> vs=[3.11333, 3.8], clipped=[0.66667, 1]."""
    vt = _vt(native)
    pi = np.array([[[.5, .5], [.8, .2]]], np.float32)
    mu = np.array([[[.25, .75], [.4, .6]]], np.float32)
    a = np.array([[1, 0]], np.int32)
    g = np.full((1, 2), .9, np.float32)
    r = np.array([[1., 2.]], np.float32)
    v = np.array([[.5, 1.]], np.float32)
    nv = np.array([[1., 2.]], np.float32)
    vs, rho = vt.from_softmax(mu, pi, a, g, r, v, nv, 2)
    np.testing.assert_allclose(vs, [[3.1133333, 3.8]], rtol=1e-5)
    np.testing.assert_allclose(rho, [[2. / 3., 1.]], rtol=1e-5)


@pytest.mark.parametrize("T,B", [(1, 1), (2, 3), (18, 32), (18, 257), (33, 5), (64, 4), (100, 1000)])
def test_from_importance_weights_random(native, T, B):
    vt = _vt(native)
    rng = np.random.default_rng(T * 1000 + B)
    kw = _rand_fiw(rng, T, B)
    vs, rho = vt.from_importance_weights(**kw)
    kw64 = {k: v.astype(np.float64) for k, v in kw.items()}
    evs, erho = vtrace_np.from_importance_weights(**kw64)
    assert _rel(vs, evs) < RTOL and _rel(rho, erho) < RTOL
    dvs, _ = vtrace_np.from_importance_weights_direct(**kw64)          # independent O(T^2) definition
    if T <= 33:
        assert _rel(vs, dvs) < RTOL


def test_from_importance_weights_no_clip_and_threshold(native):
    vt = _vt(native)
    rng = np.random.default_rng(7)
    kw = _rand_fiw(rng, 12, 9)
    for clip in (None, 0.5, 2.0):
        vs, rho = vt.from_importance_weights(clip_rho_threshold=clip, **kw)
        evs, erho = vtrace_np.from_importance_weights(clip_rho_threshold=clip,
                                                      **{k: v.astype(np.float64) for k, v in kw.items()})
        assert _rel(vs, evs) < RTOL and _rel(rho, erho) < RTOL


def test_on_policy_is_n_step_return(native):
    """App. C.2: pi == mu, no dones -> vs_t = sum gamma^k r + gamma^n V_boot."""
    vt = _vt(native)
    T, B = 18, 6
    rng = np.random.default_rng(3)
    r = rng.standard_normal((T, B)).astype(np.float32)
    v = rng.standard_normal((T, B)).astype(np.float32)
    boot = rng.standard_normal(B).astype(np.float32)
    g = np.full((T, B), 0.99, np.float32)
    vs, rho = vt.from_importance_weights(np.zeros((T, B), np.float32), g, r, v, boot)
    exp = np.zeros((T, B))
    acc = boot.astype(np.float64)
    for t in range(T - 1, -1, -1):
        acc = r[t] + 0.99 * acc
        exp[t] = acc
    assert _rel(vs, exp) < RTOL
    np.testing.assert_array_equal(rho, np.ones((T, B), np.float32))


def test_all_done(native):
    """App. C.3: gamma == 0 -> vs_t = V_t + rho_bar_t (r_t - V_t)."""
    vt = _vt(native)
    rng = np.random.default_rng(4)
    kw = _rand_fiw(rng, 9, 5)
    kw["discounts"][:] = 0
    vs, rho = vt.from_importance_weights(**kw)
    exp = kw["values"] + np.minimum(1, np.exp(kw["log_rhos"])) * (kw["rewards"] - kw["values"])
    assert _rel(vs, exp) < RTOL


@pytest.mark.parametrize("B,T,A", [(1, 1, 2), (3, 18, 18), (32, 18, 18), (5, 20, 6), (7, 33, 18), (2, 70, 3),
                                   (300, 18, 18), (4, 18, 31)])
def test_from_softmax_random(native, B, T, A):
    vt = _vt(native)
    rng = np.random.default_rng(B * 7919 + T * 31 + A)

    def sm(x):
        e = np.exp(x - x.max(-1, keepdims=True))
        return (e / e.sum(-1, keepdims=True)).astype(np.float32)
    mu = sm(rng.standard_normal((B, T, A)))
    pi = sm(rng.standard_normal((B, T, A)) * 2)
    a = rng.integers(0, A, (B, T)).astype(np.int32)
    g = ((rng.random((B, T)) > 0.1) * 0.99).astype(np.float32)
    r = rng.standard_normal((B, T)).astype(np.float32)
    v = rng.standard_normal((B, T)).astype(np.float32)
    nv = rng.standard_normal((B, T)).astype(np.float32)
    vs, rho = vt.from_softmax(mu, pi, a, g, r, v, nv, A)
    evs, erho = vtrace_np.from_softmax(mu.astype(np.float64), pi.astype(np.float64), a, g.astype(np.float64),
                                       r.astype(np.float64), v.astype(np.float64), nv.astype(np.float64), A)
    assert _rel(vs, evs) < RTOL and _rel(rho, erho) < RTOL


def test_extreme_rho(native):
    """Very off-policy actions: rho >> 1 is clipped, rho << 1 passes through."""
    vt = _vt(native)
    B, T, A = 4, 18, 18
    rng = np.random.default_rng(11)
    mu = np.full((B, T, A), 1e-6, np.float32)
    mu[..., 0] = 1 - (A - 1) * 1e-6
    pi = np.full((B, T, A), (1 - 1e-6) / (A - 1), np.float32)
    pi[..., 0] = 1e-6
    a = rng.integers(0, 2, (B, T)).astype(np.int32)      # action 0: rho = 1e-6 ; action 1: rho ~ 6e4
    g = np.full((B, T), 0.99, np.float32)
    r = rng.standard_normal((B, T)).astype(np.float32)
    v = rng.standard_normal((B, T)).astype(np.float32)
    nv = rng.standard_normal((B, T)).astype(np.float32)
    vs, rho = vt.from_softmax(mu, pi, a, g, r, v, nv, A)
    evs, erho = vtrace_np.from_softmax(mu.astype(np.float64), pi.astype(np.float64), a, g.astype(np.float64),
                                       r.astype(np.float64), v.astype(np.float64), nv.astype(np.float64), A)
    assert _rel(vs, evs) < RTOL and _rel(rho, erho) < RTOL
    assert rho.max() <= 1.0


def test_losses_and_log_probs(native):
    vt = _vt(native)
    B, T, A = 6, 18, 18
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, T, A))
    e = np.exp(x - x.max(-1, keepdims=True))
    sm = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    a = rng.integers(0, A, (B, T)).astype(np.int32)
    adv = rng.standard_normal((B, T)).astype(np.float32)
    vs = rng.standard_normal((B, T)).astype(np.float32)
    val = rng.standard_normal((B, T)).astype(np.float32)
    s64 = sm.astype(np.float64)
    assert abs(vt.compute_policy_gradient_loss(sm, a, adv, A) /
               vtrace_np.compute_policy_gradient_loss(s64, a, adv.astype(np.float64), A) - 1) < RTOL
    assert abs(vt.compute_baseline_loss(vs, val) /
               vtrace_np.compute_baseline_loss(vs.astype(np.float64), val.astype(np.float64)) - 1) < RTOL
    assert abs(vt.compute_entropy_loss(sm) / vtrace_np.compute_entropy_loss(s64) - 1) < RTOL
    lp = vt.log_probs_from_softmax_and_actions(sm, a, A)
    assert _rel(lp, vtrace_np.log_probs_from_softmax_and_actions(s64, a, A)) < RTOL
    f, m, l = vt.split_data(sm)
    assert f.shape == (B, T - 2, A) and np.shares_memory(f, sm) and np.array_equal(m, sm[:, 1:-1]) \
        and np.array_equal(l, sm[:, 2:])


def test_bad_arguments(native):
    vt = _vt(native)
    with pytest.raises(ValueError):
        vt.from_importance_weights(np.zeros((3, 2), np.float32), np.zeros((3, 3), np.float32),
                                   np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32),
                                   np.zeros(2, np.float32))
    with pytest.raises(ValueError):
        vt.from_softmax(np.zeros((2, 3, 4), np.float32), np.zeros((2, 3, 4), np.float32), np.zeros((2, 3), np.int32),
                        np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3)), 5)
