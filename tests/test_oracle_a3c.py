"""CPU tests of the A3C oracle (oracle/a3c_torch.py) and the host-side A3C mirror modules.  PARITY UNPINNED."""
import numpy as np
import pytest
import torch

from oracle import a3c_torch as at


def test_losses_known_answer_and_host_functions():
    from distributed_reinforcement_learning_b200.optimizer import a2c
    L = at.Learner(dtype=torch.float64)
    b = at.make_transitions(3, seed=12)
    o = L.losses(*[b[k] for k in at.TRAIN_FIELDS])
    pol, val, nv = o["policy"].detach().numpy(), o["value"].detach().numpy(), o["next_value"].numpy()
    r = np.clip(b["reward"].astype(np.float64), -1, 1)
    disc = np.where(b["done"], 0.0, 0.997)
    adv = r + disc * nv - val
    assert o["advantage"].numpy() == pytest.approx(adv, rel=1e-12)
    sel = pol[np.arange(3), b["action"]]
    assert o["pi_loss"].item() == pytest.approx(-np.mean(adv * sel), rel=1e-12)           # probability, not log-probability
    assert o["baseline_loss"].item() == pytest.approx(np.mean(adv ** 2), rel=1e-12)
    assert o["entropy"].item() == pytest.approx(np.mean(np.sum(pol * np.log(pol), axis=1)), rel=1e-12)
    assert o["total_loss"].item() == pytest.approx(o["pi_loss"].item() + o["baseline_loss"].item() + 0.05 * o["entropy"].item())
    assert a2c.compute_policy_loss(pol, b["action"], val, nv, disc, r, 4) == pytest.approx(o["pi_loss"].item(), rel=1e-12)
    assert a2c.compute_baseline_loss(val, nv, disc, r) == pytest.approx(o["baseline_loss"].item(), rel=1e-12)
    assert a2c.compute_entropy_loss(pol) == pytest.approx(o["entropy"].item(), rel=1e-12)


def test_head_gradients_hand_derived():
    """dL/dV = -2 bc adv / B ; dL/dlogits = pi (dpi - <pi, dpi>) with dpi = (-adv 1[a] + ec (log pi + 1)) / B."""
    L = at.Learner(dtype=torch.float64, baseline_loss_coef=0.7, entropy_coef=0.03)
    b = at.make_transitions(2, seed=8)
    res, o, g, gn = L.train(*[b[k] for k in at.TRAIN_FIELDS], return_all=True)
    pol, adv = o["policy"].detach().numpy(), o["advantage"].numpy()
    assert g["critic3.b"].numpy()[0] == pytest.approx(np.sum(-2 * 0.7 * adv / 2), rel=1e-9)
    exp = np.zeros(4)
    for i in range(2):
        dpi = 0.03 * (np.log(pol[i]) + 1.0) / 2
        dpi[b["action"][i]] += -adv[i] / 2
        exp += pol[i] * (dpi - np.dot(pol[i], dpi))
    assert g["actor3.b"].numpy() == pytest.approx(exp, rel=1e-8, abs=1e-14)


def test_soft_asymmetric_clipping_and_next_value_is_constant():
    b = at.make_transitions(3, seed=4)
    b["reward"][:] = [-10.0, 2.0, 0.0]
    L = at.Learner(dtype=torch.float64, reward_clipping="soft_asymmetric")
    o = L.losses(*[b[k] for k in at.TRAIN_FIELDS])
    r = b["reward"].astype(np.float64)
    sq = np.tanh(r / 5)
    cr = np.where(r < 0, 0.3 * sq, sq) * 5
    disc = np.where(b["done"], 0.0, 0.997)
    assert o["advantage"].numpy() == pytest.approx(cr + disc * o["next_value"].numpy() - o["value"].detach().numpy(), rel=1e-12)
    assert not o["next_value"].requires_grad


def test_a3c_agent_surface_and_no_cpu_fallback(native):
    from distributed_reinforcement_learning_b200.agent import a3c
    from distributed_reinforcement_learning_b200.model import actor_critic
    assert actor_critic.param_specs(num_action=6) == at.param_specs(num_action=6)
    kw = dict(input_shape=[84, 84, 4], num_action=4, discount_factor=0.997, start_learning_rate=1e-4,
              end_learning_rate=0.0, learning_frame=1000000000, baseline_loss_coef=1.0, entropy_coef=0.05,
              gradient_clip_norm=40.0, reward_clipping="abs_one", model_name="learner", learner_name="learner")
    ag = a3c.Agent(**kw)
    ag.set_session(None)
    a2 = a3c.Agent(**dict(kw, model_name="actor_0"))
    a2.set_session(None)
    a2.parameter_sync()
    assert np.array_equal(a2._params, ag._params)
    with pytest.raises(AssertionError):
        a3c.Agent(**dict(kw, reward_clipping="none"))
    if native.device_count() == 0:
        b = at.make_transitions(2)
        with pytest.raises(native.DrlError) as ei:
            ag.train(*[b[k] for k in at.TRAIN_FIELDS])
        assert "no CPU fallback" in str(ei.value)
