"""The committed golden vectors of the Ape-X / R2D2 / A3C steps (tests/golden/*.npz, made by make_golden.py from the
float64 oracles) pin those oracles against drift; the float32 oracle (the CPU baseline that bench.py times) must track
them within the parity bar.  PARITY UNPINNED: they pin the oracle, not TensorFlow."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
STRIDE = 61


def _check_grads(z, g, rel):
    seen = 0
    for n, v in g.items():
        v = v.detach().numpy().astype(np.float64)
        if "grad_" + n in z:
            ref = z["grad_" + n].astype(np.float64)
            assert np.max(np.abs(v - ref)) <= rel * max(np.max(np.abs(ref)), 1e-30) + 1e-12, n
        else:
            ref = z["gradsample_" + n].astype(np.float64)
            assert np.max(np.abs(v.ravel()[::STRIDE] - ref)) <= rel * max(np.max(np.abs(ref)), 1e-30) + 1e-12, n
            assert np.sqrt(np.sum(v ** 2)) == pytest.approx(float(z["gradl2_" + n]), rel=rel)
        seen += 1
    assert seen == len(g)


@pytest.mark.parametrize("dtype,rel", [(torch.float64, 1e-7), (torch.float32, 1e-4)])
def test_apex_golden(dtype, rel):
    from oracle import apex_torch as ax
    z = np.load(os.path.join(GOLD, "apex_step_B3.npz"))
    b = ax.make_transitions(int(z["B"]), A=int(z["A"]), seed=int(z["seed"]))
    L = ax.Learner(dtype=dtype, num_action=int(z["A"]))
    (loss, td), out, g, gn, lr = L.distributed_train(*[b[k] for k in ax.TRAIN_FIELDS], return_all=True)
    assert loss == pytest.approx(float(z["loss"]), rel=max(rel, 1e-6)) and gn == pytest.approx(float(z["grad_norm"]), rel=max(rel, 1e-6))
    assert td == pytest.approx(z["td_error"], rel=10 * rel, abs=1e-6)
    for k in ("main_q", "next_main_q", "target_q", "target_value", "state_action_value"):
        assert out[k].detach().numpy() == pytest.approx(z[k], rel=10 * rel, abs=1e-6), k
    _check_grads(z, g, 10 * rel)


@pytest.mark.parametrize("dtype,rel", [(torch.float64, 1e-7), (torch.float32, 1e-3)])
def test_r2d2_golden(dtype, rel):
    from oracle import r2d2_torch as rt
    z = np.load(os.path.join(GOLD, "r2d2_step_B2_S6.npz"))
    b = rt.make_sequences(int(z["B"]), S=int(z["S"]), seed=int(z["seed"]))
    L = rt.Learner(dtype=dtype, seq_len=int(z["S"]), burn_in=int(z["burn_in"]))
    (loss, td), out, g, gn = L.train(*[b[k] for k in rt.TRAIN_FIELDS], return_all=True)
    assert loss == pytest.approx(float(z["loss"]), rel=rel) and gn == pytest.approx(float(z["grad_norm"]), rel=rel)
    for k in ("main_q", "target_q", "target_value", "state_action_value"):
        assert out[k].detach().numpy() == pytest.approx(z[k], rel=rel, abs=1e-5), k
    _check_grads(z, g, 2 * rel)          # float32 h^-1 (sqrt(1.004..) - 1) costs three digits: 1e-3 for the f32 oracle


@pytest.mark.parametrize("dtype,rel", [(torch.float64, 1e-7), (torch.float32, 1e-4)])
def test_a3c_golden(dtype, rel):
    from oracle import a3c_torch as at
    z = np.load(os.path.join(GOLD, "a3c_step_B3.npz"))
    b = at.make_transitions(int(z["B"]), A=int(z["A"]), seed=int(z["seed"]))
    L = at.Learner(dtype=dtype, num_action=int(z["A"]))
    (pi, bl, en, lr), out, g, gn = L.train(*[b[k] for k in at.TRAIN_FIELDS], return_all=True)
    for got, k in ((pi, "pi_loss"), (bl, "baseline_loss"), (en, "entropy"), (gn, "grad_norm")):
        assert got == pytest.approx(float(z[k]), rel=max(10 * rel, 1e-6)), k
    for k in ("policy", "value", "next_value", "advantage"):
        assert out[k].detach().numpy() == pytest.approx(z[k], rel=10 * rel, abs=1e-6), k
    _check_grads(z, g, 10 * rel)
