"""-m gpu: the CUDA learner step through the C-ABI against the float64 CPU oracle on identical seeded
synthetic trajectories: activations, V-trace taps, the three losses, EVERY
parameter gradient, and the parameters / RMSProp slots after the update.  Bar: 1e-4 relative fp32
(BASELINE.json north_star), metric in tests/parity.py."""
import numpy as np
import pytest
import torch

import parity
from oracle import impala_torch as it
from oracle import synthetic

pytestmark = pytest.mark.gpu


def _assert_all(errs):
    bad = parity.failures(errs)
    assert not bad, "parity failures (rel err): %s" % sorted(bad.items(), key=lambda kv: -kv[1])[:12]


def test_step_small_config(native):
    """BASELINE config 1 shape: B=4, T=20."""
    _assert_all(parity.compare_step(4, T=20))


def test_step_reference_config(native):
    """BASELINE config 2 shape: B=32, T=20 (the bench workload)."""
    _assert_all(parity.compare_step(32, T=20, layers=False))


@pytest.mark.parametrize("B,T,A", [(1, 3, 2), (2, 5, 18), (3, 7, 6), (5, 32, 18), (1, 20, 31)])
def test_step_ragged_shapes(native, B, T, A):
    """Odd batch sizes (rows not a multiple of any tile), minimum / maximum trajectory, other action counts."""
    _assert_all(parity.compare_step(B, T=T, A=A))


def test_three_steps_track_oracle(native):
    """Three updates on the same batch: losses/lr/grad-norm of every step and the final parameters and
    RMSProp slots (ms0 = 1, eps inside sqrt, global-norm clip, step counter) follow the oracle."""
    _assert_all(parity.compare_step(4, T=20, steps=3, layers=False))


def test_soft_asymmetric_reward_clipping(native):
    _assert_all(parity.compare_step(3, T=9, reward_clipping="soft_asymmetric", layers=False))


def test_cuda_graph_path_matches(native):
    _assert_all(parity.compare_step(4, T=20, steps=2, layers=False, use_cuda_graph=True))


@pytest.mark.parametrize("fixture,kink_tol", [("impala_step_B2_T6.npz", parity.TOL), ("impala_step_B4_T20.npz", 2e-3)])
def test_matches_golden_fixture(native, fixture, kink_tol):
    """The committed golden vectors (tests/golden/, written by make_golden.py by EXECUTING the unmodified reference files
    agent/impala.py / optimizer/vtrace.py / model/impala_actor_critic.py over oracle/tf1_shim in float64): losses, V-trace
    taps, all 24 gradients, and the parameters / RMSProp slots after the reference's own train_op.  B4_T20 is BASELINE
    configs[0]."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", fixture))
    assert "reference files executed" in str(z["source"])
    B, T, A = int(z["B"]), int(z["T"]), int(z["A"])
    batch, params, cfg = parity.make_case(B, T, A, seed=int(z["seed"]))
    eng = parity.native_learner(batch, params, cfg)
    try:
        eng.stage(0, *[batch[k] for k in synthetic.TRAIN_FIELDS])
        out = eng.step(0)
        taps = eng.taps()
        g = eng.get_grads()
        p_after = eng.get_params()
        ms_after, step = eng.get_opt_state()
    finally:
        eng.close()
    assert step == 1
    for k in ("pi_loss", "baseline_loss", "entropy"):
        assert abs(out[k] - float(z[k])) <= parity.TOL * max(abs(float(z[k])), 1e-30), k
    assert abs(out["grad_norm"] - float(z["grad_norm"])) <= parity.TOL * float(z["grad_norm"])
    assert abs(out["learning_rate"] - float(z["learning_rate"])) < 1e-9
    for k in ("vs", "clipped_rho", "vs_plus_1", "pg_advantage"):
        assert parity.rel_err(taps[k], z[k]) < parity.TOL, k
    gd = it.unflatten_params(g, torch.float64, num_action=A)
    pd = it.unflatten_params(p_after, torch.float64, num_action=A)
    md = it.unflatten_params(ms_after, torch.float64, num_action=A)
    # ReLU kinks (see tests/parity.py): the float64 golden and a float32 forward may disagree on the mask of a
    # pre-activation that is ~0; one flip moves that image's conv gradients by O(1e-3).  The oracle-at-GPU-pattern
    # comparison (compare_step) holds the 1e-4 bar everywhere; against a FIXED golden the layers below a ReLU get
    # ``kink_tol`` (1e-4 where no flip occurs: the small case; 2e-3 for the 80-image BASELINE configs[0] case).
    for n in gd:
        tol = parity.TOL if n.split(".")[0] in ("actor3", "critic3") else kink_tol
        got = gd[n].numpy()
        if "grad_" + n in z:
            assert parity.rel_err(got, z["grad_" + n]) < tol, n
        else:       # big tensors are stored as a strided sample + l2 norm
            assert parity.rel_err(got.ravel()[::61], z["gradsample_" + n]) < tol, n
        # parameters after the update: judged against the size of the update itself
        p0 = params[n].double().numpy().ravel()[::61]
        du = z["paramsample_" + n] - p0
        assert np.max(np.abs(pd[n].numpy().ravel()[::61] - z["paramsample_" + n])) <= \
            10 * kink_tol * np.max(np.abs(du)) + 4 * np.finfo(np.float32).eps * np.max(np.abs(p0)), n
        assert parity.rel_err(md[n].numpy().ravel()[::61], z["rmssample_" + n]) < parity.TOL, n



def test_forward_windows_and_shift_identity(native):
    """build_network's first/middle/last outputs are slices of one forward (App. C.4) and match the oracle's
    REFERENCE-SHAPED graph (3 x (T-2) per-timestep network copies)."""
    B, T, A = 2, 6, 18
    batch, params, cfg = parity.make_case(B, T, A, seed=99)
    eng = parity.native_learner(batch, params, cfg)
    try:
        eng.stage(0, *[batch[k] for k in synthetic.TRAIN_FIELDS])
        pol, val = eng.forward(0)
    finally:
        eng.close()
    L = it.Learner(params, torch.float64, "reference", **cfg)
    o = L.losses(*[batch[k] for k in synthetic.TRAIN_FIELDS])
    assert parity.rel_err(pol[:, :-2], o["first_policy"].detach().numpy()) < parity.TOL
    assert parity.rel_err(pol[:, 1:-1], o["middle_policy"].detach().numpy()) < parity.TOL
    assert parity.rel_err(pol[:, 2:], o["last_policy"].detach().numpy()) < parity.TOL
    assert parity.rel_err(val[:, :-2], o["first_value"].detach().numpy()) < parity.TOL
    assert parity.rel_err(val[:, 2:], o["last_value"].detach().numpy()) < parity.TOL


def test_act_matches_single_step_network(native):
    """get_policy_and_action's forward (agent/impala.py:118-130) == network() on single frames."""
    n, A = 5, 18
    batch, params, cfg = parity.make_case(n, 3, A, seed=5)
    eng = parity.native_learner(batch, params, cfg)
    try:
        pol, h, c = eng.act(batch["state"][:, 0], batch["previous_action"][:, 0], batch["initial_h"][:, 0],
                            batch["initial_c"][:, 0])
    finally:
        eng.close()
    p64 = {k: v.double() for k, v in params.items()}
    x = torch.from_numpy((batch["state"][:, 0].astype(np.float64) / 255).astype(np.float32)).double()
    a, v, cc, hh = it.network(p64, x, torch.from_numpy(batch["previous_action"][:, 0].astype(np.int64)),
                              torch.from_numpy(batch["initial_h"][:, 0]).double(),
                              torch.from_numpy(batch["initial_c"][:, 0]).double(), A, 256)
    assert parity.rel_err(pol, a.numpy()) < parity.TOL
    assert parity.rel_err(h, hh.numpy()) < parity.TOL
    assert parity.rel_err(c, cc.numpy()) < parity.TOL


def test_data_parallel_shards_sum_to_full_batch(native):
    """sum over shards of local gradient buckets == gradient of the global batch (losses are
    batch SUMS), so one all-reduce(SUM) reproduces the reference update.  Emulated on one GPU: two B=2
    replicas, buckets added on the host, compared with a B=4 replica and with the oracle."""
    B, T, A = 4, 8, 18
    batch, params, cfg = parity.make_case(B, T, A, seed=21)
    fields = synthetic.TRAIN_FIELDS
    full = parity.native_learner(batch, params, cfg)
    try:
        full.stage(0, *[batch[k] for k in fields])
        full.forward_backward(0)
        g_full = full.get_grads()
    finally:
        full.close()
    g_sum = 0
    for lo, hi in ((0, 2), (2, 4)):
        sh = synthetic.slice_batch(batch, lo, hi)
        eng = parity.native_learner(sh, params, cfg)
        try:
            eng.stage(0, *[sh[k] for k in fields])
            eng.forward_backward(0)
            g_sum = g_sum + eng.get_grads().astype(np.float64)
        finally:
            eng.close()
    assert parity.rel_err(g_sum, g_full) < parity.TOL
    L = it.Learner(params, torch.float64, "dedup", **cfg)
    _, g = L.gradients(*[batch[k] for k in fields])
    # against the UNMASKED float64 oracle over the whole flat vector: one ReLU pre-activation within float32 rounding of
    # zero flips the mask and moves that image's conv gradients by O(1e-3) (tests/parity.py; compare_step evaluates the
    # oracle at the GPU's activation pattern and holds 1e-4) -- here the looser kink bar applies
    assert parity.rel_err(g_sum, it.flatten_grads(g)) < 2e-3


def test_errors_are_loud(native):
    from distributed_reinforcement_learning_b200 import _native as N
    from distributed_reinforcement_learning_b200.learner import NativeLearner
    with pytest.raises(N.DrlError):
        NativeLearner(batch=2, trajectory=20, input_shape=(64, 64, 4))       # unsupported geometry
    with pytest.raises(N.DrlError):
        NativeLearner(batch=2, trajectory=2)                                  # T-2 must be >= 1
    with pytest.raises(ValueError):
        NativeLearner(batch=2, reward_clipping="nope")
    eng = NativeLearner(batch=2, trajectory=5)
    try:
        with pytest.raises(N.DrlError):
            eng.step(0)                                                       # nothing staged
        with pytest.raises(ValueError):
            eng.set_params(np.zeros(10, np.float32))
        b = synthetic.make_batch(3, T=5)
        with pytest.raises(ValueError):
            eng.stage(0, *[b[k] for k in synthetic.TRAIN_FIELDS])            # wrong batch size
    finally:
        eng.close()


@pytest.mark.parametrize("graph", [False, True])
def test_two_steps_in_flight_per_slot_results(native, graph):
    """drl_learner_wait_slot: every staging slot has its own result record, so the host can enqueue step i+1 before
    it reads step i's scalars.  Results must equal the same two steps run one after the other."""
    B, T, A = 4, 6, 18
    batch, params, cfg = parity.make_case(B, T, A, seed=31)
    b2 = synthetic.make_batch(B, T=T, A=A, seed=32)
    seq = parity.native_learner(batch, params, cfg, use_cuda_graph=graph)
    seq.stage(0, *[batch[k] for k in synthetic.TRAIN_FIELDS])
    seq.stage(1, *[b2[k] for k in synthetic.TRAIN_FIELDS])
    r0, r1 = seq.step(0), seq.step(1)
    p_seq = seq.get_params()
    seq.close()
    eng = parity.native_learner(batch, params, cfg, use_cuda_graph=graph)
    eng.stage(0, *[batch[k] for k in synthetic.TRAIN_FIELDS])
    eng.stage(1, *[b2[k] for k in synthetic.TRAIN_FIELDS])
    eng.step_async(0)
    eng.step_async(1)            # second step enqueued before the first one's result is read
    q0, q1 = eng.wait(0), eng.wait(1)
    assert q0 == r0 and q1 == r1
    assert q1["step"] == q0["step"] + 1
    assert np.array_equal(eng.get_params(), p_seq)
    eng.close()
