"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel learner of SURVEY.md section 8(e) for real --
one process per GPU, torch.distributed/NCCL, each rank feeds its shard, ONE all_reduce(SUM) of the gradient
bucket (with the loss sums in its tail) between drl_learner_forward_backward and drl_learner_apply.  The
reduced gradient, the logged losses and the updated parameters of every rank must equal those of a single
replica stepping on the undivided batch."""
import os
import socket

import numpy as np
import pytest
import torch

import parity
from oracle import impala_torch as it
from oracle import synthetic

pytestmark = pytest.mark.gpu
B, T, A = 4, 6, 18


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, peer=False, Bt=B):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from distributed_reinforcement_learning_b200.learner import shard_range
    batch, params, cfg = parity.make_case(Bt, T, A, seed=77)
    lo, hi = shard_range(rank, world, Bt)
    sh = synthetic.slice_batch(batch, lo, hi)
    eng = parity.native_learner(sh, params, cfg, device=rank)
    eng.stage(0, *[sh[k] for k in synthetic.TRAIN_FIELDS])
    if peer:
        eng.enable_peer_exchange()          # the exchange becomes part of the step (csrc/peer.cu)
    out = eng.step(0)                       # forward_backward -> all_reduce(SUM) -> apply
    grads, params1 = eng.get_grads(), eng.get_params()
    if peer:                                # the fused exchange's result vs an NCCL all-reduce of the saved local buckets
        torch.cuda.synchronize()
        local, reduced = eng.bucket_tensor().clone(), eng.reduced_tensor().clone()
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        err = float((local - reduced).abs().max()) / max(float(local.abs().max()), 1e-30)
        assert err <= 1e-6, "rank %d: fused exchange differs from the NCCL all-reduce of the local buckets: %g" % (rank, err)
    out2 = eng.step(0)                      # a second step: barrier epochs, buffer reuse
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), grads=grads, params=params1, params2=eng.get_params(),
             losses=np.array([out["pi_loss"], out["baseline_loss"], out["entropy"], out["grad_norm"]]),
             losses2=np.array([out2["pi_loss"], out2["baseline_loss"], out2["entropy"], out2["grad_norm"]]))
    dist.barrier()                          # nobody unmaps / frees while a peer may still read
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_allreduce_step_equals_single_replica(native, tmp_path):
    if native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # replicas stay identical
    assert np.array_equal(r0["params"], r1["params"]) and np.array_equal(r0["grads"], r1["grads"])
    batch, params, cfg = parity.make_case(B, T, A, seed=77)
    eng = parity.native_learner(batch, params, cfg)
    eng.stage(0, *[batch[k] for k in synthetic.TRAIN_FIELDS])
    out = eng.step(0)
    g1, p1 = eng.get_grads(), eng.get_params()
    eng.close()
    assert parity.rel_err(r0["grads"], g1) < parity.TOL
    ref = np.array([out["pi_loss"], out["baseline_loss"], out["entropy"], out["grad_norm"]])
    assert np.all(np.abs(r0["losses"] - ref) <= parity.TOL * np.abs(ref))
    p0 = it.flatten_params(params)
    assert np.max(np.abs((r0["params"] - p0) - (p1 - p0))) <= 1e-3 * np.max(np.abs(p1 - p0))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_peer_exchange_equals_nccl_path(native, tmp_path, world):
    """The exchange as kernels over NVLink peer memory (CUDA IPC; csrc/peer.cu) inside the step graph: all ranks end
    with bit-identical replicas, and gradients / losses / parameters equal the NCCL all-reduce path over two steps."""
    if native.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    d_nccl, d_peer = tmp_path / "nccl", tmp_path / "peer"
    d_nccl.mkdir()
    d_peer.mkdir()
    Bt = max(B, world)                                                # one trajectory per rank at world 8
    mp.spawn(_worker, args=(world, _free_port(), str(d_nccl), False, Bt), nprocs=world, join=True)
    mp.spawn(_worker, args=(world, _free_port(), str(d_peer), True, Bt), nprocs=world, join=True)
    n0 = np.load(d_nccl / "rank0.npz")
    p0 = np.load(d_peer / "rank0.npz")
    for r in range(1, world):
        pr = np.load(d_peer / ("rank%d.npz" % r))
        for k in ("grads", "params", "params2", "losses", "losses2"):
            assert np.array_equal(p0[k], pr[k]), (r, k)               # replicas stay bit-identical
    if world == 2:
        assert np.array_equal(p0["grads"], n0["grads"])               # a + b in rank order, like the 2-rank all-reduce
    else:
        assert parity.rel_err(p0["grads"], n0["grads"]) < 1e-6        # NCCL's summation order differs for W > 2
    assert np.allclose(p0["losses"], n0["losses"], rtol=1e-6) and np.allclose(p0["losses2"], n0["losses2"], rtol=1e-5)
    upd = np.max(np.abs(n0["params2"] - it.flatten_params(parity.make_case(Bt, T, A, seed=77)[1])))
    assert np.max(np.abs(p0["params2"] - n0["params2"])) <= 1e-4 * upd


# ---- data-parallel Ape-X / A3C / R2D2 (distributed_reinforcement_learning_b200/dp.py) -------------------------------
def _family_case(family):
    """-> (make engine(B, device), batch dict, stage(eng, batch), per-sample slicer)."""
    if family == "r2d2":
        from oracle import r2d2_torch as rt
        from distributed_reinforcement_learning_b200.r2d2_learner import NativeR2D2Learner
        kwp = dict(num_action=4, lstm_size=64, input_shape=(84, 84, 1))
        pm = rt.flatten_params(rt.init_params(0, torch.float32, **kwp))
        pt = rt.flatten_params(rt.init_params(1, torch.float32, **kwp))
        batch = rt.make_sequences(4, S=6, seed=31)

        def make(Bn, dev):
            e = NativeR2D2Learner(batch=Bn, seq_len=6, burn_in=2, device=dev)
            e.set_params(pm, 0)
            e.set_params(pt, 1)
            return e

        def stage(e, b):
            e.stage(0, b["state"], b["previous_action"], b["action"], b["h"][:, 0], b["c"][:, 0], b["reward"], b["done"],
                    b["weight"])
        return make, batch, stage
    from oracle import apex_torch as ax
    pm = ax.flatten_params(ax.init_params(0, torch.float32, num_action=4))
    pt = ax.flatten_params(ax.init_params(1, torch.float32, num_action=4))
    batch = ax.make_transitions(4, seed=31)
    if family == "apex":
        from distributed_reinforcement_learning_b200.apex_learner import NativeApexLearner

        def make(Bn, dev):
            e = NativeApexLearner(batch=Bn, num_action=4, device=dev)
            e.set_params(pm, 0)
            e.set_params(pt, 1)
            return e

        def stage(e, b):
            e.stage(0, *[b[k] for k in ax.TRAIN_FIELDS])
        return make, batch, stage
    from distributed_reinforcement_learning_b200.a3c_learner import NativeA3CLearner

    def make(Bn, dev):
        e = NativeA3CLearner(batch=Bn, num_action=4, device=dev)
        e.set_params(pm)
        return e

    def stage(e, b):
        e.stage(0, *[b[k] for k in ax.TRAIN_FIELDS[:-1]])
    return make, batch, stage


def _scalars(out):
    out = out[0] if isinstance(out, tuple) else out
    return np.array([out[k] for k in sorted(out) if k not in ("step", "learning_rate")], np.float64)


def _family_worker(rank, world, port, out_dir, family):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    make, batch, stage = _family_case(family)
    n = len(batch["reward"]) // world
    sh = {k: v[rank * n:(rank + 1) * n] for k, v in batch.items()}
    eng = make(n, rank)
    stage(eng, sh)
    out = eng.step(0)                      # forward_backward -> all_reduce(SUM) -> apply(1 / world)
    st = eng.get_opt_state()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), grads=eng.get_grads(), m=st["m"], v=st["v"],
             params=eng.get_params(), scalars=_scalars(out))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("family", ["apex", "a3c", "r2d2"])
def test_two_gpu_mean_loss_learners_equal_single_replica(native, tmp_path, family):
    """Two ranks with half of the minibatch each (bucket SUM x 1/2, their losses are batch means) reproduce the
    gradient, the logged scalars and the Adam slots of one replica stepping on the undivided minibatch."""
    if native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_family_worker, args=(2, _free_port(), str(tmp_path), family), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("m", "v", "params"):
        assert np.array_equal(r0[k], r1[k]), k                       # replicas stay identical
    make, batch, stage = _family_case(family)
    eng = make(len(batch["reward"]), 0)
    p0 = eng.get_params()
    stage(eng, batch)
    out = eng.step(0)
    st = eng.get_opt_state()
    g1, p1 = eng.get_grads(), eng.get_params()
    eng.close()
    # the bucket holds the SUM of the two half-batch mean gradients: x 1/2 = the full-batch mean gradient
    assert parity.rel_err(0.5 * r0["grads"], g1) < parity.TOL
    assert parity.rel_err(r0["m"], st["m"]) < parity.TOL and parity.rel_err(r0["v"], st["v"]) < 2 * parity.TOL
    assert np.allclose(r0["scalars"], _scalars(out), rtol=1e-4)
    # Adam's first step is ~ lr * sign(g): elements with |g| near 1e-8 are ill-conditioned, so compare the bulk
    d = np.abs((r0["params"] - p0) - (p1 - p0))
    assert np.quantile(d, 0.999) <= 1e-2 * np.max(np.abs(p1 - p0)) and np.max(d) <= 2.1 * np.max(np.abs(p1 - p0))
