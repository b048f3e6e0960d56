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


def _worker(rank, world, port, out_dir, peer=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from distributed_reinforcement_learning_b200.learner import shard_range
    batch, params, cfg = parity.make_case(B, T, A, seed=77)
    lo, hi = shard_range(rank, world, B)
    sh = synthetic.slice_batch(batch, lo, hi)
    eng = parity.native_learner(sh, params, cfg, device=rank)
    eng.stage(0, *[sh[k] for k in synthetic.TRAIN_FIELDS])
    if peer:
        eng.enable_peer_exchange()          # the exchange becomes part of the step (csrc/peer.cu)
    out = eng.step(0)                       # forward_backward -> all_reduce(SUM) -> apply
    grads, params1 = eng.get_grads(), eng.get_params()
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


@pytest.mark.parametrize("world", [2, 4])
def test_fused_peer_exchange_equals_nccl_path(native, tmp_path, world):
    """The exchange as kernels over NVLink peer memory (CUDA IPC; csrc/peer.cu) inside the step graph: all ranks end
    with bit-identical replicas, and gradients / losses / parameters equal the NCCL all-reduce path over two steps."""
    if native.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    d_nccl, d_peer = tmp_path / "nccl", tmp_path / "peer"
    d_nccl.mkdir()
    d_peer.mkdir()
    mp.spawn(_worker, args=(world, _free_port(), str(d_nccl), False), nprocs=world, join=True)
    mp.spawn(_worker, args=(world, _free_port(), str(d_peer), True), nprocs=world, join=True)
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
    upd = np.max(np.abs(n0["params2"] - it.flatten_params(parity.make_case(B, T, A, seed=77)[1])))
    assert np.max(np.abs(p0["params2"] - n0["params2"])) <= 1e-4 * upd
