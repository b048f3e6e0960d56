"""CPU, world_size 2 over gloo: the data-parallel contract of SURVEY.md section 8(e).  Each rank evaluates
its shard of the global batch (shard_range), the gradient bucket -- with the three loss sums in its tail,
exactly the layout the CUDA learner all-reduces -- is reduced with ONE all_reduce(SUM), and the result must
equal the gradient / losses of the undivided batch (the reference losses are batch SUMS, optimizer/vtrace.py:
112,118,126, so no rescale).  The per-rank compute here is the CPU oracle; on GPUs the same bucket comes from
drl_learner_forward_backward (tests/test_gpu_learner.py::test_data_parallel_shards_sum_to_full_batch)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import impala_torch as it
from oracle import synthetic

B, T = 2, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from distributed_reinforcement_learning_b200.learner import shard_range
    batch = synthetic.make_batch(B, T=T)
    lo, hi = shard_range(rank, world, B)
    sh = synthetic.slice_batch(batch, lo, hi)
    L = it.Learner(it.init_params(0), torch.float64, "dedup", trajectory=T)
    out, g = L.gradients(*[sh[k] for k in synthetic.TRAIN_FIELDS])
    bucket = torch.cat([torch.from_numpy(it.flatten_grads(g)),
                        torch.stack([out["pi_loss"], out["baseline_loss"], out["entropy"]]).detach().double(),
                        torch.zeros(1, dtype=torch.float64)])
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM)          # the one collective of the step
    if rank == 0:
        np.save(os.path.join(out_dir, "bucket.npy"), bucket.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_batch(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    bucket = np.load(tmp_path / "bucket.npy")
    batch = synthetic.make_batch(B, T=T)
    L = it.Learner(it.init_params(0), torch.float64, "dedup", trajectory=T)
    out, g = L.gradients(*[batch[k] for k in synthetic.TRAIN_FIELDS])
    full = it.flatten_grads(g)
    n = full.size
    assert bucket.size == n + 4
    assert np.max(np.abs(bucket[:n] - full)) <= 1e-9 * np.max(np.abs(full))
    for i, k in enumerate(("pi_loss", "baseline_loss", "entropy")):
        assert abs(bucket[n + i] - float(out[k].detach())) <= 1e-9 * abs(float(out[k].detach()))


# ---- mean-loss learners (Ape-X, A3C, R2D2): bucket SUM x 1/world (distributed_reinforcement_learning_b200/dp.py) -----
def _mean_loss_case(family):
    """-> (full batch dict, slicer(batch, lo, hi), grads_and_scalars(batch) -> (flat float64 grads, [loss scalars]))."""
    if family == "r2d2":
        from oracle import r2d2_torch as rt
        batch = rt.make_sequences(2, S=4, seed=3)

        def run(b):
            L = rt.Learner(dtype=torch.float64, seq_len=4, burn_in=1)
            out, g = L.gradients(*[b[k] for k in rt.TRAIN_FIELDS[:-1]], weight=b["weight"])
            return it.flatten_grads(g), [float(out["value_loss"].detach())]
        return batch, run
    from oracle import a3c_torch as at
    from oracle import apex_torch as ax
    batch = ax.make_transitions(2, seed=3)
    if family == "apex":
        def run(b):
            L = ax.Learner(dtype=torch.float64)
            out, g = L.gradients(*[b[k] for k in ax.TRAIN_FIELDS[:-1]], is_weight=b["is_weight"])
            return it.flatten_grads(g), [float(out["value_loss"].detach())]
        return batch, run

    def run(b):
        L = at.Learner(dtype=torch.float64)
        out = L.losses(*[b[k] for k in at.TRAIN_FIELDS])
        names = list(L.params)
        gr = torch.autograd.grad(out["total_loss"], [L.params[n] for n in names])
        return (np.concatenate([x.detach().double().reshape(-1).numpy() for x in gr]),
                [float(out[k].detach()) for k in ("pi_loss", "baseline_loss", "entropy")])
    return batch, run


def _mean_worker(rank, world, port, out_dir, family):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    batch, run = _mean_loss_case(family)
    n = len(batch["reward"]) // world
    sh = {k: v[rank * n:(rank + 1) * n] for k, v in batch.items()}
    g, scalars = run(sh)                                    # gradient of the rank's LOCAL mean loss
    bucket = torch.cat([torch.from_numpy(g), torch.tensor(scalars, dtype=torch.float64)])
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    bucket *= 1.0 / world                                   # grad_scale of drl_<family>_apply
    if rank == 0:
        np.save(os.path.join(out_dir, "bucket.npy"), bucket.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_mean_loss_learners_sum_times_inverse_world_equals_full_batch(tmp_path):
    for family in ("apex", "a3c", "r2d2"):
        d = tmp_path / family
        d.mkdir()
        mp.spawn(_mean_worker, args=(2, _free_port(), str(d), family), nprocs=2, join=True)
        bucket = np.load(d / "bucket.npy")
        batch, run = _mean_loss_case(family)
        full, scalars = run(batch)
        n = full.size
        assert bucket.size == n + len(scalars)
        assert np.max(np.abs(bucket[:n] - full)) <= 1e-9 * np.max(np.abs(full)), family
        assert np.allclose(bucket[n:], scalars, rtol=1e-9), family
