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
