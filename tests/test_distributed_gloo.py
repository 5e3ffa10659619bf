"""world_size-2 gloo test (CPU) of the multi-GPU sampling plumbing: sharding, seeding, final gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nsamples, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffsci_b200 import distributed as D
    # a stand-in "sampler" (the kernels need a GPU): deterministic function of the white noise
    out = D.sample_sharded(lambda wn: wn * 2.0 + 1.0, nsamples, (1, 4, 4), seed=123, device="cpu")
    times = torch.tensor([float(rank + 1)])
    dist.all_reduce(times, op=dist.ReduceOp.MAX)          # bench.py's max-over-ranks timing reduction
    if rank == 0:
        q.put((out, float(times)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_sizes_and_ranges():
    from diffsci_b200 import distributed as D
    assert D.shard_sizes(10, 4) == [3, 3, 2, 2] and D.shard_sizes(8, 8) == [1] * 8 and D.shard_sizes(3, 4) == [1, 1, 1, 0]
    cover = [D.shard_range(10, 4, r) for r in range(4)]
    assert cover == [(0, 3), (3, 6), (6, 8), (8, 10)]
    a = torch.cat([D.white_noise_shard(10, (2,), 5, 4, r) for r in range(4)])
    g = torch.Generator().manual_seed(5)
    assert torch.equal(a, torch.randn(10, 2, generator=g))          # union over ranks == single-process draw


def test_two_rank_gather_matches_single_process():
    from diffsci_b200 import distributed as D
    nsamples = 5                                                     # ragged: ranks get 3 and 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nsamples, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = D.white_noise_shard(nsamples, (1, 4, 4), 123, 1, 0) * 2.0 + 1.0
    assert torch.equal(out, single) and tmax == 2.0


# ------------------------------------------------------------------------------------------------ training exchange
def test_grad_bucket_plan():
    from diffsci_b200.distributed import GradBucketer
    numels = [10, 20, 30, 40, 50]
    ready = [50, 40, 30, 20, 10]                       # the backward pass finishes the LAST parameter first
    b = GradBucketer.plan(numels, ready, bucket_elems=60)
    assert b == [(60, 150, 20), (0, 60, 50)]                                          # contiguous cover, cut from the end
    assert GradBucketer.plan(numels, ready, bucket_elems=45) == [(100, 150, 10), (30, 100, 30), (0, 30, 50)]
    assert GradBucketer.plan(numels, ready, bucket_elems=10 ** 9) == [(0, 150, 50)]


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffsci_b200.distributed import GradBucketer
    numels = [7, 5, 9, 3]
    flat = torch.arange(24, dtype=torch.float32) * (rank + 1)          # this rank's "gradients"
    bk = GradBucketer(flat, numels, ready_pos=[4, 3, 2, 1], bucket_bytes=8 * 4)
    hooks = bk.hooks()
    fired = []
    for pos in range(1, 5):                                           # stand-in for TrainGraph.run_backward
        if pos in hooks:
            hooks[pos]()
            fired.append(pos)
    scale = bk.finish()
    if rank == 0:
        q.put((flat * scale, fired, bk.buckets))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce():
    """world_size 2 over gloo: every bucket is all-reduced exactly once, as soon as it is ready, and sum * 1/world is the
    mean of the ranks' gradients (the host logic of EDMTrainer's data-parallel step)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mean, fired, buckets = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(mean, torch.arange(24, dtype=torch.float32) * 1.5)
    assert fired == sorted({r for _, _, r in buckets})
