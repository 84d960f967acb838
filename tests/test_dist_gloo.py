"""CPU, world_size 2 over gloo: the multi-rank plumbing of the benchmark (disjoint scene shards,
max-over-ranks timing, whole-job throughput) without a GPU."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ws3d_b200 import sharding, synth
    lo, hi = sharding.scene_range(rank, world, 2)
    scenes = synth.make_batch(hi - lo, 256, first_scene=lo)
    # every rank hashes its shard; rank 0 gathers to prove the shards are disjoint and complete
    digest = torch.tensor([float(scenes.sum()), float(lo), float(hi)], dtype=torch.float64)
    gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, digest)
    ms = sharding.max_over_ranks(10.0 + 5.0 * rank)            # slowest rank defines the step time
    if rank == 0:
        out.put(([g.tolist() for g in gathered], ms, sharding.aggregate_throughput(2 * 256, world, ms)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered, ms, thr = out.get()
    from ws3d_b200 import synth
    assert [g[1:] for g in gathered] == [[0.0, 2.0], [2.0, 4.0]]
    want = [float(synth.make_batch(2, 256, first_scene=s).sum()) for s in (0, 2)]
    np.testing.assert_allclose([g[0] for g in gathered], want)
    assert ms == 15.0 and thr == 2 * 256 * 2 / 0.015


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ws3d_b200 import sharding
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    grads = sharding.FlatGradients(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(100 + rank)          # every rank has its own shard of the batch
    local = []
    for step in range(2):
        x, y = torch.randn(4, 5, generator=g), torch.randn(4, 3, generator=g)
        grads.zero()
        ((net(x) - y) ** 2).mean().backward()
        assert grads.attached()                            # autograd accumulated INTO the flat buffer's views
        local.append(grads.flat.clone())
        grads.exchange()
        if step == 0:
            mean0 = grads.flat.clone()
        opt.step()
    gathered = [torch.zeros_like(local[0]) for _ in range(world)]
    dist.all_gather(gathered, local[0])
    params = torch.cat([p.detach().flatten() for p in net.parameters()])
    everyone = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(everyone, params)
    if rank == 0:
        out.put((torch.stack(gathered).mean(0).tolist(), mean0.tolist(), [e.tolist() for e in everyone], grads.nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_exchange():
    """The training path's one collective: per-rank gradients land in one flat buffer, one all-reduce averages them, and
    replicas that start identical stay identical after the optimiser steps."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want_mean, got_mean, params, nbytes = out.get()
    np.testing.assert_allclose(got_mean, want_mean, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(params[0], params[1])
    assert nbytes == 4 * (5 * 7 + 7 + 7 * 3 + 3)
