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
