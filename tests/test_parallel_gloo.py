"""N>1 host logic on the CPU: two gloo ranks shard prompts, all-reduce decoder gradients through one flat buffer and
agree on max-over-ranks timing (the GPU path uses the same code with the nccl backend)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from triplaneturbo_b200.parallel import allreduce_gradients, local_batch, max_over_ranks, shard_prompts


def test_shard_prompts_partitions_every_prompt_once():
    for world in (1, 2, 4, 8):
        seen = sorted(p for r in range(world) for p in shard_prompts(32, r, world))
        assert seen == list(range(32))
        assert all(len(shard_prompts(32, r, world)) == 32 // world for r in range(world))
    assert shard_prompts(3, 1, 2) == [1] and shard_prompts(3, 0, 2) == [0, 2]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, V = 4, 2
        g = torch.Generator().manual_seed(0)
        sc = torch.randn(P, 6, 2, 3, 3, generator=g)
        rays = torch.arange(P * V * 3, dtype=torch.float32).view(P * V, 3)
        sc_l, (rays_l,) = local_batch(sc, [rays], V, rank, world)
        mine = shard_prompts(P, rank, world)
        assert torch.equal(sc_l, sc[mine])
        assert torch.equal(rays_l, torch.cat([rays[p * V:(p + 1) * V] for p in mine]))
        # per-rank "gradients": rank-dependent so that the reduction is observable
        grads = [torch.full((64, 8), float(rank + 1)), torch.full((64, 64), float(10 * (rank + 1))), torch.ones(1, 64) * rank]
        red = allreduce_gradients(grads, average=False)
        tot = sum(range(1, world + 1))
        assert torch.equal(red[0], torch.full((64, 8), float(tot)))
        assert torch.equal(red[1], torch.full((64, 64), float(10 * tot)))
        assert torch.equal(red[2], torch.ones(1, 64) * sum(range(world)))
        avg = allreduce_gradients(grads, average=True)
        assert torch.allclose(avg[0], torch.full((64, 8), tot / world))
        assert max_over_ranks(1.0 + rank, "cpu") == float(world)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
