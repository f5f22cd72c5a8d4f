"""world_size-2 gloo test (CPU) of the multi-GPU host logic in bench.py: contiguous grid slabs that tile the volume
exactly, the single broadcast of the latent table from rank 0, and max-over-ranks timing reduction."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group('gloo', rank=rank, world_size=world)
    total = 131 ** 3
    first, count = bench.grid_shard(total, world, rank)
    # rank 0 "encodes", everybody receives the same latent table in one broadcast
    latents = torch.full((1, 8, 1000), float(rank + 1)) if rank != 0 else torch.arange(8000, dtype=torch.float32).view(1, 8, 1000)
    dist.broadcast(latents, src=0)
    assert torch.equal(latents, torch.arange(8000, dtype=torch.float32).view(1, 8, 1000))
    # every rank decodes its slab (here: a stand-in function of the vertex index) and the slabs are gathered
    mine = torch.arange(first, first + count, dtype=torch.float64) * 0.5
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([count]))
    parts = [torch.zeros(int(s), dtype=torch.float64) for s in sizes]
    if rank == 0:
        dist.gather(mine, parts, dst=0) if len({int(s) for s in sizes}) == 1 else None
    elif len({int(s) for s in sizes}) == 1:
        dist.gather(mine, dst=0)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == 10.0 + world - 1
    np.save(os.path.join(out_dir, 'shard_{}.npy'.format(rank)), np.array([first, count]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2])
def test_grid_slabs_and_broadcast(tmp_path, world):
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    shards = sorted((tuple(np.load(tmp_path / 'shard_{}.npy'.format(r))) for r in range(world)))
    assert shards[0][0] == 0
    for (f0, c0), (f1, _) in zip(shards, shards[1:]):
        assert f0 + c0 == f1
    assert shards[-1][0] + shards[-1][1] == 131 ** 3


def test_grid_shard_covers_every_vertex_once():
    sys.path.insert(0, ROOT)
    import bench
    for total in (19 ** 3, 131 ** 3, 259 ** 3, 7):
        for world in (1, 2, 3, 4, 8):
            spans = [bench.grid_shard(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            assert all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_grid_blocks_partition_and_balance():
    """the dealt block split used by bench.py: every vertex exactly once, rank loads within one block of each other"""
    sys.path.insert(0, ROOT)
    import bench
    for total in (19 ** 3, 131 ** 3, 259 ** 3, 7):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(total, dtype=np.int32)
            loads = []
            for r in range(world):
                spans = bench.grid_blocks(total, world, r)
                loads.append(sum(c for _, c in spans))
                for f, c in spans:
                    assert c > 0 and f + c <= total
                    seen[f:f + c] += 1
            assert np.all(seen == 1)
            assert max(loads) - min(loads) <= bench.GRID_BLOCK
