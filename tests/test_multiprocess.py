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


def _shard_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location('pps_sharding', os.path.join(ROOT, 'ppsurf_b200', 'sharding.py'))
    sharding = importlib.util.module_from_spec(spec)  # loaded by path: the package itself needs the CUDA library
    spec.loader.exec_module(sharding)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    shard = sharding.Shard.from_env()
    assert (shard.world, shard.rank) == (world, rank)
    # occupancy queries: every rank ends with the values of ALL queries, ragged and empty lists included
    calls = []

    def fn(q):
        calls.append(q.shape[0])
        return (q[:, 0] * 2.0 + q[:, 1] - q[:, 2]).float()

    for n in (0, 1, 5, 1000, 1001):
        q = torch.arange(3 * n, dtype=torch.float32).view(n, 3) * 0.01
        got = shard.evaluate(fn, q)
        assert torch.equal(got, fn(q)), n
    assert sum(calls[:5:1]) >= 0
    first, count = shard.slice_of(1001)
    np.save(os.path.join(out_dir, 'slice_{}.npy'.format(rank)), np.array([first, count]))
    # latent loop: passes dealt p mod G, partial sums all-reduced == one rank walking every pass (up to fp32 summation order)
    rng = np.random.default_rng(0)
    n_pts, c, sub = 500, 8, 120
    passes = [rng.permutation(n_pts)[:sub] for _ in range(11)]

    def encode(p, ids):  # a stand-in for the encoder: depends on the pass and the point
        return torch.from_numpy(np.outer(np.cos(ids + p), np.arange(1, c + 1)).astype(np.float32))

    latent, counts = torch.zeros(n_pts, c), torch.zeros(n_pts)
    for p, ids in shard.my_passes(enumerate(passes)):
        latent[ids] += encode(p, ids)
        counts[ids] += 1
    shard.reduce_latents(latent, counts)
    ref_l, ref_c = torch.zeros(n_pts, c), torch.zeros(n_pts)
    for p, ids in enumerate(passes):
        ref_l[ids] += encode(p, ids)
        ref_c[ids] += 1
    assert torch.equal(counts, ref_c) and (latent - ref_l).abs().max() < 1e-5
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_shard_collectives(tmp_path, world):
    """ppsurf_b200.sharding.Shard on gloo / CPU: sliced evaluation with the all-gather, dealt encoder passes with the all-reduce"""
    port = 29000 + (os.getpid() * 7 + world) % 900
    mp.spawn(_shard_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    slices = sorted(tuple(np.load(tmp_path / 'slice_{}.npy'.format(r))) for r in range(world))
    assert slices[0][0] == 0 and slices[-1][0] + slices[-1][1] == 1001
    assert all(a[0] + a[1] == b[0] for a, b in zip(slices, slices[1:]))


def _grad_bucket_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from ppsurf_b200 import training  # host logic only: nothing here launches a kernel
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 2))
    flat = training.flatten_gradients(net.parameters())
    params = list(net.parameters())
    assert flat.numel() == sum(p.numel() for p in params)
    assert all(p.grad.data_ptr() >= flat.data_ptr() and p.grad.shape == p.shape for p in params)
    # autograd accumulates INTO the views: the flat buffer sees the gradients
    x = torch.randn(11, 5, generator=torch.Generator().manual_seed(100 + rank))
    net(x).square().sum().backward()
    assert torch.equal(flat, torch.cat([p.grad.reshape(-1) for p in params])) and float(flat.abs().sum()) > 0
    local = flat.clone()
    training.average_gradients(flat, world)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat, torch.stack(gathered).mean(0), rtol=1e-6, atol=1e-7)
    assert torch.equal(params[0].grad.reshape(-1), flat[:params[0].numel()])  # the parameters see the averaged gradient
    if rank == 0:
        np.save(os.path.join(out_dir, 'bucket.npy'), flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2])
def test_training_gradient_bucket_gloo(tmp_path, world):
    """data-parallel fit (config 5): every gradient is a view into one flat buffer, one all-reduce averages it over the ranks
    (training.GraphedTrainStep's collective) -- host logic on gloo / CPU"""
    port = 29000 + (os.getpid() * 11 + world) % 900
    mp.spawn(_grad_bucket_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / 'bucket.npy').size == 5 * 7 + 7 + 7 + 7 + 7 * 2 + 2


def _nccl_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import ppsurf_b200
    from ppsurf_b200.sharding import Shard
    from oracle import ppsurf_oracle as oracle
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'x.txt', 'results', 0.05, 't', 256, 3, 5000, 33, 50,
                                    50000, 0, 0)
    model.network.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in oracle.make_state_dict(42).items()}, strict=True)
    model = model.to(dev)
    model.network.sampling_seed = 7
    pts = oracle.synthetic_cloud(12000, seed=9)
    pts_ms = torch.from_numpy(pts[None]).to(dev)

    def run(shard):
        model.shard = shard
        pts_bcn = pts_ms.transpose(1, 2).contiguous()
        latents = model.encode_cloud(pts_bcn, generator=torch.Generator().manual_seed(5))
        dec = model.network.decoder_for(pts_bcn, latents)
        vol = model.create_volume_device(dec, pts, 33)
        return latents.cpu().numpy(), vol.cpu().numpy()

    lat_n, vol_n = run(Shard.from_env())
    # every rank holds the same result
    both = [torch.zeros(vol_n.shape, device=dev) for _ in range(world)]
    dist.all_gather(both, torch.from_numpy(np.nan_to_num(vol_n, nan=-7.0)).to(dev))
    assert all(torch.equal(both[0], b) for b in both)
    if rank == 0:
        lat_1, vol_1 = run(Shard())  # the same cloud on one rank: only the fp32 summation order of the latent average differs
        scale = np.abs(lat_1).max()
        assert np.abs(lat_n - lat_1).max() < 2e-5 * scale
        same_mask = np.isnan(vol_n) == np.isnan(vol_1)
        assert same_mask.mean() > 0.999
        ok = ~np.isnan(vol_n) & ~np.isnan(vol_1)
        assert np.abs(vol_n[ok] - vol_1[ok]).max() < 1e-3
        np.save(os.path.join(out_dir, 'done.npy'), np.array([ok.sum()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_predict_on_two_gpus(tmp_path):
    """2 NCCL ranks: encoder passes dealt + all-reduce, every sweep's queries sliced + all-gather == the one-rank volume"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    port = 28000 + os.getpid() % 900
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert int(np.load(tmp_path / 'done.npy')[0]) > 1000
