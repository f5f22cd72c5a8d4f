#!/usr/bin/env python
"""Benchmark of the PPSurf occupancy decode (BASELINE.json metric): Mquery-pts/s over the dense (129+2)^3 marching-cubes
grid of a 100k-point synthetic cloud, ppsurf_50nn, latents and points resident on the device.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--path 0|1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One step = one pass of the hot path (kNN k=64 -> patches -> both branches -> MLP -> softmax difference) over this rank's
share of the grid.  N>1: the vertex list is dealt to the N ranks in blocks (strong scaling of the fixed 131^3 volume), rank 0
encodes the cloud and the latents travel in ONE NCCL broadcast before the timed region; the decode itself needs no
collective.  Prints one JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference (oracle/,
torch-CPU ops + scipy kd-tree, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'Mquery-pts/sec occupancy decode, ppsurf_50nn 129^3 grid'
UNIT = 'Mquery/s'
# algorithmic work of the reference formulation per query and of the formulation actually executed
FLOP_PER_QUERY_REFERENCE = 53.24e6
GEMM_FLOP_PER_ROW = 2.0 * (256 * 256 * 2 + 256 * 64)  # fc2 + fc3 + fc_query on one (query, neighbour) row


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'reference-gpu'])
    ap.add_argument('--path', type=int, default=int(os.environ.get('PPS_DECODE_PATH', '1')),
                    help='1 = tcgen05 split-fp16 kernels (default), 0 = fp32 SIMT kernels')
    ap.add_argument('--points', type=int, default=100000)
    ap.add_argument('--resolution', type=int, default=129)
    ap.add_argument('--num-pts-local', type=int, default=50)
    ap.add_argument('--chunk', type=int, default=151552,
                    help='queries per launch of the per-chunk kernels; a multiple of 37888 = 2 x 148 SMs x 128 rows (whole waves for every '
                         'tensor-core kernel); ppsurf_b200.ops.DEFAULT_CHUNK')
    ap.add_argument('--latents', default='encoder', choices=['encoder', 'random'])
    ap.add_argument('--encoder', default='sharded', choices=['sharded', 'rank0'], help='multi-GPU: who runs the latent loop')
    ap.add_argument('--cpu-sample', type=int, default=65536,
                    help='queries of the CPU baseline sample (about 10-15 s of host work on 16 cores)')
    ap.add_argument('--ref-sample', type=int, default=16384, help='queries per step of the --impl reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-gpu', action='store_true', help='skip the reference-modules-on-this-B200 leg of the b200 arm')
    ap.add_argument('--no-predict', action='store_true', help='skip the e2e_predict block (encoder + region-grown shell)')
    ap.add_argument('--refgpu-batches', type=int, default=2, help='50 000-query batches the reference-on-GPU leg times')
    ap.add_argument('--workload', default='decode', choices=['decode', 'fit'],
                    help="decode = BASELINE metric (default); fit = BASELINE config 5: one training step (bf16, batch 16 over 8 GPUs)")
    ap.add_argument('--fit-clouds-per-gpu', type=int, default=2, help='config 5: 16 clouds over 8 GPUs (weak scaling: fixed per GPU)')
    ap.add_argument('--fit-points', type=int, default=10000, help='manifold points per cloud (configs/poco.yaml:34)')
    ap.add_argument('--fit-queries', type=int, default=2000, help='query points per cloud (as in abc_minimal)')
    ap.add_argument('--fit-precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--fit-eager', action='store_true', help='config 5: eager launches + torch DDP instead of the captured CUDA graphs')
    ap.add_argument('--no-fit', action='store_true', help='skip the fit block (config 5 training step) of the default line')
    ap.add_argument('--profile-run', action='store_true', help='for ncu captures only: no minimum warm-up, no e2e leg')
    ap.add_argument('--shard-of', type=int, default=0,
                    help='debug: decode only the share rank 0 would get in a job of this many ranks (single process)')
    return ap.parse_args()


class QuietStdout:
    """Everything any library writes to file descriptor 1 while the bench runs (NCCL prints its version banner there) goes to
    stderr; `emit` writes the ONE JSON line of the contract to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {'bf16_tflops': d.get('bf16_tflops_sustained', d.get('bf16_tflops')), 'hbm_gbs': d.get('hbm_gbs'),
                'source': 'measured (MEASURED_PEAKS.json, sustained bf16)'}
    return {'bf16_tflops': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """samples nvidia-smi (every 100 ms) from before the warm-up to the end of the timed region; `mark()` is called when the timed
    region starts.  The summary uses the samples taken after the mark; a timed region shorter than the sampling period (many
    GPUs, few steps) falls back to the samples of the warm-up steps, which run the identical workload"""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.t_mark = None

    def __enter__(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def mark(self):
        self.t_mark = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        timed = [r for t, r in self.rows if self.t_mark is not None and t >= self.t_mark]
        rows = timed if timed else [r for _, r in self.rows[1:]] or [r for _, r in self.rows]  # [0] may predate the load
        sm = [float(r[0]) for r in rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == 'active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm), 'window': 'timed region' if timed else 'warm-up + timed region'}


def grid_shard(total, world, rank):
    """contiguous slab [first, first+count) of the flattened C-order vertex list"""
    first = total * rank // world
    return first, total * (rank + 1) // world - first


GRID_BLOCK = 4736  # vertices per dealt block (= 37 x 128): small enough to balance 8 ranks within 2 %, a multiple of the kNN run


def grid_blocks(total, world, rank, block=GRID_BLOCK):
    """blocks of `block` consecutive vertices of the flattened C-order list, dealt round-robin: rank r decodes blocks r, r+G,
    ...  Contiguous slabs are not balanced -- the neighbour search costs up to 10x more per vertex in the middle of the volume
    than near the surface -- a dealt split gives every rank the same mix.  Returns [(first, count), ...]; one span for G = 1"""
    if world == 1:
        return [(0, total)]
    spans = []
    for b in range(rank, (total + block - 1) // block, world):
        first = b * block
        spans.append((first, min(block, total - first)))
    return spans


def workload_name(args):
    r = args.resolution + 2
    return 'ppsurf_{}nn predict decode, {}k-pt synthetic sphere cloud, dense {}^3={} grid vertices (gen_resolution_global={})'.format(
        args.num_pts_local, args.points // 1000, r, r ** 3, args.resolution)


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle = checker / baseline only)
# ---------------------------------------------------------------------------------------------------------------------

REFERENCE_BATCH = 50000  # rec_batch_size of configs/poco.yaml:52: the reference rebuilds its kd-tree once per such batch
CPU_SUB_BATCH = 8192     # the network part runs in sub-batches so that the [Q,64,259] intermediates stay below ~1 GB


def cpu_decode(oracle, weights, pts, latents_cn, queries, num_pts_local, idx_given=None):
    """the reference's per-batch work on the host cores: kd-tree build + k=64 / k=P queries, patch normalisation, both
    branches, MLP, softmax difference (reference formulation, torch CPU ops on all threads).  `idx_given` [Q,kmax] replaces
    the kd-tree search (checker use: same neighbours as the device, so that only the network arithmetic is compared)"""
    import torch
    from oracle import ppsurf_oracle_torch as oracle_torch
    kmax = max(64, num_pts_local)
    pts_t = torch.from_numpy(pts)
    lat_t = torch.from_numpy(np.ascontiguousarray(latents_cn.T))
    out = np.empty((queries.shape[0],), dtype=np.float32)
    for b0 in range(0, queries.shape[0], REFERENCE_BATCH):
        qb = queries[b0:b0 + REFERENCE_BATCH]
        if idx_given is None:
            idx, _ = oracle.knn_kdtree(pts, qb, kmax)  # builds the tree, like source/base/proximity.py:84-89 per batch
        else:
            idx = idx_given[b0:b0 + REFERENCE_BATCH].astype(np.int64)
        loc = oracle.normalize_patches(pts[idx[:, :num_pts_local]], qb)
        for s0 in range(0, qb.shape[0], CPU_SUB_BATCH):
            sl = slice(s0, s0 + CPU_SUB_BATCH)
            occ, _ = oracle_torch.from_latent(weights, pts_t, lat_t, torch.from_numpy(qb[sl]),
                                              torch.from_numpy(idx[sl, :64].copy()), torch.from_numpy(loc[sl]))
            out[b0 + s0:b0 + s0 + occ.shape[0]] = occ.numpy()
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline must use every host core regardless of the launcher
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import ppsurf_oracle as oracle
    weights = oracle.make_state_dict(42)
    pts = oracle.synthetic_cloud(args.points, 42)
    rng = np.random.default_rng(7)
    latents = rng.standard_normal((256, args.points)).astype(np.float32)
    grid = oracle.dense_grid_queries(pts, args.resolution, 1)
    sample = args.ref_sample
    times = []
    for s in range(args.warmup + args.steps):
        q = grid[rng.choice(grid.shape[0], sample, replace=False)]
        t0 = time.perf_counter()
        cpu_decode(oracle, weights, pts, latents, q, args.num_pts_local)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3) / 1e6
    cores = torch.get_num_threads()
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'step': 'bounded sample of {} grid vertices per step'.format(sample)},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '{} random vertices of the same grid per step, torch-CPU oracle + scipy cKDTree, torch threads {} '
                                   '(set explicitly; host cores {})'.format(sample, cores, os.cpu_count())},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def reference_gpu_decode(oracle, weights_dev, pts, pts_dev, lat_dev, queries, num_pts_local, tree=None):
    """one `rec_batch_size` batch the way the reference's GPU predict path runs it: CPU kd-tree
    k=P query on the tree built once per cloud + numpy patch normalisation + H2D (source/poco_utils.py:63-72), CPU kd-tree
    REBUILD + k=64 query + H2D (source/poco_model.py:385, source/base/proximity.py:84-89), the network's torch-eager ops on the
    GPU under TF32 'high' (source/cli.py:88) and fp16 autocast (configs/poco.yaml:10 precision 16-mixed), softmax difference,
    blocking D2H (poco_utils.py:79-81).  The torch ops are the oracle's torch twin of the reference modules moved to the
    device; /root/reference itself does not exist on the GPU box."""
    import torch
    from scipy.spatial import cKDTree
    from oracle import ppsurf_oracle_torch as oracle_torch
    dev = pts_dev.device
    if tree is None:
        tree = cKDTree(pts, leafsize=10)
    _, idx_p = tree.query(queries, k=num_pts_local, workers=-1)
    loc = oracle.normalize_patches(pts[idx_p], queries)
    loc_dev = torch.from_numpy(loc).to(dev)
    _, idx64 = cKDTree(pts, leafsize=10).query(queries, k=64, workers=-1)  # rebuilt per batch like kdtree_query_oneshot
    idx_dev = torch.from_numpy(idx64.astype(np.int64)).to(dev)
    q_dev = torch.from_numpy(queries).to(dev)
    with torch.autocast('cuda', dtype=torch.float16):
        occ, _ = oracle_torch.from_latent(weights_dev, pts_dev, lat_dev, q_dev, idx_dev, loc_dev)
    return occ.float().cpu().numpy()


def time_reference_gpu(args, dev, pts_np, latents_cn, grid_queries_np, batches):
    """Mquery/s of the reference's GPU predict path on this box; `batches` batches of 50 000 (25 000 for 200nn) consecutive grid
    vertices from the middle of the volume, after one warm-up batch"""
    import torch
    from scipy.spatial import cKDTree
    from oracle import ppsurf_oracle as oracle
    from oracle import ppsurf_oracle_torch as oracle_torch
    torch.set_num_threads(os.cpu_count() or 1)
    old = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision('high')
    try:
        weights_dev = oracle_torch.to_device(oracle.make_state_dict(42), dev)
        pts_dev = torch.from_numpy(pts_np).to(dev)
        lat_dev = torch.from_numpy(np.ascontiguousarray(latents_cn.T)).to(dev)
        batch = 25000 if args.num_pts_local >= 200 else REFERENCE_BATCH
        mid = grid_queries_np.shape[0] // 2
        tree = cKDTree(pts_np, leafsize=10)
        out, times = None, []
        for b in range(-1, batches):
            q = grid_queries_np[mid + max(b, 0) * batch: mid + (max(b, 0) + 1) * batch]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = reference_gpu_decode(oracle, weights_dev, pts_np, pts_dev, lat_dev, q, args.num_pts_local, tree)
            torch.cuda.synchronize()
            if b >= 0:
                times.append(time.perf_counter() - t0)
        per_batch = float(np.mean(times))
        return {'value': batch / per_batch / 1e6, 'unit': UNIT, 'ms_per_batch': per_batch * 1e3, 'batch': batch, 'batches': batches,
                'host_threads': torch.get_num_threads(),
                'what': 'reference GPU predict path on this B200: torch-eager modules (oracle torch twin) on the device, TF32 high + fp16 '
                        'autocast, CPU cKDTree (k=P on the per-cloud tree, tree REBUILT + k=64 per batch), numpy patch normalisation, '
                        'H2D / D2H per batch'}, out
    finally:
        torch.set_float32_matmul_precision(old)


def run_reference_gpu(args):
    """`--impl reference-gpu`: the arm the north star's >= 10x target is defined against"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import ppsurf_oracle as oracle
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    pts = oracle.synthetic_cloud(args.points, 42)
    latents = np.random.default_rng(7).standard_normal((256, args.points)).astype(np.float32)
    grid = oracle.dense_grid_queries(pts, args.resolution, 1)
    res, _ = time_reference_gpu(args, dev, pts, latents, grid, max(args.steps, 1))
    print(json.dumps({
        'impl': 'reference-gpu', 'metric': METRIC, 'value': res['value'], 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
        'warmup': 1, 'ms_per_step': res['ms_per_batch'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'tf32 + fp16 autocast', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'step': 'one rec_batch_size batch of {} grid vertices'.format(res['batch'])},
        'reference_gpu': res, 'gpu_launches': 0}))


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------

def fkaconv_roofline(net, dev, peaks, batch=64):
    """secondary roofline: the FKAConv layer resnetb01.cv1 (32->32 @ 10000 points, k=16) at a batch whose
    unique bytes exceed the L2, as achieved GB/s of its UNIQUE HBM bytes
    B*[N_in*(C_in+3)*4 + N_s*16*4 + N_s*12 + N_s*C_out*4] + C_in*C_out*64 against the measured HBM peak"""
    import torch
    from ppsurf_b200 import ops, synthetic
    w = net.packed()['encoder']['resnetb01']['cv1']
    cin, cout, n = w.struct.cin, w.struct.cout, 10000
    clouds = [torch.from_numpy(synthetic.synthetic_cloud(n, 10 + i)).to(dev) for i in range(4)]
    p = torch.stack(clouds).repeat(batch // 4, 1, 1).contiguous()
    ids = torch.stack([ops.knn(c, c, 16) for c in clouds]).repeat(batch // 4, 1, 1).contiguous()
    x = torch.randn((batch, n, cin), device=dev)
    for _ in range(3):
        ops.fkaconv(w, x, p, p, ids)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 5
    for _ in range(reps):
        ops.fkaconv(w, x, p, p, ids)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    unique = batch * (n * (cin + 3) * 4 + n * 16 * 4 + n * 12 + n * cout * 4) + cin * cout * 64
    flops = 2.0 * batch * n * (16 * 16 * cin + 16 * cin * cout)
    achieved = unique / (ms * 1e-3) / 1e9
    return {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
            'traffic': None, 'kernel': 'pps_fkaconv_forward (fka_weight x3 + fka_feat + contraction), resnetb01.cv1 32->32 @ 10000 pts, '
            'k=16, batch {}'.format(batch), 'ms_per_call': ms, 'unique_bytes': unique,
            'fp32_tflops': flops / (ms * 1e-3) / 1e12,
            'note': 'arithmetic intensity {:.0f} FLOP per unique byte: above the fp32 SIMT machine balance, so the layer is '
                    'compute-bound before it is HBM-bound'.format(flops / unique)}


def predict_block(model, pts_np, dev, args):
    """secondary metric: the real predict path of one cloud through PPSurfModel.reconstruct -- encoder (latent loop),
    per-cloud decoder setup, region-grown shell decode -- wall clock, host in / host out"""
    import torch
    pts_ms = torch.from_numpy(pts_np[None].copy()).to(dev)
    model.network.sampling_seed = 42
    best = None
    for _ in range(3):  # the first two runs also warm the CUDA graphs of the encoder batches
        torch.manual_seed(42)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rec = model.reconstruct(pts_ms, resolution=args.resolution)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        if best is None or total < best['total_s']:
            st = dict(model.last_reconstruct_stats)
            best = {'total_s': total, 'encoder_s': st['encoder_s'], 'setup_ms': st['setup_s'] * 1e3, 'shell_decode_s': st['volume_s'],
                    'shell_queries': st['shell_queries'], 'sweeps': st['sweeps'],
                    'shell_mquery_per_s': st['shell_queries'] / st['volume_s'] / 1e6,
                    'decoded_fraction_of_grid': st['shell_queries'] / float((args.resolution + 2) ** 3)}
    best['what'] = 'PPSurfModel.reconstruct(pts [1,N,3]): encode_cloud (>=10 encodings per point) + decoder_for + create_volume (region ' \
                   'growing, source/poco_utils.py:178-254), volume returned to the host; best of 3'
    return best


def run_b200(args):
    import torch
    import torch.distributed as dist

    import ppsurf_b200
    from ppsurf_b200 import _lib, ops, synthetic

    quiet = QuietStdout()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    ops.require_device()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.lib

    model = ppsurf_b200.PPSurfModel(
        pointnet_latent_size=256, output_names=['imp_surf_sign'], in_channels=3, out_channels=2, k=64, lambda_l1=0.0,
        debug=False, in_file='bench', results_dir='results', padding_factor=0.05, name='ppsurf_50nn',
        network_latent_size=256, gen_subsample_manifold_iter=10, gen_subsample_manifold=10000,
        gen_resolution_global=args.resolution, num_pts_local=args.num_pts_local, rec_batch_size=50000, gen_refine_iter=10,
        workers=8)
    net = model.network
    sd = synthetic.make_state_dict(net, 42)
    net.load_state_dict(sd, strict=True)
    model = model.to(dev)
    net.decode_chunk, net.decode_path = args.chunk, args.path
    pts_np = synthetic.synthetic_cloud(args.points, 42)
    pts_bcn = torch.from_numpy(pts_np.T[None].copy()).to(dev)

    # ---- encode: the latent loop's passes are dealt to the ranks, ONE all-reduce of the partial sums (SURVEY.md §8e option B);
    # `--encoder rank0` is option A of the survey: rank 0 encodes alone, one broadcast of the latent table
    from ppsurf_b200.sharding import Shard
    encoder_s = None
    collective_ms = collective_first_ms = None
    sharded = world > 1 and args.encoder == 'sharded'
    if args.latents == 'encoder':
        net.sampling_seed = 42
        if sharded:
            model.shard = Shard.from_env()
        latents = torch.empty((1, 256, args.points), dtype=torch.float32, device=dev)
        if sharded or rank == 0:
            t0 = time.perf_counter()
            for _ in range(2):  # warm-up: a batch shape is captured into a CUDA graph on its second use, replayed from the third
                model.encode_cloud(pts_bcn, generator=torch.Generator().manual_seed(1))
            torch.cuda.synchronize()
            collective_first_ms = (time.perf_counter() - t0) * 1e3 if sharded else None  # includes NCCL set-up and the graph captures
            if sharded:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            latents = model.encode_cloud(pts_bcn, generator=torch.Generator().manual_seed(42)).contiguous()
            torch.cuda.synchronize()
            encoder_s = time.perf_counter() - t0  # 10 encodings per point on random 10k-point subsets = ~100 passes (all-reduce inside)
        model.shard = Shard()  # the dense decode below is sharded by bench.grid_blocks, not by the model
    else:
        latents = torch.from_numpy(np.random.default_rng(7).standard_normal((1, 256, args.points)).astype(np.float32)).to(dev)
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not sharded:
            dist.broadcast(latents, src=0)
        t = torch.tensor([encoder_s or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        encoder_s = float(t[0]) or None
        scratch = torch.zeros_like(latents)
        (dist.all_reduce(scratch) if sharded else dist.broadcast(scratch, src=0))
        dist.barrier()
        e0.record()
        for _ in range(5):
            (dist.all_reduce(scratch) if sharded else dist.broadcast(scratch, src=0))
        e1.record()
        torch.cuda.synchronize()
        collective_ms = e0.elapsed_time(e1) / 5  # steady state of the one collective on the 102 MB table
        del scratch

    net.decoder_for(pts_bcn, latents)  # warm (cub temp storage, first launches)
    net._decoder_cache = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec = net.decoder_for(pts_bcn, latents)  # per-cloud setup: kNN index build + hoisted fc1 table
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    r = args.resolution + 2
    total = r ** 3
    spans = grid_blocks(total, args.shard_of, 0) if (args.shard_of > 1 and world == 1) else grid_blocks(total, world, rank)
    count = sum(c for _, c in spans)
    step, bmin_pad, _ = model.grid_definition(pts_np, args.resolution, 1)
    queries = torch.cat([ops.grid_queries(r, step, bmin_pad, first=f, count=c, device=dev) for f, c in spans]) if len(spans) > 1 \
        else ops.grid_queries(r, step, bmin_pad, first=spans[0][0], count=count, device=dev)
    occ = torch.empty((count,), dtype=torch.float32, device=dev)
    ws = dec.workspace(min(args.chunk, count))
    stream = torch.cuda.current_stream()

    def one_step():
        _lib.check(lib.pps_decoder_decode(dec.packed.ref, dec.index.buf.data_ptr(), dec.pts.data_ptr(), dec.table.data_ptr(),
                                          dec.n, queries.data_ptr(), count, min(args.chunk, count), ws.data_ptr(), ws.numel(),
                                          None, occ.data_ptr(), None, args.path, stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = args.warmup if args.profile_run else max(args.warmup, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        for _ in range(warmup):
            one_step()
        barrier()
        launches0 = lib.pps_launch_count()
        lib.pps_profile_enable(1)
        barrier()
        clocks.mark()
        e0.record()
        for _ in range(args.steps):
            one_step()
        e1.record()
        barrier()
    elapsed_ms = e0.elapsed_time(e1)
    import ctypes
    dom_ms, brackets = ctypes.c_double(0), ctypes.c_longlong(0)
    _lib.check(lib.pps_profile_read(ctypes.byref(dom_ms), ctypes.byref(brackets)))
    lib.pps_profile_enable(0)
    launches = lib.pps_launch_count() - launches0

    # ---- end to end through the public host-buffer call: pinned host queries in, pinned host occupancy out
    q_host = queries.cpu().pin_memory()
    occ_host = torch.empty((count,), dtype=torch.float32).pin_memory()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_run:
        e2.record()
        e3.record()
        barrier()
    else:
        dec.decode_host(q_host, occ_host)  # warm (allocates staging)
        barrier()
        e2.record()
        for _ in range(args.steps):
            dec.decode_host(q_host, occ_host)
        e3.record()
        barrier()
        assert torch.equal(occ_host, occ.cpu()), 'host-buffer path disagrees with the device path'
    e2e_ms = max(e2.elapsed_time(e3), 1e-6)

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])

    # ---- secondary metric: the whole predict path of one cloud; on N ranks the latent loop and every sweep's queries are sharded
    predict = None
    if not args.profile_run and not args.no_predict:
        model.shard = Shard.from_env() if world > 1 else Shard()
        predict = predict_block(model, pts_np, dev, args)
        predict['ranks'] = world
        model.shard = Shard()
        if world > 1:
            t = torch.tensor([predict['total_s'], predict['encoder_s'], predict['shell_decode_s']], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            predict['total_s'], predict['encoder_s'], predict['shell_decode_s'] = (float(v) for v in t)

    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = elapsed_ms / args.steps
        value = total / (ms_per_step * 1e-3) / 1e6
        e2e_value = total / (e2e_ms / args.steps * 1e-3) / 1e6
        rows = count * 64.0 * args.steps  # (query, neighbour) rows this rank pushed through the dominant GEMMs
        dom_flops = rows * GEMM_FLOP_PER_ROW
        achieved = dom_flops / (dom_ms.value * 1e-3) / 1e12 if dom_ms.value > 0 else None
        kernel = 'linear_kernel<128,true> x2 + linear_kernel<64,true> (fp32 SIMT fc2/fc3/fc_query)' if args.path == 0 else \
            'projection_tc_kernel (tcgen05 split-fp16: gather + fc2/fc3/fc_query + softmax + attention pooling)'
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32' if args.path == 0 else 'f32 via split-fp16 tensor cores (fp32 accumulate)', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'parallelism': 'grid blocks of {} vertices dealt round-robin x{}'.format(GRID_BLOCK, world), 'chunk': args.chunk,
                       'decode_path': args.path, 'latents': args.latents,
                       'l2': 'working set per step (fc1 table {} MB + >2 GB of chunk activations) exceeds the 126 MB L2'.format(
                           args.points * 1024 // 2 ** 20),
                       'seeded_weights': 'ppsurf_b200.synthetic.make_state_dict(seed=42)'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(total * 12), 'd2h_bytes_per_step': int(total * 4),
                    'ms_per_step': e2e_ms / args.steps,
                    'call': 'pps_decoder_decode_host: pinned host queries -> device -> pinned host occupancy, chunked copies overlapped'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                         'frac': (achieved / peaks['bf16_tflops']) if achieved else None,
                         # DRAM bytes per launch from the ncu --set full capture under profiles/ (185.6 MB read + 267.7 MB written by
                         # the first launch over 300763 queries = 1507 B per query: the pooled output, 1024 B per query, minus what
                         # was still in the L2 when the capture ended, plus the cold first read of the fc1 table and the neighbour
                         # ids; in the timed loop table and weights are L2 hits), scaled to this run's queries per launch
                         'traffic': (1507.0 * count * args.steps / max(int(brackets.value), 1)) if args.path == 1 else None,
                         'traffic_source': 'profiles/r02_decode_kernels_full_session3_summary.csv' if args.path == 1 else None,
                         'kernel': kernel, 'kernel_ms_per_step': dom_ms.value / args.steps,
                         'kernel_share_of_step': dom_ms.value / elapsed_ms, 'brackets': int(brackets.value),
                         'flop_per_row_executed': GEMM_FLOP_PER_ROW, 'peak_source': peaks['source'],
                         'whole_decode_tflops_reference_formulation': value * 1e6 * FLOP_PER_QUERY_REFERENCE / 1e12},
            'clocks': clocks.summary(),
            'encoder_s': encoder_s, 'setup_ms': setup_ms,
            'encoder_parallelism': ('passes dealt to {} ranks, one all-reduce(sum) of latent sums + counts'.format(world) if sharded else
                                    ('rank 0 encodes, one broadcast' if world > 1 else 'single GPU')),
            'collective_ms': collective_ms, 'encoder_warmup_ms': collective_first_ms,
        }
        if predict is not None:
            out['e2e_predict'] = predict
        if not args.profile_run:
            out['roofline_fkaconv'] = fkaconv_roofline(net, dev, peaks)
        if not args.no_cpu_baseline and world == 1:
            from oracle import ppsurf_oracle as oracle  # CPU baseline + checker leg only
            weights = {k: v.numpy() for k, v in sd.items()}
            rng = np.random.default_rng(7)
            sel = np.sort(rng.choice(count, args.cpu_sample, replace=False))
            q_np = queries[torch.from_numpy(sel).to(dev)].cpu().numpy()
            lat_cn = latents[0].cpu().numpy()
            cpu_decode(oracle, weights, pts_np, lat_cn, q_np[:256], args.num_pts_local)  # warm BLAS threads
            t0 = time.perf_counter()
            ref = cpu_decode(oracle, weights, pts_np, lat_cn, q_np, args.num_pts_local)
            dt = time.perf_counter() - t0
            out['cpu_baseline'] = {'value': args.cpu_sample / dt / 1e6, 'unit': UNIT, 'cores': __import__('torch').get_num_threads(), 'kind': 'port',
                                   'sample': '{} random vertices of the same grid, torch-CPU oracle + scipy cKDTree (kd-tree rebuilt '
                                             'per batch like the reference), {:.1f} s'.format(args.cpu_sample, dt)}
            got = occ[torch.from_numpy(sel).to(dev)].cpu().numpy()
            # (a) the network arithmetic: oracle on the SAME neighbours as the device (the device's exact fp32 kNN is pinned by
            # the parity tests), first 8192 queries of the sample
            nchk = min(8192, args.cpu_sample)
            idx_dev = dec.index.query(torch.from_numpy(q_np[:nchk]).to(dev), max(64, args.num_pts_local)).cpu().numpy()
            ref_same = cpu_decode(oracle, weights, pts_np, lat_cn, q_np[:nchk], args.num_pts_local, idx_given=idx_dev)
            out['max_abs_err_vs_oracle'] = float(np.abs(got[:nchk] - ref_same).max())
            # (b) against the timed CPU run, whose float64 kd-tree (scipy stand-in for pykdtree) breaks near-ties of the 50th /
            # 64th neighbour differently from fp32 distances: a handful of queries get a different patch or neighbour set
            diff = np.abs(got - ref)
            out['vs_float64_kdtree'] = {'max_abs_diff': float(diff.max()), 'queries_above_1e-4': int((diff > 1e-4).sum()),
                                        'sample': int(args.cpu_sample)}
            if not args.no_reference_gpu:
                grid_np = oracle.dense_grid_queries(pts_np, args.resolution, 1)
                ref_gpu, _ = time_reference_gpu(args, dev, pts_np, lat_cn, grid_np, args.refgpu_batches)
                ref_gpu['speedup_of_this_repo'] = {'device_resident': value / ref_gpu['value'], 'e2e_host_buffers': e2e_value / ref_gpu['value']}
                out['reference_gpu'] = ref_gpu
        if not args.profile_run and not args.no_fit and world == 1:
            # BASELINE config 5 next to the headline: a short measurement of the training step (own line: --workload fit)
            del dec, ws
            ops._shared_workspace.clear()
            torch.cuda.empty_cache()
            out['fit'] = fit_measure(model, args, dev, 1, 3, 3, reference_twin=not args.no_reference_gpu)
            model.eval()
        quiet.emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---- BASELINE config 5: the training step ------------------------------------------------------------------------------------------

def fit_batch(n_clouds, n_pts, n_qry, seed):
    """synthetic training batch in the reference's collated layout (PPSurfDataset.__getitem__, source/ppsurf_data_loader.py:61-81):
    noisy-sphere clouds, query points half near the surface and half uniform, signed distance to the sphere as the label source"""
    from ppsurf_b200 import synthetic
    rng = np.random.default_rng(seed)
    pts = np.stack([synthetic.synthetic_cloud(n_pts, seed * 1000 + i) for i in range(n_clouds)])
    near = pts[:, rng.integers(0, n_pts, n_qry // 2)] + 0.03 * rng.standard_normal((n_clouds, n_qry // 2, 3))
    far = rng.uniform(-0.5, 0.5, (n_clouds, n_qry - n_qry // 2, 3))
    qry = np.concatenate([near, far], axis=1).astype(np.float32)
    dist = (np.linalg.norm(qry, axis=2) - 0.4).astype(np.float32)
    return {'pts_ms': pts.astype(np.float32), 'pts_query_ms': qry, 'imp_surf_dist_ms': dist}


def fit_measure(model, args, dev, world, steps, warmup, reference_twin=False):
    """times `steps` training steps (forward, loss, backward, DDP gradient all-reduce when world > 1, AdamW update) with CUDA
    events, max over ranks.  value: batch prepared and resident; e2e: pinned host batch -> H2D -> prepare_batch (kd-tree work of
    the reference's DataLoader workers, on the device) -> step -> loss to the host."""
    import torch
    import torch.distributed as dist
    from ppsurf_b200 import _lib, autograd as ag, data_pipeline
    net = model.network
    rank = int(os.environ.get('RANK', '0'))
    host = {k: torch.from_numpy(v).pin_memory() for k, v in fit_batch(args.fit_clouds_per_gpu, args.fit_points, args.fit_queries,
                                                                       100 + rank).items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    model.train()
    net.sampling_seed = 7
    from ppsurf_b200 import training
    ag.set_precision(args.fit_precision)
    ag.manual_seed(1234 + rank)

    def prepare():
        with torch.no_grad():
            return data_pipeline.prepare_batch(net, {k: v.to(dev, non_blocking=True) for k, v in host.items()})

    batch = prepare()
    if args.fit_eager:
        module = net
        if world > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            module = DDP(net, device_ids=[dev.index], gradient_as_bucket_view=True)
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-5, weight_decay=1e-2)

        def step(batch):
            opt.zero_grad(set_to_none=True)
            pred = module(dict(batch))
            b, c, q = pred.shape
            loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(b * q, c), batch['occ'].reshape(-1))
            loss.backward()
            opt.step()
            return loss
    else:
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-5, weight_decay=1e-2, capturable=True)
        step = training.GraphedTrainStep(net, opt, batch, world=world)

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    losses = [float(step(batch).detach()) for _ in range(max(3, warmup))]
    launches0 = _lib.lib.pps_launch_count()
    ms = timed(lambda: step(batch), steps)
    launches = (_lib.lib.pps_launch_count() - launches0) if args.fit_eager else step.kernels_per_step * steps
    # e2e: the batch of step i+1 is prepared (H2D + prepare_batch) on a side stream while step i runs, like the reference's DataLoader
    # workers prepare batches while the GPU trains; the loss of every step is read on the host
    side = torch.cuda.Stream(device=dev)

    def prepare_on_side():
        with torch.cuda.stream(side):
            return prepare()

    pending = [prepare_on_side()]

    def e2e_step():
        cur = pending.pop()
        torch.cuda.current_stream().wait_stream(side)
        loss = step(cur)                    # asynchronous (graph replay)
        pending.append(prepare_on_side())   # overlaps the step
        return float(loss.detach())         # the one host read of the step

    ms_e2e = timed(e2e_step, steps)
    torch.cuda.synchronize()
    ag.set_precision('fp32')
    clouds = args.fit_clouds_per_gpu * world
    out = {'value': clouds * steps / (ms / 1e3), 'unit': 'clouds/s', 'ms_per_step': ms / steps,
           'e2e': {'value': clouds * steps / (ms_e2e / 1e3), 'unit': 'clouds/s', 'ms_per_step': ms_e2e / steps,
                   'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': 4,
                   'call': 'pinned host batch -> H2D -> data_pipeline.prepare_batch (supports, 13 kNN index tensors, proj_ids, patches on the '
                           'device; on a side stream, overlapping the previous step) -> training step -> loss.item()'},
           'gpu_launches_per_step': int(launches // steps), 'loss_first_last_warmup': [losses[0], losses[-1]],
           'config': {'workload': 'ppsurf_50nn fit (BASELINE config 5): {} clouds/GPU x {} GPUs, {} manifold points, {} query points, '
                                  'k=64, P=50, {} GEMMs on tcgen05 (fp32 master weights / activations), AdamW, {}'.format(
                                      args.fit_clouds_per_gpu, world, args.fit_points, args.fit_queries, args.fit_precision,
                                      ('one flat NCCL gradient all-reduce' if not args.fit_eager else 'torch DDP') if world > 1 else 'single GPU') +
                                      (', eager launches' if args.fit_eager else ', step replayed from CUDA graphs'),
                      'global_batch': clouds}}
    if reference_twin and world == 1:
        # the reference's GPU training step: eager torch modules (oracle twin) on this device under autocast(bfloat16) with the same
        # optimiser and the same prepared batch
        from oracle import ppsurf_train_oracle as T
        twin = T.State({k: v.detach().cpu() for k, v in net.state_dict().items()}, device=dev)
        topt = torch.optim.AdamW(list(twin.p.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-5, weight_decay=1e-2)
        torch.backends.cuda.matmul.allow_tf32 = True

        def twin_step():
            topt.zero_grad(set_to_none=True)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                logits = T.forward(twin, batch, True, 0.3)
            T.loss_of(logits.float(), batch['occ']).backward()
            topt.step()

        for _ in range(3):
            twin_step()
        tms = timed(twin_step, steps)
        out['reference_gpu'] = {'value': clouds * steps / (tms / 1e3), 'unit': 'clouds/s', 'ms_per_step': tms / steps,
                                'what': 'torch-eager twin of the reference network (oracle/ppsurf_train_oracle.py) on this B200, '
                                        'autocast bf16 + TF32, same batch (prepared on the device), same optimiser',
                                'speedup_of_this_repo': tms / ms}
        del twin, topt
    return out


def run_fit(args):
    import torch
    import torch.distributed as dist

    import ppsurf_b200
    from ppsurf_b200 import ops, synthetic
    quiet = QuietStdout()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    ops.require_device()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'bench', 'results', 0.05, 'ppsurf_50nn', 256, 10, 10000,
                                    129, 50, 50000, 10, 8)
    model.network.load_state_dict(synthetic.make_state_dict(model.network, 42), strict=True)
    model = model.to(dev)
    with ClockSampler(local_rank) as clocks:
        clocks.mark()
        res = fit_measure(model, args, dev, world, args.steps, args.warmup, reference_twin=not args.no_reference_gpu)
    if rank == 0:
        out = {'metric': 'clouds/sec training step, ppsurf_50nn fit (BASELINE config 5)', 'value': res['value'], 'unit': 'clouds/s',
               'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': res['ms_per_step'],
               'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'bf16 GEMM operands, fp32 accumulate / master weights' if args.fit_precision == 'bf16' else 'f32', 'data': 'synthetic',
               'config': res['config'], 'e2e': res['e2e'], 'gpu_launches': res['gpu_launches_per_step'] * args.steps,
               'clocks': clocks.summary(), 'loss_first_last_warmup': res['loss_first_last_warmup']}
        if 'reference_gpu' in res:
            out['reference_gpu'] = res['reference_gpu']
        quiet.emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.impl == 'reference-gpu':
        run_reference_gpu(args)
    elif args.workload == 'fit':
        run_fit(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
