"""``PPSurfModel``: the reference LightningModule's constructor / ``predict_step`` surface
(source/ppsurf_model.py:10-36, source/poco_model.py:183-273) on the B200 network, plus the reconstruction driver
(source/poco_utils.py:26-254) with every per-query step on the device.  Selected from the reference CLI with
``--model.class_path ppsurf_b200.PPSurfModel``; ``pps.py`` itself is not edited.

On the device: the latent loop, the region-growing masks and frontier lists (``ops.RegionVolume``), marching cubes and the bisection
refinement (``ops.marching_cubes`` / ``ops.VertexRefiner``), the train / test batch preparation and the training step.  On the host, as in
the reference: file I/O and the mesh cleaning (``ppsurf_b200.mesh``, no trimesh / scikit-image needed).
"""
import os
import queue
import threading
import time
import typing

import numpy as np
import torch

from . import ops
from .network import PPSurfNetwork, _Base
from .sharding import Shard


class _NullBar:
    """stands in for Lightning's progress bar when the model is driven without a Trainer"""

    class _Bar:
        @staticmethod
        def set_postfix_str(*_a, **_k):
            pass

    predict_progress_bar = _Bar()
    test_progress_bar = _Bar()


class PPSurfModel(_Base):

    def __init__(self, pointnet_latent_size, output_names, in_channels, out_channels, k, lambda_l1, debug, in_file,
                 results_dir, padding_factor, name, network_latent_size, gen_subsample_manifold_iter,
                 gen_subsample_manifold, gen_resolution_global, num_pts_local, rec_batch_size, gen_refine_iter, workers):
        super().__init__()
        self.output_names = output_names
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.k = k
        self.lambda_l1 = lambda_l1
        self.network_latent_size = network_latent_size
        self.gen_subsample_manifold_iter = gen_subsample_manifold_iter
        self.gen_subsample_manifold = gen_subsample_manifold
        self.gen_resolution_global = gen_resolution_global
        self.rec_batch_size = rec_batch_size
        self.gen_refine_iter = gen_refine_iter
        self.workers = workers
        self.in_file = in_file
        self.results_dir = results_dir
        self.padding_factor = padding_factor
        self.debug = debug
        self.name = name
        self.num_pts_local = num_pts_local
        self.pointnet_latent_size = pointnet_latent_size
        self.network = PPSurfNetwork(in_channels=in_channels, latent_size=network_latent_size, out_channels=out_channels,
                                     k=k, num_pts_local=num_pts_local, pointnet_latent_size=pointnet_latent_size,
                                     decode_chunk=ops.DEFAULT_CHUNK)
        # rec_batch_size (50 000 / 25 000 in the reference's configs) bounds the REFERENCE's activation memory per from_latent call; here
        # it is kept for the interface only: the decode works in its own launch granularity (ops.DEFAULT_CHUNK), sized for a B200
        self.test_step_outputs = []
        self.shard = Shard()  # multi-GPU predict: set to Shard.from_env() (one process per GPU), see sharding.py

    # ---- a1: latent averaging loop (source/poco_model.py:200-237) ----------------------------------------------
    def _schedule_np(self, n: int, generator: typing.Optional[torch.Generator] = None):
        """index sets of the latent loop (source/poco_model.py:207-224) as ``(ids int64 ndarray, may_repeat)``.  The reference draws,
        pass after pass, ``sub`` points uniformly without replacement from the points still at the current visit count; consecutive
        chunks of ONE random permutation of those points are the same distribution, so one ``randperm`` per iteration replaces one per
        pass (the host schedule was on the critical path of the encoder).  Only the last chunk of an iteration is padded with random
        points of the whole cloud (``may_repeat``); padded points are visited twice in this iteration and skipped by later ones,
        exactly like the reference's ``counts == current`` test does."""
        sub = self.gen_subsample_manifold
        counts = np.zeros(n, dtype=np.int64)
        for current in range(self.gen_subsample_manifold_iter):
            while True:
                valid = np.nonzero(counts == current)[0]
                if valid.shape[0] == 0:
                    break
                if n < sub:
                    counts += 1
                    yield np.arange(n), False
                    continue
                perm = valid[torch.randperm(valid.shape[0], generator=generator).numpy()]
                for s0 in range(0, perm.shape[0], sub):
                    ids = perm[s0:s0 + sub]
                    if ids.shape[0] < sub:
                        ids = np.concatenate([ids, torch.randperm(n, generator=generator)[:sub - ids.shape[0]].numpy()])
                        counts[np.unique(ids)] += 1  # `counts[ids] += 1` counts a repeated id once
                        yield ids, True
                    else:
                        counts[ids] += 1
                        yield ids, False

    def _rotated_schedule(self, n: int, generator: typing.Optional[torch.Generator] = None, rot_seed=None):
        """the schedule with the random rotations of every pass's four support samplings, drawn in pass order from ONE generator: a
        pass gets the same rotations whichever rank or batch it lands in"""
        from .sampling import ROUNDS, random_rotations
        rot_gen = np.random.default_rng(rot_seed)
        for ids, rep in self._schedule_np(n, generator):
            yield ids, rep, random_rotations(rot_gen, 4 * ROUNDS).reshape(4, ROUNDS, 9)

    def latent_schedule(self, n: int, generator: typing.Optional[torch.Generator] = None) -> typing.Iterator[torch.Tensor]:
        """index sets of the latent loop (source/poco_model.py:207-224), lazily.  They depend only on the visit counts,
        never on network output, so the host draws the next sets while the device still works on the previous batch."""
        for ids, _ in self._schedule_np(n, generator):
            yield torch.from_numpy(ids)

    def _latent_lanes(self, dev):
        lanes = getattr(self, '_lanes', None)
        if lanes is None or lanes[0].device != dev:
            self._lanes = lanes = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        return lanes

    def encode_cloud(self, pts_bcn: torch.Tensor, generator: typing.Optional[torch.Generator] = None,
                     prog_bar=None, batch_passes: int = 16) -> torch.Tensor:
        """``pts_bcn [1,3,N]`` (device) -> latents ``[1,latent,N]``: every point is encoded at least
        ``gen_subsample_manifold_iter`` times on random ``gen_subsample_manifold``-point subsets and averaged
        (source/poco_model.py:200-237).  The passes are independent given the schedule, so ``batch_passes`` of them go
        through the encoder as one batch (InstanceNorm statistics are per sample), replayed from a CUDA graph
        (``PPSurfNetwork.latents_of_batch``); the accumulation keeps pass order."""
        pts = pts_bcn[0].transpose(0, 1).contiguous()  # [N,3]
        n = pts.shape[0]
        dev = pts.device
        net = self.network
        latent = torch.zeros((n, self.network_latent_size), dtype=torch.float32, device=dev)
        counts = torch.zeros((n,), dtype=torch.float32, device=dev)
        rot_seed = net.sampling_seed
        if self.shard.world > 1 and (generator is None or rot_seed is None):
            # every rank must walk the same schedule: without caller-supplied seeds rank 0's random seed is shared
            seed = self.shard.common_seed(dev)
            generator = torch.Generator().manual_seed(seed) if generator is None else generator
            rot_seed = seed if rot_seed is None else rot_seed
        schedule = self.shard.my_passes(self._rotated_schedule(n, generator, rot_seed))
        iteration = 0
        sub = min(self.gen_subsample_manifold, n)

        def prepare(group):
            # one host->device copy per batch: the point ids of every pass, the first occurrence of every distinct id
            # (torch semantics of `latent[ids] += x` with repeated ids: one writer wins, counted once) as rows of the
            # batch's point-major output and as destination points; plus the random rotations of the support samplings
            ids_np = [g[0] for g in group]
            firsts = [np.unique(a, return_index=True)[1] if rep else None for a, rep, _ in group]
            n_first = [sub if f is None else f.shape[0] for f in firsts]
            if all(f is None for f in firsts):
                extra = []  # no repeated id in the whole batch: rows = 0..B*sub-1, destinations = the ids themselves
            else:
                rows = np.concatenate([(np.arange(sub) if f is None else f) + k * sub for k, f in enumerate(firsts)]).astype(np.int32)
                dsts = np.concatenate([a if f is None else a[f] for a, f in zip(ids_np, firsts)]).astype(np.int32)
                extra = [rows, dsts]
            packed = torch.from_numpy(np.concatenate([np.concatenate(ids_np).astype(np.int32)] + extra)).pin_memory()
            rot = torch.from_numpy(np.stack([g[2] for g in group])).pin_memory()
            return len(group), n_first, packed, rot, not extra

        # the schedule depends only on the host-side visit counts, never on network output: a producer thread draws and
        # prepares the next batches while the device works on the current one (numpy / torch release the GIL)
        batches: queue.Queue = queue.Queue(maxsize=3)
        stop = threading.Event()

        def put(item):
            while not stop.is_set():
                try:
                    batches.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def producer():
            try:
                while not stop.is_set():
                    group = [g for _, g in zip(range(batch_passes), schedule)]
                    if not put(prepare(group) if group else None) or not group:
                        return
            except BaseException as err:  # surfaces in the consumer
                put(err)

        thread = threading.Thread(target=producer, daemon=True)
        thread.start()
        # two batches in flight: batch i runs (inputs, index generation, network -- one CUDA-graph replay) on lane i % 2 while the main
        # stream accumulates batch i-1; a lane is reused once the accumulation that read its output has been issued behind an event.
        # A batch is ~6000 small kernels over 10 000-point clouds that cannot fill the device on their own.
        main = torch.cuda.current_stream()
        lanes = self._latent_lanes(dev)
        consumed = [None, None]
        turn = 0
        try:
            while True:
                item = batches.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                b, n_first, packed, rot, plain = item
                lane = lanes[turn % 2]
                if consumed[turn % 2] is not None:
                    lane.wait_event(consumed[turn % 2])
                else:
                    lane.wait_stream(main)
                with torch.cuda.stream(lane):
                    packed = packed.to(dev, non_blocking=True)
                    rot = rot.to(dev, non_blocking=True)
                    all_ids = packed[:b * sub]
                    batch = torch.index_select(pts, 0, all_ids).view(b, sub, 3)  # [B,sub,3] point-major
                    part_pm = net.latents_of_batch(batch, rot).reshape(b * sub, -1)
                    if plain:
                        rows_dev = torch.arange(b * sub, device=dev, dtype=torch.int32)
                        dsts_dev = all_ids
                    else:
                        total_first = sum(n_first)
                        rows_dev, dsts_dev = packed[b * sub:b * sub + total_first], packed[b * sub + total_first:]
                main.wait_stream(lane)
                off = 0
                for nf in n_first:  # pass order is kept: a point revisited inside the batch accumulates in the reference's order
                    ops.latent_accumulate_rows(part_pm, rows_dev[off:off + nf], dsts_dev[off:off + nf], latent, counts)
                    off += nf
                    iteration += 1
                done = torch.cuda.Event()
                done.record(main)
                consumed[turn % 2] = done
                turn += 1
                if prog_bar is not None:
                    prog_bar.predict_progress_bar.set_postfix_str('get_latent iter: {}'.format(iteration), refresh=True)
        finally:
            stop.set()  # a failing consumer must not leave the producer blocked on the bounded queue
            thread.join()
        self.shard.reduce_latents(latent, counts)  # multi-GPU: one all-reduce of the partial sums (no-op on one rank)
        ops.latent_finalize(latent, counts)
        return latent.transpose(0, 1).unsqueeze(0)

    # ---- a11: occupancy volume (source/poco_utils.py:52-61, 178-254) ------------------------------------------------
    @staticmethod
    def grid_definition(input_points: np.ndarray, resolution: int, padding: int = 1):
        bmin, bmax = input_points.min(), input_points.max()
        step = (bmax - bmin) / (resolution - 1)
        bmin_pad = bmin - padding * step
        pts_ids = ((input_points - bmin) / step + padding).astype(np.int32)
        return np.float32(step), np.float32(bmin_pad), pts_ids

    def occupancy(self, decoder: ops.Decoder, queries: torch.Tensor) -> torch.Tensor:
        """softmax(logits)[0] - softmax(logits)[1] per query (source/poco_utils.py:74-82), on the device"""
        return decoder.decode(queries.contiguous(), want_logits=False, want_occ=True)['occ']

    def dense_volume(self, decoder: ops.Decoder, input_points: np.ndarray, resolution: int, padding: int = 1) -> torch.Tensor:
        """all ``(resolution+2*padding)^3`` vertices (the benchmark workload) -> ``[r,r,r]`` fp32 device"""
        step, bmin_pad, _ = self.grid_definition(input_points, resolution, padding)
        r = resolution + 2 * padding
        queries = ops.grid_queries(r, step, bmin_pad, device=decoder.pts.device)
        return self.occupancy(decoder, queries).view(r, r, r)

    def create_volume(self, decoder: ops.Decoder, input_points: np.ndarray, resolution: int, padding: int = 1,
                      dilation_size: int = 2, out_value: float = 1.0, prog_bar=None, pc_file_in: str = 'unknown') -> np.ndarray:
        """region-growing evaluation: only voxels within ``dilation_size`` of an input point, then of a sign change, are
        decoded (source/poco_utils.py:178-254).  Masks, frontier lists, query coordinates and the volume live on the
        device (``ops.RegionVolume``); the host reads one list length per step and the finished volume."""
        volume = self.create_volume_device(decoder, input_points, resolution, padding, dilation_size, out_value, prog_bar,
                                           pc_file_in)
        return volume.cpu().numpy().astype(np.float64)

    def create_volume_device(self, decoder: ops.Decoder, input_points: np.ndarray, resolution: int, padding: int = 1,
                             dilation_size: int = 2, out_value: float = 1.0, prog_bar=None,
                             pc_file_in: str = 'unknown') -> torch.Tensor:
        step, bmin_pad, pts_ids = self.grid_definition(input_points, resolution, padding)
        r = resolution + 2 * padding
        dev = decoder.pts.device
        region = ops.RegionVolume(r, dilation_size, dev)
        lin = np.unique((pts_ids[:, 0].astype(np.int64) * r + pts_ids[:, 1]) * r + pts_ids[:, 2]).astype(np.int32)
        seeds = torch.from_numpy(lin).to(dev)
        sweep = 0
        decoded = 0
        while seeds.shape[0] > 0:
            ids = region.pending(seeds)
            decoded += int(ids.shape[0])
            if ids.shape[0] > 0:
                region.scatter(ids, self.shard.evaluate(lambda q: self.occupancy(decoder, q), region.queries(ids, step, bmin_pad)))
            seeds = region.frontier(seeds, sweep & 1)
            sweep += 1
            if prog_bar is not None:
                prog_bar.predict_progress_bar.set_postfix_str(
                    '{}, occ sweep {}'.format(os.path.basename(pc_file_in), sweep), refresh=True)
        self.last_volume_stats = {'shell_queries': decoded, 'sweeps': sweep}
        return region.finish(padding, out_value)

    # ---- f3: marching cubes + bisection refinement on the device, mesh cleaning on the host ------------------------------------
    def extract_mesh(self, decoder: ops.Decoder, volume, step, bmin_pad, refine_iter: int, prog_bar=None,
                     pc_file_in: str = 'unknown', level: float = 0.0):
        """source/poco_utils.py:87-175 without skimage / trimesh: marching cubes of the occupancy volume and the ``refine_iter``
        bisection sweeps run on the device (``ops.marching_cubes``, ``ops.VertexRefiner``; the volume, the vertices and the bracketing
        state never leave it), then ONE copy to the host and the reference's mesh cleaning (``ppsurf_b200.mesh``).  Returns
        ``(vertices [nv,3] float64 in model space, faces [nf,3] int64)`` or ``None`` when the field has no zero crossing.

        Differences from the reference, by construction: (i) vertices are created once per crossed grid edge, i.e. already merged;
        the reference merges them in its first cleaning pass; (ii) small components are removed once, after the refinement, instead
        of before and after  --  connectivity does not change in between; (iii) ambiguous cells are triangulated by this repo's
        generated case table (``mc_tables``), the reference by skimage's Lewiner tables; the vertex positions agree."""
        vol = volume if isinstance(volume, torch.Tensor) else torch.from_numpy(np.asarray(volume, dtype=np.float32))
        vol = vol.to(decoder.pts.device, torch.float32).contiguous()
        finite = vol[~torch.isnan(vol)]
        if finite.numel() == 0 or not (float(finite.max()) > level > float(finite.min())):
            return None
        verts, vert_edge, faces = ops.marching_cubes(vol, level)
        if faces.shape[0] == 0:
            return None
        refiner = ops.VertexRefiner(vol, verts, vert_edge, step, bmin_pad)
        for it in range(refine_iter):
            if refiner.v.shape[0] > 0:
                refiner.update(self.shard.evaluate(lambda q: self.occupancy(decoder, q), refiner.v))
            if prog_bar is not None:
                prog_bar.predict_progress_bar.set_postfix_str(
                    '{}, refine iter {}'.format(os.path.basename(pc_file_in)[:16], it), refresh=True)
        verts_ms = refiner.result().cpu().numpy().astype(np.float64)
        from . import mesh as mesh_utils
        v, f = mesh_utils.clean_simple(verts_ms, faces.cpu().numpy())
        v, f = mesh_utils.remove_small_connected_components(v, f, num_faces=6)
        return (v, f) if f.shape[0] > 0 else None

    # ---- Lightning surface -----------------------------------------------------------------------------------------
    def get_prog_bar(self):
        trainer = getattr(self, '_trainer', None)
        bar = getattr(trainer, 'progress_bar_callback', None) if trainer is not None else None
        return bar if bar is not None else _NullBar()

    def forward(self, batch):
        return self.network.forward(batch)

    # ---- loss / metrics of the test and validation steps (source/poco_model.py:75-118,134-162) ----------------------
    def compute_loss(self, pred: torch.Tensor, batch_data: dict):
        """cross entropy of the 2-class logits ``pred [B,2,Q]`` against ``occ [B,Q]`` (source/poco_model.py:75-88)"""
        occ_loss = torch.nn.functional.cross_entropy(input=pred, target=batch_data['occ'], reduction='none')
        loss_components = torch.stack([occ_loss])
        loss_components_mean = torch.stack([torch.mean(occ_loss)])
        return loss_components_mean.mean(), loss_components_mean, loss_components

    @staticmethod
    def calc_metrics(pred: torch.Tensor, gt_data: dict) -> dict:
        """accuracy / precision / recall / F1 of ``argmax(pred)`` against ``occ`` with the reference's key names and NaN
        conventions (source/poco_model.py:90-102, source/base/metrics.py:10-84)"""
        predicted = (torch.argmax(pred, dim=1).squeeze() > 0)
        gt = (gt_data['occ'].squeeze() > 0)
        if gt.shape != predicted.shape:
            raise ValueError('The ground truth matrix and the predicted matrix have different sizes!')
        n = float(gt.numel())
        tp = float((predicted & gt).sum())
        fp = float((predicted & ~gt).sum())
        fn = float((~predicted & gt).sum())
        tn = n - tp - fp - fn
        nan = float('NaN')
        res = {'predictions': n, 'pred_gt': n, 'positives': tp + fp, 'pos_gt': tp + fn, 'true_neg': tn,
               'negatives': n - tp - fp, 'neg_gt': n - tp - fn, 'true_pos': tp, 'true': tp + tn, 'false_pos': fp,
               'false_neg': fn, 'false': fp + fn}
        res['accuracy'] = nan if n == 0 else (tp + tn) / n
        res['precision'] = nan if tp + fp == 0 else tp / (tp + fp)
        res['recall'] = nan if tp + fn == 0 else tp / (tp + fn)
        pr = res['precision'] + res['recall']
        res['f1_score'] = nan if pr == 0 else 2.0 * res['precision'] * res['recall'] / pr  # NaN operands propagate
        res['abs_dist_rms'] = np.nan
        return res

    def get_loss_and_metrics(self, pred, batch):
        loss, loss_components_mean, loss_components = self.compute_loss(pred=pred, batch_data=batch)
        return loss, loss_components_mean, loss_components, self.calc_metrics(pred=pred, gt_data=batch)

    def validation_step(self, batch, batch_idx):
        pred = self.network.forward(batch)
        loss, _, _, _ = self.get_loss_and_metrics(pred, batch)
        return loss

    def training_step(self, batch, batch_idx):
        """source/poco_model.py:120-125 (default_step_dict 108-118): train-mode forward through ``ppsurf_b200.training``, mean cross
        entropy over all query points; returns the loss tensor whose ``backward()`` runs the CUDA ``*_bwd`` entry points.  Works
        under Lightning's automatic optimisation, ``torch.autocast(bfloat16)`` (bf16 tensor-core GEMMs) and DDP."""
        from . import autograd as ag
        if not self.network.training:
            self.network.train()
        pred = self.network.forward(batch)  # [B,2,Q]
        b, c, q = pred.shape
        loss, _rows = ag.cross_entropy(pred.transpose(1, 2).reshape(b * q, c), batch['occ'].reshape(-1).to(pred.device, torch.int64))
        self.last_train_pred = pred.detach()
        if float(self.lambda_l1) != 0.0:
            raise NotImplementedError('lambda_l1 != 0: the reference calls self.regularize (source/poco_model.py:112-113), a method it does '
                                      'not define; every PPSurf / POCO configuration sets lambda_l1 = 0')
        if hasattr(self, 'log') and getattr(self, '_trainer', None) is not None:
            # do_logging of the reference (source/poco_model.py:302-322): total loss, the classification metrics of the step, F1 only
            # to the progress bar
            self.log('loss/train/00_all', loss.detach(), on_step=True, on_epoch=False)
            metrics = self.calc_metrics(pred.detach(), batch)
            for key in ('accuracy', 'precision', 'recall', 'f1_score'):
                value = metrics[key]
                self.log('metrics/train/{}'.format(key), 0.0 if value != value else value, on_step=True, on_epoch=False)
            self.log('metrics/train/F1', metrics['f1_score'], on_step=True, on_epoch=False, logger=False, prog_bar=False)
        return loss

    def configure_optimizers(self):
        """configs/poco.yaml:60-77: AdamW(lr 1e-3, betas (0.9, 0.999), eps 1e-5, weight_decay 1e-2) + MultiStepLR([75, 125], 0.1)
        (LightningCLI instantiates them from the yaml; this is the same pair for drivers without the CLI)"""
        opt = torch.optim.AdamW(self.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-5, weight_decay=1e-2, amsgrad=False)
        sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[75, 125], gamma=0.1)
        return {'optimizer': opt, 'lr_scheduler': sched}

    def test_step(self, batch, batch_idx):
        """source/poco_model.py:134-162: forward on the stored query points, loss and classification metrics"""
        pred = self.network.forward(batch)
        if batch['shape_id'].shape[0] != 1:
            raise NotImplementedError('batch size > 1 not supported')
        loss, loss_components_mean, loss_components = self.compute_loss(pred=pred, batch_data=batch)
        metrics_dict = self.calc_metrics(pred=pred, gt_data=batch)
        pc_file_in = batch['pc_file_in'][0]
        results = {'shape_id': batch['shape_id'].squeeze(0), 'pc_file_in': pc_file_in, 'loss': loss,
                   'loss_components_mean': loss_components_mean.squeeze(0), 'loss_components': loss_components.squeeze(0),
                   'metrics_dict': metrics_dict}
        self.test_step_outputs.append(results)
        self.get_prog_bar().test_progress_bar.set_postfix_str('pc_file: {}'.format(os.path.basename(pc_file_in)), refresh=True)
        return results

    def reconstruct(self, pts_ms: torch.Tensor, resolution: typing.Optional[int] = None, dense: bool = False,
                    prog_bar=None, pc_file_in: str = 'unknown', keep_on_device: bool = False) -> dict:
        """encoder + occupancy volume for one cloud ``pts_ms [1,N,3]`` on the device; returns the volume, the grid
        definition and the decoder state (``predict_step`` adds meshing and export on top)."""
        resolution = resolution or self.gen_resolution_global
        pts_bcn = pts_ms.to(torch.float32).transpose(1, 2).contiguous()
        self.network._decoder_cache = None  # a new cloud: nothing of the previous one may survive
        dev = pts_bcn.device
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        latents = self.encode_cloud(pts_bcn, prog_bar=prog_bar)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        decoder = self.network.decoder_for(pts_bcn, latents)
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        input_points = pts_ms[0].cpu().numpy()
        step, bmin_pad, _ = self.grid_definition(input_points, resolution, 1)
        self.last_volume_stats = {'shell_queries': (resolution + 2) ** 3, 'sweeps': 1}
        if dense:
            volume = self.dense_volume(decoder, input_points, resolution)
        else:
            volume = self.create_volume_device(decoder, input_points, resolution, prog_bar=prog_bar, pc_file_in=pc_file_in)
        if not keep_on_device:  # the reference's volume: float64 on the host
            volume = volume.cpu().numpy().astype(np.float64)
        else:
            torch.cuda.synchronize(dev)
        t3 = time.perf_counter()
        self.last_reconstruct_stats = dict(self.last_volume_stats, encoder_s=t1 - t0, setup_s=t2 - t1, volume_s=t3 - t2)
        return {'volume': volume, 'step': step, 'bmin_pad': bmin_pad, 'decoder': decoder, 'latents': latents}

    def predict_step(self, batch: dict, batch_idx, dataloader_idx=0):
        """source/poco_model.py:183-273: one cloud per batch; writes ``<results_dir>/.../<name>.ply``"""
        from . import mesh as mesh_utils
        if batch['pts_ms'].shape[0] > 1:
            raise NotImplementedError('batch size > 1 not supported')
        prog_bar = self.get_prog_bar()
        pc_file_in = batch['pc_file_in'][0] if 'pc_file_in' in batch else 'unknown'
        rec = self.reconstruct(batch['pts_ms'], prog_bar=prog_bar, pc_file_in=pc_file_in, keep_on_device=True)
        mesh = self.extract_mesh(rec['decoder'], rec['volume'], rec['step'], rec['bmin_pad'], self.gen_refine_iter,
                                 prog_bar=prog_bar, pc_file_in=pc_file_in)
        if mesh is None:
            print('No reconstruction for {}'.format(pc_file_in))
            return 0
        verts, faces = mesh
        is_dataset = os.path.splitext(str(self.in_file))[1].lower() == '.txt'
        base = os.path.basename(pc_file_in)
        if is_dataset:
            out_file = os.path.join(self.results_dir, self.name, os.path.basename(os.path.dirname(str(self.in_file))),
                                    'meshes', base)
        else:
            # a single file was normalised on load: the mesh goes back to the input frame (source/poco_model.py:256-263)
            pts_np = mesh_utils.load_pts(pc_file_in)[:, :3]
            bb_center, scale = mesh_utils.get_points_normalization_info(pts_np, self.padding_factor)
            verts = mesh_utils.denormalize_points_with_info(verts, bb_center, scale)
            out_file = os.path.join(self.results_dir, base, base + '.ply')
        mesh_utils.write_ply(out_file, verts, faces)
        return 0
