"""``PPSurfNetwork`` on the B200 kernels: same constructor, ``forward`` / ``get_latent`` / ``from_latent`` surface, dict
keys and ``state_dict`` layout as the reference network (source/ppsurf_model.py:39-117), so a reference checkpoint loads
with ``strict=True`` and the reference's reconstruction driver can call it unchanged.

The torch modules below are PARAMETER CONTAINERS ONLY (they give every tensor the reference's name and shape); their
``forward`` is never used.  All math runs through the C ABI (``ppsurf_b200.ops``) on point-major fp32 tensors; there
is no torch fallback  --  without the CUDA library or a device every entry point raises.
"""
import math
import typing

import numpy as np
import torch
from torch import nn

from . import ops, packing

try:  # the reference derives everything from LightningModule; keep that when Lightning is installed
    import pytorch_lightning as _pl

    _Base = _pl.LightningModule
except ImportError:  # not installed in the build image
    _Base = nn.Module

ENCODER_LEVELS = ((1, 1), (1, 2), (2, 2), (2, 4), (4, 4), (4, 8), (8, 8), (8, 16), (16, 16))


def _bn(c):
    return nn.BatchNorm1d(c)


class FKAConvLayerParams(_Base):
    """tensors of FKAConvLayer (source/base/nn.py:559-589)"""

    def __init__(self, cin, cout, ks=16):
        super().__init__()
        self.cv = nn.Conv2d(cin, cout, (1, ks), bias=False)
        self.register_buffer('norm_radius', torch.ones(1))
        self.norm_radius_momentum = 0.1  # train-mode update of norm_radius (source/base/nn.py:575,608-613)
        self.alpha = nn.Parameter(torch.ones(1))
        self.beta = nn.Parameter(torch.ones(1))
        self.fc1 = nn.Conv2d(3, ks, 1, bias=False)
        self.fc2 = nn.Conv2d(2 * ks, ks, 1, bias=False)
        self.fc3 = nn.Conv2d(2 * ks, ks, 1, bias=False)
        self.bn1 = nn.InstanceNorm2d(ks, affine=True)
        self.bn2 = nn.InstanceNorm2d(ks, affine=True)


class ResidualBlockParams(_Base):
    """tensors of ResidualBlock (source/base/nn.py:422-436)"""

    def __init__(self, cin, cout):
        super().__init__()
        half = cin // 2
        self.cv0, self.bn0 = nn.Conv1d(cin, half, 1), _bn(half)
        self.cv1, self.bn1 = FKAConvLayerParams(half, half), _bn(half)
        self.cv2, self.bn2 = nn.Conv1d(half, cout, 1), _bn(cout)
        self.shortcut = nn.Conv1d(cin, cout, 1) if cin != cout else nn.Identity()
        self.bn_shortcut = _bn(cout) if cin != cout else nn.Identity()


class FKAConvNetworkParams(_Base):
    """tensors of FKAConvNetwork(segmentation=True) (source/base/nn.py:455-506)"""

    def __init__(self, in_channels, out_channels, hidden=64):
        super().__init__()
        h = hidden
        self.cv0, self.bn0 = FKAConvLayerParams(in_channels, h), _bn(h)
        for (a, b), name in zip(ENCODER_LEVELS, packing.RESBLOCKS):
            setattr(self, name, ResidualBlockParams(a * h, b * h))
        self.cv5, self.bn5 = nn.Conv1d(32 * h, 16 * h, 1), _bn(16 * h)
        self.cv3d, self.bn3d = nn.Conv1d(24 * h, 8 * h, 1), _bn(8 * h)
        self.cv2d, self.bn2d = nn.Conv1d(12 * h, 4 * h, 1), _bn(4 * h)
        self.cv1d, self.bn1d = nn.Conv1d(6 * h, 2 * h, 1), _bn(2 * h)
        self.cv0d, self.bn0d = nn.Conv1d(3 * h, h, 1), _bn(h)
        self.fcout = nn.Conv1d(h, out_channels, 1)


class InterpAttentionParams(nn.Module):
    """tensors of InterpAttentionKHeadsNet (source/poco_model.py:364-379)"""

    def __init__(self, latent, out_channels, k):
        super().__init__()
        self.fc1 = nn.Conv2d(latent + 3, latent, 1)
        self.fc2 = nn.Conv2d(latent, latent, 1)
        self.fc3 = nn.Conv2d(latent, latent, 1)
        self.fc8 = nn.Conv1d(latent, out_channels, 1)
        self.fc_query = nn.Conv2d(latent, 64, 1)
        self.fc_value = nn.Conv2d(latent, latent, 1)
        self.k = k


class _AttentionParams(_Base):
    def __init__(self, c):
        super().__init__()
        self.fc_query = nn.Conv2d(c, 1, 1)
        self.fc_value = nn.Conv2d(c, c, 1)


class _STNParams(_Base):
    def __init__(self, size, dim=64):
        super().__init__()
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(dim, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, size, 1)
        self.fc1, self.fc2, self.fc3 = nn.Linear(size, size // 2), nn.Linear(size // 2, size // 4), nn.Linear(size // 4, dim * dim)
        self.bn1, self.bn2, self.bn3, self.bn4, self.bn5 = _bn(64), _bn(128), _bn(size), _bn(size // 2), _bn(size // 4)


class PointNetfeatParams(_Base):
    """tensors of PointNetfeat(use_point_stn=False, use_feat_stn=True, sym_op='att') (source/base/nn.py:256-301)"""

    def __init__(self, net_size_max, output_size):
        super().__init__()
        self.stn2 = _STNParams(net_size_max)
        self.conv0a, self.conv0b = nn.Conv1d(3, 64, 1), nn.Conv1d(64, 64, 1)
        self.bn0a, self.bn0b = _bn(64), _bn(64)
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(64, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, output_size, 1)
        self.bn1, self.bn2, self.bn3 = _bn(64), _bn(128), _bn(output_size)
        self.att = _AttentionParams(output_size)


class MLPParams(_Base):
    """tensors of MLP(num_layers=3, halving_size=False, dropout=0.3) (source/base/nn.py:377-413)"""

    def __init__(self, size, out):
        super().__init__()
        blocks = [nn.Sequential(nn.Linear(size, size), _bn(size), nn.ReLU(), nn.Dropout(0.3)) for _ in range(2)]
        self.layers = nn.Sequential(*blocks, nn.Sequential(nn.Linear(size, out)))


# ---------------------------------------------------------------------------------------------------------------------


def _pm(t: torch.Tensor) -> torch.Tensor:
    """reference ``[B,C,N]`` -> point-major contiguous ``[B,N,C]`` fp32"""
    return t.to(torch.float32).transpose(1, 2).contiguous()


def _ids32(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.int32).contiguous()


class PPSurfNetwork(_Base):

    def __init__(self, in_channels, latent_size, out_channels, k, num_pts_local, pointnet_latent_size,
                 decode_chunk=ops.DEFAULT_CHUNK, decode_path=None):
        super().__init__()
        self.latent_size = latent_size
        self.k = k
        self.num_pts_local = num_pts_local
        self.encoder = FKAConvNetworkParams(in_channels, latent_size)
        self.projection = InterpAttentionParams(latent_size, latent_size, k)
        self.point_net = PointNetfeatParams(pointnet_latent_size, latent_size)
        self.mlp = MLPParams(latent_size, out_channels)
        self.lcp_preprocess = True
        self.decode_chunk = decode_chunk
        # The decoder kernels (both paths) are built for the PPSurf configuration: latent 256, 64 attention heads, k <= 64
        # (configs/ppsurf.yaml:7-9, configs/poco.yaml:47); other sizes fail HERE rather than at the first decode.  Path 1 = tcgen05
        # split-fp16 kernels (k == 64), path 0 = fp32 SIMT kernels (any k <= 64; the reference path of the parity tests).  A cloud needs
        # at least max(k, num_pts_local) points (the reference would clamp k to the cloud size, source/poco_utils.py:259-260; a
        # 64-neighbour interpolation of fewer than 64 points is not a case its configs produce).
        if latent_size != 256 or pointnet_latent_size != 256 or k > 64 or k < 1:
            raise ValueError('ppsurf_b200 is built for network_latent_size = pointnet_latent_size = 256 and k <= 64 '
                             '(got {}, {}, {})'.format(latent_size, pointnet_latent_size, k))
        self.decode_path = (1 if k == 64 else 0) if decode_path is None else decode_path
        self.sampling_seed = None  # set for reproducible support sampling
        self.use_graphs = True     # CUDA-graph replay of the latent loop's batches (latents_of_batch)
        self.graph_seed = 12345    # baked into the captured launches; the per-round rotations re-randomise it on the device
        self._graphs = {}
        self._packed = None
        self._decoder_cache = None
        self.register_load_state_dict_post_hook(lambda module, _keys: module.invalidate())

    # ---- packed weights ------------------------------------------------------------------------------------------
    def invalidate(self):
        self._packed = None
        self._decoder_cache = None
        self._graphs = {}  # captured graphs hold pointers into the packed weights

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def packed(self):
        dev = self.mlp.layers[2][0].weight.device
        if dev.type != 'cuda':
            raise RuntimeError('PPSurfNetwork (ppsurf_b200) runs on a CUDA device only: move it with .cuda() first')
        if self._packed is None or self._packed['device'] != dev:
            ops.require_device()
            sd = self.state_dict()
            self._packed = {'device': dev,
                            'decoder': packing.pack_decoder(sd, dev, self.k, self.num_pts_local),
                            'encoder': packing.pack_encoder(sd, dev, act='silu')}
        return self._packed

    # ---- encoder ---------------------------------------------------------------------------------------------------
    def _resblock(self, blk, x, pts, support, ids):
        """x [B,Nin,C] -> [B,Ns,C'] (source/base/nn.py:438-450)"""
        b, n_in, c = x.shape
        n_s = support.shape[1]
        y = ops.linear(x.view(b * n_in, c), blk['cv0'].w, blk['cv0'].b, relu=True).view(b, n_in, -1)
        y = ops.fkaconv(blk['cv1'], y, pts, support, ids)  # bn1 + ReLU folded
        short = x
        if 'shortcut' in blk:
            short = ops.linear(x.view(b * n_in, c), blk['shortcut'].w, blk['shortcut'].b).view(b, n_in, -1)
        if n_s != n_in:
            short = ops.gather_max(short, ids)
        cout = blk['cv2'].w.shape[0]
        out = ops.linear(y.view(b * n_s, -1), blk['cv2'].w, blk['cv2'].b, residual=short.reshape(b * n_s, cout), relu=True)
        return out.view(b, n_s, cout)

    def encode(self, data: dict) -> torch.Tensor:
        """FKAConvNetwork.forward(spectral_only=True) (source/base/nn.py:508-548): needs ``pts``, ``support1-4`` and the
        13 index tensors in ``data`` (reference layouts); returns point-major latents ``[B,N0,latent]``."""
        pts = [_pm(data['pts'])] + [_pm(data['support%d' % i]) for i in (1, 2, 3, 4)]
        ids = {key: _ids32(val) for key, val in data.items() if key.startswith('ids')}
        return self._encode_pm(pts, ids)

    def _encode_pm(self, pts: list, ids: dict) -> torch.Tensor:
        """the encoder on point-major tensors: ``pts[l] [B,N_l,3]`` for the five levels, ``ids`` int32 (the C ABI's layouts)"""
        enc = self.packed()['encoder']
        b, n0, _ = pts[0].shape
        x = torch.ones_like(pts[0])  # nn.py:517
        cin0 = enc['cv0'].struct.cin  # the packed layer pads its 3 input channels to 4 (zero weights)
        if cin0 != x.shape[-1]:
            x = torch.nn.functional.pad(x, (0, cin0 - x.shape[-1]))
        x0 = ops.fkaconv(enc['cv0'], x, pts[0], pts[0], ids['ids00'])  # bn0 + ReLU folded (nn.py:519)
        x0 = self._resblock(enc['resnetb01'], x0, pts[0], pts[0], ids['ids00'])
        x1 = self._resblock(enc['resnetb10'], x0, pts[0], pts[1], ids['ids01'])
        x1 = self._resblock(enc['resnetb11'], x1, pts[1], pts[1], ids['ids11'])
        x2 = self._resblock(enc['resnetb20'], x1, pts[1], pts[2], ids['ids12'])
        x2 = self._resblock(enc['resnetb21'], x2, pts[2], pts[2], ids['ids22'])
        x3 = self._resblock(enc['resnetb30'], x2, pts[2], pts[3], ids['ids23'])
        x3 = self._resblock(enc['resnetb31'], x3, pts[3], pts[3], ids['ids33'])
        x4 = self._resblock(enc['resnetb40'], x3, pts[3], pts[4], ids['ids34'])
        x4 = self._resblock(enc['resnetb41'], x4, pts[4], pts[4], ids['ids44'])

        # U-Net decoder on the flattened batch: the 1-NN up-sampling gathers rows of the previous (deeper) level, so the row
        # indices of sample s are offset by s * N_deeper (interpolate(): ids < 0 -> 0, k = 1; nn.py:684-697)
        n4 = x4.shape[1]
        wa, wb = enc['cv5']
        glob = ops.linear(ops.global_max(x4), wb.w, wb.b)  # [B,1024]: W5b . max + b  (x4d_bug_fixed=True)
        rows4 = torch.arange(b, device=x4.device, dtype=torch.int32).repeat_interleave(n4)
        deep = ops.linear(x4.reshape(b * n4, -1), wa.w, residual=ops.gather_rows(glob, rows4), relu=True)
        n_deep = n4
        for stage, skip, key in (('cv3d', x3, 'ids43'), ('cv2d', x2, 'ids32'), ('cv1d', x1, 'ids21'), ('cv0d', x0, 'ids10')):
            wa, wb = enc[stage]
            n_l = skip.shape[1]
            up = ids[key].reshape(b, n_l).clamp_min(0) + (torch.arange(b, device=x4.device, dtype=torch.int32) * n_deep)[:, None]
            t = ops.linear(deep, wa.w, gather=up.reshape(-1).contiguous())
            deep = ops.linear(skip.reshape(b * n_l, -1), wb.w, wb.b, residual=t, relu=True)
            n_deep = n_l
        return ops.linear(deep, enc['fcout'].w, enc['fcout'].b).view(b, n0, -1)

    def spatial_ids_pm(self, pts: torch.Tensor, rot: torch.Tensor = None, seed: int = None) -> dict:
        """get_fkaconv_ids on the device, C-ABI layouts: ``pts [B,N0,3]`` -> point-major supports ``support1..4 [B,N_l,3]`` and int32
        index tensors.  ``rot [B,4,ROUNDS,9]`` device rotations (drawn from ``sampling_seed`` when absent)."""
        from .sampling import ROUNDS, random_rotations
        if rot is None:
            gen = np.random.default_rng(self.sampling_seed)
            b = pts.shape[0]
            rot = torch.from_numpy(random_rotations(gen, b * 4 * ROUNDS).reshape(b, 4, ROUNDS, 9)).to(pts.device)
            seed = int(gen.integers(0, 2 ** 31))
        return ops.encoder_ids(pts, rot, 0 if seed is None else seed)

    def spatial_ids(self, pts_bcn: torch.Tensor) -> dict:
        """get_fkaconv_ids on the device (source/poco_data_loader.py:137-209): four quantised support samplings at
        ratio 1/4 and the 13 kNN index tensors in ONE C-ABI call per batch; reference layouts on return (supports
        [B,3,Ns], ids int64)."""
        res = self.spatial_ids_pm(_pm(pts_bcn))
        out = {}
        for key, val in res.items():
            out[key] = val.transpose(1, 2).contiguous() if key.startswith('support') else val.long()
        return out

    def latents_of_batch(self, pts: torch.Tensor, rot: torch.Tensor) -> torch.Tensor:
        """``get_latent`` of the latent loop (source/poco_model.py:227) for a batch of sub-clouds ``pts [B,n,3]`` with the device
        rotations ``rot [B,4,ROUNDS,9]`` of their support samplings -> point-major latents ``[B,n,latent]``.  With ``use_graphs`` the
        whole batch (about 6000 small launches: samplings, radix sorts, index builds, 13 kNN queries per sub-cloud, the network) is
        captured ONCE per shape into a CUDA graph and replayed: the result lives in the graph's static output buffer and is valid
        until the next call with the same shape."""
        # one slot per (shape, stream): the latent loop alternates between two streams so that two batches are in flight
        key = (tuple(pts.shape), pts.device, torch.cuda.current_stream().cuda_stream)
        slot = self._graphs.get(key)
        if not self.use_graphs or slot is None:
            # eager: graphs off, or the first batch of this shape (a shape is captured when it comes back: the ragged last batch of
            # a cloud is not worth a capture unless a second cloud of the same size follows)
            if self.use_graphs:
                self._graphs[key] = {'pts': torch.empty_like(pts), 'rot': torch.empty_like(rot), 'graph': None, 'out': None}
            res = self.spatial_ids_pm(pts, rot, self.graph_seed)
            return self._encode_pm([pts] + [res['support%d' % i] for i in (1, 2, 3, 4)], res)
        slot['pts'].copy_(pts)
        slot['rot'].copy_(rot)
        if slot['graph'] is None:
            def body():
                res = self.spatial_ids_pm(slot['pts'], slot['rot'], self.graph_seed)
                return self._encode_pm([slot['pts']] + [res['support%d' % i] for i in (1, 2, 3, 4)], res)
            graph = torch.cuda.CUDAGraph()  # the eager first use of the shape was the warm-up (function attributes, cub temp sizes)
            # thread-local capture mode: the latent loop's producer thread pins host buffers (cudaHostAlloc) while this thread captures
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                slot['out'] = body()
            slot['graph'] = graph
        slot['graph'].replay()
        return slot['out']

    # ---- reference surface -------------------------------------------------------------------------------------
    def forward(self, data):
        """train/test path (source/ppsurf_model.py:70-74): ids and ``proj_ids`` come with the batch.  In TRAIN mode (``.train()``)
        the step runs through the autograd Functions of ``ppsurf_b200.training`` (batch-statistic BatchNorm, norm_radius update,
        dropout, a backward pass); in eval mode through the packed predict kernels."""
        self._decoder_cache = None
        if self.training:
            from . import training
            if 'proj_ids' not in data:  # the reference recomputes them in from_latent (ppsurf_model.py:83, has_proj_ids=False)
                pts_pm, qry = _pm(data['pts']), data['pts_query'].to(data['pts'].device, torch.float32)
                qry = qry if qry.shape[-1] == 3 else qry.transpose(1, 2)
                data['proj_ids'] = torch.stack([ops.knn(pts_pm[s], qry[s].contiguous(), self.k) for s in range(pts_pm.shape[0])]).long()
            self.invalidate()  # the optimiser is about to change the parameters the packed predict weights were built from
            return training.forward(self, data, training=True)
        data['latents'] = self.encode(data).transpose(1, 2)
        return self.from_latent(data, has_proj_ids='proj_ids' in data)

    def get_latent(self, data):
        """source/ppsurf_model.py:76-80; adds the supports / ids to ``data`` like the reference does"""
        for key, val in self.spatial_ids(data['pts']).items():
            data[key] = val
        data['latents'] = self.encode(data).transpose(1, 2)  # [B,latent,N] view of the point-major result
        data['proj_correction'] = None
        return data

    def decoder_for(self, pts_bcn: torch.Tensor, latents_bcn: torch.Tensor, sample: int = 0) -> ops.Decoder:
        """per-cloud decoder state (kNN index + fc1 table), cached while THE SAME tensor objects are passed again (the
        reference driver reuses one dict for every batch of a cloud, source/poco_utils.py:220-223).  The entry keeps strong
        references to its source tensors and compares by identity and version counter: a recycled device address of a freed
        tensor can therefore never alias another cloud.  ``forward`` and ``PPSurfModel.reconstruct`` drop the cache."""
        entry = self._decoder_cache
        if (entry is None or entry['pts'] is not pts_bcn or entry['latents'] is not latents_bcn
                or entry['versions'] != (pts_bcn._version, latents_bcn._version)
                or entry['key'] != (sample, self.decode_chunk, self.decode_path)):
            pts = pts_bcn[sample].to(torch.float32).transpose(0, 1).contiguous()
            lat = latents_bcn[sample].to(torch.float32).transpose(0, 1).contiguous()
            dec = ops.Decoder(self.packed()['decoder'], pts, lat, chunk=self.decode_chunk, path=self.decode_path)
            entry = {'pts': pts_bcn, 'latents': latents_bcn, 'versions': (pts_bcn._version, latents_bcn._version),
                     'key': (sample, self.decode_chunk, self.decode_path), 'decoder': dec}
            self._decoder_cache = entry
        return entry['decoder']

    def from_latent(self, data: typing.Dict[str, torch.Tensor], has_proj_ids: bool = False) -> torch.Tensor:
        """source/ppsurf_model.py:82-117.  ``data``: ``pts [B,3,N]``, ``latents [B,C,N]``, ``pts_query [B,Q,3]`` (CPU or
        device), optional ``pts_local_ps [B,Q,P,3]`` / ``proj_ids [B,Q,k]``.  Returns logits ``[B,2,Q]`` and, like the
        reference, stores ``proj_ids`` (int64) in ``data``.  When the caller supplies patches (the reference driver
        computes them on the CPU) they are used as given; otherwise neighbours, patches and both branches come from
        one fused ``pps_decoder_decode`` call."""
        pts, latents = data['pts'], data['latents']
        dev = pts.device
        if pts.shape[1] != 3:
            raise ValueError("from_latent: 'pts' must be [B,3,N] like the reference's, got {}".format(tuple(pts.shape)))
        if 'pts_local_ps' in data:
            loc = data['pts_local_ps']
            if loc.dim() != 4 or loc.shape[2] != self.num_pts_local or loc.shape[3] != 3:
                raise ValueError("from_latent: 'pts_local_ps' must be [B,Q,{},3] (num_pts_local of this network), got {}".format(
                    self.num_pts_local, tuple(loc.shape)))
        pts_query = data['pts_query'].to(dev, torch.float32)
        if pts_query.dim() == 2:
            pts_query = pts_query.unsqueeze(0)
        if pts_query.shape[-1] != 3:
            pts_query = pts_query.transpose(1, 2)
        outs, proj_ids = [], []
        for s in range(pts.shape[0]):
            dec = self.decoder_for(pts, latents, s)
            q = pts_query[s].contiguous()
            if 'pts_local_ps' in data:
                if has_proj_ids:
                    idx = _ids32(data['proj_ids'][s])
                else:
                    idx = dec.index.query(q, self.k)
                feat_proj = dec.projection(q, idx)
                feat_pn = ops.pointnet(dec.packed, data['pts_local_ps'][s].to(dev, torch.float32).contiguous(), self.decode_path)
                p = dec.packed.tensors
                # mlp.layers.0 on (feat_proj + feat_pn): the sum of the branches rides on the layer's linearity
                t = ops.linear(feat_proj, p['m0_w'])
                h = ops.linear(feat_pn, p['m0_w'], p['m0_b'], residual=t, relu=True)
                h = ops.linear(h, p['m1_w'], p['m1_b'], relu=True)
                outs.append(ops.linear(h, p['m2_w'], p['m2_b']))
                proj_ids.append(idx[:, :self.k])
            else:
                res = dec.decode(q, want_logits=True, want_idx=True)
                outs.append(res['logits'])
                proj_ids.append(res['idx'][:, :self.k])
        if not has_proj_ids:
            data['proj_ids'] = torch.stack(proj_ids, dim=0).long()
        return torch.stack(outs, dim=0).transpose(1, 2)
