"""Train-mode forward of ``PPSurfNetwork`` composed from the autograd Functions of ``ppsurf_b200.autograd`` (BASELINE config 5).

Mirrors the reference modules one to one (source/base/nn.py, source/poco_model.py, source/ppsurf_model.py; line numbers at each
function) but on row-major ``[rows, channels]`` activations, with the parameters read straight from the reference-named
``nn.Parameter`` containers of ``ppsurf_b200.network`` -- nothing is packed or folded in train mode: BatchNorm needs its batch
statistics, and the optimiser updates the parameters between steps.

Exact reformulations kept from the predict path (the loss and every parameter gradient are unchanged in real arithmetic):
  * fc1 of the projection acts on a per-POINT table ``latents . W_lat^T`` that is gathered per (query, neighbour) pair, plus
    ``(q - p_j) . W_xyz^T + b``  (instead of a GEMM over 259-wide gathered rows);
  * both attention poolings pool first and apply ``fc_value`` to the pooled vector (the weights of a query sum to one).
"""
import torch

from . import autograd as ag
from . import packing

LEVELS = ((0, 0, 'ids00'), (0, 1, 'ids01'), (1, 1, 'ids11'), (1, 2, 'ids12'), (2, 2, 'ids22'), (2, 3, 'ids23'), (3, 3, 'ids33'),
          (3, 4, 'ids34'), (4, 4, 'ids44'))


CONCURRENT_BRANCHES = True
_side_streams = {}


def _side_stream(device):
    key = str(device)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def _w2(conv):
    """weight of a 1x1 Conv1d / Conv2d / Linear as a ``[out, in]`` view"""
    w = conv.weight
    return w.view(w.shape[0], -1)


def _bn(x2, bn, act, training):
    """BatchNorm1d over all rows of ``x2 [rows, C]`` (batch statistics in train mode, running statistics otherwise)"""
    if training:
        y = ag.norm(x2.unsqueeze(0), bn.weight, bn.bias, act, bn.eps, (bn.running_mean, bn.running_var), bn.momentum)
        bn.num_batches_tracked += 1
        return y.squeeze(0)
    raise NotImplementedError('eval-mode BatchNorm runs on the packed predict kernels (PPSurfNetwork.forward in eval mode)')


def fkaconv_layer(layer, x, pts, support, ids, training, act='silu'):
    """FKAConvLayer.forward (nn.py:592-652): ``x [B,Nin,Cin]``, ``pts [B,Nin,3]``, ``support [B,Ns,3]``, ``ids [B,Ns,K]`` int32 ->
    ``[B*Ns, Cout]``"""
    b, n_s, kn = ids.shape
    g = b * n_s
    offs, dw = ag.FkaGeometry.apply(layer.alpha, layer.beta, pts, support, ids, layer.norm_radius, training, layer.norm_radius_momentum)

    def inorm(t, bn):  # InstanceNorm2d over (Ns, K) per sample and channel; a single neighbour skips it (nn.py:627-630)
        if kn == 1:
            return ag.act(t, act)
        return ag.norm(t.view(b, n_s * kn, 16), bn.weight, bn.bias, act, bn.eps).view(g * kn, 16)

    mat = inorm(ag.linear(offs, _w2(layer.fc1)), layer.bn1)
    mp = ag.SegMax.apply(mat.view(g, kn, 16), dw.view(g, kn))
    mat = ag.ConcatBcast.apply(mat.view(g, kn, 16), mp).view(g * kn, 32)
    mat = inorm(ag.linear(mat, _w2(layer.fc2)), layer.bn2)
    mp = ag.SegMax.apply(mat.view(g, kn, 16), dw.view(g, kn))
    mat = ag.ConcatBcast.apply(mat.view(g, kn, 16), mp).view(g * kn, 32)
    mat = ag.RowScale.apply(ag.act(ag.linear(mat, _w2(layer.fc3)), act), dw)
    feat = ag.FkaFeat.apply(x, mat, ids)
    return ag.linear(feat, _w2(layer.cv))


def residual_block(blk, x, pts, support, ids, training):
    """ResidualBlock.forward (nn.py:438-450): ``x [B,Nin,C]`` -> ``[B,Ns,C']``"""
    b, n_in, c = x.shape
    n_s = support.shape[1]
    y = _bn(ag.linear(x.reshape(b * n_in, c), _w2(blk.cv0), blk.cv0.bias), blk.bn0, 'relu', training)
    y = fkaconv_layer(blk.cv1, y.view(b, n_in, -1), pts, support, ids, training)
    y = _bn(y, blk.bn1, 'relu', training)
    y = _bn(ag.linear(y, _w2(blk.cv2), blk.cv2.bias), blk.bn2, None, training)
    short = x
    if not isinstance(blk.shortcut, torch.nn.Identity):
        short = _bn(ag.linear(x.reshape(b * n_in, c), _w2(blk.shortcut), blk.shortcut.bias), blk.bn_shortcut, None, training).view(b, n_in, -1)
    if n_s != n_in:
        short = ag.GatherMax.apply(short, ids)
    return ag.act(y + short.reshape(b * n_s, -1), 'relu').view(b, n_s, -1)


def encoder(enc, pts, ids, training):
    """FKAConvNetwork.forward(spectral_only=True) (nn.py:508-548): ``pts[l] [B,N_l,3]`` of the five levels, ``ids`` int32 -> latents
    ``[B*N0, latent]``"""
    b, n0, _ = pts[0].shape
    x = torch.ones_like(pts[0])
    x0 = _bn(fkaconv_layer(enc.cv0, x, pts[0], pts[0], ids['ids00'], training), enc.bn0, 'relu', training).view(b, n0, -1)
    feats = [None] * 5
    cur = x0
    for name, (a, c, key) in zip(packing.RESBLOCKS, LEVELS):
        cur = residual_block(getattr(enc, name), cur, pts[a], pts[c], ids[key], training)
        feats[c] = cur
    x0, x1, x2, x3, x4 = feats
    n4, c4 = x4.shape[1], x4.shape[2]
    x5 = ag.SegMax.apply(x4, None)  # [B,C4]: max over the points of a sample (nn.py:535)
    rows4 = torch.arange(b, device=x4.device, dtype=torch.int32).repeat_interleave(n4)
    cat = torch.cat([x4.reshape(b * n4, c4), ag.GatherRows.apply(x5, rows4)], dim=1)
    deep = _bn(ag.linear(cat, _w2(enc.cv5), enc.cv5.bias), enc.bn5, 'relu', training)
    n_deep = n4
    for cv, bn, skip, key in (('cv3d', 'bn3d', x3, 'ids43'), ('cv2d', 'bn2d', x2, 'ids32'), ('cv1d', 'bn1d', x1, 'ids21'),
                              ('cv0d', 'bn0d', x0, 'ids10')):
        n_l = skip.shape[1]
        up = ids[key].reshape(b, n_l).clamp_min(0) + (torch.arange(b, device=x4.device, dtype=torch.int32) * n_deep)[:, None]
        cat = torch.cat([ag.GatherRows.apply(deep, up.reshape(-1).contiguous()), skip.reshape(b * n_l, -1)], dim=1)
        conv = getattr(enc, cv)
        deep = _bn(ag.linear(cat, _w2(conv), conv.bias), getattr(enc, bn), 'relu', training)
        n_deep = n_l
    return ag.linear(deep, _w2(enc.fcout), enc.fcout.bias)


def projection(proj, latents, pts, qry, proj_ids):
    """InterpAttentionKHeadsNet.forward (poco_model.py:381-419): ``latents [B*N,C]``, ``pts [B,N,3]``, ``qry [B,Q,3]``, ``proj_ids [B,Q,k]``
    int32 -> ``[B*Q, C]``"""
    b, n, _ = pts.shape
    q, k = proj_ids.shape[1], proj_ids.shape[2]
    c = latents.shape[1]
    rows = (proj_ids + (torch.arange(b, device=pts.device, dtype=torch.int32) * n)[:, None, None]).reshape(-1).contiguous()
    w1 = _w2(proj.fc1)
    table = ag.linear(latents, w1[:, :c])  # fc1's latent columns once per point
    with torch.no_grad():
        offs = qry.reshape(b * q, 1, 3) - ag.GatherRows.apply(pts.reshape(b * n, 3), rows).view(b * q, k, 3)
    h = ag.GatherRows.apply(table, rows) + ag.linear(offs.reshape(-1, 3), w1[:, c:], proj.fc1.bias)
    h = ag.act(h, 'relu')
    h = ag.act(ag.linear(h, _w2(proj.fc2), proj.fc2.bias), 'relu')
    h = ag.act(ag.linear(h, _w2(proj.fc3), proj.fc3.bias), 'relu')
    scores = ag.linear(h, _w2(proj.fc_query), proj.fc_query.bias)
    pooled = ag.AttnPool.apply(scores.view(b * q, k, -1), h.view(b * q, k, c))
    val = ag.linear(pooled, _w2(proj.fc_value), proj.fc_value.bias)
    return ag.linear(val, _w2(proj.fc8), proj.fc8.bias)


def pointnet(pn, patches, training):
    """PointNetfeat.forward(use_feat_stn, sym_op='att') (nn.py:305-373,162-190,84-96): ``patches [M,P,3]`` -> ``[M, C]``"""
    m, p, _ = patches.shape

    def layer(x, conv, bn, act_name='relu'):
        return _bn(ag.linear(x, _w2(conv), conv.bias), bn, act_name, training)

    h = layer(patches.reshape(m * p, 3), pn.conv0a, pn.bn0a)
    h = layer(h, pn.conv0b, pn.bn0b)
    stn = pn.stn2
    t = layer(h, stn.conv1, stn.bn1)
    t = layer(t, stn.conv2, stn.bn2)
    t = layer(t, stn.conv3, stn.bn3)
    t = ag.SegMax.apply(t.view(m, p, -1), None)
    t = layer(t, stn.fc1, stn.bn4)
    t = layer(t, stn.fc2, stn.bn5)
    t = ag.linear(t, stn.fc3.weight, stn.fc3.bias) + torch.eye(64, dtype=torch.float32, device=t.device).view(1, -1)
    h = ag.Bmm.apply(h.view(m, p, 64), t.view(m, 64, 64).transpose(1, 2)).view(m * p, 64)  # x <- T x, point-major
    h = layer(h, pn.conv1, pn.bn1)
    h = layer(h, pn.conv2, pn.bn2)
    h = layer(h, pn.conv3, pn.bn3, None)
    scores = ag.linear(h, _w2(pn.att.fc_query), pn.att.fc_query.bias)
    pooled = ag.AttnPool.apply(scores.view(m, p, 1), h.view(m, p, -1))
    return ag.linear(pooled, _w2(pn.att.fc_value), pn.att.fc_value.bias)


def mlp(net, x, training):
    """MLP.forward (nn.py:376-417): Linear, BatchNorm1d, ReLU, Dropout twice, then Linear"""
    for i in (0, 1):
        blk = net.layers[i]
        x = _bn(ag.linear(x, blk[0].weight, blk[0].bias), blk[1], 'relu', training)
        if training and blk[3].p > 0:
            x = ag.Dropout.apply(x, float(blk[3].p))
    return ag.linear(x, net.layers[2][0].weight, net.layers[2][0].bias)


def forward(network, data, training=True):
    """PPSurfNetwork.forward (ppsurf_model.py:70-117) in the reference's layouts: ``pts [B,3,N]``, ``support1-4``, ``ids*``,
    ``pts_query``, ``proj_ids [B,Q,k]``, ``pts_local_ps [B,Q,P,3]`` -> logits ``[B,2,Q]``; stores ``latents [B,C,N]`` in ``data``."""
    def pm(t):
        return t.to(torch.float32).transpose(1, 2).contiguous()

    pts = [pm(data['pts'])] + [pm(data['support%d' % i]) for i in (1, 2, 3, 4)]
    ids = {key: val.to(torch.int32).contiguous() for key, val in data.items() if key.startswith('ids')}
    b, n, _ = pts[0].shape
    qry = data['pts_query'].to(pts[0].device, torch.float32)
    if qry.shape[-1] != 3:
        qry = qry.transpose(1, 2)
    qry = qry.contiguous()
    proj_ids = data['proj_ids'].to(torch.int32).contiguous()
    loc = data['pts_local_ps'].to(pts[0].device, torch.float32)
    q = qry.shape[1]
    # the local branch does not depend on the encoder: it runs on a side stream next to encoder + projection (autograd replays the
    # same fork / join in the backward pass, and a captured graph keeps the two branches parallel).  The encoder is ~800 small
    # launches over 39..10 000 points that leave most SMs idle; the PointNet layers over B*Q*P rows fill them.
    main = torch.cuda.current_stream()
    side = _side_stream(pts[0].device) if CONCURRENT_BRANCHES else main
    patches = loc.reshape(b * q, loc.shape[2], 3).contiguous()
    if side is not main:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            local = pointnet(network.point_net, patches, training)
    else:
        local = pointnet(network.point_net, patches, training)
    latents = encoder(network.encoder, pts, ids, training)
    data['latents'] = latents.view(b, n, -1).transpose(1, 2)
    glob = projection(network.projection, latents, pts[0], qry, proj_ids)
    if side is not main:
        main.wait_stream(side)  # `local` is consumed below on the main stream; the next side-stream work (the branch's backward) is
        # ordered after it by autograd's own stream synchronisation, so the allocator cannot hand its block out early
    feat = glob + local
    logits = mlp(network.mlp, feat, training)
    return logits.view(b, q, -1).transpose(1, 2)


BATCH_KEYS = ('pts', 'support1', 'support2', 'support3', 'support4', 'ids00', 'ids01', 'ids11', 'ids12', 'ids22', 'ids23', 'ids33', 'ids34',
              'ids44', 'ids43', 'ids32', 'ids21', 'ids10', 'pts_query', 'proj_ids', 'pts_local_ps', 'occ')


def flatten_gradients(params, device=None) -> torch.Tensor:
    """one flat fp32 buffer holding the gradient of every parameter, each ``p.grad`` a view into it (autograd accumulates in place):
    the data-parallel all-reduce is then a single collective.  Works on any device (the gloo tests run it on the CPU)."""
    params = [p for p in params if p.requires_grad]
    device = params[0].device if device is None else device
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=device)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return flat


def average_gradients(flat: torch.Tensor, world: int, group=None):
    """mean over the ranks of the flat gradient buffer, in place (no-op for one rank)"""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)


class GraphedTrainStep:
    """One training step (forward, cross entropy, backward, optimiser update) captured ONCE into CUDA graphs and replayed: the step is
    about 1200 small launches, which a Python thread cannot issue as fast as the device retires them.  All batches must have the shapes
    of the first one (fixed ``manifold_points`` / query count, as in the reference's training configuration).

    Data parallel (``world > 1``, one process per GPU): every parameter gradient is a view into ONE flat buffer; the backward graph
    fills it, a single NCCL all-reduce averages it over the ranks, a second graph runs the optimiser -- the reference's DDP
    (configs/device_server.yaml) with one bucket.  BatchNorm statistics stay per rank like the reference's (no SyncBatchNorm).

    ``optimizer`` must be constructed with ``capturable=True``; do not call ``zero_grad(set_to_none=True)`` afterwards (it would detach
    the gradient views; the step zeroes the flat buffer itself)."""

    def __init__(self, network, optimizer, batch: dict, world: int = 1, warmup: int = 3):
        import torch.distributed as dist
        self.network, self.optimizer, self.world = network, optimizer, world
        self.dist = dist
        dev = batch['pts'].device
        self.static = {k: batch[k].clone() for k in BATCH_KEYS}
        self.flat = flatten_gradients(network.parameters(), dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._forward_backward()
                self._reduce()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        from ._lib import lib
        launched = lib.pps_launch_count()
        self.graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fb):
            self.loss = self._forward_backward()
        self.kernels_per_step = int(lib.pps_launch_count() - launched)  # library kernels recorded in the graph (replayed every step)
        self.graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_opt):
            optimizer.step()

    def _forward_backward(self):
        self.flat.zero_()  # gradients accumulate into the flat buffer (set_to_none would detach the views)
        pred = self.network.forward(dict(self.static))
        b, c, q = pred.shape
        loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(b * q, c), self.static['occ'].reshape(-1))
        loss.backward()
        return loss.detach()

    def _reduce(self):
        average_gradients(self.flat, self.world)

    def __call__(self, batch: dict) -> torch.Tensor:
        for k in BATCH_KEYS:
            if batch[k].shape != self.static[k].shape:
                raise ValueError("GraphedTrainStep was captured for '{}' of shape {}, got {}: every batch must have the shapes of the "
                                 'first one (drop_last, fixed manifold_points / query count)'.format(
                                     k, tuple(self.static[k].shape), tuple(batch[k].shape)))
            self.static[k].copy_(batch[k], non_blocking=True)
        self.graph_fb.replay()
        self._reduce()
        self.graph_opt.replay()
        return self.loss
