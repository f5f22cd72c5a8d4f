"""Training-step twin of the reference network  --  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

A functional torch restatement of ``PPSurfNetwork.forward`` in TRAIN mode (BatchNorm on batch statistics with running-statistic
updates, the FKAConv ``norm_radius`` update, dropout) plus the cross entropy of ``PocoModel.compute_loss``; torch autograd supplies
the backward pass the CUDA ``*_bwd`` entry points are checked against.  Reference formulation throughout (no algebraic shortcuts):

    PPSurfNetwork.forward / from_latent         source/ppsurf_model.py:70-117
    FKAConvNetwork.forward(spectral_only=True)  source/base/nn.py:508-554
    ResidualBlock.forward, max_pool, interpolate source/base/nn.py:438-450,677-697
    FKAConvLayer.forward (train mode)           source/base/nn.py:592-652   (norm_radius update 608-613)
    InterpAttentionKHeadsNet.forward            source/poco_model.py:381-419
    PointNetfeat / STN / AttentionPoco          source/base/nn.py:305-373,162-190,84-96
    MLP (Linear, BatchNorm1d, ReLU, Dropout)    source/base/nn.py:376-417
    compute_loss                                source/poco_model.py:75-88

Pinned against the UNMODIFIED reference by ``tests/golden/make_golden_train.py`` (loss, logits, sampled gradient entries and
gradient norms of all 135 parameter tensors, updated buffers) -> ``tests/golden/train_step.npz`` and by
``tests/test_oracle_golden.py::test_train_oracle_*``.  Only tests/, smoke() and bench.py's reference legs may import this module.
"""
import collections

import torch
import torch.nn.functional as F

RESBLOCKS = ('resnetb01', 'resnetb10', 'resnetb11', 'resnetb20', 'resnetb21', 'resnetb30', 'resnetb31', 'resnetb40', 'resnetb41')
LEVELS = ((0, 0, 'ids00'), (0, 1, 'ids01'), (1, 1, 'ids11'), (1, 2, 'ids12'), (2, 2, 'ids22'), (2, 3, 'ids23'), (3, 3, 'ids33'),
          (3, 4, 'ids34'), (4, 4, 'ids44'))
MOMENTUM = 0.1


class State:
    """parameters (leaf tensors that require grad) and buffers (updated in place like the reference's modules do)"""

    def __init__(self, state_dict, device='cpu', dtype=torch.float32):
        self.p = collections.OrderedDict()
        self.b = collections.OrderedDict()
        for k, v in state_dict.items():
            t = torch.as_tensor(v).detach().clone().to(device)
            leaf = k.rsplit('.', 1)[1]
            if leaf in ('running_mean', 'running_var', 'norm_radius', 'num_batches_tracked'):
                self.b[k] = t if leaf == 'num_batches_tracked' else t.to(dtype)
            else:
                self.p[k] = t.to(dtype).requires_grad_(True)

    def grads(self):
        return collections.OrderedDict((k, v.grad) for k, v in self.p.items())


def _bn(s, name, x, training):
    """BatchNorm1d over [B,C,L] or [B,C] (momentum 0.1, eps 1e-5, unbiased running variance)"""
    out = F.batch_norm(x, s.b[name + '.running_mean'], s.b[name + '.running_var'], s.p[name + '.weight'], s.p[name + '.bias'],
                       training=training, momentum=MOMENTUM, eps=1e-5)
    if training:
        s.b[name + '.num_batches_tracked'] += 1
    return out


def _conv(s, name, x):
    """1x1 Conv1d / Conv2d / Linear on a channel-first tensor [B,C,...]"""
    w = s.p[name + '.weight']
    w = w.reshape(w.shape[0], -1)
    y = torch.einsum('oc,bc...->bo...', w, x)
    if (name + '.bias') in s.p:
        y = y + s.p[name + '.bias'].view(1, -1, *([1] * (x.dim() - 2)))
    return y


def _gather(data, ids):
    """batch_gather(data [B,C,N], dim=2, ids [B,M,K]) -> [B,C,M,K]"""
    b, c, _ = data.shape
    flat = ids.reshape(b, 1, -1).expand(-1, c, -1)
    return torch.gather(data, 2, flat).view(b, c, *ids.shape[1:])


def fkaconv_layer(s, name, x, pts, support, ids, training, act=F.silu):
    pg = _gather(pts, ids) - support.unsqueeze(3)
    xg = _gather(x, ids)
    dist = torch.sqrt((pg.detach() ** 2).sum(1))
    if training:
        mean_radius = dist.max(2)[0].mean()
        s.b[name + '.norm_radius'] = s.b[name + '.norm_radius'] * (1 - MOMENTUM) + mean_radius * MOMENTUM
    pg = pg / s.b[name + '.norm_radius']
    dw = torch.sigmoid(-s.p[name + '.alpha'] * dist + s.p[name + '.beta'])
    dws = dw.sum(2, keepdim=True)
    dws = dws + (dws == 0) + 1e-6
    dw = (dw / dws * dist.shape[2]).unsqueeze(1)

    def inorm(n, v):
        return F.instance_norm(v, weight=s.p[n + '.weight'], bias=s.p[n + '.bias'], eps=1e-5)

    k = pg.shape[3]
    mat = _conv(s, name + '.fc1', pg)
    mat = act(mat if k == 1 else inorm(name + '.bn1', mat))
    mp1 = torch.max(mat * dw, dim=3, keepdim=True)[0].expand(-1, -1, -1, k)
    mat = torch.cat([mat, mp1], dim=1)
    mat = _conv(s, name + '.fc2', mat)
    mat = act(mat if k == 1 else inorm(name + '.bn2', mat))
    mp2 = torch.max(mat * dw, dim=3, keepdim=True)[0].expand(-1, -1, -1, k)
    mat = torch.cat([mat, mp2], dim=1)
    mat = act(_conv(s, name + '.fc3', mat)) * dw
    feat = torch.matmul(xg.transpose(1, 2), mat.permute(0, 2, 3, 1)).transpose(1, 2)  # [B,Cin,Ns,16]
    w = s.p[name + '.cv.weight']  # [Cout,Cin,1,16]
    return torch.einsum('ocm,bcnm->bon', w[:, :, 0, :], feat)


def residual_block(s, name, x, pts, support, ids, training):
    y = F.relu(_bn(s, name + '.bn0', _conv(s, name + '.cv0', x), training))
    y = F.relu(_bn(s, name + '.bn1', fkaconv_layer(s, name + '.cv1', y, pts, support, ids, training), training))
    y = _bn(s, name + '.bn2', _conv(s, name + '.cv2', y), training)
    short = x
    if (name + '.shortcut.weight') in s.p:
        short = _bn(s, name + '.bn_shortcut', _conv(s, name + '.shortcut', x), training)
    if short.shape[2] != y.shape[2]:
        short = _gather(short, ids).max(dim=3)[0]
    return F.relu(y + short)


def encoder(s, data, training, prefix='encoder'):
    pts = [data['pts']] + [data['support%d' % i] for i in (1, 2, 3, 4)]
    e = prefix + '.'
    x = torch.ones_like(pts[0])
    x0 = F.relu(_bn(s, e + 'bn0', fkaconv_layer(s, e + 'cv0', x, pts[0], pts[0], data['ids00'], training), training))
    feats = [None] * 5
    cur = x0
    for name, (a, c, key) in zip(RESBLOCKS, LEVELS):
        cur = residual_block(s, e + name, cur, pts[a], pts[c], data[key], training)
        feats[c] = cur
    x0, x1, x2, x3, x4 = feats
    x5 = x4.max(dim=2, keepdim=True)[0].expand_as(x4)
    d = F.relu(_bn(s, e + 'bn5', _conv(s, e + 'cv5', torch.cat([x4, x5], dim=1)), training))
    for cv, bn, skip, key in (('cv3d', 'bn3d', x3, 'ids43'), ('cv2d', 'bn2d', x2, 'ids32'), ('cv1d', 'bn1d', x1, 'ids21'),
                              ('cv0d', 'bn0d', x0, 'ids10')):
        up = _gather(d, data[key].clamp_min(0)).squeeze(-1)
        d = F.relu(_bn(s, e + bn, _conv(s, e + cv, torch.cat([up, skip], dim=1)), training))
    return _conv(s, e + 'fcout', d)  # dropout p = 0 in PPSurf's encoder


def projection(s, data, latents, prefix='projection'):
    pts, ids = data['pts'], data['proj_ids']
    qry = data['pts_query']
    if qry.shape[1] != 3:
        qry = qry.transpose(1, 2)
    x = torch.cat([_gather(latents, ids), qry.unsqueeze(3) - _gather(pts, ids)], dim=1)
    for fc in ('fc1', 'fc2', 'fc3'):
        x = F.relu(_conv(s, prefix + '.' + fc, x))
    query = _conv(s, prefix + '.fc_query', x)
    value = _conv(s, prefix + '.fc_value', x)
    att = torch.softmax(query, dim=-1).mean(dim=1)  # [B,Q,k]
    x = torch.matmul(att.unsqueeze(-2), value.permute(0, 2, 3, 1)).squeeze(-2).transpose(1, 2)  # [B,C,Q]
    return _conv(s, prefix + '.fc8', x)


def pointnet(s, x, training, prefix='point_net'):
    """x [M,3,P] -> [M,C]"""
    n = prefix + '.'
    h = F.relu(_bn(s, n + 'bn0a', _conv(s, n + 'conv0a', x), training))
    h = F.relu(_bn(s, n + 'bn0b', _conv(s, n + 'conv0b', h), training))
    t = F.relu(_bn(s, n + 'stn2.bn1', _conv(s, n + 'stn2.conv1', h), training))
    t = F.relu(_bn(s, n + 'stn2.bn2', _conv(s, n + 'stn2.conv2', t), training))
    t = F.relu(_bn(s, n + 'stn2.bn3', _conv(s, n + 'stn2.conv3', t), training))
    t = t.max(dim=2)[0]
    t = F.relu(_bn(s, n + 'stn2.bn4', _conv(s, n + 'stn2.fc1', t), training))
    t = F.relu(_bn(s, n + 'stn2.bn5', _conv(s, n + 'stn2.fc2', t), training))
    t = _conv(s, n + 'stn2.fc3', t) + torch.eye(64, dtype=t.dtype, device=t.device).view(1, -1)
    h = torch.bmm(t.view(-1, 64, 64), h)
    h = F.relu(_bn(s, n + 'bn1', _conv(s, n + 'conv1', h), training))
    h = F.relu(_bn(s, n + 'bn2', _conv(s, n + 'conv2', h), training))
    h = _bn(s, n + 'bn3', _conv(s, n + 'conv3', h), training)
    w = torch.softmax(_conv(s, n + 'att.fc_query', h).squeeze(1), dim=-1)  # [M,P]
    v = _conv(s, n + 'att.fc_value', h)  # [M,C,P]
    return (v * w.unsqueeze(1)).sum(dim=2)


def mlp(s, x, training, dropout=0.3, masks=None, prefix='mlp'):
    """x [M,C]; ``masks``: optional list of two keep-masks [M,C] (bool) replacing the random dropout draw"""
    for i in (0, 1):
        x = F.relu(_bn(s, '{}.layers.{}.1'.format(prefix, i), _conv(s, '{}.layers.{}.0'.format(prefix, i), x), training))
        if training and dropout > 0:
            if masks is not None:
                x = x * masks[i].to(x.dtype) / (1.0 - dropout)
            else:
                x = F.dropout(x, dropout, True)
    return _conv(s, prefix + '.layers.2.0', x)


def forward(s, data, training=True, dropout=0.3, masks=None):
    """``data`` in the reference's layouts (pts [B,3,N], supports, ids int64, pts_query, proj_ids [B,Q,k], pts_local_ps [B,Q,P,3])
    -> logits [B,2,Q]"""
    latents = encoder(s, data, training)
    fp = projection(s, data, latents)  # [B,C,Q]
    loc = data['pts_local_ps']
    b, q = loc.shape[:2]
    fl = pointnet(s, loc.reshape(b * q, loc.shape[2], 3).transpose(1, 2), training).view(b, q, -1)
    feat = fp.transpose(1, 2) + fl
    out = mlp(s, feat.reshape(b * q, -1), training, dropout, masks)
    return out.view(b, q, -1).transpose(1, 2)


def loss_of(logits, occ):
    """compute_loss (source/poco_model.py:75-88): mean cross entropy over all query points"""
    return F.cross_entropy(logits, occ, reduction='none').mean()


def training_step(s, data, dropout=0.3, masks=None):
    """one forward + backward; returns (loss, logits); gradients in ``s.grads()``, updated buffers in ``s.b``"""
    for v in s.p.values():
        v.grad = None
    logits = forward(s, data, True, dropout, masks)
    loss = loss_of(logits, data['occ'])
    loss.backward()
    return loss.detach(), logits.detach()
