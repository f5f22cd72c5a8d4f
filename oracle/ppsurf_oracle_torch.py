"""Multi-threaded CPU timing twin of the decode oracle  --  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Same reference formulation as ``ppsurf_oracle.from_latent`` (no algebraic shortcuts), restated with torch CPU tensor ops
so that every step (GEMMs AND the element-wise work) uses all host threads the way the reference's own torch-eager CPU
path does; numpy's element-wise kernels are single-threaded and would understate the reference.  Used only by bench.py's
``cpu_baseline`` / ``--impl reference`` legs; ``tests/test_oracle_golden.py`` pins it to the numpy oracle and the golden
vectors.  Citations as in ``ppsurf_oracle.py`` (source/poco_model.py:381-419, source/base/nn.py:305-373,162-190,84-96,
376-417, source/ppsurf_model.py:82-117, source/poco_utils.py:74-82).
"""
import numpy as np
import torch

BN_EPS = 1e-5


def _t(p, name):
    v = p[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def to_device(p, device):
    """the state dict as tensors on ``device`` (bench.py's reference-on-GPU arm: the reference modules moved with .cuda())"""
    return {k: _t(p, k).to(device) for k in p}


def _lin(p, name, x):
    """channel-last pointwise layer: x [..., Cin] -> [..., Cout]"""
    w = _t(p, name + '.weight')
    y = x @ w.reshape(w.shape[0], -1).T
    return y + _t(p, name + '.bias') if (name + '.bias') in p else y


def _bn(p, name, x):
    s = _t(p, name + '.weight') / torch.sqrt(_t(p, name + '.running_var') + BN_EPS)
    return (x - _t(p, name + '.running_mean')) * s + _t(p, name + '.bias')


def from_latent(p, pts, latents, queries, proj_ids, patches):
    """pts [N,3], latents [N,C], queries [Q,3], proj_ids [Q,k] int64, patches [Q,P,3] (all CPU float32 tensors)
    -> occupancy [Q] = softmax(logits)[0] - softmax(logits)[1]"""
    with torch.inference_mode():
        x = torch.cat([latents[proj_ids], queries[:, None, :] - pts[proj_ids]], dim=2)  # [Q,k,C+3]
        for fc in ('fc1', 'fc2', 'fc3'):
            x = torch.relu(_lin(p, 'projection.' + fc, x))
        query = _lin(p, 'projection.fc_query', x)  # [Q,k,64]
        value = _lin(p, 'projection.fc_value', x)  # [Q,k,C]
        att = torch.softmax(query, dim=1).mean(dim=2)  # [Q,k]
        feat_proj = _lin(p, 'projection.fc8', (att[:, :, None] * value).sum(dim=1))

        n = 'point_net.'
        h = torch.relu(_bn(p, n + 'bn0a', _lin(p, n + 'conv0a', patches)))
        h = torch.relu(_bn(p, n + 'bn0b', _lin(p, n + 'conv0b', h)))  # [Q,P,64]
        t = torch.relu(_bn(p, n + 'stn2.bn1', _lin(p, n + 'stn2.conv1', h)))
        t = torch.relu(_bn(p, n + 'stn2.bn2', _lin(p, n + 'stn2.conv2', t)))
        t = torch.relu(_bn(p, n + 'stn2.bn3', _lin(p, n + 'stn2.conv3', t))).max(dim=1).values
        t = torch.relu(_bn(p, n + 'stn2.bn4', _lin(p, n + 'stn2.fc1', t)))
        t = torch.relu(_bn(p, n + 'stn2.bn5', _lin(p, n + 'stn2.fc2', t)))
        t = (_lin(p, n + 'stn2.fc3', t) + torch.eye(64, device=t.device, dtype=t.dtype).reshape(1, -1)).view(-1, 64, 64)
        h = torch.einsum('qij,qpj->qpi', t, h)
        h = torch.relu(_bn(p, n + 'bn1', _lin(p, n + 'conv1', h)))
        h = torch.relu(_bn(p, n + 'bn2', _lin(p, n + 'conv2', h)))
        h = _bn(p, n + 'bn3', _lin(p, n + 'conv3', h))  # [Q,P,C]
        w = torch.softmax(_lin(p, n + 'att.fc_query', h)[..., 0], dim=1)
        feat_pn = (w[:, :, None] * _lin(p, n + 'att.fc_value', h)).sum(dim=1)

        f = feat_proj + feat_pn
        for i in (0, 1):
            f = torch.relu(_bn(p, 'mlp.layers.{}.1'.format(i), _lin(p, 'mlp.layers.{}.0'.format(i), f)))
        logits = _lin(p, 'mlp.layers.2.0', f)
        s = torch.softmax(logits, dim=1)
        return s[:, 0] - s[:, 1], logits
