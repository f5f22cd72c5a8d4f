"""CPU oracle for the PPSurf occupancy hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  The product path (``ppsurf_b200``) never does and fails loudly without its CUDA library.

This is a plain-numpy restatement of the reference algorithm (cg-tuwien/ppsurf @ 060675d); every function cites
the reference ``file:line`` it follows (paths relative to the reference root).  It is written from the behaviour
of the reference, shares no code with it, and runs in float32 (bit-comparable rounding class with the
reference's CPU path) or float64 (``dtype=np.float64``, used to arbitrate who is closer when fp32 results differ).

Parity pinning: the reference ships no golden vectors or known-answer tests for this path.
The oracle is therefore pinned against outputs of the UNMODIFIED reference run in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``, checked by ``tests/test_oracle_golden.py``).
Two boundaries stay "parity unpinned" because the third-party code is neither vendored nor installable offline:
``pykdtree`` (>=1.3, requirements.txt:17; stand-in: scipy cKDTree in float64) and ``torch_geometric``/
``torch_cluster`` ``voxel_grid`` (requirements.txt:4-5) used by ``sampling_quantized``.

Layouts follow the reference at this API so the tests read like reference calls:
``pts [B,3,N]``, ``latents [B,C,N]``, ``ids [B,Ns,K]`` int64, ``pts_query [B,Q,3]``, ``pts_local_ps [B,Q,P,3]``.
Parameters are passed as a ``dict name -> np.ndarray`` with the reference ``state_dict`` names of
``PPSurfNetwork`` (``encoder.*``, ``projection.*``, ``point_net.*``, ``mlp.*``).
"""
import collections
import math
import typing

import numpy as np

BN_EPS = 1e-5  # torch.nn.BatchNorm1d / InstanceNorm2d default
Params = typing.Dict[str, np.ndarray]


# --------------------------------------------------------------------------------------------------------------
# parameter inventory (source/ppsurf_model.py:39-68, source/base/nn.py:453-506,557-589,133-160,255-301,376-413,
#                      source/poco_model.py:364-379)
# --------------------------------------------------------------------------------------------------------------

def _bn_spec(spec, name, c):
    spec[name + '.weight'] = (c,)
    spec[name + '.bias'] = (c,)
    spec[name + '.running_mean'] = (c,)
    spec[name + '.running_var'] = (c,)
    spec[name + '.num_batches_tracked'] = ()


def _fka_spec(spec, name, cin, cout, ks=16):
    spec[name + '.alpha'] = (1,)
    spec[name + '.beta'] = (1,)
    spec[name + '.norm_radius'] = (1,)
    spec[name + '.cv.weight'] = (cout, cin, 1, ks)
    spec[name + '.fc1.weight'] = (ks, 3, 1, 1)
    spec[name + '.fc2.weight'] = (ks, 2 * ks, 1, 1)
    spec[name + '.fc3.weight'] = (ks, 2 * ks, 1, 1)
    spec[name + '.bn1.weight'] = (ks,)
    spec[name + '.bn1.bias'] = (ks,)
    spec[name + '.bn2.weight'] = (ks,)
    spec[name + '.bn2.bias'] = (ks,)


def _conv1d_spec(spec, name, cin, cout):
    spec[name + '.weight'] = (cout, cin, 1)
    spec[name + '.bias'] = (cout,)


def _conv2d_spec(spec, name, cin, cout):
    spec[name + '.weight'] = (cout, cin, 1, 1)
    spec[name + '.bias'] = (cout,)


def _linear_spec(spec, name, cin, cout):
    spec[name + '.weight'] = (cout, cin)
    spec[name + '.bias'] = (cout,)


def _resblock_spec(spec, name, cin, cout):
    _conv1d_spec(spec, name + '.cv0', cin, cin // 2)
    _bn_spec(spec, name + '.bn0', cin // 2)
    _fka_spec(spec, name + '.cv1', cin // 2, cin // 2)
    _bn_spec(spec, name + '.bn1', cin // 2)
    _conv1d_spec(spec, name + '.cv2', cin // 2, cout)
    _bn_spec(spec, name + '.bn2', cout)
    if cin != cout:
        _conv1d_spec(spec, name + '.shortcut', cin, cout)
        _bn_spec(spec, name + '.bn_shortcut', cout)


RESBLOCKS = (('resnetb01', 1, 1), ('resnetb10', 1, 2), ('resnetb11', 2, 2), ('resnetb20', 2, 4),
             ('resnetb21', 4, 4), ('resnetb30', 4, 8), ('resnetb31', 8, 8), ('resnetb40', 8, 16),
             ('resnetb41', 16, 16))


def param_spec(in_channels=3, latent_size=256, out_channels=2, pointnet_latent_size=256, hidden=64) \
        -> 'collections.OrderedDict[str, tuple]':
    """Names and shapes of ``PPSurfNetwork.state_dict()`` in registration order (455 entries for PPSurf 50NN)."""
    spec = collections.OrderedDict()
    e = 'encoder'
    _fka_spec(spec, e + '.cv0', in_channels, hidden)
    _bn_spec(spec, e + '.bn0', hidden)
    for name, a, b in RESBLOCKS:
        _resblock_spec(spec, '{}.{}'.format(e, name), a * hidden, b * hidden)
    _conv1d_spec(spec, e + '.cv5', 32 * hidden, 16 * hidden)
    _bn_spec(spec, e + '.bn5', 16 * hidden)
    _conv1d_spec(spec, e + '.cv3d', 24 * hidden, 8 * hidden)
    _bn_spec(spec, e + '.bn3d', 8 * hidden)
    _conv1d_spec(spec, e + '.cv2d', 12 * hidden, 4 * hidden)
    _bn_spec(spec, e + '.bn2d', 4 * hidden)
    _conv1d_spec(spec, e + '.cv1d', 6 * hidden, 2 * hidden)
    _bn_spec(spec, e + '.bn1d', 2 * hidden)
    _conv1d_spec(spec, e + '.cv0d', 3 * hidden, hidden)
    _bn_spec(spec, e + '.bn0d', hidden)
    _conv1d_spec(spec, e + '.fcout', hidden, latent_size)

    p = 'projection'
    _conv2d_spec(spec, p + '.fc1', latent_size + 3, latent_size)
    _conv2d_spec(spec, p + '.fc2', latent_size, latent_size)
    _conv2d_spec(spec, p + '.fc3', latent_size, latent_size)
    _conv1d_spec(spec, p + '.fc8', latent_size, latent_size)
    _conv2d_spec(spec, p + '.fc_query', latent_size, 64)
    _conv2d_spec(spec, p + '.fc_value', latent_size, latent_size)

    n = 'point_net'
    s = pointnet_latent_size
    _conv1d_spec(spec, n + '.stn2.conv1', 64, 64)
    _conv1d_spec(spec, n + '.stn2.conv2', 64, 128)
    _conv1d_spec(spec, n + '.stn2.conv3', 128, s)
    _linear_spec(spec, n + '.stn2.fc1', s, s // 2)
    _linear_spec(spec, n + '.stn2.fc2', s // 2, s // 4)
    _linear_spec(spec, n + '.stn2.fc3', s // 4, 64 * 64)
    _bn_spec(spec, n + '.stn2.bn1', 64)
    _bn_spec(spec, n + '.stn2.bn2', 128)
    _bn_spec(spec, n + '.stn2.bn3', s)
    _bn_spec(spec, n + '.stn2.bn4', s // 2)
    _bn_spec(spec, n + '.stn2.bn5', s // 4)
    _conv1d_spec(spec, n + '.conv0a', 3, 64)
    _conv1d_spec(spec, n + '.conv0b', 64, 64)
    _bn_spec(spec, n + '.bn0a', 64)
    _bn_spec(spec, n + '.bn0b', 64)
    _conv1d_spec(spec, n + '.conv1', 64, 64)
    _conv1d_spec(spec, n + '.conv2', 64, 128)
    _conv1d_spec(spec, n + '.conv3', 128, latent_size)
    _bn_spec(spec, n + '.bn1', 64)
    _bn_spec(spec, n + '.bn2', 128)
    _bn_spec(spec, n + '.bn3', latent_size)
    _conv2d_spec(spec, n + '.att.fc_query', latent_size, 1)
    _conv2d_spec(spec, n + '.att.fc_value', latent_size, latent_size)

    m = 'mlp.layers'
    _linear_spec(spec, m + '.0.0', latent_size, latent_size)
    _bn_spec(spec, m + '.0.1', latent_size)
    _linear_spec(spec, m + '.1.0', latent_size, latent_size)
    _bn_spec(spec, m + '.1.1', latent_size)
    _linear_spec(spec, m + '.2.0', latent_size, out_channels)
    return spec


DEFAULT_GAINS = {'encoder': 0.68, 'projection': 1.5, 'point_net': 1.21, 'mlp': 1.2}


def make_state_dict(seed=42, gains: typing.Optional[dict] = None, **spec_kwargs) -> Params:
    """Deterministic synthetic weights (no checkpoint is vendored; SURVEY.md §8d 'weights').

    numpy ``default_rng`` is platform independent, so the container and the GPU box build identical tensors.
    Linear/conv weights ~ U(+-gain*sqrt(3/fan_in)) with a per-sub-network gain chosen so that latents are O(1) and
    logits O(1-10) on the synthetic cloud (the 1e-4 abs tolerance is meaningless on O(0.01) logits), biases ~ U(+-0.1), BatchNorm
    ``running_mean ~ N(0,0.1)``, ``running_var ~ U(0.5,1.5)``, affine ``weight ~ U(0.75,1.25)``, ``bias ~ N(0,0.1)``,
    FKAConv ``alpha,beta ~ U(0.5,1.5)``, ``norm_radius ~ U(0.05,0.2)``  --  non-trivial values everywhere so that
    folding mistakes are visible.
    """
    rng = np.random.default_rng(seed)
    spec = param_spec(**spec_kwargs)
    gains = dict(DEFAULT_GAINS, **(gains or {}))
    sd = collections.OrderedDict()
    for name, shape in spec.items():
        leaf = name.rsplit('.', 1)[1]
        owner = name.rsplit('.', 1)[0]
        parent = owner.rsplit('.', 1)[0] if '.' in owner else ''
        is_norm = (owner + '.running_mean') in spec or \
            (owner.endswith(('.bn1', '.bn2')) and (parent + '.alpha') in spec)  # BatchNorm / FKAConv InstanceNorm
        if leaf == 'num_batches_tracked':
            v = np.array(7, dtype=np.int64)
        elif leaf in ('alpha', 'beta'):
            v = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == 'norm_radius':
            v = rng.uniform(0.05, 0.2, size=shape).astype(np.float32)
        elif leaf == 'running_mean':
            v = (0.1 * rng.standard_normal(size=shape)).astype(np.float32)
        elif leaf == 'running_var':
            v = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif is_norm and leaf == 'weight':
            v = rng.uniform(0.75, 1.25, size=shape).astype(np.float32)
        elif is_norm and leaf == 'bias':
            v = (0.1 * rng.standard_normal(size=shape)).astype(np.float32)
        elif leaf == 'weight':
            fan_in = int(np.prod(shape[1:]))
            bound = gains[name.split('.', 1)[0]] * math.sqrt(3.0 / fan_in)
            v = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        elif leaf == 'bias':
            v = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        else:
            raise KeyError(name)
        sd[name] = v
    return sd


# --------------------------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------------------------

def _w(p: Params, name: str, dtype):
    return np.asarray(p[name]).astype(dtype, copy=False)


def _relu(x):
    return np.maximum(x, 0)


def _sigmoid(x):
    return 1 / (1 + np.exp(-x))


def _silu(x):
    return x * _sigmoid(x)


_ACT = {'relu': _relu, 'silu': _silu}


def batch_gather(data: np.ndarray, index: np.ndarray) -> np.ndarray:
    """``data [B,C,N]``, ``index [B,M,K]`` -> ``[B,C,M,K]``  (source/base/nn.py:655-674, dim=2)."""
    b = np.arange(data.shape[0])[:, None, None]
    return np.moveaxis(data[b, :, index], 3, 1)


def _pointwise(p: Params, name: str, x: np.ndarray, dtype) -> np.ndarray:
    """1x1 Conv1d / Conv2d / Linear over channel axis 1: ``x [B,Cin,...]`` -> ``[B,Cout,...]``."""
    w = _w(p, name + '.weight', dtype)
    w = w.reshape(w.shape[0], -1)
    y = np.moveaxis(np.tensordot(w, x, axes=([1], [1])), 0, 1)  # BLAS gemm: [o,b,...] -> [b,o,...]
    if (name + '.bias') in p:
        bias = _w(p, name + '.bias', dtype)
        y = y + bias.reshape((1, -1) + (1,) * (x.ndim - 2))
    return y


def _batchnorm_eval(p: Params, name: str, x: np.ndarray, dtype) -> np.ndarray:
    """BatchNorm1d in eval mode: running statistics (torch.nn.BatchNorm1d; used at source/base/nn.py:440-444 etc.)."""
    shape = (1, -1) + (1,) * (x.ndim - 2)
    mean = _w(p, name + '.running_mean', dtype).reshape(shape)
    var = _w(p, name + '.running_var', dtype).reshape(shape)
    g = _w(p, name + '.weight', dtype).reshape(shape)
    b = _w(p, name + '.bias', dtype).reshape(shape)
    return (x - mean) / np.sqrt(var + dtype(BN_EPS)) * g + b


def _instancenorm(p: Params, name: str, x: np.ndarray, dtype) -> np.ndarray:
    """InstanceNorm2d(affine=True), per-(sample,channel) biased statistics over (Ns,K) in train AND eval
    (source/base/nn.py:586-587)."""
    mean = x.mean(axis=(2, 3), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(2, 3), keepdims=True)
    g = _w(p, name + '.weight', dtype).reshape(1, -1, 1, 1)
    b = _w(p, name + '.bias', dtype).reshape(1, -1, 1, 1)
    return (x - mean) / np.sqrt(var + dtype(BN_EPS)) * g + b


# --------------------------------------------------------------------------------------------------------------
# a6: exact k nearest neighbours  (source/poco_utils.py:257-273, source/base/proximity.py:40-89)
# --------------------------------------------------------------------------------------------------------------

def sq_dist_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Squared distance in float32 with pykdtree's float path expression ((dx*dx + dy*dy) + dz*dz, each op rounded,
    no FMA).  ``a [...,3]``, ``b [...,3]`` broadcastable."""
    a = a.astype(np.float32, copy=False)
    b = b.astype(np.float32, copy=False)
    d = a - b
    dd = d * d
    return (dd[..., 0] + dd[..., 1]) + dd[..., 2]


def knn(points: np.ndarray, queries: np.ndarray, k: int, chunk: int = 2048) \
        -> typing.Tuple[np.ndarray, np.ndarray]:
    """Exact k-NN, ascending distance, ``k <- min(k, N)`` (source/poco_utils.py:259-260).

    ``points [N,3]``, ``queries [Q,3]`` float32 -> ``idx [Q,k]`` int64, ``dist2 [Q,k]`` float32.
    Brute force (the kd-tree is an accelerator, not part of the contract).  Ties: lower point index first
    (a kd-tree may break exact ties differently; the tests compare index sets modulo equal distances).
    """
    points = np.ascontiguousarray(points, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    n = points.shape[0]
    k = min(k, n)
    q = queries.shape[0]
    idx = np.empty((q, k), dtype=np.int64)
    d2 = np.empty((q, k), dtype=np.float32)
    for s in range(0, q, chunk):
        qs = queries[s:s + chunk]
        dist = sq_dist_f32(qs[:, None, :], points[None, :, :])  # [c,N]
        if k < n:
            part = np.argpartition(dist, k - 1, axis=1)[:, :k]
            kth = np.take_along_axis(dist, part, axis=1).max(axis=1, keepdims=True)
            # all candidates with dist <= kth, then stable sort by (dist, index)
            order = np.empty((qs.shape[0], k), dtype=np.int64)
            for r in range(qs.shape[0]):
                cand = np.nonzero(dist[r] <= kth[r, 0])[0]
                o = np.lexsort((cand, dist[r, cand]))[:k]
                order[r] = cand[o]
        else:
            order = np.stack([np.lexsort((np.arange(n), dist[r])) for r in range(qs.shape[0])], axis=0)
        idx[s:s + chunk] = order
        d2[s:s + chunk] = np.take_along_axis(dist, order, axis=1)
    return idx, d2


def knn_kdtree(points: np.ndarray, queries: np.ndarray, k: int) -> typing.Tuple[np.ndarray, np.ndarray]:
    """kd-tree search as the reference runs it (``make_kdtree`` + ``query_kdtree``, source/base/proximity.py:40-81) with
    scipy's cKDTree(leafsize=10, all cores) standing in for pykdtree, which is not installable offline.  Used by the CPU
    timing legs of bench.py (a brute-force search would misrepresent the reference's cost); the parity tests use the
    exact float32 brute force above."""
    from scipy.spatial import cKDTree
    k = min(k, points.shape[0])
    d, i = cKDTree(points, leafsize=10).query(queries, k=k, workers=-1)
    if k == 1:
        d, i = d[:, None], i[:, None]
    return i.astype(np.int64), (d * d).astype(np.float32)


def knn_batched(points: np.ndarray, support: np.ndarray, k: int) -> np.ndarray:
    """Reference-shaped wrapper: ``points [B,3,N]``, ``support [B,3,Q]`` -> ``[B,Q,k]`` int64
    (source/poco_utils.py:257-273; ``k==1`` keeps the trailing axis as the reference's ``unsqueeze(2)`` does)."""
    out = []
    for b in range(points.shape[0]):
        i, _ = knn(points[b].T, support[b].T, k)
        out.append(i)
    return np.stack(out, axis=0)


# --------------------------------------------------------------------------------------------------------------
# a7: local patches  (source/poco_utils.py:67-72, source/ppsurf_data_loader.py:83-123)
# --------------------------------------------------------------------------------------------------------------

def normalize_patches(pts_local_ms: np.ndarray, pts_query_ms: np.ndarray) -> np.ndarray:
    """``pts_local_ms [Q,P,3]``, ``pts_query_ms [Q,3]`` -> ``(p - q) / max_j ||p_j - q||``  in float32
    (source/ppsurf_data_loader.py:91-123)."""
    pts_local_ms = pts_local_ms.astype(np.float32, copy=False)
    pts_query_ms = pts_query_ms.astype(np.float32, copy=False)
    diff = pts_local_ms - pts_query_ms[:, None, :]
    dd = diff * diff
    dist = np.sqrt((dd[..., 0] + dd[..., 1]) + dd[..., 2])
    radius = dist.max(axis=-1)
    return diff / radius[:, None, None]


def get_pts_local_ps(pts_raw: np.ndarray, pts_query: np.ndarray, num_pts_local: int) -> np.ndarray:
    """k=P nearest raw points of every query, patch-normalised (source/poco_utils.py:67-72)."""
    ids, _ = knn(pts_raw, pts_query, num_pts_local)
    return normalize_patches(pts_raw[ids], pts_query)


# --------------------------------------------------------------------------------------------------------------
# a3: FKAConv layer  (source/base/nn.py:592-652)
# --------------------------------------------------------------------------------------------------------------

def fkaconv_layer(p: Params, name: str, x: np.ndarray, pts: np.ndarray, support: np.ndarray, ids: np.ndarray,
                  act: str = 'silu', dtype=np.float32) -> np.ndarray:
    """``x [B,Cin,Nin]``, ``pts [B,3,Nin]``, ``support [B,3,Ns]``, ``ids [B,Ns,K]`` -> ``[B,Cout,Ns]`` (eval mode)."""
    f = _ACT[act]
    x = x.astype(dtype, copy=False)
    ks = ids.shape[2]
    pts_g = batch_gather(pts.astype(dtype, copy=False), ids)  # [B,3,Ns,K]   nn.py:597
    x_g = batch_gather(x, ids)  # [B,Cin,Ns,K]                                nn.py:598
    pts_g = pts_g - support.astype(dtype, copy=False)[:, :, :, None]  # nn.py:601
    dist = np.sqrt((pts_g ** 2).sum(axis=1))  # [B,Ns,K]              nn.py:605
    pts_g = pts_g / _w(p, name + '.norm_radius', dtype)  # nn.py:616
    alpha = _w(p, name + '.alpha', dtype)
    beta = _w(p, name + '.beta', dtype)
    dw = _sigmoid(-alpha * dist + beta)  # nn.py:619
    dws = dw.sum(axis=2, keepdims=True)
    dws = dws + (dws == 0).astype(dtype) + dtype(1e-6)  # nn.py:621
    dw = (dw / dws * dtype(ks))[:, None, :, :]  # [B,1,Ns,K]            nn.py:622-624

    w1 = _w(p, name + '.fc1.weight', dtype).reshape(16, 3)
    w2 = _w(p, name + '.fc2.weight', dtype).reshape(16, 32)
    w3 = _w(p, name + '.fc3.weight', dtype).reshape(16, 32)
    mat = np.einsum('oc,bcnk->bonk', w1, pts_g, optimize=True)
    mat = f(mat) if ks == 1 else f(_instancenorm(p, name + '.bn1', mat, dtype))  # nn.py:627-630
    mp1 = np.broadcast_to((mat * dw).max(axis=3, keepdims=True), mat.shape)  # nn.py:631-633
    mat = np.concatenate([mat, mp1], axis=1)
    mat = np.einsum('oc,bcnk->bonk', w2, mat, optimize=True)
    mat = f(mat) if ks == 1 else f(_instancenorm(p, name + '.bn2', mat, dtype))  # nn.py:635-638
    mp2 = np.broadcast_to((mat * dw).max(axis=3, keepdims=True), mat.shape)
    mat = np.concatenate([mat, mp2], axis=1)
    mat = f(np.einsum('oc,bcnk->bonk', w3, mat, optimize=True)) * dw  # [B,16,Ns,K]   nn.py:643

    feat = np.einsum('bcnk,bmnk->bcnm', x_g, mat, optimize=True)  # [B,Cin,Ns,16]  nn.py:647-649
    wc = _w(p, name + '.cv.weight', dtype)[:, :, 0, :]  # [Cout,Cin,16]
    return np.einsum('ocm,bcnm->bon', wc, feat, optimize=True)  # nn.py:650


# --------------------------------------------------------------------------------------------------------------
# a4: residual block + max_pool  (source/base/nn.py:438-450, 677-680)
# --------------------------------------------------------------------------------------------------------------

def max_pool(data: np.ndarray, ids: np.ndarray) -> np.ndarray:
    return batch_gather(data, ids).max(axis=3)


def residual_block(p: Params, name: str, x: np.ndarray, pts, support, ids, act='silu', dtype=np.float32):
    x = x.astype(dtype, copy=False)
    x_short = x
    y = _relu(_batchnorm_eval(p, name + '.bn0', _pointwise(p, name + '.cv0', x, dtype), dtype))
    y = fkaconv_layer(p, name + '.cv1', y, pts, support, ids, act=act, dtype=dtype)
    y = _relu(_batchnorm_eval(p, name + '.bn1', y, dtype))
    y = _batchnorm_eval(p, name + '.bn2', _pointwise(p, name + '.cv2', y, dtype), dtype)
    if (name + '.shortcut.weight') in p:
        x_short = _batchnorm_eval(p, name + '.bn_shortcut', _pointwise(p, name + '.shortcut', x_short, dtype), dtype)
    if x_short.shape[2] != y.shape[2]:
        x_short = max_pool(x_short, ids)
    return _relu(y + x_short)


# --------------------------------------------------------------------------------------------------------------
# a5: FKAConv U-Net  (source/base/nn.py:508-554, 684-697; PPSurf: SiLU, x4d_bug_fixed=True, ppsurf_model.py:49-50)
# --------------------------------------------------------------------------------------------------------------

def interpolate(x: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """1-NN (or mean of K) up-sampling; negative ids are clamped to 0 as the reference does in place (nn.py:684-697)."""
    ids = np.where(ids > -1, ids, 0)
    g = batch_gather(x, ids)
    return g.mean(axis=-1) if ids.shape[-1] > 1 else g[..., 0]


def fkaconv_network(p: Params, data: dict, prefix='encoder', act='silu', x4d_bug_fixed=True, dtype=np.float32):
    """``FKAConvNetwork.forward(data, spectral_only=True)`` with ``segmentation=True``, dropout inactive.
    ``data``: ``pts [B,3,N0]``, ``support1..4``, ``ids00,01,11,12,22,23,33,34,44`` (K=16), ``ids43,32,21,10`` (K=1)."""
    e = prefix
    pts = data['pts'].astype(dtype, copy=False)
    s1, s2, s3, s4 = (data['support%d' % i].astype(dtype, copy=False) for i in (1, 2, 3, 4))
    x = np.ones_like(pts)  # nn.py:517

    def rb(name, xin, a, b, ids):
        return residual_block(p, '{}.{}'.format(e, name), xin, a, b, data[ids], act=act, dtype=dtype)

    def cbr(cv, bn, xin):
        return _relu(_batchnorm_eval(p, '{}.{}'.format(e, bn), _pointwise(p, '{}.{}'.format(e, cv), xin, dtype), dtype))

    x0 = fkaconv_layer(p, e + '.cv0', x, pts, pts, data['ids00'], act=act, dtype=dtype)
    x0 = _relu(_batchnorm_eval(p, e + '.bn0', x0, dtype))  # nn.py:519
    x0 = rb('resnetb01', x0, pts, pts, 'ids00')
    x1 = rb('resnetb10', x0, pts, s1, 'ids01')
    x1 = rb('resnetb11', x1, s1, s1, 'ids11')
    x2 = rb('resnetb20', x1, s1, s2, 'ids12')
    x2 = rb('resnetb21', x2, s2, s2, 'ids22')
    x3 = rb('resnetb30', x2, s2, s3, 'ids23')
    x3 = rb('resnetb31', x3, s3, s3, 'ids33')
    x4 = rb('resnetb40', x3, s3, s4, 'ids34')
    x4 = rb('resnetb41', x4, s4, s4, 'ids44')

    x5 = np.broadcast_to(x4.max(axis=2, keepdims=True), x4.shape)  # nn.py:531
    x4d = cbr('cv5', 'bn5', np.concatenate([x4, x5], axis=1))
    if not x4d_bug_fixed:
        x4d = x4  # nn.py:533-534 (POCO behaviour)
    x3d = cbr('cv3d', 'bn3d', np.concatenate([interpolate(x4d, data['ids43']), x3], axis=1))
    x2d = cbr('cv2d', 'bn2d', np.concatenate([interpolate(x3d, data['ids32']), x2], axis=1))
    x1d = cbr('cv1d', 'bn1d', np.concatenate([interpolate(x2d, data['ids21']), x1], axis=1))
    xout = cbr('cv0d', 'bn0d', np.concatenate([interpolate(x1d, data['ids10']), x0], axis=1))
    return _pointwise(p, e + '.fcout', xout, dtype)  # [B,latent,N0]


# --------------------------------------------------------------------------------------------------------------
# a2: support sampling + index tensors  (source/poco_data_loader.py:59-134, 137-209)
# --------------------------------------------------------------------------------------------------------------

def _rotation(axis: int, deg: float) -> np.ndarray:
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    m = np.eye(3)
    a, b = [(1, 2), (0, 2), (0, 1)][axis]
    m[a, a], m[a, b], m[b, a], m[b, b] = c, s, -s, c
    return m


def sampling_quantized(pts: np.ndarray, n_support: int, rng: np.random.Generator) -> np.ndarray:
    """Indices of ``n_support`` points of ``pts [N,3]``: one representative per voxel of a randomly rotated grid,
    voxel size ``||bbox||_2 / sqrt(n_support)``, halved until enough points were picked, random trim of the
    last round (source/poco_data_loader.py:85-127).  Non-deterministic by construction in the reference
    (random rotations, randperm, which voxel member wins) -> only distributional parity is defined."""
    n = pts.shape[0]
    if n_support == n:
        return np.arange(n)
    vox = float(np.linalg.norm(pts.max(axis=0) - pts.min(axis=0))) / math.sqrt(n_support)
    cur = pts.astype(np.float64)
    ids = np.arange(n)
    picked, count = [], 0
    while True:
        rot = _rotation(2, rng.uniform(-180, 180)) @ _rotation(1, rng.uniform(-180, 180)) @ \
            _rotation(0, rng.uniform(-180, 180))
        pr = cur @ rot.T
        cell = np.floor((pr - pr.min(axis=0)) / vox).astype(np.int64)
        dims = cell.max(axis=0) + 1
        key = (cell[:, 2] * dims[1] + cell[:, 1]) * dims[0] + cell[:, 0]
        _, first = np.unique(key, return_index=True)
        if count + first.shape[0] < n_support:
            picked.append(ids[first])
            count += first.shape[0]
            keep = np.ones(cur.shape[0], dtype=bool)
            keep[first] = False
            cur, ids = cur[keep], ids[keep]
            vox = vox / 2
        else:
            sel = rng.permutation(first.shape[0])[:n_support - count]
            picked.append(ids[first[sel]])
            break
    return np.concatenate(picked)


def get_fkaconv_ids(pts: np.ndarray, rng: np.random.Generator) -> dict:
    """``pts [B,3,N]`` -> supports (ratio 1/4 four times) and the 13 index tensors
    (source/poco_data_loader.py:137-209)."""
    b = pts.shape[0]
    sup = [pts]
    for _ in range(4):
        prev = sup[-1]
        n_sup = max(1, int(prev.shape[2] * 0.25))
        sel = [sampling_quantized(prev[i].T, n_sup, rng) for i in range(b)]
        sup.append(np.stack([prev[i][:, sel[i]] for i in range(b)], axis=0))
    out = {'support%d' % i: sup[i] for i in (1, 2, 3, 4)}
    for a, c in ((0, 0), (0, 1), (1, 1), (1, 2), (2, 2), (2, 3), (3, 3), (3, 4), (4, 4)):
        out['ids%d%d' % (a, c)] = knn_batched(sup[a], sup[c], 16)
    for a, c in ((4, 3), (3, 2), (2, 1), (1, 0)):
        out['ids%d%d' % (a, c)] = knn_batched(sup[a], sup[c], 1)
    return out


# --------------------------------------------------------------------------------------------------------------
# a8: global branch  (source/poco_model.py:381-419)
# --------------------------------------------------------------------------------------------------------------

def interp_attention(p: Params, latents: np.ndarray, pts: np.ndarray, pts_query: np.ndarray, proj_ids: np.ndarray,
                     prefix='projection', dtype=np.float32) -> np.ndarray:
    """``latents [B,C,N]``, ``pts [B,3,N]``, ``pts_query [B,Q,3]``, ``proj_ids [B,Q,k]`` -> ``[B,C,Q]``."""
    q = np.swapaxes(pts_query.astype(dtype, copy=False), 1, 2)  # [B,3,Q]
    x = batch_gather(latents.astype(dtype, copy=False), proj_ids)  # [B,C,Q,k]      poco_model.py:400
    rel = q[:, :, :, None] - batch_gather(pts.astype(dtype, copy=False), proj_ids)  # poco_model.py:401-402
    x = np.concatenate([x, rel], axis=1)
    x = _relu(_pointwise(p, prefix + '.fc1', x, dtype))
    x = _relu(_pointwise(p, prefix + '.fc2', x, dtype))
    x = _relu(_pointwise(p, prefix + '.fc3', x, dtype))
    query = _pointwise(p, prefix + '.fc_query', x, dtype)  # [B,64,Q,k]
    value = _pointwise(p, prefix + '.fc_value', x, dtype)  # [B,C,Q,k]
    query = query - query.max(axis=-1, keepdims=True)
    e = np.exp(query)
    att = (e / e.sum(axis=-1, keepdims=True)).mean(axis=1)  # [B,Q,k]                poco_model.py:412
    out = np.einsum('bqk,bcqk->bcq', att, value, optimize=True)  # poco_model.py:413-414
    return _pointwise(p, prefix + '.fc8', out, dtype)  # poco_model.py:417


# --------------------------------------------------------------------------------------------------------------
# a9: local branch  (source/base/nn.py:305-373 with use_point_stn=False, use_feat_stn=True, sym_op='att';
#                    STN nn.py:162-190; AttentionPoco nn.py:84-96)
# --------------------------------------------------------------------------------------------------------------

def pointnet_feat(p: Params, x: np.ndarray, prefix='point_net', dtype=np.float32) -> np.ndarray:
    """``x [Q,3,P]`` (patch-space points) -> ``[Q,latent]``; BatchNorm in eval mode."""
    n = prefix
    x = x.astype(dtype, copy=False)

    def cbr(cv, bn, xin, relu=True):
        y = _batchnorm_eval(p, '{}.{}'.format(n, bn), _pointwise(p, '{}.{}'.format(n, cv), xin, dtype), dtype)
        return _relu(y) if relu else y

    x = cbr('conv0a', 'bn0a', x)
    x = cbr('conv0b', 'bn0b', x)  # [Q,64,P]
    t = cbr('stn2.conv1', 'stn2.bn1', x)
    t = cbr('stn2.conv2', 'stn2.bn2', t)
    t = cbr('stn2.conv3', 'stn2.bn3', t)
    t = t.max(axis=2)  # MaxPool1d(P)                                               nn.py:170
    t = cbr('stn2.fc1', 'stn2.bn4', t)
    t = cbr('stn2.fc2', 'stn2.bn5', t)
    t = _pointwise(p, n + '.stn2.fc3', t, dtype)
    t = (t + np.eye(64, dtype=dtype).reshape(1, -1)).reshape(-1, 64, 64)  # nn.py:187-189
    x = np.einsum('qij,qjp->qip', t, x, optimize=True)  # nn.py:329
    x = cbr('conv1', 'bn1', x)
    x = cbr('conv2', 'bn2', x)
    x = cbr('conv3', 'bn3', x, relu=False)  # [Q,C,P]                               nn.py:336
    wq = _w(p, n + '.att.fc_query.weight', dtype).reshape(1, -1)
    bq = _w(p, n + '.att.fc_query.bias', dtype)
    query = np.einsum('oc,qcp->qop', wq, x, optimize=True)[:, 0, :] + bq  # [Q,P]    nn.py:88
    wv = _w(p, n + '.att.fc_value.weight', dtype).reshape(x.shape[1], x.shape[1])
    bv = _w(p, n + '.att.fc_value.bias', dtype)
    value = np.einsum('oc,qcp->qpo', wv, x, optimize=True) + bv  # [Q,P,C]           nn.py:89
    query = query - query.max(axis=-1, keepdims=True)
    e = np.exp(query)
    w = e / e.sum(axis=-1, keepdims=True)  # nn.py:91
    return (value * w[:, :, None]).sum(axis=1)  # nn.py:93


# --------------------------------------------------------------------------------------------------------------
# a10/a11: sum of branches, MLP, occupancy  (source/ppsurf_model.py:82-117, source/base/nn.py:376-417,
#                                            source/poco_utils.py:74-82)
# --------------------------------------------------------------------------------------------------------------

def mlp(p: Params, x: np.ndarray, prefix='mlp', dtype=np.float32) -> np.ndarray:
    """``[Q,C]`` -> ``[Q,out]``: (Linear, BN, ReLU, Dropout(eval: identity)) x2, Linear."""
    for i in (0, 1):
        x = _pointwise(p, '{}.layers.{}.0'.format(prefix, i), x, dtype)
        x = _relu(_batchnorm_eval(p, '{}.layers.{}.1'.format(prefix, i), x, dtype))
    return _pointwise(p, '{}.layers.2.0'.format(prefix), x, dtype)


def from_latent(p: Params, data: dict, k: int = 64, dtype=np.float32) -> np.ndarray:
    """``PPSurfNetwork.from_latent``: ``data`` has ``pts [B,3,N]``, ``latents [B,C,N]``, ``pts_query [B,Q,3]``,
    ``pts_local_ps [B,Q,P,3]`` and optionally ``proj_ids``; returns logits ``[B,2,Q]``."""
    if 'proj_ids' not in data:
        data['proj_ids'] = knn_batched(data['pts'], np.swapaxes(data['pts_query'], 1, 2), k)
    feat_proj = interp_attention(p, data['latents'], data['pts'], data['pts_query'], data['proj_ids'], dtype=dtype)
    loc = data['pts_local_ps']
    b, q, npl, _ = loc.shape
    feat_pn = pointnet_feat(p, np.swapaxes(loc.reshape(b * q, npl, 3), 1, 2), dtype=dtype).reshape(b, q, -1)
    feat = np.swapaxes(feat_proj, 1, 2) + feat_pn  # ppsurf_model.py:100
    out = mlp(p, feat.reshape(b * q, -1), dtype=dtype).reshape(b, q, -1)
    return np.swapaxes(out, 1, 2)


def occupancy_from_logits(logits: np.ndarray) -> np.ndarray:
    """``[B,2,Q]`` -> ``softmax(dim=1)[:,0] - softmax(dim=1)[:,1]`` (source/poco_utils.py:79-80)."""
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m)
    s = e / e.sum(axis=1, keepdims=True)
    return s[:, 0] - s[:, 1]


def network_forward(p: Params, data: dict, k: int = 64, dtype=np.float32) -> np.ndarray:
    """``PPSurfNetwork.forward`` (train/test path): encoder with supplied ids, then ``from_latent`` with the
    supplied ``proj_ids`` (source/ppsurf_model.py:70-74)."""
    data['latents'] = fkaconv_network(p, data, dtype=dtype)
    return from_latent(p, data, k=k, dtype=dtype)


# --------------------------------------------------------------------------------------------------------------
# a11: query grid and region growing  (source/poco_utils.py:52-61, 178-254)
# --------------------------------------------------------------------------------------------------------------

def grid_definition(input_points: np.ndarray, resolution: int, padding: int = 1):
    """``bmin``/``bmax`` are the scalar min/max over ALL coordinates; returns ``(step, bmin_pad, pts_ids)`` in the
    reference's dtypes (float32 scalars under numpy>=2 promotion, int32 voxel ids)  (source/poco_utils.py:52-61)."""
    bmin = input_points.min()
    bmax = input_points.max()
    step = (bmax - bmin) / (resolution - 1)
    bmin_pad = bmin - padding * step
    pts_ids = ((input_points - bmin) / step + padding).astype(np.int32)
    return step, bmin_pad, pts_ids


def dense_grid_queries(input_points: np.ndarray, resolution: int, padding: int = 1) -> np.ndarray:
    """All ``(resolution+2*padding)^3`` vertices in C order, coordinates ``idx*step + bmin_pad`` in float32
    (the benchmark workload of SURVEY.md §8d; expression from source/poco_utils.py:212-213)."""
    step, bmin_pad, _ = grid_definition(input_points, resolution, padding)
    r = resolution + 2 * padding
    coord = np.stack(np.meshgrid(np.arange(r), np.arange(r), np.arange(r), indexing='ij'), axis=-1)
    coord = coord.reshape(-1, 3).astype(np.float32)
    return (coord * np.float32(step) + np.float32(bmin_pad)).astype(np.float32)


def create_volume(predict: typing.Callable[[np.ndarray], np.ndarray], input_points: np.ndarray, resolution: int,
                  padding: int = 1, dilation_size: int = 2, out_value: float = 1.0, batch: int = 50000) -> np.ndarray:
    """Region-growing evaluation of the occupancy field (source/poco_utils.py:178-254).
    ``predict(queries [q,3] float32) -> occupancy [q]``."""
    step, bmin_pad, pts_ids = grid_definition(input_points, resolution, padding)
    r = resolution + 2 * padding
    shape = (r, r, r)

    def dilate(ids):
        m = np.zeros(shape, dtype=bool)
        lo = np.maximum(0, ids - dilation_size)
        hi = np.minimum(r, ids + dilation_size + 1)
        for a, c in zip(lo, hi):
            m[a[0]:c[0], a[1]:c[1], a[2]:c[2]] = True
        return m

    volume = np.full(shape, np.nan, dtype=np.float64)
    to_see = np.ones(shape, dtype=bool)
    pts_ids = pts_ids.astype(np.int64)
    while pts_ids.shape[0] > 0:
        mask = dilate(pts_ids)
        coord = np.argwhere(mask).astype(np.float32)
        queries = (coord * np.float32(step) + np.float32(bmin_pad)).astype(np.float32)
        z = np.concatenate([predict(queries[s:s + batch]) for s in range(0, queries.shape[0], batch)], axis=0)
        volume[mask] = z.astype(np.float64)
        to_see[pts_ids[:, 0], pts_ids[:, 1], pts_ids[:, 2]] = False
        v = volume[pts_ids[:, 0], pts_ids[:, 1], pts_ids[:, 2]]
        mask_neg = dilate(pts_ids[v <= 0])
        mask_pos = dilate(pts_ids[v >= 0])
        with np.errstate(invalid='ignore'):
            new_mask = (mask_neg & (volume >= 0) & to_see) | (mask_pos & (volume <= 0) & to_see)
        pts_ids = np.argwhere(new_mask).astype(np.int64)
    for ax in range(3):
        sl = [slice(None)] * 3
        sl[ax] = slice(0, padding)
        volume[tuple(sl)] = out_value
        sl[ax] = slice(-padding, None)
        volume[tuple(sl)] = out_value
    return volume


# --------------------------------------------------------------------------------------------------------------
# a1: latent averaging loop  (source/poco_model.py:200-237)
# --------------------------------------------------------------------------------------------------------------

def latent_loop_schedule(n_pts: int, subsample: int, iters: int, rng: np.random.Generator) -> typing.List[np.ndarray]:
    """The index sets of the reference's latent loop (they do not depend on network output): every point gets
    >= ``iters`` encodings on random ``subsample``-point subsets; padding ids come from a fresh permutation of all
    points; ``counts[ids] += 1`` with duplicate ids is last-writer-wins, i.e. +1 per distinct id."""
    counts = np.zeros(n_pts, dtype=np.int64)
    sched = []
    for cur in range(iters):
        while counts.min() < cur + 1:
            valid = np.nonzero(counts == cur)[0]
            if n_pts >= subsample:
                ids = valid[rng.permutation(valid.shape[0])[:subsample]]
                if ids.shape[0] < subsample:
                    ids = np.concatenate([ids, rng.permutation(n_pts)[:subsample - ids.shape[0]]])
            else:
                ids = np.arange(n_pts)
            counts[np.unique(ids)] += 1
            sched.append(ids)
    return sched


def accumulate_latents(n_pts: int, passes: typing.Iterable[typing.Tuple[np.ndarray, np.ndarray]], latent_size: int):
    """``latent[ids] += partial; counts[ids] += 1; latent /= counts`` with torch advanced-indexing semantics
    (last writer wins for duplicate ids inside one pass)  (source/poco_model.py:228-234).
    ``passes`` yields ``(ids [n], partial [n,C])``."""
    latent = np.zeros((n_pts, latent_size), dtype=np.float32)
    counts = np.zeros((n_pts,), dtype=np.float32)
    for ids, partial in passes:
        latent[ids] = latent[ids] + partial
        counts[ids] = counts[ids] + 1
    return latent / counts[:, None], counts


# --------------------------------------------------------------------------------------------------------------
# synthetic workload of SURVEY.md §8d
# --------------------------------------------------------------------------------------------------------------

def state_dict_digest(p: Params) -> str:
    """sha256 over names, shapes and raw bytes: proves both sides of a parity test hold the same weights."""
    import hashlib
    h = hashlib.sha256()
    for name in p:
        a = np.ascontiguousarray(p[name])
        h.update(name.encode())
        h.update(str(a.shape).encode())
        h.update(str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def synthetic_cloud(n: int, seed: int = 42, radius: float = 0.4, noise: float = 0.005) -> np.ndarray:
    """Noisy sphere, ``[n,3]`` float32, inside [-0.5,0.5]^3 like reference-normalised data."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return (radius * d + noise * rng.standard_normal((n, 3))).astype(np.float32)


# --------------------------------------------------------------------------------------------------------------
# f3: marching cubes (this repo's generated case table, vertices on grid edges like skimage's) and the
#     bisection refinement of the vertices  (source/poco_utils.py:96, 111-168)
# --------------------------------------------------------------------------------------------------------------

def marching_cubes(volume: np.ndarray, level: float = 0.0):
    """numpy restatement of csrc/mcubes.cu: ``volume [r,r,r]`` -> ``verts [nv,3]`` float32 in volume-index coordinates (one per crossed
    grid edge, ordered by edge number 3 * linear index + axis), ``vert_edge [nv]``, ``faces [nt,3]`` (cells in C order).  The reference
    uses skimage's Lewiner tables (third party, not installed); both place a vertex by linear interpolation on every crossed grid
    edge, so the vertex SET is the same, the triangulation of ambiguous cells may differ.  The case table is the product's
    (ppsurf_b200/mc_tables.py, generated); tests pin it through table-independent properties: closed 2-manifold, orientation, Euler
    characteristic, vertices exactly on the level set of the trilinear edge interpolant."""
    from ppsurf_b200 import mc_tables as T
    vol = np.asarray(volume, dtype=np.float32)
    r = vol.shape[0]
    c = np.zeros((r - 1, r - 1, r - 1), dtype=np.int32)
    nan = np.zeros_like(c, dtype=bool)
    for i in range(8):
        ox, oy, oz = i & 1, (i >> 1) & 1, (i >> 2) & 1
        v = vol[ox:ox + r - 1, oy:oy + r - 1, oz:oz + r - 1]
        nan |= np.isnan(v)
        c |= (v < np.float32(level)).astype(np.int32) << i
    c[nan] = 0
    cells = np.argwhere(T.TRI_COUNT[c] > 0)
    tri_edges = []
    for x, y, z in cells:
        row = T.TRI_TABLE[c[x, y, z]]
        for e in row[row >= 0]:
            o = T.EDGE_ORIGIN[e]
            g = ((x + o[0]) * r + (y + o[1])) * r + (z + o[2])
            tri_edges.append(3 * g + T.EDGE_AXIS[e])
    tri_edges = np.asarray(tri_edges, dtype=np.int64)
    vert_edge = np.unique(tri_edges)
    faces = np.searchsorted(vert_edge, tri_edges).reshape(-1, 3).astype(np.int32)
    g, axis = vert_edge // 3, vert_edge % 3
    xyz = np.stack([g // (r * r), (g // r) % r, g % r], axis=1)
    stride = np.where(axis == 0, r * r, np.where(axis == 1, r, 1))
    flat = vol.reshape(-1)
    va, vb = flat[g], flat[g + stride]
    t = np.clip((np.float32(level) - va) / (vb - va), np.float32(0), np.float32(1)).astype(np.float32)
    verts = xyz.astype(np.float32)
    verts[np.arange(verts.shape[0]), axis] += t
    return verts, vert_edge.astype(np.int32), faces


def refine_vertices(predict: typing.Callable[[np.ndarray], np.ndarray], volume: np.ndarray, verts: np.ndarray, step, bmin_pad,
                    refine_iter: int = 10) -> np.ndarray:
    """source/poco_utils.py:111-168: vertices strictly inside a grid edge are bisected ``refine_iter`` times between the edge's two
    grid vertices; returns all vertices in model space.  ``verts`` in volume-index coordinates, ``predict(q [n,3] f32) -> [n]``."""
    step, bmin_pad = np.float32(step), np.float32(bmin_pad)
    verts = np.asarray(verts, dtype=np.float32)
    dirs = verts - np.floor(verts)
    dirs = (dirs > 0).astype(verts.dtype)
    mask = np.logical_and(dirs.sum(axis=1) > 0, dirs.sum(axis=1) < 2)
    v = verts[mask]
    dirs = dirs[mask]
    v1 = np.floor(v)
    v2 = v1 + dirs
    v1 = v1.astype(int)
    v2 = v2.astype(int)
    preds1 = np.asarray(volume)[v1[:, 0], v1[:, 1], v1[:, 2]].astype(np.float32)
    preds2 = np.asarray(volume)[v2[:, 0], v2[:, 1], v2[:, 2]].astype(np.float32)
    v1 = v1.astype(np.float32) * step + bmin_pad
    v2 = v2.astype(np.float32) * step + bmin_pad
    mask_tmp = np.logical_and(~np.isnan(preds1), ~np.isnan(preds2))
    v, v1, v2, preds1, preds2 = v[mask_tmp], v1[mask_tmp], v2[mask_tmp], preds1[mask_tmp], preds2[mask_tmp]
    mask[mask] = mask_tmp
    out = verts * step + bmin_pad
    v = v * step + bmin_pad
    for _ in range(refine_iter):
        preds = np.asarray(predict(v.astype(np.float32)), dtype=np.float32)
        mask1 = (preds * preds1) > 0
        v1[mask1] = v[mask1]
        preds1[mask1] = preds[mask1]
        mask2 = (preds * preds2) > 0
        v2[mask2] = v[mask2]
        preds2[mask2] = preds[mask2]
        v = (v2 + v1) / 2
        out[mask] = v
    return out
