"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors made by the reference.
Run on the B200 box:  python -m pytest tests -m gpu"""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-4  # north-star tolerance: logits within 1e-4 abs of the reference's fp32 CPU path


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'the gpu tests need a CUDA device'
    from ppsurf_b200 import ops
    ops.require_device()
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def net(dev, weights):
    import ppsurf_b200
    n = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
    n.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    return n.to(dev).eval()


def cu(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)


def assert_knn_equal(oracle, pts, qry, idx, d2, ref_idx):
    """identical distance multiset per query (bit exact), identical index set except inside exact ties"""
    idx = idx.astype(np.int64)
    got = oracle.sq_dist_f32(qry[:, None, :], pts[idx])
    if d2 is not None:
        np.testing.assert_array_equal(got, d2)  # the reported dist2 is the fp32 distance of the reported neighbour
        assert np.all(np.diff(d2, axis=1) >= 0)
    ref = oracle.sq_dist_f32(qry[:, None, :], pts[ref_idx])
    np.testing.assert_array_equal(np.sort(got, axis=1), np.sort(ref, axis=1))
    for r in np.nonzero(np.any(np.sort(idx, axis=1) != np.sort(ref_idx, axis=1), axis=1))[0]:
        diff = list(set(idx[r]) ^ set(ref_idx[r]))
        assert np.all(oracle.sq_dist_f32(qry[r][None], pts[diff]) == np.sort(ref[r])[-1]), 'row {}'.format(r)


# ---- a6 kNN ---------------------------------------------------------------------------------------------------------

def test_knn_golden(dev, oracle):
    from ppsurf_b200 import ops
    g = load_golden('knn')
    index = ops.KnnIndex(cu(g['pts'], dev))
    idx, d2 = index.query(cu(g['qry'], dev), 64, return_dist=True)
    assert_knn_equal(oracle, g['pts'], g['qry'], idx.cpu().numpy(), d2.cpu().numpy(), g['idx64'].astype(np.int64))
    idx1 = index.query(cu(g['qry'], dev), 1)
    assert_knn_equal(oracle, g['pts'], g['qry'], idx1.cpu().numpy(), None, g['idx1'].astype(np.int64))
    # k > N clamps to N like the reference (source/poco_utils.py:259-260)
    small = ops.knn(cu(g['pts'][:10], dev), cu(g['qry'][:5], dev), 16)
    assert tuple(small.shape) == (5, 10)
    np.testing.assert_array_equal(small.cpu().numpy(), g['idx_small'])


@pytest.mark.parametrize('n,q,k', [(20000, 3000, 64), (3000, 500, 200), (2500, 2500, 16), (39, 39, 16), (700, 100, 300), (5000, 800, 100),
                                   (150, 64, 100)])
def test_knn_vs_oracle(dev, oracle, n, q, k):
    from ppsurf_b200 import ops
    rng = np.random.default_rng(n + k)
    pts = oracle.synthetic_cloud(n, seed=n)
    qry = np.concatenate([pts[rng.integers(0, n, q // 2)] + 0.01 * rng.standard_normal((q // 2, 3)),
                          rng.uniform(-0.6, 0.6, (q - q // 2, 3))]).astype(np.float32)
    idx, d2 = ops.knn(cu(pts, dev), cu(qry, dev), k, return_dist=True)
    ref_idx, _ = oracle.knn(pts, qry, k)
    assert_knn_equal(oracle, pts, qry, idx.cpu().numpy(), d2.cpu().numpy(), ref_idx)


@pytest.mark.parametrize('k', [64, 100, 200])
def test_knn_grid_ordered_queries(dev, oracle, k):
    """consecutive grid vertices (the decoder's query order): every query is seeded with its predecessor's neighbour list and,
    near the surface, goes straight to the cells around its search ball -- the result must still be the exact kNN"""
    from ppsurf_b200 import ops
    n = 30000
    pts = oracle.synthetic_cloud(n, seed=77)
    ax = np.linspace(-0.5, 0.5, 24, dtype=np.float32)
    qry = np.stack(np.meshgrid(ax, ax, ax, indexing='ij'), axis=-1).reshape(-1, 3)  # C order: z fastest, like the volume
    idx, d2 = ops.knn(cu(pts, dev), cu(qry, dev), k, return_dist=True)
    ref_idx, _ = oracle.knn(pts, qry, k)
    assert_knn_equal(oracle, pts, qry, idx.cpu().numpy(), d2.cpu().numpy(), ref_idx)


def test_knn_tuning_knobs_do_not_change_the_result(dev, oracle):
    """run length, finest cells per point and the scan-whole-node threshold only steer the traversal: indices and distances are bit
    identical under every setting (unique total order (dist2, index)), and equal to the oracle's"""
    from ppsurf_b200 import _lib, ops
    pts = oracle.synthetic_cloud(20000, seed=5)
    ax = np.linspace(-0.5, 0.5, 20, dtype=np.float32)
    qry = np.stack(np.meshgrid(ax, ax, ax, indexing='ij'), axis=-1).reshape(-1, 3)
    lib = _lib.lib
    old = (lib.pps_debug_knn_run(-1), lib.pps_debug_knn_cells(-1), lib.pps_debug_knn_scan_child(-1))
    try:
        ref = None
        for run, cells, sc in ((16, 2, 192), (1, 2, 0), (8, 1, 32), (32, 8, 1024), (16, 4, 64)):
            lib.pps_debug_knn_run(run)
            lib.pps_debug_knn_cells(cells)
            lib.pps_debug_knn_scan_child(sc)
            idx, d2 = ops.knn(cu(pts, dev), cu(qry, dev), 64, return_dist=True)
            got = (idx.cpu().numpy(), d2.cpu().numpy())
            if ref is None:
                ref = got
                assert_knn_equal(oracle, pts, qry, got[0], got[1], oracle.knn(pts, qry, 64)[0])
            np.testing.assert_array_equal(got[0], ref[0])
            np.testing.assert_array_equal(got[1], ref[1])
    finally:
        lib.pps_debug_knn_run(old[0])
        lib.pps_debug_knn_cells(old[1])
        lib.pps_debug_knn_scan_child(old[2])


def test_knn_edge_cases(dev, oracle):
    from ppsurf_b200 import ops
    rng = np.random.default_rng(5)
    # duplicated points and exact ties: the (dist2, index) order makes the result unique -> equal to the oracle's
    base = rng.uniform(-0.5, 0.5, (300, 3)).astype(np.float32)
    pts = np.concatenate([base, base, base[:50]])
    qry = base[:64].copy()
    idx, d2 = ops.knn(cu(pts, dev), cu(qry, dev), 8, return_dist=True)
    ref_idx, ref_d2 = oracle.knn(pts, qry, 8)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    np.testing.assert_array_equal(d2.cpu().numpy(), ref_d2)
    # all points identical (degenerate bounding box), queries far outside the box
    same = np.tile(np.array([[0.1, -0.2, 0.3]], dtype=np.float32), (100, 1))
    far = np.array([[10, 10, 10], [-7, 0, 3], [0.1, -0.2, 0.3]], dtype=np.float32)
    idx = ops.knn(cu(same, dev), cu(far, dev), 5).cpu().numpy()
    np.testing.assert_array_equal(idx, np.tile(np.arange(5), (3, 1)))
    # empty query set
    assert ops.knn(cu(base, dev), torch.empty((0, 3), device=dev), 4).shape == (0, 4)
    # planar cloud (zero extent on one axis)
    plane = base.copy()
    plane[:, 2] = 0.25
    q2 = rng.uniform(-0.5, 0.5, (200, 3)).astype(np.float32)
    idx, d2 = ops.knn(cu(plane, dev), cu(q2, dev), 16, return_dist=True)
    assert_knn_equal(oracle, plane, q2, idx.cpu().numpy(), d2.cpu().numpy(), oracle.knn(plane, q2, 16)[0])


def test_knn_bad_arguments(dev):
    from ppsurf_b200 import _lib, ops
    pts = torch.zeros((10, 3), device=dev)
    index = ops.KnnIndex(pts)
    with pytest.raises(_lib.PpsError):
        _lib.check(_lib.lib.pps_knn_query(index.buf.data_ptr(), 10, pts.data_ptr(), 10, 11, pts.data_ptr(), None, None))
    with pytest.raises(_lib.PpsError):
        _lib.check(_lib.lib.pps_knn_build(pts.data_ptr(), 10, index.buf.data_ptr(), 16, None))


# ---- a7 patches -----------------------------------------------------------------------------------------------------

def test_patches_golden(dev, oracle):
    from ppsurf_b200 import ops
    g = load_golden('patches')
    pts, qry = cu(g['pts'], dev), cu(g['qry'], dev)
    idx, d2 = ops.knn(pts, qry, 64, return_dist=True)
    loc = ops.patch_normalize(pts, qry, idx, d2, 50).cpu().numpy()
    # same neighbours, same fp32 expression: bit exact up to the order inside exact distance ties
    np.testing.assert_array_equal(np.sort(loc, axis=1), np.sort(g['pts_local_ps'], axis=1))
    same_order = np.all(loc == g['pts_local_ps'], axis=(1, 2))
    assert same_order.mean() > 0.99


# ---- generic linear ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('m,n,k', [(1000, 256, 256), (333, 64, 64), (129, 2, 256), (77, 4096, 64), (500, 128, 3), (64, 70, 37),
                                   (39, 512, 8192), (156, 256, 4096), (625, 40, 2048)])  # the last three take the split-K route
def test_linear(dev, m, n, k):
    from ppsurf_b200 import ops
    gen = torch.Generator().manual_seed(m + n + k)
    x = torch.randn((m, k), generator=gen)
    w = torch.randn((n, k), generator=gen) / k ** 0.5
    b = torch.randn((n,), generator=gen)
    r = torch.randn((m, n), generator=gen)
    ref = torch.relu(x.double() @ w.double().T + b.double() + r.double())
    out = ops.linear(x.to(dev), w.to(dev), b.to(dev), residual=r.to(dev), relu=True).cpu().double()
    assert (out - ref).abs().max() < 2e-5
    src = torch.randn((50, k), generator=gen)
    gi = torch.randint(0, 50, (m,), generator=gen)
    ref = src.double()[gi] @ w.double().T
    out = ops.linear(src.to(dev), w.to(dev), gather=gi.to(dev, torch.int32)).cpu().double()
    assert (out - ref).abs().max() < 2e-5


# ---- a3/a4/a5 encoder ---------------------------------------------------------------------------------------------------

def test_fkaconv_and_resblock_golden(dev, net, weights_digest):
    from ppsurf_b200 import ops
    g = load_golden('fkaconv')
    assert str(g['digest']) == weights_digest
    enc = net.packed()['encoder']
    pts, sup = cu(g['pts'][None], dev), cu(g['support'][None], dev)
    ids = cu(g['ids'][None], dev, torch.int32)
    # the packed layer includes bn1 + ReLU of the enclosing block: compare against the same composition of the golden
    sd = net.state_dict()
    s = (sd['encoder.resnetb01.bn1.weight'] / torch.sqrt(sd['encoder.resnetb01.bn1.running_var'] + 1e-5)).cpu().numpy()
    sh = (sd['encoder.resnetb01.bn1.bias'].cpu().numpy() - sd['encoder.resnetb01.bn1.running_mean'].cpu().numpy() * s)
    ref = np.maximum(g['y_fka'][0].T * s + sh, 0)
    x32 = cu(g['x32'][0].T[None], dev)
    y = ops.fkaconv(enc['resnetb01']['cv1'], x32, pts, sup, ids)[0].cpu().numpy()
    assert np.abs(y - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    x64 = cu(g['x64'][0].T[None], dev)
    y = net._resblock(enc['resnetb10'], x64, pts, sup, ids)[0].cpu().numpy()
    assert np.abs(y - g['y_rb'][0].T).max() < 2e-5 * max(1.0, np.abs(g['y_rb']).max())
    ids_same = cu(g['ids_same'][None], dev, torch.int32)
    y = net._resblock(enc['resnetb01'], x64, pts, pts, ids_same)[0].cpu().numpy()
    assert np.abs(y - g['y_rb_same'][0].T).max() < 2e-5 * max(1.0, np.abs(g['y_rb_same']).max())


def test_encoder_golden(dev, net, oracle, weights):
    g = load_golden('encoder')
    data = {k: cu(v, dev, torch.int64 if k.startswith('ids') else None) for k, v in g.items() if k not in ('latents', 'digest')}
    lat = net.encode(data)[0].cpu().numpy().T  # [C,N]
    scale = np.abs(g['latents']).max()
    assert np.abs(lat - g['latents'][0]).max() < 5e-5 * scale
    # batch of two different clouds == two single runs (per-sample InstanceNorm statistics)
    d2 = {k: torch.cat([v, v.flip(-1) if k in ('pts',) else v], dim=0) for k, v in data.items()}
    d2['pts'] = torch.cat([data['pts'], data['pts'] * 0.9], dim=0)
    for i in (1, 2, 3, 4):
        d2['support%d' % i] = torch.cat([data['support%d' % i], data['support%d' % i] * 0.9], dim=0)
    lat2 = net.encode(d2)
    assert np.abs(lat2[0].cpu().numpy().T - g['latents'][0]).max() < 5e-5 * scale
    ref1 = oracle.fkaconv_network(weights, {k: v[1:2].cpu().numpy() for k, v in d2.items()})
    assert np.abs(lat2[1].cpu().numpy().T - ref1[0]).max() < 5e-5 * max(scale, np.abs(ref1).max())


def test_get_latent_indices_and_latents(dev, net, oracle, weights):
    """spatial_ids: supports are subsets of the right sizes, the 13 index tensors equal the oracle's kNN on those
    supports, and get_latent's output equals the oracle encoder on the same supports/ids"""
    pts = oracle.synthetic_cloud(1500, seed=21)
    net.sampling_seed = 7
    data = net.spatial_ids(cu(pts.T[None], dev))
    sizes = [1500, 375, 93, 23, 5]
    lv = [pts] + [data['support%d' % i][0].T.cpu().numpy() for i in (1, 2, 3, 4)]
    for i in range(1, 5):
        assert lv[i].shape == (sizes[i], 3)
        prev = {tuple(r) for r in lv[i - 1].tolist()}
        assert all(tuple(r) in prev for r in lv[i].tolist())
    for a, c, k in ((0, 0, 16), (0, 1, 16), (1, 1, 16), (1, 2, 16), (2, 2, 16), (2, 3, 16), (3, 3, 16), (3, 4, 16), (4, 4, 16),
                    (4, 3, 1), (3, 2, 1), (2, 1, 1), (1, 0, 1)):
        got = data['ids%d%d' % (a, c)][0].cpu().numpy()
        assert got.dtype == np.int64 and got.shape == (sizes[c], min(k, sizes[a]))
        assert_knn_equal(oracle, lv[a], lv[c], got, None, oracle.knn(lv[a], lv[c], k)[0])
    # with only 5 points in support4 the ids have 5 columns and the reference's 16-wide kernel cannot run either: use a
    # cloud large enough (support4 = 16 points) for the comparison of values
    pts = oracle.synthetic_cloud(4200, seed=22)
    data = net.get_latent({'pts': cu(pts.T[None], dev)})
    assert data['proj_correction'] is None and tuple(data['latents'].shape) == (1, 256, 4200)
    ref = oracle.fkaconv_network(weights, {k: v.cpu().numpy() for k, v in data.items() if k.startswith(('pts', 'support', 'ids'))})
    got = data['latents'].cpu().numpy()
    assert np.abs(got - ref).max() < 5e-5 * max(1.0, np.abs(ref).max())


def test_device_support_sampling(dev, oracle):
    """quantised support sampling on the device: right count, unique ids, reproducible, blue-noise-like spread comparable
    to the oracle's restatement of the reference algorithm (the reference is random by construction: distributional parity)"""
    from ppsurf_b200 import ops
    from ppsurf_b200.sampling import ROUNDS, random_rotations
    pts = oracle.synthetic_cloud(4000, 5)
    rot = torch.from_numpy(random_rotations(np.random.default_rng(1), ROUNDS)).to(dev)
    sel = ops.sample_quantized(cu(pts, dev), 1000, rot, seed=7).cpu().numpy()
    assert sel.shape == (1000,) and np.unique(sel).shape[0] == 1000 and sel.min() >= 0 and sel.max() < 4000
    again = ops.sample_quantized(cu(pts, dev), 1000, rot, seed=7).cpu().numpy()
    np.testing.assert_array_equal(sel, again)
    other = ops.sample_quantized(cu(pts, dev), 1000, rot, seed=8).cpu().numpy()
    assert not np.array_equal(np.sort(sel), np.sort(other))

    def spread(ids):  # mean nearest-neighbour distance inside the sample
        _, d2 = oracle.knn(pts[ids], pts[ids], 2)
        return float(np.sqrt(d2[:, 1]).mean())

    ref = oracle.sampling_quantized(pts, 1000, np.random.default_rng(1))
    rnd = spread(np.random.default_rng(2).permutation(4000)[:1000])
    assert spread(sel) > 1.08 * rnd and abs(spread(sel) - spread(ref)) < 0.2 * spread(ref), (spread(sel), spread(ref), rnd)
    # edge cases: tiny clouds, n_support == n, heavy duplication (hash table / voxel halving cannot separate the points)
    tiny = cu(pts[:7], dev)
    assert sorted(ops.sample_quantized(tiny, 1, rot, 1).cpu().tolist())[0] in range(7)
    np.testing.assert_array_equal(np.sort(ops.sample_quantized(tiny, 7, rot, 1).cpu().numpy()), np.arange(7))
    dup = cu(np.repeat(pts[:10], 40, axis=0), dev)
    s = ops.sample_quantized(dup, 100, rot, 3).cpu().numpy()
    assert np.unique(s).shape[0] == 100


# ---- a8-a11 decoder -------------------------------------------------------------------------------------------------------

def _decode_inputs(g):
    return np.random.default_rng(int(g['latents_seed'])).standard_normal((1, 256, g['pts'].shape[0])).astype(np.float32)


@pytest.mark.parametrize('path', [0, 1])
def test_decode_golden(dev, net, oracle, weights, path):
    from ppsurf_b200 import ops
    g = load_golden('decode')
    latents = _decode_inputs(g)
    dec = ops.Decoder(net.packed()['decoder'], cu(g['pts'], dev), cu(latents[0].T, dev), chunk=100, path=path)
    qry = cu(g['qry'], dev)
    feat_proj = dec.projection(qry, cu(g['proj_ids'], dev, torch.int32)).cpu().numpy()
    assert np.abs(feat_proj - g['feat_proj'][0].T).max() < 5e-5 * max(1.0, np.abs(g['feat_proj']).max())
    feat_pn = ops.pointnet(dec.packed, cu(g['pts_local_ps'], dev), path).cpu().numpy()
    assert np.abs(feat_pn - g['feat_pn']).max() < 5e-5 * max(1.0, np.abs(g['feat_pn']).max())
    res = dec.decode(qry, want_logits=True, want_occ=True, want_idx=True)  # 192 queries in chunks of 100: ragged tail
    logits = res['logits'].cpu().numpy().T[None]
    assert np.abs(logits - g['logits']).max() < LOGIT_TOL
    assert np.abs(res['occ'].cpu().numpy() - g['occ'][0]).max() < LOGIT_TOL
    assert_knn_equal(oracle, g['pts'], g['qry'], res['idx'].cpu().numpy(), None, g['proj_ids'].astype(np.int64))
    # float64 oracle as the arbiter
    data = {'pts': g['pts'].T[None], 'latents': latents, 'pts_query': g['qry'][None],
            'pts_local_ps': g['pts_local_ps'][None], 'proj_ids': g['proj_ids'].astype(np.int64)[None]}
    ref64 = oracle.from_latent(weights, data, dtype=np.float64)
    assert np.abs(logits - ref64).max() < LOGIT_TOL
    # host-buffer entry point == device entry point
    occ_host = dec.decode_host(torch.from_numpy(g['qry']).pin_memory())
    np.testing.assert_array_equal(occ_host.numpy(), res['occ'].cpu().numpy())


def test_from_latent_reference_interface(dev, net):
    """the dict the reference driver builds (source/poco_utils.py:220-223): CPU pts_query, device patches"""
    g = load_golden('decode')
    latents = _decode_inputs(g)
    data = {'pts': cu(g['pts'].T[None], dev), 'latents': cu(latents, dev), 'pts_query': torch.from_numpy(g['qry'][None]),
            'pts_local_ps': cu(g['pts_local_ps'][None], dev)}
    out = net.from_latent(data)
    assert tuple(out.shape) == (1, 2, g['qry'].shape[0]) and out.device.type == 'cuda'
    assert np.abs(out.cpu().numpy() - g['logits']).max() < LOGIT_TOL
    assert data['proj_ids'].dtype == torch.int64 and tuple(data['proj_ids'].shape) == (1, g['qry'].shape[0], 64)
    # without caller-supplied patches everything comes from the fused call
    data2 = {'pts': data['pts'], 'latents': data['latents'], 'pts_query': torch.from_numpy(g['qry'][None])}
    out2 = net.from_latent(data2)
    assert np.abs(out2.cpu().numpy() - g['logits']).max() < LOGIT_TOL
    # supplied proj_ids (the train/test path)
    data3 = dict(data)
    data3['proj_ids'] = cu(g['proj_ids'][None], dev, torch.int64)
    out3 = net.from_latent(data3, has_proj_ids=True)
    assert np.abs(out3.cpu().numpy() - g['logits']).max() < LOGIT_TOL


@pytest.mark.parametrize('npl', [200, 100, 64, 130])
def test_decode_large_patches(dev, oracle, npl):
    """P > 64 (ppsurf_200nn: P = 200): the kNN runs with k = max(64, P), the global branch uses the first 64; on the tensor-core
    path a patch spans ceil(P/64) half-tiles whose maxima / attention partials are merged (atomic max, online softmax)"""
    import ppsurf_b200
    w = oracle.make_state_dict(43)
    net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, npl, 256)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}, strict=True)
    net = net.to(dev)
    rng = np.random.default_rng(9)
    pts = oracle.synthetic_cloud(3000, seed=31)
    latents = rng.standard_normal((1, 256, 3000)).astype(np.float32)
    qry = (pts[rng.integers(0, 3000, 45)] + 0.03 * rng.standard_normal((45, 3))).astype(np.float32)
    data = {'pts': pts.T[None], 'latents': latents, 'pts_query': qry[None],
            'pts_local_ps': oracle.get_pts_local_ps(pts, qry, npl)[None]}
    ref = oracle.from_latent(w, data, dtype=np.float64)
    for path in (1, 0):
        net.decode_path = path
        net._decoder_cache = None
        out = net.from_latent({'pts': cu(pts.T[None], dev), 'latents': cu(latents, dev), 'pts_query': torch.from_numpy(qry[None])})
        # unit-variance latents and seed-43 weights give logits up to |l| ~ 9 here: the absolute tolerance applies to the occupancy the
        # volume stores, the logits are held to the same accuracy relative to their size (2e-5 of the largest, i.e. 1e-4 at |l| = 5)
        got = out.cpu().numpy()
        assert np.abs(got - ref).max() < max(LOGIT_TOL, 2e-5 * np.abs(ref).max()), 'path {}'.format(path)
        assert np.abs(oracle.occupancy_from_logits(got) - oracle.occupancy_from_logits(ref)).max() < LOGIT_TOL, 'path {}'.format(path)


def test_grid_queries_bit_exact(dev, oracle):
    from ppsurf_b200 import ops
    import ppsurf_b200
    pts = oracle.synthetic_cloud(3000, seed=4)
    step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts, 17, 1)
    ref = oracle.dense_grid_queries(pts, 17, 1)
    got = ops.grid_queries(19, step, bmin_pad, device=dev).cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    part = ops.grid_queries(19, step, bmin_pad, first=1000, count=777, device=dev).cpu().numpy()
    np.testing.assert_array_equal(part, ref[1000:1777])


def test_region_growing_golden(dev, net):
    """create_volume's bookkeeping against the reference's _create_volume on an analytic field"""
    import ppsurf_b200
    g = load_golden('volume')
    model = ppsurf_b200.PPSurfModel(
        pointnet_latent_size=256, output_names=['imp_surf_sign'], in_channels=3, out_channels=2, k=64, lambda_l1=0.0,
        debug=False, in_file='x.txt', results_dir='results', padding_factor=0.05, name='t', network_latent_size=256,
        gen_subsample_manifold_iter=1, gen_subsample_manifold=10000, gen_resolution_global=17, num_pts_local=50,
        rec_batch_size=50000, gen_refine_iter=0, workers=1)

    class FakeDecoder:
        pts = torch.zeros((1, 3), device=dev)

    model.occupancy = lambda dec, q: torch.tanh(20.0 * (q.double().norm(dim=1) - 0.4)).float()
    vol = model.create_volume(FakeDecoder(), g['pts'], 17)
    np.testing.assert_array_equal(np.isnan(vol), np.isnan(g['volume']))
    assert np.nanmax(np.abs(vol - g['volume'])) < 1e-5


def test_region_growing_device_vs_oracle(dev, oracle):
    """device bookkeeping (ops.RegionVolume) against the oracle's create_volume on a field with several sign changes, zeros
    exactly on grid vertices and seeds on the volume border (clipped dilation boxes)"""
    import ppsurf_b200
    model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'x.txt', 'results', 0.05, 't', 256, 1, 10000, 33, 50,
                                    50000, 0, 1)

    class FakeDecoder:
        pts = torch.zeros((1, 3), device=dev)

    rng = np.random.default_rng(3)
    pts = np.concatenate([oracle.synthetic_cloud(700, seed=5), rng.uniform(-0.5, 0.5, (40, 3)).astype(np.float32),
                          np.array([[-0.5, -0.5, -0.5], [0.5, 0.5, 0.5]], dtype=np.float32)])

    def field_np(q):
        q = q.astype(np.float64)
        v = np.sin(9.0 * q[:, 0]) * np.cos(7.0 * q[:, 1]) + 0.5 * (np.linalg.norm(q, axis=1) - 0.4)
        return np.where(np.abs(q[:, 2]) < 1e-6, 0.0, v).astype(np.float32)  # a plane of exact zeros

    model.occupancy = lambda dec, q: torch.from_numpy(field_np(q.cpu().numpy())).to(dev)
    for res in (17, 33):
        vol = model.create_volume(FakeDecoder(), pts, res)
        ref = oracle.create_volume(field_np, pts, res)
        np.testing.assert_array_equal(np.isnan(vol), np.isnan(ref))
        np.testing.assert_array_equal(vol[~np.isnan(vol)], ref[~np.isnan(ref)])


def test_predict_pipeline_small(dev, net, oracle, weights):
    """encode_cloud (latent loop) + region-grown volume on a small cloud; the decode inside the volume is checked
    against the float64 oracle on the same latents"""
    import ppsurf_b200
    model = ppsurf_b200.PPSurfModel(
        pointnet_latent_size=256, output_names=['imp_surf_sign'], in_channels=3, out_channels=2, k=64, lambda_l1=0.0,
        debug=False, in_file='x.txt', results_dir='results', padding_factor=0.05, name='t', network_latent_size=256,
        gen_subsample_manifold_iter=2, gen_subsample_manifold=4200, gen_resolution_global=17, num_pts_local=50,
        rec_batch_size=5000, gen_refine_iter=0, workers=1)
    model.network.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    model = model.to(dev)
    pts = oracle.synthetic_cloud(7000, seed=12)
    rec = model.reconstruct(cu(pts[None], dev), resolution=17)
    vol = rec['volume']
    assert vol.shape == (19, 19, 19) and np.isfinite(vol[~np.isnan(vol)]).all() and (~np.isnan(vol)).sum() > 500
    latents = rec['latents'].cpu().numpy()
    assert np.isfinite(latents).all()
    coords = np.argwhere(~np.isnan(vol))
    coords = coords[np.all((coords > 0) & (coords < 18), axis=1)][::37][:40]
    qry = (coords.astype(np.float32) * rec['step'] + rec['bmin_pad']).astype(np.float32)
    data = {'pts': pts.T[None], 'latents': latents, 'pts_query': qry[None],
            'pts_local_ps': oracle.get_pts_local_ps(pts, qry, 50)[None]}
    ref = oracle.occupancy_from_logits(oracle.from_latent(weights, data, dtype=np.float64))[0]
    got = vol[coords[:, 0], coords[:, 1], coords[:, 2]]
    assert np.abs(got - ref).max() < 2e-4
    dense = model.dense_volume(rec['decoder'], pts, 17).cpu().numpy()
    assert np.abs(dense[coords[:, 0], coords[:, 1], coords[:, 2]] - ref).max() < 2e-4


def test_full_size_properties(dev, net, oracle):
    """BASELINE config 2 sizes (100k points, 131^3 grid): size-independent properties on a slab of the dense grid"""
    import ppsurf_b200
    from ppsurf_b200 import ops
    pts = oracle.synthetic_cloud(100000, seed=42)
    rng = np.random.default_rng(1)
    latents = torch.from_numpy(rng.standard_normal((100000, 256)).astype(np.float32)).to(dev)
    dec = ops.Decoder(net.packed()['decoder'], cu(pts, dev), latents, chunk=16384)
    step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts, 129, 1)
    first, count = 131 * 131 * 60, 131 * 131 * 3  # three z-slabs through the middle of the sphere: 51 483 queries
    qry = ops.grid_queries(131, step, bmin_pad, first=first, count=count, device=dev)
    res = dec.decode(qry, want_logits=True, want_occ=True, want_idx=True)
    idx, occ, logits = res['idx'].cpu().numpy(), res['occ'].cpu().numpy(), res['logits'].cpu().numpy()
    assert np.isfinite(logits).all() and np.all(np.abs(occ) <= 1.0)
    assert np.all(np.sort(idx, axis=1)[:, 1:] != np.sort(idx, axis=1)[:, :-1])  # no duplicate neighbours
    q = qry.cpu().numpy()
    d2 = oracle.sq_dist_f32(q[:, None, :], pts[idx.astype(np.int64)])
    assert np.all(np.diff(d2, axis=1) >= 0)  # ascending
    sample = rng.integers(0, count, 64)  # brute-force check of a sample (includes queries deep inside the sphere)
    ref_idx, _ = oracle.knn(pts, q[sample], 64)
    assert_knn_equal(oracle, pts, q[sample], idx[sample], None, ref_idx)
    # occupancy is a deterministic function of the query: decoding the same slab in other chunk sizes is bit identical
    dec2 = ops.Decoder(net.packed()['decoder'], cu(pts, dev), latents, chunk=5000)
    occ2 = dec2.decode(qry, want_logits=False, want_occ=True)['occ'].cpu().numpy()
    np.testing.assert_array_equal(occ, occ2)
    # the tensor-core path at full chunk size (128 tiles per launch, every SM busy) agrees with the fp32 path
    dec_tc = ops.Decoder(net.packed()['decoder'], cu(pts, dev), latents, chunk=16384, path=1)
    res_tc = dec_tc.decode(qry, want_logits=True, want_occ=True)
    assert np.abs(res_tc['logits'].cpu().numpy() - logits).max() < LOGIT_TOL
    assert np.abs(res_tc['occ'].cpu().numpy() - occ).max() < LOGIT_TOL
    # logits of the sample against the float64 oracle
    data = {'pts': pts.T[None], 'latents': latents.cpu().numpy().T[None], 'pts_query': q[sample][None],
            'pts_local_ps': oracle.get_pts_local_ps(pts, q[sample], 50)[None], 'proj_ids': ref_idx[None]}
    ref = oracle.from_latent(weights_for(net), data, dtype=np.float64)
    assert np.abs(logits[sample].T[None] - ref).max() < LOGIT_TOL


def test_tensor_core_path_matches_fp32_path(dev, net, oracle):
    """path 1 (tcgen05 split-fp16) against path 0 (fp32 SIMT) and the float64 oracle on 3001 queries: odd count (half-filled
    last tile), more tiles than SMs (persistent loop + ring wrap-around), far-away queries"""
    from ppsurf_b200 import ops
    rng = np.random.default_rng(77)
    pts = oracle.synthetic_cloud(6000, seed=5)
    latents = torch.from_numpy(rng.standard_normal((6000, 256)).astype(np.float32)).to(dev)
    qry = np.concatenate([pts[rng.integers(0, 6000, 2000)] + 0.02 * rng.standard_normal((2000, 3)),
                          rng.uniform(-0.6, 0.6, (1001, 3))]).astype(np.float32)
    out = []
    for path in (0, 1):
        dec = ops.Decoder(net.packed()['decoder'], cu(pts, dev), latents, chunk=1024, path=path)
        out.append(dec.decode(cu(qry, dev), want_logits=True, want_idx=True))
    l0, l1 = out[0]['logits'].cpu().numpy(), out[1]['logits'].cpu().numpy()
    assert np.isfinite(l1).all()
    assert np.abs(l0 - l1).max() < LOGIT_TOL
    sel = rng.integers(0, 3001, 48)
    data = {'pts': pts.T[None], 'latents': latents.cpu().numpy().T[None], 'pts_query': qry[sel][None],
            'pts_local_ps': oracle.get_pts_local_ps(pts, qry[sel], 50)[None]}
    ref = oracle.from_latent(weights_for(net), data, dtype=np.float64)
    assert np.abs(l1[sel].T[None] - ref).max() < LOGIT_TOL
    # the projection alone, so that a tensor-core error is not hidden behind the MLP
    dec0 = ops.Decoder(net.packed()['decoder'], cu(pts, dev), latents, path=0)
    dec1 = ops.Decoder(net.packed()['decoder'], cu(pts, dev), latents, path=1)
    idx = out[0]['idx'][:, :64].contiguous()
    f0, f1 = dec0.projection(cu(qry, dev), idx).cpu().numpy(), dec1.projection(cu(qry, dev), idx).cpu().numpy()
    assert np.abs(f0 - f1).max() < 2e-5 * max(1.0, np.abs(f0).max())


def weights_for(net):
    return {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}


def test_test_step_matches_oracle_forward(dev, oracle, weights):
    """`pps.py test` path (source/poco_model.py:134-162): network.forward on a batch that carries its own ids, cross-entropy
    loss and classification metrics against the oracle's forward on the same batch"""
    import ppsurf_b200
    model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'x.txt', 'results', 0.05, 'test', 256, 10, 10000,
                                    129, 50, 50000, 10, 0)
    model.network.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    model = model.to(dev).eval()  # Trainer.test puts the module in eval mode; in train mode forward() is the training path
    rng = np.random.default_rng(17)
    pts = oracle.synthetic_cloud(4200, seed=23)
    model.network.sampling_seed = 3
    batch = model.network.spatial_ids(cu(pts.T[None], dev))
    batch['pts'] = cu(pts.T[None], dev)
    qry = (pts[rng.integers(0, 4200, 300)] + 0.02 * rng.standard_normal((300, 3))).astype(np.float32)
    occ = (np.linalg.norm(qry, axis=1) < 0.4).astype(np.int64)
    batch['pts_query'] = cu(qry[None], dev)
    batch['pts_local_ps'] = cu(oracle.get_pts_local_ps(pts, qry, 50)[None], dev)
    batch['occ'] = cu(occ[None], dev)
    batch['shape_id'] = torch.tensor([0])
    batch['pc_file_in'] = ['synthetic.xyz']
    ref_in = {k: v.cpu().numpy() for k, v in batch.items() if isinstance(v, torch.Tensor) and k not in ('occ', 'shape_id')}
    res = model.test_step(batch, 0)
    ref_logits = oracle.network_forward(weights, ref_in)
    ref_loss = torch.nn.functional.cross_entropy(torch.from_numpy(ref_logits), torch.from_numpy(occ[None]))
    assert abs(float(res['loss']) - float(ref_loss)) < 1e-4
    margin = np.abs(ref_logits[0, 0] - ref_logits[0, 1]) > 1e-3  # labels are only defined away from exact ties
    ref_lab = np.argmax(ref_logits[0], axis=0)
    assert res['metrics_dict']['predictions'] == 300.0
    if margin.all():
        assert res['metrics_dict']['true_pos'] == float(((ref_lab == 1) & (occ == 1)).sum())
    assert len(model.test_step_outputs) == 1 and res['pc_file_in'] == 'synthetic.xyz'


# ---- round 2: binary freshness, real data on the CUDA path, config 3 at size, rank stitching, decoder cache ---------------

def test_binary_matches_sources(dev):
    """the prebuilt libppsurf_b200.so that travels to the GPU box was built from exactly the sources in the tree"""
    from ppsurf_b200 import _lib
    _lib.assert_binary_matches_sources()
    assert _lib.lib.pps_version() >= 200


def _bare_model(resolution=17, npl=50):
    import ppsurf_b200
    return ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'x.txt', 'results', 0.05, 't', 256, 1, 10000,
                                   resolution, npl, 50000, 0, 1)


@pytest.mark.parametrize('path', [0, 1])
def test_real_cloud_region_grown_volume_cuda(dev, net, weights_digest, path):
    """BASELINE configs[0] plumbing on REAL data through the CUDA path: 3000 vertices of an abc_minimal cloud decoded and
    region-grown at gen_resolution_global = 17; the golden volume was made by the UNMODIFIED reference (_create_volume,
    from_latent, normalize_patches; tests/golden/make_golden.py).  Same NaN mask, values within the north-star tolerance."""
    from ppsurf_b200 import ops
    g = load_golden('real_volume')
    assert str(g['digest']) == weights_digest
    pts = g['pts']
    latents = np.random.default_rng(int(g['latents_seed'])).standard_normal((1, 256, pts.shape[0])).astype(np.float32)
    dec = ops.Decoder(net.packed()['decoder'], cu(pts, dev), cu(latents[0].T, dev), chunk=1000, path=path)
    vol = _bare_model().create_volume(dec, pts, 17)
    np.testing.assert_array_equal(np.isnan(vol), np.isnan(g['volume']))
    assert np.isnan(vol).any() and np.nanmin(vol) < -0.5 and np.nanmax(vol) > 0.5
    assert np.nanmax(np.abs(vol - g['volume'])) < LOGIT_TOL


def test_config3_at_size(dev, oracle):
    """BASELINE config 3 at its real sizes: ppsurf_200nn (P = 200), 250k-point cloud, a slab of the dense 131^3 grid.
    Properties that do not need the oracle at size (ascending exact neighbours, chunk-size invariance, agreement of the
    tensor-core and the fp32 path) plus a 256-query sample against the float64 oracle."""
    import ppsurf_b200
    from ppsurf_b200 import ops
    w = oracle.make_state_dict(43)
    net200 = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 200, 256)
    net200.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}, strict=True)
    net200 = net200.to(dev)
    n = 250000
    pts = oracle.synthetic_cloud(n, seed=42)
    rng = np.random.default_rng(2)
    latents = torch.from_numpy(rng.standard_normal((n, 256)).astype(np.float32)).to(dev)
    step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts, 129, 1)
    first, count = 131 * 131 * 24, 131 * 131 * 2  # two z-slabs that cut the sphere's surface: 34 322 queries
    qry = ops.grid_queries(131, step, bmin_pad, first=first, count=count, device=dev)
    packed = net200.packed()['decoder']
    dec = ops.Decoder(packed, cu(pts, dev), latents, chunk=12500, path=1)
    res = dec.decode(qry, want_logits=True, want_occ=True, want_idx=True)
    idx, occ, logits = res['idx'].cpu().numpy(), res['occ'].cpu().numpy(), res['logits'].cpu().numpy()
    assert idx.shape == (count, 200) and np.isfinite(logits).all() and np.all(np.abs(occ) <= 1.0)
    q = qry.cpu().numpy()
    d2 = oracle.sq_dist_f32(q[:, None, :], pts[idx.astype(np.int64)])
    assert np.all(np.diff(d2, axis=1) >= 0)
    srt = np.sort(idx, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])
    # chunk-size invariance (the reference batches 25 000 queries for 200nn, configs/ppsurf_200nn.yaml:8): bit identical
    occ2 = ops.Decoder(packed, cu(pts, dev), latents, chunk=4097, path=1).decode(qry, want_logits=False, want_occ=True)['occ']
    np.testing.assert_array_equal(occ, occ2.cpu().numpy())
    # fp32 SIMT path on a part of the slab: two fp32-grade evaluations of the same field, each within LOGIT_TOL of the exact value
    part = slice(5000, 9000)
    l0 = ops.Decoder(packed, cu(pts, dev), latents, chunk=2000, path=0).decode(qry[part].contiguous(), want_logits=True)['logits']
    assert np.abs(l0.cpu().numpy() - logits[part]).max() < 2 * LOGIT_TOL
    # a 256-query sample against the float64 oracle (exact neighbours from the brute-force oracle kNN)
    sample = np.sort(rng.choice(count, 256, replace=False))
    ref_idx, _ = oracle.knn(pts, q[sample], 200)
    assert_knn_equal(oracle, pts, q[sample], idx[sample], None, ref_idx)
    data = {'pts': pts.T[None], 'latents': latents.cpu().numpy().T[None], 'pts_query': q[sample][None],
            'pts_local_ps': oracle.get_pts_local_ps(pts, q[sample], 200)[None], 'proj_ids': ref_idx[:, :64][None]}
    ref = oracle.from_latent(w, data, dtype=np.float64)
    # unit-variance random latents drive the logits of this stress case to |l| ~ 9 (a trained network stays around 1, where the
    # absolute north-star tolerance applies): the logits are held to the same RELATIVE accuracy, 2e-5 of the largest logit, and the
    # occupancy value the volume stores (bounded by 1) to the absolute tolerance
    assert np.abs(ref).max() > 4.0
    assert np.abs(logits[sample].T[None] - ref).max() < 2e-5 * np.abs(ref).max()
    assert np.abs(occ[sample] - oracle.occupancy_from_logits(ref)[0]).max() < LOGIT_TOL


def test_dealt_blocks_stitch_to_the_one_rank_volume(dev, net, oracle):
    """multi-GPU decode (bench.grid_blocks): the vertex list dealt to G ranks in blocks, each share decoded on its own, stitched
    back == the volume one rank decodes, bit for bit (the occupancy of a vertex does not depend on its launch neighbours)"""
    import bench
    import ppsurf_b200
    from ppsurf_b200 import ops
    pts = oracle.synthetic_cloud(20000, seed=8)
    latents = torch.from_numpy(np.random.default_rng(3).standard_normal((20000, 256)).astype(np.float32)).to(dev)
    dec = ops.Decoder(net.packed()['decoder'], cu(pts, dev), latents, chunk=37888, path=1)
    step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts, 33, 1)
    r, total = 35, 35 ** 3
    whole = dec.decode(ops.grid_queries(r, step, bmin_pad, device=dev), want_logits=False, want_occ=True)['occ'].cpu().numpy()
    for world in (2, 8):
        stitched = np.full((total,), np.nan, dtype=np.float32)
        for rank in range(world):
            spans = bench.grid_blocks(total, world, rank)
            q = torch.cat([ops.grid_queries(r, step, bmin_pad, first=f, count=c, device=dev) for f, c in spans])
            occ = dec.decode(q, want_logits=False, want_occ=True)['occ'].cpu().numpy()
            off = 0
            for f, c in spans:
                stitched[f:f + c] = occ[off:off + c]
                off += c
        np.testing.assert_array_equal(stitched, whole)


def test_two_same_shaped_clouds_back_to_back(dev, net, oracle):
    """the per-cloud decoder cache must not serve the previous cloud's kNN index / fc1 table to a second
    cloud of the same shape whose tensors landed on the recycled device addresses"""
    rng = np.random.default_rng(11)
    qry = torch.from_numpy(rng.uniform(-0.45, 0.45, (1, 300, 3)).astype(np.float32))
    outs = []
    for seed in (1, 2, 1):
        pts = oracle.synthetic_cloud(5000, seed=seed) * (1.0 if seed == 1 else 0.8)
        lat = np.random.default_rng(seed).standard_normal((1, 256, 5000)).astype(np.float32)
        data = {'pts': cu(pts.T[None], dev), 'latents': cu(lat, dev), 'pts_query': qry}
        a = net.from_latent(data).cpu().numpy()
        b = net.from_latent(data).cpu().numpy()  # same dict again: served from the cache
        np.testing.assert_array_equal(a, b)
        outs.append(a)
        del data
        torch.cuda.empty_cache()
    np.testing.assert_array_equal(outs[0], outs[2])
    assert np.abs(outs[0] - outs[1]).max() > 1e-3  # a different cloud gives a different field
    with pytest.raises(ValueError):
        net.from_latent({'pts': cu(pts.T[None], dev), 'latents': cu(lat, dev), 'pts_query': qry,
                         'pts_local_ps': torch.zeros((1, 300, 20, 3), device=dev)})


# ---- a3: fused FKAConv (csrc/fka_tc.cu) -----------------------------------------------------------------------------------

FKA_LAYERS = [  # (layer prefix, n_in, n_s, batch)
    ('encoder.cv0', 3000, 3000, 2),            # 3 (padded to 4) -> 64
    ('encoder.resnetb01.cv1', 2500, 2500, 1),  # 32 -> 32, one sample, ragged last tile
    ('encoder.resnetb10.cv1', 2000, 500, 3),   # 32 -> 32 strided
    ('encoder.resnetb21.cv1', 625, 625, 2),    # 128 -> 128: tiles straddle samples
    ('encoder.resnetb30.cv1', 625, 156, 5),    # 128 -> 128 strided, n_s < tile
    ('encoder.resnetb31.cv1', 156, 156, 4),    # 256 -> 256 (N = 256)
    ('encoder.resnetb41.cv1', 39, 39, 16),     # 512 -> 512: two N slices, 3.3 samples per tile
]


@pytest.mark.parametrize('name,n_in,n_s,batch', FKA_LAYERS)
def test_fkaconv_fused_vs_oracle(dev, net, oracle, weights, name, n_in, n_s, batch):
    """the fused tensor-core FKAConv against the float64 oracle of FKAConvLayer.forward (source/base/nn.py:592-652) and
    against the unfused fp32 kernels, for every channel configuration of the encoder, batches whose 128-row tiles straddle
    sample borders (per-sample InstanceNorm statistics), ragged last tiles and both N slices of the 512-wide layer"""
    from ppsurf_b200 import _lib, ops
    enc = net.packed()['encoder']
    w = enc['cv0'] if name == 'encoder.cv0' else enc[name.split('.')[1]]['cv1']
    cin_ref = weights[name + '.cv.weight'].shape[1]
    cin, cout = w.struct.cin, w.struct.cout
    rng = np.random.default_rng(n_in + cin)
    pts = np.stack([oracle.synthetic_cloud(n_in, seed=100 + i) * (1.0 + 0.1 * i) for i in range(batch)])  # [B,N,3]
    sup = pts[:, :n_s].copy()
    ids = np.stack([oracle.knn(pts[i], sup[i], 16)[0] for i in range(batch)])  # [B,Ns,16]
    x = rng.standard_normal((batch, n_in, cin_ref)).astype(np.float32)
    ref = oracle.fkaconv_layer(weights, name, x.transpose(0, 2, 1), pts.transpose(0, 2, 1), sup.transpose(0, 2, 1), ids,
                               dtype=np.float64)  # [B,Cout,Ns], before the BatchNorm that the packed layer folds in
    bn = 'encoder.bn0' if name == 'encoder.cv0' else name.rsplit('.', 1)[0] + '.bn1'
    s = weights[bn + '.weight'].astype(np.float64) / np.sqrt(weights[bn + '.running_var'].astype(np.float64) + 1e-5)
    sh = weights[bn + '.bias'].astype(np.float64) - weights[bn + '.running_mean'].astype(np.float64) * s
    ref = np.maximum(ref.transpose(0, 2, 1) * s + sh, 0.0)
    xp = np.concatenate([x, np.zeros((batch, n_in, cin - cin_ref), np.float32)], axis=2) if cin != cin_ref else x
    args = (w, cu(xp, dev), cu(pts, dev), cu(sup, dev), cu(ids, dev, torch.int32))
    launches0 = _lib.lib.pps_launch_count()
    fused = ops.fkaconv(*args).cpu().numpy()
    assert _lib.lib.pps_launch_count() - launches0 == 3  # two statistics passes + the fused kernel
    _lib.lib.pps_debug_fka_fused(0)
    try:
        unfused = ops.fkaconv(*args).cpu().numpy()
    finally:
        _lib.lib.pps_debug_fka_fused(1)
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(unfused - ref).max() < 2e-5 * scale
    # the tensor cores accumulate in fp32 WITHOUT round-to-nearest, so the error of a contraction grows with its length: 3 x K/16
    # accumulations per output.  K = 16 cin <= 4096 stays inside the 2e-5 of the fp32 kernels; the 512-channel layer (K = 8192, 1536
    # accumulations; identical error with and without the power-of-two weight scaling, i.e. not an operand-split effect) gets 4e-5.
    # The encoder as a whole is held to 5e-5 of the latent scale by test_encoder_golden with these kernels in place.
    tol = 2e-5 if 16 * cin <= 4096 else 4e-5
    assert np.abs(fused - ref).max() < tol * scale, np.abs(fused - ref).max() / scale


# ---- f3: marching cubes + refinement on the device, predict_step end to end ---------------------------------------------------

def _mc_field(r=41):
    ax = np.linspace(-0.6, 0.6, r, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing='ij')
    a = np.sqrt((x + 0.15) ** 2 + y ** 2 + z ** 2) - 0.3
    b = np.sqrt((x - 0.3) ** 2 + (y - 0.1) ** 2 + z ** 2) - 0.2
    return np.minimum(a, b).astype(np.float32)


def test_marching_cubes_device_vs_oracle(dev, oracle):
    """csrc/mcubes.cu against its numpy restatement: same vertices (bit exact, same order), same edge numbers, same faces; NaN cells
    emit nothing; a volume without a crossing gives an empty mesh"""
    from ppsurf_b200 import ops
    rng = np.random.default_rng(4)
    for vol in (_mc_field(), rng.standard_normal((12, 12, 12)).astype(np.float32)):
        for with_nan in (False, True):
            v = vol.copy()
            if with_nan:
                v[: v.shape[0] // 3] = np.nan
            verts, vert_edge, faces = ops.marching_cubes(cu(v, dev), 0.0)
            rv, re, rf = oracle.marching_cubes(v, 0.0)
            np.testing.assert_array_equal(vert_edge.cpu().numpy(), re)
            np.testing.assert_array_equal(verts.cpu().numpy(), rv)
            np.testing.assert_array_equal(faces.cpu().numpy(), rf)
    verts, _, faces = ops.marching_cubes(torch.ones((8, 8, 8), device=dev), 0.0)
    assert verts.shape == (0, 3) and faces.shape == (0, 3)


def test_vertex_refinement_device_vs_oracle(dev, oracle):
    """ten bisection sweeps on the device (ops.VertexRefiner) against the oracle's restatement of source/poco_utils.py:111-168 with
    the same analytic occupancy function: identical vertices"""
    from ppsurf_b200 import ops
    vol = _mc_field(33)
    vol[:4, :4, :4] = np.nan
    step, bmin_pad = np.float32(1.2 / 32), np.float32(-0.6)

    def field(q):
        q = q.astype(np.float64)
        a = np.sqrt((q[:, 0] + 0.15) ** 2 + q[:, 1] ** 2 + q[:, 2] ** 2) - 0.3
        b = np.sqrt((q[:, 0] - 0.3) ** 2 + (q[:, 1] - 0.1) ** 2 + q[:, 2] ** 2) - 0.2
        return np.minimum(a, b).astype(np.float32)

    verts, vert_edge, faces = ops.marching_cubes(cu(vol, dev), 0.0)
    ref = oracle.refine_vertices(field, vol, verts.cpu().numpy(), step, bmin_pad, 10)
    refiner = ops.VertexRefiner(cu(vol, dev), verts, vert_edge, step, bmin_pad)
    for _ in range(10):
        refiner.update(cu(field(refiner.v.cpu().numpy()), dev))
    got = refiner.result().cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    assert np.abs(field(got)).max() < 1e-4


def test_predict_step_writes_a_mesh(dev, net, oracle, weights, tmp_path):
    """`pps.py rec` path end to end on the device (source/poco_model.py:183-273): encoder, region-grown volume, marching cubes,
    refinement, cleaning, de-normalisation of a single-file input, PLY export  --  with an occupancy head whose sign is the analytic
    sphere (random-init weights have no surface), so that the mesh can be checked: closed, radius 0.4 in the input frame"""
    import ppsurf_b200
    from ppsurf_b200 import mesh as mesh_utils
    pts_file = str(tmp_path / 'sphere.npy')
    raw = oracle.synthetic_cloud(9000, seed=3) * 3.0 + np.array([10.0, -2.0, 5.0], dtype=np.float32)
    np.save(pts_file, raw)
    center, scale = mesh_utils.get_points_normalization_info(raw.astype(np.float64), 0.05)
    pts_ms = ((raw - center) / scale).astype(np.float32)
    model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, pts_file, str(tmp_path / 'results'), 0.05, 't', 256,
                                    2, 5000, 33, 50, 50000, 10, 0)
    model.network.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    model = model.to(dev)
    r_ms = 0.4 * 3.0 / scale
    real_occupancy = model.occupancy
    calls = []

    def occupancy(decoder, q):
        calls.append(int(q.shape[0]))
        real = real_occupancy(decoder, q)  # the real decode runs (and must be finite); its sign is replaced by the sphere's
        assert torch.isfinite(real).all()
        return torch.tanh(40.0 * (q.norm(dim=1) - r_ms))

    model.occupancy = occupancy
    assert model.predict_step({'pts_ms': cu(pts_ms[None], dev), 'pc_file_in': [pts_file]}, 0) == 0
    out = os.path.join(str(tmp_path / 'results'), 'sphere.npy', 'sphere.npy.ply')
    assert os.path.exists(out) and len(calls) > 10
    verts = mesh_utils.read_ply_vertices(out)
    radius = np.linalg.norm(verts - center, axis=1)
    assert verts.shape[0] > 500 and np.abs(radius - 1.2).max() < 2e-3  # bisection: step / 2^10 in model space, times the scale


def test_prepare_batch_on_device(dev, net, oracle, weights):
    """f4: the DataLoader workers' per-item CPU work (kd-tree patches, patch normalisation, get_data_poco) for a collated batch of two
    clouds on the device: labels, exact projection neighbours, bit-exact patches, index tensors equal to the oracle kNN on the
    sampled supports; the prepared dict drives network.forward like the reference's batch does"""
    from ppsurf_b200 import data_pipeline
    rng = np.random.default_rng(31)
    n, q = 4200, 150
    pts = np.stack([oracle.synthetic_cloud(n, seed=41), oracle.synthetic_cloud(n, seed=42) * 0.9])
    raw = np.stack([np.concatenate([p, p[:800] + 0.001]) for p in pts]).astype(np.float32)  # the raw cloud is denser than the subsample
    qry = np.stack([(p[rng.integers(0, n, q)] + 0.03 * rng.standard_normal((q, 3))).astype(np.float32) for p in pts])
    dist = (np.linalg.norm(qry, axis=2) - np.array([0.4, 0.36])[:, None]).astype(np.float32)
    dist[0, :5] = 0.0
    net.sampling_seed = 5
    batch = data_pipeline.prepare_batch(net, {'pts_ms': torch.from_numpy(pts), 'pts_query_ms': torch.from_numpy(qry),
                                              'pts_raw_ms': torch.from_numpy(raw), 'imp_surf_dist_ms': torch.from_numpy(dist)})
    np.testing.assert_array_equal(batch['occ'].cpu().numpy(), (dist > 0).astype(np.int64))
    assert tuple(batch['pts'].shape) == (2, 3, n) and tuple(batch['pts_query'].shape) == (2, 3, q)
    assert batch['proj_ids'].dtype == torch.int64 and tuple(batch['pts_local_ps'].shape) == (2, q, 50, 3)
    for b in range(2):
        assert_knn_equal(oracle, pts[b], qry[b], batch['proj_ids'][b].cpu().numpy(), None, oracle.knn(pts[b], qry[b], 64)[0])
        ref_loc = oracle.get_pts_local_ps(raw[b], qry[b], 50)
        np.testing.assert_array_equal(np.sort(batch['pts_local_ps'][b].cpu().numpy(), axis=1), np.sort(ref_loc, axis=1))
        sup1 = batch['support1'][b].T.cpu().numpy()
        assert_knn_equal(oracle, pts[b], sup1, batch['ids01'][b].cpu().numpy(), None, oracle.knn(pts[b], sup1, 16)[0])
    logits = net.forward(dict(batch))
    assert tuple(logits.shape) == (2, 2, q) and torch.isfinite(logits).all()
    ref_in = {k: v.cpu().numpy() for k, v in batch.items() if isinstance(v, torch.Tensor)
              and k.startswith(('pts', 'support', 'ids', 'proj_ids'))}
    ref_in['pts_query'] = qry  # the oracle takes the queries as [B,Q,3]
    # encoder AND decoder in one go: the latents carry the encoder's tolerance (5e-5 of their scale, test_encoder_golden) into a
    # decoder whose logits move by a few times that; the float64 oracle is the arbiter
    ref = oracle.network_forward(weights, ref_in, dtype=np.float64)
    assert np.abs(logits.cpu().numpy() - ref).max() < 3 * LOGIT_TOL
