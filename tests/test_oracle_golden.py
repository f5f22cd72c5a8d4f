"""Pins the CPU oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by
tests/golden/make_golden.py in the build container).  CPU only."""
import os

import numpy as np
import pytest

from conftest import load_golden


def _knn_sets_equal(pts, qry, idx_a, idx_b, oracle):
    """index sets identical except where the k-th distance ties"""
    da = np.sort(oracle.sq_dist_f32(qry[:, None, :], pts[idx_a]), axis=1)
    db = np.sort(oracle.sq_dist_f32(qry[:, None, :], pts[idx_b]), axis=1)
    np.testing.assert_array_equal(da, db)
    for r in range(idx_a.shape[0]):
        if set(idx_a[r]) != set(idx_b[r]):
            diff = set(idx_a[r]) ^ set(idx_b[r])
            kth = da[r, -1]
            d = oracle.sq_dist_f32(qry[r][None], pts[list(diff)])
            assert np.all(d == kth), 'row {} differs outside a tie'.format(r)


def test_param_inventory(oracle, weights):
    spec = oracle.param_spec()
    assert len(spec) == 455
    assert sum(int(np.prod(s)) for s in spec.values()) == 13774258
    assert list(spec) == list(weights)


def test_knn(oracle):
    g = load_golden('knn')
    idx, d2 = oracle.knn(g['pts'], g['qry'], 64)
    assert np.all(np.diff(d2, axis=1) >= 0)
    _knn_sets_equal(g['pts'], g['qry'], idx, g['idx64'].astype(np.int64), oracle)
    idx1, _ = oracle.knn(g['pts'], g['qry'], 1)
    _knn_sets_equal(g['pts'], g['qry'], idx1, g['idx1'].astype(np.int64), oracle)
    # k > N is clamped to N (source/poco_utils.py:259-260)
    idx_s, _ = oracle.knn(g['pts'][:10], g['qry'][:5], 16)
    assert idx_s.shape == (5, 10)
    np.testing.assert_array_equal(idx_s, g['idx_small'])


def test_patches(oracle):
    g = load_golden('patches')
    loc = oracle.get_pts_local_ps(g['pts'], g['qry'], 50)
    # the same 50 neighbours in the same order and the same fp32 arithmetic: bit exact unless ties reorder
    np.testing.assert_allclose(np.sort(loc, axis=1), np.sort(g['pts_local_ps'], axis=1), rtol=0, atol=0)
    assert np.abs(np.linalg.norm(loc, axis=2).max(axis=1) - 1).max() < 1e-6


def test_fkaconv_layer_and_resblock(oracle, weights, weights_digest):
    g = load_golden('fkaconv')
    assert str(g['digest']) == weights_digest
    pts, sup = g['pts'].T[None], g['support'].T[None]
    ids = g['ids'].astype(np.int64)[None]
    y = oracle.fkaconv_layer(weights, 'encoder.resnetb01.cv1', g['x32'], pts, sup, ids)
    np.testing.assert_allclose(y, g['y_fka'], rtol=1e-4, atol=1e-5 * np.abs(g['y_fka']).max())
    y = oracle.residual_block(weights, 'encoder.resnetb10', g['x64'], pts, sup, ids)
    np.testing.assert_allclose(y, g['y_rb'], rtol=1e-4, atol=1e-5 * np.abs(g['y_rb']).max())
    ids_same = g['ids_same'].astype(np.int64)[None]
    y = oracle.residual_block(weights, 'encoder.resnetb01', g['x64'], pts, pts, ids_same)
    np.testing.assert_allclose(y, g['y_rb_same'], rtol=1e-4, atol=1e-5 * np.abs(g['y_rb_same']).max())


def test_encoder(oracle, weights, weights_digest):
    g = load_golden('encoder')
    assert str(g['digest']) == weights_digest
    data = {k: (v.astype(np.int64) if k.startswith('ids') else v) for k, v in g.items() if k not in ('latents', 'digest')}
    lat = oracle.fkaconv_network(weights, data)
    scale = np.abs(g['latents']).max()
    assert np.abs(lat - g['latents']).max() < 2e-5 * scale


def test_decode(oracle, weights, weights_digest):
    g = load_golden('decode')
    assert str(g['digest']) == weights_digest
    latents = np.random.default_rng(int(g['latents_seed'])).standard_normal((1, 256, g['pts'].shape[0])).astype(np.float32)
    assert abs(latents.astype(np.float64).sum() - float(g['latents_sum'])) < 1e-6
    pts = g['pts'].T[None]
    ids = g['proj_ids'].astype(np.int64)[None]
    fp = oracle.interp_attention(weights, latents, pts, g['qry'][None], ids)
    assert np.abs(fp - g['feat_proj']).max() < 2e-5 * max(1.0, np.abs(g['feat_proj']).max())
    fn = oracle.pointnet_feat(weights, np.swapaxes(g['pts_local_ps'], 1, 2))
    assert np.abs(fn - g['feat_pn']).max() < 2e-5 * max(1.0, np.abs(g['feat_pn']).max())
    data = {'pts': pts, 'latents': latents, 'pts_query': g['qry'][None],
            'pts_local_ps': oracle.get_pts_local_ps(g['pts'], g['qry'], 50)[None]}
    logits = oracle.from_latent(weights, data)
    assert np.abs(logits - g['logits']).max() < 1e-4  # the north-star tolerance, fp32 vs fp32
    assert np.abs(oracle.occupancy_from_logits(logits) - g['occ']).max() < 1e-4
    # the float64 oracle is at least as close to the reference as the fp32 one is
    logits64 = oracle.from_latent(weights, dict(data), dtype=np.float64)
    assert np.abs(logits64 - g['logits']).max() < 1e-4


def test_region_growing(oracle):
    g = load_golden('volume')

    def predict(q):
        return np.tanh(20.0 * (np.linalg.norm(q, axis=1) - 0.4)).astype(np.float32)

    step, bmin_pad, _ = oracle.grid_definition(g['pts'], 17, 1)
    assert np.float32(step) == g['step'] and np.float32(bmin_pad) == g['bmin_pad']
    vol = oracle.create_volume(predict, g['pts'], 17, padding=1)
    np.testing.assert_array_equal(np.isnan(vol), np.isnan(g['volume']))
    np.testing.assert_array_equal(np.nan_to_num(vol), np.nan_to_num(g['volume']))
    dense = oracle.dense_grid_queries(g['pts'], 17, 1)
    assert dense.shape == (19 ** 3, 3)


def test_latent_schedule(oracle):
    rng = np.random.default_rng(3)
    sched = oracle.latent_loop_schedule(25000, 10000, 3, rng)
    counts = np.zeros(25000, dtype=np.int64)
    for ids in sched:
        assert ids.shape[0] == 10000
        counts[np.unique(ids)] += 1
    assert counts.min() >= 3
    small = oracle.latent_loop_schedule(500, 10000, 2, rng)
    assert len(small) == 2 and all(np.array_equal(s, np.arange(500)) for s in small)


def test_torch_timing_twin_matches(oracle, weights):
    """the multi-threaded torch restatement used for CPU timing gives the numpy oracle's / the reference's numbers"""
    import torch
    from oracle import ppsurf_oracle_torch as OT
    g = load_golden('decode')
    latents = np.random.default_rng(int(g['latents_seed'])).standard_normal((1, 256, g['pts'].shape[0])).astype(np.float32)
    occ, logits = OT.from_latent(weights, torch.from_numpy(g['pts']), torch.from_numpy(latents[0].T.copy()),
                                 torch.from_numpy(g['qry']), torch.from_numpy(g['proj_ids'].astype(np.int64)),
                                 torch.from_numpy(g['pts_local_ps']))
    assert np.abs(logits.numpy().T[None] - g['logits']).max() < 1e-4
    assert np.abs(occ.numpy() - g['occ'][0]).max() < 1e-4


def test_real_cloud_region_grown_volume(oracle, weights, weights_digest):
    """BASELINE configs[0] plumbing on REAL data: 3000 vertices of an abc_minimal cloud, normalised, decoded and region-grown at
    gen_resolution_global = 17 by the unmodified reference (its _create_volume, from_latent, normalize_patches and the kd-tree
    stand-in); the oracle's restatement of the same chain must produce the same volume"""
    g = load_golden('real_volume')
    assert str(g['digest']) == weights_digest
    pts = g['pts']
    latents = np.random.default_rng(int(g['latents_seed'])).standard_normal((1, 256, pts.shape[0])).astype(np.float32)
    assert abs(float(latents.astype(np.float64).sum()) - float(g['latents_sum'])) < 1e-6

    def predict(q):
        data = {'pts': pts.T[None], 'latents': latents, 'pts_query': q[None],
                'pts_local_ps': oracle.get_pts_local_ps(pts, q, 50)[None]}
        return oracle.occupancy_from_logits(oracle.from_latent(weights, data))[0]

    vol = oracle.create_volume(predict, pts, 17)
    np.testing.assert_array_equal(np.isnan(vol), np.isnan(g['volume']))
    finite = vol[~np.isnan(vol)]
    assert np.isnan(vol).any() and finite.min() < -0.5 and finite.max() > 0.5  # region-grown, with an inside and an outside
    assert np.nanmax(np.abs(vol - g['volume'])) < 2e-5


def test_train_oracle_twin_matches_the_reference_training_step(weights, weights_digest):
    """config 5: the torch twin of the TRAINING step (oracle/ppsurf_train_oracle.py, the checker of the CUDA backward) against the
    golden vectors of the unmodified reference in train mode (tests/golden/make_golden_train.py): loss, logits, gradient norms / sums /
    sampled entries of all 298 parameter tensors in float64, and the buffers the step updates"""
    import sys
    import torch
    from oracle import ppsurf_train_oracle as T
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from make_golden_train import sample_positions
    g = load_golden('train_step')
    assert str(g['digest']) == weights_digest

    def batch(dtype):
        out = {}
        for key in g:
            if key.startswith('in_'):
                t = torch.from_numpy(g[key])
                out[key[3:]] = t.long() if t.dtype == torch.int32 else t.to(dtype)
        return out

    s = T.State(weights, dtype=torch.float64)
    loss, logits = T.training_step(s, batch(torch.float64), dropout=0.0)
    assert abs(float(loss) - float(g['loss64'])) < 1e-12
    assert np.abs(logits.numpy() - g['logits64']).max() < 1e-6
    grads = s.grads()
    assert list(grads) == [str(n) for n in g['grad_names']]
    for i, (name, gr) in enumerate(grads.items()):
        flat = gr.numpy().reshape(-1)
        scale = max(float(g['grad_norm'][i]), 1e-12)
        assert abs(np.sqrt((flat * flat).sum()) - g['grad_norm'][i]) <= 1e-9 * scale + 1e-15, name
        assert np.abs(flat[sample_positions(i, flat.size)] - g['grad_samples'][i]).max() <= 1e-9 * scale + 1e-15, name
    # the float32 run of the twin reproduces the reference's float32 loss and buffers (BatchNorm running statistics, norm_radius)
    s32 = T.State(weights)
    loss32, _ = T.training_step(s32, batch(torch.float32), dropout=0.0)
    assert abs(float(loss32) - float(g['loss'])) < 1e-6
    for key in g:
        if key.startswith('buf_'):
            assert np.abs(s32.b[key[4:]].numpy() - g[key]).max() < 1e-5, key
