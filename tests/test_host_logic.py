"""CPU-only tests of the host side: the C-ABI library loads and exports every declared symbol, the weight packing is
algebraically equal to the reference formulation, the module keeps the reference's state_dict layout, and the product
path refuses to run without a device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ppsurf_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pps_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from ppsurf_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), '{} declared in include/ppsurf_b200.h but not exported'.format(name)
        assert name in _lib.SIGNATURES, '{} has no ctypes signature'.format(name)
    assert sorted(_lib.SIGNATURES) == declared
    assert _lib.lib.pps_compiled_arch() == 100
    assert _lib.lib.pps_knn_index_bytes(100000) > 100000 * 16  # pure size arithmetic, no device needed


def test_struct_layouts_match_header():
    """field order of the ctypes mirrors == field order of the C structs"""
    from ppsurf_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'ppsurf_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    for cname, mirror in (('pps_decoder_weights', _lib.DecoderWeights), ('pps_fkaconv_weights', _lib.FKAConvWeights)):
        body = re.search(r'typedef struct ' + cname + r' \{(.*?)\} ' + cname, text, flags=re.S).group(1)
        fields = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if not decl.startswith('const') else decl.split('*', 1)[1]
            fields += [n.strip().lstrip('*') for n in names.split(',')]
        assert fields == [f[0] for f in mirror._fields_], cname


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-device behaviour')
def test_no_cpu_fallback(weights):
    import ppsurf_b200
    net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
    data = {'pts': torch.zeros(1, 3, 100), 'latents': torch.zeros(1, 256, 100), 'pts_query': torch.zeros(1, 4, 3)}
    with pytest.raises(RuntimeError):
        net.from_latent(data)
    with pytest.raises(Exception):
        ppsurf_b200.ops.require_device()


def test_state_dict_layout_matches_reference(oracle, weights):
    import ppsurf_b200
    net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
    sd = net.state_dict()
    spec = oracle.param_spec()
    assert list(sd.keys()) == list(spec.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(spec[k]), k
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    model = ppsurf_b200.PPSurfModel(
        pointnet_latent_size=256, output_names=['imp_surf_sign'], in_channels=3, out_channels=2, k=64, lambda_l1=0.0,
        debug=False, in_file='x.txt', results_dir='results', padding_factor=0.05, name='ppsurf_50nn',
        network_latent_size=256, gen_subsample_manifold_iter=10, gen_subsample_manifold=10000,
        gen_resolution_global=129, num_pts_local=50, rec_batch_size=50000, gen_refine_iter=10, workers=8)
    assert [k for k in model.state_dict()] == ['network.' + k for k in spec]


def test_packed_decoder_equals_reference_formulation(oracle, weights):
    from ppsurf_b200 import packing
    from packed_math import decode_packed
    g = load_golden('decode')
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
    p = packing.pack_decoder(sd, 'cpu', k=64, num_pts_local=50)
    latents = np.random.default_rng(int(g['latents_seed'])).standard_normal((1, 256, g['pts'].shape[0])).astype(np.float32)
    fp, fn, logits = decode_packed(p, torch.from_numpy(g['pts']), torch.from_numpy(latents[0].T.copy()),
                                   torch.from_numpy(g['qry']), torch.from_numpy(g['proj_ids'].astype(np.int64)),
                                   torch.from_numpy(g['pts_local_ps']))
    assert np.abs(fp.numpy() - g['feat_proj'][0].T).max() < 5e-5
    assert np.abs(fn.numpy() - g['feat_pn']).max() < 5e-5
    assert np.abs(logits.numpy().T[None] - g['logits']).max() < 1e-4


def test_packed_fkaconv_kernel_layout(weights):
    """cv_w column m*cin + c holds cv.weight[o,c,0,m] scaled by the folded BatchNorm"""
    from ppsurf_b200 import packing
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
    p = packing.pack_fkaconv(sd, 'encoder.resnetb01.cv1', 'cpu', 'silu', bn='encoder.resnetb01.bn1')
    cv = sd['encoder.resnetb01.cv1.cv.weight'].double()
    s = sd['encoder.resnetb01.bn1.weight'].double() / torch.sqrt(sd['encoder.resnetb01.bn1.running_var'].double() + 1e-5)
    w = p.tensors['cv_w'].double().view(32, 16, 32)
    for o, m, c in ((0, 0, 0), (3, 5, 7), (31, 15, 31)):
        assert abs(float(w[o, m, c]) - float(cv[o, c, 0, m] * s[o])) < 1e-6
    assert p.struct.cin == 32 and p.struct.cout == 32 and p.struct.act == 1 and p.struct.out_relu == 1
    enc = packing.pack_encoder(sd, 'cpu')
    assert enc['cv3d'][0].w.shape == (512, 1024) and enc['cv3d'][1].w.shape == (512, 512)
    assert enc['cv5'][0].w.shape == (1024, 1024) and enc['cv0d'][0].w.shape == (64, 128)


def test_grid_definition_matches_oracle(oracle):
    import ppsurf_b200
    pts = oracle.synthetic_cloud(5000, 3)
    step, bmin_pad, ids = ppsurf_b200.PPSurfModel.grid_definition(pts, 129, 1)
    s2, b2, i2 = oracle.grid_definition(pts, 129, 1)
    assert step == np.float32(s2) and bmin_pad == np.float32(b2)
    np.testing.assert_array_equal(ids, i2)


def test_sampling_rotations():
    """host side of the support sampling: proper rotations, reproducible from the seed"""
    from ppsurf_b200.sampling import random_rotations
    r = random_rotations(np.random.default_rng(3), 12).reshape(12, 3, 3).astype(np.float64)
    for m in r:
        assert np.abs(m @ m.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(m) - 1) < 1e-6
    np.testing.assert_array_equal(random_rotations(np.random.default_rng(3), 12).reshape(12, 3, 3), r.astype(np.float32))
    assert np.abs(r[0] - r[1]).max() > 1e-3


def test_synthetic_generators_match_oracle(oracle, weights):
    import ppsurf_b200
    from ppsurf_b200 import synthetic
    net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
    sd = synthetic.make_state_dict(net, 42)
    assert list(sd) == list(weights)
    for k in sd:
        np.testing.assert_array_equal(sd[k].numpy(), np.asarray(weights[k]), err_msg=k)
    np.testing.assert_array_equal(synthetic.synthetic_cloud(1000, 3), oracle.synthetic_cloud(1000, 3))


def test_loss_and_metrics_follow_the_reference_formulas():
    """compute_loss / calc_metrics of the module mirror against a direct restatement of source/poco_model.py:75-102 and
    source/base/metrics.py:41-84 (int32 sums of predicted / ground-truth indicator vectors)"""
    import ppsurf_b200
    g = torch.Generator().manual_seed(5)
    pred = torch.randn((1, 2, 500), generator=g)
    occ = (torch.rand((1, 500), generator=g) > 0.4).to(torch.int64)
    model = ppsurf_b200.PPSurfModel.__new__(ppsurf_b200.PPSurfModel)
    loss, mean, comps = ppsurf_b200.PPSurfModel.compute_loss(model, pred, {'occ': occ})
    ref = torch.nn.functional.cross_entropy(pred, occ, reduction='none')
    assert comps.shape == (1, 1, 500) and torch.equal(comps[0], ref)
    assert float(loss) == pytest.approx(float(ref.mean())) and mean.shape == (1,)
    m = ppsurf_b200.PPSurfModel.calc_metrics(pred, {'occ': occ})
    p_int = (torch.argmax(pred, dim=1).to(torch.float32).squeeze() > 0).to(torch.int32)
    g_int = (occ.squeeze() > 0).to(torch.int32)
    tp = float(((p_int + g_int) == 2).sum())
    fp = float(((p_int * 2 + g_int) == 2).sum())
    fn = float(((p_int + 2 * g_int) == 2).sum())
    tn = 500.0 - float(torch.nonzero(p_int + g_int).shape[0])
    assert (m['true_pos'], m['false_pos'], m['false_neg'], m['true_neg']) == (tp, fp, fn, tn)
    assert m['accuracy'] == (tp + tn) / 500.0 and m['precision'] == tp / (tp + fp) and m['recall'] == tp / (tp + fn)
    assert m['f1_score'] == pytest.approx(2.0 * m['precision'] * m['recall'] / (m['precision'] + m['recall']))
    # degenerate case: nothing predicted positive -> precision NaN, F1 NaN (reference conventions)
    m0 = ppsurf_b200.PPSurfModel.calc_metrics(torch.tensor([[[1.0, 1.0], [0.0, 0.0]]]), {'occ': torch.tensor([[1, 0]])})
    assert np.isnan(m0['precision']) and m0['recall'] == 0.0 and np.isnan(m0['f1_score'])


def test_bench_emits_one_clean_stdout_line():
    """bench.py's contract: ONE JSON line on stdout.  Library chatter on fd 1 (NCCL prints its version banner there) must land
    on stderr"""
    import subprocess
    import sys
    code = ('import os, sys; sys.path.insert(0, {!r}); import bench; q = bench.QuietStdout(); os.write(1, b"banner\\n"); '
            'print("chatter"); q.emit("{{\\"ok\\": 1}}"); print("more")').format(ROOT)
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True)
    assert p.stdout == '{"ok": 1}\n' and 'banner' in p.stderr and 'more' in p.stderr


def test_tensor_core_pack_layout(weights):
    """The projection kernel's weight pack, decoded on the CPU exactly the way the kernel addresses it: stage s of layer l holds,
    for each CTA r of the pair, rows [128r, 128r+128) of W as [hi k8-block 0 | hi k8-block 1 | lo block 0 | lo block 1] with
    blocks of [128 rows][8 fp16]; fc_query is the N = 64 operand: slot c (one 64-column chunk) holds, for each CTA r, heads [32r, 32r+32)
    as four k16 steps of 2 KB [hi block 0 | hi block 1 | lo block 0 | lo block 1], blocks of [32 rows][8 fp16].  hi + lo must reproduce
    the fp32 weights to 2^-21 relative."""
    from ppsurf_b200 import packing
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
    p = packing.pack_decoder(sd, 'cpu', 64, 50)
    pack = p.tensors['tc_wpack'].numpy().view(np.float16)  # 2 bytes per element
    stage_elems = 8192 // 2  # one CTA's ring slot

    def block(stage_base, r0, rows=128):  # -> [rows, 16] float32 (hi + lo) of one 8 KB slot
        slot = pack[stage_base:stage_base + stage_elems].astype(np.float32)
        hi = slot[:2048].reshape(2, rows, 8)   # [k8 block][row][8]
        lo = slot[2048:].reshape(2, rows, 8)
        return (hi + lo).transpose(1, 0, 2).reshape(rows, 16)

    for layer, name in enumerate(('projection.fc2.weight', 'projection.fc3.weight')):
        w = np.asarray(weights[name]).reshape(256, 256).astype(np.float32)
        for s in (0, 7, 15):
            for r in (0, 1):
                got = block((layer * 16 + s) * 2 * stage_elems + r * stage_elems, 0)
                ref = w[128 * r:128 * r + 128, 16 * s:16 * s + 16]
                assert np.abs(got - ref).max() <= 2.0 ** -20 * np.abs(ref).max()
    wq = np.asarray(weights['projection.fc_query.weight']).reshape(64, 256).astype(np.float32)
    base = 2 * 16 * 2 * stage_elems
    for c in (0, 3):
        for r in (0, 1):
            slot = pack[base + (c * 2 + r) * stage_elems:base + (c * 2 + r + 1) * stage_elems].astype(np.float32)
            for sub in (0, 2):
                step = slot[sub * 1024:(sub + 1) * 1024]  # 2 KB = 1024 fp16
                hi, lo = step[:512].reshape(2, 32, 8), step[512:].reshape(2, 32, 8)
                got = (hi + lo).transpose(1, 0, 2).reshape(32, 16)
                k0 = 64 * c + 16 * sub
                assert np.abs(got - wq[32 * r:32 * r + 32, k0:k0 + 16]).max() <= 2.0 ** -20 * np.abs(wq).max()
    assert pack.size * 2 == 2 * 16 * 16384 + 4 * 16384


def test_pointnet_pack_layouts(weights):
    """The PointNet kernels' weight packs, decoded the way the kernels address them (csrc/pointnet_tc.cu).
    pn_stn (CTA pairs): 7 ring-slot fills per pair-tile, each stored as [CTA 0's slot | CTA 1's slot]: conv0b and stn.conv1 (8 KB per
    CTA: 4 k16 steps x its 32 of the 64 weight rows), stn.conv2 (16 KB: 4 steps x 64 of 128 rows), stn.conv3 (4 fills of 16 KB: 2 steps x
    its 128 of the 256 features); a step is [hi k8-block 0 | hi block 1 | lo block 0 | lo block 1], blocks of [rows][8 fp16].
    pn_feat: conv1 (4 steps x 64 rows = 16 KB), conv2 (4 steps x 128 rows, 8 KB each: steps 0,1 at +16 KB, steps 2,3 at +32 KB)."""
    from ppsurf_b200 import packing
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
    p = packing.pack_decoder(sd, 'cpu', 64, 50)
    t = p.tensors

    def steps(buf, rows, nsteps):  # fp16 view of nsteps k16 steps of `rows` weight rows -> [rows, 16 * nsteps] float32 (hi + lo)
        out = []
        per = rows * 32  # fp16 elements per step: 4 blocks of rows x 8
        for s in range(nsteps):
            st = buf[s * per:(s + 1) * per].astype(np.float32)
            hi, lo = st[:per // 2].reshape(2, rows, 8), st[per // 2:].reshape(2, rows, 8)
            out.append((hi + lo).transpose(1, 0, 2).reshape(rows, 16))
        return np.concatenate(out, axis=1)

    def close(got, ref):
        assert np.abs(got - ref).max() <= 2.0 ** -20 * max(np.abs(ref).max(), 1e-30)

    stn = t['tc_pn_stn'].numpy().view(np.float16)
    off = 0  # in fp16 elements
    for name, n, k, fills in (('pn0b_w', 64, 64, 1), ('stn1_w', 64, 64, 1), ('stn2_w', 128, 64, 1), ('stn3_w', 256, 128, 4)):
        w = t[name].numpy().astype(np.float32).reshape(n, k)
        h = n // 2
        slot = h * 32 * (k // 16) // fills  # fp16 elements of one CTA's slot
        for f in range(fills):
            for r in (0, 1):
                got = steps(stn[off:off + slot], h, (k // 16) // fills)
                kk = k // fills
                close(got, w[h * r:h * r + h, kk * f:kk * f + kk])
                off += slot
    assert off * 2 == stn.size * 2 == 196608

    feat = t['tc_pn_feat'].numpy().view(np.float16)
    w1 = t['pn1_w'].numpy().astype(np.float32).reshape(64, 64)
    w2 = t['pn2_w'].numpy().astype(np.float32).reshape(128, 64)
    close(steps(feat[:8192], 64, 4), w1)
    close(steps(feat[8192:16384], 128, 2), w2[:, :32])
    close(steps(feat[16384:24576], 128, 2), w2[:, 32:])
    assert feat.size * 2 == 49152


def test_merged_stn_fc3_pack_equals_transform_then_conv1(weights):
    """The tensor-core path packs the STN's last FC merged with the local branch's conv1 (packing.py): the kernel chain
    f2 -> fc3' -> M_q, conv1 = M_q . a1 must equal the reference's T_q = fc3(f2) + I, x' = T_q . a1, conv1 = W1 x' (nn.py:338-351,
    162-190) -- checked here on the CPU with the matrices decoded from the pack exactly as the kernels address them."""
    from ppsurf_b200 import packing
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
    p = packing.pack_decoder(sd, 'cpu', 64, 50)
    t = p.tensors
    raw = t['tc_stn_fc'].numpy()
    head = 4 * (128 * 256 + 64 * 128)          # fc1, fc2 packs
    body = raw[head:head + 4 * 4096 * 64].view(np.float16)
    bias = raw[head + 4 * 4096 * 64:].view(np.float32)
    assert bias.size == 4096
    # 16 blocks of 256 output rows, K = 64: 4 k16 steps of [hi kb0 | hi kb1 | lo kb0 | lo kb1], blocks of [256 rows][8 fp16]
    w3m = np.zeros((4096, 64), np.float64)
    per = 256 * 32
    for nb in range(16):
        for s_ in range(4):
            st = body[(nb * 4 + s_) * per:(nb * 4 + s_ + 1) * per].astype(np.float64)
            hi, lo = st[:per // 2].reshape(2, 256, 8), st[per // 2:].reshape(2, 256, 8)
            w3m[nb * 256:(nb + 1) * 256, 16 * s_:16 * s_ + 16] = (hi + lo).transpose(1, 0, 2).reshape(256, 16)
    rng = np.random.default_rng(3)
    f2 = np.abs(rng.standard_normal(64))                      # output of a ReLU layer
    a1 = np.abs(rng.standard_normal((50, 64)))                # [points, channels]
    w3 = t['stnf3_w'].numpy().astype(np.float64)              # unmerged, the fp32 path's
    b3i = t['stnf3_b'].numpy().astype(np.float64)             # bias + identity
    w1 = t['pn1_w'].numpy().astype(np.float64)
    T = (w3 @ f2 + b3i).reshape(64, 64)
    want = (a1 @ T.T) @ w1.T                                  # x'[p,i] = sum_j T[i,j] a1[p,j];  conv1[p,o] = sum_i W1[o,i] x'[p,i]
    M = (w3m @ f2 + bias).reshape(64, 64)
    got = a1 @ M.T                                            # conv1[p,o] = sum_j M[o,j] a1[p,j]
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_latent_schedule_of_the_module():
    """PPSurfModel.latent_schedule (source/poco_model.py:207-224): subsets of gen_subsample_manifold points until every point has
    been visited gen_subsample_manifold_iter times; a cloud smaller than the subset is encoded whole, once per iteration"""
    import ppsurf_b200
    model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'x.txt', 'results', 0.05, 't', 256, 3, 10000, 17, 50,
                                    50000, 0, 0)
    n = 25000
    counts = np.zeros(n, dtype=np.int64)
    sched = list(model.latent_schedule(n, torch.Generator().manual_seed(3)))
    for ids in sched:
        assert ids.shape[0] == 10000 and int(ids.min()) >= 0 and int(ids.max()) < n
        counts[np.unique(ids.numpy())] += 1
    assert counts.min() >= 3 and len(sched) >= 8
    again = list(model.latent_schedule(n, torch.Generator().manual_seed(3)))
    assert len(again) == len(sched) and all(torch.equal(a, b) for a, b in zip(again, sched))  # a seed reproduces the schedule
    small = list(model.latent_schedule(500, torch.Generator().manual_seed(1)))
    assert len(small) == 3 and all(torch.equal(s, torch.arange(500)) for s in small)


def test_fka_tc_pack_layout():
    """operand pack of the fused FKAConv kernel (csrc/fka_tc.cu): k16 stages of [hi kb0 | hi kb1 | lo kb0 | lo kb1] per slice of
    min(cout, 256) rows, K order k = ((c/2)*4 + m/4)*8 + (m%4)*2 + c%2; hi + lo reproduces the weight to 2^-21"""
    from ppsurf_b200 import packing
    gen = torch.Generator().manual_seed(5)
    for cout, cin in ((32, 4), (64, 32), (512, 8)):
        w = torch.randn((cout, cin, 16), generator=gen, dtype=torch.float64)  # [o][c][m]
        pack = packing.fka_tc_pack(w)
        nsl, k = min(cout, 256), 16 * cin
        assert pack.numel() == 4 * k * cout
        halves = pack.view(torch.float16).view(cout // nsl, k // 16, 2, 2, nsl, 8).to(torch.float64)  # [slice][stage][hl][kb][row][8]
        both = halves[:, :, 0] + halves[:, :, 1]  # hi + lo: [slice][stage][kb][row][8]
        wk = both.permute(0, 3, 1, 2, 4).reshape(cout, k)  # [o][k]
        for c in range(cin):
            for m in (0, 5, 15):
                kk = ((c // 2) * 4 + m // 4) * 8 + (m % 4) * 2 + c % 2
                assert torch.allclose(wk[:, kk], w[:, c, m], rtol=2.0 ** -20, atol=1e-7)
