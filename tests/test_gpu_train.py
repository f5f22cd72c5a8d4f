"""Training step (BASELINE config 5) on the CUDA path: every ``*_fwd`` / ``*_bwd`` primitive of the C ABI against torch autograd of
the same op, and the whole ``training_step`` (loss, logits, all 298 parameter gradients, updated buffers) against the golden vectors
of the UNMODIFIED reference (``tests/golden/train_step.npz``) and the float64 oracle twin.  Run on the B200 box: pytest -m gpu"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'the gpu tests need a CUDA device'
    from ppsurf_b200 import ops
    ops.require_device()
    return torch.device('cuda:0')


def _rand(gen, *shape, dev=None):
    return torch.randn(*shape, generator=gen).to(dev)


def _close(a, b, tol, what=''):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = float((a - b).abs().max())
    scale = max(float(b.abs().max()), 1e-30)
    assert err <= tol * max(scale, 1.0), '{}: max abs err {:.3e} (scale {:.3e})'.format(what, err, scale)


def _grad_pair(ours_fn, ref_fn, inputs, tol, what):
    """forward + backward of both callables on the same leaves with the same upstream gradient"""
    a = [t.detach().clone().requires_grad_(t.is_floating_point()) for t in inputs]
    b = [t.detach().clone().requires_grad_(t.is_floating_point()) for t in inputs]
    ya, yb = ours_fn(*a), ref_fn(*b)
    _close(ya, yb, tol, what + ' forward')
    up = torch.randn(yb.shape, generator=torch.Generator().manual_seed(5)).to(yb.device)
    ya.backward(up)
    yb.backward(up)
    for i, (ta, tb) in enumerate(zip(a, b)):
        if tb.grad is not None:
            assert ta.grad is not None, '{}: no gradient for input {}'.format(what, i)
            _close(ta.grad, tb.grad, tol, '{} grad {}'.format(what, i))


# ---- primitives ----------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('m,n,k', [(1000, 256, 259), (77, 16, 3), (5000, 64, 256), (300, 4096, 64), (64, 2, 256), (5000, 16, 32),
                                   (5001, 16, 3), (40000, 32, 16)])
def test_linear_fwd_bwd_fp32(dev, m, n, k):
    from ppsurf_b200 import autograd as ag
    ag.set_precision('fp32')
    gen = torch.Generator().manual_seed(m + n + k)
    x, w, b = _rand(gen, m, k, dev=dev), _rand(gen, n, k, dev=dev) / k ** 0.5, _rand(gen, n, dev=dev)
    _grad_pair(lambda x, w, b: ag.linear(x, w, b), lambda x, w, b: F.linear(x, w, b), [x, w, b], 2e-5, 'linear {}x{}x{}'.format(m, n, k))
    # a column slice of a wider parameter (fc1 of the projection: latent and xyz columns) -- strided weight views
    wide = _rand(gen, n, k + 5, dev=dev)
    _grad_pair(lambda x, w: ag.linear(x, w[:, 2:2 + k], None), lambda x, w: F.linear(x, w[:, 2:2 + k]), [x, wide], 1e-4, 'sliced weight')


@pytest.mark.parametrize('m,n,k', [(4096, 256, 256), (1000, 64, 256), (130, 256, 512), (5000, 64, 64), (2048, 4096, 64)])
def test_linear_fwd_bwd_bf16_tensor_cores(dev, m, n, k):
    """bf16 tcgen05 path against torch with bf16-rounded operands and fp32 accumulation (same arithmetic contract); shapes whose three
    GEMMs (forward, data gradient, weight gradient) all qualify for the tensor-core kernel (m >= 64, n >= 16, k >= 32)"""
    from ppsurf_b200 import autograd as ag
    gen = torch.Generator().manual_seed(m + n + k)
    x, w, b = _rand(gen, m, k, dev=dev), _rand(gen, n, k, dev=dev) / k ** 0.5, _rand(gen, n, dev=dev)
    up = _rand(gen, m, n, dev=dev)

    def r(t):
        return t.bfloat16().float()

    ag.set_precision('bf16')
    try:
        xa, wa, ba = (t.clone().requires_grad_(True) for t in (x, w, b))
        y = ag.linear(xa, wa, ba)
        y.backward(up)
    finally:
        ag.set_precision('fp32')
    _close(y, r(x) @ r(w).t() + b, 1e-5, 'bf16 forward')
    _close(xa.grad, r(up) @ r(w), 1e-5, 'bf16 dgrad')
    _close(wa.grad, r(up).t() @ r(x), 2e-5, 'bf16 wgrad (split over the rows)')
    _close(ba.grad, up.sum(0), 1e-5, 'bias grad')


def test_bmm_fwd_bwd(dev):
    from ppsurf_b200 import autograd as ag
    ag.set_precision('fp32')
    gen = torch.Generator().manual_seed(1)
    h, t = _rand(gen, 37, 50, 64, dev=dev), _rand(gen, 37, 64, 64, dev=dev)
    _grad_pair(lambda h, t: ag.Bmm.apply(h, t.transpose(1, 2)), lambda h, t: torch.bmm(h, t.transpose(1, 2)), [h, t], 2e-5, 'bmm')


@pytest.mark.parametrize('groups,rows,c,act', [(1, 5000, 64, 'relu'), (3, 1600, 16, 'silu'), (1, 33, 1024, None), (2, 48, 16, 'silu')])
def test_norm_fwd_bwd(dev, groups, rows, c, act):
    from ppsurf_b200 import autograd as ag
    gen = torch.Generator().manual_seed(rows)
    x = _rand(gen, groups, rows, c, dev=dev) * 2 + 0.5
    gamma, beta = _rand(gen, c, dev=dev) * 0.2 + 1, _rand(gen, c, dev=dev) * 0.1
    fn = {'relu': F.relu, 'silu': F.silu, None: lambda v: v}[act]

    def ref(x, gamma, beta):
        m = x.mean(dim=1, keepdim=True)
        v = x.var(dim=1, unbiased=False, keepdim=True)
        return fn((x - m) / torch.sqrt(v + 1e-5) * gamma + beta)

    _grad_pair(lambda x, g, b: ag.norm(x, g, b, act), ref, [x, gamma, beta], 5e-5, 'norm')
    if groups == 1:  # BatchNorm running statistics like torch's module
        bn = torch.nn.BatchNorm1d(c).to(dev).train()
        rm, rv = bn.running_mean.clone() + 0.3, bn.running_var.clone() * 1.7
        bn.running_mean.copy_(rm)
        bn.running_var.copy_(rv)
        bn(x[0])
        ag.norm(x, gamma, beta, act, 1e-5, (rm, rv), 0.1)
        _close(rm, bn.running_mean, 1e-5, 'running_mean')
        _close(rv, bn.running_var, 1e-5, 'running_var')


def test_segment_ops_fwd_bwd(dev):
    from ppsurf_b200 import autograd as ag
    gen = torch.Generator().manual_seed(2)
    x, w = _rand(gen, 500, 16, 16, dev=dev), torch.rand(500, 16, generator=gen).to(dev) + 0.1
    _grad_pair(lambda x, w: ag.SegMax.apply(x, w), lambda x, w: (x * w[:, :, None]).max(dim=1)[0], [x, w], 1e-5, 'weighted seg max')
    _grad_pair(lambda x: ag.SegMax.apply(x, None), lambda x: x.max(dim=1)[0], [x], 1e-6, 'seg max')
    v = _rand(gen, 500, 16, dev=dev)
    _grad_pair(lambda x, v: ag.ConcatBcast.apply(x, v), lambda x, v: torch.cat([x, v[:, None, :].expand(-1, 16, -1)], dim=2), [x, v], 1e-6,
               'concat')
    x2, w2 = _rand(gen, 3000, 16, dev=dev), _rand(gen, 3000, dev=dev)
    _grad_pair(lambda x, w: ag.RowScale.apply(x, w), lambda x, w: x * w[:, None], [x2, w2], 1e-5, 'rowscale')
    for name, fn in (('relu', F.relu), ('silu', F.silu)):
        _grad_pair(lambda x: ag.act(x, name), fn, [x2], 1e-5, name)
    idx = torch.randint(0, 3000, (7000,), generator=gen).to(dev)
    _grad_pair(lambda x: ag.GatherRows.apply(x, idx.int()), lambda x: x[idx], [x2], 1e-5, 'gather rows')
    xb = _rand(gen, 2, 600, 24, dev=dev)
    ids = torch.randint(0, 600, (2, 150, 16), generator=gen).to(dev)
    _grad_pair(lambda x: ag.GatherMax.apply(x, ids.int()),
               lambda x: torch.stack([x[s][ids[s]].max(dim=1)[0] for s in range(2)]), [xb], 1e-6, 'gather max')


@pytest.mark.parametrize('s,h,c', [(64, 64, 256), (50, 1, 256), (200, 1, 64)])
def test_attention_pooling_fwd_bwd(dev, s, h, c):
    from ppsurf_b200 import autograd as ag
    gen = torch.Generator().manual_seed(s + h)
    scores, v = _rand(gen, 77, s, h, dev=dev) * 2, _rand(gen, 77, s, c, dev=dev)
    _grad_pair(lambda a, b: ag.AttnPool.apply(a, b), lambda a, b: (torch.softmax(a, dim=1).mean(dim=2)[:, :, None] * b).sum(dim=1),
               [scores, v], 2e-5, 'attention pooling')


def test_fkaconv_geometry_and_feature_product(dev, oracle):
    from ppsurf_b200 import autograd as ag
    gen = torch.Generator().manual_seed(3)
    b, n_in, n_s, kn, cin = 2, 400, 100, 16, 24
    pts = _rand(gen, b, n_in, 3, dev=dev) * 0.2
    sup = pts[:, :n_s].contiguous()
    ids = torch.randint(0, n_in, (b, n_s, kn), generator=gen).to(dev)
    alpha, beta = torch.tensor([0.8], device=dev), torch.tensor([1.2], device=dev)
    radius = torch.tensor([0.15], device=dev)

    def ours(alpha, beta):
        offs, dw = ag.FkaGeometry.apply(alpha, beta, pts, sup, ids.int(), radius.clone(), False, 0.1)
        return torch.cat([offs.reshape(-1), dw])

    def ref(alpha, beta):
        pg = torch.stack([pts[s][ids[s]] for s in range(b)]) - sup[:, :, None, :]
        dist = torch.sqrt((pg ** 2).sum(-1))
        dw = torch.sigmoid(-alpha * dist + beta)
        dws = dw.sum(2, keepdim=True)
        dw = dw / (dws + (dws == 0) + 1e-6) * kn
        return torch.cat([(pg / radius).reshape(-1), dw.reshape(-1)])

    _grad_pair(ours, ref, [alpha, beta], 2e-5, 'fka geometry')
    # train mode moves norm_radius towards the mean neighbourhood radius (nn.py:608-613)
    r2 = radius.clone()
    ag.FkaGeometry.apply(alpha, beta, pts, sup, ids.int(), r2, True, 0.1)
    pg = torch.stack([pts[s][ids[s]] for s in range(b)]) - sup[:, :, None, :]
    want = radius * 0.9 + torch.sqrt((pg ** 2).sum(-1)).max(2)[0].mean() * 0.1
    _close(r2, want, 1e-6, 'norm_radius update')
    x, mat = _rand(gen, b, n_in, cin, dev=dev), _rand(gen, b * n_s * kn, 16, dev=dev)

    def ref_feat(x, mat):
        xg = torch.stack([x[s][ids[s]] for s in range(b)])  # [B,Ns,K,Cin]
        return torch.einsum('bnjc,bnjm->bncm', xg, mat.view(b, n_s, kn, 16)).reshape(b * n_s, cin * 16)

    _grad_pair(lambda x, mat: ag.FkaFeat.apply(x, mat, ids.int()), ref_feat, [x, mat], 2e-5, 'fka feature product')


def test_cross_entropy_and_dropout(dev):
    from ppsurf_b200 import autograd as ag
    gen = torch.Generator().manual_seed(4)
    logits = _rand(gen, 999, 2, dev=dev) * 3
    target = torch.randint(0, 2, (999,), generator=gen).to(dev)
    _grad_pair(lambda l: ag.cross_entropy(l, target)[0], lambda l: F.cross_entropy(l, target), [logits], 1e-6, 'cross entropy')
    x = torch.ones(200000, device=dev, requires_grad=True)
    ag.manual_seed(7)
    y = ag.Dropout.apply(x, 0.3)
    kept = (y > 0).float().mean().item()
    assert abs(kept - 0.7) < 0.01 and abs(y.max().item() - 1 / 0.7) < 1e-6
    y.sum().backward()
    assert torch.equal(x.grad > 0, y > 0)
    y2 = ag.Dropout.apply(x, 0.3)
    assert not torch.equal(y2 > 0, y > 0), 'every draw uses a fresh mask'


def test_training_abi_rejects_bad_arguments(dev):
    """the training entry points return a status and a message instead of launching on inconsistent arguments"""
    import ctypes
    from ppsurf_b200 import _lib, autograd as ag
    lib = _lib.lib
    x = torch.zeros(64, 32, device=dev)
    y = torch.zeros(64, 16, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # precision outside {0, 1}; ldc smaller than n; null operand
    assert lib.pps_gemm(p(x), 0, 32, 1, p(x), 0, 1, 32, p(y), 0, 16, 1, 64, 16, 32, None, 0, 7, st) == -1
    assert b'precision' in lib.pps_last_error()
    assert lib.pps_gemm(p(x), 0, 32, 1, p(x), 0, 1, 32, p(y), 0, 8, 1, 64, 16, 32, None, 0, 0, st) == -1
    assert lib.pps_gemm(None, 0, 32, 1, p(x), 0, 1, 32, p(y), 0, 16, 1, 64, 16, 32, None, 0, 0, st) == -1
    # norm: workspace too small -> PPS_ERR_WORKSPACE; zero groups -> invalid
    ws = torch.zeros(8, dtype=torch.uint8, device=dev)
    g = torch.ones(32, device=dev)
    m = torch.zeros(32, device=dev)
    assert lib.pps_norm_fwd(p(x), 1, 64, 32, p(g), p(g), 1e-5, 0, p(x), p(m), p(m), p(ws), ws.numel(), st) == -2
    assert lib.pps_norm_fwd(p(x), 0, 64, 32, p(g), p(g), 1e-5, 0, p(x), p(m), p(m), p(ws), ws.numel(), st) == -1
    # attention pooling: a group that does not fit the shared-memory tile
    assert lib.pps_attn_pool_fwd(p(x), p(x), 1, 4096, 64, 8, p(x), p(x), p(x), st) == -1
    assert b'shared-memory' in lib.pps_last_error()
    # dropout probability out of range, more than 16 neighbours per FKAConv neighbourhood
    assert lib.pps_dropout_fwd(p(x), 10, 1.5, 1, None, p(x), p(ws), st) == -1
    assert lib.pps_fka_feat_fwd(p(x), p(x), p(x), 1, 4, 4, 17, 8, p(x), st) == -1
    # the Python layer turns a status into an exception
    with pytest.raises(_lib.PpsError):
        ag.gemm(x, torch.zeros(32, 16, device=dev), prec=5)
    torch.cuda.synchronize()


# ---- the whole step against the reference ----------------------------------------------------------------------------------------------

def _fixture_batch(g, dev):
    data = {}
    for key in g:
        if key.startswith('in_'):
            t = torch.from_numpy(g[key])
            data[key[3:]] = (t.long() if t.dtype == torch.int32 else t).to(dev)
    return data


def _train_net(dev, weights, dropout):
    import ppsurf_b200
    net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    net = net.to(dev).train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = dropout
    return net


def _oracle_grads64(weights, g):
    """float64 gradients of the oracle twin on the fixture batch (CPU, a few seconds)"""
    from oracle import ppsurf_train_oracle as T
    s = T.State(weights, dtype=torch.float64)
    data = {}
    for key in g:
        if key.startswith('in_'):
            t = torch.from_numpy(g[key])
            data[key[3:]] = t.long() if t.dtype == torch.int32 else t.double()
    loss, logits = T.training_step(s, data, dropout=0.0)
    return float(loss), s.grads()


def test_training_step_matches_reference(dev, weights, weights_digest):
    """loss / logits / buffers against the reference's golden vectors; every parameter gradient against the float64 twin with a
    tolerance tied to the error of the REFERENCE's own float32 gradients (stored in the fixture)"""
    from ppsurf_b200 import autograd as ag
    g = load_golden('train_step')
    assert str(g['digest']) == weights_digest
    ag.set_precision('fp32')
    net = _train_net(dev, weights, dropout=0.0)
    data = _fixture_batch(g, dev)
    launches0 = __import__('ppsurf_b200')._lib.lib.pps_launch_count()
    pred = net.forward(data)
    loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(-1, 2), data['occ'].reshape(-1))
    loss.backward()
    torch.cuda.synchronize()
    assert __import__('ppsurf_b200')._lib.lib.pps_launch_count() - launches0 > 500, 'the step must run on the library kernels'
    assert abs(float(loss.detach()) - float(g['loss64'])) < 2e-5
    assert np.abs(pred.detach().cpu().numpy() - g['logits64']).max() < 1e-4
    _loss64, ref = _oracle_grads64(weights, g)
    names = [str(n) for n in g['grad_names']]
    params = dict(net.named_parameters())
    report = []
    for i, name in enumerate(names):
        got = params[name].grad
        assert got is not None, 'no gradient for ' + name
        got = got.detach().double().cpu().reshape(-1)
        want = ref[name].reshape(-1)
        norm = float(want.norm())
        err = float((got - want).norm())
        assert abs(norm - float(g['grad_norm'][i])) <= 1e-6 * max(norm, 1e-12) + 1e-14, 'twin vs reference fixture: ' + name
        # tolerance: 3e-3 of the tensor's norm, or 4x what the reference's OWN float32 gradient is off by (its median error is 1.2e-3
        # of the norm on this batch: the gradients of this deep BatchNorm network are ill-conditioned), plus an absolute floor of a few
        # float32 roundings of the largest gradients (2.2) for the structurally-zero gradients (biases in front of a BatchNorm) and the
        # scalar alpha / beta, which are sums of 10^4..10^5 cancelling terms
        tol = max(3e-3 * norm, 4.0 * float(g['grad_ref32_l2err'][i]), 2e-7 * np.sqrt(want.numel()), 2e-6 if want.numel() == 1 else 0.0)
        report.append((err / tol, name, err, tol, norm))
    report.sort(reverse=True)
    print('largest gradient error / tolerance:', ['{} {:.2f} (err {:.2e}, norm {:.2e})'.format(r[1], r[0], r[2], r[4]) for r in report[:6]])
    rel = [r[2] / r[4] for r in report if r[4] > 1e-6]
    print('relative L2 error of the {} non-zero gradients: median {:.2e}, max {:.2e}'.format(len(rel), np.median(rel), max(rel)))
    assert report[0][0] <= 1.0, report[:8]
    assert np.median(rel) < 3e-3
    bufs = dict(net.named_buffers())
    for key in g:
        if key.startswith('buf_'):
            _close(bufs[key[4:]], torch.from_numpy(g[key]), 2e-5, key)


def test_training_step_bf16_and_optimizer(dev, weights):
    """the bf16 tensor-core step agrees with the fp32 step to bf16 accuracy, and a few AdamW steps through PPSurfModel.training_step
    lower the loss"""
    import ppsurf_b200
    from ppsurf_b200 import autograd as ag
    g = load_golden('train_step')
    data = _fixture_batch(g, dev)
    grads = {}
    for mode in ('fp32', 'bf16'):
        ag.set_precision(mode)
        try:
            net = _train_net(dev, weights, dropout=0.0)
            pred = net.forward(dict(data))
            loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(-1, 2), data['occ'].reshape(-1))
            loss.backward()
            grads[mode] = (float(loss.detach()), {k: v.grad.detach().flatten() for k, v in net.named_parameters()})
        finally:
            ag.set_precision('fp32')
    assert abs(grads['fp32'][0] - grads['bf16'][0]) < 2e-2
    # how far bf16 moves the gradients of THIS network is a property of the network (ill-conditioned: see the fp32 test), so the
    # yardstick is the oracle twin under torch.autocast(bfloat16) on the same device: per parameter group, the cosine between this
    # repo's bf16 and fp32 gradients must not be worse than the twin's by more than 0.03
    from oracle import ppsurf_train_oracle as T
    twin = {}
    for mode in ('fp32', 'bf16'):
        s = T.State(weights, device=dev)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(mode == 'bf16')):
            logits = T.forward(s, data, True, 0.0)
        T.loss_of(logits.float(), data['occ']).backward()
        twin[mode] = {k: v.flatten() for k, v in s.grads().items()}
    big = [k for k, v in grads['fp32'][1].items() if float(v.norm()) > 1e-3 and v.numel() >= 256]
    for group in ('encoder', 'projection', 'point_net', 'mlp'):
        keys = [k for k in big if k.startswith(group)]
        ours = np.median([float(F.cosine_similarity(grads['fp32'][1][k], grads['bf16'][1][k], dim=0)) for k in keys])
        ref = np.median([float(F.cosine_similarity(twin['fp32'][k], twin['bf16'][k], dim=0)) for k in keys])
        exact = min(float(F.cosine_similarity(grads['fp32'][1][k], twin['fp32'][k], dim=0)) for k in keys)
        print('{}: bf16-vs-fp32 gradient cosine, this repo {:.4f}, torch autocast twin {:.4f}; fp32 vs twin fp32 min {:.6f}'.format(
            group, ours, ref, exact))
        assert ours >= ref - 0.03 and ours > 0.75 and exact > 0.9999, group

    model = ppsurf_b200.PPSurfModel(256, ['occ'], 3, 2, 64, 0.0, False, 'x.txt', 'results', 0.05, 'ppsurf', 256, 10, 10000, 17, 50, 50000, 10, 0)
    model.network.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    model = model.to(dev).train()
    opt = model.configure_optimizers()['optimizer']
    ag.manual_seed(11)
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            loss = model.training_step(dict(data), 0)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    # after training steps the eval path must see the NEW parameters (packed predict weights are rebuilt)
    model.eval()
    with torch.no_grad():
        out = model.network.forward(dict(data))
    assert out.shape == (2, 2, data['occ'].shape[1]) and torch.isfinite(out).all()


def test_training_step_small_ragged_cloud_vs_twin(dev, weights, oracle):
    """one cloud of 2304 points and 33 queries: the deepest encoder level holds 9 points, so its neighbourhoods have 9 members (the
    kn < 16 paths of the FKAConv primitives, BatchNorm over 9 rows; smaller levels make BatchNorm itself ill-conditioned: two rows
    normalise to +-1 whatever their values); compared with the float64 twin computed here"""
    sys.path.insert(0, GOLDEN)
    from make_golden_train import make_batch
    from oracle import ppsurf_train_oracle as T
    from ppsurf_b200 import autograd as ag
    batch = make_batch(seed=321, b=1, n=2304, q=33)
    assert batch['ids44'].shape[1:] == (9, 9)
    data = {k: torch.from_numpy(v).to(dev) for k, v in batch.items()}
    ag.set_precision('fp32')
    net = _train_net(dev, weights, dropout=0.0)
    pred = net.forward(dict(data))
    loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(-1, 2), data['occ'].reshape(-1))
    loss.backward()
    s = T.State(weights, dtype=torch.float64)
    ref_loss, ref_logits = T.training_step(s, {k: (torch.from_numpy(v).double() if v.dtype == np.float32 else torch.from_numpy(v))
                                               for k, v in batch.items()}, dropout=0.0)
    # BatchNorm over 2 and 10 rows amplifies float32 rounding: looser than the 1200-point fixture
    assert abs(float(loss.detach()) - float(ref_loss)) < 1e-3
    assert float((pred.detach().double().cpu() - ref_logits).abs().max()) < 5e-3
    ref = s.grads()
    rel = []
    for name, par in net.named_parameters():
        want = ref[name].reshape(-1)
        if float(want.norm()) > 1e-4:
            rel.append((float((par.grad.detach().double().cpu().reshape(-1) - want).norm() / want.norm()), name))
    rel.sort(reverse=True)
    print('small cloud: median relative gradient error {:.2e}, worst {}'.format(np.median([r[0] for r in rel]), rel[:3]))
    assert np.median([r[0] for r in rel]) < 5e-3 and sum(r[0] > 0.1 for r in rel) <= 3, rel[:6]


def _ddp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from oracle import ppsurf_oracle as O
    from ppsurf_b200 import autograd as ag
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    dev = torch.device('cuda', rank)
    weights = O.make_state_dict(42)
    g = dict(np.load(os.path.join(GOLDEN, 'train_step.npz')))
    full = _fixture_batch(g, dev)
    mine = {k: v[rank:rank + 1].contiguous() for k, v in full.items()}  # one cloud per rank
    ag.set_precision('fp32')
    # local gradient without DDP
    net = _train_net(dev, weights, 0.0)
    pred = net.forward(dict(mine))
    loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(-1, 2), mine['occ'].reshape(-1))
    loss.backward()
    local = torch.cat([p.grad.flatten() for p in net.parameters()])
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = torch.stack(gathered).mean(0)
    # the same step under DDP: gradients are all-reduced (mean) by DDP's hooks while the backward kernels run
    net2 = DDP(_train_net(dev, weights, 0.0), device_ids=[rank])
    pred = net2(dict(mine))
    loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(-1, 2), mine['occ'].reshape(-1))
    loss.backward()
    got = torch.cat([p.grad.flatten() for p in net2.module.parameters()])
    err = float((got - want).norm() / want.norm())
    torch.save({'err': err, 'loss': float(loss.detach())}, os.path.join(tmp, 'rank{}.pt'.format(rank)))
    dist.destroy_process_group()


def test_ddp_training_step_two_ranks(dev, tmp_path):
    """data-parallel fit: per-rank batch shards, NCCL gradient all-reduce by torch DDP around the CUDA backward"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    mp.spawn(_ddp_worker, args=(2, 29533, str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(os.path.join(str(tmp_path), 'rank{}.pt'.format(r))) for r in range(2)]
    # two executions of the same step differ by the gradient noise of this network (atomic summation order, ~1e-3 of the norm, see
    # test_training_step_matches_reference); gradients that were NOT averaged over the ranks would be off by ~1
    print('DDP gradient vs mean of the local gradients, relative L2 error per rank:', [r['err'] for r in res])
    assert all(r['err'] < 2e-2 for r in res), res
