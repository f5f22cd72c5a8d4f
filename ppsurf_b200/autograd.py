"""``torch.autograd.Function`` pairs over the training primitives of the C ABI (``include/ppsurf_b200.h``, "config 5").

torch contributes the tape (which backward to call, in which order, and the accumulation of ``.grad``), device memory and the
current stream; every forward and every backward below is one or a few calls into ``libppsurf_b200.so``.  With these Functions the
reference's ``training_step`` (source/poco_model.py:120-125) runs under Lightning's automatic optimisation, AMP and DDP: the
parameters are ordinary ``nn.Parameter`` leaves, so DDP's gradient hooks and any torch optimiser work unchanged.

Precision of the dense contractions: ``set_precision('fp32' | 'bf16')``; inside ``torch.autocast(dtype=torch.bfloat16)`` the bf16
tensor-core path is used as well.  Activations, weights and gradients are fp32 in memory in both modes.
"""
import ctypes

import torch

from ._lib import check, lib
from .ops import _ptr, _stream

_state = {'precision': 0, 'seed': 0x1234567}
ACT = {None: 0, 'none': 0, 'relu': 1, 'silu': 2}


def set_precision(name: str):
    """'fp32' (SIMT, the path of the gradient parity tests) or 'bf16' (tcgen05, fp32 accumulation)"""
    _state['precision'] = {'fp32': 0, 'bf16': 1}[name]


def precision() -> int:
    if torch.is_autocast_enabled():
        return 1
    return _state['precision']


def manual_seed(seed: int):
    """seed of the dropout masks (advanced by every draw)"""
    _state['seed'] = int(seed) & 0xFFFFFFFF


def _draw_counter(device):
    key = ('counter', str(device))
    if key not in _state:
        _state[key] = torch.zeros((1,), dtype=torch.int32, device=device)
    return _state[key]


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def gemm(a: torch.Tensor, b: torch.Tensor, bias=None, out=None, accumulate=False, prec=None) -> torch.Tensor:
    """``a [m,k]`` (or ``[batch,m,k]``) times ``b [k,n]`` (or ``[batch,k,n]``): arbitrary-stride VIEWS, no copies.  Returns a
    contiguous ``[m,n]`` / ``[batch,m,n]`` fp32 tensor (``out`` when given)."""
    batched = a.dim() == 3
    if batched:
        batch, m, k = a.shape
        sa = a.stride()
        sb = b.stride()
        n = b.shape[2]
        assert b.shape[0] == batch and b.shape[1] == k
    else:
        batch = 1
        m, k = a.shape
        n = b.shape[1]
        assert b.shape[0] == k, (a.shape, b.shape)
        sa = (0,) + tuple(a.stride())
        sb = (0,) + tuple(b.stride())
    if out is None:
        out = torch.empty((batch, m, n) if batched else (m, n), dtype=torch.float32, device=a.device)
    assert out.is_contiguous() and a.dtype == torch.float32 and b.dtype == torch.float32 and a.is_cuda and b.is_cuda
    check(lib.pps_gemm(ctypes.c_void_p(a.data_ptr()), sa[0], sa[1], sa[2], ctypes.c_void_p(b.data_ptr()), sb[0], sb[1], sb[2],
                       _ptr(out), m * n, n, batch, m, n, k, _ptr(bias, torch.float32) if bias is not None else None,
                       1 if accumulate else 0, precision() if prec is None else prec, _stream()))
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty((x.shape[1],), dtype=torch.float32, device=x.device)
    check(lib.pps_colsum(ctypes.c_void_p(x.data_ptr()), x.shape[0], x.shape[1], x.stride(0), _ptr(out), 0, _stream()))
    return out


class Linear(torch.autograd.Function):
    """``y = x @ w.T + bias``: ``x [M,K]`` (row-strided views allowed), ``w [N,K]`` (any view of a parameter), ``bias [N]`` or None.
    Backward: ``dx = dy @ w``, ``dw = dy.T @ x`` (split over the rows), ``db = colsum(dy)`` -- three calls of the same GEMM."""

    @staticmethod
    def forward(ctx, x, w, bias):
        x, w = _f32(x), _f32(w)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return gemm(x, w.t(), bias=None if bias is None else _f32(bias).contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32(dy).contiguous()
        dx = gemm(dy, w) if ctx.needs_input_grad[0] else None
        dw = gemm(dy.t(), x) if ctx.needs_input_grad[1] else None
        db = colsum(dy) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


class Bmm(torch.autograd.Function):
    """``out[b] = a[b] @ bm[b]`` for views ``a [B,M,K]``, ``bm [B,K,N]`` (torch.bmm(trans2, x) of the PointNet, nn.py:347)"""

    @staticmethod
    def forward(ctx, a, bm):
        ctx.save_for_backward(a, bm)
        return gemm(a, bm)

    @staticmethod
    def backward(ctx, dout):
        a, bm = ctx.saved_tensors
        dout = dout.contiguous()
        da = gemm(dout, bm.transpose(1, 2)) if ctx.needs_input_grad[0] else None
        db = gemm(a.transpose(1, 2), dout) if ctx.needs_input_grad[1] else None
        return da, db


class Norm(torch.autograd.Function):
    """BatchNorm (groups = 1) / InstanceNorm (groups = samples) over ``x [G,R,C]`` with the activation fused; in train mode the
    BatchNorm running statistics are updated in place like torch's module does."""

    @staticmethod
    def forward(ctx, x, gamma, beta, act, eps, running_mean, running_var, momentum):
        x = x.contiguous()
        g, r, c = x.shape
        gamma, beta = _f32(gamma).contiguous(), _f32(beta).contiguous()
        y = torch.empty_like(x)
        mean = torch.empty((g, c), dtype=torch.float32, device=x.device)
        var = torch.empty_like(mean)
        ws = torch.empty(lib.pps_norm_workspace_bytes(g, c), dtype=torch.uint8, device=x.device)
        check(lib.pps_norm_fwd(_ptr(x, torch.float32), g, r, c, _ptr(gamma), _ptr(beta), eps, act, _ptr(y), _ptr(mean), _ptr(var), _ptr(ws),
                               ws.numel(), _stream()))
        if running_mean is not None:
            check(lib.pps_bn_running_update(_ptr(mean), _ptr(var), r, momentum, c, _ptr(running_mean, torch.float32),
                                            _ptr(running_var, torch.float32), _stream()))
        ctx.save_for_backward(x, gamma, beta, mean, var)
        ctx.act, ctx.eps = act, eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, var = ctx.saved_tensors
        g, r, c = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dgamma = torch.empty((c,), dtype=torch.float32, device=x.device)
        dbeta = torch.empty_like(dgamma)
        ws = torch.empty(lib.pps_norm_workspace_bytes(g, c), dtype=torch.uint8, device=x.device)
        check(lib.pps_norm_bwd(_ptr(x), _ptr(dy, torch.float32), g, r, c, _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(var), ctx.eps, ctx.act,
                               _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(ws), ws.numel(), _stream()))
        return dx, dgamma, dbeta, None, None, None, None, None


class Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.pps_act_fwd(_ptr(x, torch.float32), x.numel(), act, _ptr(y), _stream()))
        ctx.save_for_backward(x)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        check(lib.pps_act_bwd(_ptr(x), _ptr(dy, torch.float32), x.numel(), ctx.act, _ptr(dx), _stream()))
        return dx, None


class Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        x = x.contiguous()
        y = torch.empty_like(x)
        mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        _state['seed'] = (_state['seed'] * 1664525 + 1013904223) & 0xFFFFFFFF
        counter = _draw_counter(x.device)
        check(lib.pps_dropout_fwd(_ptr(x, torch.float32), x.numel(), p, _state['seed'], _ptr(counter), _ptr(y), _ptr(mask), _stream()))
        counter.add_(1)  # a device op: inside a captured graph every replay advances the draw
        ctx.save_for_backward(mask)
        ctx.p = p
        return y

    @staticmethod
    def backward(ctx, dy):
        mask, = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        check(lib.pps_dropout_bwd(_ptr(dy, torch.float32), _ptr(mask), dy.numel(), ctx.p, _ptr(dx), _stream()))
        return dx, None


class RowScale(torch.autograd.Function):
    """``y[r,:] = x[r,:] * w[r]``"""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.contiguous(), w.contiguous()
        y = torch.empty_like(x)
        check(lib.pps_rowscale_fwd(_ptr(x, torch.float32), _ptr(w, torch.float32), x.shape[0], x.shape[1], _ptr(y), _stream()))
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        check(lib.pps_rowscale_bwd(_ptr(x), _ptr(w), _ptr(dy, torch.float32), x.shape[0], x.shape[1], _ptr(dx), _ptr(dw), _stream()))
        return dx, dw


class ConcatBcast(torch.autograd.Function):
    """``x [G,S,C]``, ``v [G,C]`` -> ``[G,S,2C]`` = cat(x, v broadcast over S)"""

    @staticmethod
    def forward(ctx, x, v):
        x, v = x.contiguous(), v.contiguous()
        g, s, c = x.shape
        out = torch.empty((g, s, 2 * c), dtype=torch.float32, device=x.device)
        check(lib.pps_concat_bcast_fwd(_ptr(x, torch.float32), _ptr(v, torch.float32), g, s, c, _ptr(out), _stream()))
        ctx.shape = (g, s, c)
        return out

    @staticmethod
    def backward(ctx, dout):
        g, s, c = ctx.shape
        dout = dout.contiguous()
        dx = torch.empty((g, s, c), dtype=torch.float32, device=dout.device)
        dv = torch.empty((g, c), dtype=torch.float32, device=dout.device)
        check(lib.pps_concat_bcast_bwd(_ptr(dout, torch.float32), g, s, c, _ptr(dx), _ptr(dv), _stream()))
        return dx, dv


class GatherRows(torch.autograd.Function):
    """``x [N,C]``, ``idx [M]`` int32 -> ``x[idx] [M,C]``; backward = atomic scatter-add"""

    @staticmethod
    def forward(ctx, x, idx):
        x = x.contiguous()
        y = torch.empty((idx.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
        check(lib.pps_gather_rows(_ptr(x, torch.float32), _ptr(idx, torch.int32), idx.shape[0], x.shape[1], _ptr(y), _stream()))
        ctx.save_for_backward(idx)
        ctx.n = x.shape[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        idx, = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.zeros((ctx.n, dy.shape[1]), dtype=torch.float32, device=dy.device)
        check(lib.pps_scatter_add_rows(_ptr(dy, torch.float32), _ptr(idx), idx.shape[0], dy.shape[1], _ptr(dx), _stream()))
        return dx, None


class SegMax(torch.autograd.Function):
    """``x [G,S,C]``, optional ``w [G,S]`` -> ``max_s x*w  [G,C]``"""

    @staticmethod
    def forward(ctx, x, w):
        x = x.contiguous()
        w = None if w is None else w.contiguous()
        g, s, c = x.shape
        y = torch.empty((g, c), dtype=torch.float32, device=x.device)
        arg = torch.empty((g, c), dtype=torch.int32, device=x.device)
        check(lib.pps_seg_max_fwd(_ptr(x, torch.float32), _ptr(w, torch.float32) if w is not None else None, g, s, c, _ptr(y), _ptr(arg),
                                  _stream()))
        ctx.save_for_backward(x, arg, *(() if w is None else (w,)))
        ctx.has_w = w is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, arg = ctx.saved_tensors[:2]
        w = ctx.saved_tensors[2] if ctx.has_w else None
        g, s, c = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        want_dw = ctx.has_w and ctx.needs_input_grad[1]
        dw = torch.empty((g, s), dtype=torch.float32, device=x.device) if want_dw else None
        check(lib.pps_seg_max_bwd(_ptr(dy, torch.float32), _ptr(arg), _ptr(x), _ptr(w) if w is not None else None, g, s, c, _ptr(dx),
                                  _ptr(dw) if dw is not None else None, _stream()))
        return dx, dw


class GatherMax(torch.autograd.Function):
    """max_pool (nn.py:677-680): ``x [B,Nin,C]``, ``ids [B,Ns,K]`` int32 -> ``[B,Ns,C]``"""

    @staticmethod
    def forward(ctx, x, ids):
        x = x.contiguous()
        b, n_in, c = x.shape
        n_s, kn = ids.shape[1], ids.shape[2]
        y = torch.empty((b, n_s, c), dtype=torch.float32, device=x.device)
        arg = torch.empty((b, n_s, c), dtype=torch.int32, device=x.device)
        check(lib.pps_gather_max_fwd(_ptr(x, torch.float32), _ptr(ids, torch.int32), b, n_in, n_s, c, kn, _ptr(y), _ptr(arg), _stream()))
        ctx.save_for_backward(arg)
        ctx.shape = (b, n_in, c)
        return y

    @staticmethod
    def backward(ctx, dy):
        arg, = ctx.saved_tensors
        b, n_in, c = ctx.shape
        dy = dy.contiguous()
        dx = torch.zeros((b, n_in, c), dtype=torch.float32, device=dy.device)
        check(lib.pps_gather_max_bwd(_ptr(dy, torch.float32), _ptr(arg), arg.numel() // c, c, _ptr(dx), _stream()))
        return dx, None


class AttnPool(torch.autograd.Function):
    """``scores [G,S,H]``, ``v [G,S,C]`` -> ``sum_s mean_h softmax_s(scores) * v  [G,C]``"""

    @staticmethod
    def forward(ctx, scores, v):
        scores, v = scores.contiguous(), v.contiguous()
        g, s, h = scores.shape
        c = v.shape[2]
        prob = torch.empty_like(scores)
        a = torch.empty((g, s), dtype=torch.float32, device=v.device)
        out = torch.empty((g, c), dtype=torch.float32, device=v.device)
        check(lib.pps_attn_pool_fwd(_ptr(scores, torch.float32), _ptr(v, torch.float32), g, s, h, c, _ptr(prob), _ptr(a), _ptr(out), _stream()))
        ctx.save_for_backward(prob, a, v)
        return out

    @staticmethod
    def backward(ctx, dout):
        prob, a, v = ctx.saved_tensors
        g, s, h = prob.shape
        c = v.shape[2]
        dout = dout.contiguous()
        dscores, dv = torch.empty_like(prob), torch.empty_like(v)
        check(lib.pps_attn_pool_bwd(_ptr(dout, torch.float32), _ptr(prob), _ptr(a), _ptr(v), g, s, h, c, _ptr(dscores), _ptr(dv), _stream()))
        return dscores, dv


class FkaGeometry(torch.autograd.Function):
    """neighbourhood offsets and distance weights of FKAConvLayer (nn.py:598-624).  Inputs that carry a gradient: ``alpha``, ``beta``.
    ``norm_radius`` (buffer) is updated in place in train mode.  Returns ``offs [R,3]`` (no gradient: the reference detaches nothing here
    but its points never require one) and ``dw [R]``."""

    @staticmethod
    def forward(ctx, alpha, beta, pts, support, ids, norm_radius, training, momentum):
        b, n_in, _ = pts.shape
        n_s, kn = ids.shape[1], ids.shape[2]
        r = b * n_s * kn
        dev = pts.device
        offs = torch.empty((r, 3), dtype=torch.float32, device=dev)
        dist, sig, dw = (torch.empty((r,), dtype=torch.float32, device=dev) for _ in range(3))
        scratch = torch.empty((1,), dtype=torch.float64, device=dev)
        alpha, beta = _f32(alpha).contiguous(), _f32(beta).contiguous()
        check(lib.pps_fka_geometry_fwd(_ptr(pts, torch.float32), _ptr(support, torch.float32), _ptr(ids, torch.int32), b, n_in, n_s, kn,
                                       _ptr(alpha), _ptr(beta), _ptr(norm_radius, torch.float32), momentum, 1 if training else 0, _ptr(offs),
                                       _ptr(dist), _ptr(sig), _ptr(dw), _ptr(scratch), _stream()))
        ctx.save_for_backward(dist, sig)
        ctx.kn = kn
        ctx.mark_non_differentiable(offs)
        return offs, dw

    @staticmethod
    def backward(ctx, _doffs, ddw):
        dist, sig = ctx.saved_tensors
        ddw = ddw.contiguous()
        out = torch.empty((2,), dtype=torch.float64, device=ddw.device)
        check(lib.pps_fka_weights_bwd(_ptr(ddw, torch.float32), _ptr(sig), _ptr(dist), dist.numel() // ctx.kn, ctx.kn, _ptr(out), _stream()))
        g = out.float()
        return g[0:1], g[1:2], None, None, None, None, None, None


class FkaFeat(torch.autograd.Function):
    """``feat[p, c*16+m] = sum_j x[ids[p,j], c] * mat[p,j,m]``: ``x [B,Nin,Cin]``, ``mat [B*Ns*K,16]`` -> ``[B*Ns, Cin*16]``"""

    @staticmethod
    def forward(ctx, x, mat, ids):
        x, mat = x.contiguous(), mat.contiguous()
        b, n_in, cin = x.shape
        n_s, kn = ids.shape[1], ids.shape[2]
        feat = torch.empty((b * n_s, cin * 16), dtype=torch.float32, device=x.device)
        check(lib.pps_fka_feat_fwd(_ptr(x, torch.float32), _ptr(ids, torch.int32), _ptr(mat, torch.float32), b, n_in, n_s, kn, cin, _ptr(feat),
                                   _stream()))
        ctx.save_for_backward(x, mat, ids)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        x, mat, ids = ctx.saved_tensors
        b, n_in, cin = x.shape
        n_s, kn = ids.shape[1], ids.shape[2]
        dfeat = dfeat.contiguous()
        dx, dmat = torch.empty_like(x), torch.empty_like(mat)
        check(lib.pps_fka_feat_bwd(_ptr(dfeat, torch.float32), _ptr(x), _ptr(ids), _ptr(mat), b, n_in, n_s, kn, cin, _ptr(dx), _ptr(dmat),
                                   _stream()))
        return dx, dmat, None


class CrossEntropy(torch.autograd.Function):
    """mean over the rows of ``-log softmax(logits)[target]``: ``logits [M,C]``, ``target [M]`` int64 (compute_loss, poco_model.py:75-88)"""

    @staticmethod
    def forward(ctx, logits, target):
        logits = logits.contiguous()
        m, c = logits.shape
        rows = torch.empty((m,), dtype=torch.float32, device=logits.device)
        total = torch.empty((1,), dtype=torch.float64, device=logits.device)
        target = target.contiguous()
        check(lib.pps_ce_fwd(_ptr(logits, torch.float32), _ptr(target, torch.int64), m, c, _ptr(rows), _ptr(total), _stream()))
        ctx.save_for_backward(logits, target)
        ctx.mark_non_differentiable(rows)
        return (total / m).float().reshape(()), rows

    @staticmethod
    def backward(ctx, dloss, _drows):
        logits, target = ctx.saved_tensors
        m, c = logits.shape
        scale = (dloss.reshape(1).float() / m).expand(m).contiguous()
        dlogits = torch.empty_like(logits)
        check(lib.pps_ce_bwd(_ptr(logits), _ptr(target), _ptr(scale), m, c, _ptr(dlogits), _stream()))
        return dlogits, None


# ---- functional front ends ---------------------------------------------------------------------------------------------------------

def linear(x, w, bias=None):
    return Linear.apply(x, w, bias)


def norm(x3, gamma, beta, act=None, eps=1e-5, running=None, momentum=0.1):
    """``x3 [G,R,C]``; ``running = (running_mean, running_var)`` buffers to update (BatchNorm in train mode) or None"""
    rm, rv = running if running is not None else (None, None)
    return Norm.apply(x3, gamma, beta, ACT[act], eps, rm, rv, momentum)


def act(x, name):
    return Act.apply(x, ACT[name])


def cross_entropy(logits_rows, target_rows):
    """-> (mean loss, per-row losses)"""
    return CrossEntropy.apply(logits_rows, target_rows)
