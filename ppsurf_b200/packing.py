"""Host-side weight packing: reference ``state_dict`` tensors -> the flat fp32 matrices the C ABI consumes.

Done once per ``load_state_dict`` in float64 on the CPU, then cast to fp32 and moved to the device:
  * eval-mode BatchNorm folded into the preceding linear map,
  * ``fc8 . fc_value`` and ``att.fc_value . bn3 . conv3`` merged (the attention weights sum to one, so the pooling
    commutes with every affine map that follows it),
  * FKAConv ``cv.weight [cout,cin,1,16]`` repacked to ``[cout, 16*cin]`` with column ``m*cin + c``.
Names on the left are the reference's (source/ppsurf_model.py:39-68, source/base/nn.py, source/poco_model.py:364-379).
"""
import ctypes

import torch

from . import _lib

BN_EPS = 1e-5


def _f64(sd, name):
    return sd[name].detach().to('cpu', torch.float64)


def _mat(sd, name):
    w = _f64(sd, name)
    return w.reshape(w.shape[0], -1)


def _fold_bn(sd, lin, bn, has_bias=True):
    """(W, b) of ``bn(lin(x))`` in eval mode"""
    w = _mat(sd, lin + '.weight')
    b = _f64(sd, lin + '.bias') if has_bias else torch.zeros(w.shape[0], dtype=torch.float64)
    s = _f64(sd, bn + '.weight') / torch.sqrt(_f64(sd, bn + '.running_var') + BN_EPS)
    return w * s[:, None], (b - _f64(sd, bn + '.running_mean')) * s + _f64(sd, bn + '.bias')


class Packed:
    """A ctypes struct plus the device tensors that keep its pointers alive."""

    def __init__(self, struct):
        self.struct = struct
        self.tensors = {}

    def put(self, field, value, device):
        t = value.to(torch.float32).contiguous().to(device)
        self.tensors[field] = t
        setattr(self.struct, field, t.data_ptr())
        return t

    @property
    def ref(self):
        return ctypes.byref(self.struct)


def pack_decoder(sd, device, k, num_pts_local, prefix='') -> Packed:
    p = Packed(_lib.DecoderWeights())
    g = prefix + 'projection.'
    w1 = _mat(sd, g + 'fc1.weight')
    latent = w1.shape[0]
    st = p.struct
    st.latent, st.heads, st.k, st.num_pts_local = latent, _mat(sd, g + 'fc_query.weight').shape[0], int(k), int(num_pts_local)
    p.put('w1_lat', w1[:, :latent], device)
    p.put('w1_xyz', w1[:, latent:latent + 3], device)
    p.put('b1', _f64(sd, g + 'fc1.bias'), device)
    p.put('w2', _mat(sd, g + 'fc2.weight'), device)
    p.put('b2', _f64(sd, g + 'fc2.bias'), device)
    p.put('w3', _mat(sd, g + 'fc3.weight'), device)
    p.put('b3', _f64(sd, g + 'fc3.bias'), device)
    p.put('wq', _mat(sd, g + 'fc_query.weight'), device)
    p.put('bq', _f64(sd, g + 'fc_query.bias'), device)
    w8, wv = _mat(sd, g + 'fc8.weight'), _mat(sd, g + 'fc_value.weight')
    p.put('wv8', w8 @ wv, device)
    p.put('bv8', w8 @ _f64(sd, g + 'fc_value.bias') + _f64(sd, g + 'fc8.bias'), device)

    n = prefix + 'point_net.'
    for field, lin, bn in (('pn0a', 'conv0a', 'bn0a'), ('pn0b', 'conv0b', 'bn0b'), ('stn1', 'stn2.conv1', 'stn2.bn1'),
                           ('stn2', 'stn2.conv2', 'stn2.bn2'), ('stn3', 'stn2.conv3', 'stn2.bn3'),
                           ('stnf1', 'stn2.fc1', 'stn2.bn4'), ('stnf2', 'stn2.fc2', 'stn2.bn5'),
                           ('pn1', 'conv1', 'bn1'), ('pn2', 'conv2', 'bn2')):
        w, b = _fold_bn(sd, n + lin, n + bn)
        p.put(field + '_w', w, device)
        p.put(field + '_b', b, device)
    st.stn_size = _mat(sd, n + 'stn2.conv3.weight').shape[0]
    p.put('stnf3_w', _mat(sd, n + 'stn2.fc3.weight'), device)
    p.put('stnf3_b', _f64(sd, n + 'stn2.fc3.bias') + torch.eye(64, dtype=torch.float64).reshape(-1), device)
    a3, c3 = _fold_bn(sd, n + 'conv3', n + 'bn3')  # x3 = a3 h + c3 (no ReLU before the attention pooling)
    wq = _mat(sd, n + 'att.fc_query.weight')  # [1,C]
    p.put('pnq_w', (wq @ a3).reshape(-1), device)
    st.pnq_b = float((wq @ c3).reshape(()) + _f64(sd, n + 'att.fc_query.bias').reshape(()))
    wv_att = _mat(sd, n + 'att.fc_value.weight')
    p.put('pnv_w', wv_att @ a3, device)
    p.put('pnv_b', wv_att @ c3 + _f64(sd, n + 'att.fc_value.bias'), device)

    m = prefix + 'mlp.layers.'
    for i in (0, 1):
        w, b = _fold_bn(sd, m + '{}.0'.format(i), m + '{}.1'.format(i))
        p.put('m{}_w'.format(i), w, device)
        p.put('m{}_b'.format(i), b, device)
    p.put('m2_w', _mat(sd, m + '2.0.weight'), device)
    p.put('m2_b', _f64(sd, m + '2.0.bias'), device)
    if latent == 256 and st.heads == 64:
        def pair_pack(w):  # per k16 stage: [CTA 0: features 0..127 | CTA 1: features 128..255], each hi 4 KB + lo 4 KB
            a, b = tc_pack_matrix(w[:128]).view(16, -1), tc_pack_matrix(w[128:]).view(16, -1)
            return torch.stack([a, b], dim=1).reshape(-1)

        def query_pack(w):  # fc_query as the N = 64 operand: per 64-column chunk [CTA 0: heads 0..31 | CTA 1: heads 32..63], a CTA's
            # slot = its 4 k16 steps of 2 KB (hi kb0 | hi kb1 | lo kb0 | lo kb1, 32 rows each)
            a, b = tc_pack_matrix(w[:32]).view(4, -1), tc_pack_matrix(w[32:]).view(4, -1)
            return torch.stack([a, b], dim=1).reshape(-1)

        pack = torch.cat([pair_pack(_mat(sd, g + 'fc2.weight')), pair_pack(_mat(sd, g + 'fc3.weight')),
                          query_pack(_mat(sd, g + 'fc_query.weight'))])
        assert pack.numel() == _lib.lib.pps_decoder_tc_pack_bytes()
        p.tensors['tc_wpack'] = pack.to(device)
        st.tc_wpack = p.tensors['tc_wpack'].data_ptr()
    else:
        st.tc_wpack = None
    if latent == 256 and st.stn_size == 256 and num_pts_local <= 256:
        t = p.tensors
        w3 = t['stn3_w'].to('cpu', torch.float64)
        def pair_slots(w, nslots):  # pn_stn_kernel runs as CTA pairs: per ring slot [CTA 0: its half of the weight rows | CTA 1]
            h = w.shape[0] // 2
            a, b = tc_pack_matrix(w[:h]).view(nslots, -1), tc_pack_matrix(w[h:]).view(nslots, -1)
            return torch.stack([a, b], dim=1).reshape(-1)

        # conv0b, stn.conv1 (8 KB per CTA), stn.conv2 (16 KB): one slot each; stn.conv3: 4 slots of two k16 steps (16 KB per CTA)
        stn = torch.cat([pair_slots(t['pn0b_w'].cpu(), 1), pair_slots(t['stn1_w'].cpu(), 1), pair_slots(t['stn2_w'].cpu(), 1),
                         pair_slots(w3, 4)])
        feat = torch.cat([tc_pack_matrix(t['pn1_w'].cpu()), tc_pack_matrix(t['pn2_w'].cpu())])
        assert stn.numel() == _lib.lib.pps_decoder_tc_pn_stn_bytes() and feat.numel() == _lib.lib.pps_decoder_tc_pn_feat_bytes()
        t['tc_pn_stn'], t['tc_pn_feat'] = stn.to(device), feat.to(device)
        st.tc_pn_stn, st.tc_pn_feat = t['tc_pn_stn'].data_ptr(), t['tc_pn_feat'].data_ptr()
    else:
        st.tc_pn_stn = st.tc_pn_feat = None
    if latent == 256 and st.stn_size == 256:
        t = p.tensors
        # The STN's last FC is packed MERGED with the local branch's conv1: conv1(T_q . a1) = (W1 T_q) . a1 and T_q = fc3(f2) + I is
        # linear in f2, so M_q[o,j] = sum_i W1[o,i] T_q[i,j] = sum_c (sum_i W1[o,i] W3[(i,j),c]) f2[c] + sum_i W1[o,i] (b3 + I)[(i,j)]:
        # stn_fc_tc_kernel emits M_q (same shape as T_q) and pn_feat_kernel's first MMA is conv1 with a per-query matrix -- exact in
        # real arithmetic like the reformulations of DESIGN.md section 3, merged in float64.  (The fp32 path keeps the unmerged stnf3 / pn1.)
        w1f = _fold_bn(sd, n + 'conv1', n + 'bn1')[0]                                        # [64 o, 64 i]
        w3 = _mat(sd, n + 'stn2.fc3.weight').reshape(64, 64, -1)                             # [i, j, c]
        b3i = (_f64(sd, n + 'stn2.fc3.bias') + torch.eye(64, dtype=torch.float64).reshape(-1)).reshape(64, 64)
        f3 = torch.einsum('oi,ijc->ojc', w1f, w3).reshape(4096, -1)
        b3m = (w1f @ b3i).reshape(-1)
        stn_fc = torch.cat([tc_pack_matrix(t['stnf1_w'].cpu()), tc_pack_matrix(t['stnf2_w'].cpu())] +
                           [tc_pack_matrix(f3[nb * 256:(nb + 1) * 256]) for nb in range(16)] +
                           [b3m.to(torch.float32).contiguous().view(torch.uint8)])
        mlp = torch.cat([tc_pack_matrix(t[name].cpu()) for name in ('wv8', 'pnv_w', 'm0_w', 'm1_w')])
        assert stn_fc.numel() == _lib.lib.pps_decoder_tc_stn_fc_bytes() and mlp.numel() == _lib.lib.pps_decoder_tc_mlp_bytes()
        t['tc_stn_fc'], t['tc_mlp'] = stn_fc.to(device), mlp.to(device)
        st.tc_stn_fc, st.tc_mlp = t['tc_stn_fc'].data_ptr(), t['tc_mlp'].data_ptr()
        p.put('tc_bias_feat', t['bv8'].cpu().double() + t['pnv_b'].cpu().double(), device)
    else:
        st.tc_stn_fc = st.tc_mlp = st.tc_bias_feat = None
    return p


def tc_split(w: torch.Tensor):
    """fp32 value -> (fp16 hi, fp16 lo) with hi + lo carrying 22 mantissa bits"""
    w32 = w.to(torch.float32)
    hi = w32.to(torch.float16)
    lo = (w32 - hi.to(torch.float32)).to(torch.float16)
    return hi, lo


def tc_pack_matrix(w: torch.Tensor) -> torch.Tensor:
    """``w [N,256]`` -> uint8 tensor: 16 k16 stages of [hi block 0 | hi block 1 | lo block 0 | lo block 1], block = [N rows, 8
    fp16] (the UMMA canonical K-major no-swizzle layout the kernel copies verbatim into shared memory)"""
    n, k = w.shape
    assert k % 16 == 0
    hi, lo = tc_split(w)
    both = torch.stack([hi, lo], dim=0).view(2, n, k // 16, 2, 8)  # [hl, n, stage, kb, 8]
    return both.permute(2, 0, 3, 1, 4).contiguous().view(torch.uint8).reshape(-1)


def fka_tc_pack(w_cm: torch.Tensor) -> torch.Tensor:
    """``w_cm [cout, cin, 16]`` (c, m) -> operand pack of the fused kernel (csrc/fka_tc.cu): K order
    k = ((c / 2) * 4 + m / 4) * 8 + (m % 4) * 2 + c % 2, k16 stages of [hi kb0 | hi kb1 | lo kb0 | lo kb1] per slice of
    min(cout, 256) output rows (slices back to back)"""
    cout, cin, _ = w_cm.shape
    assert cin % 4 == 0
    # [cout, cpair, cc, s, mm] -> [cout, cpair, s, mm, cc]
    wk = w_cm.reshape(cout, cin // 2, 2, 4, 4).permute(0, 1, 3, 4, 2).reshape(cout, 16 * cin)
    nsl = min(cout, 256)
    return torch.cat([tc_pack_matrix(wk[i:i + nsl]) for i in range(0, cout, nsl)])


def pack_fkaconv(sd, name, device, act, bn=None) -> Packed:
    """``name`` = FKAConvLayer prefix; ``bn`` = the BatchNorm that follows it (folded, with its ReLU).  An input width that is
    not a multiple of 4 (cv0: 3 channels) is zero-padded to the next multiple: the caller pads ``x`` the same way."""
    p = Packed(_lib.FKAConvWeights())
    cv = _f64(sd, name + '.cv.weight')  # [cout,cin,1,16]
    if cv.shape[1] % 4:
        cv = torch.cat([cv, torch.zeros(cv.shape[0], 4 - cv.shape[1] % 4, 1, 16, dtype=cv.dtype)], dim=1)
    cout, cin = cv.shape[0], cv.shape[1]
    w = cv[:, :, 0, :].permute(0, 2, 1).reshape(cout, 16 * cin)
    st = p.struct
    st.cin, st.cout, st.act = cin, cout, {'relu': 0, 'silu': 1}[act]
    st.alpha = float(_f64(sd, name + '.alpha'))
    st.beta = float(_f64(sd, name + '.beta'))
    st.norm_radius = float(_f64(sd, name + '.norm_radius'))
    if bn is not None:
        s = _f64(sd, bn + '.weight') / torch.sqrt(_f64(sd, bn + '.running_var') + BN_EPS)
        w = w * s[:, None]
        p.put('out_bias', _f64(sd, bn + '.bias') - _f64(sd, bn + '.running_mean') * s, device)
        st.out_relu = 1
    else:
        st.out_bias = None
        st.out_relu = 0
    p.put('cv_w', w, device)
    nsl = min(cout, 256)
    if nsl in (32, 64, 128, 256) and cout % nsl == 0:
        # the fp16 hi/lo split needs lo = w - fp16(w) ~ 2^-12 |w| in fp16's NORMAL range (>= 6.1e-5): the kernels of the wide layers
        # are ~0.01, whose lo parts would be denormals with 2 % precision.  A power-of-two scale (exact) moves max|w| to ~2048; the
        # fused kernel's epilogue multiplies the accumulator by its inverse.
        wmax = float(w.abs().max())
        shift = int(torch.floor(torch.log2(torch.tensor(2048.0 / wmax)))) if wmax > 0 else 0
        shift = max(-20, min(20, shift))
        p.tensors['tc_pack'] = fka_tc_pack((w * 2.0 ** shift).reshape(cout, 16, cin).permute(0, 2, 1).contiguous()).to(device)
        st.tc_pack = p.tensors['tc_pack'].data_ptr()
        st.tc_out_scale = 2.0 ** -shift
    else:
        st.tc_pack = None
        st.tc_out_scale = 1.0
    p.put('fc1', _mat(sd, name + '.fc1.weight'), device)
    p.put('fc2', _mat(sd, name + '.fc2.weight'), device)
    p.put('fc3', _mat(sd, name + '.fc3.weight'), device)
    host = torch.cat([_mat(sd, name + '.fc' + i + '.weight').reshape(-1) for i in '123']).to(torch.float32).contiguous()
    p.tensors['mlp_host'] = host  # stays on the host: passed by value to the fused kernels
    st.mlp_host = host.data_ptr()
    p.put('in1_w', _f64(sd, name + '.bn1.weight'), device)
    p.put('in1_b', _f64(sd, name + '.bn1.bias'), device)
    p.put('in2_w', _f64(sd, name + '.bn2.weight'), device)
    p.put('in2_b', _f64(sd, name + '.bn2.bias'), device)
    return p


class PackedLinear:
    def __init__(self, w, b, device):
        self.w = w.to(torch.float32).contiguous().to(device)
        self.b = None if b is None else b.to(torch.float32).contiguous().to(device)


RESBLOCKS = ('resnetb01', 'resnetb10', 'resnetb11', 'resnetb20', 'resnetb21', 'resnetb30', 'resnetb31', 'resnetb40',
             'resnetb41')


def pack_encoder(sd, device, act='silu', prefix='encoder.') -> dict:
    """All encoder layers (source/base/nn.py:453-554).  Concat inputs of the U-Net decoder are split into one matrix
    per source so that ``cat`` never materialises."""
    e = prefix
    out = {'cv0': pack_fkaconv(sd, e + 'cv0', device, act, bn=e + 'bn0')}
    for rb in RESBLOCKS:
        r = e + rb
        blk = {'cv0': PackedLinear(*_fold_bn(sd, r + '.cv0', r + '.bn0'), device),
               'cv1': pack_fkaconv(sd, r + '.cv1', device, act, bn=r + '.bn1'),
               'cv2': PackedLinear(*_fold_bn(sd, r + '.cv2', r + '.bn2'), device)}
        if (r + '.shortcut.weight') in sd:
            blk['shortcut'] = PackedLinear(*_fold_bn(sd, r + '.shortcut', r + '.bn_shortcut'), device)
        out[rb] = blk
    for cv, bn in (('cv5', 'bn5'), ('cv3d', 'bn3d'), ('cv2d', 'bn2d'), ('cv1d', 'bn1d'), ('cv0d', 'bn0d')):
        # input = cat([first, second]): cv5 sees [x4, global max]; cvXd sees [interpolate(deeper), skip] where the skip
        # level has as many channels as the layer's output, so the first block is (in - out) columns wide
        w, b = _fold_bn(sd, e + cv, e + bn)
        c_first = w.shape[1] - w.shape[0]
        out[cv] = (PackedLinear(w[:, :c_first], None, device), PackedLinear(w[:, c_first:], b, device))
    out['fcout'] = PackedLinear(_mat(sd, e + 'fcout.weight'), _f64(sd, e + 'fcout.bias'), device)
    return out
