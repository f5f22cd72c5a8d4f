"""Test helper: evaluates the PACKED formulation (ppsurf_b200.packing output: folded BatchNorm, hoisted fc1 table,
pool-before-value merges, repacked FKAConv kernels) with plain float64 torch ops on the CPU.  It lets the CPU suite
prove that the algebraic restructuring the CUDA kernels implement is equal to the reference formulation (the oracle)
before any GPU time is spent.  Not product code."""
import torch


def t64(p, name):
    return p.tensors[name].to(torch.float64)


def decode_packed(p, pts, latents, queries, idx, patches):
    """pts [N,3], latents [N,C], queries [Q,3], idx [Q,k] long, patches [Q,P,3] -> (feat_proj, feat_pn, logits)"""
    pts, latents, queries, patches = (x.to(torch.float64) for x in (pts, latents, queries, patches))
    table = latents @ t64(p, 'w1_lat').T + t64(p, 'b1') - pts @ t64(p, 'w1_xyz').T
    h = torch.relu(table[idx] + (queries @ t64(p, 'w1_xyz').T)[:, None, :])
    h = torch.relu(h @ t64(p, 'w2').T + t64(p, 'b2'))
    h = torch.relu(h @ t64(p, 'w3').T + t64(p, 'b3'))
    score = h @ t64(p, 'wq').T + t64(p, 'bq')  # [Q,k,heads]
    att = torch.softmax(score, dim=1).mean(dim=2)  # [Q,k]
    pooled = (att[:, :, None] * h).sum(dim=1)
    feat_proj = pooled @ t64(p, 'wv8').T + t64(p, 'bv8')

    a0 = torch.relu(patches @ t64(p, 'pn0a_w').T + t64(p, 'pn0a_b'))
    a1 = torch.relu(a0 @ t64(p, 'pn0b_w').T + t64(p, 'pn0b_b'))
    s = torch.relu(a1 @ t64(p, 'stn1_w').T + t64(p, 'stn1_b'))
    s = torch.relu(s @ t64(p, 'stn2_w').T + t64(p, 'stn2_b'))
    s = torch.relu(s @ t64(p, 'stn3_w').T + t64(p, 'stn3_b'))
    g = s.max(dim=1).values
    g = torch.relu(g @ t64(p, 'stnf1_w').T + t64(p, 'stnf1_b'))
    g = torch.relu(g @ t64(p, 'stnf2_w').T + t64(p, 'stnf2_b'))
    tm = (g @ t64(p, 'stnf3_w').T + t64(p, 'stnf3_b')).view(-1, 64, 64)
    x = torch.einsum('qij,qpj->qpi', tm, a1)
    x = torch.relu(x @ t64(p, 'pn1_w').T + t64(p, 'pn1_b'))
    x = torch.relu(x @ t64(p, 'pn2_w').T + t64(p, 'pn2_b'))
    logit = x @ t64(p, 'pnq_w') + p.struct.pnq_b
    w = torch.softmax(logit, dim=1)
    pooled128 = (w[:, :, None] * x).sum(dim=1)
    feat_pn = pooled128 @ t64(p, 'pnv_w').T + t64(p, 'pnv_b')

    f = feat_proj + feat_pn
    f = torch.relu(f @ t64(p, 'm0_w').T + t64(p, 'm0_b'))
    f = torch.relu(f @ t64(p, 'm1_w').T + t64(p, 'm1_b'))
    logits = f @ t64(p, 'm2_w').T + t64(p, 'm2_b')
    return feat_proj, feat_pn, logits
