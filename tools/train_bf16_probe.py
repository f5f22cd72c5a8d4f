"""bf16 training step: gradient agreement with the fp32 step, for this repo's kernels and for the torch twin under autocast
(how much of the deviation is inherent to bf16 on this network)."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ppsurf_b200
from ppsurf_b200 import autograd as ag
from oracle import ppsurf_oracle as O, ppsurf_train_oracle as T
from test_gpu_train import _fixture_batch, _train_net
dev = torch.device('cuda:0')
w = O.make_state_dict(42)
g = dict(np.load(os.path.join(ROOT, 'tests/golden/train_step.npz')))
data = _fixture_batch(g, dev)
res = {}
for mode in ('fp32', 'bf16'):
    ag.set_precision(mode)
    net = _train_net(dev, w, 0.0)
    pred = net.forward(dict(data))
    loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(-1, 2), data['occ'].reshape(-1))
    loss.backward()
    res[mode] = (float(loss), {k: v.grad.detach().flatten().double() for k, v in net.named_parameters()})
ag.set_precision('fp32')
tw = {}
for mode in ('fp32', 'bf16'):
    s = T.State(w, device=dev)
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(mode == 'bf16')):
        logits = T.forward(s, data, True, 0.0)
        loss = T.loss_of(logits.float(), data['occ'])
    loss.backward()
    tw[mode] = (float(loss), {k: v.flatten().double() for k, v in s.grads().items()})
print('loss ours fp32/bf16', res['fp32'][0], res['bf16'][0], 'twin fp32/bf16', tw['fp32'][0], tw['bf16'][0])
rows = []
for k, v in res['fp32'][1].items():
    if float(v.norm()) < 1e-3 or v.numel() < 256: continue
    c1 = float(F.cosine_similarity(v, res['bf16'][1][k], dim=0))
    c2 = float(F.cosine_similarity(tw['fp32'][1][k], tw['bf16'][1][k], dim=0))
    c3 = float(F.cosine_similarity(v, tw['fp32'][1][k], dim=0))
    rows.append((k, c1, c2, c3))
for grp in ('encoder', 'projection', 'point_net', 'mlp'):
    sel = [r for r in rows if r[0].startswith(grp)]
    print(grp, len(sel), 'ours bf16 vs fp32: median %.4f min %.4f | twin autocast vs fp32: median %.4f min %.4f | ours fp32 vs twin fp32 min %.6f' % (
        np.median([r[1] for r in sel]), min(r[1] for r in sel), np.median([r[2] for r in sel]), min(r[2] for r in sel), min(r[3] for r in sel)))
