"""Encoder-side measurements: time of one 10k-point encoder pass split into support sampling, the 13 kNN
index tensors and the network, and the FKAConv layer at batch 64 (unique bytes exceed the 126 MB L2) as achieved GB/s of
its unique HBM bytes  B*[N_in*(C_in+3)*4 + N_s*16*4 + N_s*12 + N_s*C_out*4] + C_in*C_out*64."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import ops, synthetic

dev = torch.device('cuda:0')
net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
net.load_state_dict(synthetic.make_state_dict(net, 42))
net = net.to(dev)
net.sampling_seed = 1
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists('MEASURED_PEAKS.json') else {'hbm_gbs': 6650.0}


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


pts = torch.from_numpy(synthetic.synthetic_cloud(10000, 3).T[None].copy()).to(dev)
t_ids, data = timed(lambda: net.spatial_ids(pts))
data['pts'] = pts
t_net, _ = timed(lambda: net.encode(data))
print('one 10k-point pass: supports + 13 kNN index tensors {:.2f} ms, network {:.2f} ms'.format(t_ids * 1e3, t_net * 1e3))

enc = net.packed()['encoder']
rng = np.random.default_rng(0)
for name, cin, n_in, n_s in (('resnetb01.cv1', 32, 10000, 10000), ('resnetb10.cv1', 32, 10000, 2500), ('resnetb21.cv1', 128, 625, 625)):
    b = 64
    blk, layer = name.split('.')
    w = enc[blk][layer]
    cout = w.struct.cout
    p = torch.from_numpy(np.stack([synthetic.synthetic_cloud(n_in, 10 + i) for i in range(4)])).to(dev).repeat(b // 4, 1, 1).contiguous()
    sup = p[:, :n_s].contiguous()
    ids = torch.stack([ops.knn(p[i].contiguous(), sup[i].contiguous(), 16) for i in range(4)]).repeat(b // 4, 1, 1).contiguous()
    x = torch.randn((b, n_in, cin), device=dev)
    t, _ = timed(lambda: ops.fkaconv(w, x, p, sup, ids))
    unique = b * (n_in * (cin + 3) * 4 + n_s * 16 * 4 + n_s * 12 + n_s * cout * 4) + cin * cout * 64
    print('FKAConv {} ({}->{} @ {} -> {}, B={}): {:.3f} ms, unique bytes {:.1f} MB -> {:.0f} GB/s = {:.3f} of measured HBM peak'.format(
        name, cin, cout, n_in, n_s, b, t * 1e3, unique / 1e6, unique / t / 1e9, unique / t / 1e9 / peaks['hbm_gbs']))
