"""FKAConv layer alone at batch 64 (unique bytes > L2): timing per layer shape, fused kernel vs unfused kernels, and the
statistics passes separately.  Also the ncu target for `-k regex:fka_fused`:  python tools/fka_bench.py [--one]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppsurf_b200
from ppsurf_b200 import _lib, ops, synthetic

dev = torch.device('cuda:0')
net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
net.load_state_dict(synthetic.make_state_dict(net, 42))
net = net.to(dev)
enc = net.packed()['encoder']
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


shapes = [('resnetb01.cv1', 10000, 10000, 64)]
if '--one' not in sys.argv:
    shapes += [('resnetb10.cv1', 10000, 2500, 64), ('resnetb11.cv1', 2500, 2500, 64), ('resnetb21.cv1', 625, 625, 64),
               ('resnetb31.cv1', 156, 156, 64), ('resnetb41.cv1', 39, 39, 64)]
for name, n_in, n_s, b in shapes:
    blk, layer = name.split('.')
    w = enc[blk][layer]
    cin, cout = w.struct.cin, w.struct.cout
    p = torch.from_numpy(np.stack([synthetic.synthetic_cloud(n_in, 10 + i) for i in range(4)])).to(dev).repeat(b // 4, 1, 1).contiguous()
    sup = p[:, :n_s].contiguous()
    ids = torch.stack([ops.knn(p[i].contiguous(), sup[i].contiguous(), 16) for i in range(4)]).repeat(b // 4, 1, 1).contiguous()
    x = torch.randn((b, n_in, cin), device=dev)
    t_fused = timed(lambda: ops.fkaconv(w, x, p, sup, ids))
    unique = b * (n_in * (cin + 3) * 4 + n_s * 16 * 4 + n_s * 12 + n_s * cout * 4) + cin * cout * 64
    line = 'FKAConv {} ({}->{} @ {} -> {}, B={}): fused {:.3f} ms, unique {:.1f} MB -> {:.0f} GB/s = {:.3f} of measured HBM peak'.format(
        name, cin, cout, n_in, n_s, b, t_fused, unique / 1e6, unique / t_fused / 1e6, unique / t_fused / 1e6 / peaks['hbm_gbs'])
    if '--one' not in sys.argv:
        _lib.lib.pps_debug_fka_fused(0)
        t_unfused = timed(lambda: ops.fkaconv(w, x, p, sup, ids))
        _lib.lib.pps_debug_fka_fused(1)
        line += '; unfused {:.3f} ms'.format(t_unfused)
    print(line, flush=True)
