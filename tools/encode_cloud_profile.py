"""Where the time of PPSurfModel.encode_cloud goes on the 100k-point bench cloud (debug aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import synthetic

dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(
    pointnet_latent_size=256, output_names=['imp_surf_sign'], in_channels=3, out_channels=2, k=64, lambda_l1=0.0, debug=False,
    in_file='bench', results_dir='results', padding_factor=0.05, name='ppsurf_50nn', network_latent_size=256,
    gen_subsample_manifold_iter=10, gen_subsample_manifold=10000, gen_resolution_global=129, num_pts_local=50,
    rec_batch_size=50000, gen_refine_iter=10, workers=8)
net = model.network
net.load_state_dict(synthetic.make_state_dict(net, 42))
model = model.to(dev)
NPTS = 20000 if '--small' in sys.argv else 100000
pts = torch.from_numpy(synthetic.synthetic_cloud(NPTS, 42).T[None].copy()).to(dev)
acc = {'spatial_ids': 0.0, 'encode': 0.0}
orig_ids, orig_enc = net.spatial_ids, net.encode


def wrap(name, fn):
    def inner(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        acc[name] += time.perf_counter() - t0
        return out
    return inner


net.spatial_ids = wrap('spatial_ids', orig_ids)
net.encode = wrap('encode', orig_enc)
for bp in ((16,) if '--small' in sys.argv else (16, 16)):
    for k in acc:
        acc[k] = 0.0
    net.sampling_seed = 42
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sched = list(model.latent_schedule(NPTS, torch.Generator().manual_seed(42)))
    t_sched = time.perf_counter() - t0
    t0 = time.perf_counter()
    model.encode_cloud(pts, generator=torch.Generator().manual_seed(42), batch_passes=bp)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print('batch_passes {}: {} passes, total {:.3f} s (schedule alone {:.3f} s): spatial_ids {:.3f} s, encode {:.3f} s, rest {:.3f} s'.format(
        bp, len(sched), total, t_sched, acc['spatial_ids'], acc['encode'], total - acc['spatial_ids'] - acc['encode']))
