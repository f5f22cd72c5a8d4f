"""kNN timing on the bench workload: 100k-point cloud, dense 131^3 grid, k=64 (debug aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import ops, synthetic

dev = torch.device('cuda:0')
pts_np = synthetic.synthetic_cloud(100000, 42)
pts = torch.from_numpy(pts_np).to(dev)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts_np, 129, 1)
qry = ops.grid_queries(131, step, bmin_pad, device=dev)
index = ops.KnnIndex(pts)
for k in (64, 64, 16):
    index.query(qry[:100000], k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx, d2 = index.query(qry, k, return_dist=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('k={} full grid {} queries: {:.1f} ms = {:.1f} ns/query'.format(k, qry.shape[0], ms, ms * 1e6 / qry.shape[0]))
    # by distance to the surface: slab near the sphere centre vs slab near the surface
    for name, first in (('centre slab', 131 * 131 * 65), ('surface slab', 131 * 131 * 12), ('outside slab', 131 * 131 * 2)):
        q = qry[first:first + 131 * 131 * 2]
        index.query(q, k)
        torch.cuda.synchronize()
        e0.record()
        index.query(q, k)
        e1.record()
        torch.cuda.synchronize()
        print('   {:13s} {:.2f} ms = {:.1f} ns/query'.format(name, e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e6 / q.shape[0]))
# build time
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    ops.KnnIndex(pts)
e1.record()
torch.cuda.synchronize()
print('index build: {:.3f} ms'.format(e0.elapsed_time(e1) / 10))
