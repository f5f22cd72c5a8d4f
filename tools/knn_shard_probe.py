import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import ppsurf_b200, bench
from ppsurf_b200 import ops, synthetic
dev = torch.device('cuda:0')
pts_np = synthetic.synthetic_cloud(100000, 42)
pts = torch.from_numpy(pts_np).to(dev)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts_np, 129, 1)
r = 131
total = r ** 3
qry = ops.grid_queries(r, step, bmin_pad, device=dev)
index = ops.KnnIndex(pts)
def t(q, name):
    index.query(q, 64); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); index.query(q, 64); e1.record(); torch.cuda.synchronize()
    print('{:40s} {:8d} queries {:7.2f} ms {:6.1f} ns/query'.format(name, q.shape[0], e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e6 / q.shape[0]))
t(qry, 'full grid')
for blk in (4736, 37888, 1184):
    spans = bench.grid_blocks(total, 8, 0, blk)
    q = torch.cat([qry[f:f + c] for f, c in spans])
    t(q, 'dealt blocks of {} (rank 0 of 8)'.format(blk))
n8 = total // 8
for k in (0, 3, 4):
    t(qry[k * n8:(k + 1) * n8].contiguous(), 'contiguous slab {} of 8'.format(k))
t(qry[:606208].contiguous(), 'first 606k (one super-chunk)')
