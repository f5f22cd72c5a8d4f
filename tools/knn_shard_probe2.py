"""One rank's dealt share of the 131^3 grid at 8 / 4 GPUs (the launches that use the deferral pass) for several scan caps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ppsurf_b200
from ppsurf_b200 import _lib, ops, synthetic

dev = torch.device('cuda:0')
pts_np = synthetic.synthetic_cloud(100000, 42)
pts = torch.from_numpy(pts_np).to(dev)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts_np, 129, 1)
total = 131 ** 3
qry = ops.grid_queries(131, step, bmin_pad, device=dev)
index = ops.KnnIndex(pts)
shares = {g: torch.cat([qry[f:f + c] for f, c in bench.grid_blocks(total, g, 0, 4736)]) for g in (8, 4)}
for cap in (16384, 32768, 65536, 131072):
    _lib.lib.pps_debug_knn_scan_cap(cap)
    line = 'scan cap {:6d}:'.format(cap)
    for g, q in shares.items():
        index.query(q, 64)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        index.query(q, 64)
        e1.record()
        torch.cuda.synchronize()
        line += '  rank 0 of {}: {} queries {:.2f} ms'.format(g, q.shape[0], e0.elapsed_time(e1))
    print(line)
