"""kNN run length (queries per warp) against launch size on the bench grid: full grid, one rank's dealt blocks at 2 / 4 / 8 ranks."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200, bench
from ppsurf_b200 import ops, synthetic, _lib
dev = torch.device('cuda:0')
pts_np = synthetic.synthetic_cloud(100000, 42)
pts = torch.from_numpy(pts_np).to(dev)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts_np, 129, 1)
r = 131; total = r ** 3
qry = ops.grid_queries(r, step, bmin_pad, device=dev)
index = ops.KnnIndex(pts)
def t(q):
    index.query(q, 64); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); index.query(q, 64); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
sets = [('full grid', qry)]
for world in (2, 4, 8):
    spans = bench.grid_blocks(total, world, 0)
    sets.append(('rank 0 of %d' % world, torch.cat([qry[f:f + c] for f, c in spans])))
for name, q in sets:
    row = []
    for run in (16, 8, 4, 2, 1):
        _lib.lib.pps_debug_knn_run(run)
        row.append('run %2d: %6.2f ms' % (run, t(q)))
    print('%-14s %8d queries  ' % (name, q.shape[0]) + '  '.join(row))
_lib.lib.pps_debug_knn_run(16)
