"""kNN (k=64) over the dense bench grid for several run lengths (consecutive queries per warp in the seeded search)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import _lib, ops, synthetic

dev = torch.device('cuda:0')
pts_np = synthetic.synthetic_cloud(100000, 42)
pts = torch.from_numpy(pts_np).to(dev)
index = ops.KnnIndex(pts)
small = torch.from_numpy(synthetic.synthetic_cloud(10000, 7)).to(dev)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts_np, 129, 1)
qry = ops.grid_queries(131, step, bmin_pad, device=dev)
part = qry[:281011].contiguous()  # one rank's share at 8 GPUs is dealt in blocks; a contiguous eighth is the harder (slab) case
ref = None
for run in (8, 16, 32, 64):
    _lib.lib.pps_debug_knn_run(run)
    for name, q in (('full grid', qry), ('first eighth', part)):
        index.query(q, 64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            idx = index.query(q, 64)
        e1.record()
        torch.cuda.synchronize()
        if name == 'full grid':
            if ref is None:
                ref = idx.clone()
            same = bool(torch.equal(ref, idx))
        print('run {:3d}  {:12s} {:8.2f} ms  ({:.1f} ns/query){}'.format(run, name, e0.elapsed_time(e1) / 3, e0.elapsed_time(e1) / 3 * 1e6 / q.shape[0],
                                                                  '  identical to run 8: {}'.format(same) if name == 'full grid' else ''), flush=True)
_lib.lib.pps_debug_knn_run(16)
for factor in (16, 8, 4, 2, 1):
    _lib.lib.pps_debug_knn_cells(factor)
    index = ops.KnnIndex(pts)
    small_index = ops.KnnIndex(small)
    for name, ix, q, k in (('grid k=64', index, qry, 64), ('encoder 10k self k=16', small_index, small, 16)):
        ix.query(q, k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            idx = ix.query(q, k)
        e1.record()
        torch.cuda.synchronize()
        same = bool(torch.equal(ref, idx)) if k == 64 else ''
        print('cells/point {:2d}  {:22s} {:8.3f} ms  {}'.format(factor, name, e0.elapsed_time(e1) / 3, same), flush=True)
_lib.lib.pps_debug_knn_cells(2)
