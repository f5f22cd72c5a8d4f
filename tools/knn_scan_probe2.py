"""kNN of the dense 131^3 grid: sweep of (finest cells per point, run length, scan-per-child threshold) -- all result-neutral knobs."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import _lib, ops, synthetic

dev = torch.device('cuda:0')
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
pts_np = synthetic.synthetic_cloud(npts, 42)
pts = torch.from_numpy(pts_np).to(dev)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts_np, 129, 1)
qry = ops.grid_queries(131, step, bmin_pad, device=dev)
ref = None
for cells, run, sc in itertools.product((1, 2, 4, 8), (8, 16, 32), (128, 192, 320)):
    _lib.lib.pps_debug_knn_cells(cells)
    _lib.lib.pps_debug_knn_run(run)
    _lib.lib.pps_debug_knn_scan_child(sc)
    index = ops.KnnIndex(pts)
    index.query(qry[:200000], k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx, d2 = index.query(qry, k, return_dist=True)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = (idx.clone(), d2.clone())
    same = bool((idx == ref[0]).all()) and bool((d2 == ref[1]).all())
    print('cells {} run {:2d} scan_child {:3d}: {:.1f} ms  identical {}'.format(cells, run, sc, e0.elapsed_time(e1), same))
