"""kNN of the dense 131^3 grid (k = 64) for several values of the scan-whole-node threshold (pps_debug_knn_scan_child); checks that the
results do not depend on it."""
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
index = ops.KnnIndex(pts)
ref = None
for v in (0, 32, 64, 96, 128, 192, 256, 512):
    _lib.lib.pps_debug_knn_scan_child(v)
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
    q = qry[131 * 131 * 30:131 * 131 * 32]
    index.query(q, k)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    index.query(q, k)
    e3.record()
    torch.cuda.synchronize()
    print('scan_child {:4d}: full grid {:.1f} ms, mid slab {:.2f} ms, identical to scan_child 0: {}'.format(
        v, e0.elapsed_time(e1), e2.elapsed_time(e3), same))
