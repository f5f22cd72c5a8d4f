"""Per-phase cycle counters of CTA 0 of projection_tc_kernel on one chunk of the bench workload (debug aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import _lib, ops, synthetic

dev = torch.device('cuda:0')
net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
net.load_state_dict(synthetic.make_state_dict(net, 42))
net = net.to(dev)
n = 100000
pts = torch.from_numpy(synthetic.synthetic_cloud(n, 42)).to(dev)
lat = torch.from_numpy(np.random.default_rng(7).standard_normal((n, 256)).astype(np.float32)).to(dev)
dec = ops.Decoder(net.packed()['decoder'], pts, lat, chunk=18944, path=1)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts.cpu().numpy(), 129, 1)
qry = ops.grid_queries(131, step, bmin_pad, first=131 * 131 * 60, count=18944, device=dev)
idx = dec.index.query(qry, 64)
counters = torch.zeros(128, dtype=torch.int64, device=dev)
tiles = 18944 // 4 // 74
names = ['mma_total', 'mma_wait_chunk', 'mma_wait_weights', 'mma_wait_sgroup', 'g_gather', 'g_wait', 'g_E2E3', 's_wait', 's_softmax', 's_pool', 'mma_issue32', 'mma_commit32']
ref = None
print('CTA pairs resident at once:', _lib.lib.pps_debug_tc_max_clusters())
for cs, mask in ((0, 0x1FF), (0, 0x049), (0, 0x0DB)):  # all three terms, hh only, hh + hl
    _lib.lib.pps_debug_tc_cluster(cs)
    _lib.lib.pps_decoder_tc_terms(mask)
    for it in range(3):
        _lib.lib.pps_debug_tc_profile(counters.data_ptr())
        out = dec.projection(qry, idx)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5):
        dec.projection(qry, idx)
    e1.record()
    torch.cuda.synchronize()
    c = counters.cpu().numpy()
    ref = out if ref is None else ref
    print('term mask {:#x}: {:.1f} us per chunk (incl. value matrix), max |diff| vs all terms {:.2e}'.format(
        mask, e0.elapsed_time(e1) * 200, float((out - ref).abs().max())))
    print('   ' + '  '.join('{}={:.0f}'.format(k, v / tiles) for k, v in zip(names, c)))
    per_pair = c[32:32 + 74] / tiles
    print('   per-pair cycles per tile: min {:.0f} median {:.0f} max {:.0f}; slowest pairs {}'.format(per_pair.min(), np.median(per_pair), per_pair.max(), np.argsort(per_pair)[-5:]))
_lib.lib.pps_debug_tc_profile(None)
_lib.lib.pps_debug_tc_cluster(0)
_lib.lib.pps_decoder_tc_terms(0x0FF)
