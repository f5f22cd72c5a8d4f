"""Per-phase cycle counters of CTA 0 of pn_stn_kernel on one chunk of patches (debug aid)."""
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
q = 18944
patches = torch.from_numpy(np.random.default_rng(1).standard_normal((q, 50, 3)).astype(np.float32) * 0.3).to(dev)
packed = net.packed()['decoder']
counters = torch.zeros(32, dtype=torch.int64, device=dev)
for it in range(3):
    _lib.lib.pps_debug_tc_profile(counters.data_ptr())
    ops.pointnet(packed, patches, 1)
    torch.cuda.synchronize()
c = counters.cpu().numpy()
tiles = q // 2 // 296
names = ['mma_total', 'mma_wait_ready', 'e_gather', 'e_wait_acc012', 'e_epi012', 'e_wait_stn3', 'e_max']
print('pn_stn per tile (CTA 0, {} tiles): '.format(tiles) + '  '.join('{}={:.0f}'.format(k, v / tiles) for k, v in zip(names, c[16:23])))
_lib.lib.pps_debug_tc_profile(None)
