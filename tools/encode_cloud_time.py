"""Wall-clock of PPSurfModel.encode_cloud on the 100k-point bench cloud (no instrumentation)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import synthetic

dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'bench', 'results', 0.05, 'p', 256, 10, 10000, 129, 50,
                                50000, 10, 8)
net = model.network
net.load_state_dict(synthetic.make_state_dict(net, 42))
model = model.to(dev)
pts = torch.from_numpy(synthetic.synthetic_cloud(100000, 42).T[None].copy()).to(dev)
for rep in range(3):
    net.sampling_seed = 42
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lat = model.encode_cloud(pts, generator=torch.Generator().manual_seed(42))
    torch.cuda.synchronize()
    print('encode_cloud: {:.3f} s, checksum {:.6f}'.format(time.perf_counter() - t0, float(lat.double().abs().mean())))
