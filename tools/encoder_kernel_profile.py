"""Per-kernel device time (torch.profiler) of one encode_cloud call on the 100k-point bench cloud (graph replays included)."""
import collections, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import synthetic
dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(256, ['x'], 3, 2, 64, 0.0, False, 'bench', 'results', 0.05, 'p', 256, 10, 10000, 129, 50, 50000, 10, 8)
model.network.load_state_dict(synthetic.make_state_dict(model.network, 42), strict=True)
model = model.to(dev).eval()
pts = torch.from_numpy(synthetic.synthetic_cloud(100000, 42).T[None].copy()).to(dev)
model.network.sampling_seed = 42
for _ in range(3):
    model.encode_cloud(pts, generator=torch.Generator().manual_seed(1))
torch.cuda.synchronize(); t0 = time.perf_counter()
model.encode_cloud(pts, generator=torch.Generator().manual_seed(2)); torch.cuda.synchronize()
print('encode_cloud %.1f ms' % ((time.perf_counter() - t0) * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.encode_cloud(pts, generator=torch.Generator().manual_seed(3)); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name.split('(')[0][:60]][0] += 1; agg[e.name.split('(')[0][:60]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print('%-62s %6d %9.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
print('total device time %.1f us in %d kernels' % (tot, sum(v[0] for v in agg.values())))
