"""Where the region-grown shell decode of predict_step spends its time: per-sweep query counts / times and the per-kernel device time
(torch.profiler) of one create_volume_device call on the bench cloud.  python tools/predict_profile.py"""
import collections, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppsurf_b200
from ppsurf_b200 import ops, synthetic
dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(256, ['x'], 3, 2, 64, 0.0, False, 'bench', 'results', 0.05, 'p', 256, 10, 10000, 129, 50, 50000, 10, 8)
model.network.load_state_dict(synthetic.make_state_dict(model.network, 42), strict=True)
model = model.to(dev).eval()
net = model.network
pts_np = synthetic.synthetic_cloud(100000, 42)
pts = torch.from_numpy(pts_np.T[None].copy()).to(dev)
net.sampling_seed = 42
lat = model.encode_cloud(pts, generator=torch.Generator().manual_seed(1)).contiguous()
dec = net.decoder_for(pts, lat)
log = []
orig = model.occupancy
def timed_occ(decoder, q):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = orig(decoder, q)
    torch.cuda.synchronize(); log.append((q.shape[0], (time.perf_counter() - t0) * 1e3))
    return out
for rep in range(2):
    log.clear()
    model.occupancy = timed_occ
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vol = model.create_volume_device(dec, pts_np, 129)
    torch.cuda.synchronize(); total = (time.perf_counter() - t0) * 1e3
print('sweeps', len(log), 'queries', sum(n for n, _ in log), 'decode ms', sum(t for _, t in log), 'total ms (with syncs)', total)
for n, t in log: print('  %8d queries %8.2f ms  %6.2f Mq/s' % (n, t, n / t / 1e3))
model.occupancy = orig
torch.cuda.synchronize(); t0 = time.perf_counter()
vol = model.create_volume_device(dec, pts_np, 129)
torch.cuda.synchronize(); print('untimed-sweeps total ms', (time.perf_counter() - t0) * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    vol = model.create_volume_device(dec, pts_np, 129); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name.split('(')[0][:60]][0] += 1; agg[e.name.split('(')[0][:60]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print('%-62s %5d %9.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
print('total device time %.1f us' % tot)
