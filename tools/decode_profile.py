"""Per-kernel device time (torch.profiler) of one dense decode: python tools/decode_profile.py [points] [num_pts_local] [resolution]"""
import collections, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import ops, synthetic
n, npl, res = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 100000), (2, 50), (3, 129)))
dev = torch.device('cuda:0')
net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, npl, 256)
net.load_state_dict(synthetic.make_state_dict(net, 42), strict=True)
net = net.to(dev).eval()
pts_np = synthetic.synthetic_cloud(n, 42)
pts = torch.from_numpy(pts_np).to(dev)
lat = torch.from_numpy(np.random.default_rng(7).standard_normal((n, 256)).astype(np.float32)).to(dev)
dec = ops.Decoder(net.packed()['decoder'], pts, lat, chunk=37888, path=1)
step, bmin_pad, _ = ppsurf_b200.PPSurfModel.grid_definition(pts_np, res, 1)
q = ops.grid_queries(res + 2, step, bmin_pad, device=dev)
dec.decode(q, want_logits=False, want_occ=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dec.decode(q, want_logits=False, want_occ=True); e1.record(); torch.cuda.synchronize()
print('%d points, P=%d, %d queries: %.1f ms, %.2f Mquery/s' % (n, npl, q.shape[0], e0.elapsed_time(e1), q.shape[0] / e0.elapsed_time(e1) / 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    dec.decode(q, want_logits=False, want_occ=True); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name.split('(')[0][:60]][0] += 1; agg[e.name.split('(')[0][:60]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print('%-62s %5d %9.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
