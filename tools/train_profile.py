"""Per-kernel device time of one training step (BASELINE config 5) from torch.profiler (CUPTI): python tools/train_profile.py [fp32|bf16]"""
import os, sys, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, ppsurf_b200
from ppsurf_b200 import autograd as ag, data_pipeline, synthetic
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(256, ['x'], 3, 2, 64, 0.0, False, 'bench', 'results', 0.05, 'p', 256, 10, 10000, 129, 50, 50000, 10, 8)
model.network.load_state_dict(synthetic.make_state_dict(model.network, 42), strict=True)
model = model.to(dev).train()
net = model.network
net.sampling_seed = 7
host = {k: torch.from_numpy(v) for k, v in bench.fit_batch(2, 10000, 2000, 100).items()}
with torch.no_grad():
    batch = data_pipeline.prepare_batch(net, {k: v.to(dev) for k, v in host.items()})
opt = torch.optim.AdamW(net.parameters(), lr=1e-3, eps=1e-5, weight_decay=1e-2)
ag.set_precision(prec)
def step():
    opt.zero_grad(set_to_none=True)
    pred = net.forward(dict(batch))
    b, c, q = pred.shape
    loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(b * q, c), batch['occ'].reshape(-1))
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.split('(')[0][:60]
        agg[n][0] += 1; agg[n][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print('%-62s %5d %9.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
print('total device time %.1f us in %d kernels' % (tot, sum(v[0] for v in agg.values())))
big = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: -(e.device_time if hasattr(e, 'device_time') else e.cuda_time))[:14]
for e in big: print('  %-60s %8.1f us' % (e.name[:60], e.device_time if hasattr(e, 'device_time') else e.cuda_time))
