"""Which GEMM shapes of the training step cost what: every pps_gemm call of one step timed with CUDA events (serialised)."""
import collections, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, ppsurf_b200
from ppsurf_b200 import autograd as ag, data_pipeline, synthetic, training
dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(256, ['x'], 3, 2, 64, 0.0, False, 'bench', 'results', 0.05, 'p', 256, 10, 10000, 129, 50, 50000, 10, 8)
model.network.load_state_dict(synthetic.make_state_dict(model.network, 42), strict=True)
model = model.to(dev).train()
net = model.network
net.sampling_seed = 7
training.CONCURRENT_BRANCHES = False
host = {k: torch.from_numpy(v) for k, v in bench.fit_batch(2, 10000, 2000, 100).items()}
with torch.no_grad():
    batch = data_pipeline.prepare_batch(net, {k: v.to(dev) for k, v in host.items()})
ag.set_precision('bf16')
def step():
    for p in net.parameters(): p.grad = None
    pred = net.forward(dict(batch))
    b, c, q = pred.shape
    loss, _ = ag.cross_entropy(pred.transpose(1, 2).reshape(b * q, c), batch['occ'].reshape(-1))
    loss.backward()
for _ in range(2): step()
orig = ag.gemm
log = []
def timed(a, b, **kw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(a, b, **kw); e1.record(); torch.cuda.synchronize()
    kind = 'A:%s B:%s' % ('k' if a.stride(-1) == 1 else 'm', 'k' if b.stride(-2) == 1 else 'n')
    log.append((tuple(a.shape), b.shape[-1], kind, e0.elapsed_time(e1) * 1e3))
    return out
ag.gemm = timed
step()
ag.gemm = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for shp, n, kind, us in log:
    agg[(shp, n, kind)][0] += 1; agg[(shp, n, kind)][1] += us
tot = sum(v[1] for v in agg.values())
print('%d gemm calls, %.1f us in total (serialised, includes ~8 us of launch + event overhead each)' % (len(log), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print('%-22s n=%-5d %-8s x%-3d %9.1f us (%6.1f each)' % (str(k[0]), k[1], k[2], v[0], v[1], v[1] / v[0]))
