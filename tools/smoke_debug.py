"""Which branch of the tensor-core decode carries the largest deviation from the fp32 path on the smoke workload."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import ops, synthetic, _lib
dev = torch.device('cuda:0')
net = ppsurf_b200.PPSurfNetwork(3, 256, 2, 64, 50, 256)
net.load_state_dict(synthetic.make_state_dict(net, 42), strict=True)
net = net.to(dev).eval()
rng = np.random.default_rng(0)
pts = synthetic.synthetic_cloud(4200, seed=1)
net.sampling_seed = 5
data = net.get_latent({'pts': torch.from_numpy(pts.T[None].copy()).to(dev)})
lat = data['latents'][0].t().contiguous()
qry = torch.from_numpy((pts[rng.integers(0, 4200, 256)] + 0.02 * rng.standard_normal((256, 3))).astype(np.float32)).to(dev)
p = torch.from_numpy(pts).to(dev)
packed = net.packed()['decoder']
res = {}
for path in (0, 1):
    dec = ops.Decoder(packed, p, lat, chunk=256, path=path)
    idx, d2 = dec.index.query(qry, 64, return_dist=True)
    fp = dec.projection(qry, idx)
    patches = ops.patch_normalize(p, qry, idx, d2, 50)
    fl = ops.pointnet(packed, patches, path)
    out = dec.decode(qry, want_logits=True)['logits']
    res[path] = (fp, fl, out)
for name, i in (('projection', 0), ('pointnet', 1), ('logits', 2)):
    a, b = res[0][i], res[1][i]
    e = (a - b).abs()
    q = int(e.max(dim=1)[0].argmax())
    print('%-10s max |path1 - path0| %.3e at query %d (row scale %.3f, row max err %.3e), median %.3e, scale %.3f' % (
        name, float(e.max()), q, float(a[q].abs().max()), float(e[q].max()), float(e.median()), float(a.abs().max())))
for mask in (0x1ff, 0x0ff, 0x17f):
    _lib.lib.pps_decoder_tc_terms(mask)
    dec = ops.Decoder(packed, p, lat, chunk=256, path=1)
    idx = dec.index.query(qry, 64)
    fp = dec.projection(qry, idx)
    print('term mask %03x: projection max |path1 - path0| %.3e' % (mask, float((fp - res[0][0]).abs().max())))
_lib.lib.pps_decoder_tc_terms(0x1ff)
