"""Roofline position of the training GEMM (pps_gemm, bf16 tcgen05 path) on the layer shapes of the step: CUDA-event time,
algorithmic HBM bytes (fp32 operands in + fp32 result out, weights counted once) and FLOPs.  python tools/train_gemm_bench.py [--once]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppsurf_b200 import autograd as ag
dev = torch.device('cuda:0')
once = '--once' in sys.argv
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
hbm = peaks.get('hbm_gbs', 6650.0)
shapes = [('projection fc2/fc3 forward', 256000, 256, 256, 'fwd'), ('projection fc2/fc3 data gradient', 256000, 256, 256, 'dgrad'),
          ('projection fc2/fc3 weight gradient', 256000, 256, 256, 'wgrad'), ('fc_query forward', 256000, 64, 256, 'fwd'),
          ('stn.conv3 forward', 200000, 256, 128, 'fwd'), ('stn.conv3 weight gradient', 200000, 256, 128, 'wgrad'),
          ('pointnet conv1 forward', 200000, 64, 64, 'fwd'), ('encoder cv0 1x1 forward', 20000, 32, 64, 'fwd')]
rows = []
for name, m, n, k, kind in shapes:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); dy = torch.randn(m, n, device=dev)
    if kind == 'fwd':
        fn = lambda: ag.gemm(x, w.t(), prec=1); byts = (m * k + n * k + m * n) * 4
    elif kind == 'dgrad':
        fn = lambda: ag.gemm(dy, w, prec=1); byts = (m * n + n * k + m * k) * 4
    else:
        fn = lambda: ag.gemm(dy.t(), x, prec=1); byts = (m * n + m * k + n * k) * 4
    reps = 1 if once else 20
    for _ in range(0 if once else 3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    rows.append((name, m, n, k, us, byts / us / 1e3, byts / us / 1e3 / hbm, 2.0 * m * n * k / us / 1e6))
    print('%-36s m=%-7d n=%-4d k=%-4d %8.1f us %7.0f GB/s (%.2f of %d)  %6.1f TFLOP/s' % rows[-1][:4] + '' if False else
          '%-36s m=%-7d n=%-4d k=%-4d %8.1f us %7.0f GB/s (%.2f of HBM peak) %6.1f TFLOP/s' % rows[-1])
