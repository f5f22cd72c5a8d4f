"""Breakdown of get_fkaconv_ids on the device for one 10k-point pass (debug aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppsurf_b200
from ppsurf_b200 import ops, synthetic
from ppsurf_b200.sampling import sampling_quantized

dev = torch.device('cuda:0')
pts = torch.from_numpy(synthetic.synthetic_cloud(10000, 3)).to(dev)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


gen = np.random.default_rng(1)
levels = [pts]
for i in range(4):
    n_sup = levels[-1].shape[0] // 4
    t, sel = timed(lambda: sampling_quantized(levels[-1], n_sup, gen))
    print('sampling level {} ({} -> {}): {:.3f} ms'.format(i, levels[-1].shape[0], n_sup, t))
    levels.append(levels[-1][sel].contiguous())
for a, c, k in ((0, 0, 16), (0, 1, 16), (1, 1, 16), (1, 2, 16), (2, 2, 16), (2, 3, 16), (3, 3, 16), (3, 4, 16), (4, 4, 16), (4, 3, 1), (3, 2, 1), (2, 1, 1), (1, 0, 1)):
    tb, index = timed(lambda: ops.KnnIndex(levels[a]))
    tq, _ = timed(lambda: index.query(levels[c], k))
    print('knn {}->{} k={}: build {:.3f} ms, query {:.3f} ms'.format(a, c, k, tb, tq))
