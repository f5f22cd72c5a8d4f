"""Per-layer term ablation of the split-fp16 GEMMs of the global branch (projection_tc_kernel): which of the three products
x_hi w_hi + x_lo w_hi + x_hi w_lo each layer (fc2, fc3, fc_query) needs to stay inside the 1e-4 logit tolerance, and what a dropped
term buys.  Error = max |logit - float64 oracle| over 4096 queries (same neighbours), real encoder latents of the bench cloud.
    python tools/pass_ablation.py  > profiles/r02_pass_ablation.md"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppsurf_b200
from ppsurf_b200 import _lib, ops, synthetic
from oracle import ppsurf_oracle as oracle  # checker

dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(
    pointnet_latent_size=256, output_names=['imp_surf_sign'], in_channels=3, out_channels=2, k=64, lambda_l1=0.0, debug=False,
    in_file='bench', results_dir='results', padding_factor=0.05, name='ppsurf_50nn', network_latent_size=256,
    gen_subsample_manifold_iter=10, gen_subsample_manifold=10000, gen_resolution_global=129, num_pts_local=50,
    rec_batch_size=50000, gen_refine_iter=10, workers=8)
net = model.network
sd = synthetic.make_state_dict(net, 42)
net.load_state_dict(sd)
model = model.to(dev)
weights = {k: v.numpy() for k, v in sd.items()}
n = 100000
pts_np = synthetic.synthetic_cloud(n, 42)
pts_bcn = torch.from_numpy(pts_np.T[None].copy()).to(dev)
net.sampling_seed = 42
latents = model.encode_cloud(pts_bcn, generator=torch.Generator().manual_seed(42)).contiguous()
dec = net.decoder_for(pts_bcn, latents)
step, bmin_pad, _ = model.grid_definition(pts_np, 129, 1)
rng = np.random.default_rng(3)
grid = oracle.dense_grid_queries(pts_np, 129, 1)
near = grid[np.abs(np.linalg.norm(grid, axis=1) - 0.4) < 0.03]
qs = np.concatenate([near[rng.choice(near.shape[0], 3072, replace=False)], grid[rng.choice(grid.shape[0], 1024, replace=False)]]).astype(np.float32)
q_dev = torch.from_numpy(qs).to(dev)
idx = dec.index.query(q_dev, 64).cpu().numpy().astype(np.int64)
lat_np = latents.cpu().numpy()
ref = []
for s0 in range(0, qs.shape[0], 512):
    sl = slice(s0, s0 + 512)
    data = {'pts': pts_np.T[None], 'latents': lat_np, 'pts_query': qs[sl][None], 'pts_local_ps': oracle.get_pts_local_ps(pts_np, qs[sl], 50)[None],
            'proj_ids': idx[sl][None]}
    ref.append(oracle.from_latent(weights, data, dtype=np.float64)[0].T)
ref = np.concatenate(ref)  # [Q,2]
big = ops.grid_queries(131, step, bmin_pad, first=131 * 131 * 40, count=37888 * 8, device=dev)

names = {0: 'hh', 1: 'hh+lh', 2: 'hh+hl', 3: 'hh+lh+hl'}


def mask_of(fc2, fc3, fcq):  # each: bit0 = x_lo w_hi, bit1 = x_hi w_lo
    return 0x49 | (fc2 << 1) | (fc3 << 4) | (fcq << 7)


print('| fc2 | fc3 | fc_query | max abs logit err | max abs occupancy err | projection ms / 303k queries |')
print('|---|---|---|---|---|---|')
for fc2, fc3, fcq in ((3, 3, 3), (3, 3, 2), (3, 3, 1), (3, 3, 0), (3, 2, 3), (3, 1, 3), (3, 2, 2), (2, 3, 3), (1, 3, 3), (2, 2, 2), (3, 2, 0), (2, 2, 0), (0, 0, 0)):
    _lib.lib.pps_decoder_tc_terms(mask_of(fc2, fc3, fcq))
    res = dec.decode(q_dev, want_logits=True, want_occ=True)
    err = float(np.abs(res['logits'].cpu().numpy() - ref).max())
    occ_ref = oracle.occupancy_from_logits(ref.T[None])[0]
    err_o = float(np.abs(res['occ'].cpu().numpy() - occ_ref).max())
    for _ in range(2):
        dec.decode(big, want_logits=False, want_occ=True)
    _lib.lib.pps_profile_enable(1)
    dec.decode(big, want_logits=False, want_occ=True)
    import ctypes
    ms, br = ctypes.c_double(0), ctypes.c_longlong(0)
    _lib.lib.pps_profile_read(ctypes.byref(ms), ctypes.byref(br))
    _lib.lib.pps_profile_enable(0)
    print('| {} | {} | {} | {:.2e} | {:.2e} | {:.2f} |'.format(names[fc2], names[fc3], names[fcq], err, err_o, ms.value), flush=True)
_lib.lib.pps_decoder_tc_terms(0x1FF)
